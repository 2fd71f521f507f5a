set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_bc7amd.py tests/test_golden.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python tools/amd_mode_times.py 2048 2>&1 | tail -20
B200IC_EXTRA_DEFS=B200IC_AMD_TIMING B200IC_BUILD_DIR=/tmp/tl python -m gfx_imagecompress_b200.build > /tmp/tl.log 2>&1; tail -2 /tmp/tl.log
B200IC_LIB=/tmp/tl/libgfx_imagecompress_b200.so timeout 300 python tools/amd_phase_times.py 1024
