set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_bc7amd.py tests/test_golden.py tests/test_parity_wide.py -q -m gpu -k "not headline" 2>&1 | tail -4
timeout 300 python tools/amd_mode_times.py 2048 2>&1 | tail -20
