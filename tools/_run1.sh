set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
timeout 900 python -m pytest tests/test_bc7amd.py tests/test_golden.py -x -q -m gpu 2>&1 | tail -15
timeout 300 python tools/amd_mode_times.py 2048 2>&1 | tail -20
