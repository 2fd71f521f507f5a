set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 300 python tools/amd_mode_times.py 2048 2>&1 | tail -20
