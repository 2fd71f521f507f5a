set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_bc7rg.py tests/test_golden.py -q -m gpu -k "rg or golden" 2>&1 | tail -3
python tools/rg_time.py bc7_rg 8192
python tools/rg_time.py bc7_rg 4096 opaque
python tools/rg_time.py bc7_rg 4096 ramp
