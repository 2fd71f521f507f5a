set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_bc1.py tests/test_bc23.py tests/test_golden.py -q -m gpu -k "bc1 or bc2 or bc3 or golden" 2>&1 | tail -3
python tools/rg_time.py bc1 4096
python tools/rg_time.py bc1 1024
