set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_bc6h.py tests/test_golden.py -q -m gpu -k "bc6h or hdr or golden" 2>&1 | tail -3
python tools/rg_time.py bc6h 4096
python tools/rg_time.py bc1 4096
python tools/rg_time.py bc1 1024
