cd $GRAFT_REPO_ROOT
timeout 300 python tools/amd_mode_times.py 2048 2>&1 | grep "opaque  mask 0x01\|opaque  mask 0x10\|mask 0xff"
