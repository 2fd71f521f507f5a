set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_bc7amd.py tests/test_golden.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python tools/amd_mode_times.py 2048 2>&1 | tail -20
B200IC_EXTRA_DEFS=B200IC_AMD_TIMING B200IC_BUILD_DIR=/tmp/tl python -m gfx_imagecompress_b200.build > /tmp/tl.log 2>&1; tail -2 /tmp/tl.log
B200IC_LIB=/tmp/tl/libgfx_imagecompress_b200.so timeout 300 python tools/amd_phase_times.py 1024
cat > /tmp/opq.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
import gfx_imagecompress_b200 as g
from gfx_imagecompress_b200 import synth
g.load_library(); g.init(0)
n = int(sys.argv[1]); kind = sys.argv[2]; mask = int(sys.argv[3], 0)
dev = torch.device("cuda", 0)
px = torch.from_numpy(synth.rgba8_gradnoise(n, n, 3, kind)).to(dev)
out = torch.empty((n * n // 16, 16), dtype=torch.uint8, device=dev)
o = g.Opts.default(amd_mode_mask=mask)
for _ in range(2):
    g.encode_device(g.BC7_AMD, px, synth.FMT_RGBA8, n, n, 1, opts=o, out=out)
torch.cuda.synchronize()
PY
ncu --set full --import-source on --clock-control none -k regex:amd_cube_kernel -s 1 -c 1 -o gpurun_out/r2_cube_mode0 python /tmp/opq.py 1024 opaque 0x01 > gpurun_out/ncu_c.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:amd_window_kernel -s 1 -c 1 -o gpurun_out/r2_window_mode1 python /tmp/opq.py 1024 opaque 0x02 > gpurun_out/ncu_f.log 2>&1
