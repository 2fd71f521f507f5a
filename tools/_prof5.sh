cd $GRAFT_REPO_ROOT
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
ncu --set full --import-source on --clock-control none -k regex:amd_cube_kernel -s 1 -c 1 -o gpurun_out/r2g_cube_mode0 python /tmp/opq.py 1024 opaque 0x01 > gpurun_out/ncu_g0.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:amd_quant_kernel -s 1 -c 1 -o gpurun_out/r2g_quant_mode1 python /tmp/opq.py 1024 opaque 0x02 > gpurun_out/ncu_g1.log 2>&1
