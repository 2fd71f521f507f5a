set -x
cd $GRAFT_REPO_ROOT
cat > /tmp/opq.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
import gfx_imagecompress_b200 as g
from gfx_imagecompress_b200 import synth
g.load_library(); g.init(0)
n = int(sys.argv[1]); kind = sys.argv[2]; mask = int(sys.argv[3], 0)
codec = g.BC7_RG if len(sys.argv) > 4 else g.BC7_AMD
dev = torch.device("cuda", 0)
px = torch.from_numpy(synth.rgba8_gradnoise(n, n, 3, kind)).to(dev)
out = torch.empty((n * n // 16, 16), dtype=torch.uint8, device=dev)
o = g.Opts.default(amd_mode_mask=mask)
for _ in range(2):
    g.encode_device(codec, px, synth.FMT_RGBA8, n, n, 1, opts=o, out=out)
torch.cuda.synchronize()
PY
i=0
for spec in "amd_cube_kernel 0x10 r2f_cube_mode4" "amd_window_kernel 0x01 r2f_window_mode0" "amd_cube_kernel 0x01 r2f_cube_mode0" "amd_window_kernel 0x10 r2f_window_mode4"; do
  set -- $spec
  ncu --set full --import-source on --clock-control none -k regex:$1 -s 1 -c 1 -o gpurun_out/$3 python /tmp/opq.py 1024 opaque $2 > gpurun_out/ncu_f$i.log 2>&1
  i=$((i+1))
done
ncu --set full --import-source on --clock-control none -k regex:bc7rg_kernel -s 1 -c 1 -o gpurun_out/r2f_bc7rg python /tmp/opq.py 2048 lefthalf 0xff rg > gpurun_out/ncu_f9.log 2>&1
