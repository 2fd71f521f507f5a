#!/usr/bin/env python3
"""Device-only timing of one codec on the benchmark pattern (quick iteration aid): rg_time.py codec size [kind]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gfx_imagecompress_b200 as g
from gfx_imagecompress_b200 import synth

g.load_library(); g.init(0)
codec = {"bc7_rg": g.BC7_RG, "bc1": g.BC1, "bc7_amd": g.BC7_AMD, "bc6h": g.BC6H}[sys.argv[1]]
n = int(sys.argv[2]); kind = sys.argv[3] if len(sys.argv) > 3 else "lefthalf"
dev = torch.device("cuda", 0)
hdr = codec == g.BC6H
px = torch.from_numpy(synth.hdr_rgba16f(n, n, 5).view("uint16") if hdr else synth.rgba8_gradnoise(n, n, 3, kind)).to(dev)
fmt = synth.FMT_RGBA16UF if hdr else synth.FMT_RGBA8
bb = 8 if codec == g.BC1 else 16
out = torch.empty((n * n // 16, bb), dtype=torch.uint8, device=dev)
for _ in range(3):
    g.encode_device(codec, px, fmt, n, n, 1, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    g.encode_device(codec, px, fmt, n, n, 1, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"{sys.argv[1]} {n}^2 {kind}: {ms:.3f} ms  {n * n / ms / 1e3:.1f} Mpix/s")
