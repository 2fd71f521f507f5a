#!/usr/bin/env python3
"""Phase breakdown of the AMD BC7 warp kernel per mode (needs a build with B200IC_EXTRA_DEFS=B200IC_AMD_TIMING):
clock64 deltas of lane 0 summed over all warps, plus item / round counts of the shake phases."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gfx_imagecompress_b200 as g
from gfx_imagecompress_b200 import synth

L = g.load_library(); g.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda", 0)
names = ["quantise", "rank", "cube", "window", "window2", "pick+pack", "cube items", "cube batches", "win items", "win rounds", "cube passes"]
for kind, modes in (("opaque", (0, 1, 2, 3, 4, 5)), ("ramp", (4, 5, 7))):
    px = torch.from_numpy(synth.rgba8_gradnoise(n, n, 3, kind)).to(dev)
    out = torch.empty((n * n // 16, 16), dtype=torch.uint8, device=dev)
    nb = n * n // 16
    for m in modes:
        o = g.Opts.default(amd_mode_mask=1 << m)
        buf = (C.c_ulonglong * 16)()
        L.b200ic_amd_timing(buf, 1)
        g.encode_device(g.BC7_AMD, px, synth.FMT_RGBA8, n, n, 1, opts=o, out=out)
        L.b200ic_amd_timing(buf, 0)
        v = list(buf)
        tot = sum(v[:6]) or 1
        print(f"{kind} mode {m}: " + ", ".join(f"{names[i]} {100 * v[i] / tot:.0f}%" for i in range(6)) +
              " | per block: " + ", ".join(f"{names[i]} {v[i] / nb:.1f}" for i in range(6, 11)), flush=True)
