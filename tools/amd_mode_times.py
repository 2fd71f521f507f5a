#!/usr/bin/env python3
"""Times the AMD BC7 kernel per mode (ModeMask) on opaque / translucent inputs: shows where the time goes and whether
the single-mode times add up to the all-modes time (instruction-cache / local-memory interference if they do not)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gfx_imagecompress_b200 as g
from gfx_imagecompress_b200 import synth

g.load_library(); g.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda", 0)
for kind in ("opaque", "ramp"):
    px = torch.from_numpy(synth.rgba8_gradnoise(n, n, 3, kind)).to(dev)
    out = torch.empty((n * n // 16, 16), dtype=torch.uint8, device=dev)
    tot = 0.0
    for mask in [1 << m for m in range(8)] + [0xFF]:
        o = g.Opts.default(amd_mode_mask=mask)
        g.encode_device(g.BC7_AMD, px, synth.FMT_RGBA8, n, n, 1, opts=o, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            g.encode_device(g.BC7_AMD, px, synth.FMT_RGBA8, n, n, 1, opts=o, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        if mask != 0xFF:
            tot += ms
        import subprocess
        clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
        print(f"[{clk}] {kind:7s} mask {mask:#04x}: {ms:8.2f} ms  {n * n / 16 / ms / 1e3:8.2f} Mblocks/s" + (f"   (sum of single modes {tot:.2f} ms)" if mask == 0xFF else ""), flush=True)
