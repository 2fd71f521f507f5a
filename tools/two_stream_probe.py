#!/usr/bin/env python3
"""Does running two independent AMD BC7 encodes on two streams at once beat running them back to back?  (Probe for
co-scheduling the FP64-latency-bound quantise kernels with the ALU-bound cube kernels.)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gfx_imagecompress_b200 as g
from gfx_imagecompress_b200 import synth

g.load_library(); g.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
dev = torch.device("cuda", 0)
imgs = [torch.from_numpy(synth.rgba8_gradnoise(n, n, 3 + i, "lefthalf")).to(dev) for i in range(2)]
outs = [torch.empty((n * n // 16, 16), dtype=torch.uint8, device=dev) for _ in range(2)]
streams = [torch.cuda.Stream(dev) for _ in range(2)]


def run(concurrent):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for rep in range(2):
        for i in range(2):
            st = streams[i] if concurrent else streams[0]
            st.wait_event(e0)
            g.encode_device(g.BC7_AMD, imgs[i], synth.FMT_RGBA8, n, n, 1, out=outs[i], stream=st.cuda_stream)
    for st in streams:
        torch.cuda.current_stream(dev).wait_stream(st)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 4


for _ in range(2):
    a, b = run(False), run(True)
    print(f"{n}^2 lefthalf: back to back {a:.2f} ms per encode, two streams {b:.2f} ms per encode ({a / b:.3f}x)", flush=True)
