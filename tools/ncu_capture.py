#!/usr/bin/env python3
"""Workload for `ncu --set full` captures: two encodes of one synthetic image through the host C-ABI
(the first is the warm-up that ncu skips with -s).  usage: ncu_capture.py codec size"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gfx_imagecompress_b200 as g
from gfx_imagecompress_b200 import synth

g.load_library()
g.init(0)
codec = {"bc1": 1, "bc4": 4, "bc5": 5, "bc6h": 6, "bc7_amd": 7, "bc7_rg": 8}[sys.argv[1]]
n = int(sys.argv[2])
if codec == 6:
    px, fmt = synth.hdr_rgba16f(n, n, 4), synth.FMT_RGBA16UF
elif codec in (4, 5):
    px, fmt = synth.height_rg8(n, n, 2), synth.FMT_RG8
else:
    px, fmt = synth.rgba8_gradnoise(n, n, 3, "punch" if codec == 1 else "lefthalf"), synth.FMT_RGBA8
for _ in range(2):
    g.encode_host(codec, px, fmt)
