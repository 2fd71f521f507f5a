#!/usr/bin/env python3
"""In-process multi-GPU sharding of ONE image behind the reference's entry point: Image_CompressAMDBC7 (pageable image in,
malloc'd image out) with b200ic_set_devices(1, 2, ... visible).  One JSON line per device count.
usage: inprocess_scaling.py [size=8192] [steps=3]"""
import ctypes as C
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gfx_imagecompress_b200 as g
from gfx_imagecompress_b200 import synth

size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
L = g.load_library()
g.init(0)
px = synth.rgba8_gradnoise(size, size, 3, "lefthalf")
img = g.Image(px, synth.FMT_RGBA8)
libc = C.CDLL(None)
libc.free.argtypes = [C.c_void_p]
base = None
n = 1
while n <= g.device_count():
    g.set_devices(n)
    addr = L.Image_CompressAMDBC7(img.ptr, None, None, None)  # warm-up: contexts, tables, staging buffers of every device
    hdr = g.api._ImageHeader.from_address(addr)
    blocks = np.ctypeslib.as_array((C.c_uint8 * hdr.dataSize).from_address(addr + 48)).copy()
    libc.free(C.c_void_p(addr))
    if base is None:
        base = blocks
    t0 = time.perf_counter()
    for _ in range(steps):
        libc.free(C.c_void_p(L.Image_CompressAMDBC7(img.ptr, None, None, None)))
    dt = (time.perf_counter() - t0) / steps
    print(json.dumps({"path": "Image_CompressAMDBC7 (pageable Image in, malloc'd Image out)", "image": [size, size], "devices": n,
                      "ms_per_image": dt * 1e3, "mpix_per_s": size * size / 1e6 / dt, "identical_to_one_device": bool(np.array_equal(blocks, base))}), flush=True)
    n *= 2
