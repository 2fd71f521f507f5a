#!/usr/bin/env python3
"""Generates tests/golden/*.npz: small seeded inputs together with the blocks the UNMODIFIED reference
(oracle/_ref/libref_oracle.so, compiled from /root/reference/src by oracle/Makefile) produces for them.

The reference's own tests hold no golden vectors, so these fixtures -- outputs of the reference itself run
in the build container -- are what pins parity where the reference library cannot travel.  Re-run after changing
the compat shim or the synthetic generators:

    python tests/golden/make_golden.py

Every fixture stores: `pixels` (H, W, C) in the source format, `fmt` (b200ic_format), `codec` (b200ic_codec), `blocks`
(nblocks, 8|16) uint8 and, for option variants, `opts` as a json string.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from oracle.ref import BC1, BC4, BC5, BC6H, BC7, BC7_RG, default_opts  # noqa: E402
from gfx_imagecompress_b200 import synth  # noqa: E402


def fixtures():
    rgba = synth.rgba8_gradnoise(64, 32, 3, "lefthalf")
    punch = synth.rgba8_gradnoise(32, 32, 1, "punch")
    npot = synth.rgba8_gradnoise(37, 21, 5, "ramp")
    pat_rgb, f_rgb = synth.pattern("RGB", 33, 33)
    pat_pt, f_pt = synth.pattern("RGB_Punchthrough", 32, 32)
    pat_rgba, f_rgba = synth.pattern("RGBA", 32, 32)
    rg = synth.height_rg8(64, 64, 2)
    hdr = synth.hdr_rgba16f(32, 32, 4)
    hdr_npot = synth.hdr_rgba16f(18, 10, 9)
    F8 = synth.FMT_RGBA8
    yield "bc1_gradnoise_punch", BC1, punch, F8, {}
    yield "bc1_npot", BC1, npot, F8, {}
    yield "bc1_pattern_punchthrough_usealpha", BC1, pat_pt, f_pt, dict(bc1_use_alpha=1, bc1_alpha_threshold=128)
    yield "bc1_steps2_nothreshold", BC1, punch, F8, dict(bc1_use_alpha=0, bc1_alpha_threshold=0, amd_refinement_steps=2)
    yield "bc4_height_rg8", BC4, rg, synth.FMT_RG8, {}
    yield "bc5_height_rg8", BC5, rg, synth.FMT_RG8, {}
    yield "bc5_npot_rgba8", BC5, npot, F8, {}
    yield "bc7rg_lefthalf", BC7_RG, rgba, F8, {}
    yield "bc7rg_pattern_rgb_npot", BC7_RG, pat_rgb, f_rgb, {}
    yield "bc7rg_linear_fast", BC7_RG, rgba, F8, dict(rg_perceptual=0, rg_fast=1)
    yield "bc7amd_lefthalf", BC7, rgba, F8, {}
    yield "bc7amd_pattern_rgba", BC7, pat_rgba, f_rgba, {}
    yield "bc7amd_npot", BC7, npot, F8, {}
    yield "bc7amd_mask_0x42", BC7, rgba, F8, dict(amd_mode_mask=0x42)
    yield "bc6h_hdr", BC6H, hdr, synth.FMT_RGBA16UF, {}
    yield "bc6h_hdr_npot", BC6H, hdr_npot, synth.FMT_RGBA16UF, {}


def main():
    oracle.build()
    if not oracle.have_ref():
        raise SystemExit("needs oracle/_ref/libref_oracle.so (build container with /root/reference)")
    ref = oracle.RefOracle()
    total = 0
    for name, codec, px, fmt, opts in fixtures():
        px = np.ascontiguousarray(px)
        blocks = ref.encode(codec, px, fmt, opts=default_opts(**opts) if opts else None)
        stored = px.view(np.uint16) if px.dtype == np.float16 else px  # npz without pickles: halves as bit patterns
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, pixels=stored, is_half=np.array(px.dtype == np.float16), fmt=np.array(fmt),
                            codec=np.array(codec), blocks=blocks, opts=np.array(json.dumps(opts)))
        total += os.path.getsize(path)
        print(f"{name}: {px.shape} -> {blocks.shape}")
    print("total bytes", total)


if __name__ == "__main__":
    main()
