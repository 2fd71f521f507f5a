"""GPU: image layouts around the block search -- slices, padded rows (row pitch), narrow source formats, sRGB tags --
the gather of reference src/block_utils.cpp:7-41,116-144 (Image_CalculateIndex is slice-major, row-major)."""
import numpy as np
import pytest

import cases
from gfx_imagecompress_b200 import synth
from oracle.ref import BC1, BC5, BC7, BC7_RG

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("codec,refc", [(5, BC5), (8, BC7_RG), (1, BC1)])
def test_slices_are_independent_images(engine, ref, codec, refc):
    a = synth.rgba8_gradnoise(37, 21, 5, "ramp")
    b = synth.rgba8_gradnoise(37, 21, 6, "lefthalf")
    both = np.ascontiguousarray(np.stack([a, b]))
    got = engine.encode_host(codec, both, synth.FMT_RGBA8)
    want = np.concatenate([ref.encode(refc, a, synth.FMT_RGBA8), ref.encode(refc, b, synth.FMT_RGBA8)])
    assert np.array_equal(got, want)


def test_row_and_slice_pitch(engine, ref):
    import torch
    dev = torch.device("cuda", 0)
    img = synth.rgba8_gradnoise(50, 18, 7, "lefthalf")           # (18, 50, 4)
    padded = np.zeros((2, 24, 64, 4), np.uint8)                      # rows padded to 64 texels, slices to 24 rows
    padded[0, :18, :50] = img
    padded[1, :18, :50] = img[::-1]
    t = torch.from_numpy(padded).to(dev)
    out = engine.encode_device(engine.BC7_RG, t, synth.FMT_RGBA8, 50, 18, slices=2, row_pitch=64 * 4, slice_pitch=24 * 64 * 4)
    torch.cuda.synchronize()
    want = np.concatenate([ref.encode(BC7_RG, img, synth.FMT_RGBA8), ref.encode(BC7_RG, np.ascontiguousarray(img[::-1]), synth.FMT_RGBA8)])
    assert np.array_equal(out.cpu().numpy(), want)


def test_narrow_sources(engine, ref):
    """R8 / RG8 / RGB8 sources: missing channels are g = b = 0, a = 1 (compat shim), for every colour codec."""
    rgba = synth.rgba8_gradnoise(32, 16, 9, "opaque")
    for nch, fmt in ((1, synth.FMT_R8), (2, synth.FMT_RG8), (3, synth.FMT_RGB8)):
        px = np.ascontiguousarray(rgba[..., :nch])
        for codec, refc in ((1, BC1), (8, BC7_RG), (7, BC7)):
            got = engine.encode_host(codec, px, fmt)
            want = ref.encode(refc, px, fmt)
            assert np.array_equal(got, want), (nch, codec)


def test_srgb_tags(engine):
    px = synth.rgba8_gradnoise(16, 16, 2, "opaque")
    lin = engine.Image_CompressAMDBC7(engine.Image(px, synth.FMT_RGBA8))
    srgb = engine.Image_CompressAMDBC7(engine.Image(px, synth.FMT_RGBA8_SRGB))
    assert lin.format == 26 and srgb.format == 27          # DXBC7_UNORM / DXBC7_SRGB: the reference only retags
    assert np.array_equal(lin.blocks(16), srgb.blocks(16))


def test_depth_greater_than_one_is_rejected(engine):
    import ctypes as C
    px = synth.rgba8_gradnoise(8, 8, 2, "opaque")
    img = engine.Image(px, synth.FMT_RGBA8)
    img.header.depth = 2
    assert engine.Image_CompressAMDBC7(img) is None and engine.Image_CompressAMDBC1(img) is None
