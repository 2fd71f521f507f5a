"""bc7enc16 path (bit-exact). CPU: the host build of the kernel's per-block core vs the compiled reference.
GPU: the CUDA kernel through the C-ABI vs the compiled reference."""
import numpy as np
import pytest

import cases
from gfx_imagecompress_b200 import synth
from oracle.ref import BC7_RG, default_opts


def _ref_block(ref, b, perceptual, fast):
    out = np.zeros(16, np.uint8)
    ref.lib.Image_CompressRichGel999BC7enc16(b.ctypes.data, fast, perceptual, out.ctypes.data)
    return out


@pytest.mark.parametrize("perceptual,fast", [(True, False), (False, False), (True, True), (False, True)])
def test_core_hostbuild_matches_reference(ref, perceptual, fast):
    import hostbuild
    L = hostbuild.load()
    for name, px, fmt in cases.rgba_cases(small=True):
        if fmt != synth.FMT_RGBA8 or px.shape[0] % 4 or px.shape[1] % 4:
            continue
        blocks = cases.to_blocks_rgba8(px)
        got = hostbuild.bc7rg_blocks(L, blocks, perceptual, fast)
        want = np.stack([_ref_block(ref, b, perceptual, fast) for b in blocks])
        bad = np.flatnonzero((got != want).any(axis=1))
        assert bad.size == 0, f"{name}: {bad.size}/{len(want)} blocks differ, first {bad[:5]}"


@pytest.mark.gpu
def test_images_bit_exact(engine, ref):
    for name, px, fmt in cases.rgba_cases():
        got = engine.encode_host(engine.BC7_RG, px, fmt)
        want = ref.encode(BC7_RG, px, fmt)
        bad = np.flatnonzero((got != want).any(axis=1))
        assert bad.size == 0, f"{name}: {bad.size}/{len(want)} blocks differ, first {bad[:5]}"


@pytest.mark.gpu
@pytest.mark.parametrize("perceptual,fast", [(False, False), (True, True), (False, True)])
def test_options_bit_exact(engine, ref, perceptual, fast):
    px = synth.rgba8_gradnoise(128, 64, 21, "lefthalf")
    got = engine.encode_host(engine.BC7_RG, px, synth.FMT_RGBA8, engine.Opts.default(rg_perceptual=int(perceptual), rg_fast=int(fast)))
    want = ref.encode(BC7_RG, px, synth.FMT_RGBA8, opts=default_opts(rg_perceptual=perceptual, rg_fast=fast))
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_float_source_and_block_api(engine, ref):
    p, f = synth.pattern("FloatRGBA", 32, 32)
    assert np.array_equal(engine.encode_host(engine.BC7_RG, p, f), ref.encode(BC7_RG, p, f))
    blocks = cases.to_blocks_rgba8(synth.rgba8_gradnoise(64, 64, 8, "lefthalf"))
    got = engine.encode_blocks(engine.BC7_RG, blocks, 107)
    want = np.stack([_ref_block(ref, b, True, False) for b in blocks])
    assert np.array_equal(got, want)
    assert np.array_equal(engine.Image_CompressRichGel999BC7enc16(blocks[3]), want[3])


@pytest.mark.gpu
def test_image_api(engine, ref):
    px = synth.rgba8_gradnoise(257, 257, 3, "lefthalf")
    dst = engine.Image_CompressRichGel999BC7(engine.Image(px, synth.FMT_RGBA8))
    assert dst is not None and (dst.width, dst.height) == (260, 260) and dst.format == 26  # DXBC7_UNORM
    assert np.array_equal(dst.blocks(16), ref.encode(BC7_RG, px, synth.FMT_RGBA8))
    fast = engine.ImageCompress_Compress(7, True, engine.Image(px, synth.FMT_RGBA8))  # Image_CT_DXBC7, fast -> bc7enc16
    assert np.array_equal(fast.blocks(16), dst.blocks(16))


@pytest.mark.gpu
def test_large_image_sampled_rows(engine, ref):
    """2048^2: every block of 16 evenly spaced block-rows must match the reference; whole image checked by checksum
    stability across two runs (determinism)."""
    px = synth.rgba8_gradnoise(2048, 2048, 3, "lefthalf")
    got = engine.encode_host(engine.BC7_RG, px, synth.FMT_RGBA8)
    again = engine.encode_host(engine.BC7_RG, px, synth.FMT_RGBA8)
    assert np.array_equal(got, again)
    bx = 512
    for r in range(0, 512, 32):
        want = ref.encode(BC7_RG, px, synth.FMT_RGBA8, rows=(r, r + 1))
        assert np.array_equal(got[r * bx:(r + 1) * bx], want), f"block-row {r}"
