"""BC2 / BC3 (reference src/amd_bc2_compressor.cpp:11-60, src/amd_bc3_compressor.cpp:11-60).

Alpha halves (bytes 0..7) are bit-exact with the compiled reference: BC3 = Image_CompressAMDAlphaSingleModeBlock on the
alpha channel, BC2 = Image_CompressAMDExplictAlphaSingleModeBlock's 4-bit rounding.  The reference's colour half is
stack-layout dependent (CompRGBBlock, src/amd_bcx_body.cpp:1299-1362, reads its stride-3 input with stride 4 and writes
past fBlk[48]); the defined behaviour here is the BC1 4-point fit it is a copy of -- checked against the host build of
that fit, against the BC1 block API where the 4-point fit wins there, and by decoded PSNR against the reference's."""
import numpy as np
import pytest

import cases
from gfx_imagecompress_b200 import synth
from oracle import metrics
from oracle.ref import BC2, BC3


def _alpha_blocks_f32(px):
    a = px[..., 3] if px.shape[2] == 4 else np.full(px.shape[:2], 255, np.uint8)
    h, w = a.shape
    b = a.reshape(h // 4, 4, w // 4, 4).transpose(0, 2, 1, 3).reshape(-1, 16)
    return np.ascontiguousarray(b.astype(np.float32) / np.float32(255))


def test_reference_alpha_halves_are_the_block_functions(ref):
    """What the oracle pins: the alpha half of the reference's BC3 / BC2 image output is its own block function."""
    px = synth.rgba8_gradnoise(64, 32, 3, "lefthalf")
    ab = _alpha_blocks_f32(px)
    ref.lib.Image_CompressAMDExplictAlphaSingleModeBlock.argtypes = ref.lib.Image_CompressAMDAlphaSingleModeBlock.argtypes
    want3 = ref.encode(BC3, px, synth.FMT_RGBA8)
    want2 = ref.encode(BC2, px, synth.FMT_RGBA8)
    for i, b in enumerate(ab):
        assert np.array_equal(ref.alpha_block(b), want3[i, :8])
        out = np.zeros(8, np.uint8)
        ref.lib.Image_CompressAMDExplictAlphaSingleModeBlock(b.ctypes.data, out.ctypes.data)
        assert np.array_equal(out, want2[i, :8])


@pytest.mark.gpu
@pytest.mark.parametrize("codec,refc", [(3, BC3), (2, BC2)], ids=["bc3", "bc2"])
def test_images(engine, ref, codec, refc):
    import hostbuild
    L = hostbuild.load()
    for name, px, fmt in cases.rgba_cases(small=True):
        got = engine.encode_host(codec, px, fmt)
        want = ref.encode(refc, px, fmt)
        assert got.shape == want.shape == (((px.shape[0] + 3) // 4) * ((px.shape[1] + 3) // 4), 16)
        assert np.array_equal(got[:, :8], want[:, :8]), f"{name}: alpha halves differ from the reference"
        if px.shape[0] % 4 == 0 and px.shape[1] % 4 == 0 and px.shape[2] == 4:
            fb = cases.to_blocks_f32(px)
            colour = hostbuild.bc23_colour_blocks(L, fb)
            assert np.array_equal(got[:, 8:], colour), f"{name}: colour halves differ from the host build of the 4-point fit"
            # decoded colour quality vs the reference's (its colour bytes are whatever its stack held: usually worse)
            rgb = px[..., :3]
            p_got = metrics.psnr_rgb_bc1(got[:, 8:], rgb)
            p_ref = metrics.psnr_rgb_bc1(want[:, 8:], rgb)
            assert p_got >= p_ref - 0.5, f"{name}: colour PSNR {p_got:.2f} dB vs reference {p_ref:.2f} dB"


@pytest.mark.gpu
def test_image_api_and_pick_then_compress(engine, ref):
    px = synth.rgba8_gradnoise(64, 32, 3, "lefthalf")
    img = engine.Image(px, synth.FMT_RGBA8)
    d3 = engine.Image_CompressAMDBC3(img)
    d2 = engine.Image_CompressAMDBC2(img)
    assert d3 is not None and d2 is not None and (d3.width, d3.height) == (64, 32)
    assert np.array_equal(d3.blocks(16), engine.encode_host(3, px, synth.FMT_RGBA8))
    assert np.array_equal(d2.blocks(16), engine.encode_host(2, px, synth.FMT_RGBA8))
    # the reference's pick-then-compress flow: RGBA + AllowDXBC1to5 picks DXBC3 (src/imagecompress.cpp:104-110)
    t = engine.ImageCompress_PickCompressionType(1, img)
    assert t == 3
    out = engine.ImageCompress_Compress(t, False, img)
    assert out is not None and np.array_equal(out.blocks(16), d3.blocks(16))


@pytest.mark.gpu
def test_block_api(engine, ref):
    px = synth.rgba8_gradnoise(32, 32, 9, "lefthalf")
    fb = cases.to_blocks_f32(px).reshape(-1, 16, 4)
    ref.lib.Image_CompressAMDExplictAlphaSingleModeBlock.argtypes = ref.lib.Image_CompressAMDAlphaSingleModeBlock.argtypes
    whole = engine.encode_host(2, px, synth.FMT_RGBA8)
    for i in (0, 5, 17, 63):
        a = np.ascontiguousarray(fb[i, :, 3])
        want = np.zeros(8, np.uint8)
        ref.lib.Image_CompressAMDExplictAlphaSingleModeBlock(a.ctypes.data, want.ctypes.data)
        assert np.array_equal(engine.Image_CompressAMDExplictAlphaSingleModeBlock(a), want)
        rgb = np.ascontiguousarray(fb[i, :, :3])
        assert np.array_equal(engine.Image_CompressAMDRGBSingleModeBlock(rgb), whole[i, 8:])


@pytest.mark.gpu
def test_unsupported_block_arguments_fail_loudly(engine, capfd):
    """void block functions cannot return an error: the block is filled with 0xFF (never left uninitialised) and the
    reason goes to stderr / b200ic_last_error."""
    fb = cases.to_blocks_f32(synth.rgba8_gradnoise(8, 8, 1, "opaque"))[0]
    out = engine.Image_CompressAMDMultiModeLDRBlock(fb, quality=0.5)
    assert (out == 0xFF).all()
    out = engine.Image_CompressAMDBC1Block(fb, adaptive=True)
    assert (out == 0xFF).all()
    err = capfd.readouterr().err
    assert "Image_CompressAMDMultiModeLDRBlock" in err and "Image_CompressAMDBC1Block" in err
