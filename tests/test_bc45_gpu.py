"""GPU parity (bit-exact): BC4/BC5 kernels through the C-ABI vs the oracle."""
import numpy as np
import pytest

import cases
from gfx_imagecompress_b200 import synth

pytestmark = pytest.mark.gpu


def _oracle_bc45(ref_or_none, restated, codec, px, fmt):
    if px.dtype == np.uint8:
        return restated.bc4(px) if codec == 4 else restated.bc5(px)
    return ref_or_none.encode(codec, px, fmt)


@pytest.mark.parametrize("codec", [4, 5])
def test_images_bit_exact(engine, restated, codec):
    for name, px, fmt in cases.scalar_cases():
        if px.dtype != np.uint8:
            continue
        want = restated.bc4(px) if codec == 4 else restated.bc5(px)
        got = engine.encode_host(codec, px, fmt)
        bad = np.flatnonzero((got != want).any(axis=1))
        assert bad.size == 0, f"{name}: {bad.size}/{len(want)} blocks differ, first {bad[:5]}"


def test_against_compiled_reference(engine, ref):
    px = synth.height_rg8(512, 512, 11)
    for codec in (4, 5):
        assert np.array_equal(engine.encode_host(codec, px, synth.FMT_RG8), ref.encode(codec, px, synth.FMT_RG8))


def test_block_api_bit_exact(engine, restated):
    blocks = cases.random_scalar_blocks(4096, seed=1)
    got = engine.encode_blocks(engine.BC4, blocks, 101)
    want = np.stack([restated.alpha_block(b) for b in blocks])
    bad = np.flatnonzero((got != want).any(axis=1))
    assert bad.size == 0, f"{bad.size} blocks differ, first {bad[:5]}"
    one = engine.Image_CompressAMDAlphaSingleModeBlock(blocks[5])
    assert np.array_equal(one, want[5])


def test_image_api_shape_and_format(engine, restated):
    # mirrors reference tests/test_imagecompress.cpp: non-null, dims rounded up to 4, output format tag
    px = synth.height_rg8(257, 257, 3)
    img = engine.Image(px, synth.FMT_RG8)
    dst = engine.Image_CompressAMDBC5(img)
    assert dst is not None and (dst.width, dst.height, dst.depth, dst.slices) == (260, 260, 1, 1)
    assert dst.format == 22  # TinyImageFormat_DXBC5_UNORM
    assert np.array_equal(dst.blocks(16), restated.bc5(px))
    dst4 = engine.ImageCompress_Compress(4, False, img)  # Image_CT_DXBC4
    assert dst4.format == 20 and np.array_equal(dst4.blocks(8), restated.bc4(px))


def test_progress_cancel(engine):
    px = synth.height_rg8(64, 64, 3)
    calls = []
    assert engine.Image_CompressAMDBC4(engine.Image(px, synth.FMT_RG8), progress=lambda p: calls.append(p) or True) is None
    assert calls


def test_config2_full_size(engine, ref):
    """BASELINE config 2: 4096x4096 RG8 height/normal map, every block bit-identical to the reference."""
    px = synth.height_rg8(4096, 4096, 2)
    for codec in (4, 5):
        got = engine.encode_host(codec, px, synth.FMT_RG8)
        want = ref.encode(codec, px, synth.FMT_RG8)
        assert np.array_equal(got, want)
