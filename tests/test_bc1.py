"""BC1 path (bit-exact). CPU: host build of the kernel's per-block core vs the reference block API.
GPU: CUDA kernel through the C-ABI vs the compiled reference (BASELINE config[0]: 1024^2 gradient+noise with a
punch-through quadrant, through the reference's C API)."""
import numpy as np
import pytest

import cases
from gfx_imagecompress_b200 import synth
from oracle.ref import BC1, default_opts


@pytest.mark.parametrize("thr,steps,r3d", [(128 / 255.0, 1, False), (0.0, 1, False), (128 / 255.0, 2, False), (128 / 255.0, 1, True)])
def test_core_hostbuild_matches_reference(ref, thr, steps, r3d):
    """r3d = the b3DRefinement option (Refine3D, src/amd_bcx_body.cpp:808-932): bit 8 of the core's `steps` argument."""
    import hostbuild
    L = hostbuild.load()
    for name, px, fmt in cases.rgba_cases(small=True):
        if px.shape[2] != 4 or px.shape[0] % 4 or px.shape[1] % 4:
            continue
        fb = cases.to_blocks_f32(px)
        if r3d:
            fb = fb[:48]  # 729 candidate ramps per fit: keep the CPU suite short
        got = hostbuild.bc1_blocks(L, fb, thr, steps | (0x100 if r3d else 0))
        want = np.stack([_ref_block(ref, b, thr, steps, r3d) for b in fb])
        bad = np.flatnonzero((got != want).any(axis=1))
        assert bad.size == 0, f"{name}: {bad.size}/{len(want)} blocks differ, first {bad[:5]}"


def _ref_block(ref, b, thr, steps, r3d=False):
    out = np.zeros(8, np.uint8)
    ref.lib.Image_CompressAMDBC1Block(np.ascontiguousarray(b, np.float32).ctypes.data, False, r3d, steps, thr, out.ctypes.data)
    return out


@pytest.mark.gpu
def test_images_bit_exact(engine, ref):
    for name, px, fmt in cases.rgba_cases():
        got = engine.encode_host(engine.BC1, px, fmt)
        want = ref.encode(BC1, px, fmt)
        bad = np.flatnonzero((got != want).any(axis=1))
        assert bad.size == 0, f"{name}: {bad.size}/{len(want)} blocks differ, first {bad[:5]}"


@pytest.mark.gpu
def test_config1_full_size(engine, ref):
    px = synth.rgba8_gradnoise(1024, 1024, 1, "punch")
    got = engine.encode_host(engine.BC1, px, synth.FMT_RGBA8)
    want = ref.encode(BC1, px, synth.FMT_RGBA8)
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_options_and_image_api(engine, ref):
    px, fmt = synth.pattern("RGB_Punchthrough", 256, 256)
    # reference test: UseAlpha = true, AlphaThreshold = 128 (tests/test_imagecompress.cpp:173-175)
    dst = engine.Image_CompressAMDBC1(engine.Image(px, fmt), options=(True, 128))
    assert dst is not None and (dst.width, dst.height) == (256, 256) and dst.format == 14  # DXBC1_RGBA_UNORM
    want = ref.encode(BC1, px, fmt, opts=default_opts(bc1_use_alpha=1, bc1_alpha_threshold=128))
    assert np.array_equal(dst.blocks(8), want)
    # alpha threshold 0 disables punch-through; RefinementSteps = 2
    dst = engine.Image_CompressAMDBC1(engine.Image(px, fmt), amdOptions=(False, False, 2, 0xFF), options=(False, 0))
    assert dst.format == 12  # DXBC1_RGB_UNORM
    want = ref.encode(BC1, px, fmt, opts=default_opts(bc1_use_alpha=0, bc1_alpha_threshold=0, amd_refinement_steps=2))
    assert np.array_equal(dst.blocks(8), want)
    d1 = engine.ImageCompress_Compress(1, False, engine.Image(px, fmt))  # Image_CT_DXBC1
    assert np.array_equal(d1.blocks(8), ref.encode(BC1, px, fmt))
    # b3DRefinement (Refine3D): 729 candidate ramps per fit, still the reference's bytes
    small = np.ascontiguousarray(px[:64, :64])
    dst = engine.Image_CompressAMDBC1(engine.Image(small, fmt), amdOptions=(True, False, 1, 0xFF))
    want = ref.encode(BC1, small, fmt, opts=default_opts(amd_3d_refinement=1))
    assert dst is not None and np.array_equal(dst.blocks(8), want)
    # unsupported knobs fail loudly (NULL), never a silent fallback: AdaptiveColourWeights reads uninitialised memory in the reference
    assert engine.Image_CompressAMDBC1(engine.Image(px, fmt), amdOptions=(False, True, 1, 0xFF)) is None


@pytest.mark.gpu
def test_block_api(engine, ref):
    fb = cases.to_blocks_f32(synth.rgba8_gradnoise(64, 64, 5, "punch"))
    got = engine.encode_blocks(engine.BC1, fb.reshape(-1, 16, 4), 104)
    want = np.stack([ref.bc1_block(b) for b in fb])
    assert np.array_equal(got, want)
    assert np.array_equal(engine.Image_CompressAMDBC1Block(fb[9]), want[9])
