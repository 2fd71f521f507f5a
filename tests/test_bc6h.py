"""BC6H path (unsigned half sources; BASELINE config[3]). The search is FP32 in the reference's operation order, so the
blocks are expected to be bit-identical to the compiled reference; the contractual gate is <= 0.02 dB PSNR.
CPU: host build of the kernel's per-block core vs the reference's BC6HBlockEncoder::CompressBlock.
GPU: CUDA kernel through the C-ABI vs the compiled reference's image API."""
import numpy as np
import pytest

from gfx_imagecompress_b200 import synth
from oracle.ref import BC6H


def hdr_blocks(img16: np.ndarray) -> np.ndarray:
    h, w, _ = img16.shape
    f = img16.astype(np.float32)
    return np.ascontiguousarray(f.reshape(h // 4, 4, w // 4, 4, 4).transpose(0, 2, 1, 3, 4).reshape(-1, 64))


def hdr_cases(n: int = 64):
    """(name, (H, W, 4) float16). No exactly-flat subsets (the reference reads an uninitialised direction there)."""
    rng = np.random.default_rng(21)
    out = [("hdr_ramp_sun", synth.hdr_rgba16f(n, n, 4))]
    # wide exponent spread incl. values close to the half maximum (clampF16Max) and below the 1e-5 flush
    e = rng.uniform(-18, 15.9, (n // 2, n // 2, 4))
    wide = (2.0 ** e).astype(np.float16)
    wide[..., 3] = 1.0
    out.append(("wide_exponents", wide))
    lowv = (rng.uniform(0, 3e-5, (n // 2, n // 2, 4))).astype(np.float16)
    lowv[::2, ::2, :3] += np.float16(0.25)
    out.append(("tiny_and_quarter", lowv))
    smooth = np.empty((n // 2, n // 2, 4), np.float32)
    y, x = np.mgrid[0:n // 2, 0:n // 2]
    smooth[..., 0] = 0.5 + x / 64.0 + y / 512.0
    smooth[..., 1] = 2.0 + y / 16.0 + x / 1024.0
    smooth[..., 2] = 0.01 + (x + y) / 4096.0
    smooth[..., 3] = 1.0
    out.append(("smooth", smooth.astype(np.float16)))
    edge = np.where(((x // 3 + y // 5) & 1)[..., None] > 0, np.float32(40.0), np.float32(0.125)) + rng.uniform(0, 0.01, (n // 2, n // 2, 1))
    edge = np.repeat(edge, 4, axis=2) * np.array([1.0, 0.5, 0.25, 1.0])
    out.append(("hard_edges", edge.astype(np.float16)))
    return out


def test_core_hostbuild_matches_reference(ref):
    import hostbuild
    L = hostbuild.load()
    for name, img in hdr_cases(32):
        got = hostbuild.bc6h_blocks(L, hdr_blocks(img))
        want = ref.encode(BC6H, np.ascontiguousarray(img), synth.FMT_RGBA16UF)
        bad = np.flatnonzero((got != want).any(axis=1))
        assert bad.size == 0, f"{name}: {bad.size}/{len(want)} blocks differ, first {bad[:5]}"


@pytest.mark.gpu
def test_images_match_reference(engine, ref):
    for name, img in hdr_cases(64):
        img = np.ascontiguousarray(img)
        got = engine.encode_host(engine.BC6H, img, synth.FMT_RGBA16UF)
        want = ref.encode(BC6H, img, synth.FMT_RGBA16UF)
        bad = np.flatnonzero((got != want).any(axis=1))
        assert bad.size == 0, f"{name}: {bad.size}/{len(want)} blocks differ, first {bad[:5]}"


@pytest.mark.gpu
def test_npot_and_float32_source(engine, ref):
    img = synth.hdr_rgba16f(37, 21, 9)
    got = engine.encode_host(engine.BC6H, img, synth.FMT_RGBA16UF)
    want = ref.encode(BC6H, img, synth.FMT_RGBA16UF)
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_config4_sampled_rows(engine, ref):
    """BASELINE config[3]: 4096^2 unsigned half environment map. The reference needs ~5 h of CPU for the whole image,
    so it encodes evenly spaced block-rows; the GPU encodes the full image and must agree on those rows."""
    w = h = 4096
    img = synth.hdr_rgba16f(w, h, 4)
    got = engine.encode_host(engine.BC6H, img, synth.FMT_RGBA16UF).reshape(h // 4, w // 4, 16)
    for by in (0, 255, 256, 700, 1023):
        want = ref.encode(BC6H, img, synth.FMT_RGBA16UF, rows=(by, by + 1))
        assert np.array_equal(got[by], want), f"block-row {by} differs"


@pytest.mark.gpu
def test_block_api_and_image_api(engine, ref):
    img = synth.hdr_rgba16f(64, 32, 6)
    want = ref.encode(BC6H, img, synth.FMT_RGBA16UF)
    got = engine.encode_blocks(engine.BC6H, hdr_blocks(img).reshape(-1, 16, 4), 104)
    assert np.array_equal(got, want)
    dst = engine.Image_CompressAMDBC6H(engine.Image(img, synth.FMT_RGBA16UF))
    assert dst is not None and (dst.width, dst.height) == (64, 32)
    assert np.array_equal(dst.blocks(16), want)


def signed_cases(n: int = 32):
    """(name, (H, W, 4) float32) for the signed path (any source that is "float && signed" in the reference's terms:
    R16G16B16A16_SFLOAT / R32G32B32A32_SFLOAT, e.g. its own FloatRGBA test pattern)."""
    rng = np.random.default_rng(5)
    out = [("pattern_FloatRGBA", synth.pattern("FloatRGBA", n, n)[0].astype(np.float32))]
    a = rng.normal(0, 2.0, (n, n, 4)).astype(np.float16).astype(np.float32)
    a[..., 3] = 1
    out.append(("gauss_plus_minus", a))
    b = (rng.uniform(-1, 1, (n, n, 4)) * np.array([100, 1, 0.01, 1])).astype(np.float16).astype(np.float32)
    out.append(("scaled_plus_minus", b))
    out.append(("hdr_positive", synth.hdr_rgba16f(n, n, 4).astype(np.float32)))
    return out


def _ref_signed_blocks(ref, fb):
    want = np.zeros((len(fb), 16), np.uint8)
    for i, b in enumerate(fb):
        b = np.ascontiguousarray(b)
        ref.lib.ref_bc6h_block(b.ctypes.data, 1, 0xFF, want[i].ctypes.data)
    return want


def test_core_hostbuild_signed_matches_reference(ref):
    import hostbuild
    L = hostbuild.load()
    for name, img in signed_cases(16):
        fb = hdr_blocks(img)
        got = hostbuild.bc6h_blocks(L, fb, True)
        want = _ref_signed_blocks(ref, fb)
        bad = np.flatnonzero((got != want).any(axis=1))
        assert bad.size == 0, f"{name}: {bad.size}/{len(want)} blocks differ, first {bad[:5]}"


@pytest.mark.gpu
def test_signed_sources_match_reference(engine, ref):
    """RGBA32F / RGBA16F sources take the reference's signed path (src/amd_bc6h_compressor.cpp:18-28)."""
    for name, img in signed_cases(32):
        img = np.ascontiguousarray(img, np.float32)
        got = engine.encode_host(engine.BC6H, img, synth.FMT_RGBA32F)
        want = ref.encode(BC6H, img, synth.FMT_RGBA32F)
        bad = np.flatnonzero((got != want).any(axis=1))
        assert bad.size == 0, f"{name}: {bad.size}/{len(want)} blocks differ, first {bad[:5]}"
    p, f = synth.pattern("FloatRGBA", 64, 64)
    dst = engine.Image_CompressAMDBC6H(engine.Image(p, f))
    assert dst is not None and dst.format == 25  # DXBC6H_SFLOAT
    assert np.array_equal(dst.blocks(16), ref.encode(BC6H, p, f))
