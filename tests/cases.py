"""Shared parity inputs: seeded synthetic images and the reference's own test patterns."""
import numpy as np

from gfx_imagecompress_b200 import synth


def scalar_cases():
    """(name, pixels, fmt) for BC4/BC5."""
    out = [
        ("height_rg8_256", synth.height_rg8(256, 256, 2), synth.FMT_RG8),
        ("gradnoise_rgba8_npot", synth.rgba8_gradnoise(257, 131, 1, "ramp"), synth.FMT_RGBA8),
        ("height_r8", np.ascontiguousarray(synth.height_rg8(128, 64, 5)[..., :1]), synth.FMT_R8),
        ("tiny_3x5", synth.rgba8_gradnoise(3, 5, 9, "ramp"), synth.FMT_RGBA8),
        ("one_texel", synth.rgba8_gradnoise(1, 1, 9, "ramp"), synth.FMT_RGBA8),
    ]
    for name in ("R", "G", "RGB", "RGB_Punchthrough", "RGBA"):
        p, f = synth.pattern(name, 64, 64)
        out.append(("pattern_" + name, p, f))
    p, f = synth.pattern("RGB", 257, 257)
    out.append(("pattern_RGB_257", p, f))
    rng = np.random.default_rng(7)
    out.append(("uniform_noise_rg8", rng.integers(0, 256, (64, 64, 2), dtype=np.uint8), synth.FMT_RG8))
    ends = rng.choice(np.array([0, 1, 2, 253, 254, 255, 128], np.uint8), (64, 64, 2))
    out.append(("endpoint_values_rg8", np.ascontiguousarray(ends), synth.FMT_RG8))
    return out


def random_scalar_blocks(n=4096, seed=0):
    rng = np.random.default_rng(seed)
    blocks = np.empty((n, 16), np.float32)
    for i in range(n):
        k = i % 4
        if k == 0:
            v = rng.random(16, dtype=np.float32)
        elif k == 1:
            v = rng.integers(0, 256, 16).astype(np.float32) / np.float32(255)
        elif k == 2:
            v = (rng.integers(0, 40, 16) + rng.integers(0, 216)).astype(np.float32) / np.float32(255)
        else:
            v = rng.choice(np.array([0, 1 / 255, 2 / 255, 254 / 255, 1.0, 0.5, 0.3], np.float32), 16)
        blocks[i] = v
    return blocks


def to_blocks_rgba8(img: np.ndarray) -> np.ndarray:
    """(H, W, 4) uint8 with H, W multiples of 4 -> (nblocks, 16) uint32 packed RGBA8 in block order."""
    h, w, _ = img.shape
    return np.ascontiguousarray(img.reshape(h // 4, 4, w // 4, 4, 4).transpose(0, 2, 1, 3, 4).reshape(-1, 16, 4)).view(np.uint32).reshape(-1, 16)


def rgba_cases(small: bool = False):
    """(name, (H, W, C) uint8, fmt) colour images for BC1 / BC7: seeded gradient+noise with the alpha layouts of
    SURVEY.md 8d, the reference's own test patterns (tests/test_imagecompress.cpp:14-126) incl. the NPOT size,
    white noise, smooth low-contrast ramps and flat blocks."""
    n = 32 if small else 64
    rng = np.random.default_rng(11)
    out = [
        ("gradnoise_lefthalf", synth.rgba8_gradnoise(n * 2, n, 3, "lefthalf"), synth.FMT_RGBA8),
        ("gradnoise_opaque", synth.rgba8_gradnoise(n, n, 4, "opaque"), synth.FMT_RGBA8),
        ("gradnoise_punch", synth.rgba8_gradnoise(n, n, 1, "punch"), synth.FMT_RGBA8),
        ("gradnoise_npot", synth.rgba8_gradnoise(37, 21, 5, "ramp"), synth.FMT_RGBA8),
        ("noise_rgba", rng.integers(0, 256, (n // 2, n // 2, 4), dtype=np.uint8), synth.FMT_RGBA8),
        ("noise_rgb", rng.integers(0, 256, (n // 2, n // 2, 3), dtype=np.uint8), synth.FMT_RGB8),
    ]
    for name in ("R", "RGB", "RGB_Punchthrough", "RGBA"):
        p, f = synth.pattern(name, n, n)
        out.append(("pattern_" + name, p, f))
    p, f = synth.pattern("RGB", 33 if small else 257, 33 if small else 257)
    out.append(("pattern_RGB_npot", p, f))
    x = np.arange(n)[None, :]
    y = np.arange(n)[:, None]
    smooth = np.stack([(x + y) // 2 + 100, (x * 2) % 256 + 0 * y, (y * 3) % 256 + 0 * x, np.full((n, n), 255)], 2).astype(np.uint8)
    out.append(("smooth_ramps", np.ascontiguousarray(smooth), synth.FMT_RGBA8))
    lowc = (rng.integers(0, 6, (n // 2, n // 2, 4)) + np.array([120, 60, 200, 250])).astype(np.uint8)
    out.append(("low_contrast", lowc, synth.FMT_RGBA8))
    return out


def to_blocks_f32(img: np.ndarray) -> np.ndarray:
    """(H, W, 4) uint8 -> (nblocks, 64) float32 exactly as the shim's Image_GetPixelAtF decodes (x / 255.0f)."""
    u = to_blocks_rgba8(np.ascontiguousarray(img)).view(np.uint8).reshape(-1, 16, 4)
    return np.ascontiguousarray((u.astype(np.float32) / np.float32(255)).reshape(-1, 64))
