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
