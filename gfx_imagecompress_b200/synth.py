"""Deterministic synthetic inputs for the BCn encode path.

Everything derives from a stateless 32-bit integer hash of (seed, texel index), so any host (and any
rank of a sharded run) regenerates identical bytes.  Also holds the reference's own test-pattern
generators (reference tests/test_imagecompress.cpp:14-126), restated on u8/f32 arrays.
"""
from __future__ import annotations

import numpy as np

# TinyImageFormat tags of the compat shim (compat/tiny_imageformat/tinyimageformat_base.h)
FMT_R8 = 1
FMT_RG8 = 3
FMT_RGB8 = 5
FMT_RGBA8 = 7
FMT_RGBA8_SRGB = 8
FMT_RGBA16F = 9
FMT_RGBA32F = 10
FMT_RGBA16UF = 11  # private "unsigned half" tag -> BC6H unsigned path

_BPP = {FMT_R8: 1, FMT_RG8: 2, FMT_RGB8: 3, FMT_RGBA8: 4, FMT_RGBA8_SRGB: 4, FMT_RGBA16F: 8, FMT_RGBA16UF: 8,
        FMT_RGBA32F: 16}


def bytes_per_texel(fmt: int) -> int:
    return _BPP[fmt]


def hash32(idx: np.ndarray, seed: int) -> np.ndarray:
    """xorshift-multiply hash of (seed, idx) -> uint32 (vectorised)."""
    with np.errstate(over="ignore"):
        h = idx.astype(np.uint32) * np.uint32(0x9E3779B1) + np.uint32((seed * 0x85EBCA6B + 0xC2B2AE35) & 0xFFFFFFFF)
        h ^= h >> np.uint32(16)
        h *= np.uint32(0x7FEB352D)
        h ^= h >> np.uint32(15)
        h *= np.uint32(0x846CA68B)
        h ^= h >> np.uint32(16)
    return h


def _grid(w: int, h: int, y0: int = 0, rows: int | None = None):
    rows = h if rows is None else rows
    y = np.arange(y0, y0 + rows, dtype=np.int64)[:, None]
    x = np.arange(w, dtype=np.int64)[None, :]
    return x, y


def rgba8_gradnoise(w: int, h: int, seed: int, alpha: str = "opaque", y0: int = 0, rows: int | None = None) -> np.ndarray:
    """Gradient + uniform noise RGBA8.  alpha: 'opaque' | 'punch' (A=0 in the bottom-right quadrant, C1)
    | 'lefthalf' (noisy alpha ramp over the left half, opaque right half, C3) | 'ramp' (alpha ramp everywhere).
    Rows [y0, y0+rows) of the full image are produced (used for block-row sharding)."""
    x, y = _grid(w, h, y0, rows)
    hsh = hash32((y * w + x).astype(np.uint64) & 0xFFFFFFFF, seed)
    n = (hsh & np.uint32(31)).astype(np.int64) - 16
    n2 = ((hsh >> np.uint32(8)) & np.uint32(31)).astype(np.int64) - 16
    out = np.empty((x.shape[1] * 0 + y.shape[0], w, 4), np.uint8)
    out[..., 0] = np.clip((255 * x) // w + n, 0, 255)
    out[..., 1] = np.clip((255 * y) // h + n, 0, 255)
    out[..., 2] = np.clip((255 * (x + y)) // (w + h) - n, 0, 255)
    if alpha == "opaque":
        out[..., 3] = 255
    elif alpha == "punch":
        out[..., 3] = np.where((x >= w // 2) & (y >= h // 2), 0, 255)
    elif alpha == "lefthalf":
        a = np.clip(255 - (255 * x) // w + n2, 0, 255)
        out[..., 3] = np.where(x < w // 2, a, 255)
    elif alpha == "ramp":
        out[..., 3] = np.clip(255 - (255 * x) // w + n2, 0, 255)
    else:
        raise ValueError(alpha)
    return out


def _value_noise(w: int, h: int, seed: int, octaves: int = 8) -> np.ndarray:
    """Integer-lattice value noise in [0,1), bilinear, summed over octaves (float64, deterministic)."""
    x, y = _grid(w, h)
    acc = np.zeros((h, w), np.float64)
    amp, tot = 1.0, 0.0
    for o in range(octaves):
        cell = max(1, 256 >> o)
        gx, gy = x // cell, y // cell
        fx, fy = (x % cell) / cell, (y % cell) / cell

        def lat(ix, iy):
            return (hash32(((iy * 65537 + ix) & 0xFFFFFFFF).astype(np.uint64), seed + 17 * o) >> np.uint32(8)).astype(np.float64) / 16777216.0

        v = (lat(gx, gy) * (1 - fx) + lat(gx + 1, gy) * fx) * (1 - fy) + (lat(gx, gy + 1) * (1 - fx) + lat(gx + 1, gy + 1) * fx) * fy
        acc += amp * v
        tot += amp
        amp *= 0.5
    return acc / tot


def height_rg8(w: int, h: int, seed: int) -> np.ndarray:
    """C2 input: ch0 = value-noise height (contrast-stretched so that many blocks exceed the 48/256 range
    that enables the reference's global search), ch1 = x-derivative remapped to 0..255; a band of flat
    and two-valued blocks is stamped in for the <=2-unique-values early-outs."""
    hgt = _value_noise(w, h, seed)
    hs = np.clip((hgt - 0.5) * 2.2 + 0.5, 0.0, 1.0)
    hsh = hash32((np.arange(h, dtype=np.uint64)[:, None] * w + np.arange(w, dtype=np.uint64)[None, :]) & 0xFFFFFFFF, seed + 1)
    grain = ((hsh & np.uint32(63)).astype(np.float64) - 32.0) / 255.0
    sharp = ((hsh >> np.uint32(6)) & np.uint32(63)) == 0  # 1/64 of texels get strong grain -> wide-range blocks
    ch0 = np.clip(hs + np.where(sharp, grain * 3.0, grain * 0.15), 0, 1)
    dx = np.zeros_like(hs)
    dx[:, 1:-1] = (hs[:, 2:] - hs[:, :-2]) * 6.0
    sharp1 = ((hsh >> np.uint32(12)) & np.uint32(63)) == 0
    ch1 = np.clip(0.5 + dx + np.where(sharp1, grain * 3.0, grain * 0.1), 0, 1)
    out = np.empty((h, w, 2), np.uint8)
    out[..., 0] = np.floor(ch0 * 255.0 + 0.5)
    out[..., 1] = np.floor(ch1 * 255.0 + 0.5)
    # flat + two-valued band (about 1/64 of the rows at the top, block aligned)
    band = max(4, (h // 64) & ~3)
    out[:band, : w // 2, :] = 77
    xx = np.arange(w // 2, w)[None, :]
    out[:band, w // 2:, 0] = np.where((xx // 2) & 1, 30, 200)
    out[:band, w // 2:, 1] = np.where((xx // 3) & 1, 0, 255)
    return out


def hdr_rgba16f(w: int, h: int, seed: int) -> np.ndarray:
    """C4 input: smooth HDR ramp + per-texel dither, a 'sun' up to ~1000, stored as RGBA half (A=1). No
    exactly-flat blocks (the reference's BC6H has UB on flat subsets, SURVEY.md 7 hard part 5)."""
    x, y = _grid(w, h)
    hsh = hash32(((y * w + x) & 0xFFFFFFFF).astype(np.uint64), seed)
    g = 4.0 * x / w
    out = np.empty((h, w, 4), np.float32)
    for c in range(3):
        u = ((hsh >> np.uint32(8 * c)) & np.uint32(255)).astype(np.float64)
        out[..., c] = g * (c + 1) / 3.0 + u / 1024.0 + y / (16.0 * h)
    cx, cy, r = 0.75 * w, 0.25 * h, 0.06 * min(w, h)
    d2 = ((x - cx) ** 2 + (y - cy) ** 2) / (r * r)
    out[..., :3] += (1000.0 * np.exp(-d2 * 3.0))[..., None].astype(np.float32)
    out[..., 3] = 1.0
    return out.astype(np.float16)


# ---- reference test patterns (tests/test_imagecompress.cpp:14-126) ------------------------------------

def _pattern_rgb_f(w: int, h: int) -> np.ndarray:
    x, y = _grid(w, h)
    col = np.zeros((h, w, 4), np.float64)
    col[..., 1] = 1.0
    col[..., 3] = 1.0
    red = (((x // 2) & 2) != 0) | (((y // 2) & 2) != 0)
    col[red] = (1, 0, 0, 1)
    blue = (((x // 3) % 3) != 0) & (((y // 3) % 3) != 0)
    col[blue] = (0, 0, 1, 1)
    return col


def _to_u8(col: np.ndarray, nch: int) -> np.ndarray:
    return np.floor(np.clip(col[..., :nch], 0, 1) * 255.0 + 0.5).astype(np.uint8)


def pattern(name: str, w: int, h: int):
    """Returns (array, fmt).  Names: R,G,B (solid, RGB8), RGB (checker/lattice RGB8),
    RGB_Punchthrough (RGBA8), RGBA (alpha = x/width, RGBA8), FloatRGBA (RGBA32F)."""
    x, y = _grid(w, h)
    if name in ("R", "G", "B"):
        col = np.zeros((h, w, 4), np.float64)
        col[..., "RGB".index(name)] = 1.0
        col[..., 3] = 1.0
        return _to_u8(col, 3), FMT_RGB8
    col = _pattern_rgb_f(w, h)
    if name == "RGB":
        return _to_u8(col, 3), FMT_RGB8
    if name == "RGB_Punchthrough":
        nada = (x > w // 2) & (y > h // 2)
        col[nada] = (0, 0, 0, 0)
        return _to_u8(col, 4), FMT_RGBA8
    if name == "RGBA":
        col[..., 3] = (x.astype(np.float32) / np.float32(w)) + 0 * y
        return _to_u8(col, 4), FMT_RGBA8
    if name == "FloatRGBA":
        col[..., 3] = (x.astype(np.float32) / np.float32(w)) + 0 * y
        return col.astype(np.float32), FMT_RGBA32F
    raise ValueError(name)
