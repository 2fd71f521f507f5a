"""TEST INFRASTRUCTURE ONLY: decoded-texel quality metrics (PSNR) built on the spec decoders of oracle/bcdec.c."""
from __future__ import annotations

import numpy as np

from .ref import Decoders

_dec = None


def decoders() -> Decoders:
    global _dec
    if _dec is None:
        _dec = Decoders()
    return _dec


def psnr_u8(a: np.ndarray, b: np.ndarray) -> float:
    d = a.astype(np.float64) - b.astype(np.float64)
    mse = float(np.mean(d * d))
    return 99.0 if mse == 0 else 10.0 * np.log10(255.0 * 255.0 / mse)


def psnr_bc7(blocks: np.ndarray, rgba: np.ndarray) -> float:
    """RGBA PSNR of decoded BC7 `blocks` against the (H, W, 4) uint8 source."""
    h, w = rgba.shape[:2]
    return psnr_u8(decoders().bc7(blocks, w, h), rgba)


def psnr_bc1(blocks: np.ndarray, rgba: np.ndarray) -> float:
    h, w = rgba.shape[:2]
    return psnr_u8(decoders().bc1(blocks, w, h)[..., :3], rgba[..., :3])


def psnr_rgb_bc1(colour_blocks: np.ndarray, rgb: np.ndarray) -> float:
    """RGB PSNR of the 8-byte COLOUR halves of BC2 / BC3 blocks (always 4-colour mode) against (H, W, 3) uint8."""
    h, w = rgb.shape[:2]
    full = np.zeros((len(colour_blocks), 16), np.uint8)
    full[:, 8:] = colour_blocks
    return psnr_u8(decoders().bc23(full, w, h, True)[..., :3], rgb[..., :3])


def psnr_bc23(blocks: np.ndarray, rgba: np.ndarray, explicit_alpha: bool) -> float:
    h, w = rgba.shape[:2]
    return psnr_u8(decoders().bc23(blocks, w, h, explicit_alpha), rgba)


def psnr_bc6h(blocks: np.ndarray, half_rgba: np.ndarray, is_signed: bool = False) -> float:
    """RGB PSNR of decoded BC6H blocks against the (H, W, 4) float16 source, peak = the source's largest finite value
    (the stated <= 0.02 dB gate of BASELINE.json compares two encoders on the same image, so the peak cancels)."""
    h, w = half_rgba.shape[:2]
    dec = decoders().bc6h(blocks, w, h, is_signed).astype(np.float64)
    src = half_rgba[..., :3].astype(np.float64)
    mse = float(np.mean((dec - src) ** 2))
    peak = float(np.max(np.abs(src))) or 1.0
    return 99.0 if mse == 0 else 10.0 * np.log10(peak * peak / mse)
