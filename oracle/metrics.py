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
