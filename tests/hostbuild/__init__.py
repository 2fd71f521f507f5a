"""TEST/DEBUG AID: g++ build of the host-portable per-block encoder cores (csrc/*_core.cuh).

Lets the CPU-only container check the encoder logic against the oracle before GPU time is spent. It is not part of
the product: the package never loads it, and the GPU tests go through the CUDA library only.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_CSRC = os.path.join(_ROOT, "gfx_imagecompress_b200", "csrc")
_LIB = os.path.join(_HERE, "libhostbuild.so")
_DEFS = ["-DHB_BC7RG"]


def _stale():
    if not os.path.exists(_LIB):
        return True
    t = os.path.getmtime(_LIB)
    deps = [os.path.join(_HERE, "hostbuild.cpp")] + [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith(".cuh")]
    return any(os.path.getmtime(d) > t for d in deps)


def load():
    if _stale():
        defs = list(_DEFS)
        for name, d in (("bc1_core.cuh", "-DHB_BC1"), ("bc7amd_core.cuh", "-DHB_BC7AMD"), ("bc6h_core.cuh", "-DHB_BC6H")):
            if os.path.exists(os.path.join(_CSRC, name)):
                defs.append(d)
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-shared", "-pthread"] + defs +
                       ["-I" + _CSRC, "-I" + os.path.join(_ROOT, "include"), "-o", _LIB, os.path.join(_HERE, "hostbuild.cpp")], check=True)
    L = C.CDLL(_LIB)
    L.hb_bc7rg_blocks.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p]
    if hasattr(L, "hb_bc6h_blocks"):
        L.hb_bc6h_blocks.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
    if hasattr(L, "hb_bc1_blocks"):
        L.hb_bc1_blocks.argtypes = [C.c_void_p, C.c_uint64, C.c_float, C.c_int, C.c_void_p]
    if hasattr(L, "hb_bc7amd_blocks"):
        L.hb_bc7amd_blocks.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    if hasattr(L, "hb_bc23_colour_blocks"):
        L.hb_bc23_colour_blocks.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
    if hasattr(L, "hb_cube_lane_check"):
        L.hb_cube_lane_check.argtypes = [C.c_uint64, C.c_int]
    return L


def bc7amd_blocks(L, blocks_f32: np.ndarray, mode_mask: int = 0xFF, threads: int = 8, u8_path: bool = False):
    """blocks_f32: (N, 64) float32 RGBA 0..1 in texel order. Returns (blocks uint8 (N,16), encoder SSE (N,))."""
    b = np.ascontiguousarray(blocks_f32, np.float32).reshape(-1, 64)
    out = np.zeros((len(b), 16), np.uint8)
    err = np.zeros(len(b), np.float64)
    L.hb_bc7amd_blocks(b.ctypes.data, len(b), mode_mask, out.ctypes.data, err.ctypes.data, threads, int(u8_path))
    return out, err


def bc7rg_blocks(L, blocks_u32: np.ndarray, perceptual=True, fast=False) -> np.ndarray:
    b = np.ascontiguousarray(blocks_u32, np.uint32).reshape(-1, 16)
    out = np.zeros((len(b), 16), np.uint8)
    L.hb_bc7rg_blocks(b.ctypes.data, len(b), int(perceptual), int(fast), out.ctypes.data)
    return out


def bc1_blocks(L, blocks_f32: np.ndarray, alpha_threshold: float = 128 / 255.0, steps: int = 1) -> np.ndarray:
    b = np.ascontiguousarray(blocks_f32, np.float32).reshape(-1, 64)
    out = np.zeros((len(b), 8), np.uint8)
    L.hb_bc1_blocks(b.ctypes.data, len(b), alpha_threshold, steps, out.ctypes.data)
    return out


def bc6h_blocks(L, blocks_f32: np.ndarray, is_signed: bool = False) -> np.ndarray:
    b = np.ascontiguousarray(blocks_f32, np.float32).reshape(-1, 64)
    out = np.zeros((len(b), 16), np.uint8)
    L.hb_bc6h_blocks(b.ctypes.data, len(b), int(is_signed), out.ctypes.data)
    return out


def bc23_colour_blocks(L, blocks_f32: np.ndarray, steps: int = 1) -> np.ndarray:
    b = np.ascontiguousarray(blocks_f32, np.float32).reshape(-1, 64)
    out = np.zeros((len(b), 8), np.uint8)
    L.hb_bc23_colour_blocks(b.ctypes.data, len(b), steps, out.ctypes.data)
    return out
