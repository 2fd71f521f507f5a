"""ctypes bindings for the oracle libraries (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REFDIR = os.path.join(_HERE, "_ref")

BC1, BC2, BC3, BC4, BC5, BC6H, BC7, BC7_RG = 1, 2, 3, 4, 5, 6, 7, 8
BLOCK_BYTES = {BC1: 8, BC2: 16, BC3: 16, BC4: 8, BC5: 16, BC6H: 16, BC7: 16, BC7_RG: 16}


def build(verbose: bool = False) -> None:
    """Builds whatever can be built: restatements + decoders always; the reference .so only when
    /root/reference is present (it is not on the GPU box, which uses the prebuilt file)."""
    targets = ["restate", "decoders"]
    if os.path.isdir("/root/reference/src"):
        targets.append("ref")
    subprocess.run(["make", "-s", "-j8", "-C", _HERE] + targets, check=True,
                   stdout=None if verbose else subprocess.DEVNULL)


def have_ref() -> bool:
    return os.path.exists(os.path.join(_REFDIR, "libref_oracle.so"))


def have_restated() -> bool:
    return os.path.exists(os.path.join(_REFDIR, "librestate.so"))


class RefOpts(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "bc1_use_alpha", "bc1_alpha_threshold", "amd_3d_refinement", "amd_adaptive_weights",
        "amd_refinement_steps", "amd_mode_mask", "rg_perceptual", "rg_fast", "use_defaults")]


def default_opts(**kw) -> RefOpts:
    o = RefOpts(0, 128, 0, 0, 1, 0xFF, 1, 0, 1)
    for k, v in kw.items():
        setattr(o, k, int(v))
        o.use_defaults = 0
    return o


def _nblocks(w, h):
    return ((w + 3) // 4) * ((h + 3) // 4)


class RefOracle:
    """The compiled, unmodified reference (image-level API + block API)."""

    def __init__(self):
        self.lib = C.CDLL(os.path.join(_REFDIR, "libref_oracle.so"))
        L = self.lib
        L.ref_encode_rows.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32,
                                      C.c_void_p, C.c_int, C.c_void_p]
        L.ref_encode_rows.restype = C.c_int
        L.ref_hw_threads.restype = C.c_int
        L.Image_CompressAMDAlphaSingleModeBlock.argtypes = [C.c_void_p, C.c_void_p]
        L.Image_CompressAMDBC1Block.argtypes = [C.c_void_p, C.c_bool, C.c_bool, C.c_uint8, C.c_float, C.c_void_p]
        L.Image_CompressAMDMultiModeLDRBlock.argtypes = [C.c_void_p, C.c_uint8, C.c_bool, C.c_float, C.c_bool,
                                                         C.c_bool, C.c_float, C.c_void_p]
        L.Image_CompressRichGel999BC7enc16.argtypes = [C.c_void_p, C.c_bool, C.c_bool, C.c_void_p]
        L.ref_bc6h_block.argtypes = [C.c_void_p, C.c_int, C.c_uint8, C.c_void_p]
        L.Image_CompressInit()

    def hw_threads(self) -> int:
        return int(self.lib.ref_hw_threads())

    def encode(self, codec: int, pixels: np.ndarray, fmt: int, threads: int = 0, opts: RefOpts | None = None,
               rows: tuple[int, int] | None = None) -> np.ndarray:
        """pixels: (H, W, C) contiguous array in format `fmt`. Returns uint8 (nblocks, blockBytes)."""
        assert pixels.flags.c_contiguous
        h, w = pixels.shape[:2]
        by0, by1 = rows if rows is not None else (0, (h + 3) // 4)
        bx = (w + 3) // 4
        out = np.zeros(((by1 - by0) * bx, BLOCK_BYTES[codec]), np.uint8)
        if threads <= 0:
            threads = self.hw_threads()
        rc = self.lib.ref_encode_rows(codec, pixels.ctypes.data, w, h, fmt, by0, by1, out.ctypes.data, threads,
                                      C.byref(opts) if opts is not None else None)
        if rc != 0:
            raise RuntimeError(f"reference encode failed rc={rc}")
        return out

    def alpha_block(self, vals: np.ndarray) -> np.ndarray:
        v = np.ascontiguousarray(vals, np.float32)
        out = np.zeros(8, np.uint8)
        self.lib.Image_CompressAMDAlphaSingleModeBlock(v.ctypes.data, out.ctypes.data)
        return out

    def bc1_block(self, rgba: np.ndarray, alpha_threshold: float = 128 / 255.0, steps: int = 1) -> np.ndarray:
        v = np.ascontiguousarray(rgba, np.float32)
        out = np.zeros(8, np.uint8)
        self.lib.Image_CompressAMDBC1Block(v.ctypes.data, False, False, steps, alpha_threshold, out.ctypes.data)
        return out


class Restated:
    """Plain-C restatements (oracle/restate_*.c)."""

    def __init__(self):
        self.lib = C.CDLL(os.path.join(_REFDIR, "librestate.so"))
        L = self.lib
        L.restate_alpha_block.argtypes = [C.c_void_p, C.c_void_p]
        for n in ("restate_bc4_image", "restate_bc5_image"):
            getattr(L, n).argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]

    def alpha_block(self, vals: np.ndarray) -> np.ndarray:
        v = np.ascontiguousarray(vals, np.float32)
        out = np.zeros(8, np.uint8)
        self.lib.restate_alpha_block(v.ctypes.data, out.ctypes.data)
        return out

    def _img(self, fn, pixels, bb):
        assert pixels.dtype == np.uint8 and pixels.flags.c_contiguous
        h, w = pixels.shape[:2]
        nch = pixels.shape[2] if pixels.ndim == 3 else 1
        out = np.zeros((_nblocks(w, h), bb), np.uint8)
        getattr(self.lib, fn)(pixels.ctypes.data, w, h, nch, out.ctypes.data)
        return out

    def bc4(self, pixels):
        return self._img("restate_bc4_image", pixels, 8)

    def bc5(self, pixels):
        return self._img("restate_bc5_image", pixels, 16)


class Decoders:
    """Spec decoders (oracle/bcdec.c): blocks -> texels, for PSNR."""

    def __init__(self):
        self.lib = C.CDLL(os.path.join(_REFDIR, "libbcdec.so"))
        L = self.lib
        L.bcdec_bc1.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.bcdec_bc7.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.bcdec_bc45.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
        L.bcdec_bc23.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
        L.bcdec_bc6h.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]

    def bc1(self, blocks: np.ndarray, w: int, h: int) -> np.ndarray:
        b = np.ascontiguousarray(blocks, np.uint8)
        out = np.zeros((h, w, 4), np.uint8)
        self.lib.bcdec_bc1(b.ctypes.data, w, h, out.ctypes.data)
        return out

    def bc7(self, blocks: np.ndarray, w: int, h: int, with_modes: bool = False):
        b = np.ascontiguousarray(blocks, np.uint8)
        out = np.zeros((h, w, 4), np.uint8)
        hist = np.zeros(9, np.uint32)
        self.lib.bcdec_bc7(b.ctypes.data, w, h, out.ctypes.data, hist.ctypes.data)
        return (out, hist) if with_modes else out

    def bc45(self, blocks: np.ndarray, w: int, h: int, nch: int) -> np.ndarray:
        b = np.ascontiguousarray(blocks, np.uint8)
        out = np.zeros((h, w, nch), np.uint8)
        self.lib.bcdec_bc45(b.ctypes.data, w, h, nch, out.ctypes.data)
        return out

    def bc23(self, blocks: np.ndarray, w: int, h: int, explicit_alpha: bool) -> np.ndarray:
        """BC2 (explicit_alpha) / BC3 blocks -> (h, w, 4) uint8."""
        b = np.ascontiguousarray(blocks, np.uint8)
        out = np.zeros((h, w, 4), np.uint8)
        self.lib.bcdec_bc23(b.ctypes.data, w, h, int(explicit_alpha), out.ctypes.data)
        return out

    def bc6h(self, blocks: np.ndarray, w: int, h: int, is_signed: bool = False, with_modes: bool = False):
        """BC6H blocks -> (h, w, 3) float16."""
        b = np.ascontiguousarray(blocks, np.uint8)
        out = np.zeros((h, w, 3), np.uint16)
        hist = np.zeros(15, np.uint32)
        self.lib.bcdec_bc6h(b.ctypes.data, w, h, int(is_signed), out.ctypes.data, hist.ctypes.data)
        return (out.view(np.float16), hist) if with_modes else out.view(np.float16)
