"""ctypes host mirror of the C-ABI (include/b200ic.h) and of the reference-facing Image_* API."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import synth

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgfx_imagecompress_b200.so")

BC1, BC2, BC3, BC4, BC5, BC6H, BC7_AMD, BC7_RG = 1, 2, 3, 4, 5, 6, 7, 8
BLOCK_BYTES = {BC1: 8, BC2: 16, BC3: 16, BC4: 8, BC5: 16, BC6H: 16, BC7_AMD: 16, BC7_RG: 16}
FMT_BLOCKS_F32X1, FMT_BLOCKS_F32X3, FMT_BLOCKS_F32X4, FMT_BLOCKS_RGBA8 = 101, 103, 104, 107

# Image_CompressType (include/gfx_imagecompress/imagecompress.h)
Image_CT_None, Image_CT_DXBC1, Image_CT_DXBC2, Image_CT_DXBC3, Image_CT_DXBC4, Image_CT_DXBC5, Image_CT_DXBC6H, \
    Image_CT_DXBC7 = range(8)


class B200Error(RuntimeError):
    pass


class Opts(C.Structure):
    """b200ic_opts"""
    _fields_ = [("bc1_alpha_threshold", C.c_float), ("amd_refinement_steps", C.c_int32),
                ("amd_3d_refinement", C.c_int32), ("amd_adaptive_weights", C.c_int32), ("amd_mode_mask", C.c_int32),
                ("src_has_alpha", C.c_int32), ("rg_perceptual", C.c_int32), ("rg_fast", C.c_int32),
                ("bc6h_signed", C.c_int32), ("bc4_channel", C.c_int32), ("reserved", C.c_int32 * 6)]

    @staticmethod
    def default(**kw) -> "Opts":
        o = Opts()
        library().b200ic_default_opts(C.byref(o))
        for k, v in kw.items():
            setattr(o, k, v)
        return o


class _ImageHeader(C.Structure):
    """Image_ImageHeader of compat/gfx_image/image.h (48 bytes, texels follow)."""
    _fields_ = [("dataSize", C.c_uint64), ("width", C.c_uint32), ("height", C.c_uint32), ("depth", C.c_uint32),
                ("slices", C.c_uint32), ("format", C.c_uint32), ("flags", C.c_uint16), ("nextType", C.c_uint8),
                ("pad8", C.c_uint8), ("pad", C.c_uint64), ("pad2", C.c_uint64)]


assert C.sizeof(_ImageHeader) == 48


class _AmdOptions(C.Structure):
    _fields_ = [("b3DRefinement", C.c_bool), ("AdaptiveColourWeights", C.c_bool), ("RefinementSteps", C.c_uint8),
                ("ModeMask", C.c_uint8)]


class _Bc1Options(C.Structure):
    _fields_ = [("UseAlpha", C.c_bool), ("AlphaThreshold", C.c_uint8)]


class _RgOptions(C.Structure):
    _fields_ = [("perceptual", C.c_bool), ("fast", C.c_bool)]


_PROGRESS = C.CFUNCTYPE(C.c_bool, C.c_void_p, C.c_float)

_lib = None


def load_library(path: str | None = None):
    """Loads the CUDA library. Raises if it has not been built -- there is deliberately no fallback."""
    global _lib
    path = path or os.environ.get("B200IC_LIB") or LIB_PATH  # B200IC_LIB: a debug build (tools/amd_phase_times.py)
    if not os.path.exists(path):
        raise B200Error(f"{path} is missing: run `python -m gfx_imagecompress_b200.build` (or __graft_entry__.build())")
    L = C.CDLL(path)
    L.b200ic_default_opts.argtypes = [C.c_void_p]
    L.b200ic_init.argtypes = [C.c_int]
    L.b200ic_last_error.restype = C.c_char_p
    L.b200ic_output_bytes.restype = C.c_uint64
    L.b200ic_output_bytes.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32]
    L.b200ic_launch_count.restype = C.c_uint64
    L.b200ic_encode_device.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64,
                                       C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.b200ic_encode_host.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.b200ic_encode_host_sharded.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.b200ic_set_devices.argtypes = [C.c_int]
    L.b200ic_device_count.restype = C.c_int
    L.b200ic_box_mip_rgba8_device.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p]
    L.b200ic_decode_device.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
    L.b200ic_write_dds.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    L.b200ic_encode_blocks.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p]
    L.b200ic_plan_shards.restype = C.c_uint64
    L.b200ic_plan_shards.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p,
                                     C.c_uint64]
    L.b200ic_encode_batch_device.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    for name in ("Image_CompressAMDBC1", "Image_CompressAMDBC2", "Image_CompressAMDBC3", "Image_CompressAMDBC4",
                 "Image_CompressAMDBC5", "Image_CompressAMDBC6H", "Image_CompressAMDBC7", "Image_CompressRichGel999BC7",
                 "ImageCompress_Compress"):
        getattr(L, name).restype = C.c_void_p
    L.Image_CompressAMDBC1.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.Image_CompressAMDBC4.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.Image_CompressAMDBC5.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    for name in ("Image_CompressAMDBC2", "Image_CompressAMDBC3", "Image_CompressAMDBC6H", "Image_CompressAMDBC7",
                 "Image_CompressRichGel999BC7"):
        getattr(L, name).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ImageCompress_Compress.argtypes = [C.c_int, C.c_bool, C.c_void_p]
    L.ImageCompress_PickCompressionType.argtypes = [C.c_int, C.c_void_p]
    L.Image_CompressAMDAlphaSingleModeBlock.argtypes = [C.c_void_p, C.c_void_p]
    L.Image_CompressAMDBC1Block.argtypes = [C.c_void_p, C.c_bool, C.c_bool, C.c_uint8, C.c_float, C.c_void_p]
    L.Image_CompressAMDMultiModeLDRBlock.argtypes = [C.c_void_p, C.c_uint8, C.c_bool, C.c_float, C.c_bool, C.c_bool,
                                                     C.c_float, C.c_void_p]
    L.Image_CompressRichGel999BC7enc16.argtypes = [C.c_void_p, C.c_bool, C.c_bool, C.c_void_p]
    L.Image_CompressAMDRGBSingleModeBlock.argtypes = [C.c_void_p, C.c_bool, C.c_bool, C.c_uint8, C.c_void_p]
    L.Image_CompressAMDExplictAlphaSingleModeBlock.argtypes = [C.c_void_p, C.c_void_p]
    _lib = L
    return L


def library():
    return _lib if _lib is not None else load_library()


def _check(rc: int, what: str):
    if rc != 0:
        msg = library().b200ic_last_error().decode() or f"rc={rc}"
        raise B200Error(f"{what}: {msg}")


def init(device: int = 0) -> None:
    _check(library().b200ic_init(device), "b200ic_init")


def codec_available(codec: int) -> bool:
    return bool(library().b200ic_codec_available(codec))


def launch_count() -> int:
    return int(library().b200ic_launch_count())


def _geometry(pixels: np.ndarray, fmt: int):
    assert pixels.flags.c_contiguous
    if pixels.ndim == 4:
        s, h, w = pixels.shape[:3]
    else:
        s = 1
        h, w = pixels.shape[:2]
    assert pixels.nbytes == s * h * w * synth.bytes_per_texel(fmt), "array does not match format"
    return s, h, w


def device_count() -> int:
    return int(library().b200ic_device_count())


def set_devices(n: int) -> None:
    """How many GPUs the Image_Compress* entry points shard a large image over (0 = all visible)."""
    library().b200ic_set_devices(n)


def encode_host(codec: int, pixels: np.ndarray, fmt: int, opts: Opts | None = None, out: np.ndarray | None = None,
                progress=None, devices: int = 1) -> np.ndarray:
    """Host-buffer encode through b200ic_encode_host (H2D + kernels + D2H inside the call); devices != 1 shards the
    block-rows over that many GPUs of this process (b200ic_encode_host_sharded, 0 = all visible).
    pixels: ([S,] H, W, C) C-contiguous array in `fmt`. Returns uint8 (nblocks, blockBytes)."""
    L = library()
    s, h, w = _geometry(pixels, fmt)
    nb = ((w + 3) // 4) * ((h + 3) // 4) * s
    if out is None:
        out = np.empty((nb, BLOCK_BYTES[codec]), np.uint8)
    cb = _HostProgress(progress) if progress is not None else None
    if devices == 1:
        rc = L.b200ic_encode_host(codec, pixels.ctypes.data, fmt, w, h, 0, s, C.byref(opts) if opts is not None else None,
                                  out.ctypes.data, C.cast(cb.fn, C.c_void_p) if cb else None, None)
    else:
        rc = L.b200ic_encode_host_sharded(codec, pixels.ctypes.data, fmt, w, h, 0, s, C.byref(opts) if opts is not None else None,
                                          out.ctypes.data, C.cast(cb.fn, C.c_void_p) if cb else None, None, devices)
    if rc == 1:
        return None  # cancelled
    _check(rc, "b200ic_encode_host")
    return out


class _HostProgress:
    def __init__(self, fn):
        self._py = fn
        self.fn = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_float)(lambda user, pct: 1 if fn(pct) else 0)


def encode_device(codec: int, src, fmt: int, width: int, height: int, slices: int = 1, opts: Opts | None = None,
                  out=None, stream=None, row_pitch: int = 0, slice_pitch: int = 0):
    """Device-resident encode through b200ic_encode_device. `src` / `out` are CUDA torch tensors (plumbing only:
    torch provides the memory and the stream). Asynchronous on torch's current stream unless `stream` is given."""
    import torch
    L = library()
    assert src.is_cuda and src.is_contiguous()
    nb = ((width + 3) // 4) * ((height + 3) // 4) * slices
    if out is None:
        out = torch.empty((nb, BLOCK_BYTES[codec]), dtype=torch.uint8, device=src.device)
    st = stream if stream is not None else torch.cuda.current_stream(src.device).cuda_stream
    with torch.cuda.device(src.device):
        rc = L.b200ic_encode_device(codec, src.data_ptr(), fmt, width, height, row_pitch, slice_pitch, slices,
                                    C.byref(opts) if opts is not None else None, out.data_ptr(), st)
    _check(rc, "b200ic_encode_device")
    return out


def encode_blocks(codec: int, blocks: np.ndarray, fmt: int, opts: Opts | None = None) -> np.ndarray:
    """Batched block API (b200ic_encode_blocks): blocks is (N, 16[, C]) float32 or (N,16) uint32 RGBA8."""
    L = library()
    assert blocks.flags.c_contiguous
    n = blocks.shape[0]
    out = np.empty((n, BLOCK_BYTES[codec]), np.uint8)
    _check(L.b200ic_encode_blocks(codec, blocks.ctypes.data, fmt, n, C.byref(opts) if opts is not None else None,
                                  out.ctypes.data), "b200ic_encode_blocks")
    return out


def box_mip_chain(top, stream=None):
    """Full mip chain of an (H, W, 4) uint8 CUDA tensor down to 1x1 with b200ic_box_mip_rgba8_device."""
    import torch
    L = library()
    assert top.is_cuda and top.is_contiguous() and top.dtype == torch.uint8 and top.shape[2] == 4
    chain = [top]
    st = stream if stream is not None else torch.cuda.current_stream(top.device).cuda_stream
    with torch.cuda.device(top.device):
        while chain[-1].shape[0] > 1 or chain[-1].shape[1] > 1:
            s = chain[-1]
            h, w = int(s.shape[0]), int(s.shape[1])
            d = torch.empty((max(1, h // 2), max(1, w // 2), 4), dtype=torch.uint8, device=top.device)
            _check(L.b200ic_box_mip_rgba8_device(s.data_ptr(), w, h, 0, d.data_ptr(), 0, st), "b200ic_box_mip_rgba8_device")
            chain.append(d)
    return chain


def decode_device(codec: int, blocks, width: int, height: int, slices: int = 1, is_signed: bool = False, stream=None):
    """b200ic_decode_device: `blocks` = CUDA uint8 tensor (nblocks, blockBytes). Returns (slices*height, width, C) texels:
    uint8 RGBA / R / RG, or float16 RGBA for BC6H."""
    import torch
    assert blocks.is_cuda and blocks.is_contiguous()
    ch = {BC4: 1, BC5: 2}.get(codec, 4)
    out = torch.empty((slices * height, width, ch), dtype=torch.float16 if codec == BC6H else torch.uint8, device=blocks.device)
    st = stream if stream is not None else torch.cuda.current_stream(blocks.device).cuda_stream
    with torch.cuda.device(blocks.device):
        _check(library().b200ic_decode_device(codec, blocks.data_ptr(), width, height, slices, int(is_signed), out.data_ptr(), 0, st),
               "b200ic_decode_device")
    return out


def write_dds(path: str, codec: int, width: int, height: int, levels, srgb: bool = False, is_signed: bool = False) -> None:
    """b200ic_write_dds: `levels` = list of (nblocks, blockBytes) uint8 arrays, top level first."""
    arrs = [np.ascontiguousarray(a, np.uint8) for a in levels]
    ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    _check(library().b200ic_write_dds(path.encode(), codec, int(srgb), int(is_signed), width, height, len(arrs), ptrs), "b200ic_write_dds")


class ImageDesc(C.Structure):
    """b200ic_image_desc"""
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("format", C.c_int32), ("width", C.c_uint32),
                ("height", C.c_uint32), ("reserved", C.c_uint32), ("row_pitch_bytes", C.c_uint64)]


class Shard(C.Structure):
    """b200ic_shard: block-rows [row0, row1) of image `image`"""
    _fields_ = [("image", C.c_uint32), ("row0", C.c_uint32), ("row1", C.c_uint32), ("reserved", C.c_uint32)]


def plan_shards(dims, world: int, rank: int, chunk_rows: int = 0):
    """b200ic_plan_shards: dims = [(width, height), ...] -> [(image, row0, row1), ...] of `rank` (host only, no GPU)."""
    L = library()
    n = len(dims)
    w = (C.c_uint32 * max(n, 1))(*[d[0] for d in dims])
    h = (C.c_uint32 * max(n, 1))(*[d[1] for d in dims])
    count = L.b200ic_plan_shards(w, h, n, chunk_rows, world, rank, None, 0)
    out = (Shard * max(count, 1))()
    L.b200ic_plan_shards(w, h, n, chunk_rows, world, rank, out, count)
    return [(out[i].image, out[i].row0, out[i].row1) for i in range(count)]


def encode_batch_device(codec: int, images, fmt: int, outs=None, shards=None, opts: Opts | None = None, stream=None):
    """b200ic_encode_batch_device: `images` = CUDA torch tensors (H, W, C) in `fmt`, one per texture / mip level; `outs` =
    matching (nblocks, blockBytes) uint8 tensors (allocated if None); `shards` = [(image, row0, row1)] or None (whole
    images). Asynchronous on torch's current stream."""
    import torch
    L = library()
    if outs is None:
        outs = [torch.empty((((t.shape[0] + 3) // 4) * ((t.shape[1] + 3) // 4), BLOCK_BYTES[codec]), dtype=torch.uint8,
                            device=t.device) for t in images]
    descs = (ImageDesc * max(len(images), 1))()
    for i, (t, o) in enumerate(zip(images, outs)):
        assert t.is_cuda and t.is_contiguous() and o.is_cuda and o.is_contiguous()
        descs[i] = ImageDesc(t.data_ptr(), o.data_ptr(), fmt, t.shape[1], t.shape[0], 0, 0)
    sh = None
    if shards is not None:
        sh = (Shard * max(len(shards), 1))(*[Shard(i, a, b, 0) for i, a, b in shards])
    dev = images[0].device if images else None
    st = stream if stream is not None else (torch.cuda.current_stream(dev).cuda_stream if dev is not None else None)
    with torch.cuda.device(dev):
        rc = L.b200ic_encode_batch_device(codec, descs, len(images), sh, len(shards) if shards is not None else 0,
                                          C.byref(opts) if opts is not None else None, st)
    _check(rc, "b200ic_encode_batch_device")
    return outs


# ---- the reference-facing image API ----------------------------------------------------------------------

class Image:
    """An Image_ImageHeader + texels in one buffer, as gfx_image lays it out (compat/gfx_image/image.h)."""

    def __init__(self, pixels: np.ndarray, fmt: int):
        s, h, w = _geometry(pixels, fmt)
        self.buf = (C.c_uint8 * (48 + pixels.nbytes))()
        hdr = _ImageHeader.from_buffer(self.buf)
        hdr.dataSize, hdr.width, hdr.height, hdr.depth, hdr.slices, hdr.format = pixels.nbytes, w, h, 1, s, fmt
        C.memmove(C.addressof(self.buf) + 48, pixels.ctypes.data, pixels.nbytes)
        self.header = hdr

    @property
    def ptr(self):
        return C.addressof(self.buf)


class CompressedImage:
    """Result image returned by the C API (owned: freed with libc free, i.e. Image_Destroy of the shim)."""

    def __init__(self, addr: int):
        hdr = _ImageHeader.from_address(addr)
        self.width, self.height, self.depth, self.slices, self.format = hdr.width, hdr.height, hdr.depth, hdr.slices, hdr.format
        self.data = np.ctypeslib.as_array((C.c_uint8 * hdr.dataSize).from_address(addr + 48)).copy()
        C.CDLL(None).free(C.c_void_p(addr))

    def blocks(self, block_bytes: int) -> np.ndarray:
        return self.data.reshape(-1, block_bytes)


def _wrap(addr):
    return CompressedImage(addr) if addr else None


def _cb(progress):
    if progress is None:
        return None, None
    fn = _PROGRESS(lambda user, pct: bool(progress(pct)))
    return fn, C.cast(fn, C.c_void_p)


def Image_CompressAMDBC1(src: Image, amdOptions=None, options=None, progress=None):
    a = _AmdOptions(*amdOptions) if amdOptions else None
    o = _Bc1Options(*options) if options else None
    keep, cb = _cb(progress)
    return _wrap(library().Image_CompressAMDBC1(src.ptr, C.byref(a) if a else None, C.byref(o) if o else None, cb, None))


def Image_CompressAMDBC2(src: Image, amdOptions=None, progress=None):
    a = _AmdOptions(*amdOptions) if amdOptions else None
    keep, cb = _cb(progress)
    return _wrap(library().Image_CompressAMDBC2(src.ptr, C.byref(a) if a else None, cb, None))


def Image_CompressAMDBC3(src: Image, amdOptions=None, progress=None):
    a = _AmdOptions(*amdOptions) if amdOptions else None
    keep, cb = _cb(progress)
    return _wrap(library().Image_CompressAMDBC3(src.ptr, C.byref(a) if a else None, cb, None))


def ImageCompress_PickCompressionType(flags: int, src: Image) -> int:
    return int(library().ImageCompress_PickCompressionType(flags, src.ptr))


def Image_CompressAMDBC4(src: Image, progress=None):
    keep, cb = _cb(progress)
    return _wrap(library().Image_CompressAMDBC4(src.ptr, cb, None))


def Image_CompressAMDBC5(src: Image, progress=None):
    keep, cb = _cb(progress)
    return _wrap(library().Image_CompressAMDBC5(src.ptr, cb, None))


def Image_CompressAMDBC6H(src: Image, amdOptions=None, progress=None):
    a = _AmdOptions(*amdOptions) if amdOptions else None
    keep, cb = _cb(progress)
    return _wrap(library().Image_CompressAMDBC6H(src.ptr, C.byref(a) if a else None, cb, None))


def Image_CompressAMDBC7(src: Image, amdOptions=None, progress=None):
    a = _AmdOptions(*amdOptions) if amdOptions else None
    keep, cb = _cb(progress)
    return _wrap(library().Image_CompressAMDBC7(src.ptr, C.byref(a) if a else None, cb, None))


def Image_CompressRichGel999BC7(src: Image, richOptions=None, progress=None):
    r = _RgOptions(*richOptions) if richOptions else None
    keep, cb = _cb(progress)
    return _wrap(library().Image_CompressRichGel999BC7(src.ptr, C.byref(r) if r else None, cb, None))


def ImageCompress_Compress(type_: int, fast: bool, src: Image):
    addr = library().ImageCompress_Compress(type_, fast, src.ptr)
    if addr == src.ptr:
        return src
    return _wrap(addr)


def Image_CompressAMDAlphaSingleModeBlock(values) -> np.ndarray:
    v = np.ascontiguousarray(values, np.float32).reshape(16)
    out = np.zeros(8, np.uint8)
    library().Image_CompressAMDAlphaSingleModeBlock(v.ctypes.data, out.ctypes.data)
    return out


def Image_CompressAMDBC1Block(rgba, adaptive=False, refine3d=False, steps=1, alpha_threshold=128 / 255.0) -> np.ndarray:
    v = np.ascontiguousarray(rgba, np.float32).reshape(64)
    out = np.zeros(8, np.uint8)
    library().Image_CompressAMDBC1Block(v.ctypes.data, adaptive, refine3d, steps, alpha_threshold, out.ctypes.data)
    return out


def Image_CompressAMDMultiModeLDRBlock(rgba, mode_mask=0xFF, src_has_alpha=True, quality=1.0, colour_restrict=True,
                                       alpha_restrict=True, performance=1.0) -> np.ndarray:
    v = np.ascontiguousarray(rgba, np.float32).reshape(64)
    out = np.zeros(16, np.uint8)
    library().Image_CompressAMDMultiModeLDRBlock(v.ctypes.data, mode_mask, src_has_alpha, quality, colour_restrict,
                                                 alpha_restrict, performance, out.ctypes.data)
    return out


def Image_CompressRichGel999BC7enc16(rgba8, fast=False, perceptual=True) -> np.ndarray:
    v = np.ascontiguousarray(rgba8, np.uint32).reshape(16)
    out = np.zeros(16, np.uint8)
    library().Image_CompressRichGel999BC7enc16(v.ctypes.data, fast, perceptual, out.ctypes.data)
    return out


def Image_CompressAMDRGBSingleModeBlock(rgb, adaptive=False, refine3d=False, steps=1) -> np.ndarray:
    v = np.ascontiguousarray(rgb, np.float32).reshape(48)
    out = np.zeros(8, np.uint8)
    library().Image_CompressAMDRGBSingleModeBlock(v.ctypes.data, adaptive, refine3d, steps, out.ctypes.data)
    return out


def Image_CompressAMDExplictAlphaSingleModeBlock(values) -> np.ndarray:
    v = np.ascontiguousarray(values, np.float32).reshape(16)
    out = np.zeros(8, np.uint8)
    library().Image_CompressAMDExplictAlphaSingleModeBlock(v.ctypes.data, out.ctypes.data)
    return out
