"""CPU: the drop-in library loads without a GPU and exports every entry point the two public headers declare
(include/b200ic.h = the thin C-ABI, include/gfx_imagecompress/imagecompress.h = the reference's own API surface,
reference include/gfx_imagecompress/imagecompress.h:57-142).  No compute call is made here; without a CUDA device the
encode entries must FAIL (there is no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header: str, pattern: str):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    return sorted(set(re.findall(pattern, src)))


def test_headers_declare_and_library_exports():
    import gfx_imagecompress_b200 as g
    lib = g.load_library()
    b200 = declared("b200ic.h", r"\b(b200ic_[a-z_0-9]+)\s*\(")
    image = declared(os.path.join("gfx_imagecompress", "imagecompress.h"), r"\b(Image(?:Compress)?_[A-Za-z0-9_]+)\s*\(")
    image = [s for s in image if not s.endswith("Func")]
    assert len(b200) >= 12 and len(image) >= 18, (b200, image)
    missing = [s for s in b200 + image if not hasattr(lib, s)]
    assert not missing, f"declared in include/ but not exported by the library: {missing}"


def test_reference_api_surface_is_complete():
    """Every symbol of the reference's public header exists under the same name."""
    want = ["Image_CompressInit", "Image_CompressDeinit", "ImageCompress_Compress", "ImageCompress_PickCompressionType",
            "Image_CompressAMDBC1", "Image_CompressAMDBC2", "Image_CompressAMDBC3", "Image_CompressAMDBC4",
            "Image_CompressAMDBC5", "Image_CompressAMDBC6H", "Image_CompressAMDBC7", "Image_CompressRichGel999BC7",
            "Image_CompressAMDRGBSingleModeBlock", "Image_CompressAMDAlphaSingleModeBlock",
            "Image_CompressAMDExplictAlphaSingleModeBlock", "Image_CompressAMDBC1Block",
            "Image_CompressAMDMultiModeLDRBlock", "Image_CompressRichGel999BC7enc16"]
    import gfx_imagecompress_b200 as g
    lib = g.load_library()
    assert not [s for s in want if not hasattr(lib, s)]


def test_sizes_and_defaults_without_gpu():
    import gfx_imagecompress_b200 as g
    lib = g.load_library()
    lib.b200ic_block_bytes.restype = C.c_uint32
    lib.b200ic_output_bytes.restype = C.c_uint64
    assert [lib.b200ic_block_bytes(c) for c in (1, 2, 3, 4, 5, 6, 7, 8)] == [8, 16, 16, 8, 16, 16, 16, 16]
    assert lib.b200ic_output_bytes(7, 257, 257, 1) == 65 * 65 * 16
    o = g.Opts.default()
    assert abs(o.bc1_alpha_threshold - 128 / 255.0) < 1e-7 and o.amd_refinement_steps == 1 and o.amd_mode_mask == 0xFF
    assert o.rg_perceptual == 1 and o.rg_fast == 0 and o.bc4_channel == 1


def test_no_cpu_fallback():
    """On a machine without a CUDA device every encode entry fails loudly; with one this test has nothing to check."""
    import gfx_imagecompress_b200 as g
    lib = g.load_library()
    if lib.b200ic_device_count() > 0:
        pytest.skip("CUDA device present")
    px = np.zeros((8, 8, 4), np.uint8)
    with pytest.raises(g.B200Error):
        g.encode_host(g.BC7_RG, px, 7)
    assert g.Image_CompressAMDBC4(g.Image(px, 7)) is None
