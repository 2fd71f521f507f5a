"""gfx_imagecompress_b200 -- B200 (sm_100a) BCn block-compression engine behind the gfx_imagecompress C API.

The product is the shared library `lib/libgfx_imagecompress_b200.so` (CUDA kernels + C-ABI, built in-tree by
`build.py`).  This package is the thin Python host mirror used by tests / bench: ctypes bindings of the C-ABI
(`include/b200ic.h`) and of the drop-in `Image_Compress*` API (`include/gfx_imagecompress/imagecompress.h`).
There is no CPU fallback: importing works without a GPU, encoding raises.
"""
from .api import (  # noqa: F401
    BC1, BC2, BC3, BC4, BC5, BC6H, BC7_AMD, BC7_RG, BLOCK_BYTES, B200Error, Opts, Image, library, load_library,
    encode_host, encode_device, encode_blocks, encode_batch_device, plan_shards, launch_count, init, codec_available,
    Image_CompressAMDBC1, Image_CompressAMDBC2, Image_CompressAMDBC3, Image_CompressAMDBC4, Image_CompressAMDBC5, Image_CompressAMDBC6H, Image_CompressAMDBC7,
    Image_CompressRichGel999BC7, ImageCompress_Compress,
    Image_CompressAMDAlphaSingleModeBlock, Image_CompressAMDBC1Block, Image_CompressAMDMultiModeLDRBlock,
    Image_CompressRichGel999BC7enc16, Image_CompressAMDRGBSingleModeBlock, Image_CompressAMDExplictAlphaSingleModeBlock,
    ImageCompress_PickCompressionType, device_count, set_devices, box_mip_chain, write_dds, decode_device,
)
from . import synth  # noqa: F401
