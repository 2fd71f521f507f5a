// Compatibility shim for `tiny_imageformat`: only the format tags and predicates the BCn encode
// path touches (call sites: reference src/amd_bc*_compressor.cpp:16-39, src/imagecompress.cpp:56-69,
// src/block_utils.cpp:154). Enum values are private to this shim.
#pragma once
#include <stdint.h>
#include <stdbool.h>

// the private unsigned-half tag below exists in THIS shim only; code that must also compile against upstream
// tiny_imageformat (csrc/image_shim.cpp) tests this macro
#define B200IC_COMPAT_HAS_R16G16B16A16_UFLOAT 1

typedef enum TinyImageFormat {
	TinyImageFormat_UNDEFINED = 0,
	TinyImageFormat_R8_UNORM,
	TinyImageFormat_R8_SNORM,
	TinyImageFormat_R8G8_UNORM,
	TinyImageFormat_R8G8_SNORM,
	TinyImageFormat_R8G8B8_UNORM,
	TinyImageFormat_R8G8B8_SRGB,
	TinyImageFormat_R8G8B8A8_UNORM,
	TinyImageFormat_R8G8B8A8_SRGB,
	TinyImageFormat_R16G16B16A16_SFLOAT,
	TinyImageFormat_R32G32B32A32_SFLOAT,
	// private tag (not in upstream tiny_imageformat): half-float data whose values are >= 0, so that
	// Image_CompressAMDBC6H takes its unsigned path (BASELINE.json config 4). See SURVEY.md 8c.
	TinyImageFormat_R16G16B16A16_UFLOAT,

	TinyImageFormat_DXBC1_RGB_UNORM,
	TinyImageFormat_DXBC1_RGB_SRGB,
	TinyImageFormat_DXBC1_RGBA_UNORM,
	TinyImageFormat_DXBC1_RGBA_SRGB,
	TinyImageFormat_DXBC2_UNORM,
	TinyImageFormat_DXBC2_SRGB,
	TinyImageFormat_DXBC3_UNORM,
	TinyImageFormat_DXBC3_SRGB,
	TinyImageFormat_DXBC4_UNORM,
	TinyImageFormat_DXBC4_SNORM,
	TinyImageFormat_DXBC5_UNORM,
	TinyImageFormat_DXBC5_SNORM,
	TinyImageFormat_DXBC6H_UFLOAT,
	TinyImageFormat_DXBC6H_SFLOAT,
	TinyImageFormat_DXBC7_UNORM,
	TinyImageFormat_DXBC7_SRGB,
	TinyImageFormat_Count
} TinyImageFormat;
