#pragma once
#include "tiny_imageformat/tinyimageformat_base.h"

#ifdef __cplusplus
extern "C" {
#endif

static inline bool TinyImageFormat_IsCompressed(TinyImageFormat f) {
	return f >= TinyImageFormat_DXBC1_RGB_UNORM && f <= TinyImageFormat_DXBC7_SRGB;
}
static inline bool TinyImageFormat_IsSRGB(TinyImageFormat f) {
	switch (f) {
	case TinyImageFormat_R8G8B8_SRGB: case TinyImageFormat_R8G8B8A8_SRGB:
	case TinyImageFormat_DXBC1_RGB_SRGB: case TinyImageFormat_DXBC1_RGBA_SRGB:
	case TinyImageFormat_DXBC2_SRGB: case TinyImageFormat_DXBC3_SRGB: case TinyImageFormat_DXBC7_SRGB: return true;
	default: return false;
	}
}
static inline bool TinyImageFormat_IsSigned(TinyImageFormat f) {
	switch (f) {
	case TinyImageFormat_R8_SNORM: case TinyImageFormat_R8G8_SNORM:
	case TinyImageFormat_R16G16B16A16_SFLOAT: case TinyImageFormat_R32G32B32A32_SFLOAT:
	case TinyImageFormat_DXBC4_SNORM: case TinyImageFormat_DXBC5_SNORM: case TinyImageFormat_DXBC6H_SFLOAT: return true;
	default: return false;
	}
}
static inline bool TinyImageFormat_IsFloat(TinyImageFormat f) {
	switch (f) {
	case TinyImageFormat_R16G16B16A16_SFLOAT: case TinyImageFormat_R32G32B32A32_SFLOAT:
	case TinyImageFormat_R16G16B16A16_UFLOAT:
	case TinyImageFormat_DXBC6H_UFLOAT: case TinyImageFormat_DXBC6H_SFLOAT: return true;
	default: return false;
	}
}
static inline bool TinyImageFormat_IsNormalised(TinyImageFormat f) {
	return f != TinyImageFormat_UNDEFINED && !TinyImageFormat_IsFloat(f);
}
static inline uint32_t TinyImageFormat_ChannelCount(TinyImageFormat f) {
	switch (f) {
	case TinyImageFormat_R8_UNORM: case TinyImageFormat_R8_SNORM:
	case TinyImageFormat_DXBC4_UNORM: case TinyImageFormat_DXBC4_SNORM: return 1;
	case TinyImageFormat_R8G8_UNORM: case TinyImageFormat_R8G8_SNORM:
	case TinyImageFormat_DXBC5_UNORM: case TinyImageFormat_DXBC5_SNORM: return 2;
	case TinyImageFormat_R8G8B8_UNORM: case TinyImageFormat_R8G8B8_SRGB:
	case TinyImageFormat_DXBC1_RGB_UNORM: case TinyImageFormat_DXBC1_RGB_SRGB:
	case TinyImageFormat_DXBC6H_UFLOAT: case TinyImageFormat_DXBC6H_SFLOAT: return 3;
	case TinyImageFormat_UNDEFINED: case TinyImageFormat_Count: return 0;
	default: return 4;
	}
}
// bytes per texel for uncompressed formats
static inline uint32_t TinyImageFormat_BytesPerPixel(TinyImageFormat f) {
	switch (f) {
	case TinyImageFormat_R8_UNORM: case TinyImageFormat_R8_SNORM: return 1;
	case TinyImageFormat_R8G8_UNORM: case TinyImageFormat_R8G8_SNORM: return 2;
	case TinyImageFormat_R8G8B8_UNORM: case TinyImageFormat_R8G8B8_SRGB: return 3;
	case TinyImageFormat_R8G8B8A8_UNORM: case TinyImageFormat_R8G8B8A8_SRGB: return 4;
	case TinyImageFormat_R16G16B16A16_SFLOAT: case TinyImageFormat_R16G16B16A16_UFLOAT: return 8;
	case TinyImageFormat_R32G32B32A32_SFLOAT: return 16;
	default: return 0;
	}
}
static inline uint32_t TinyImageFormat_BitSizeOfBlock(TinyImageFormat f) {
	switch (f) {
	case TinyImageFormat_DXBC1_RGB_UNORM: case TinyImageFormat_DXBC1_RGB_SRGB:
	case TinyImageFormat_DXBC1_RGBA_UNORM: case TinyImageFormat_DXBC1_RGBA_SRGB:
	case TinyImageFormat_DXBC4_UNORM: case TinyImageFormat_DXBC4_SNORM: return 64;
	case TinyImageFormat_DXBC2_UNORM: case TinyImageFormat_DXBC2_SRGB:
	case TinyImageFormat_DXBC3_UNORM: case TinyImageFormat_DXBC3_SRGB:
	case TinyImageFormat_DXBC5_UNORM: case TinyImageFormat_DXBC5_SNORM:
	case TinyImageFormat_DXBC6H_UFLOAT: case TinyImageFormat_DXBC6H_SFLOAT:
	case TinyImageFormat_DXBC7_UNORM: case TinyImageFormat_DXBC7_SRGB: return 128;
	default: return TinyImageFormat_BytesPerPixel(f) * 8;
	}
}
static inline uint32_t TinyImageFormat_WidthOfBlock(TinyImageFormat f) { return TinyImageFormat_IsCompressed(f) ? 4 : 1; }
static inline uint32_t TinyImageFormat_HeightOfBlock(TinyImageFormat f) { return TinyImageFormat_IsCompressed(f) ? 4 : 1; }

#ifdef __cplusplus
}
#endif
