// Compatibility shim for DeanoC `gfx_image` (image container; external to the reference tree).
// Provides the subset of the container API the BCn encode path and its callers touch:
//   reference src/block_utils.cpp:24-26,155-157 (index / pixel fetch / block index / raw pointer),
//   reference src/amd_bc1_compressor.cpp:36 (Image_CreateNoClear), tests/test_imagecompress.cpp (create/destroy).
// Layout + conversions below are OUR definition and are shared by the CPU oracle build and the
// B200 host shim, so both sides of every parity test see identical texel values:
//   * pixel data follows the header contiguously; row-major, slice-major, tightly packed
//   * compressed images keep the caller's width/height rounded UP to a multiple of 4
//     (reference tests expect 257 -> 260, tests/test_imagecompress.cpp:155-156)
//   * u8 UNORM/SRGB channel -> float is `x / 255.0f` (no sRGB curve), missing g,b = 0, a = 1
//   * half -> float is exact
#pragma once
#include "al2o3_platform/platform.h"
#include "al2o3_cmath/scalar.h"
#include "tiny_imageformat/tinyimageformat_base.h"
#include "tiny_imageformat/tinyimageformat_query.h"

typedef struct Image_ImageHeader {
	uint64_t dataSize;
	uint32_t width;
	uint32_t height;
	uint32_t depth;
	uint32_t slices;
	union {
		TinyImageFormat format;
		uint32_t fmtSizer;
	};
	uint16_t flags;
	uint8_t nextType;
	uint8_t pad8;
	union {
		uint64_t pad;
		struct Image_ImageHeader const *nextImage;
	};
	uint64_t pad2;
} Image_ImageHeader; // 48 bytes; texel data starts right after

typedef struct Image_PixelF { float r, g, b, a; } Image_PixelF;
typedef struct Image_PixelD { double r, g, b, a; } Image_PixelD;

#ifdef __cplusplus
extern "C" {
#endif

static inline void *Image_RawDataPtr(Image_ImageHeader const *image) {
	return (void *) (image + 1);
}

static inline uint64_t Image_Shim_ByteCount(uint32_t w, uint32_t h, uint32_t d, uint32_t s, TinyImageFormat fmt) {
	if (TinyImageFormat_IsCompressed(fmt)) {
		uint64_t const bx = (w + 3) / 4, by = (h + 3) / 4;
		return bx * by * d * s * (TinyImageFormat_BitSizeOfBlock(fmt) / 8);
	}
	return (uint64_t) w * h * d * s * TinyImageFormat_BytesPerPixel(fmt);
}

static inline Image_ImageHeader const *Image_CreateNoClear(uint32_t width, uint32_t height, uint32_t depth,
																													 uint32_t slices, TinyImageFormat format) {
	if (TinyImageFormat_IsCompressed(format)) {
		width = (width + 3u) & ~3u;
		height = (height + 3u) & ~3u;
	}
	uint64_t const bytes = Image_Shim_ByteCount(width, height, depth, slices, format);
	if (bytes == 0) return NULL;
	Image_ImageHeader *img = (Image_ImageHeader *) malloc(sizeof(Image_ImageHeader) + bytes);
	if (!img) return NULL;
	memset(img, 0, sizeof(Image_ImageHeader));
	img->dataSize = bytes;
	img->width = width;
	img->height = height;
	img->depth = depth;
	img->slices = slices;
	img->format = format;
	return img;
}
static inline Image_ImageHeader const *Image_Create(uint32_t width, uint32_t height, uint32_t depth,
																										uint32_t slices, TinyImageFormat format) {
	Image_ImageHeader const *img = Image_CreateNoClear(width, height, depth, slices, format);
	if (img) memset(Image_RawDataPtr(img), 0, img->dataSize);
	return img;
}
static inline Image_ImageHeader const *Image_Create2D(uint32_t width, uint32_t height, TinyImageFormat format) {
	return Image_Create(width, height, 1, 1, format);
}
static inline void Image_Destroy(Image_ImageHeader const *image) {
	free((void *) image);
}

static inline size_t Image_CalculateIndex(Image_ImageHeader const *image, uint32_t x, uint32_t y, uint32_t z,
																					uint32_t slice) {
	return ((((size_t) slice * image->depth + z) * image->height + y) * image->width) + x;
}
// block index of the 4x4 block containing texel (x,y); blocks are row-major per slice
static inline size_t Image_GetBlockIndex(Image_ImageHeader const *image, uint32_t x, uint32_t y, uint32_t z,
																				 uint32_t slice) {
	size_t const bx = (image->width + 3) / 4, by = (image->height + 3) / 4;
	return ((((size_t) slice * image->depth + z) * by + (y / 4)) * bx) + (x / 4);
}

static inline void Image_GetPixelAtF(Image_ImageHeader const *image, float *pixel, size_t index) {
	uint8_t const *raw = (uint8_t const *) Image_RawDataPtr(image);
	pixel[0] = 0.0f; pixel[1] = 0.0f; pixel[2] = 0.0f; pixel[3] = 1.0f;
	switch (image->format) {
	case TinyImageFormat_R8_UNORM:
		pixel[0] = raw[index] / 255.0f;
		break;
	case TinyImageFormat_R8G8_UNORM:
		pixel[0] = raw[index * 2 + 0] / 255.0f;
		pixel[1] = raw[index * 2 + 1] / 255.0f;
		break;
	case TinyImageFormat_R8G8B8_UNORM: case TinyImageFormat_R8G8B8_SRGB:
		for (int c = 0; c < 3; ++c) pixel[c] = raw[index * 3 + c] / 255.0f;
		break;
	case TinyImageFormat_R8G8B8A8_UNORM: case TinyImageFormat_R8G8B8A8_SRGB:
		for (int c = 0; c < 4; ++c) pixel[c] = raw[index * 4 + c] / 255.0f;
		break;
	case TinyImageFormat_R16G16B16A16_SFLOAT: case TinyImageFormat_R16G16B16A16_UFLOAT: {
		uint16_t const *h = (uint16_t const *) raw + index * 4;
		for (int c = 0; c < 4; ++c) pixel[c] = Math_Half2Float(h[c]);
		break;
	}
	case TinyImageFormat_R32G32B32A32_SFLOAT: {
		float const *f = (float const *) raw + index * 4;
		for (int c = 0; c < 4; ++c) pixel[c] = f[c];
		break;
	}
	default: break;
	}
}
static inline void Image_GetPixelAtD(Image_ImageHeader const *image, double *pixel, size_t index) {
	float f[4];
	Image_GetPixelAtF(image, f, index);
	for (int c = 0; c < 4; ++c) pixel[c] = (double) f[c];
}

#ifdef __cplusplus
}
#endif
