// common.cuh -- shared device helpers for the sm_100a BCn kernels: source-image descriptor,
// replicate-edge texel gather (replaces reference src/block_utils.cpp:7-41,116-144), warp utilities.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "b200ic.h"

namespace b200ic {

struct SrcImage {
	const uint8_t *base;
	uint64_t row_pitch;   // bytes
	uint64_t slice_pitch; // bytes
	uint32_t width, height, slices;
	uint32_t blocks_x, blocks_y;
	int32_t format;       // b200ic_format
};

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

// u8 -> float exactly as the compat shim's Image_GetPixelAtF: IEEE `x / 255.0f`
__device__ __forceinline__ float unorm8(uint32_t v) { return __fdiv_rn((float) v, 255.0f); }

__device__ __forceinline__ float half_bits_to_float(uint16_t h) {
	return __half2float(__ushort_as_half(h));
}

// One channel of the texel at (x,y) of block-local index `i` in block (bx,by,slice) with replicate-edge clamp
// (src/block_utils.cpp:19,22). Missing channels: g=b=0, a=1 (shim definition).
__device__ __forceinline__ float fetch_channel(const SrcImage &img, uint64_t block_linear, uint32_t bx, uint32_t by,
																							 uint32_t slice, int i, int ch) {
	const int fmt = img.format;
	if (fmt >= 100) { // pre-gathered blocks
		const int nch = fmt - 100 == 7 ? 4 : fmt - 100;
		if (ch >= nch) return ch == 3 ? 1.0f : 0.0f;
		if (fmt == B200IC_FMT_BLOCKS_RGBA8) return unorm8(img.base[(block_linear * 16 + i) * 4 + ch]);
		return reinterpret_cast<const float *>(img.base)[(block_linear * 16 + i) * nch + ch];
	}
	uint32_t x = bx * 4 + (i & 3), y = by * 4 + (i >> 2);
	x = min(x, img.width - 1);
	y = min(y, img.height - 1);
	const uint8_t *row = img.base + (uint64_t) slice * img.slice_pitch + (uint64_t) y * img.row_pitch;
	switch (fmt) {
	case B200IC_FMT_R8: return ch == 0 ? unorm8(__ldg(row + x)) : (ch == 3 ? 1.0f : 0.0f);
	case B200IC_FMT_RG8: return ch < 2 ? unorm8(__ldg(row + x * 2 + ch)) : (ch == 3 ? 1.0f : 0.0f);
	case B200IC_FMT_RGB8:
	case B200IC_FMT_RGB8_SRGB: return ch < 3 ? unorm8(__ldg(row + x * 3 + ch)) : 1.0f;
	case B200IC_FMT_RGBA8:
	case B200IC_FMT_RGBA8_SRGB: return unorm8(__ldg(row + x * 4 + ch));
	case B200IC_FMT_RGBA16F:
	case B200IC_FMT_RGBA16UF: return half_bits_to_float(__ldg(reinterpret_cast<const uint16_t *>(row) + x * 4 + ch));
	case B200IC_FMT_RGBA32F: return __ldg(reinterpret_cast<const float *>(row) + x * 4 + ch);
	default: return 0.0f;
	}
}

// Whole RGBA texel; 8-bit 4-channel sources use one 32-bit load, half sources one 64-bit load.
__device__ __forceinline__ float4 fetch_rgba(const SrcImage &img, uint64_t block_linear, uint32_t bx, uint32_t by,
																						 uint32_t slice, int i) {
	const int fmt = img.format;
	if (fmt < 100) {
		uint32_t x = bx * 4 + (i & 3), y = by * 4 + (i >> 2);
		x = min(x, img.width - 1);
		y = min(y, img.height - 1);
		const uint8_t *row = img.base + (uint64_t) slice * img.slice_pitch + (uint64_t) y * img.row_pitch;
		if (fmt == B200IC_FMT_RGBA8 || fmt == B200IC_FMT_RGBA8_SRGB) {
			const uint32_t p = __ldg(reinterpret_cast<const uint32_t *>(row) + x);
			return make_float4(unorm8(p & 255u), unorm8((p >> 8) & 255u), unorm8((p >> 16) & 255u), unorm8(p >> 24));
		}
		if (fmt == B200IC_FMT_RGBA16F || fmt == B200IC_FMT_RGBA16UF) {
			const uint2 p = __ldg(reinterpret_cast<const uint2 *>(row) + x);
			return make_float4(half_bits_to_float(p.x & 0xffff), half_bits_to_float(p.x >> 16),
												 half_bits_to_float(p.y & 0xffff), half_bits_to_float(p.y >> 16));
		}
		if (fmt == B200IC_FMT_RGBA32F) return __ldg(reinterpret_cast<const float4 *>(row) + x);
	}
	return make_float4(fetch_channel(img, block_linear, bx, by, slice, i, 0),
										 fetch_channel(img, block_linear, bx, by, slice, i, 1),
										 fetch_channel(img, block_linear, bx, by, slice, i, 2),
										 fetch_channel(img, block_linear, bx, by, slice, i, 3));
}

// Raw RGBA8 texel (bc7enc16 consumes bytes). For non-u8 sources the float value is re-encoded exactly as the
// shim's TinyImageFormat_EncodeLogicalPixelsF does: clamp, (uint8)(v*255.0f+0.5f)  (src/richgel999_bc7enc16.cpp:52-55).
__device__ __forceinline__ uint32_t f2u8(float v) {
	if (!(v > 0.0f)) v = 0.0f;
	if (v > 1.0f) v = 1.0f;
	return (uint32_t) __fadd_rn(__fmul_rn(v, 255.0f), 0.5f);
}

// al2o3 Math_MaxF / Math_MinF as defined by the compat shim (plain ternaries; compat/al2o3_cmath/scalar.h)
__device__ __forceinline__ float Math_MaxF_dev(float a, float b) { return a > b ? a : b; }
__device__ __forceinline__ float Math_MinF_dev(float a, float b) { return a < b ? a : b; }

// lexicographic (error, index) arg-min over an aligned power-of-two lane group: the reference keeps the
// FIRST strict minimum in scan order (`err < best`), so ties must go to the lowest scan index.
template <int GROUP>
__device__ __forceinline__ void group_argmin(float &err, int &idx, unsigned mask = FULL) {
#pragma unroll
	for (int d = GROUP / 2; d > 0; d >>= 1) {
		const float e2 = __shfl_xor_sync(mask, err, d);
		const int i2 = __shfl_xor_sync(mask, idx, d);
		if (e2 < err || (e2 == err && i2 < idx)) { err = e2; idx = i2; }
	}
}

} // namespace b200ic
