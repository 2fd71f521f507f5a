// bc1.cu -- sm_100a kernel for the AMD-Compressonator-compatible BC1 path (bit-exact with the reference).
//
// Replaces the image loop of reference src/amd_bc1_compressor.cpp:44-71 (gather via block_utils.cpp:7-41) and
// Image_CompressAMDBC1Block (src/amd_bcx_helpers.cpp:51-105); the search is bc1_core.cuh.
//
// Mapping: one 4x4 block per PAIR of adjacent threads (the 3-point fit on one, the 4-point fit on the other). The search is a data-dependent fixed-point iteration (axis refit until the error
// stops improving by 0.001) over FP32 sums whose order is part of the result, so the block axis is the parallel
// axis: consecutive threads take consecutive blocks of a block-row, texel rows of a warp are contiguous 512-byte
// runs (128-bit load per thread for RGBA8), each thread stores its 8-byte block as one 64-bit vector.
// Compiled with --fmad=false: the reference's output changes under FP32 contraction.
#include "common.cuh"
#include "kernels.h"
#include "bc1_core.cuh"

namespace b200ic {

namespace {

constexpr int kThreads = 128;

struct Bc1Params {
	SrcImage img;
	uint2 *dst;
	uint64_t n_blocks;
	float alpha_threshold;
	int32_t steps;
	int32_t explicit_alpha; // bc23_colour_kernel: which halves to produce (Bc23Part)
};

// Texels of block (bx, by, slice) as normalised floats (src/block_utils.cpp:7-41)
__device__ __forceinline__ void gather_block(const SrcImage &img, uint64_t block, uint32_t bx, uint32_t by, uint32_t slice, float in[64]) {
	const int fmt = img.format;
	const bool rgba8 = (fmt == B200IC_FMT_RGBA8 || fmt == B200IC_FMT_RGBA8_SRGB);
	if (rgba8 && bx * 4 + 4 <= img.width && ((img.row_pitch | (uintptr_t) img.base | img.slice_pitch) & 15u) == 0) {
#pragma unroll
		for (int r = 0; r < 4; r++) {
			const uint32_t y = min(by * 4 + r, img.height - 1);
			const uint4 v = __ldg(reinterpret_cast<const uint4 *>(img.base + (uint64_t) slice * img.slice_pitch + (uint64_t) y * img.row_pitch) + bx);
			const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
			for (int c = 0; c < 4; c++) {
				in[(r * 4 + c) * 4 + 0] = unorm8(w[c] & 255u);
				in[(r * 4 + c) * 4 + 1] = unorm8((w[c] >> 8) & 255u);
				in[(r * 4 + c) * 4 + 2] = unorm8((w[c] >> 16) & 255u);
				in[(r * 4 + c) * 4 + 3] = unorm8(w[c] >> 24);
			}
		}
	} else {
#pragma unroll 1
		for (int i = 0; i < 16; i++) {
			const float4 f = fetch_rgba(img, block, bx, by, slice, i);
			in[i * 4 + 0] = f.x; in[i * 4 + 1] = f.y; in[i * 4 + 2] = f.z; in[i * 4 + 3] = f.w;
		}
	}
}

__global__ void __launch_bounds__(kThreads, 8) bc1_kernel(const Bc1Params p) {
	// two adjacent lanes per block: the 3-point and the 4-point fit are independent until the final comparison.
	// No early return: the pair meets in a shuffle below, so out-of-range pairs work on a clamped block and only skip
	// the store (kThreads is even: pairs never straddle warps).
	const uint64_t tid = (uint64_t) blockIdx.x * kThreads + threadIdx.x;
	const bool live = (tid >> 1) < p.n_blocks;
	const uint64_t block = live ? (tid >> 1) : p.n_blocks - 1;
	const int which = (int) (tid & 1u);
	const uint64_t per_slice = (uint64_t) p.img.blocks_x * p.img.blocks_y;
	const uint32_t slice = (uint32_t) (block / per_slice);
	const uint32_t rem = (uint32_t) (block - (uint64_t) slice * per_slice);
	const uint32_t by = rem / p.img.blocks_x, bx = rem - by * p.img.blocks_x;

	float in[64];
	gather_block(p.img, block, bx, by, slice, in);
	uint8_t ep[3][2], idx[16];
	const double e = bc1::fit_half(in, which, p.alpha_threshold, p.steps, ep, idx);
	__syncwarp(FULL); // the data-dependent loops of the two fits end at different times: reconverge before the exchange
	const double other = __shfl_xor_sync(FULL, e, 1);
	const double e3 = which ? other : e, e4 = which ? e : other;
	const int m = (e3 <= e4) ? 0 : 1; // (:89) -- an exact 3-point fit (e3 == 0) wins whatever the 4-point error is
	if (live && m == which) {
		uint32_t out[2];
		bc1::pack_fit(m, ep, idx, out);
		p.dst[block] = make_uint2(out[0], out[1]);
	}
}

// ---- BC2 / BC3: 16-byte blocks = 8 bytes of alpha, then 8 bytes of colour -------------------------------
// Colour half = Image_CompressAMDRGBSingleModeBlock (src/amd_bcx_helpers.cpp:142-179).  Its CompRGBBlock
// (src/amd_bcx_body.cpp:1299-1362) is a copy of the BC1 colour fit that indexes the caller's stride-3 float[48] with
// stride 4 and writes 15 floats past its own fBlk[48]: the bytes it produces depend on the caller's stack.  The defined
// behaviour here is the one the code is a copy OF: the BC1 4-point fit without punch-through (CompRGBABlock with
// dwNumPoints = 4), packed like :167-178 (c0 <= c1 swaps the end points).  One block per thread.
__global__ void __launch_bounds__(kThreads) bc23_colour_kernel(const Bc1Params p) {
	const uint64_t block = (uint64_t) blockIdx.x * kThreads + threadIdx.x;
	if (block >= p.n_blocks) return;
	const uint64_t per_slice = (uint64_t) p.img.blocks_x * p.img.blocks_y;
	const uint32_t slice = (uint32_t) (block / per_slice);
	const uint32_t rem = (uint32_t) (block - (uint64_t) slice * per_slice);
	const uint32_t by = rem / p.img.blocks_x, bx = rem - by * p.img.blocks_x;
	float in[64];
	gather_block(p.img, block, bx, by, slice, in);
	const int part = p.explicit_alpha; // kBc3Colour / kBc2Both (16-byte blocks), kColourOnly / kAlphaOnly (block API: 8 bytes)
	if (part != kAlphaOnly) {
		uint8_t ep[3][2], idx[16];
		bc1::compress(in, 4, false, 0.0f, p.steps, ep, idx);
		uint32_t out[2];
		bc1::pack_fit(1, ep, idx, out);
		p.dst[part == kColourOnly ? block : block * 2 + 1] = make_uint2(out[0], out[1]);
	}
	if (part == kBc2Both || part == kAlphaOnly) { // Image_CompressAMDExplictAlphaSingleModeBlock (src/amd_bcx_helpers.cpp:107-123)
		uint32_t a[2] = {0, 0};
		const int ch = p.img.format == B200IC_FMT_BLOCKS_F32X1 ? 0 : 3;
#pragma unroll
		for (int i = 0; i < 16; i++) a[i >> 3] |= bc1::explicit_alpha4(in[i * 4 + ch]) << ((i & 7) * 4);
		p.dst[part == kAlphaOnly ? block : block * 2] = make_uint2(a[0], a[1]);
	}
}

} // namespace

cudaError_t launch_bc1(const SrcImage &img, const b200ic_opts &opts, void *dst, cudaStream_t stream) {
	Bc1Params p;
	p.img = img;
	p.dst = static_cast<uint2 *>(dst);
	p.n_blocks = (uint64_t) img.blocks_x * img.blocks_y * img.slices;
	p.alpha_threshold = opts.bc1_alpha_threshold;
	p.steps = (opts.amd_refinement_steps & 0xff) | (opts.amd_3d_refinement ? bc1::kRefine3D : 0);
	p.explicit_alpha = 0;
	if (p.n_blocks == 0) return cudaSuccess;
	const uint64_t grid = (2 * p.n_blocks + kThreads - 1) / kThreads;
	bc1_kernel<<<(unsigned) grid, kThreads, 0, stream>>>(p);
	return cudaGetLastError();
}

// BC3: alpha half = the BC4 search on channel 3 (Image_CompressAMDAlphaSingleModeBlock, src/amd_bc3_compressor.cpp:40),
// BC2: alpha half = 4-bit rounding; colour half see bc23_colour_kernel.  A source without alpha encodes alpha = 1
// (forceAlphaTo1, src/block_utils.cpp:100-104) -- the gather already returns 1.0 for missing channels.
cudaError_t launch_bc23(const SrcImage &img, const b200ic_opts &opts, int part, void *dst, cudaStream_t stream) {
	Bc1Params p;
	p.img = img;
	p.dst = static_cast<uint2 *>(dst);
	p.n_blocks = (uint64_t) img.blocks_x * img.blocks_y * img.slices;
	p.alpha_threshold = 0.0f;
	p.steps = (opts.amd_refinement_steps & 0xff) | (opts.amd_3d_refinement ? bc1::kRefine3D : 0);
	p.explicit_alpha = part;
	if (p.n_blocks == 0) return cudaSuccess;
	if (part == kBc3Colour) {
		const cudaError_t e = launch_bc45(img, 1, 3, dst, stream, 16);
		if (e != cudaSuccess) return e;
		count_launches(1);
	}
	bc23_colour_kernel<<<(unsigned) ((p.n_blocks + kThreads - 1) / kThreads), kThreads, 0, stream>>>(p);
	return cudaGetLastError();
}

} // namespace b200ic
