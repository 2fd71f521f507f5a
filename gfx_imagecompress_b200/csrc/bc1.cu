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
};

__global__ void __launch_bounds__(kThreads) bc1_kernel(const Bc1Params p) {
	// two adjacent lanes per block: the 3-point and the 4-point fit are independent until the final comparison
	const uint64_t tid = (uint64_t) blockIdx.x * kThreads + threadIdx.x;
	const uint64_t block = tid >> 1;
	const int which = (int) (tid & 1u);
	if (block >= p.n_blocks) return; // (pairs never straddle: kThreads is even)
	const uint64_t per_slice = (uint64_t) p.img.blocks_x * p.img.blocks_y;
	const uint32_t slice = (uint32_t) (block / per_slice);
	const uint32_t rem = (uint32_t) (block - (uint64_t) slice * per_slice);
	const uint32_t by = rem / p.img.blocks_x, bx = rem - by * p.img.blocks_x;

	float in[64];
	const int fmt = p.img.format;
	const bool rgba8 = (fmt == B200IC_FMT_RGBA8 || fmt == B200IC_FMT_RGBA8_SRGB);
	if (rgba8 && bx * 4 + 4 <= p.img.width && ((p.img.row_pitch | (uintptr_t) p.img.base | p.img.slice_pitch) & 15u) == 0) {
#pragma unroll
		for (int r = 0; r < 4; r++) {
			const uint32_t y = min(by * 4 + r, p.img.height - 1);
			const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p.img.base + (uint64_t) slice * p.img.slice_pitch + (uint64_t) y * p.img.row_pitch) + bx);
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
			const float4 f = fetch_rgba(p.img, block, bx, by, slice, i);
			in[i * 4 + 0] = f.x; in[i * 4 + 1] = f.y; in[i * 4 + 2] = f.z; in[i * 4 + 3] = f.w;
		}
	}
	uint8_t ep[3][2], idx[16];
	const double e = bc1::fit_half(in, which, p.alpha_threshold, p.steps, ep, idx);
	const unsigned pm = __activemask();
	const double other = __shfl_xor_sync(pm, e, 1);
	const double e3 = which ? other : e, e4 = which ? e : other;
	const int m = (e3 <= e4) ? 0 : 1; // (:89) -- an exact 3-point fit (e3 == 0) wins whatever the 4-point error is
	if (m == which) {
		uint32_t out[2];
		bc1::pack_fit(m, ep, idx, out);
		p.dst[block] = make_uint2(out[0], out[1]);
	}
}

} // namespace

cudaError_t launch_bc1(const SrcImage &img, const b200ic_opts &opts, void *dst, cudaStream_t stream) {
	Bc1Params p;
	p.img = img;
	p.dst = static_cast<uint2 *>(dst);
	p.n_blocks = (uint64_t) img.blocks_x * img.blocks_y * img.slices;
	p.alpha_threshold = opts.bc1_alpha_threshold;
	p.steps = opts.amd_refinement_steps;
	if (p.n_blocks == 0) return cudaSuccess;
	const uint64_t grid = (2 * p.n_blocks + kThreads - 1) / kThreads;
	bc1_kernel<<<(unsigned) grid, kThreads, 0, stream>>>(p);
	return cudaGetLastError();
}

} // namespace b200ic
