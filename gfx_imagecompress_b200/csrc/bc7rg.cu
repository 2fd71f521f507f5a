// bc7rg.cu -- sm_100a kernel for the bc7enc16-compatible BC7 path (bit-exact with the reference).
//
// Replaces the image loop + per-block call of reference src/richgel999_bc7enc16.cpp:21-97 (gather via
// block_utils.cpp:7-41, re-encode to RGBA8 :52-55, bc7enc16_compress_block :1517-1547). The per-block search
// lives in bc7rg_core.cuh.
//
// Mapping: one 4x4 block per thread. The search is a short chain of dependent least-squares refits whose float
// accumulations must stay in texel order for bit-exactness, so the parallel axis that costs nothing is the
// block axis: consecutive threads take consecutive blocks of a block-row, each texel row of a warp is one
// contiguous 512-byte run (128-bit load per thread), and each thread stores its 16-byte block as one vector.
// Compiled with --fmad=false: no FP32 contraction anywhere (see build.py).
#include "common.cuh"
#include "kernels.h"
#include "bc7rg_core.cuh"

namespace b200ic {

namespace {

constexpr int kThreads = 128;

__device__ rg::OptimalEndpoint g_opt1[512];

struct RgParams {
	SrcImage img;
	uint4 *dst;
	uint64_t n_blocks;
	rg::Params enc;
};

__global__ void __launch_bounds__(kThreads) bc7rg_kernel(const RgParams p) {
	const uint64_t block = (uint64_t) blockIdx.x * kThreads + threadIdx.x;
	if (block >= p.n_blocks) return;
	const uint64_t per_slice = (uint64_t) p.img.blocks_x * p.img.blocks_y;
	const uint32_t slice = (uint32_t) (block / per_slice);
	const uint32_t rem = (uint32_t) (block - (uint64_t) slice * per_slice);
	const uint32_t by = rem / p.img.blocks_x, bx = rem - by * p.img.blocks_x;

	uint32_t px[16];
	const int fmt = p.img.format;
	const bool rgba8 = (fmt == B200IC_FMT_RGBA8 || fmt == B200IC_FMT_RGBA8_SRGB);
	if (rgba8 && bx * 4 + 4 <= p.img.width && ((p.img.row_pitch | (uintptr_t) p.img.base | p.img.slice_pitch) & 15u) == 0) {
#pragma unroll
		for (int r = 0; r < 4; r++) {
			const uint32_t y = min(by * 4 + r, p.img.height - 1);
			const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p.img.base + (uint64_t) slice * p.img.slice_pitch + (uint64_t) y * p.img.row_pitch) + bx);
			px[r * 4 + 0] = v.x; px[r * 4 + 1] = v.y; px[r * 4 + 2] = v.z; px[r * 4 + 3] = v.w;
		}
	} else if (fmt == B200IC_FMT_BLOCKS_RGBA8) {
#pragma unroll
		for (int r = 0; r < 4; r++) {
			const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p.img.base) + block * 4 + r);
			px[r * 4 + 0] = v.x; px[r * 4 + 1] = v.y; px[r * 4 + 2] = v.z; px[r * 4 + 3] = v.w;
		}
	} else {
		// any other source: gather as float exactly like ReadNxNBlockF, then the shim's float -> u8 re-encode
#pragma unroll 1
		for (int i = 0; i < 16; i++) {
			const float4 f = fetch_rgba(p.img, block, bx, by, slice, i);
			px[i] = f2u8(f.x) | (f2u8(f.y) << 8) | (f2u8(f.z) << 16) | (f2u8(f.w) << 24);
		}
	}
	uint64_t out[2];
	rg::encode_block(px, p.enc, out);
	p.dst[block] = make_uint4((uint32_t) out[0], (uint32_t) (out[0] >> 32), (uint32_t) out[1], (uint32_t) (out[1] >> 32));
}

} // namespace

cudaError_t init_bc7rg_tables() {
	static rg::OptimalEndpoint host[512];
	rg::build_mode1_single_colour_table(host);
	return cudaMemcpyToSymbol(g_opt1, host, sizeof(host));
}

cudaError_t launch_bc7rg(const SrcImage &img, const b200ic_opts &opts, void *dst, cudaStream_t stream) {
	RgParams p;
	p.img = img;
	p.dst = static_cast<uint4 *>(dst);
	p.n_blocks = (uint64_t) img.blocks_x * img.blocks_y * img.slices;
	if (p.n_blocks == 0) return cudaSuccess;
	rg::OptimalEndpoint *table = nullptr;
	cudaError_t e = cudaGetSymbolAddress(reinterpret_cast<void **>(&table), g_opt1);
	if (e != cudaSuccess) return e;
	rg::make_params(p.enc, opts.rg_perceptual != 0, opts.rg_fast != 0, table);
	const uint64_t grid = (p.n_blocks + kThreads - 1) / kThreads;
	bc7rg_kernel<<<(unsigned) grid, kThreads, 0, stream>>>(p);
	return cudaGetLastError();
}

} // namespace b200ic
