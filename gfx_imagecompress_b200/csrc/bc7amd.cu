// bc7amd.cu -- sm_100a kernel for the AMD-Compressonator-compatible BC7 path (all eight modes, quality 1).
//
// Replaces the image loop of reference src/amd_bc7_compressor.cpp:25-80 (gather via block_utils.cpp:7-41) and the
// BC7BlockEncoder::CompressBlock tree (src/amd_bc7_body.cpp:1289-1465); the search itself is bc7amd_core.cuh.
//
// Mapping: one warp per 4x4 block, candidates -> lanes.
//   quantise phase : one (partition, subset) of the current mode per lane (16..192 independent optQuantAnD problems)
//   selection      : rank of every partition's error computed in parallel (stable: ties keep partition order),
//                    the 8 lowest are shaken
//   shake phase    : one (attempt, subset) per lane -- ep_shaker_d + ep_shaker_2_d chains are independent
//   dual-index     : one (rotation, index-selection, vector|scalar) per lane
//   winners        : first strict minimum in the reference's scan order, by one lane, then broadcast
// Arithmetic is FP64 like the reference (B200 keeps a full-rate FP64 pipe), which makes the blocks bit-identical
// to the reference's wherever its qsort tie order does not matter.
#include "common.cuh"
#include "kernels.h"
#include "bc7amd_block.cuh"

namespace b200ic {

namespace {

using namespace amd7;

constexpr int kWarps = 4;

__device__ uint32_t *g_sp_table = nullptr;
uint32_t *g_sp_table_host[16] = {};

struct ShakeOut {
	real err;
	uint64_t idx; // 4 bits per subset-local entry
	uint32_t ep[2]; // 4 x 8-bit endpoint codes each
};

struct WarpScratch {
	float in[64];
	BlockInput B;
	real serr[64][3];
	real perr[64];
	int top[8];
	ShakeOut so[24];
	uint64_t blk[2];
	real blk_err;
};

struct AmdParams {
	SrcImage img;
	uint4 *dst;
	uint64_t n_blocks;
	const uint32_t *sp;
	uint32_t mode_mask;
};

__device__ __forceinline__ uint64_t pack_idx(const int *idx, int n) {
	uint64_t v = 0;
	for (int i = 0; i < n; i++) v |= (uint64_t) (idx[i] & 15) << (4 * i);
	return v;
}
__device__ __forceinline__ uint32_t pack_ep(const int e[4]) {
	return (uint32_t) (e[0] & 255) | ((uint32_t) (e[1] & 255) << 8) | ((uint32_t) (e[2] & 255) << 16) | ((uint32_t) (e[3] & 255) << 24);
}

template <bool U8>
__global__ void __launch_bounds__(kWarps * 32) bc7amd_kernel(const AmdParams p) {
	__shared__ WarpScratch scratch[kWarps];
	const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
	const uint64_t block = (uint64_t) blockIdx.x * kWarps + warp;
	if (block >= p.n_blocks) return; // whole warp
	WarpScratch &ws = scratch[warp];
	const Tables T{p.sp};

	const uint64_t per_slice = (uint64_t) p.img.blocks_x * p.img.blocks_y;
	const uint32_t slice = (uint32_t) (block / per_slice);
	const uint32_t rem = (uint32_t) (block - (uint64_t) slice * per_slice);
	const uint32_t by = rem / p.img.blocks_x, bx = rem - by * p.img.blocks_x;
	if (lane < 16) {
		const float4 t = fetch_rgba(p.img, block, bx, by, slice, (int) lane);
		ws.in[lane * 4 + 0] = t.x; ws.in[lane * 4 + 1] = t.y; ws.in[lane * 4 + 2] = t.z; ws.in[lane * 4 + 3] = t.w;
	}
	__syncwarp();
	if (lane == 0) prepare_block(ws.in, p.mode_mask, ws.B);
	__syncwarp();
	const uint32_t mask = ws.B.mode_mask;

	real best = A7_HUGE;
	uint64_t out0 = 0, out1 = 0;
	for (int vi = 0; vi < 8; vi++) {
		const int mode = mode_visit_order(vi);
		if (!(mask & (1u << mode))) continue;
		const ModeInfo mi = mode_info(mode);
		if (mi.alpha != 2) {
			const ShakeParams sp = single_index_shake_params(mode);
			const int nparts = 1 << mi.partition_bits, subsets = mi.subsets;
			for (int t = (int) lane; t < nparts * subsets; t += 32) {
				const int part = t / subsets, s = t - part * subsets;
				real sub[kMaxEntries][4];
				int n, idx[kMaxEntries];
				gather_subset(ws.B, subsets, part, s, sp.dim, sub, n);
				ws.serr[part][s] = n ? quantise_subset(sub, n, sp.clusters, idx, sp.dim) : 0;
			}
			__syncwarp();
			for (int part = (int) lane; part < nparts; part += 32) {
				real e = 0;
				for (int s = 0; s < subsets; s++) e += ws.serr[part][s];
				ws.perr[part] = e;
			}
			__syncwarp();
			const int attempts = nparts < 8 ? nparts : 8;
			for (int part = (int) lane; part < nparts; part += 32) {
				const real e = ws.perr[part];
				int rank = 0;
				for (int q = 0; q < nparts; q++) {
					const real eq = ws.perr[q];
					rank += ((e - eq > 0) || (!(eq - e > 0) && q < part)) ? 1 : 0;
				}
				if (rank < attempts) ws.top[rank] = part;
			}
			__syncwarp();
			if ((int) lane < attempts * subsets) {
				const int a = (int) lane / subsets, s = (int) lane - a * subsets;
				const int part = ws.top[a];
				real sub[kMaxEntries][4];
				int n, idx[kMaxEntries], ep[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
				gather_subset(ws.B, subsets, part, s, sp.dim, sub, n);
				ShakeOut o;
				o.err = 0; o.idx = 0; o.ep[0] = o.ep[1] = 0;
				if (n) {
					quantise_subset(sub, n, sp.clusters, idx, sp.dim);
					if (U8) {
						U8Subset S;
						make_u8_subset(sub, n, sp.dim, S);
						o.err = shake_subset_u8(T, sp, S, idx, ep);
					} else {
						o.err = shake_subset(T, sp, sub, n, idx, ep);
					}
					o.idx = pack_idx(idx, n);
					o.ep[0] = pack_ep(ep[0]);
					o.ep[1] = pack_ep(ep[1]);
				}
				ws.so[lane] = o;
			}
			__syncwarp();
			if (lane == 0) {
				real be = A7_HUGE;
				int ba = 0;
				for (int a = 0; a < attempts; a++) {
					real e = 0;
					for (int s = 0; s < subsets; s++) e += ws.so[a * subsets + s].err;
					if (e < be) { be = e; ba = a; }
				}
				SingleIndexResult r;
				r.partition = ws.top[ba];
				for (int s = 0; s < subsets; s++) {
					const ShakeOut &o = ws.so[ba * subsets + s];
					for (int k = 0; k < 4; k++) {
						r.ep[s][0][k] = (int) ((o.ep[0] >> (8 * k)) & 255u);
						r.ep[s][1][k] = (int) ((o.ep[1] >> (8 * k)) & 255u);
					}
					for (int i = 0; i < 16; i++) r.idx[s][i] = (int) ((o.idx >> (4 * i)) & 15u);
				}
				uint64_t blk[2];
				pack_single_index(mode, r, blk);
				ws.blk[0] = blk[0];
				ws.blk[1] = blk[1];
				ws.blk_err = be;
			}
			__syncwarp();
		} else {
			const int nrot = 1 << mi.rotation_bits, nsel = 1 << mi.index_mode_bits;
			const int combos = nrot * nsel;
			if ((int) lane < combos * 2) {
				const int combo = (int) lane >> 1, which = (int) lane & 1;
				const int rot = combo / nsel, isel = combo - rot * nsel;
				real blkv[16][4];
				for (int i = 0; i < 16; i++) {
					if (which == 0) {
						blkv[i][0] = ws.B.px[i][rotation_channel(rot, 1)];
						blkv[i][1] = ws.B.px[i][rotation_channel(rot, 2)];
						blkv[i][2] = ws.B.px[i][rotation_channel(rot, 3)];
					} else {
						blkv[i][0] = blkv[i][1] = blkv[i][2] = ws.B.px[i][rotation_channel(rot, 0)];
					}
					blkv[i][3] = 0;
				}
				const int ib = which == 0 ? (isel ? mi.index_bits1 : mi.index_bits0) : (isel ? mi.index_bits0 : mi.index_bits1);
				const int cb = which == 0 ? mi.vector_bits / 3 : mi.scalar_bits;
				const int bits[4] = {cb, cb, cb, 6 * cb};
				int idx[16], ep[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
				quantise_subset(blkv, 16, 1 << ib, idx, 3);
				ShakeOut o;
				if (U8) {
					U8Subset S;
					make_u8_subset(blkv, 16, 3, S);
					shake_cube_u8_any(T, S, idx, ib, bits, CART);
					o.err = shake_window_u8_any(T, S, idx, ep, 6, ib, bits[3], 3);
				} else {
					shake_cube(T, blkv, 16, idx, (1 << ib) - 1, bits, CART);
					o.err = shake_window(T, blkv, 16, idx, ep, 6, (1 << ib) - 1, bits[3], 3);
				}
				o.idx = pack_idx(idx, 16);
				o.ep[0] = pack_ep(ep[0]);
				o.ep[1] = pack_ep(ep[1]);
				ws.so[lane] = o;
			}
			__syncwarp();
			if (lane == 0) {
				real be = A7_HUGE;
				int bc = 0;
				for (int c = 0; c < combos; c++) {
					real e = 0;
					e += ws.so[2 * c].err;
					e += ws.so[2 * c + 1].err / 3.;
					if (e < be) { be = e; bc = c; }
				}
				int ep[2][2][4], idx[2][16];
				for (int w = 0; w < 2; w++) {
					const ShakeOut &o = ws.so[2 * bc + w];
					for (int k = 0; k < 4; k++) {
						ep[w][0][k] = (int) ((o.ep[0] >> (8 * k)) & 255u);
						ep[w][1][k] = (int) ((o.ep[1] >> (8 * k)) & 255u);
					}
					for (int i = 0; i < 16; i++) idx[w][i] = (int) ((o.idx >> (4 * i)) & 15u);
				}
				uint64_t blk[2];
				pack_dual_index(mode, bc % nsel, bc / nsel, ep, idx, blk);
				ws.blk[0] = blk[0];
				ws.blk[1] = blk[1];
				ws.blk_err = be;
			}
			__syncwarp();
		}
		const real e = ws.blk_err;
		if (e < best) {
			best = e;
			out0 = ws.blk[0];
			out1 = ws.blk[1];
		}
		__syncwarp();
	}
	if (lane == 0) p.dst[block] = make_uint4((uint32_t) out0, (uint32_t) (out0 >> 32), (uint32_t) out1, (uint32_t) (out1 >> 32));
}

} // namespace

cudaError_t init_bc7amd_tables() {
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess) return e;
	if (dev < 0 || dev >= 16) return cudaErrorInvalidDevice;
	if (g_sp_table_host[dev]) return cudaSuccess;
	static uint32_t *host = nullptr;
	if (!host) {
		host = new uint32_t[kSpEntries];
		build_single_point_table(host);
	}
	uint32_t *d = nullptr;
	e = cudaMalloc(&d, kSpEntries * sizeof(uint32_t));
	if (e != cudaSuccess) return e;
	e = cudaMemcpy(d, host, kSpEntries * sizeof(uint32_t), cudaMemcpyHostToDevice);
	if (e != cudaSuccess) return e;
	g_sp_table_host[dev] = d;
	return cudaSuccess;
}

cudaError_t launch_bc7amd(const SrcImage &img, const b200ic_opts &opts, void *dst, cudaStream_t stream) {
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess) return e;
	if (dev < 0 || dev >= 16 || !g_sp_table_host[dev]) return cudaErrorInitializationError;
	AmdParams p;
	p.img = img;
	p.dst = static_cast<uint4 *>(dst);
	p.n_blocks = (uint64_t) img.blocks_x * img.blocks_y * img.slices;
	p.sp = g_sp_table_host[dev];
	p.mode_mask = (uint32_t) opts.amd_mode_mask & 0xffu;
	if (p.n_blocks == 0) return cudaSuccess;
	const uint64_t grid = (p.n_blocks + kWarps - 1) / kWarps;
	// 8-bit sources: every component is an exact integer, the exact INT32 shakers apply (bc7amd_int.cuh)
	const bool u8 = img.format == B200IC_FMT_R8 || img.format == B200IC_FMT_RG8 || img.format == B200IC_FMT_RGB8 ||
									img.format == B200IC_FMT_RGB8_SRGB || img.format == B200IC_FMT_RGBA8 || img.format == B200IC_FMT_RGBA8_SRGB ||
									img.format == B200IC_FMT_BLOCKS_RGBA8;
	if (u8) bc7amd_kernel<true><<<(unsigned) grid, kWarps * 32, 0, stream>>>(p);
	else bc7amd_kernel<false><<<(unsigned) grid, kWarps * 32, 0, stream>>>(p);
	return cudaGetLastError();
}

} // namespace b200ic
