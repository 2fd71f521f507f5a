// bc7amd.cu -- sm_100a kernels for the AMD-Compressonator-compatible BC7 path (all eight modes, quality 1).
//
// Replaces the image loop of reference src/amd_bc7_compressor.cpp:25-80 (gather via block_utils.cpp:7-41) and the
// BC7BlockEncoder::CompressBlock tree (src/amd_bc7_body.cpp:1289-1465); the search itself is bc7amd_core.cuh /
// bc7amd_int.cuh.
//
// One pass per mode in the reference's visiting order {6,4,3,1,2,0,7,5} (src/amd_bc7_body.cpp:1400); the running best
// block lives in dst and its error in a per-block scratch word, a later mode replaces the block only on a strictly lower
// error (= the reference's first strict minimum).  8-bit sources (every component an exact integer) take the phase
// kernels below; float sources the generic FP64 kernel at the end of the file.
//
// Modes 0 .. 5 and 7 -- THREE kernels per mode (two for mode 7, which has no cube walk), state handed over through HBM
// (<= 584 B per block: traffic that is 10^-3 of the ALU time, and every phase gets its own register budget / occupancy):
//   quantise : lane = one optQuantAnD problem (FP64): a (partition, subset) of the block -- 48 .. 192 per block, repeated
//              subsets of the 3-subset tables once -- or a (rotation, index selection, vector | scalar) of 2 / 4 blocks
//              (modes 4 / 5); ranking of the partitions, the 8 best are kept          -> q_top, q_idx
//   cube     : ep_shaker_d (82 % of the reference's time).  Persistent warps take the next block -- the next 2 / 4 blocks
//              in the modes with 16 / 8 tasks per block -- from the launch's work counter.  A work item = (task, (q,p)
//              re-indexing); set-up (cluster statistics, least-squares fit, lattice floors) one item per lane, then the
//              items one after the other with lane = (lattice, corner): all 32 lanes share the texels, the trip count and
//              the ramp tables (shared memory) -- no divergence in the 3-instruction inner loop (VABSDIFF4, IDP.4A,
//              VIMNMX3); winners by redux.sync min of (error, scan position) keys = the reference's first strict minimum;
//              mode 0 prunes its second pass with per-channel bounds               -> c_idx, c_err
//   window   : ep_shaker_2_d chains (fit: lane = item; search: lane = (item, channel, parity combination)), best attempt,
//              bit packing, compare-and-store
// The kernels are instantiated per index width (and pruning): the code a mode never runs is compiled out (instruction cache).
// Mode 6 (ONE task per block, 16 index levels): thread per block, serial form.
#include "common.cuh"
#include "kernels.h"
#include "bc7amd_block.cuh"
#include <mutex>
#include <type_traits>
#include <vector>

namespace b200ic {

namespace {

using namespace amd7;

constexpr int kWarps = 4;
constexpr int kMaxTasks = 24;   // tasks per block -- single-index: 8 attempts x 3 subsets; dual-index: 8 combos x 2
constexpr int kWarpTasks = 32;  // tasks per warp: a warp of the cube / window kernels takes 2 (mode 4) or 4 (mode 5) blocks at a time
constexpr uint32_t kChunkBlocks = 1u << 19; // blocks per pass over the phase kernels (bounds the scratch: 584 B per block)

uint32_t *g_sp_table_host[16] = {};
int g_serial_modes = 0x40; // modes run by the thread-per-block kernel (debug builds: b200ic_amd_serial_modes)

// Quantise-phase task order: (partition, subset) pairs of each partition table sorted by subset size (descending),
// so that the 32 lanes of one round run optQuantAnD problems of (nearly) the same size. Pure scheduling data.
__constant__ uint8_t c_qorder[48 + 192 + 128];
__device__ __forceinline__ const uint8_t *quantise_order(int subsets, int nparts) {
	return subsets == 3 ? (nparts == 16 ? c_qorder : c_qorder + 48) : c_qorder + 240;
}
// The 3-subset tables repeat subsets (36 distinct texel masks among the 48 of mode 0, 140 among the 192 of mode 2) and the
// quantiser is a pure function of the texel list: c_qfirst[order position] = the first position (same table, same order)
// with that mask; only those are quantised, the repeats copy.  The 2-subset masks are all distinct.
__constant__ uint8_t c_qfirst[48 + 192];

// Optional phase timing (build with -DB200IC_AMD_TIMING, read with b200ic_amd_timing): clock64 deltas of lane 0 summed
// per phase over all warps. Slots: 0 quantise, 1 rank, 2 cube, 3 window, 4 window (2nd), 5 pick+pack,
// 6 cube items, 7 cube item batches, 8 window items, 9 window rounds, 10 cube passes
#ifdef B200IC_AMD_TIMING
__device__ unsigned long long g_amd_timing[16];
#define AMD_T0() long long t__ = clock64()
#define AMD_T(slot)                                                         \
	do {                                                                      \
		const long long n__ = clock64();                                        \
		if (lane == 0) atomicAdd(&g_amd_timing[slot], (unsigned long long) (n__ - t__)); \
		t__ = n__;                                                              \
	} while (0)
#define AMD_COUNT(slot, v)                                                  \
	do {                                                                      \
		if (lane == 0) atomicAdd(&g_amd_timing[slot], (unsigned long long) (v)); \
	} while (0)
#else
#define AMD_T0()
#define AMD_T(slot)
#define AMD_COUNT(slot, v)
#endif

struct ShakeOut {
	real err;
	uint64_t idx; // 4 bits per subset-local entry
	uint32_t ep[2]; // 4 x 8-bit endpoint codes each
};

// One shake problem of the u8 path: a subset of a partition (single-index modes) or a (rotation, index selection,
// vector | scalar) part of a dual-index mode
struct Task {
	uint32_t d[16];     // packed texels
	uint32_t plane[16]; // the same, channel-planar (window_planes_u8; filled by the window kernel only)
	uint64_t idx_q;     // quantiser indices
	uint64_t cur;       // collapsed indices of the running pass
	uint64_t best_idx;  // index_io of the reference's ep_shaker_d
	uint64_t pass_key;  // best (err << 16 | qp << 8 | lattice << 6 | gray) of the running pass
	uint64_t pass_idx;
	real err_o;
	// ep_shaker_2_d state
	uint64_t w_index;   // running indices (uncollapsed)
	uint64_t w_best_idx, w_best_ep;
	real w_err_o;
	uint16_t item_base, item_count;
	uint8_t n, clog, bits, type, Mi, done, all_same, dim;
	uint8_t w_bits_total, w_size, w_tries, w_active;
};

// HBM hand-over between the phase kernels, indexed by the block's position in the running chunk
struct AmdScratch {
	uint64_t *q_idx; // [block][24] quantiser indices of task (attempt, subset)
	uint8_t *q_top;  // [block][8]  partitions to shake, best first; q_top[block][0] == 0xff: mode not searched for this block
	uint64_t *c_idx; // [block][24] ep_shaker_d's indices per task
	real *c_err;     // [block][24] ep_shaker_d's error per task
};

struct AmdParams {
	SrcImage img;
	uint4 *dst;
	uint64_t block0;       // first block of the chunk
	uint32_t n_blocks;     // blocks in the chunk
	const uint32_t *sp;
	uint32_t mode_mask;    // the caller's ModeMask (input of the reference's mode filter)
	uint32_t launch_modes; // modes searched by THIS pass
	int mode;              // phase kernels: the mode of this pass
	real *best_err;        // per block of the image: error of the block currently in dst (carried from pass to pass)
	int first;             // first pass of the sequence: nothing to compare with
	uint32_t *next_block;  // persistent kernels: the launch's work counter (zero at launch), see fetch_block
	AmdScratch s;
};

__device__ __forceinline__ uint64_t pack_idx(const int *idx, int n) {
	uint64_t v = 0;
	for (int i = 0; i < n; i++) v |= (uint64_t) (idx[i] & 15) << (4 * i);
	return v;
}
__device__ __forceinline__ void unpack_idx(uint64_t v, int *idx, int n) {
	for (int i = 0; i < n; i++) idx[i] = (int) ((v >> (4 * i)) & 15u);
}
__device__ __forceinline__ uint32_t pack_ep(const int e[4]) {
	return (uint32_t) (e[0] & 255) | ((uint32_t) (e[1] & 255) << 8) | ((uint32_t) (e[2] & 255) << 16) | ((uint32_t) (e[3] & 255) << 24);
}

// The persistent kernels (cube, window) hand out the blocks of the chunk one at a time from a counter: the block times
// are data dependent (a mode is skipped for half of the benchmark image), so any static split leaves SMs idle at the end
// of a launch (4 .. 6 % of every launch with 8 CTAs per resident slot).
__device__ __forceinline__ uint32_t fetch_block(uint32_t *counter, unsigned lane) {
	uint32_t b = 0;
	if (lane == 0) b = atomicAdd(counter, 1u);
	return __shfl_sync(FULL, b, 0);
}

struct BlockCoord {
	uint64_t gblock;
	uint32_t bx, by, slice;
};
__device__ __forceinline__ BlockCoord block_coord(const AmdParams &p, uint32_t local) {
	BlockCoord c;
	c.gblock = p.block0 + local;
	const uint64_t per_slice = (uint64_t) p.img.blocks_x * p.img.blocks_y;
	c.slice = (uint32_t) (c.gblock / per_slice);
	const uint32_t rem = (uint32_t) (c.gblock - (uint64_t) c.slice * per_slice);
	c.by = rem / p.img.blocks_x;
	c.bx = rem - c.by * p.img.blocks_x;
	return c;
}

// One texel of an 8-bit source as packed bytes (the values the encoder sees are exactly the source bytes:
// (x / 255.0f) * 255.0f == x; missing channels g = b = 0, a = 255)
__device__ __forceinline__ uint32_t fetch_rgba_u8(const SrcImage &img, const BlockCoord &c, int i) {
	const int fmt = img.format;
	if (fmt == B200IC_FMT_RGBA8 || fmt == B200IC_FMT_RGBA8_SRGB) {
		const uint32_t x = min(c.bx * 4 + (i & 3), img.width - 1), y = min(c.by * 4 + (i >> 2), img.height - 1);
		const uint8_t *row = img.base + (uint64_t) c.slice * img.slice_pitch + (uint64_t) y * img.row_pitch;
		return __ldg(reinterpret_cast<const uint32_t *>(row) + x);
	}
	if (fmt == B200IC_FMT_BLOCKS_RGBA8) return __ldg(reinterpret_cast<const uint32_t *>(img.base) + c.gblock * 16 + i);
	const float4 t = fetch_rgba(img, c.gblock, c.bx, c.by, c.slice, i);
	return (uint32_t) (t.x * 255.0f) | ((uint32_t) (t.y * 255.0f) << 8) | ((uint32_t) (t.z * 255.0f) << 16) | ((uint32_t) (t.w * 255.0f) << 24);
}

// the subset of a task as the single-colour path wants it
__device__ __forceinline__ void task_subset(const Task &t, U8Subset &S) {
	int sum[4] = {0, 0, 0, 0};
	for (int i = 0; i < t.n; i++) {
		S.d[i] = t.d[i];
#pragma unroll
		for (int j = 0; j < 4; j++) sum[j] += (int) ((t.d[i] >> (8 * j)) & 255u);
	}
	S.n = t.n;
	S.all_same = t.all_same != 0;
	for (int j = 0; j < 4; j++) S.mean[j] = j < t.dim ? (real) sum[j] / (real) t.n : 0;
}

// Task table of a single-index mode: lane = attempt * subsets + subset takes the texels of its subset in texel order
__device__ __forceinline__ void build_single_index_task(Task &t, const uint32_t *px, int subsets, int part, int s, const ShakeParams &sp, uint64_t idx_q) {
	const uint32_t keep = sp.dim == 3 ? 0x00ffffffu : 0xffffffffu;
	int n = 0;
	bool same = true;
	for (int i = 0; i < 16; i++)
		if (subset_of(subsets, part, i) == s) {
			const uint32_t v = px[i] & keep;
			t.d[n] = v;
			same = same && (v == t.d[0]);
			n++;
		}
	for (int i = n; i < 16; i++) t.d[i] = 0;
	t.idx_q = idx_q;
	t.n = (uint8_t) n;
	t.clog = (uint8_t) ilog2(sp.clusters);
	t.bits = (uint8_t) sp.bits[0];
	t.type = (uint8_t) sp.parity;
	t.all_same = same ? 1 : 0;
	t.dim = (uint8_t) sp.dim;
	t.w_bits_total = (uint8_t) sp.bits[3];
	t.w_size = (uint8_t) sp.shake_size;
	t.w_index = idx_q;
	t.w_active = 1;
	t.done = 0;
	t.item_base = t.item_count = 0;
}

// Task table of a dual-index mode (4, 5): lane = 2 * (rotation * selections + index selection) + (0 vector | 1 scalar).
// The scalar channel is searched as a 3-vector with the channel replicated (src/amd_bc7_body.cpp:1094-1096).
struct DualShape {
	int nsel, combos, ntasks;
};
__device__ __forceinline__ DualShape dual_shape(const ModeInfo &mi) {
	DualShape s;
	s.nsel = 1 << mi.index_mode_bits;
	s.combos = (1 << mi.rotation_bits) * s.nsel;
	s.ntasks = s.combos * 2;
	return s;
}
__device__ __forceinline__ int dual_index_bits(const ModeInfo &mi, int isel, int which) {
	return which == 0 ? (isel ? mi.index_bits1 : mi.index_bits0) : (isel ? mi.index_bits0 : mi.index_bits1);
}
__device__ __forceinline__ void build_dual_index_task(Task &t, const uint32_t *px, const ModeInfo &mi, int lane, uint64_t idx_q) {
	const DualShape ds = dual_shape(mi);
	const int combo = lane >> 1, which = lane & 1;
	const int rot = combo / ds.nsel, isel = combo - rot * ds.nsel;
	const int c0 = rotation_channel(rot, 0), c1 = rotation_channel(rot, 1), c2 = rotation_channel(rot, 2), c3 = rotation_channel(rot, 3);
	bool same = true;
	for (int i = 0; i < 16; i++) {
		const uint32_t v = px[i];
		uint32_t w;
		if (which == 0) w = ((v >> (8 * c1)) & 255u) | (((v >> (8 * c2)) & 255u) << 8) | (((v >> (8 * c3)) & 255u) << 16);
		else w = ((v >> (8 * c0)) & 255u) * 0x010101u;
		t.d[i] = w;
		same = same && (w == t.d[0]);
	}
	const int cb = which == 0 ? mi.vector_bits / 3 : mi.scalar_bits;
	t.idx_q = idx_q;
	t.n = 16;
	t.clog = (uint8_t) dual_index_bits(mi, isel, which);
	t.bits = (uint8_t) cb;
	t.type = CART;
	t.all_same = same ? 1 : 0;
	t.dim = 3;
	t.w_bits_total = (uint8_t) (6 * cb);
	t.w_size = 6;
	t.w_index = idx_q;
	t.w_active = 1;
	t.done = 0;
	t.item_base = t.item_count = 0;
}

// Tasks with the same texel set.  The 3-subset partition tables repeat subsets (36 distinct masks among the 48 of mode 0, 140
// among the 192 of mode 2), and the 8 partitions the quantiser ranks best are similar: 11 .. 12 % of the 24 tasks of a block
// repeat an earlier one on the benchmark image.  Quantiser and shakers are pure functions of the texel list and the mode, so
// the repeat takes the first task's result.  Returns the first task with this lane's texel mask (== lane: not a repeat).
__device__ __forceinline__ int first_task_with_mask(int subsets, int part, int s, int ntasks, unsigned lane) {
	uint32_t mask = 0;
	for (int i = 0; i < 16; i++) mask |= (subset_of(subsets, part, i) == s ? 1u : 0u) << i;
	const uint32_t key = (int) lane < ntasks ? mask : 0x10000u + lane; // (idle lanes match nobody)
	return __ffs(__match_any_sync(FULL, key)) - 1;
}

// Start (or restart) a pass of ep_shaker_d for one task: collapse the indices, handle the single-index case.
__device__ __noinline__ void cube_begin_pass(const Tables &T, Task &t, uint64_t from) {
	int index[kMaxEntries];
	unpack_idx(from, index, t.n);
	const int Mi = collapse_indices(index, t.n);
	if (Mi == 0) {
		U8Subset S;
		task_subset(t, S);
		const int bits[3] = {t.bits, t.bits, t.bits};
		int e0[2][4];
		const real e = shake_single_index_u8(T, S, t.clog, bits, t.type, 3, index, e0);
		if (e < t.err_o) {
			t.err_o = e;
			t.best_idx = pack_idx(index, t.n);
		}
		t.done = 1;
		t.item_count = 0;
		return;
	}
	t.cur = pack_idx(index, t.n);
	t.Mi = (uint8_t) Mi;
	t.pass_key = ~0ull;
	t.pass_idx = 0;
}

// Lay out the work items of the running pass: item_base / item_count per task (prefix sum over the tasks in `order`).
// Returns the total number of items.
template <typename WS> __device__ __forceinline__ int layout_items(WS &ws, int ntasks, int count, int sortkey, unsigned lane) {
	int rank = 0;
	for (int o = 0; o < ntasks; o++) {
		const int ok = __shfl_sync(FULL, sortkey, o);
		rank += (ok > sortkey || (ok == sortkey && o < (int) lane)) ? 1 : 0;
	}
	if ((int) lane < ntasks) ws.order[rank] = (uint8_t) lane;
	__syncwarp();
	const int owner = (int) lane < ntasks ? ws.order[lane] : 0;
	int mine = __shfl_sync(FULL, count, owner);
	if ((int) lane >= ntasks) mine = 0;
	int incl = mine;
	for (int dlt = 1; dlt < 32; dlt <<= 1) {
		const int v = __shfl_up_sync(FULL, incl, dlt);
		if ((int) lane >= dlt) incl += v;
	}
	const int total = __shfl_sync(FULL, incl, 31);
	if ((int) lane < ntasks) {
		ws.task[owner].item_base = (uint16_t) (incl - mine);
		ws.task[owner].item_count = (uint16_t) mine;
	}
	__syncwarp();
	return total;
}
template <typename WS> __device__ __forceinline__ int item_owner(const WS &ws, int ntasks, int it) {
	int ti = 0;
	for (int r = 0; r < ntasks; r++) {
		const int cand = ws.order[r];
		const int base = ws.task[cand].item_base;
		if (it >= base && it < base + ws.task[cand].item_count) ti = cand;
	}
	return ti;
}

// =====================================================================================================================
// quantise kernel
// =====================================================================================================================
struct QuantScratch {
	real pxc[4][16];      // the block, channel-major, 0..255 (quantiser input, see QuantIO)
	real serr[64][3];
	uint64_t qidx[64][3]; // quantiser indices of every (partition, subset) of the running mode
	union {
		real qs[2][15][32]; // the two lane-strided FP64 work arrays of QuantIO (element k of lane l at [k][l]); a subset has <= 15 texels
		real perr[64];      // (after the quantiser rounds) error per partition
	};
	uint32_t mode_mask;   // after the filter of :1340-1380
	int top[8];
};
constexpr int kQuantCtasPerSm = 5;
static_assert(kWarps * sizeof(QuantScratch) <= 233472 / kQuantCtasPerSm - 1024, "quantise scratch: 5 CTAs per SM");

__global__ void __launch_bounds__(kWarps * 32, kQuantCtasPerSm) amd_quant_kernel(const AmdParams p) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	QuantScratch *scratch = reinterpret_cast<QuantScratch *>(smem_raw);
	const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
	const uint32_t block = blockIdx.x * kWarps + warp;
	if (block >= p.n_blocks) return; // whole warp
	QuantScratch &ws = scratch[warp];
	const BlockCoord bc = block_coord(p, block);
	{ // prepare_block (bc7amd_core.cuh) with one texel per lane
		bool na = false, zo = false;
		real v[4] = {0, 0, 0, 0};
		if (lane < 16) {
			const float4 t = fetch_rgba(p.img, bc.gblock, bc.bx, bc.by, bc.slice, (int) lane);
			const float in4[4] = {t.x, t.y, t.z, t.w};
			if (t.w < 1.0) na = true;
			else if (((double) t.w >= 0.99999) || ((double) t.w < 0.00001)) zo = true;
#pragma unroll
			for (int j = 0; j < 4; j++) {
				v[j] = (real) (in4[j] * 255.0f);
				ws.pxc[j][lane] = v[j];
			}
		}
		const bool needs_alpha = __any_sync(FULL, na), zero_one = __any_sync(FULL, zo);
		real range = 0;
#pragma unroll
		for (int j = 0; j < 4; j++) {
			real mn = lane < 16 ? v[j] : A7_HUGE, mx = lane < 16 ? v[j] : 0;
			mx = mx > 0 ? mx : 0; // the reference's running maximum starts at 0
			for (int d = 8; d > 0; d >>= 1) {
				const real mn2 = __shfl_xor_sync(FULL, mn, d), mx2 = __shfl_xor_sync(FULL, mx, d);
				mn = mn2 < mn ? mn2 : mn;
				mx = mx2 > mx ? mx2 : mx;
			}
			const real r = mx - mn;
			range = j == 0 ? r : (range > r ? range : r);
		}
		if (lane == 0) ws.mode_mask = filter_modes(p.mode_mask, needs_alpha, zero_one, range < 1e-10);
	}
	__syncwarp();
	const int mode = p.mode;
	if (!(ws.mode_mask & p.launch_modes & (1u << mode))) { // whole warp
		if (lane == 0) p.s.q_top[(size_t) block * 8] = 0xffu;
		return;
	}
	const ModeInfo mi = mode_info(mode);
	AMD_T0();
	const ShakeParams sp = single_index_shake_params(mode);
	const int nparts = 1 << mi.partition_bits, subsets = mi.subsets;
	const uint8_t *qorder = quantise_order(subsets, nparts);
	QuantIOShared io;
	io.px = (uint32_t) __cvta_generic_to_shared(&ws.pxc[0][0]);
	io.chan = 0xE4u;
	io.proj = (uint32_t) __cvta_generic_to_shared(&ws.qs[0][0][lane]);
	io.dev = (uint32_t) __cvta_generic_to_shared(&ws.qs[1][0][lane]);
	io.stride = 32;
	// the distinct subsets come first in the order (sorted by size among themselves), the repeats after them
	const uint8_t *qfirst = subsets == 3 ? (nparts == 16 ? c_qfirst : c_qfirst + 48) : nullptr;
	const int nproblems = nparts * subsets;
	int ndistinct = nproblems;
	if (qfirst) {
		ndistinct = 0;
		for (int tt = (int) lane; tt < nproblems; tt += 32) ndistinct += qfirst[tt] == tt ? 1 : 0;
		ndistinct = __reduce_add_sync(FULL, ndistinct);
	}
	for (int tt = (int) lane; tt < ndistinct; tt += 32) {
		const int t = qorder[tt];
		const int part = t / subsets, s = t - part * subsets;
		uint32_t smask = 0;
		for (int i = 0; i < 16; i++) smask |= (subset_of(subsets, part, i) == s ? 1u : 0u) << i;
		int n;
		io.texels = texels_of_mask(smask, n);
		uint64_t packed = 0;
		ws.serr[part][s] = n ? quantise_subset(io, n, sp.clusters, sp.dim, packed) : 0;
		ws.qidx[part][s] = packed;
	}
	__syncwarp();
	for (int tt = ndistinct + (int) lane; tt < nproblems; tt += 32) { // repeats
		const int t = qorder[tt], f = qorder[qfirst[tt]];
		ws.serr[t / subsets][t % subsets] = ws.serr[f / subsets][f % subsets];
		ws.qidx[t / subsets][t % subsets] = ws.qidx[f / subsets][f % subsets];
	}
	__syncwarp();
	AMD_T(0);
	for (int part = (int) lane; part < nparts; part += 32) {
		real e = 0;
		for (int s = 0; s < subsets; s++) e += ws.serr[part][s];
		ws.perr[part] = e;
	}
	__syncwarp();
	for (int part = (int) lane; part < nparts; part += 32) { // stable rank (ties keep partition order); the 8 lowest are shaken
		const real e = ws.perr[part];
		int rank = 0;
		for (int q = 0; q < nparts; q++) {
			const real eq = ws.perr[q];
			rank += ((e - eq > 0) || (!(eq - e > 0) && q < part)) ? 1 : 0;
		}
		if (rank < 8) ws.top[rank] = part;
	}
	__syncwarp();
	if ((int) lane < 8 * subsets) {
		const int a = (int) lane / subsets, s = (int) lane - a * subsets;
		p.s.q_idx[(size_t) block * kMaxTasks + lane] = ws.qidx[ws.top[a]][s];
	}
	if (lane < 8) p.s.q_top[(size_t) block * 8 + lane] = (uint8_t) ws.top[lane];
	AMD_T(1);
}

// Dual-index modes 4 / 5: 16 / 8 quantiser problems per block, all on the 16 texels of the block -- a warp takes 2 / 4
// blocks so that its lanes are full: lane = (block of the warp, rotation, index selection, vector | scalar).
struct DualQuantScratch {
	float in[4][64];
	BlockInput B[4];
	real qs[2][16][32];
};
__global__ void __launch_bounds__(kWarps * 32, 4) amd_quant_dual_kernel(const AmdParams p) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	DualQuantScratch *scratch = reinterpret_cast<DualQuantScratch *>(smem_raw);
	const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
	const ModeInfo mi = mode_info(p.mode);
	const DualShape ds = dual_shape(mi);
	const int bpw = 32 / ds.ntasks; // blocks per warp
	const uint32_t block0 = (blockIdx.x * kWarps + warp) * (uint32_t) bpw;
	if (block0 >= p.n_blocks) return; // whole warp
	DualQuantScratch &ws = scratch[warp];
	AMD_T0();
	for (int pass = 0; pass < (bpw + 1) / 2; pass++) { // block set-up, two blocks (16 lanes each) at a time
		const int sub = pass * 2 + ((int) lane >> 4), tx = (int) lane & 15;
		const bool have = sub < bpw && block0 + sub < p.n_blocks;
		bool na = false, zo = false;
		real v[4] = {0, 0, 0, 0};
		if (have) {
			const BlockCoord bc = block_coord(p, block0 + sub);
			const float4 t = fetch_rgba(p.img, bc.gblock, bc.bx, bc.by, bc.slice, tx);
			const float in4[4] = {t.x, t.y, t.z, t.w};
			if (t.w < 1.0) na = true;
			else if (((double) t.w >= 0.99999) || ((double) t.w < 0.00001)) zo = true;
#pragma unroll
			for (int j = 0; j < 4; j++) {
				v[j] = (real) (in4[j] * 255.0f);
				ws.B[sub].px[tx][j] = v[j];
				ws.B[sub].pxc[j][tx] = v[j];
			}
		}
		const unsigned half = 0xffffu << (lane & 16u);
		const bool needs_alpha = (__ballot_sync(FULL, na) & half) != 0, zero_one = (__ballot_sync(FULL, zo) & half) != 0;
		real range = 0;
#pragma unroll
		for (int j = 0; j < 4; j++) {
			real mn = v[j], mx = v[j] > 0 ? v[j] : 0; // the reference's running maximum starts at 0
			for (int d = 8; d > 0; d >>= 1) { // (xor distances below 16 stay inside the block's 16 lanes)
				const real mn2 = __shfl_xor_sync(FULL, mn, d), mx2 = __shfl_xor_sync(FULL, mx, d);
				mn = mn2 < mn ? mn2 : mn;
				mx = mx2 > mx ? mx2 : mx;
			}
			const real r = mx - mn;
			range = j == 0 ? r : (range > r ? range : r);
		}
		if (have && tx == 0) {
			const uint32_t mm = filter_modes(p.mode_mask, needs_alpha, zero_one, range < 1e-10);
			ws.B[sub].mode_mask = mm;
			p.s.q_top[(size_t) (block0 + sub) * 8] = (mm & p.launch_modes & (1u << p.mode)) ? 0 : 0xffu;
		}
	}
	__syncwarp();
	const int sub = (int) lane / ds.ntasks, tk = (int) lane - sub * ds.ntasks;
	if (sub < bpw && block0 + sub < p.n_blocks && (ws.B[sub].mode_mask & p.launch_modes & (1u << p.mode))) {
		const int combo = tk >> 1, which = tk & 1;
		const int rot = combo / ds.nsel, isel = combo - rot * ds.nsel;
		const uint32_t c0 = (uint32_t) rotation_channel(rot, 0), c1 = (uint32_t) rotation_channel(rot, 1),
									 c2 = (uint32_t) rotation_channel(rot, 2), c3 = (uint32_t) rotation_channel(rot, 3);
		QuantIOShared io;
		io.px = (uint32_t) __cvta_generic_to_shared(&ws.B[sub].pxc[0][0]);
		io.texels = 0xFEDCBA9876543210ull;
		io.chan = which == 0 ? (c1 | (c2 << 2) | (c3 << 4)) : (c0 | (c0 << 2) | (c0 << 4));
		io.proj = (uint32_t) __cvta_generic_to_shared(&ws.qs[0][0][lane]);
		io.dev = (uint32_t) __cvta_generic_to_shared(&ws.qs[1][0][lane]);
		io.stride = 32;
		uint64_t qpacked = 0;
		quantise_subset(io, 16, 1 << dual_index_bits(mi, isel, which), 3, qpacked);
		p.s.q_idx[(size_t) (block0 + sub) * kMaxTasks + tk] = qpacked;
	}
	AMD_T(0);
}

// =====================================================================================================================
// cube kernel (ep_shaker_d)
// =====================================================================================================================
struct __align__(16) CubeScratch {
	Task task[kWarpTasks];
	uint64_t tab[4 * 12];     // ramp tables of the running item: [lattice][channel * 4 + endpoint combination]
	uint32_t item_ep[32][6];  // expanded endpoint candidates of the batch's items (cube_item_setup_u8)
	uint32_t lbs[4 * 12];     // second pass: per-ramp bounds of the running item (cube_bound_u8)
	uint32_t plane[12];       //              channel planes of the running task
	uint32_t px[4][16];       // texels of the warp's blocks
	uint8_t surv[256];        //              corner ids of the running item whose bound can still beat the first pass
	uint8_t item_ti[32], item_qp[32];
	uint8_t order[kWarpTasks];
	uint8_t same_as[kWarpTasks]; // first task with the same texels (first_task_with_mask); 0xff: no work (block absent / mode skipped)
};

// Index vector of the pass winner: the palette travels from the winning lane by shuffles, texel i is classified by lane i.
template <int CLOG>
__device__ __forceinline__ void cube_finish_task(Task &t, uint64_t lane_best, const uint32_t *best_pal, unsigned lane) {
	constexpr int C = 1 << CLOG;
	const uint32_t hi = (uint32_t) (lane_best >> 32), lo = (uint32_t) lane_best;
	const uint32_t mh = __reduce_min_sync(FULL, hi);
	const uint32_t ml = __reduce_min_sync(FULL, hi == mh ? lo : 0xffffffffu);
	const int src = __ffs(__ballot_sync(FULL, hi == mh && lo == ml)) - 1;
	uint32_t pal[C];
#pragma unroll
	for (int c = 0; c < C; c++) pal[c] = __shfl_sync(FULL, best_pal[c], src);
	uint32_t nib = 0;
	if ((int) lane < t.n) {
		const uint32_t di = t.d[lane];
		uint32_t m = 0xffffffffu;
#pragma unroll
		for (int c = 0; c < C; c++) m = umin32(m, (sq_dist4(pal[c], di) << 4) | (uint32_t) c);
		nib = m & 15u;
	}
	const uint32_t w0 = __reduce_or_sync(FULL, lane < 8 ? nib << (4 * lane) : 0u);
	const uint32_t w1 = __reduce_or_sync(FULL, (lane >= 8 && lane < 16) ? nib << (4 * (lane - 8)) : 0u);
	if (lane == 0) {
		t.pass_key = ((uint64_t) mh << 32) | ml;
		t.pass_idx = ((uint64_t) w1 << 32) | w0;
	}
}

// All corners of one item, one per lane; folds the item's best into the lane's running best of the task.
template <int CLOG>
__device__ __forceinline__ void cube_item_corners(const CubeScratch &ws, const uint32_t *d, int n, int nlb, int qp, unsigned lane, uint64_t &lane_best,
																									uint32_t *best_pal) {
	uint32_t key, xy;
	cube_lane_corners<CLOG>(ws.tab, d, n, nlb, lane, key, xy);
	const uint64_t k64 = ((uint64_t) (key >> 8) << 16) | ((uint64_t) qp << 8) | (uint64_t) (key & 255u);
	if (k64 < lane_best) {
		lane_best = k64;
		cube_lane_palette<CLOG>(ws.tab, nlb, lane, xy, best_pal);
	}
}

// The same for an item of the SECOND pass, where only corners strictly better than the first pass's error matter: the
// 12 bounds per lattice over the lanes, the corners whose bound does not exceed `thr` compacted into a list (ballot-free:
// per-lane masks + prefix sum), one surviving corner per lane and round.  thr follows the best error found.
template <int CLOG>
__device__ __forceinline__ void cube_item_pruned(CubeScratch &ws, const uint32_t *d, int n, int nlb, int qp, unsigned lane, uint32_t &thr,
																								 uint64_t &lane_best, uint32_t *best_pal) {
	constexpr int C = 1 << CLOG;
	const int nl = 1 << nlb;
	for (int b = (int) lane; b < nl * 12; b += 32) ws.lbs[b] = cube_bound_u8<CLOG>(ws.tab[b], ws.plane + 4 * ((b >> 2) % 3), n);
	__syncwarp();
	uint32_t mask = 0;
#pragma unroll
	for (int r = 0; r < 8; r++)
		if (r < 2 * nl && cube_cid_bound(ws.lbs, (int) lane + 32 * r) <= thr) mask |= 1u << r;
	const int mine = __popc(mask);
	int incl = mine;
#pragma unroll
	for (int dlt = 1; dlt < 32; dlt <<= 1) {
		const int v = __shfl_up_sync(FULL, incl, dlt);
		if ((int) lane >= dlt) incl += v;
	}
	const int S = __shfl_sync(FULL, incl, 31);
	if (S == 0) return; // (uniform) nothing in this item can beat the first pass
	int at = incl - mine;
	while (mask) {
		const int r = __ffs(mask) - 1;
		mask &= mask - 1;
		ws.surv[at++] = (uint8_t) (lane + 32u * r);
	}
	__syncwarp();
	uint32_t item_min = 0xffffffffu;
#pragma unroll 1
	for (int s0 = 0; s0 < S; s0 += 32) {
		if (s0 + (int) lane < S) {
			const int cid = ws.surv[s0 + lane];
			uint32_t pal[C];
			cube_cid_palette<CLOG>(ws.tab, cid, pal);
			const uint32_t e = cube_corner_error_u8<CLOG>(pal, d, n);
			const uint32_t key = cube_cid_key(e, cid);
			const uint64_t k64 = ((uint64_t) (key >> 8) << 16) | ((uint64_t) qp << 8) | (uint64_t) (key & 255u);
			item_min = umin32(item_min, e);
			if (k64 < lane_best) {
				lane_best = k64;
#pragma unroll
				for (int c = 0; c < C; c++) best_pal[c] = pal[c];
			}
		}
	}
	thr = umin32(thr, __reduce_min_sync(FULL, item_min));
}

// The index widths a launch can meet (CLOGS: bit 2 = 2-bit, bit 3 = 3-bit indices) are a property of the mode; only the
// dual-index mode 4 has both.  The code of the other width is compiled out: the cube and window kernels stall on
// instruction fetch more than on anything else when their loop bodies do not fit the instruction caches.
template <int CLOGS> __device__ __forceinline__ bool is_clog2(int clog) { return CLOGS == 4 || (CLOGS == 12 && clog == 2); }
__host__ __device__ constexpr int mode_clogs(int mode) { return mode == 4 ? 12 : ((mode == 0 || mode == 1) ? 8 : 4); }

// ep_shaker_d for all tasks of the warp. On return task[i].err_o / best_idx hold its result.
template <int CLOGS, bool PRUNE2>
__device__ __forceinline__ void cube_phase(const Tables &T, CubeScratch &ws, const uint32_t *lut2, const uint32_t *lut3, int ntasks, unsigned lane) {
	constexpr bool prune2 = PRUNE2;
	if ((int) lane < ntasks) {
		Task &t = ws.task[lane];
		t.err_o = A7_HUGE;
		t.best_idx = t.idx_q;
		t.done = 0;
		if (ws.same_as[lane] == lane) {
			cube_begin_pass(T, t, t.idx_q);
		} else { // a repeat: takes the first task's result afterwards
			t.done = 1;
			t.item_count = 0;
		}
	}
	__syncwarp();
#pragma unroll 1
	for (int pass = 0; pass < 2; pass++) {
		int count = 0;
		if ((int) lane < ntasks && !ws.task[lane].done) count = qp_count(ws.task[lane].Mi, (1 << ws.task[lane].clog) - 1);
		const int total = layout_items(ws, ntasks, count, 0, lane); // equal sort keys: task order
		if (total == 0) break;
		AMD_COUNT(6, total);
		AMD_COUNT(10, 1);
		int cur_ti = -1, n = 0, clog = 3, nlb = 0, bcc = 0;
		uint32_t thr = 0xffffffffu; // second pass: what a corner has to beat (the first pass's error - 1, then the best found)
		const bool prune = prune2 && pass == 1;
		const uint32_t *d = ws.task[0].d; // texels of the running task (shared memory: the lanes read them together)
		uint64_t lane_best = ~0ull;
		uint32_t best_pal[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 1
		for (int b0 = 0; b0 < total; b0 += 32) {
			AMD_COUNT(7, 1);
			{ // set-up: one item per lane
				const int it = b0 + (int) lane;
				if (it < total) {
					const int ti = item_owner(ws, ntasks, it);
					const Task &t = ws.task[ti];
					const int qp = it - t.item_base;
					int q, p;
					qp_decode(qp, t.Mi, (1 << t.clog) - 1, q, p);
					const int use_par = (t.type == BCC || t.type == SAME_PAR) ? 1 : 0;
					uint32_t ep[6];
					if (is_clog2<CLOGS>(t.clog)) cube_item_setup_u8<2>(t.d, t.n, t.cur, q, p, t.bits, use_par, ep);
					else cube_item_setup_u8<3>(t.d, t.n, t.cur, q, p, t.bits, use_par, ep);
#pragma unroll
					for (int k = 0; k < 6; k++) ws.item_ep[lane][k] = ep[k];
					ws.item_ti[lane] = (uint8_t) ti;
					ws.item_qp[lane] = (uint8_t) qp;
				}
			}
			__syncwarp();
			const int cnt = min(32, total - b0);
#pragma unroll 1
			for (int j = 0; j < cnt; j++) {
				const int ti = ws.item_ti[j];
				if (ti != cur_ti) { // (uniform) next task: close the previous one, take the new texels into registers
					if (cur_ti >= 0) {
						if (is_clog2<CLOGS>(clog)) cube_finish_task<2>(ws.task[cur_ti], lane_best, best_pal, lane);
						else cube_finish_task<3>(ws.task[cur_ti], lane_best, best_pal, lane);
					}
					const Task &t = ws.task[ti];
					d = t.d;
					n = t.n;
					clog = t.clog;
					bcc = t.type == BCC ? 1 : 0;
					nlb = t.type == BCC ? 2 : (t.type == SAME_PAR ? 1 : 0);
					lane_best = ~0ull;
					cur_ti = ti;
					if (prune) {
						__syncwarp();
						if (lane < 12) {
							const int k = (int) lane >> 2, w = (int) lane & 3;
							uint32_t v = 0;
							for (int b = 0; b < 4; b++)
								if (4 * w + b < t.n) v |= ((t.d[4 * w + b] >> (8 * k)) & 255u) << (8 * b);
							ws.plane[lane] = v;
						}
						thr = t.err_o >= 1. ? (uint32_t) t.err_o - 1u : 0u;
					}
				}
				// ramp tables of the item's lattices, 4 entries per lane and step
				{
					const uint32_t *ep = ws.item_ep[j];
					uint32_t *tw = reinterpret_cast<uint32_t *>(ws.tab);
					if (is_clog2<CLOGS>(clog)) {
						for (int id = (int) lane; id < (12 << nlb); id += 32) tw[2 * id] = cube_tab_word_lut<2>(lut2, ep, bcc, id);
					} else {
						for (int id = (int) lane; id < (24 << nlb); id += 32) tw[id] = cube_tab_word_lut<3>(lut3, ep, bcc, id);
					}
				}
				__syncwarp();
				const int qp = ws.item_qp[j];
				if (prune) {
					if (is_clog2<CLOGS>(clog)) cube_item_pruned<2>(ws, d, n, nlb, qp, lane, thr, lane_best, best_pal);
					else cube_item_pruned<3>(ws, d, n, nlb, qp, lane, thr, lane_best, best_pal);
				} else {
					if (is_clog2<CLOGS>(clog)) cube_item_corners<2>(ws, d, n, nlb, qp, lane, lane_best, best_pal);
					else cube_item_corners<3>(ws, d, n, nlb, qp, lane, lane_best, best_pal);
				}
				__syncwarp();
			}
		}
		if (cur_ti >= 0) {
			if (is_clog2<CLOGS>(clog)) cube_finish_task<2>(ws.task[cur_ti], lane_best, best_pal, lane);
			else cube_finish_task<3>(ws.task[cur_ti], lane_best, best_pal, lane);
		}
		__syncwarp();
		// ---- per task: the reference's change / better logic (:1372-1400)
		if ((int) lane < ntasks && !ws.task[lane].done && ws.task[lane].pass_key == ~0ull) ws.task[lane].done = 1; // pruned pass: nothing could beat the first pass
		if ((int) lane < ntasks && !ws.task[lane].done) {
			Task &t = ws.task[lane];
			const real err_2 = (real) (uint32_t) (t.pass_key >> 16);
			const int qp = (int) ((t.pass_key >> 8) & 255u);
			int q0, p0;
			qp_decode(qp, t.Mi, (1 << t.clog) - 1, q0, p0);
			int change = 0;
			for (int k = 0; k < t.n; k++)
				change = change || ((int) ((t.cur >> (4 * k)) & 15u) * q0 + p0 != (int) ((t.pass_idx >> (4 * k)) & 15u));
			const int better = err_2 < t.err_o;
			if (better) {
				t.best_idx = t.pass_idx;
				t.err_o = err_2;
			}
			if (!(change && better) || pass == 1) t.done = 1;
			else cube_begin_pass(T, t, t.pass_idx);
		}
		__syncwarp();
	}
}

// Persistent: one CTA per resident slot, every warp takes the next block of the chunk from the launch's counter; the two
// difference tables of the ramps (6 KB, bc7amd_int.cuh) are built once per CTA.
constexpr int kCubeCtasPerSm = 5;
template <int CLOGS, bool PRUNE2> __global__ void __launch_bounds__(kWarps * 32, kCubeCtasPerSm) amd_cube_kernel(const AmdParams p) {
	__shared__ CubeScratch scratch[kWarps];
	__shared__ uint32_t lut2[RampLutShape<2>::kWords], lut3[RampLutShape<3>::kWords];
	if (CLOGS & 4) ramp_lut_fill<2>(lut2, (int) threadIdx.x, kWarps * 32);
	if (CLOGS & 8) ramp_lut_fill<3>(lut3, (int) threadIdx.x, kWarps * 32);
	__syncthreads();
	const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
	CubeScratch &ws = scratch[warp];
	const Tables T{p.sp};
	const ModeInfo mi = mode_info(p.mode);
	const ShakeParams sp = single_index_shake_params(p.mode);
	const int subsets = mi.subsets, ntasks = mi.alpha == 2 ? dual_shape(mi).ntasks : 8 * subsets;
	// modes with 16 / 8 tasks per block (two subsets; dual index) -- the warp takes 2 / 4 blocks at a time so that the
	// lane = task and lane = item stages run on full warps
	const int bpw = 32 / ntasks, wtasks = bpw * ntasks;
	const int sub = (int) lane / ntasks, tk = (int) lane - sub * ntasks; // this lane's task: block of the group, task of the block
#pragma unroll 1
	for (;;) {
		const uint32_t block0 = fetch_block(p.next_block, lane) * (uint32_t) bpw;
		if (block0 >= p.n_blocks) break;
		const uint32_t block = block0 + (uint32_t) sub;
		const bool have = sub < bpw && block < p.n_blocks && p.s.q_top[(size_t) block * 8] != 0xffu; // (0xff: mode not searched for this block)
		if (!__any_sync(FULL, have)) continue;
		__syncwarp();
		for (int b = (int) lane >> 4; b < bpw; b += 2) // 16 lanes per block
			if (block0 + b < p.n_blocks) ws.px[b][lane & 15] = fetch_rgba_u8(p.img, block_coord(p, block0 + b), (int) lane & 15);
		__syncwarp();
		AMD_T0();
		if (have) {
			const uint64_t idx_q = p.s.q_idx[(size_t) block * kMaxTasks + tk];
			if (mi.alpha == 2) {
				build_dual_index_task(ws.task[lane], ws.px[sub], mi, tk, idx_q);
			} else {
				const int a = tk / subsets, s = tk - a * subsets;
				build_single_index_task(ws.task[lane], ws.px[sub], subsets, p.s.q_top[(size_t) block * 8 + a], s, sp, idx_q);
			}
		}
		{
			int first = have ? (int) lane : 0xff;
			if (subsets == 3) { // (uniform; the 2-subset tables and the dual-index tasks have no repeats)
				const int a = (int) lane < ntasks ? (int) lane / subsets : 0, s = (int) lane - a * subsets;
				first = first_task_with_mask(subsets, p.s.q_top[(size_t) block0 * 8 + a], s, ntasks, lane);
			}
			if ((int) lane < wtasks) ws.same_as[lane] = (uint8_t) first;
		}
		__syncwarp();
		cube_phase<CLOGS, PRUNE2>(T, ws, lut2, lut3, wtasks, lane);
		if (have) {
			const Task &t = ws.task[ws.same_as[lane]];
			p.s.c_idx[(size_t) block * kMaxTasks + tk] = t.best_idx;
			p.s.c_err[(size_t) block * kMaxTasks + tk] = t.err_o;
		}
		AMD_T(2);
	}
}

// =====================================================================================================================
// window kernel (ep_shaker_2_d, best attempt, packing)
// =====================================================================================================================
constexpr int kWinBatch = 32; // work items per batch of the window phase
// NT tasks per warp in NB blocks; wres holds ND channels with RS slots each (4: one per parity combination; 1: the
// dual-index modes, whose lattices have no parity)
template <int NT, int RS, int ND, int NB> struct __align__(16) WindowScratchT {
	static constexpr int kResStride = RS, kBlocks = NB;
	Task task[NT];
	uint64_t item_key[kWinBatch];
	uint64_t item_idx[kWinBatch];
	union {
		uint64_t wres[kWinBatch][ND * RS]; // window_sub_search_u8 results of the batch: [item][channel * RS + pp0 * 2 + pp1]
		ShakeOut so[NT];                   // the tasks' results (after the window phases)
	};
	real wepa[kWinBatch][8];             // window_item_fit_u8: least-squares endpoints [endpoint * 4 + channel]
	uint32_t wsel[kWinBatch][4];         //                     byte-permute selectors
	uint32_t px[NB][16];                 // texels of the warp's blocks
	uint8_t item_ti[kWinBatch], item_qp[kWinBatch];
	uint8_t order[NT];
	uint8_t same_as[NT]; // first task with the same texels (first_task_with_mask)
	uint8_t top[NB][8];
};
using WindowScratch1 = WindowScratchT<kMaxTasks, 4, 4, 1>;     // modes 0, 2 (24 tasks per block) and 7 (4 channels): one block per warp
using WindowScratch2 = WindowScratchT<kWarpTasks, 4, 3, 2>;    // modes 1, 3 (16 tasks per block): 2 blocks per warp
using WindowScratchDual = WindowScratchT<kWarpTasks, 1, 3, 4>; // dual-index modes: 2 (mode 4) / 4 (mode 5) blocks per warp

// Start a round of ep_shaker_2_d for one task (:785-827): collapse, single-index case.
__device__ __noinline__ void window_begin_round(const Tables &T, Task &t) {
	int index[kMaxEntries];
	unpack_idx(t.w_index, index, t.n);
	const int Mi = collapse_indices(index, t.n);
	if (Mi == 0) {
		U8Subset S;
		task_subset(t, S);
		const int mb = (t.w_bits_total + 2 * t.dim - 1) / (2 * t.dim);
		const int bits[4] = {mb, mb, mb, mb};
		int e0[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
		const real e = shake_single_index_u8(T, S, t.clog, bits, t.w_bits_total % (2 * t.dim), t.dim, index, e0);
		if (e < t.w_err_o) {
			t.w_err_o = e;
			t.w_best_idx = pack_idx(index, t.n);
			t.w_best_ep = pack_ep8(e0);
		}
		t.done = 1;
		t.item_count = 0;
		return;
	}
	t.cur = pack_idx(index, t.n);
	t.Mi = (uint8_t) Mi;
	t.pass_key = ~0ull;
	t.pass_idx = 0;
}

// ep_shaker_2_d for the tasks with w_active set, starting from task.w_index. Results in w_err_o / w_best_idx / w_best_ep.
template <int CLOGS, typename WS>
__device__ __noinline__ void window_phase(const Tables &T, WS &ws, const uint32_t *lut2, const uint32_t *lut3, int ntasks, int dim, int type, unsigned lane) {
	if ((int) lane < ntasks) {
		Task &t = ws.task[lane];
		t.done = t.w_active ? 0 : 1;
		if (t.w_active) {
			t.w_err_o = A7_HUGE;
			t.w_best_idx = t.w_index;
			t.w_best_ep = 0;
			t.w_tries = 8;
			window_begin_round(T, t);
		}
	}
	__syncwarp();
	for (int round = 0; round < 9; round++) {
		int count = 0, sortkey = -1;
		if ((int) lane < ntasks && !ws.task[lane].done) {
			const Task &t = ws.task[lane];
			count = qp_count(t.Mi, (1 << t.clog) - 1);
			sortkey = t.clog * 32 + t.n;
		}
		const int total = layout_items(ws, ntasks, count, sortkey, lane);
		if (total == 0) break;
		AMD_COUNT(8, total);
		AMD_COUNT(9, (total + 31) / 32);
		// every task of a launch has the same dimension and parity type (the endpoint bits differ between the vector and
		// scalar tasks of the dual-index modes): dim x (1 | 2 | 4) independent searches per item
		const int use_par = type != 0;
		const int ncombo = type == BCC ? 4 : (type == SAME_PAR ? 2 : 1), per = dim * ncombo;
		const int nbatch = (total + kWinBatch - 1) / kWinBatch, batch = (total + nbatch - 1) / nbatch; // (even batches)
		for (int b0 = 0; b0 < total; b0 += batch) {
			const int cnt = min(batch, total - b0);
			if ((int) lane < cnt) { // fit: one item per lane
				const int it = b0 + (int) lane, ti = item_owner(ws, ntasks, it);
				const Task &t = ws.task[ti];
				const int qp = it - t.item_base;
				int q, p;
				qp_decode(qp, t.Mi, (1 << t.clog) - 1, q, p);
				real epa[8];
				uint32_t sel[4];
				if (is_clog2<CLOGS>(t.clog)) window_item_fit_u8<2>(t.d, t.n, t.cur, q, p, dim, epa, sel);
				else window_item_fit_u8<3>(t.d, t.n, t.cur, q, p, dim, epa, sel);
#pragma unroll
				for (int k = 0; k < 8; k++) ws.wepa[lane][k] = epa[k];
#pragma unroll
				for (int k = 0; k < 4; k++) ws.wsel[lane][k] = sel[k];
				ws.item_ti[lane] = (uint8_t) ti;
				ws.item_qp[lane] = (uint8_t) qp;
			}
			__syncwarp();
			for (int sub = (int) lane; sub < cnt * per; sub += 32) { // search: one (item, channel, parity combination) per lane
				const int item = sub / per, rest = sub - item * per, j = rest / ncombo, combo = rest - j * ncombo;
				const int pp0 = type == SAME_PAR ? combo : (combo >> 1), pp1 = type == SAME_PAR ? combo : (combo & 1);
				const Task &t = ws.task[ws.item_ti[item]];
				const int mb = (t.w_bits_total + 2 * dim - 1) / (2 * dim);
				const uint4 sv = *reinterpret_cast<const uint4 *>(ws.wsel[item]);
				const uint4 pv = *reinterpret_cast<const uint4 *>(t.plane + 4 * j);
				const uint32_t sel[4] = {sv.x, sv.y, sv.z, sv.w}, plw[4] = {pv.x, pv.y, pv.z, pv.w};
				const real ep0 = ws.wepa[item][j], ep1 = ws.wepa[item][4 + j];
				ws.wres[item][j * WS::kResStride + pp0 * 2 + pp1] = is_clog2<CLOGS>(t.clog) ? window_sub_search_u8<2>(lut2, ep0, ep1, sel, plw, t.n, mb, use_par, t.w_size, pp0, pp1)
																													 : window_sub_search_u8<3>(lut3, ep0, ep1, sel, plw, t.n, mb, use_par, t.w_size, pp0, pp1);
			}
			__syncwarp();
			if ((int) lane < cnt) { // combine
				uint64_t epo;
				const uint32_t err = window_item_combine(ws.wres[lane], type, dim, epo, WS::kResStride);
				ws.item_key[lane] = ((uint64_t) err << 8) | (uint64_t) (255 - ws.item_qp[lane]); // `<=`: the LAST minimum wins
				ws.item_idx[lane] = epo;
			}
			__syncwarp();
			if ((int) lane < ntasks && !ws.task[lane].done) {
				Task &t = ws.task[lane];
				const int lo = max(b0, (int) t.item_base), hi = min(b0 + cnt, (int) t.item_base + (int) t.item_count);
				for (int it = lo; it < hi; it++)
					if (ws.item_key[it - b0] < t.pass_key) {
						t.pass_key = ws.item_key[it - b0];
						t.pass_idx = ws.item_idx[it - b0];
					}
			}
			__syncwarp();
		}
		if ((int) lane < ntasks && !ws.task[lane].done) {
			Task &t = ws.task[lane];
			const int qp = 255 - (int) (t.pass_key & 255u);
			int q0, p0;
			qp_decode(qp, t.Mi, (1 << t.clog) - 1, q0, p0);
			const int mb = (t.w_bits_total + 2 * t.dim - 1) / (2 * t.dim);
			uint64_t idg;
			uint32_t err_r;
			if (is_clog2<CLOGS>(t.clog)) err_r = recluster_u8<2>(t.d, t.n, t.pass_idx, mb, t.dim, idg);
			else err_r = recluster_u8<3>(t.d, t.n, t.pass_idx, mb, t.dim, idg);
			int change = 0;
			for (int k = 0; k < t.n; k++) change = change || ((int) ((t.cur >> (4 * k)) & 15u) * q0 + p0 != (int) ((idg >> (4 * k)) & 15u));
			const int better = (real) err_r < t.w_err_o;
			if (better) {
				t.w_best_idx = t.w_index = idg;
				t.w_best_ep = t.pass_idx;
				t.w_err_o = (real) err_r;
			}
			if (!(change && better) || t.w_tries == 0) {
				t.done = 1;
			} else {
				t.w_tries--;
				window_begin_round(T, t);
			}
		}
		__syncwarp();
	}
}

// EncodeSingleIndexBlock (pack_single_index, bc7amd_core.cuh) spread over the warp: lane i < 16 = texel i (its index, the
// anchor flip of its subset, its bit position by a prefix sum of the index widths), lanes 0 .. = the endpoint fields and
// p-bits, lane 31 = mode and partition; the 128-bit block is the OR of the lanes' contributions.
__device__ __forceinline__ void put128(uint64_t &lo, uint64_t &hi, uint32_t v, int n, int pos) {
	if (n <= 0) return;
	v &= (n >= 32) ? 0xffffffffu : ((1u << n) - 1u);
	if (pos < 64) {
		lo |= (uint64_t) v << pos;
		if (pos + n > 64) hi |= (uint64_t) v >> (64 - pos);
	} else {
		hi |= (uint64_t) v << (pos - 64);
	}
}
__device__ __forceinline__ void pack_single_index_warp(int mode, int partition, const ShakeOut *so, unsigned lane, uint64_t &out0, uint64_t &out1) {
	const ModeInfo mi = mode_info(mode);
	const int dim = mi.alpha == 0 ? 3 : 4, cbits = mi.alpha == 0 ? mi.vector_bits / 3 : mi.vector_bits / 4;
	const int ib = mi.index_bits0, subsets = mi.subsets;
	const int fix1 = subsets == 3 ? kBc7Anchor3a[partition] : (subsets == 2 ? kBc7Anchor2[partition] : 0);
	const int fix2 = subsets == 3 ? kBc7Anchor3b[partition] : 0;
	const int i = (int) lane & 15;
	const int p = subset_of(subsets, partition, i);
	const unsigned m0 = __ballot_sync(FULL, lane < 16 && p == 0), m1 = __ballot_sync(FULL, lane < 16 && p == 1), m2 = __ballot_sync(FULL, lane < 16 && p == 2);
	const unsigned mine = p == 0 ? m0 : (p == 1 ? m1 : m2);
	const int cnt = __popc(mine & ((1u << i) - 1u));
	uint32_t blk = (uint32_t) ((so[p].idx >> (4 * cnt)) & 15u);
	const int f0 = (int) ((__shfl_sync(FULL, blk, 0) >> (ib - 1)) & 1u), f1 = (int) ((__shfl_sync(FULL, blk, fix1) >> (ib - 1)) & 1u),
						f2 = (int) ((__shfl_sync(FULL, blk, fix2) >> (ib - 1)) & 1u);
	const int flip0 = f0, flip1 = subsets > 1 ? f1 : 0, flip2 = subsets > 2 ? f2 : 0;
	if (p == 0 ? flip0 : (p == 1 ? flip1 : flip2)) blk = (uint32_t) ((1 << ib) - 1) - blk;
	const int myfix = p == 0 ? 0 : (p == 1 ? fix1 : fix2);
	const int width = lane < 16 ? ib - (i == myfix ? 1 : 0) : 0;
	int incl = width;
#pragma unroll
	for (int d = 1; d < 16; d <<= 1) {
		const int v = __shfl_up_sync(FULL, incl, d);
		if ((int) lane >= d) incl += v;
	}
	const int header = mode + 1 + mi.partition_bits, nfields = dim * subsets * 2;
	const int npb = mi.parity == SAME_PAR ? subsets : (mi.parity == BCC ? 2 * subsets : 0);
	uint64_t lo = 0, hi = 0;
	if (lane < 16) put128(lo, hi, blk, width, header + nfields * cbits + npb + incl - width);
	if ((int) lane < nfields) { // endpoint field (channel k, subset s, endpoint e), stream order k, s, e
		const int k = (int) lane / (subsets * 2), rem = (int) lane - k * subsets * 2, s = rem >> 1, e = rem & 1;
		const int a = s == 0 ? flip0 : (s == 1 ? flip1 : flip2);
		const uint32_t v = (so[s].ep[e ^ a] >> (8 * k)) & 255u;
		put128(lo, hi, mi.parity != CART ? v >> 1 : v, cbits, header + (int) lane * cbits);
	} else if ((int) lane < nfields + npb) { // p-bits; ONE_PBIT quirk: both from endpoint 1 after the flip (:443-448)
		const int j = (int) lane - nfields, s = mi.parity == SAME_PAR ? j : (j >> 1), e = mi.parity == SAME_PAR ? 1 : (j & 1);
		const int a = s == 0 ? flip0 : (s == 1 ? flip1 : flip2);
		put128(lo, hi, so[s].ep[e ^ a] & 1u, 1, header + nfields * cbits + j);
	}
	if (lane == 31) {
		put128(lo, hi, 1u << mode, mode + 1, 0);
		put128(lo, hi, (uint32_t) partition, mi.partition_bits, mode + 1);
	}
	const uint32_t w0 = __reduce_or_sync(FULL, (uint32_t) lo), w1 = __reduce_or_sync(FULL, (uint32_t) (lo >> 32)),
								 w2 = __reduce_or_sync(FULL, (uint32_t) hi), w3 = __reduce_or_sync(FULL, (uint32_t) (hi >> 32));
	out0 = (uint64_t) w0 | ((uint64_t) w1 << 32);
	out1 = (uint64_t) w2 | ((uint64_t) w3 << 32);
}

// A group of 2 / 4 blocks of a dual-index mode (4 / 5): lane = (block of the group, task of the block)
template <int CLOGS>
__device__ __forceinline__ void window_group_dual(const AmdParams &p, WindowScratchDual &ws, const Tables &T, const uint32_t *lut2, const uint32_t *lut3,
																									uint32_t group, unsigned lane) {
	const int mode = p.mode;
	const ModeInfo mi = mode_info(mode);
	const DualShape ds = dual_shape(mi);
	const int ntasks = ds.ntasks, bpw = 32 / ntasks;
	const int sub = (int) lane / ntasks, tk = (int) lane - sub * ntasks;
	const uint32_t block0 = group * (uint32_t) bpw, block = block0 + (uint32_t) sub;
	const bool present = block < p.n_blocks;
	const bool have = present && p.s.q_top[(size_t) block * 8] != 0xffu; // (0xff: mode not searched for this block)
	if (p.first && present && !have && tk == 0) {
		p.dst[p.block0 + block] = make_uint4(0, 0, 0, 0);
		p.best_err[p.block0 + block] = A7_HUGE;
	}
	if (!__any_sync(FULL, have)) return;
	__syncwarp();
	for (int b = (int) lane >> 4; b < bpw; b += 2) // 16 lanes per block
		if (block0 + b < p.n_blocks) ws.px[b][lane & 15] = fetch_rgba_u8(p.img, block_coord(p, block0 + b), (int) lane & 15);
	__syncwarp();
	AMD_T0();
	{
		Task &t = ws.task[lane];
		if (have) {
			build_dual_index_task(t, ws.px[sub], mi, tk, p.s.q_idx[(size_t) block * kMaxTasks + tk]);
			t.best_idx = p.s.c_idx[(size_t) block * kMaxTasks + tk];
			t.err_o = p.s.c_err[(size_t) block * kMaxTasks + tk];
			t.w_index = t.best_idx; // ep_shaker_2_d runs on ep_shaker_d's indices only (src/amd_bc7_body.cpp:1120-1160)
			window_planes_u8(t.d, t.n, t.plane);
		} else {
			t.w_active = 0;
		}
	}
	__syncwarp();
	window_phase<CLOGS>(T, ws, lut2, lut3, kWarpTasks, 3, CART, lane);
	AMD_T(3);
	{
		const Task &t = ws.task[lane];
		ShakeOut o;
		o.err = t.w_err_o;
		o.idx = t.w_best_idx;
		o.ep[0] = (uint32_t) t.w_best_ep;
		o.ep[1] = (uint32_t) (t.w_best_ep >> 32);
		__syncwarp(); // (so shares its storage with the search results of the phase)
		ws.so[lane] = o;
	}
	__syncwarp();
	if (have && tk == 0) { // one lane per block: best combination, packing
		const ShakeOut *so = ws.so + sub * ntasks;
		real be = A7_HUGE;
		int bcm = 0;
		for (int c = 0; c < ds.combos; c++) {
			real e = 0;
			e += so[2 * c].err;
			e += so[2 * c + 1].err / 3.;
			if (e < be) { be = e; bcm = c; }
		}
		const uint64_t gblock = p.block0 + block;
		const real carried = p.first ? A7_HUGE : p.best_err[gblock];
		if (p.first || be < carried) {
			int epp[2][2][4], idxp[2][16];
			for (int w = 0; w < 2; w++) {
				const ShakeOut &o = so[2 * bcm + w];
				for (int k = 0; k < 4; k++) {
					epp[w][0][k] = (int) ((o.ep[0] >> (8 * k)) & 255u);
					epp[w][1][k] = (int) ((o.ep[1] >> (8 * k)) & 255u);
				}
				for (int i = 0; i < 16; i++) idxp[w][i] = (int) ((o.idx >> (4 * i)) & 15u);
			}
			uint64_t blk[2];
			pack_dual_index(mode, bcm % ds.nsel, bcm / ds.nsel, epp, idxp, blk);
			p.dst[gblock] = make_uint4((uint32_t) blk[0], (uint32_t) (blk[0] >> 32), (uint32_t) blk[1], (uint32_t) (blk[1] >> 32));
			p.best_err[gblock] = be;
		}
	}
	AMD_T(5);
}

// A group of WS::kBlocks blocks of a single-index mode (0 .. 3, 7): lane = (block of the group, task of the block)
template <int CLOGS, typename WS>
__device__ __forceinline__ void window_group_single(const AmdParams &p, WS &ws, const Tables &T, const uint32_t *lut2, const uint32_t *lut3, uint32_t group,
																										unsigned lane) {
	constexpr int bpw = WS::kBlocks;
	const int mode = p.mode;
	const ModeInfo mi = mode_info(mode);
	const ShakeParams sp = single_index_shake_params(mode);
	const int subsets = mi.subsets, ntasks = 8 * subsets, wtasks = bpw * ntasks;
	const int sub = (int) lane / ntasks, tk = (int) lane - sub * ntasks;
	const uint32_t block0 = group * (uint32_t) bpw, block = block0 + (uint32_t) sub;
	const bool present = sub < bpw && block < p.n_blocks;
	const bool have = present && p.s.q_top[(size_t) block * 8] != 0xffu; // (0xff: mode not searched for this block)
	if (p.first && present && !have && tk == 0) {
		p.dst[p.block0 + block] = make_uint4(0, 0, 0, 0);
		p.best_err[p.block0 + block] = A7_HUGE;
	}
	if (!__any_sync(FULL, have)) return;
	__syncwarp();
	for (int b = (int) lane >> 4; b < bpw; b += 2) // 16 lanes per block
		if (block0 + b < p.n_blocks) ws.px[b][lane & 15] = fetch_rgba_u8(p.img, block_coord(p, block0 + b), (int) lane & 15);
	if ((int) lane < 8 * bpw && block0 + (lane >> 3) < p.n_blocks) ws.top[lane >> 3][lane & 7] = p.s.q_top[(size_t) (block0 + (lane >> 3)) * 8 + (lane & 7)];
	__syncwarp();
	const bool cube = sp.dim == 3;
	const int wtype = sp.bits[3] % (2 * sp.dim);
	AMD_T0();
	if (have) {
		const int a = tk / subsets, s = tk - a * subsets;
		Task &t = ws.task[lane];
		const uint64_t idx_q = p.s.q_idx[(size_t) block * kMaxTasks + tk];
		build_single_index_task(t, ws.px[sub], subsets, ws.top[sub][a], s, sp, idx_q);
		if (cube) {
			t.best_idx = p.s.c_idx[(size_t) block * kMaxTasks + tk];
			t.err_o = p.s.c_err[(size_t) block * kMaxTasks + tk];
		}
		window_planes_u8(t.d, t.n, t.plane);
	}
	{
		int first = (int) lane;
		if (subsets == 3) { // (uniform; one block per warp) repeats of a texel set take the first task's result
			const int a = (int) lane < ntasks ? (int) lane / subsets : 0, s = (int) lane - a * subsets;
			first = first_task_with_mask(subsets, ws.top[0][a], s, ntasks, lane);
		}
		if ((int) lane < wtasks) {
			ws.same_as[lane] = (uint8_t) first;
			if (!have || first != (int) lane) ws.task[lane].w_active = 0;
		}
	}
	__syncwarp();
	// shake_subset (:709-805): ep_shaker_2_d on the quantiser's indices and, where ep_shaker_d won, again on its indices
	window_phase<CLOGS>(T, ws, lut2, lut3, wtasks, sp.dim, wtype, lane);
	AMD_T(3);
	if (cube) {
		if ((int) lane < wtasks) {
			Task &t = ws.task[lane];
			t.w_active = (have && ws.same_as[lane] == lane && t.err_o < t.w_err_o) ? 1 : 0;
			t.w_index = t.best_idx;
		}
		__syncwarp();
		window_phase<CLOGS>(T, ws, lut2, lut3, wtasks, sp.dim, wtype, lane);
		AMD_T(4);
	}
	ShakeOut o;
	if ((int) lane < wtasks) {
		const Task &t = ws.task[ws.same_as[lane]];
		o.err = t.w_err_o;
		o.idx = t.w_best_idx;
		o.ep[0] = (uint32_t) t.w_best_ep;
		o.ep[1] = (uint32_t) (t.w_best_ep >> 32);
	}
	__syncwarp(); // (so shares its storage with the search results of the phases)
	if ((int) lane < wtasks) ws.so[lane] = o;
	__syncwarp();
#pragma unroll 1
	for (int b = 0; b < bpw; b++) { // (uniform) best attempt of each block, packing
		if (!__shfl_sync(FULL, have ? 1 : 0, b * ntasks)) continue;
		const ShakeOut *so = ws.so + b * ntasks;
		const uint64_t gblock = p.block0 + block0 + (uint32_t) b;
		int ba = 0, store = 0;
		real be = A7_HUGE;
		if (lane == 0) {
			for (int a = 0; a < 8; a++) {
				real e = 0;
				for (int s = 0; s < subsets; s++) e += so[a * subsets + s].err;
				if (e < be) { be = e; ba = a; }
			}
			const real carried = p.first ? A7_HUGE : p.best_err[gblock];
			store = (p.first || be < carried) ? 1 : 0;
		}
		ba = __shfl_sync(FULL, ba, 0);
		store = __shfl_sync(FULL, store, 0);
		if (store) { // (uniform) this mode beats what the earlier modes left in dst
			uint64_t b0, b1;
			pack_single_index_warp(mode, ws.top[b][ba], &so[ba * subsets], lane, b0, b1);
			if (lane == 0) {
				p.dst[gblock] = make_uint4((uint32_t) b0, (uint32_t) (b0 >> 32), (uint32_t) b1, (uint32_t) (b1 >> 32));
				p.best_err[gblock] = be;
			}
		}
	}
	AMD_T(5);
}

// 8 warps per CTA: 2 CTAs (16 warps) fit the 227 KB of shared memory next to one copy of the ramp tables each
constexpr int kWindowWarps = 8, kWindowCtasPerSm = 2;
template <typename WS, int CLOGS> __global__ void __launch_bounds__(kWindowWarps * 32, kWindowCtasPerSm) amd_window_kernel(const AmdParams p) {
	constexpr bool kDual = std::is_same<WS, WindowScratchDual>::value;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	WS *scratch = reinterpret_cast<WS *>(smem_raw);
	__shared__ uint32_t lut2[RampLutShape<2>::kWords], lut3[RampLutShape<3>::kWords];
	if (CLOGS & 4) ramp_lut_fill<2>(lut2, (int) threadIdx.x, kWindowWarps * 32);
	if (CLOGS & 8) ramp_lut_fill<3>(lut3, (int) threadIdx.x, kWindowWarps * 32);
	__syncthreads();
	const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
	WS &ws = scratch[warp];
	const Tables T{p.sp};
	const uint32_t bpw = kDual ? (p.mode == 4 ? 2u : 4u) : (uint32_t) WS::kBlocks;
	const uint32_t groups = (p.n_blocks + bpw - 1) / bpw;
#pragma unroll 1
	for (;;) {
		const uint32_t group = fetch_block(p.next_block, lane);
		if (group >= groups) break;
		if constexpr (kDual) window_group_dual<CLOGS>(p, ws, T, lut2, lut3, group, lane);
		else window_group_single<CLOGS>(p, ws, T, lut2, lut3, group, lane);
	}
}
// (2 CTAs of 8 warps and their tables fit the shared memory of an SM)
constexpr size_t kWindowSmemBudget = (233472 / kWindowCtasPerSm - 1024 - 6144) / kWindowWarps;
static_assert(sizeof(WindowScratch1) <= kWindowSmemBudget && sizeof(WindowScratch2) <= kWindowSmemBudget && sizeof(WindowScratchDual) <= kWindowSmemBudget,
							"window scratch: 2 CTAs per SM");

// =====================================================================================================================
// Mode 6 has ONE task per block (one partition, one subset, 16 index levels: outside the 8-entry ramp words of the window
// kernel): every THREAD takes a block and runs the serial form of the same search (bc7amd_block.cuh, the code the host
// build checks against the reference).  1.4 % of the benchmark step.  (Debug builds can route modes 4 / 5 here too.)
// =====================================================================================================================
__global__ void __launch_bounds__(128) bc7amd_serial_kernel(const AmdParams p, const int mode) {
	const uint32_t block = blockIdx.x * blockDim.x + threadIdx.x;
	if (block >= p.n_blocks) return;
	const Tables T{p.sp};
	const BlockCoord bc = block_coord(p, block);
	float in[64];
#pragma unroll 1
	for (int i = 0; i < 16; i++) {
		const float4 t = fetch_rgba(p.img, bc.gblock, bc.bx, bc.by, bc.slice, i);
		in[i * 4 + 0] = t.x; in[i * 4 + 1] = t.y; in[i * 4 + 2] = t.z; in[i * 4 + 3] = t.w;
	}
	BlockInput B;
	prepare_block(in, p.mode_mask, B);
	const real carried = p.first ? A7_HUGE : p.best_err[bc.gblock];
	real best = carried;
	uint64_t out[2] = {0, 0};
	if (B.mode_mask & p.launch_modes & (1u << mode)) {
		uint64_t tmp[2];
		const real e = (mode_info(mode).alpha != 2) ? compress_single_index<true>(T, B, mode, tmp) : compress_dual_index<true>(T, B, mode, tmp);
		if (e < best) { best = e; out[0] = tmp[0]; out[1] = tmp[1]; }
	}
	if (p.first || best < carried) {
		p.dst[bc.gblock] = make_uint4((uint32_t) out[0], (uint32_t) (out[0] >> 32), (uint32_t) out[1], (uint32_t) (out[1] >> 32));
		p.best_err[bc.gblock] = best;
	}
}

// =====================================================================================================================
// Float sources (RGBA16F / RGBA32F / float block API): the generic FP64 shakers, one warp per block, lanes = tasks.
// =====================================================================================================================
struct FloatScratch {
	float in[64];
	BlockInput B;
	real serr[64][3];
	uint64_t qidx[64][3];
	real perr[64];
	int top[8];
	ShakeOut so[kMaxTasks];
	real qs[2][16][32];
	uint64_t blk[2];
	real blk_err;
};

__global__ void __launch_bounds__(kWarps * 32, 2) bc7amd_float_kernel(const AmdParams p) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	FloatScratch *scratch = reinterpret_cast<FloatScratch *>(smem_raw);
	const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
	const uint32_t block = blockIdx.x * kWarps + warp;
	if (block >= p.n_blocks) return; // whole warp
	FloatScratch &ws = scratch[warp];
	const Tables T{p.sp};
	const BlockCoord bc = block_coord(p, block);
	if (lane < 16) {
		const float4 t = fetch_rgba(p.img, bc.gblock, bc.bx, bc.by, bc.slice, (int) lane);
		ws.in[lane * 4 + 0] = t.x; ws.in[lane * 4 + 1] = t.y; ws.in[lane * 4 + 2] = t.z; ws.in[lane * 4 + 3] = t.w;
	}
	__syncwarp();
	if (lane == 0) prepare_block(ws.in, p.mode_mask, ws.B);
	__syncwarp();
	const uint32_t mask = ws.B.mode_mask & p.launch_modes;
	if (mask == 0 && !p.first) return; // whole warp
	const real carried = p.first ? A7_HUGE : p.best_err[bc.gblock];
	real best = carried;
	uint64_t out0 = 0, out1 = 0;
	for (int vi = 0; vi < 8; vi++) {
		const int mode = mode_visit_order(vi);
		if (!(mask & (1u << mode))) continue;
		const ModeInfo mi = mode_info(mode);
		if (mi.alpha != 2) {
			const ShakeParams sp = single_index_shake_params(mode);
			const int nparts = 1 << mi.partition_bits, subsets = mi.subsets;
			const uint8_t *qorder = subsets > 1 ? quantise_order(subsets, nparts) : nullptr;
			QuantIO io;
			io.px = &ws.B.pxc[0][0];
			io.chan = 0xE4u;
			io.proj = &ws.qs[0][0][lane];
			io.dev = &ws.qs[1][0][lane];
			io.stride = 32;
			for (int tt = (int) lane; tt < nparts * subsets; tt += 32) {
				const int t = qorder ? qorder[tt] : tt;
				const int part = t / subsets, s = t - part * subsets;
				uint32_t smask = 0;
				for (int i = 0; i < 16; i++) smask |= (subset_of(subsets, part, i) == s ? 1u : 0u) << i;
				int n;
				io.texels = texels_of_mask(smask, n);
				uint64_t packed = 0;
				ws.serr[part][s] = n ? quantise_subset(io, n, sp.clusters, sp.dim, packed) : 0;
				ws.qidx[part][s] = packed;
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
			const int ntasks = attempts * subsets;
			if ((int) lane < ntasks) {
				real sub[kMaxEntries][4];
				int n = 0, idx[kMaxEntries], ep[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
				const int a = (int) lane / subsets, s = (int) lane - a * subsets;
				gather_subset(ws.B, subsets, ws.top[a], s, sp.dim, sub, n);
				unpack_idx(ws.qidx[ws.top[a]][s], idx, n);
				ShakeOut o;
				o.err = shake_subset(T, sp, sub, n, idx, ep);
				o.idx = pack_idx(idx, n);
				o.ep[0] = pack_ep(ep[0]);
				o.ep[1] = pack_ep(ep[1]);
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
			const int ntasks = combos * 2;
			uint64_t qpacked = 0;
			if ((int) lane < ntasks) {
				const int combo = (int) lane >> 1, which = (int) lane & 1;
				const int rot = combo / nsel, isel = combo - rot * nsel;
				const uint32_t c0 = (uint32_t) rotation_channel(rot, 0), c1 = (uint32_t) rotation_channel(rot, 1),
											 c2 = (uint32_t) rotation_channel(rot, 2), c3 = (uint32_t) rotation_channel(rot, 3);
				QuantIO io;
				io.px = &ws.B.pxc[0][0];
				io.texels = 0xFEDCBA9876543210ull;
				io.chan = which == 0 ? (c1 | (c2 << 2) | (c3 << 4)) : (c0 | (c0 << 2) | (c0 << 4));
				io.proj = &ws.qs[0][0][lane];
				io.dev = &ws.qs[1][0][lane];
				io.stride = 32;
				const int qb = which == 0 ? (isel ? mi.index_bits1 : mi.index_bits0) : (isel ? mi.index_bits0 : mi.index_bits1);
				quantise_subset(io, 16, 1 << qb, 3, qpacked);
			}
			__syncwarp();
			if ((int) lane < ntasks) {
				real blkv[16][4];
				int idx[16], ep[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
				const int combo = (int) lane >> 1, which = (int) lane & 1;
				const int rot = combo / nsel, isel = combo - rot * nsel;
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
				unpack_idx(qpacked, idx, 16);
				const int bits[4] = {cb, cb, cb, 6 * cb};
				ShakeOut o;
				shake_cube(T, blkv, 16, idx, (1 << ib) - 1, bits, CART);
				o.err = shake_window(T, blkv, 16, idx, ep, 6, (1 << ib) - 1, bits[3], 3);
				o.idx = pack_idx(idx, 16);
				o.ep[0] = pack_ep(ep[0]);
				o.ep[1] = pack_ep(ep[1]);
				ws.so[lane] = o;
			}
			__syncwarp();
			if (lane == 0) {
				real be = A7_HUGE;
				int bcm = 0;
				for (int c = 0; c < combos; c++) {
					real e = 0;
					e += ws.so[2 * c].err;
					e += ws.so[2 * c + 1].err / 3.;
					if (e < be) { be = e; bcm = c; }
				}
				int epp[2][2][4], idxp[2][16];
				for (int w = 0; w < 2; w++) {
					const ShakeOut &o = ws.so[2 * bcm + w];
					for (int k = 0; k < 4; k++) {
						epp[w][0][k] = (int) ((o.ep[0] >> (8 * k)) & 255u);
						epp[w][1][k] = (int) ((o.ep[1] >> (8 * k)) & 255u);
					}
					for (int i = 0; i < 16; i++) idxp[w][i] = (int) ((o.idx >> (4 * i)) & 15u);
				}
				uint64_t blk[2];
				pack_dual_index(mode, bcm % nsel, bcm / nsel, epp, idxp, blk);
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
	if (lane == 0 && (p.first || best < carried)) {
		p.dst[bc.gblock] = make_uint4((uint32_t) out0, (uint32_t) (out0 >> 32), (uint32_t) out1, (uint32_t) (out1 >> 32));
		p.best_err[bc.gblock] = best;
	}
}

} // namespace

#ifdef B200IC_AMD_TIMING
extern "C" __attribute__((visibility("default"))) void b200ic_amd_serial_modes(int mask) { g_serial_modes = (mask & 0x30) | 0x40; }
extern "C" __attribute__((visibility("default"))) int b200ic_amd_timing(unsigned long long *out, int reset) {
	cudaDeviceSynchronize();
	if (out) cudaMemcpyFromSymbol(out, g_amd_timing, sizeof(g_amd_timing));
	if (reset) {
		unsigned long long z[16] = {};
		cudaMemcpyToSymbol(g_amd_timing, z, sizeof(z));
	}
	return 0;
}
#endif

// ---- per-kernel timing for bench.py's roofline (b200ic_profile / b200ic_profile_read, include/b200ic.h): with profiling
// on, every kernel of the AMD BC7 pipeline is bracketed by CUDA events on the launching stream
namespace {
struct ProfEvent {
	int mode, kind; // kind: 0 quantise, 1 cube, 2 window, 3 thread-per-block / float kernel
	cudaEvent_t a, b;
};
std::mutex g_prof_mu;
std::vector<ProfEvent> g_prof_events;
bool g_prof_on = false;
struct ProfScope {
	cudaStream_t st;
	ProfEvent ev;
	bool on;
	ProfScope(cudaStream_t s, int mode, int kind) : st(s), on(g_prof_on) {
		if (!on) return;
		ev.mode = mode;
		ev.kind = kind;
		cudaEventCreate(&ev.a);
		cudaEventCreate(&ev.b);
		cudaEventRecord(ev.a, st);
	}
	~ProfScope() {
		if (!on) return;
		cudaEventRecord(ev.b, st);
		std::lock_guard<std::mutex> lock(g_prof_mu);
		g_prof_events.push_back(ev);
	}
};
} // namespace
extern "C" __attribute__((visibility("default"))) void b200ic_profile(int enable) { g_prof_on = enable != 0; }
extern "C" __attribute__((visibility("default"))) int b200ic_profile_read(double *ms, uint64_t *launches) {
	std::lock_guard<std::mutex> lock(g_prof_mu);
	for (int i = 0; i < 32; i++) {
		if (ms) ms[i] = 0;
		if (launches) launches[i] = 0;
	}
	int n = 0;
	for (ProfEvent &e : g_prof_events) {
		float t = 0;
		if (cudaEventSynchronize(e.b) == cudaSuccess && cudaEventElapsedTime(&t, e.a, e.b) == cudaSuccess) {
			if (ms) ms[e.mode * 4 + e.kind] += (double) t;
			if (launches) launches[e.mode * 4 + e.kind]++;
			n++;
		}
		cudaEventDestroy(e.a);
		cudaEventDestroy(e.b);
	}
	g_prof_events.clear();
	return n;
}

// The window kernel of a mode: sets its shared-memory attribute (p == nullptr) or launches it
namespace {
template <typename WS, int CLOGS> cudaError_t window_kernel_do(const AmdParams *p, unsigned grid, cudaStream_t stream) {
	if (!p) return cudaFuncSetAttribute(amd_window_kernel<WS, CLOGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (kWindowWarps * sizeof(WS)));
	amd_window_kernel<WS, CLOGS><<<grid, kWindowWarps * 32, kWindowWarps * sizeof(WS), stream>>>(*p);
	return cudaSuccess;
}
cudaError_t window_kernel_launch(int mode, const AmdParams *p, unsigned grid, cudaStream_t stream) {
	switch (mode) {
	case 0: return window_kernel_do<WindowScratch1, mode_clogs(0)>(p, grid, stream);
	case 1: return window_kernel_do<WindowScratch2, mode_clogs(1)>(p, grid, stream);
	case 2: return window_kernel_do<WindowScratch1, mode_clogs(2)>(p, grid, stream);
	case 3: return window_kernel_do<WindowScratch2, mode_clogs(3)>(p, grid, stream);
	case 4: return window_kernel_do<WindowScratchDual, mode_clogs(4)>(p, grid, stream);
	case 5: return window_kernel_do<WindowScratchDual, mode_clogs(5)>(p, grid, stream);
	case 7: return window_kernel_do<WindowScratch1, mode_clogs(7)>(p, grid, stream);
	default: return cudaErrorInvalidValue;
	}
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
	{
		uint8_t order[48 + 192 + 128], first[48 + 192];
		int pos = 0;
		for (int table = 0; table < 3; table++) {
			const int subsets = table < 2 ? 3 : 2, nparts = table == 0 ? 16 : 64, base = pos;
			auto mask_of = [&](int t) {
				uint32_t m = 0;
				for (int i = 0; i < 16; i++) m |= (subset_of(subsets, t / subsets, i) == t % subsets ? 1u : 0u) << i;
				return m;
			};
			auto repeats = [&](int t) { // an earlier (partition, subset) of the table has the same texels
				for (int u = 0; u < t; u++)
					if (mask_of(u) == mask_of(t)) return true;
				return false;
			};
			for (int rep = 0; rep < 2; rep++) // the distinct subsets first, then the repeats; each part by size, descending
				for (int size = 16; size >= 0; size--)
					for (int t = 0; t < nparts * subsets; t++) {
						int n = 0;
						for (int i = 0; i < 16; i++) n += subset_of(subsets, t / subsets, i) == t % subsets ? 1 : 0;
						if (n == size && (repeats(t) ? 1 : 0) == rep) order[pos++] = (uint8_t) t;
					}
			if (table < 2)
				for (int i = base; i < pos; i++) {
					int f = i;
					for (int j = base; j < i; j++)
						if (mask_of(order[j]) == mask_of(order[i])) { f = j; break; }
					first[i] = (uint8_t) (f - base);
				}
		}
		e = cudaMemcpyToSymbol(c_qfirst, first, sizeof(first));
		if (e != cudaSuccess) return e;
		e = cudaMemcpyToSymbol(c_qorder, order, sizeof(order));
		if (e != cudaSuccess) return e;
	}
	{
		double step[17 * 16] = {};
		for (int n = 1; n <= 16; n++)
			for (int i = 0; i < 16; i++) step[n * 16 + i] = (2. * (double) i + 1 - (double) n) / 2. / (double) n;
		e = cudaMemcpyToSymbol(c_simplex_step, step, sizeof(step));
		if (e != cudaSuccess) return e;
	}
	e = cudaFuncSetAttribute(amd_quant_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (kWarps * sizeof(QuantScratch)));
	if (e != cudaSuccess) return e;
	e = cudaFuncSetAttribute(amd_quant_dual_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (kWarps * sizeof(DualQuantScratch)));
	if (e != cudaSuccess) return e;
	for (int mode = 0; mode < 8; mode++) {
		if (mode == 6) continue;
		e = window_kernel_launch(mode, nullptr, 0, nullptr);
		if (e != cudaSuccess) return e;
	}
	e = cudaFuncSetAttribute(bc7amd_float_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (kWarps * sizeof(FloatScratch)));
	if (e != cudaSuccess) return e;
	g_sp_table_host[dev] = d;
	return cudaSuccess;
}

cudaError_t launch_bc7amd(const SrcImage &img, const b200ic_opts &opts, void *dst, cudaStream_t stream) {
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess) return e;
	if (dev < 0 || dev >= 16 || !g_sp_table_host[dev]) return cudaErrorInitializationError;
	const uint64_t total_blocks = (uint64_t) img.blocks_x * img.blocks_y * img.slices;
	if (total_blocks == 0) return cudaSuccess;
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	AmdParams p;
	p.img = img;
	p.dst = static_cast<uint4 *>(dst);
	p.sp = g_sp_table_host[dev];
	p.mode_mask = (uint32_t) opts.amd_mode_mask & 0xffu;
	p.mode = 0;
	p.next_block = nullptr;
	p.s = AmdScratch{nullptr, nullptr, nullptr, nullptr};
	// 8-bit sources: every component is an exact integer, the exact INT32 shakers apply (bc7amd_int.cuh)
	const bool u8 = img.format == B200IC_FMT_R8 || img.format == B200IC_FMT_RG8 || img.format == B200IC_FMT_RGB8 ||
									img.format == B200IC_FMT_RGB8_SRGB || img.format == B200IC_FMT_RGBA8 || img.format == B200IC_FMT_RGBA8_SRGB ||
									img.format == B200IC_FMT_BLOCKS_RGBA8;
	const uint32_t chunk = (uint32_t) (total_blocks < kChunkBlocks ? total_blocks : kChunkBlocks);
	unsigned char *scratch = nullptr;
	const size_t per_block = sizeof(uint64_t) * kMaxTasks * 2 + sizeof(real) * kMaxTasks + 8;
	constexpr size_t kCounterBytes = 256; // work counters of the persistent launches of one chunk (<= 15 launches)
	e = cudaMallocAsync((void **) &p.best_err, total_blocks * sizeof(real), stream);
	if (e != cudaSuccess) return e;
	if (u8) {
		e = cudaMallocAsync((void **) &scratch, kCounterBytes + (size_t) chunk * per_block, stream);
		if (e != cudaSuccess) {
			cudaFreeAsync(p.best_err, stream);
			return e;
		}
		p.s.q_idx = reinterpret_cast<uint64_t *>(scratch + kCounterBytes);
		p.s.c_idx = p.s.q_idx + (size_t) chunk * kMaxTasks;
		p.s.c_err = reinterpret_cast<real *>(p.s.c_idx + (size_t) chunk * kMaxTasks);
		p.s.q_top = reinterpret_cast<uint8_t *>(p.s.c_err + (size_t) chunk * kMaxTasks);
	}
	const uint32_t user = p.mode_mask ? p.mode_mask : 0xCFu;
	int launches = 0;
	for (uint64_t b0 = 0; b0 < total_blocks && e == cudaSuccess; b0 += chunk) {
		p.block0 = b0;
		p.n_blocks = (uint32_t) (total_blocks - b0 < chunk ? total_blocks - b0 : chunk);
		const unsigned warp_grid = (p.n_blocks + kWarps - 1) / kWarps;
		const unsigned window_ctas = (p.n_blocks + kWindowWarps - 1) / kWindowWarps;
		const unsigned window_grid = window_ctas < (unsigned) (sms * kWindowCtasPerSm) ? window_ctas : (unsigned) (sms * kWindowCtasPerSm);
		const unsigned cube_grid = warp_grid < (unsigned) (sms * kCubeCtasPerSm) ? warp_grid : (unsigned) (sms * kCubeCtasPerSm);
		int passes = 0;
		uint32_t *counter = reinterpret_cast<uint32_t *>(scratch);
		if (u8) {
			e = cudaMemsetAsync(scratch, 0, kCounterBytes, stream);
			if (e != cudaSuccess) break;
		}
		for (int vi = 0; vi < 8; vi++) {
			const int mode = mode_visit_order(vi);
			if (!u8) { // one fused launch: the generic kernel walks the modes itself
				if (vi) break;
				p.launch_modes = 0xFFu;
				p.first = 1;
				ProfScope ps(stream, 0, 3);
				bc7amd_float_kernel<<<warp_grid, kWarps * 32, kWarps * sizeof(FloatScratch), stream>>>(p);
				launches++;
				continue;
			}
			if (!(user & (1u << mode)) && passes) continue; // (the first pass always runs: it initialises dst)
			p.launch_modes = 1u << mode;
			p.mode = mode;
			p.first = passes == 0;
			if (mode == 6 || ((g_serial_modes >> mode) & 1)) {
				ProfScope ps(stream, mode, 3);
				bc7amd_serial_kernel<<<(p.n_blocks + 127) / 128, 128, 0, stream>>>(p, mode);
				launches++;
			} else {
				{
					ProfScope ps(stream, mode, 0);
					if (mode == 4 || mode == 5) {
						const unsigned per_cta = kWarps * (mode == 4 ? 2u : 4u);
						amd_quant_dual_kernel<<<(p.n_blocks + per_cta - 1) / per_cta, kWarps * 32, kWarps * sizeof(DualQuantScratch), stream>>>(p);
					} else {
						amd_quant_kernel<<<warp_grid, kWarps * 32, kWarps * sizeof(QuantScratch), stream>>>(p);
					}
				}
				if (mode != 7) {
					ProfScope ps(stream, mode, 1);
					p.next_block = counter++;
					// (second-pass pruning pays where an item has 256 corners: mode 0; mode 2 with 64 measured 2x slower)
					if (mode == 0) amd_cube_kernel<mode_clogs(0), true><<<cube_grid, kWarps * 32, 0, stream>>>(p);
					else if (mode == 1) amd_cube_kernel<mode_clogs(1), false><<<cube_grid, kWarps * 32, 0, stream>>>(p);
					else if (mode == 4) amd_cube_kernel<mode_clogs(4), false><<<cube_grid, kWarps * 32, 0, stream>>>(p);
					else amd_cube_kernel<4, false><<<cube_grid, kWarps * 32, 0, stream>>>(p);
				}
				{
					ProfScope ps(stream, mode, 2);
					p.next_block = counter++;
					e = window_kernel_launch(mode, &p, window_grid, stream);
					if (e != cudaSuccess) break;
				}
				launches += mode != 7 ? 3 : 2;
			}
			passes++;
		}
	}
	count_launches(launches - 1);
	if (e == cudaSuccess) e = cudaGetLastError();
	if (scratch) cudaFreeAsync(scratch, stream);
	cudaFreeAsync(p.best_err, stream);
	return e;
}

} // namespace b200ic
