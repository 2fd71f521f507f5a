// bc7amd.cu -- sm_100a kernel for the AMD-Compressonator-compatible BC7 path (all eight modes, quality 1).
//
// Replaces the image loop of reference src/amd_bc7_compressor.cpp:25-80 (gather via block_utils.cpp:7-41) and the
// BC7BlockEncoder::CompressBlock tree (src/amd_bc7_body.cpp:1289-1465); the search itself is bc7amd_core.cuh.
//
// Mapping: one warp per 4x4 block, candidates -> lanes.
//   quantise phase : one (partition, subset) of the current mode per lane (16..192 independent optQuantAnD problems)
//   selection      : rank of every partition's error computed in parallel (stable: ties keep partition order),
//                    the 8 lowest are shaken
//   shake phase    : "tasks" = (attempt, subset) or, for the dual-index modes, (rotation, selection, vector|scalar).
//                    ep_shaker_d (82 % of the reference's time) is cut into work items (task, (q,p) re-indexing,
//                    z-slice of the endpoint cube); the items of all tasks are laid out sorted by subset size and
//                    dealt to the lanes round-robin, so that the lanes of one round run the same trip counts.
//                    Every item is evaluated with the exact INT32 form of bc7amd_int.cuh; the per-task winner is
//                    the minimum (error, scan position) key = the reference's first strict minimum.
//                    ep_shaker_2_d chains run one task per lane.
//   winners        : first strict minimum in the reference's scan order, by one lane, then broadcast
// Arithmetic: FP64 where the reference is FP64 and the operands are not integers (quantiser, endpoint fit), exact
// INT32 elsewhere; the blocks are bit-identical to the reference's wherever its qsort tie order does not matter.
// Float sources take the generic FP64 shakers (template parameter U8 = false).
#include "common.cuh"
#include "kernels.h"
#include "bc7amd_block.cuh"
#include <stdlib.h>

namespace b200ic {

namespace {

using namespace amd7;

constexpr int kWarps = 4;
constexpr int kMaxTasks = 24;   // single-index: 8 attempts x 3 subsets; dual-index: 8 combos x 2
constexpr int kItemBatch = 160; // work items evaluated between two per-task reductions

uint32_t *g_sp_table_host[16] = {};

// Quantise-phase task order: (partition, subset) pairs of each partition table sorted by subset size (descending),
// so that the 32 lanes of one round run optQuantAnD problems of (nearly) the same size. Pure scheduling data.
__constant__ uint8_t c_qorder[48 + 192 + 128];
__device__ __forceinline__ const uint8_t *quantise_order(int subsets, int nparts) {
	return subsets == 3 ? (nparts == 16 ? c_qorder : c_qorder + 48) : c_qorder + 240;
}

// Optional phase timing (build with -DB200IC_AMD_TIMING, read with b200ic_amd_timing): clock64 deltas of lane 0 summed
// per phase over all warps. Slots: 0 quantise, 1 rank+task setup, 2 cube, 3 window, 4 window (2nd), 5 pick+pack,
// 6 cube items evaluated, 7 cube rounds of 32 lanes, 8 window items, 9 window rounds, 10 cube passes
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

// One ep_shaker_d problem (u8 path)
struct CubeTask {
	uint32_t d[16];     // packed texels
	uint64_t idx_q;     // quantiser indices
	uint64_t cur;       // collapsed indices of the running pass
	uint64_t best_idx;  // index_io of the reference
	uint64_t pass_key;  // best (err << 16 | qp << 8 | lattice << 6 | gray) of the running pass
	uint64_t pass_idx;
	real err_o;
	real mean[4];
	// ep_shaker_2_d state
	uint64_t w_index;   // running indices (uncollapsed)
	uint64_t w_best_idx, w_best_ep;
	real w_err_o;
	uint16_t item_base, item_count;
	uint8_t n, clog, bits, type, Mi, done, all_same, dim;
	uint8_t w_bits_total, w_size, w_tries, w_active;
};

// One batch (<= 32 items, one per lane) of the pruned cube walk, one lattice at a time (cube_batch)
constexpr int kCubeBatch = 32;
struct CubeBatch {
	uint64_t tab[kCubeBatch][12];        // this lattice's ramp tables per item: [channel * 4 + endpoint combination]
	uint64_t mask[kCubeBatch];           // this lattice's surviving corners per item
	unsigned long long best[kCubeBatch]; // running best per item: key << 32 | lattice << 6 | corner
	uint16_t unit_base[kCubeBatch + 1];
	uint8_t task_of[kCubeBatch];
};

struct WarpScratch {
	float in[64];
	BlockInput B;
	union {
		struct { // quantise phase -> ranking -> task set-up
			real serr[64][3];
			uint64_t qidx[64][3]; // quantiser indices of every (partition, subset) of the running mode
			real perr[64];
		};
		CubeBatch cb; // cube phase (the quantiser results are dead by then)
	};
	int top[8];
	ShakeOut so[kMaxTasks];
	union {
		struct { // shake phases
			CubeTask task[kMaxTasks];
			uint64_t item_key[kItemBatch];
			uint64_t item_idx[kItemBatch];
		};
		real qs[2][16][32]; // quantise phases: the two lane-strided FP64 work arrays of QuantIO (element k of lane l at [k][l])
	};
	uint8_t order[kMaxTasks];
	uint64_t blk[2];
	real blk_err;
};

struct AmdParams {
	SrcImage img;
	uint4 *dst;
	uint64_t n_blocks;
	const uint32_t *sp;
	uint32_t mode_mask;    // the caller's ModeMask (input of the reference's mode filter)
	uint32_t launch_modes; // modes searched by THIS launch (one launch per mode, see launch_bc7amd)
	real *best_err;        // per block: error of the block currently in dst (carried from launch to launch)
	int first;             // first launch of the sequence: nothing to compare with
	int zsplit_single, zsplit_dual; // experiment knobs: 0 = auto / built-in choice
	int split_n;                    // cube items of subsets with >= split_n texels are shared by two lanes (0 = never)
	int prune;                      // cube walk: 0 exhaustive, 1 per-lane branch-and-bound, 2 two-phase branch-and-bound (cube_batch)
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

// Start (or restart) a pass of ep_shaker_d for one task: collapse the indices, handle the single-index case.
__device__ __noinline__ void cube_begin_pass(const Tables &T, CubeTask &t, uint64_t from) {
	int index[kMaxEntries];
	unpack_idx(from, index, t.n);
	const int Mi = collapse_indices(index, t.n);
	if (Mi == 0) {
		U8Subset S;
		for (int i = 0; i < t.n; i++) S.d[i] = t.d[i];
		S.n = t.n;
		S.all_same = t.all_same != 0;
		for (int j = 0; j < 4; j++) S.mean[j] = t.mean[j];
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

__device__ __forceinline__ int item_owner(const WarpScratch &ws, int ntasks, int it) {
	int ti = 0;
	for (int r = 0; r < ntasks; r++) {
		const int cand = ws.order[r];
		const int base = ws.task[cand].item_base;
		if (it >= base && it < base + ws.task[cand].item_count) ti = cand;
	}
	return ti;
}

// ---- exact branch-and-bound of the cube walk, organised for SIMT ------------------------------------------------
// (bound and proof: cube_search_pruned_u8 in bc7amd_int.cuh.)  A batch is <= 32 items, one per lane.  Per lattice:
//   A  every lane: the ramp tables of its item (to shared memory), the twelve per-channel bounds, the exact error of
//      the corner with the smallest bound, and the mask of corners whose bound does not exceed the item's best so far
//      (over the lattices already finished: later lattices are pruned almost completely);
//   B  the surviving corners of ALL items form one list that is cut into 32 equal runs -- every lane evaluates the
//      same number of corners whatever its own item pruned; results meet in a 64-bit atomicMin per item.
// Only corners that can neither win nor tie are skipped, so the keys are those of the exhaustive walk.
template <int CLOG>
__device__ __forceinline__ void cube_batch_t(WarpScratch &ws, int ntasks, int b0, int b1, unsigned lane) {
	constexpr int C = 1 << CLOG;
	CubeBatch &cb = ws.cb;
	const int it = b0 + (int) lane;
	const bool have = it < b1;
	int qp = 0, ti = 0, nl = 0, use_par = 0, bcc = 0;
	int fl[2][3][2];
	int bits[3] = {0, 0, 0};
	if (have) {
		ti = item_owner(ws, ntasks, it);
		const CubeTask &t = ws.task[ti];
		qp = it - t.item_base;
		int q, p;
		qp_decode(qp, t.Mi, C - 1, q, p);
		cb.task_of[lane] = (uint8_t) ti;
		ClusterAcc<CLOG> cs;
		cluster_acc<CLOG>(t.d, t.n, t.cur, q, p, cs);
		real epa[2][4];
		fit_endpoints_acc<CLOG>(cs, 3, epa);
		bits[0] = bits[1] = bits[2] = t.bits;
		use_par = (t.type == BCC || t.type == SAME_PAR) ? 1 : 0;
		bcc = (t.type == BCC) ? 1 : 0;
		cube_floors(epa, bits, use_par, fl);
		nl = (use_par + 1) * (bcc + 1);
	}
	const int nl_max = __reduce_max_sync(FULL, nl);
	uint32_t best_key = 0xffffffffu, best_code = 0xffffffffu;
	uint32_t win_pal[C];
#pragma unroll
	for (int c = 0; c < C; c++) win_pal[c] = 0;
#pragma unroll 1
	for (int l = 0; l < nl_max; l++) {
		// ---- phase A
		uint64_t mask = 0;
		if (have && l < nl) {
			const CubeTask &t = ws.task[ti];
			const int odd = l / (bcc + 1), flip = l - odd * (bcc + 1);
			uint32_t lb[12], ep_unused[3];
			cube_lattice_setup<CLOG>(t.d, t.n, bits, fl, use_par, odd, flip, cb.tab[lane], lb, ep_unused);
			const int sc = cube_seed_corner(lb);
			if (cube_corner_bound(lb, sc) <= (best_key >> 8)) {
				const uint32_t before = best_key;
				cube_corner_u8<CLOG>(t.d, t.n, reinterpret_cast<const uint64_t(*)[4]>(cb.tab[lane]), sc & 3, (sc >> 2) & 3, sc >> 4, l, best_key, win_pal);
				if (best_key != before) best_code = (uint32_t) ((l << 6) | sc);
			}
			mask = cube_survivors(lb, best_key >> 8) & ~(1ull << sc);
		}
		cb.mask[lane] = mask;
		cb.best[lane] = ((unsigned long long) best_key << 32) | best_code;
		const int mine = __popcll(mask);
		int incl = mine;
		for (int dlt = 1; dlt < 32; dlt <<= 1) {
			const int v = __shfl_up_sync(FULL, incl, dlt);
			if ((int) lane >= dlt) incl += v;
		}
		const int units = __shfl_sync(FULL, incl, 31);
		cb.unit_base[lane] = (uint16_t) (incl - mine);
		if (lane == 31) cb.unit_base[kCubeBatch] = (uint16_t) units;
		__syncwarp();
		// ---- phase B
		const int chunk = (units + 31) >> 5;
		const int u0 = (int) lane * chunk, u1 = min(units, u0 + chunk);
		if (u0 < u1) {
			int lo = 0, hi = kCubeBatch;
			while (hi - lo > 1) {
				const int mid = (lo + hi) >> 1;
				if ((int) cb.unit_base[mid] <= u0) lo = mid;
				else hi = mid;
			}
			int item = lo;
			uint64_t rem = cb.mask[item];
			for (int k = u0 - (int) cb.unit_base[item]; k > 0; k--) rem &= rem - 1;
			for (int u = u0; u < u1; u++) {
				while (rem == 0) {
					item++;
					rem = cb.mask[item];
				}
				const int corner = __ffsll((long long) rem) - 1;
				rem &= rem - 1;
				const CubeTask &t = ws.task[cb.task_of[item]];
				uint32_t key = 0xffffffffu, pal_unused[C];
				cube_corner_u8<CLOG>(t.d, t.n, reinterpret_cast<const uint64_t(*)[4]>(cb.tab[item]), corner & 3, (corner >> 2) & 3, corner >> 4, l, key,
														 pal_unused);
				atomicMin(&cb.best[item], ((unsigned long long) key << 32) | (unsigned long long) ((l << 6) | corner));
			}
		}
		__syncwarp();
		// ---- a corner of this lattice won: keep its palette (the tables are overwritten by the next lattice)
		if (have) {
			const unsigned long long v = cb.best[lane];
			const uint32_t code = (uint32_t) v;
			if ((uint32_t) (v >> 32) != best_key || code != best_code) {
				best_key = (uint32_t) (v >> 32);
				best_code = code;
				const int corner = (int) (code & 63u);
				const uint64_t t0 = cb.tab[lane][corner & 3], t1 = cb.tab[lane][4 + ((corner >> 2) & 3)], t2 = cb.tab[lane][8 + (corner >> 4)];
#pragma unroll
				for (int c = 0; c < C; c++) win_pal[c] = byte_of(t0, c) | (byte_of(t1, c) << 8) | (byte_of(t2, c) << 16);
			}
		}
		__syncwarp();
	}
	if (have) {
		const CubeTask &t = ws.task[ti];
		ws.item_key[it - b0] = ((uint64_t) (best_key >> 8) << 16) | ((uint64_t) qp << 8) | (uint64_t) (best_key & 255u);
		ws.item_idx[it - b0] = palette_indices_u8<CLOG>(t.d, t.n, win_pal);
	}
	__syncwarp();
}
__device__ __noinline__ void cube_batch(WarpScratch &ws, int ntasks, int clog, int b0, int b1, unsigned lane) {
	if (clog == 2) cube_batch_t<2>(ws, ntasks, b0, b1, lane);
	else cube_batch_t<3>(ws, ntasks, b0, b1, lane);
}

// ep_shaker_d for all tasks of the warp (u8 path). On return task[i].err_o / best_idx hold its result.
__device__ __noinline__ void cube_phase(const Tables &T, WarpScratch &ws, int ntasks, int zsplit_in, unsigned lane, int prune, int split_n) {
	int zsplit = zsplit_in > 0 ? zsplit_in : 1;
	if ((int) lane < ntasks) {
		CubeTask &t = ws.task[lane];
		t.err_o = A7_HUGE;
		t.best_idx = t.idx_q;
		t.done = 0;
		cube_begin_pass(T, t, t.idx_q);
	}
	__syncwarp();
	for (int pass = 0; pass < 2; pass++) {
		// ---- lay the items out: tasks sorted by (clog, n) descending so that the lanes of a round agree on trip counts
		int count = 0, sortkey = -1, split = 1;
		if ((int) lane < ntasks && !ws.task[lane].done) {
			const CubeTask &t = ws.task[lane];
			count = qp_count(t.Mi, (1 << t.clog) - 1);
			sortkey = t.clog * 32 + t.n;
			split = (split_n > 0 && t.n >= split_n) ? 2 : 1;
		}
		if (zsplit_in <= 0) {
			// cut every item into z-slices of the endpoint cube so that the rounds of 32 lanes are as full as possible:
			// cost(z) = rounds(z) * (4 / z) quarter-cubes; ties go to the coarser split (the endpoint fit is per item)
			int tot1 = count * split;
			for (int dlt = 16; dlt > 0; dlt >>= 1) tot1 += __shfl_xor_sync(FULL, tot1, dlt);
			const int c1 = ((tot1 + 31) >> 5) * 4, c2 = ((2 * tot1 + 31) >> 5) * 2, c4 = (4 * tot1 + 31) >> 5;
			zsplit = c4 < c2 ? (c4 < c1 ? 4 : 1) : (c2 < c1 ? 2 : 1);
		}
		// an item of a big subset is shared by two ADJACENT lanes, each summing half of the texels: the tasks are laid out
		// by descending size, so the split ones come first and every pair starts on an even item (batches and rounds
		// are even too)
		if (prune == 2) { zsplit = 1; split = 1; } // corners, not z-slices or texel halves, are the grain of the pruned walk
		count *= zsplit * split;
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
		if (total == 0) break;
		AMD_COUNT(6, total);
		AMD_COUNT(7, (total + 31) / 32);
		AMD_COUNT(10, 1);
		const int batch = prune == 2 ? kCubeBatch : kItemBatch;
		for (int b0 = 0; b0 < total; b0 += batch) {
			const int b1 = min(total, b0 + batch);
			if (prune == 2) cube_batch(ws, ntasks, ws.task[0].clog, b0, b1, lane);
			else for (int it = b0 + (int) lane; it < b1; it += 32) {
				int ti = 0;
				for (int r = 0; r < ntasks; r++) {
					const int cand = ws.order[r];
					const int base = ws.task[cand].item_base;
					if (it >= base && it < base + ws.task[cand].item_count) ti = cand;
				}
				const CubeTask &t = ws.task[ti];
				const int local = it - t.item_base;
				const int hs = (split_n > 0 && t.n >= split_n) ? 2 : 1;
				const int qp = local / (zsplit * hs), rest = local - qp * (zsplit * hs);
				const int zp = rest / hs, half = hs == 2 ? rest - zp * hs : -1;
				int q, p;
				qp_decode(qp, t.Mi, (1 << t.clog) - 1, q, p);
				const int bits[3] = {t.bits, t.bits, t.bits};
				const int zn = 4 / zsplit;
				uint32_t key;
				uint64_t idx;
				const unsigned pm = split_n > 0 ? __activemask() : 0u;
				if (t.clog == 2) cube_item_u8<2>(t.d, t.n, t.cur, q, p, bits, t.type, zp * zn, zp * zn + zn, key, idx, prune == 1, half, pm);
				else cube_item_u8<3>(t.d, t.n, t.cur, q, p, bits, t.type, zp * zn, zp * zn + zn, key, idx, prune == 1, half, pm);
				ws.item_key[it - b0] = ((uint64_t) (key >> 8) << 16) | ((uint64_t) qp << 8) | (uint64_t) (key & 255u);
				ws.item_idx[it - b0] = idx;
			}
			__syncwarp();
			if ((int) lane < ntasks && !ws.task[lane].done) {
				CubeTask &t = ws.task[lane];
				const int lo = max(b0, (int) t.item_base), hi = min(b1, (int) t.item_base + (int) t.item_count);
				for (int it = lo; it < hi; it++)
					if (ws.item_key[it - b0] < t.pass_key) {
						t.pass_key = ws.item_key[it - b0];
						t.pass_idx = ws.item_idx[it - b0];
					}
			}
			__syncwarp();
		}
		// ---- per task: the reference's change / better logic (:1372-1400)
		if ((int) lane < ntasks && !ws.task[lane].done) {
			CubeTask &t = ws.task[lane];
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

// Lay out the work items of the running pass: tasks sorted by (clog, n) descending so that the lanes of a round agree
// on trip counts; task[i].item_base / item_count, ws.order. Returns the total number of items.
__device__ __forceinline__ int layout_items(WarpScratch &ws, int ntasks, int count, int sortkey, unsigned lane) {
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

// Start a round of ep_shaker_2_d for one task (:785-827): collapse, single-index case.
__device__ __noinline__ void window_begin_round(const Tables &T, CubeTask &t) {
	int index[kMaxEntries];
	unpack_idx(t.w_index, index, t.n);
	const int Mi = collapse_indices(index, t.n);
	if (Mi == 0) {
		U8Subset S;
		for (int i = 0; i < t.n; i++) S.d[i] = t.d[i];
		S.n = t.n;
		S.all_same = t.all_same != 0;
		for (int j = 0; j < 4; j++) S.mean[j] = t.mean[j];
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

// ep_shaker_2_d for the tasks with w_active set, starting from task.w_index (u8 path). Results in w_err_o /
// w_best_idx / w_best_ep.
__device__ __noinline__ void window_phase(const Tables &T, WarpScratch &ws, int ntasks, unsigned lane) {
	if ((int) lane < ntasks) {
		CubeTask &t = ws.task[lane];
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
			const CubeTask &t = ws.task[lane];
			count = qp_count(t.Mi, (1 << t.clog) - 1);
			sortkey = t.clog * 32 + t.n;
		}
		const int total = layout_items(ws, ntasks, count, sortkey, lane);
		if (total == 0) break;
		AMD_COUNT(8, total);
		AMD_COUNT(9, (total + 31) / 32);
		for (int b0 = 0; b0 < total; b0 += kItemBatch) {
			const int b1 = min(total, b0 + kItemBatch);
			for (int it = b0 + (int) lane; it < b1; it += 32) {
				const CubeTask &t = ws.task[item_owner(ws, ntasks, it)];
				const int qp = it - t.item_base;
				int q, p;
				qp_decode(qp, t.Mi, (1 << t.clog) - 1, q, p);
				uint64_t epo;
				uint32_t err;
				if (t.clog == 2) err = window_item_u8<2>(t.d, t.n, t.cur, q, p, t.w_size, t.w_bits_total, t.dim, epo);
				else if (t.clog == 3) err = window_item_u8<3>(t.d, t.n, t.cur, q, p, t.w_size, t.w_bits_total, t.dim, epo);
				else err = window_item_u8<4>(t.d, t.n, t.cur, q, p, t.w_size, t.w_bits_total, t.dim, epo);
				ws.item_key[it - b0] = ((uint64_t) err << 8) | (uint64_t) (255 - qp); // `<=`: the LAST minimum wins
				ws.item_idx[it - b0] = epo;
			}
			__syncwarp();
			if ((int) lane < ntasks && !ws.task[lane].done) {
				CubeTask &t = ws.task[lane];
				const int lo = max(b0, (int) t.item_base), hi = min(b1, (int) t.item_base + (int) t.item_count);
				for (int it = lo; it < hi; it++)
					if (ws.item_key[it - b0] < t.pass_key) {
						t.pass_key = ws.item_key[it - b0];
						t.pass_idx = ws.item_idx[it - b0];
					}
			}
			__syncwarp();
		}
		if ((int) lane < ntasks && !ws.task[lane].done) {
			CubeTask &t = ws.task[lane];
			const int qp = 255 - (int) (t.pass_key & 255u);
			int q0, p0;
			qp_decode(qp, t.Mi, (1 << t.clog) - 1, q0, p0);
			const int mb = (t.w_bits_total + 2 * t.dim - 1) / (2 * t.dim);
			uint64_t idg;
			uint32_t err_r;
			if (t.clog == 2) err_r = recluster_u8<2>(t.d, t.n, t.pass_idx, mb, t.dim, idg);
			else if (t.clog == 3) err_r = recluster_u8<3>(t.d, t.n, t.pass_idx, mb, t.dim, idg);
			else err_r = recluster_u8<4>(t.d, t.n, t.pass_idx, mb, t.dim, idg);
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

template <bool U8, int MINB>
__global__ void __launch_bounds__(kWarps * 32, MINB) bc7amd_kernel(const AmdParams p) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	WarpScratch *scratch = reinterpret_cast<WarpScratch *>(smem_raw);
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
	{ // prepare_block (bc7amd_core.cuh) with one texel per lane
		bool na = false, zo = false;
		real v[4] = {0, 0, 0, 0};
		if (lane < 16) {
			const float a = ws.in[lane * 4 + 3];
			if (a < 1.0) na = true;
			else if (((double) a >= 0.99999) || ((double) a < 0.00001)) zo = true;
#pragma unroll
			for (int j = 0; j < 4; j++) {
				v[j] = (real) (ws.in[lane * 4 + j] * 255.0f);
				ws.B.px[lane][j] = v[j];
				ws.B.pxc[j][lane] = v[j];
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
		if (lane == 0) ws.B.mode_mask = filter_modes(p.mode_mask, needs_alpha, zero_one, range < 1e-10);
	}
	__syncwarp();
	const uint32_t mask = ws.B.mode_mask & p.launch_modes;
	if (mask == 0 && !p.first) return; // whole warp

	const real carried = p.first ? A7_HUGE : p.best_err[block];
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
			AMD_T0();
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
			AMD_T(0);
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
			const bool cube_u8 = U8 && sp.dim == 3;
			real sub[kMaxEntries][4];
			int n = 0, idx[kMaxEntries], ep[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
			if ((int) lane < ntasks) {
				const int a = (int) lane / subsets, s = (int) lane - a * subsets;
				gather_subset(ws.B, subsets, ws.top[a], s, sp.dim, sub, n);
				unpack_idx(ws.qidx[ws.top[a]][s], idx, n);
				if (U8) {
					U8Subset S;
					make_u8_subset(sub, n, sp.dim, S);
					CubeTask &t = ws.task[lane];
					for (int i = 0; i < 16; i++) t.d[i] = i < n ? S.d[i] : 0u;
					t.idx_q = pack_idx(idx, n);
					for (int j = 0; j < 4; j++) t.mean[j] = S.mean[j];
					t.n = (uint8_t) n;
					t.clog = (uint8_t) ilog2(sp.clusters);
					t.bits = (uint8_t) sp.bits[0];
					t.type = (uint8_t) sp.parity;
					t.all_same = S.all_same ? 1 : 0;
					t.dim = (uint8_t) sp.dim;
					t.w_bits_total = (uint8_t) sp.bits[3];
					t.w_size = (uint8_t) sp.shake_size;
					t.w_index = t.idx_q;
					t.w_active = 1;
				}
			}
			__syncwarp();
			if (U8) {
				// shake_subset (:709-805): ep_shaker_d, ep_shaker_2_d on the quantiser's indices, and where the former
				// won, ep_shaker_2_d again on its indices
				AMD_T(1);
				if (cube_u8) cube_phase(T, ws, ntasks, p.zsplit_single ? p.zsplit_single : (subsets == 3 ? 1 : 0), lane, p.prune, p.split_n); // 3 subsets: small n, the per-item endpoint fit outweighs fuller rounds (measured)
				AMD_T(2);
				window_phase(T, ws, ntasks, lane);
				AMD_T(3);
				if (cube_u8) {
					if ((int) lane < ntasks) {
						CubeTask &t = ws.task[lane];
						t.w_active = (t.err_o < t.w_err_o) ? 1 : 0;
						t.w_index = t.best_idx;
					}
					__syncwarp();
					window_phase(T, ws, ntasks, lane);
					AMD_T(4);
				}
			}
			if ((int) lane < ntasks) {
				ShakeOut o;
				if (!U8) {
					o.err = shake_subset(T, sp, sub, n, idx, ep);
					o.idx = pack_idx(idx, n);
					o.ep[0] = pack_ep(ep[0]);
					o.ep[1] = pack_ep(ep[1]);
				} else {
					const CubeTask &t = ws.task[lane];
					o.err = t.w_err_o;
					o.idx = t.w_best_idx;
					o.ep[0] = (uint32_t) t.w_best_ep;
					o.ep[1] = (uint32_t) (t.w_best_ep >> 32);
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
			AMD_T(5);
		} else {
			const int nrot = 1 << mi.rotation_bits, nsel = 1 << mi.index_mode_bits;
			const int combos = nrot * nsel;
			const int ntasks = combos * 2;
			real blkv[16][4];
			int idx[16], ep[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
			int ib = 2, cb = 5;
			U8Subset S;
			uint64_t qpacked = 0;
			if ((int) lane < ntasks) { // quantise (work arrays overlay the task table: finish on all lanes first)
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
				ib = which == 0 ? (isel ? mi.index_bits1 : mi.index_bits0) : (isel ? mi.index_bits0 : mi.index_bits1);
				cb = which == 0 ? mi.vector_bits / 3 : mi.scalar_bits;
				unpack_idx(qpacked, idx, 16);
				if (U8) {
					make_u8_subset(blkv, 16, 3, S);
					CubeTask &t = ws.task[lane];
					for (int i = 0; i < 16; i++) t.d[i] = S.d[i];
					t.idx_q = pack_idx(idx, 16);
					for (int j = 0; j < 4; j++) t.mean[j] = S.mean[j];
					t.n = 16;
					t.clog = (uint8_t) ib;
					t.bits = (uint8_t) cb;
					t.type = CART;
					t.all_same = S.all_same ? 1 : 0;
					t.dim = 3;
					t.w_bits_total = (uint8_t) (6 * cb);
					t.w_size = 6;
				}
			}
			__syncwarp();
			if (U8) {
				cube_phase(T, ws, ntasks, p.zsplit_dual ? p.zsplit_dual : (ntasks <= 8 ? 4 : 2), lane, p.prune == 1 ? 1 : 0, 0);
				if ((int) lane < ntasks) {
					CubeTask &t = ws.task[lane];
					t.w_index = t.best_idx;
					t.w_active = 1;
				}
				__syncwarp();
				window_phase(T, ws, ntasks, lane);
			}
			if ((int) lane < ntasks) {
				const int bits[4] = {cb, cb, cb, 6 * cb};
				ShakeOut o;
				if (U8) {
					const CubeTask &t = ws.task[lane];
					o.err = t.w_err_o;
					o.idx = t.w_best_idx;
					o.ep[0] = (uint32_t) t.w_best_ep;
					o.ep[1] = (uint32_t) (t.w_best_ep >> 32);
				} else {
					shake_cube(T, blkv, 16, idx, (1 << ib) - 1, bits, CART);
					o.err = shake_window(T, blkv, 16, idx, ep, 6, (1 << ib) - 1, bits[3], 3);
					o.idx = pack_idx(idx, 16);
					o.ep[0] = pack_ep(ep[0]);
					o.ep[1] = pack_ep(ep[1]);
				}
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
				int epp[2][2][4], idxp[2][16];
				for (int w = 0; w < 2; w++) {
					const ShakeOut &o = ws.so[2 * bc + w];
					for (int k = 0; k < 4; k++) {
						epp[w][0][k] = (int) ((o.ep[0] >> (8 * k)) & 255u);
						epp[w][1][k] = (int) ((o.ep[1] >> (8 * k)) & 255u);
					}
					for (int i = 0; i < 16; i++) idxp[w][i] = (int) ((o.idx >> (4 * i)) & 15u);
				}
				uint64_t blk[2];
				pack_dual_index(mode, bc % nsel, bc / nsel, epp, idxp, blk);
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
	// first strict minimum over the modes in the reference's visiting order: a later launch only replaces the block
	// when its error is strictly lower
	if (lane == 0 && (p.first || best < carried)) {
		p.dst[block] = make_uint4((uint32_t) out0, (uint32_t) (out0 >> 32), (uint32_t) out1, (uint32_t) (out1 >> 32));
		p.best_err[block] = best;
	}
}

// Modes with few independent tasks per block -- mode 6: ONE partition, ONE subset; modes 4 / 5: 16 / 8 (rotation,
// index selection, vector | scalar) tasks -- leave most lanes of a warp-per-block mapping idle.  For these every
// THREAD takes a block and runs the serial form of the same search (bc7amd_block.cuh, the code the host build checks
// against the reference).  Measured at 1024^2 (64 Ki blocks): mode 6 16.7 -> 1.6 ms, mode 4 20.9 -> 14.2 ms, mode 5
// 10.7 -> 5.0 ms; the partitioned modes 0-3 and 7 are 1.5 .. 6x SLOWER this way and stay warp-per-block.
__global__ void __launch_bounds__(128) bc7amd_serial_kernel(const AmdParams p, const int mode) {
	const uint64_t block = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (block >= p.n_blocks) return;
	const Tables T{p.sp};
	const uint64_t per_slice = (uint64_t) p.img.blocks_x * p.img.blocks_y;
	const uint32_t slice = (uint32_t) (block / per_slice);
	const uint32_t rem = (uint32_t) (block - (uint64_t) slice * per_slice);
	const uint32_t by = rem / p.img.blocks_x, bx = rem - by * p.img.blocks_x;
	float in[64];
#pragma unroll 1
	for (int i = 0; i < 16; i++) {
		const float4 t = fetch_rgba(p.img, block, bx, by, slice, i);
		in[i * 4 + 0] = t.x; in[i * 4 + 1] = t.y; in[i * 4 + 2] = t.z; in[i * 4 + 3] = t.w;
	}
	BlockInput B;
	prepare_block(in, p.mode_mask, B);
	const real carried = p.first ? A7_HUGE : p.best_err[block];
	real best = carried;
	uint64_t out[2] = {0, 0};
	if (B.mode_mask & p.launch_modes & (1u << mode)) {
		uint64_t tmp[2];
		const real e = (mode_info(mode).alpha != 2) ? compress_single_index<true>(T, B, mode, tmp) : compress_dual_index<true>(T, B, mode, tmp);
		if (e < best) { best = e; out[0] = tmp[0]; out[1] = tmp[1]; }
	}
	if (p.first || best < carried) {
		p.dst[block] = make_uint4((uint32_t) out[0], (uint32_t) (out[0] >> 32), (uint32_t) out[1], (uint32_t) (out[1] >> 32));
		p.best_err[block] = best;
	}
}

} // namespace

#ifdef B200IC_AMD_TIMING
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
		uint8_t order[48 + 192 + 128];
		int pos = 0;
		for (int table = 0; table < 3; table++) {
			const int subsets = table < 2 ? 3 : 2, nparts = table == 0 ? 16 : 64;
			for (int size = 16; size >= 0; size--)
				for (int t = 0; t < nparts * subsets; t++) {
					const int part = t / subsets, s = t % subsets;
					int n = 0;
					for (int i = 0; i < 16; i++) n += subset_of(subsets, part, i) == s ? 1 : 0;
					if (n == size) order[pos++] = (uint8_t) t;
				}
		}
		e = cudaMemcpyToSymbol(c_qorder, order, sizeof(order));
		if (e != cudaSuccess) return e;
	}
	e = cudaFuncSetAttribute(bc7amd_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (kWarps * sizeof(WarpScratch)));
	if (e != cudaSuccess) return e;
	e = cudaFuncSetAttribute(bc7amd_kernel<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (kWarps * sizeof(WarpScratch)));
	if (e != cudaSuccess) return e;
	e = cudaFuncSetAttribute(bc7amd_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (kWarps * sizeof(WarpScratch)));
	if (e != cudaSuccess) return e;
	e = cudaFuncSetAttribute(bc7amd_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (kWarps * sizeof(WarpScratch)));
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
	p.zsplit_single = getenv("B200IC_AMD_ZS") ? atoi(getenv("B200IC_AMD_ZS")) : 0;
	p.zsplit_dual = getenv("B200IC_AMD_ZD") ? atoi(getenv("B200IC_AMD_ZD")) : 0;
	p.split_n = getenv("B200IC_AMD_SPLITN") ? atoi(getenv("B200IC_AMD_SPLITN")) : 0; // measured: 8 / 6 are 3-4 % slower than no split
	p.prune = getenv("B200IC_AMD_PRUNE") ? atoi(getenv("B200IC_AMD_PRUNE")) : 0;
	if (p.n_blocks == 0) return cudaSuccess;
	const uint64_t grid = (p.n_blocks + kWarps - 1) / kWarps;
	const size_t smem = kWarps * sizeof(WarpScratch);
	// 8-bit sources: every component is an exact integer, the exact INT32 shakers apply (bc7amd_int.cuh)
	const bool u8 = img.format == B200IC_FMT_R8 || img.format == B200IC_FMT_RG8 || img.format == B200IC_FMT_RGB8 ||
									img.format == B200IC_FMT_RGB8_SRGB || img.format == B200IC_FMT_RGBA8 || img.format == B200IC_FMT_RGBA8_SRGB ||
									img.format == B200IC_FMT_BLOCKS_RGBA8;
	static const int variant = getenv("B200IC_AMD_VARIANT") ? atoi(getenv("B200IC_AMD_VARIANT")) : 3;
	const int fused = getenv("B200IC_AMD_FUSED") ? atoi(getenv("B200IC_AMD_FUSED")) : 0;
	const int serial_mask = getenv("B200IC_AMD_SERIAL") ? (int) strtol(getenv("B200IC_AMD_SERIAL"), nullptr, 0) : 0x70;
	// One launch per mode, in the reference's visiting order {6,4,3,1,2,0,7,5} (src/amd_bc7_body.cpp:1400), the running
	// best block and its error carried in dst / best_err: every SM then runs ONE mode's code at a time.  The fused
	// all-modes launch is 13 % (opaque) to 26 % (translucent) slower than the sum of its single-mode launches
	// (instruction-cache and local-memory interference between warps in different modes, profiles/).
	e = cudaMallocAsync((void **) &p.best_err, p.n_blocks * sizeof(real), stream);
	if (e != cudaSuccess) return e;
	const uint32_t user = p.mode_mask ? p.mode_mask : 0xCFu;
	int launches = 0;
	for (int vi = 0; vi < 8; vi++) {
		const int mode = mode_visit_order(vi);
		if (fused) {
			if (vi) break;
			p.launch_modes = 0xFFu;
		} else {
			if (!(user & (1u << mode)) && launches) continue; // (the first launch always runs: it initialises dst)
			p.launch_modes = 1u << mode;
		}
		p.first = launches == 0;
		if (u8 && !fused && ((serial_mask >> mode) & 1)) bc7amd_serial_kernel<<<(unsigned) ((p.n_blocks + 127) / 128), 128, 0, stream>>>(p, mode);
		else if (u8 && variant == 4) bc7amd_kernel<true, 4><<<(unsigned) grid, kWarps * 32, smem, stream>>>(p);
		else if (u8 && variant == 2) bc7amd_kernel<true, 2><<<(unsigned) grid, kWarps * 32, smem, stream>>>(p);
		else if (u8) bc7amd_kernel<true, 3><<<(unsigned) grid, kWarps * 32, smem, stream>>>(p);
		else bc7amd_kernel<false, 4><<<(unsigned) grid, kWarps * 32, smem, stream>>>(p);
		launches++;
	}
	count_launches(launches - 1);
	e = cudaGetLastError();
	cudaFreeAsync(p.best_err, stream);
	return e;
}

} // namespace b200ic
