// bc7amd_core.cuh -- AMD-Compressonator-compatible BC7 encoder search (all eight modes), scalar building blocks.
//
// Follows the search of the reference's BC7BlockEncoder at quality = performance = 1.0, colourRestrict =
// alphaRestrict = true (the only configuration its image API uses, src/amd_bc7_compressor.cpp:58-65):
//   block set-up / mode filter / mode order      src/amd_bc7_body.cpp:1289-1465
//   single-index modes 0,1,2,3,6,7               src/amd_bc7_body.cpp:548-890
//   dual-index modes 4,5                         src/amd_bc7_body.cpp:1059-1278
//   bit packing                                  src/amd_bc7_body.cpp:333-538, 902-1056
//   optQuantAnD_d + helpers                      src/amd_bc7_3dquant_vpc.cpp:138-420, 686-695, 1201-1286, 1874-2045
//   ep_shaker_d / ep_shaker_2_d / single point   src/amd_shake.cpp:351-367, 513-538, 546-1404
//
// Differences in construction (results are equal up to sort tie order, see below):
//   * the reference's 100 MB `ramp[clog][bits][p1][p2][i]` double LUT (src/amd_shake.cpp:225,283-286) is an exact
//     integer expression of the expanded endpoints (ramp_int below); it is computed in registers.
//   * the single-colour tables sp_idx / sp_err (src/amd_shake.cpp:230-232,293-345) are packed to 4 bytes per entry
//     (p1, p2, distance) and built once on the host by build_single_point_table().
//   * ep_shaker_d's cached per-texel error cube `ce` is recomputed (same operations, same order of additions).
//   * the outputs nobody reads (reconstructed points `out`, `direction`, `step`, `epo`, ep_shaker_d's endpoint codes
//     -- overwritten by the ep_shaker_2_d call that always follows) are not produced.
//   * qsort is replaced by a stable insertion sort; the reference's glibc qsort orders EQUAL keys differently, which
//     is one reason this path is PSNR-gated and not bit-gated (SURVEY.md 7 hard part 4).
//   * trace quantiser: never reached at quality 1 (m_blockMaxRange <= 255 always), not built.
// `real` is the arithmetic type of the search; double reproduces the reference.
#pragma once
#include <stdint.h>
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define A7_HD __host__ __device__ __forceinline__
#define A7_HDN __host__ __device__ __noinline__
#else
#define A7_HD inline
#define A7_HDN
#endif
#if defined(__CUDA_ARCH__)
#define B7T_QUAL __constant__ const
#else
#define B7T_QUAL const
#endif
#include "bc7_tables.h"

namespace b200ic {
namespace amd7 {

typedef double real;
#define A7_HUGE DBL_MAX
#ifndef A7_STATS_IT
#define A7_STATS_IT(it)
#define A7_STATS_F()
#define A7_STATS_G()
#define A7_STATS_REPLAY()
#define A7_STATS_OVF(v)
#endif

constexpr int kMaxEntries = 16;
constexpr int kQuantMaxTry = 200; // optQuantAnD_d maxTry (:1876)

// ---- mode description (bti[], src/amd_bc7_body.cpp:84-94) ---------------------------------------------
enum Parity { CART = 0, SAME_PAR = 1, BCC = 2 };
struct ModeInfo {
	uint8_t alpha;        // 0 none, 1 combined, 2 separate
	uint8_t partition_bits, rotation_bits, index_mode_bits, scalar_bits, vector_bits, parity, subsets, index_bits0, index_bits1;
};
A7_HD ModeInfo mode_info(int m) {
	switch (m) {
	case 0: return {0, 4, 0, 0, 0, 12, BCC, 3, 3, 0};
	case 1: return {0, 6, 0, 0, 0, 18, SAME_PAR, 2, 3, 0};
	case 2: return {0, 6, 0, 0, 0, 15, CART, 3, 2, 0};
	case 3: return {0, 6, 0, 0, 0, 21, BCC, 2, 2, 0};
	case 4: return {2, 0, 2, 1, 6, 15, CART, 1, 2, 3};
	case 5: return {2, 0, 2, 0, 8, 21, CART, 1, 2, 2};
	case 6: return {1, 0, 0, 0, 0, 28, BCC, 1, 4, 0};
	default: return {1, 6, 0, 0, 0, 20, BCC, 2, 2, 0};
	}
}

// ---- ramps ---------------------------------------------------------------------------------------------
// REFERENCE QUIRK (kept for parity): src/amd_shake.cpp tests `#if USE_FINAL_BC7_WEIGHTS` (:236) but only
// src/amd_bc7_body.cpp defines that macro (:60), so the shakers' ramp table is built with the "pure linear" weights
// i/(2^clog - 1) (:245-252), not the BC7 hardware weights {0,21,43,64}/64 ... that a decoder applies. Every
// endpoint / index refinement below therefore optimises against floor(e1 + i/(C-1) * (e2-e1) + 0.5).
// Exact integer form: the fractional part of the exact value is an odd multiple of 1/(2(C-1)), never within
// rounding distance of an integer, so the reference's double evaluation and this integer one always agree.
A7_HD int expand_bits(int bits, int v) { return (v << (8 - bits)) | (v >> (2 * bits - 8)); } // expand_ (:254-257)
A7_HD int ramp_int(int e1, int e2, int i, int clog) {
	const int D = (1 << clog) - 1;
	return (2 * (D * e1 + i * (e2 - e1)) + D) / (2 * D);
}

// ---- single-colour table ----------------------------------------------------------------------------------
// entry (clog-2, bits-5, value, par1, par2, index) -> p1 | p2<<8 | distance<<16 ; error = distance^2
struct Tables {
	const uint32_t *sp;
};
constexpr size_t kSpEntries = 3u * 4u * 256u * 2u * 2u * 16u;
A7_HD size_t sp_slot(int clog, int bits, int value, int o1, int o2, int i) {
	return ((((size_t) ((clog - 2) * 4 + (bits - 5)) * 256 + value) * 2 + o1) * 2 + o2) * 16 + i;
}
inline void build_single_point_table(uint32_t *sp) { // init_ramps (:293-345)
	const uint32_t kEmpty = 0xffffffffu;
#pragma unroll 1
	for (size_t i = 0; i < kSpEntries; i++) sp[i] = kEmpty;
#pragma unroll 1
	for (int clog = 2; clog < 5; clog++)
#pragma unroll 1
		for (int bits = 5; bits < 9; bits++) {
#pragma unroll 1
			for (int p1 = 0; p1 < (1 << bits); p1++)
#pragma unroll 1
				for (int p2 = 0; p2 < (1 << bits); p2++)
#pragma unroll 1
					for (int i = 0; i < (1 << clog); i++) {
						const int v = ramp_int(expand_bits(bits, p1), expand_bits(bits, p2), i, clog);
						sp[sp_slot(clog, bits, v, p1 & 1, p2 & 1, i)] = (uint32_t) p1 | ((uint32_t) p2 << 8); // last writer wins
					}
#pragma unroll 1
			for (int o1 = 0; o1 < 2; o1++)
#pragma unroll 1
				for (int o2 = 0; o2 < 2; o2++)
#pragma unroll 1
					for (int i = 0; i < (1 << clog); i++) {
						bool exact[256];
#pragma unroll 1
						for (int v = 0; v < 256; v++) exact[v] = sp[sp_slot(clog, bits, v, o1, o2, i)] != kEmpty;
#pragma unroll 1
						for (int v = 0; v < 256; v++) {
							if (exact[v]) continue;
							int k = 1;
#pragma unroll 1
							for (; k < 256; k++)
								if ((v - k >= 0 && exact[v - k]) || (v + k < 256 && exact[v + k])) break;
							uint32_t src = 0;
							if (v - k >= 0 && exact[v - k]) src = sp[sp_slot(clog, bits, v - k, o1, o2, i)];
							else if (v + k < 256 && exact[v + k]) src = sp[sp_slot(clog, bits, v + k, o1, o2, i)];
							sp[sp_slot(clog, bits, v, o1, o2, i)] = (src & 0xffffu) | ((uint32_t) k << 16);
						}
					}
		}
}

// ---- small utilities ---------------------------------------------------------------------------------------
// order[] = permutation sorting key[] ascending, equal keys in original order (what an insertion sort with the
// reference's comparator `a - b > 0` produces). Rank counting: fixed trip counts, no data-dependent branches.
A7_HD void sort_order(const real *key, int *order, int n) {
#pragma unroll 1
	for (int i = 0; i < n; i++) {
		const real k = key[i];
		int rank = 0;
#pragma unroll 1
		for (int j = 0; j < n; j++) {
			const real kj = key[j];
			rank += ((k - kj > 0) || (!(kj - k > 0) && j < i)) ? 1 : 0;
		}
		order[rank] = i;
	}
}
A7_HD int ilog2(int v) { int c = 0; while (v >>= 1) c++; return c; }

// eigenVector_d (:336-420): dominant eigenvector by repeated squaring, 3 rounds of (normalise by the largest
// diagonal entry, square 8 times). The covariance is bitwise symmetric and stays so (division by one scalar; a
// product term c[i][k]*c[k][j] of (i,j) equals the term c[j][k]*c[k][i] of (j,i) and both sums run over k in the
// same order), so only the upper triangle is computed -- identical bits, two thirds of the work, all in registers.
template <int DIM> A7_HD void dominant_axis_t(const real cov[4][4], real axis[4]) {
	real c[DIM][DIM];
#pragma unroll
	for (int i = 0; i < DIM; i++)
#pragma unroll
		for (int j = 0; j < DIM; j++) c[i][j] = cov[i][j];
#pragma unroll 1
	for (int round = 0; round < 3; round++) {
		real md = 0;
#pragma unroll
		for (int i = 0; i < DIM; i++) md = c[i][i] > md ? c[i][i] : md;
		if (md <= 0) return;
#pragma unroll
		for (int i = 0; i < DIM; i++)
#pragma unroll
			for (int j = i; j < DIM; j++) {
				c[i][j] /= md;
				c[j][i] = c[i][j];
			}
#pragma unroll 1
		for (int m = 0; m < 8; m++) {
			real nx[DIM][DIM];
#pragma unroll
			for (int i = 0; i < DIM; i++)
#pragma unroll
				for (int j = i; j < DIM; j++) {
					real t = 0;
#pragma unroll
					for (int k = 0; k < DIM; k++) t += c[i][k] * c[k][j];
					nx[i][j] = t;
				}
#pragma unroll
			for (int i = 0; i < DIM; i++)
#pragma unroll
				for (int j = i; j < DIM; j++) {
					c[i][j] = nx[i][j];
					c[j][i] = nx[i][j];
				}
		}
	}
	real md = 0;
	int k = 0;
#pragma unroll
	for (int i = 0; i < DIM; i++) {
		k = c[i][i] > md ? i : k;
		md = c[i][i] > md ? c[i][i] : md;
	}
	real row[DIM];
#pragma unroll
	for (int i = 0; i < DIM; i++) {
		row[i] = c[0][i];
#pragma unroll
		for (int r = 1; r < DIM; r++)
			if (k == r) row[i] = c[r][i];
	}
	real t = 0;
#pragma unroll
	for (int i = 0; i < DIM; i++) {
		t += row[i] * row[i];
		axis[i] = row[i];
	}
	t = sqrt(t);
	if (t <= 0) return;
#pragma unroll
	for (int i = 0; i < DIM; i++) axis[i] /= t;
}
A7_HDN void dominant_axis(const real cov[4][4], real axis[4], int dim) {
	if (dim == 3) dominant_axis_t<3>(cov, axis);
	else dominant_axis_t<4>(cov, axis);
}

// ---- quantiser I/O --------------------------------------------------------------------------------------------
// A quantiser call reads its n points straight from the block (channel-major, so that the 32 lanes of a warp -- each
// working on another subset of the same block -- hit 16 different bank pairs or broadcast) and keeps its two
// n-element FP64 work arrays in caller-provided storage: lane-strided shared memory on the GPU (element k at
// [k * stride]: conflict-free, no local-memory traffic), plain local arrays on the host.
struct QuantIO { // generic pointers: host build and the thread-per-block kernel (local work arrays)
	const real *px;   // px[channel * 16 + texel], 0..255
	uint64_t texels;  // 4 bits per entry: texel of entry k
	uint32_t chan;    // 2 bits per component: source channel of component j
	real *proj, *dev; // work arrays
	int stride;
	A7_HD real point(uint32_t idx) const { return px[idx]; }
	A7_HD real get_proj(int i) const { return proj[i * stride]; }
	A7_HD void set_proj(int i, real v) const { proj[i * stride] = v; }
	A7_HD real get_dev(int i) const { return dev[i * stride]; }
	A7_HD void set_dev(int i, real v) const { dev[i * stride] = v; }
};
#if defined(__CUDACC__)
// The same three arrays in SHARED memory, named by 32-bit shared-window addresses and touched with ld.shared / st.shared:
// the warp kernels' quantiser then issues LDS / STS with 32-bit address arithmetic instead of generic 64-bit loads.
struct QuantIOShared {
	uint32_t px, proj, dev; // byte addresses in the shared window (__cvta_generic_to_shared)
	uint64_t texels;
	uint32_t chan;
	int stride;             // elements between consecutive entries of proj / dev
	static __device__ __forceinline__ real lds(uint32_t a) {
		real v;
		asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
		return v;
	}
	static __device__ __forceinline__ void sts(uint32_t a, real v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
	static __device__ __forceinline__ real lds_const(uint32_t a) { // the block's texels do not change during a call: the compiler may move / merge these loads
		real v;
		asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
		return v;
	}
	__device__ __forceinline__ real point(uint32_t idx) const { return lds_const(px + idx * 8u); }
	__device__ __forceinline__ real get_proj(int i) const { return lds(proj + (uint32_t) (i * stride) * 8u); }
	__device__ __forceinline__ void set_proj(int i, real v) const { sts(proj + (uint32_t) (i * stride) * 8u, v); }
	__device__ __forceinline__ real get_dev(int i) const { return lds(dev + (uint32_t) (i * stride) * 8u); }
	__device__ __forceinline__ void set_dev(int i, real v) const { sts(dev + (uint32_t) (i * stride) * 8u, v); }
};
#endif
template <class IO> A7_HD real quant_point(const IO &io, int k, int j) {
	return io.point(((io.chan >> (2 * j)) & 3u) * 16u + (uint32_t) ((io.texels >> (4 * k)) & 15u));
}
A7_HD uint64_t texels_of_mask(uint32_t mask16, int &n) { // entries in texel order
	uint64_t t = 0;
	n = 0;
#pragma unroll 1
	for (int i = 0; i < 16; i++)
		if (mask16 & (1u << i)) {
			t |= (uint64_t) i << (4 * n);
			n++;
		}
	return t;
}

// (2 i + 1 - n) / 2 / n of quant_AnD_Shell's move to the centre of the fundamental simplex (:1258-1259): one IEEE division
// per entry on the host; on the GPU a 17 x 16 table of the same correctly rounded quotients (uploaded by
// init_bc7amd_tables) instead of a 40-instruction FP64 division per entry and call.
#if defined(__CUDACC__)
static __constant__ double c_simplex_step[17 * 16];
#endif
#if defined(__CUDA_ARCH__)
#define A7_SIMPLEX_STEP(n, i) c_simplex_step[(n) * 16 + (i)]
#else
#define A7_SIMPLEX_STEP(n, i) ((2. * (real) (i) + 1 - (real) (n)) / 2. / (real) (n))
#endif
// quant_AnD_Shell (:1201-1286): optimal uniform k-level quantisation of the n scalars of io's proj array (lattice A_n* decoding).
// Returns the indices packed 4 bits per entry (values taken & 15 like every consumer does).
template <class IO> A7_HDN uint64_t lattice_quantise(const IO &io, int k, int n) {
	real m = io.get_proj(0), M = m;
#pragma unroll 1
	for (int i = 1; i < n; i++) {
		const real v = io.get_proj(i);
		m = m < v ? m : v;
		M = M > v ? M : v;
	}
	if (M == m) return 0;
	const real s = (real) (k - 1) / (M - m);
	real dm = 0, r = 0;
	uint64_t z4 = 0; // floor values, one nibble each (0 .. k-1)
#pragma unroll 1
	for (int i = 0; i < n; i++) {
		const real v = io.get_proj(i) * s;
		const real z = floor(v + 0.5 - m * s);
		z4 |= (uint64_t) ((int) z & 15) << (4 * i);
		const real d = v - z - m * s;
		io.set_dev(i, d);
		dm += d;
		r += d * d;
	}
	uint32_t inc = 0; // entries moved one level up
	if ((real) n * r - dm * dm >= (real) (n - 1) / 4 / 2) {
		dm /= (real) n;
#pragma unroll 1
		for (int i = 0; i < n; i++) io.set_dev(i, io.get_dev(i) - dm);
		// stable rank of every deviation (what an insertion sort with the comparator `a - b > 0` produces)
		// stable rank of every deviation (what an insertion sort with the comparator `a - b > 0` produces), one compare
		// per PAIR: for i < j exactly one of the two gains a rank -- entry i if d_i > d_j, else entry j (a - b > 0 is
		// a > b for these finite values: a difference of distinct finite numbers never rounds to zero)
		uint64_t rank4 = 0;
#pragma unroll 1
		for (int i = 0; i < n; i++) {
			const real ki = io.get_dev(i);
#pragma unroll 1
			for (int j = i + 1; j < n; j++) rank4 += 1ull << (4 * (ki > io.get_dev(j) ? i : j));
		}
		uint64_t ord = 0;
#pragma unroll 1
		for (int i = 0; i < n; i++) ord |= (uint64_t) i << (4 * (int) ((rank4 >> (4 * i)) & 15u));
		real mm = 0, l = 0;
		int j = -1;
#pragma unroll 1
		for (int i = 0; i < n; i++) {
			l += io.get_dev((int) ((ord >> (4 * i)) & 15u)) - A7_SIMPLEX_STEP(n, i);
			if (l < mm) { mm = l; j = i; }
		}
		j = (j + 1) % n;
#pragma unroll 1
		for (int i = j; i < n; i++) inc |= 1u << (int) ((ord >> (4 * i)) & 15u);
	}
	int mi = 99;
#pragma unroll 1
	for (int i = 0; i < n; i++) {
		const int v = (int) ((z4 >> (4 * i)) & 15u) + (int) ((inc >> i) & 1u);
		mi = mi < v ? mi : v;
	}
	uint64_t out = 0;
#pragma unroll 1
	for (int i = 0; i < n; i++) {
		const int v = (int) ((z4 >> (4 * i)) & 15u) + (int) ((inc >> i) & 1u) - mi;
		A7_STATS_OVF(v);
		out |= (uint64_t) (v & 15) << (4 * i);
	}
	return out;
}

// refit (:1923-1956): direction through the index-weighted centred points, projections into io.proj; s, t out
template <int DIM, class IO> A7_HD void quant_refit(const IO &io, const real *mean, int n, uint64_t a, bool want_st, real &s, real &t) {
	real dir[DIM], q = 0;
	real ss = 0, tt = 0;
#pragma unroll
	for (int j = 0; j < DIM; j++) dir[j] = 0;
#pragma unroll 1
	for (int k = 0; k < n; k++) {
		const int ik = (int) ((a >> (4 * k)) & 15u);
		ss += ik;
		tt += ik * ik;
#pragma unroll
		for (int j = 0; j < DIM; j++) dir[j] += (quant_point(io, k, j) - mean[j]) * ik;
	}
#pragma unroll
	for (int j = 0; j < DIM; j++) q += dir[j] * dir[j];
	if (want_st) {
		ss /= (real) n;
		tt = tt - ss * ss * (real) n;
		tt = (tt == 0 ? 0. : 1 / tt);
	}
	q = sqrt(q);
	if (want_st) tt *= q;
	if (q != 0)
#pragma unroll
		for (int j = 0; j < DIM; j++) dir[j] /= q;
#pragma unroll 1
	for (int k = 0; k < n; k++) {
		real p = 0;
#pragma unroll
		for (int j = 0; j < DIM; j++) p += (quant_point(io, k, j) - mean[j]) * dir[j];
		io.set_proj(k, p);
	}
	s = ss;
	t = tt;
}

// optQuantAnD_d (:1874-2045): PCA line + iterative optimal uniform quantiser over the n points of `io`.
// Returns the SSE; index_out = final indices, 4 bits per entry.
template <int DIM, class IO> A7_HD real quantise_points(const IO &io, int n, int clusters, uint64_t &index_out) {
	index_out = 0;
	if (n == 0) return 0;
	real mean[DIM];
#pragma unroll
	for (int j = 0; j < DIM; j++) mean[j] = 0;
#pragma unroll 1
	for (int k = 0; k < n; k++)
#pragma unroll
		for (int j = 0; j < DIM; j++) mean[j] += quant_point(io, k, j);
#pragma unroll
	for (int j = 0; j < DIM; j++) mean[j] /= (real) n;
	real cov[4][4];
#pragma unroll
	for (int i = 0; i < 4; i++)
#pragma unroll
		for (int j = 0; j < 4; j++) cov[i][j] = 0;
#pragma unroll 1
	for (int k = 0; k < n; k++) {
		real c[DIM];
#pragma unroll
		for (int j = 0; j < DIM; j++) c[j] = quant_point(io, k, j) - mean[j];
#pragma unroll
		for (int i = 0; i < DIM; i++)
#pragma unroll
			for (int j = 0; j <= i; j++) cov[i][j] += c[i] * c[j];
	}
	real t = 0;
#pragma unroll
	for (int i = 0; i < DIM; i++) {
		t += cov[i][i];
#pragma unroll
		for (int j = 0; j < i; j++) cov[j][i] = cov[i][j];
	}
	if (t < (1. / 256.)) return 0;
	{
		real dir[4] = {0, 0, 0, 0};
		dominant_axis(cov, dir, DIM);
#pragma unroll 1
		for (int k = 0; k < n; k++) {
			real p = 0;
#pragma unroll
			for (int j = 0; j < DIM; j++) p += (quant_point(io, k, j) - mean[j]) * dir[j];
			io.set_proj(k, p);
		}
	}
	// The loop below is the reference's (:1911-2007), including its quirks: the convergence test compares against
	// the indices of iteration 1 (the `index_[j]=index_[j]` no-op, :1997), so an assignment that oscillates never
	// "converges" and runs all 200 iterations (0.3 .. 1.4 % of the calls, 100x the cost of the rest).  Both steps of an
	// iteration are PURE functions of the current index vector -- refit+reassign F(index) and the lattice quantiser
	// G(projection(index)) -- so
	//   * the state is carried as one packed 64-bit word `cur` (4 bits per entry) and F / G are memoised on it (four
	//     slots held in registers): a replayed step costs a few compares;
	//   * an iteration is a pure function of (cur, try_two): once the state after G repeats one of the last 8 states
	//     with try_two unchanged (or already negative: `try_two--` then never cuts a refit chain again) the remaining
	//     iterations are periodic and no iteration of the period passed the convergence test, so the final state is
	//     read from the history instead of being replayed (kHist covers every period seen; longer ones just replay).
	uint64_t mk0 = 0, mk1 = 0, mk2 = 0, mk3 = 0, mf0 = 0, mf1 = 0, mf2 = 0, mf3 = 0, mg0 = 0, mg1 = 0, mg2 = 0, mg3 = 0;
	uint32_t gvalid = 0;
	int memo_n = 0, memo_next = 0;
	constexpr int kHist = 8;
	uint64_t hist[kHist];
	int hist_try[kHist];
	uint64_t first = 0;
	int try_two = 50;
	real s;
	A7_STATS_G();
	uint64_t cur = lattice_quantise(io, clusters, n); // iteration 0
	int it = 1;
#pragma unroll 1
	for (; it < kQuantMaxTry; it++) {
		int last = -1;
		bool have_proj = false;
		int done;
		do {
			const uint64_t a = cur;
			int slot = -1;
			if (memo_n > 0 && mk0 == a) slot = 0;
			if (memo_n > 1 && mk1 == a) slot = 1;
			if (memo_n > 2 && mk2 == a) slot = 2;
			if (memo_n > 3 && mk3 == a) slot = 3;
			uint64_t b;
			if (slot >= 0) {
				A7_STATS_REPLAY();
				b = slot == 0 ? mf0 : (slot == 1 ? mf1 : (slot == 2 ? mf2 : mf3));
				have_proj = false;
			} else {
				A7_STATS_F();
				quant_refit<DIM>(io, mean, n, a, true, s, t);
				// The reference sorts the projections and walks the cluster boundaries (k + 0.5 - s) * t with one
				// running k (:1977-1984). The boundaries are non-decreasing in k (t >= 0), so for sorted input the
				// running k of an element equals the NUMBER of boundaries it exceeds: no sort, no dependent loop.
				b = 0;
				if (clusters <= 8) { // the (up to 7) boundaries in registers, one pass over the projections
					real bound[7];
#pragma unroll
					for (int c = 0; c < 7; c++) bound[c] = c < clusters - 1 ? ((real) c + 0.5 - s) * t : 0;
#pragma unroll 1
					for (int j = 0; j < n; j++) {
						const real pj = io.get_proj(j);
						int cnt = 0;
#pragma unroll
						for (int c = 0; c < 7; c++) cnt += (c < clusters - 1 && pj > bound[c]) ? 1 : 0;
						b |= (uint64_t) cnt << (4 * j);
					}
				} else {
#pragma unroll 1
					for (int c = 0; c < clusters - 1; c++) {
						const real bound = ((real) c + 0.5 - s) * t;
#pragma unroll 1
						for (int j = 0; j < n; j++) b += (uint64_t) (io.get_proj(j) > bound ? 1 : 0) << (4 * j);
					}
				}
				slot = memo_next;
				memo_next = (memo_next + 1) & 3;
				memo_n = memo_n < 4 ? memo_n + 1 : 4;
				if (slot == 0) { mk0 = a; mf0 = b; }
				else if (slot == 1) { mk1 = a; mf1 = b; }
				else if (slot == 2) { mk2 = a; mf2 = b; }
				else { mk3 = a; mf3 = b; }
				gvalid &= ~(1u << slot);
				have_proj = true;
			}
			cur = b;
			done = (b == a);
			last = slot;
		} while (!done && try_two--);
		if (it == 1) first = cur;
		else if (first == cur) { A7_STATS_IT(it); break; }
		if ((gvalid >> last) & 1u) {
			cur = last == 0 ? mg0 : (last == 1 ? mg1 : (last == 2 ? mg2 : mg3));
		} else {
			if (!have_proj) { // projection of the memoised refit's INPUT indices
				const uint64_t a = last == 0 ? mk0 : (last == 1 ? mk1 : (last == 2 ? mk2 : mk3));
				real s2, t2;
				quant_refit<DIM>(io, mean, n, a, false, s2, t2);
			}
			A7_STATS_G();
			cur = lattice_quantise(io, clusters, n);
			if (last == 0) mg0 = cur;
			else if (last == 1) mg1 = cur;
			else if (last == 2) mg2 = cur;
			else mg3 = cur;
			gvalid |= 1u << last;
		}
		// `cur` is now the state at the top of iteration it + 1; states from the top of iteration 2 on are recorded
		// (every iteration >= 2 runs the convergence test, so a repeat proves that the test fails forever)
		if (it >= 2) {
			int period = 0;
#pragma unroll 1
			for (int pd = 1; pd <= kHist && pd <= it - 1; pd++) {
				const int h = (it + 1 - pd) & (kHist - 1);
				if (hist[h] == cur && (hist_try[h] == try_two || hist_try[h] < 0)) { period = pd; break; }
			}
			if (period) {
				const int r = (kQuantMaxTry - 1 - it) % period; // iterations still to run, modulo the period
				cur = hist[(it + 1 - period + r) & (kHist - 1)];
				A7_STATS_IT(kQuantMaxTry + 1);
				break;
			}
		}
		hist[(it + 1) & (kHist - 1)] = cur;
		hist_try[(it + 1) & (kHist - 1)] = try_two;
		if (it == kQuantMaxTry - 1) { A7_STATS_IT(kQuantMaxTry); }
	}
	// final reconstruction error (:2010-2043)
	real dir[DIM];
	s = t = 0;
#pragma unroll
	for (int j = 0; j < DIM; j++) dir[j] = 0;
#pragma unroll 1
	for (int k = 0; k < n; k++) {
		const int ik = (int) ((cur >> (4 * k)) & 15u);
		s += ik;
		t += ik * ik;
#pragma unroll
		for (int j = 0; j < DIM; j++) dir[j] += (quant_point(io, k, j) - mean[j]) * ik;
	}
	s /= (real) n;
	t = t - s * s * (real) n;
	t = (t == 0 ? 0. : 1 / t);
	real err = 0;
#pragma unroll 1
	for (int i = 0; i < n; i++) {
		const int ii = (int) ((cur >> (4 * i)) & 15u);
#pragma unroll
		for (int j = 0; j < DIM; j++) {
			const real v = quant_point(io, i, j);
			const real o = mean[j] + dir[j] * t * ((real) ii - s);
			err += (v - o) * (v - o);
		}
	}
	index_out = cur;
	return err;
}
template <class IO> A7_HDN real quantise_subset(const IO &io, int n, int clusters, int dim, uint64_t &index_out) {
	return dim == 3 ? quantise_points<3>(io, n, clusters, index_out) : quantise_points<4>(io, n, clusters, index_out);
}

// ep_find_floor (:351-367)
A7_HD int endpoint_floor(real v, int bits, int use_par, int odd) {
	int i1 = 0, i2 = 1 << (bits - use_par);
	odd = use_par ? odd : 0;
	while (i2 - i1 > 1) {
		const int j = (i1 + i2) / 2;
		if (v >= (real) expand_bits(bits, (j << use_par) + odd)) i1 = j;
		else i2 = j;
	}
	return (i1 << use_par) + odd;
}

// index_collapse_ (:513-538); returns the new maximum index.
// The reference keeps the LARGEST d in [2, max - min] that divides every index[k] - min (else 1): that is the gcd of the
// differences (every common divisor divides the gcd, and the gcd does not exceed the largest difference).  Here: AND of
// per-value divisor masks (bit d set when d divides v; v = 0: every d), highest common bit; the division by D <= 15 is a
// multiply and shift.
static B7T_QUAL uint16_t kDivisorMask[16] = {0xfffe, 0x0002, 0x0006, 0x000a, 0x0016, 0x0022, 0x004e, 0x0082,
																						 0x0116, 0x020a, 0x0426, 0x0802, 0x105e, 0x2002, 0x4086, 0x802a};
static B7T_QUAL uint16_t kInv12[16] = {0, 4096, 2048, 1366, 1024, 820, 683, 586, 512, 456, 410, 373, 342, 316, 293, 274}; // ceil(4096 / d)
A7_HD int collapse_indices(int *index, int n) {
	int mi = index[0], Mi = index[0];
#pragma unroll 1
	for (int k = 1; k < n; k++) {
		mi = mi < index[k] ? mi : index[k];
		Mi = Mi > index[k] ? Mi : index[k];
	}
	uint32_t common = 0xfffeu;
#pragma unroll 1
	for (int k = 0; k < n; k++) common &= kDivisorMask[(index[k] - mi) & 15];
	// divisors above the largest difference only survive when every difference is 0: D = 1 then (the loop of :520 is empty)
	common &= (2u << (Mi - mi)) - 1u;
	int D = 1;
	while (common >> (D + 1)) D++;
	const uint32_t inv = kInv12[D];
	int top = 0;
#pragma unroll 1
	for (int k = 0; k < n; k++) {
		index[k] = (int) (((uint32_t) (index[k] - mi) * inv) >> 12); // exact for 0 <= x <= 15, 1 <= D <= 15
		top = top > index[k] ? top : index[k];
	}
	return top;
}

// quant_single_point_d (:546-701): best (endpoints, index) reproducing ONE colour `pt`. Returns per-texel error.
A7_HDN real single_point(const Tables &T, const real pt[4], int clog, const int *bits, int type, int dim, int epo[2][4], int &best_index) {
	const int use_par = type != 0;
	const int npv = 1 << type; // npv_nd for types 0..2
	real err0 = A7_HUGE, err1 = A7_HUGE;
	int idx = 0, idx1 = 0;
	int e0[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
#pragma unroll 1
	for (int pn = 0; pn < npv; pn++) {
		// parity of endpoint 0 / 1 for this vector (identical across channels for CART / SAME_PAR / BCC)
		const int par0 = type == SAME_PAR ? pn : (pn >> 1), par1 = type == SAME_PAR ? pn : (pn & 1);
		const int a0 = use_par ? par0 : 0, a1 = use_par ? par0 + 1 : 2;
		const int b0 = use_par ? par1 : 0, b1 = use_par ? par1 + 1 : 2;
#pragma unroll 1
		for (int i = 0; i < (1 << clog); i++) {
			real t = 0;
			int t1o[4], t2o[4], dr0[4];
#pragma unroll 1
			for (int j = 0; j < dim; j++) {
				real tbest = A7_HUGE;
				int tf = (int) floor(pt[j]), tc = (int) ceil(pt[j]);
				tf = tf < 0 ? 0 : tf;
				tc = tc > 255 ? 255 : tc;
#pragma unroll 1
				for (int t1 = a0; t1 < a1; t1++)
#pragma unroll 1
					for (int t2 = b0; t2 < b1; t2++) {
						const uint32_t ef = T.sp[sp_slot(clog, bits[j], tf, t1, t2, i)] >> 16, ec = T.sp[sp_slot(clog, bits[j], tc, t1, t2, i)] >> 16;
						int dr;
						if (ef > ec) dr = tc;
						else if (ef < ec) dr = tf;
						else dr = (int) floor(pt[j] + 0.5);
						const real k = (real) (T.sp[sp_slot(clog, bits[j], dr, t1, t2, i)] >> 16);
						const real tr = k * k + 2 * k * fabs((real) dr - pt[j]) + ((real) dr - pt[j]) * ((real) dr - pt[j]);
						if (tr < tbest) { tbest = tr; t1o[j] = t1; t2o[j] = t2; dr0[j] = dr; }
					}
				t += tbest;
			}
			if (t < err0) {
				idx = i;
#pragma unroll 1
				for (int j = 0; j < dim; j++) {
					const uint32_t e = T.sp[sp_slot(clog, bits[j], dr0[j], t1o[j], t2o[j], i)];
					e0[0][j] = (int) (e & 255u);
					e0[1][j] = (int) ((e >> 8) & 255u);
				}
				err0 = t;
			}
			if (err0 == 0) break;
		}
		if (err0 < err1) {
			idx1 = idx;
#pragma unroll 1
			for (int j = 0; j < dim; j++) { epo[0][j] = e0[0][j]; epo[1][j] = e0[1][j]; }
			err1 = err0;
		}
		if (err1 == 0) break;
	}
	best_index = idx1;
	return err1;
}

// Handles the "every texel takes the same index" case shared by both shakers (:789-826, :1114-1140)
A7_HDN real shake_single_index(const Tables &T, const real data[][4], int n, bool all_same, const real mean[4], int clog, const int *bits,
															int type, int dim, int *index, int epo[2][4]) {
	int bi;
	real t;
	if (all_same) {
		t = single_point(T, data[0], clog, bits, type, dim, epo, bi) * (real) n;
	} else {
		single_point(T, mean, clog, bits, type, dim, epo, bi);
		t = 0;
#pragma unroll 1
		for (int i = 0; i < n; i++)
#pragma unroll 1
			for (int j = 0; j < dim; j++) {
				const real o = (real) ramp_int(expand_bits(bits[j], epo[0][j]), expand_bits(bits[j], epo[1][j]), bi, clog);
				t += (data[i][j] - o) * (data[i][j] - o);
			}
	}
#pragma unroll 1
	for (int i = 0; i < n; i++) index[i] = bi;
	return t;
}

// Least-squares endpoints for index assignment cidx[] against rounded cluster means (:858-916 == :1160-1219)
A7_HDN void fit_endpoints(const real data[][4], int n, const int *cidx, int Mi_, int dim, real epa[2][4]) {
	real cc[16][4];
	int cnt[16];
#pragma unroll 1
	for (int c = 0; c <= Mi_; c++) {
		cnt[c] = 0;
#pragma unroll 1
		for (int j = 0; j < dim; j++) cc[c][j] = 0;
	}
#pragma unroll 1
	for (int i = 0; i < n; i++) {
#pragma unroll 1
		for (int j = 0; j < dim; j++) cc[cidx[i]][j] += data[i][j];
		cnt[cidx[i]]++;
	}
	// mean + round of every populated cluster (clusters are independent, so the visit order is irrelevant)
#pragma unroll 1
	for (int c = 0; c <= Mi_; c++)
		if (cnt[c])
#pragma unroll 1
			for (int j = 0; j < dim; j++) cc[c][j] = floor(cc[c][j] / (real) cnt[c] + 0.5);
	real im00 = 0, im01 = 0, im11 = 0, rp[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
#pragma unroll 1
	for (int k = 0; k < n; k++) {
		const int a = Mi_ - cidx[k], b = cidx[k];
		im00 += a * a;
		im01 += b * a;
		im11 += b * b;
#pragma unroll 1
		for (int j = 0; j < dim; j++) {
			rp[0][j] += (real) a * cc[b][j];
			rp[1][j] += (real) b * cc[b][j];
		}
	}
	const real dd = im00 * im11 - im01 * im01;
	const real i00 = im11 / dd, i11 = im00 / dd, i01 = -im01 / dd;
#pragma unroll 1
	for (int j = 0; j < dim; j++) {
		epa[0][j] = (i00 * rp[0][j] + i01 * rp[1][j]) * (real) Mi_;
		epa[1][j] = (i01 * rp[0][j] + i11 * rp[1][j]) * (real) Mi_;
	}
}

// ep_shaker_2_d (:703-1053): per-channel window search around the least-squares endpoints with fixed indices,
// combined over parity vectors, then re-clustering; up to 9 rounds. index[] in/out, epo_code out. Returns SSE.
A7_HDN real shake_window(const Tables &T, const real data[][4], int n, int *index_io, int epo_code[2][4], int size, int Mi_, int bits_total,
												 int dim) {
	const int type = bits_total % (2 * dim);
	const int use_par = type != 0;
	const int mb = (bits_total + 2 * dim - 1) / (2 * dim);
	int max_bits[4] = {mb, mb, mb, mb};
	const int clog = ilog2(Mi_ + 1);
	const int C = 1 << clog;
	int index[kMaxEntries];
#pragma unroll 1
	for (int k = 0; k < n; k++) index[k] = index_io[k];
	bool alls = true;
#pragma unroll 1
	for (int i = 1; i < n; i++)
#pragma unroll 1
		for (int j = 0; j < dim; j++) alls = alls && (data[0][j] == data[i][j]);
	real mean[4] = {0, 0, 0, 0};
#pragma unroll 1
	for (int j = 0; j < dim; j++) {
		real m = 0;
#pragma unroll 1
		for (int i = 0; i < n; i++) m += data[i][j];
		mean[j] = m / (real) n;
	}
	real err_o = A7_HUGE;
	int maxTry = 8;
	int done;
	do {
		const int Mi = collapse_indices(index, n);
		if (Mi == 0) {
			int e0[2][4];
			const real t = shake_single_index(T, data, n, alls, mean, clog, max_bits, type, dim, index, e0);
			if (t < err_o) {
#pragma unroll 1
				for (int k = 0; k < n; k++) index_io[k] = index[k];
#pragma unroll 1
				for (int j = 0; j < dim; j++) { epo_code[0][j] = e0[0][j]; epo_code[1][j] = e0[1][j]; }
				err_o = t;
			}
			return err_o;
		}
		int p0 = -1, q0 = -1;
		real err_0 = A7_HUGE;
		int epo_0[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
#pragma unroll 1
		for (int q = 1; q * Mi <= Mi_; q++)
#pragma unroll 1
			for (int p = 0; p <= Mi_ - q * Mi; p++) {
				int cidx[kMaxEntries];
#pragma unroll 1
				for (int k = 0; k < n; k++) cidx[k] = index[k] * q + p;
				real epa[2][4];
				fit_endpoints(data, n, cidx, Mi_, dim, epa);
				real ed[2][2][4];
				int ep2[2][2][2][4];
				const int rr = use_par ? 2 : 1, step = 1 << use_par, top = (1 << mb) - 1;
#pragma unroll 1
				for (int j = 0; j < dim; j++)
#pragma unroll 1
					for (int pp0 = 0; pp0 < rr; pp0++)
#pragma unroll 1
						for (int pp1 = 0; pp1 < rr; pp1++) {
							int lo[2], hi[2];
#pragma unroll 1
							for (int i = 0; i < 2; i++) {
								const int f = endpoint_floor(epa[i][j], mb, use_par, i ? pp1 : pp0);
								lo[i] = f - ((f < (size >> 1) - 1 ? f : (size >> 1) - 1) & ~use_par);
								hi[i] = f + ((top - f < (size >> 1) ? top - f : (size >> 1)) & ~use_par);
							}
							real best = A7_HUGE;
							int b1 = 0, b2 = 0;
#pragma unroll 1
							for (int p1 = lo[0]; p1 <= hi[0]; p1 += step) {
								const int e1 = expand_bits(mb, p1);
#pragma unroll 1
								for (int p2 = lo[1]; p2 <= hi[1]; p2 += step) {
									const int e2 = expand_bits(mb, p2);
									real t = 0;
#pragma unroll 1
									for (int m = n - 1; m >= 0; m--) {
										const real d = (real) ramp_int(e1, e2, cidx[m], clog) - data[m][j];
										t += d * d;
									}
									if (t < best) { best = t; b1 = p1; b2 = p2; }
								}
							}
							ed[pp0][pp1][j] = best;
							ep2[pp0][pp1][0][j] = b1;
							ep2[pp0][pp1][1][j] = b2;
						}
				real err_1 = A7_HUGE;
				int epo_1[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
#pragma unroll 1
				for (int pn = 0; pn < (1 << type); pn++) {
					const int v0 = type == SAME_PAR ? pn : (pn >> 1), v1 = type == SAME_PAR ? pn : (pn & 1);
					real e2 = 0;
#pragma unroll 1
					for (int j = 0; j < dim; j++) e2 += ed[v0][v1][j];
					if (e2 < err_1) {
						err_1 = e2;
#pragma unroll 1
						for (int j = 0; j < dim; j++) { epo_1[0][j] = ep2[v0][v1][0][j]; epo_1[1][j] = ep2[v0][v1][1][j]; }
					}
				}
				if (err_1 <= err_0) {
					err_0 = err_1;
					p0 = p;
					q0 = q;
#pragma unroll 1
					for (int j = 0; j < dim; j++) { epo_0[0][j] = epo_1[0][j]; epo_0[1][j] = epo_1[1][j]; }
				}
			}
		// re-cluster against the chosen endpoints
		int e1[4], e2[4];
#pragma unroll 1
		for (int j = 0; j < dim; j++) { e1[j] = expand_bits(mb, epo_0[0][j]); e2[j] = expand_bits(mb, epo_0[1][j]); }
		int idg[kMaxEntries];
		real err_r = 0;
#pragma unroll 1
		for (int i = 0; i < n; i++) {
			real cmin = A7_HUGE;
			int ci = 0;
#pragma unroll 1
			for (int c = 0; c < C; c++) {
				real t = 0;
#pragma unroll 1
				for (int k = 0; k < dim; k++) {
					const real d = (real) ramp_int(e1[k], e2[k], c, clog) - data[i][k];
					t += d * d;
				}
				if (t < cmin) { cmin = t; ci = c; }
			}
			idg[i] = ci;
			err_r += cmin;
		}
		int change = 0;
#pragma unroll 1
		for (int k = 0; k < n; k++) change = change || (index[k] * q0 + p0 != idg[k]);
		const int better = err_r < err_o;
		if (better) {
#pragma unroll 1
			for (int k = 0; k < n; k++) index_io[k] = index[k] = idg[k];
#pragma unroll 1
			for (int j = 0; j < dim; j++) { epo_code[0][j] = epo_0[0][j]; epo_code[1][j] = epo_0[1][j]; }
			err_o = err_r;
		}
		done = !(change && better);
	} while (!done && maxTry--);
	return err_o;
}

// One (odd, flip) lattice of ep_shaker_d (:1230-1353): the 2x2x2 endpoint-neighbour cube of both endpoints walked
// in the reference's Gray-code order; full re-clustering at each of the 64 corners. Updates (err_1, idx_1) with
// first-strict-minimum semantics.
A7_HD void shake_cube_lattice(const real data[][4], int n, int clog, const int *bits, const real epa[2][4], int use_par, int odd, int flip,
															real &err_1, int *idx_1) {
	const int C = 1 << clog;
	int epi[2][3][2];
#pragma unroll 1
	for (int j = 0; j < 3; j++)
#pragma unroll 1
		for (int i = 0; i < 2; i++) {
			const int f = endpoint_floor(epa[i][j], bits[j], use_par, (odd ^ (flip & i)) & 1);
			const int top = (1 << bits[j]) - 1;
			epi[i][j][0] = f;
			epi[i][j][1] = f + ((top - f < (1 << use_par) ? top - f : (1 << use_par)) & ~use_par);
		}
	real r[3][16];
#pragma unroll 1
	for (int j = 0; j < 3; j++) {
		const int e1 = expand_bits(bits[j], epi[0][j][0]), e2 = expand_bits(bits[j], epi[1][j][0]);
#pragma unroll 1
		for (int c = 0; c < C; c++) r[j][c] = (real) ramp_int(e1, e2, c, clog);
	}
	int s = 0;
#pragma unroll 1
	for (int p1 = 0; p1 < 64; p1++) {
		const int g = p1 & (-p1);
		int j0 = 0, ei0 = 0, ei1 = 0;
#pragma unroll 1
		for (int j = 0; j < 3; j++)
			if (((g >> (2 * j)) & 3) != 0) {
				j0 = j;
				ei0 = ((s ^ g) >> (2 * j)) & 1;
				ei1 = ((s ^ g) >> (2 * j + 1)) & 1;
			}
		s ^= g;
		{
			const int e1 = expand_bits(bits[j0], epi[0][j0][ei0]), e2 = expand_bits(bits[j0], epi[1][j0][ei1]);
#pragma unroll 1
			for (int c = 0; c < C; c++) r[j0][c] = (real) ramp_int(e1, e2, c, clog);
		}
		real err_0 = 0;
		int idx_0[kMaxEntries];
#pragma unroll 1
		for (int i = 0; i < n; i++) {
			real cmin = A7_HUGE;
			int ci = 0;
#pragma unroll 1
			for (int c = 0; c < C; c++) {
				real t = 0;
#pragma unroll 1
				for (int k = 0; k < 3; k++) t += (r[k][c] - data[i][k]) * (r[k][c] - data[i][k]);
				if (t < cmin) { cmin = t; ci = c; }
			}
			idx_0[i] = ci;
			err_0 += cmin;
		}
		if (err_0 < err_1) {
#pragma unroll 1
			for (int i = 0; i < n; i++) idx_1[i] = idx_0[i];
			err_1 = err_0;
		}
	}
}

// ep_shaker_d (:1058-1404), dimension 3 only (its only use). index[] in/out. Returns SSE.
A7_HDN real shake_cube(const Tables &T, const real data[][4], int n, int *index_io, int Mi_, const int *bits, int type) {
	const int dim = 3;
	const int use_par = (type == BCC || type == SAME_PAR), bcc = (type == BCC);
	const int clog = ilog2(Mi_ + 1);
	int index[kMaxEntries];
#pragma unroll 1
	for (int k = 0; k < n; k++) index[k] = index_io[k];
	bool alls = true;
#pragma unroll 1
	for (int i = 1; i < n; i++)
#pragma unroll 1
		for (int j = 0; j < dim; j++) alls = alls && (data[0][j] == data[i][j]);
	real mean[4] = {0, 0, 0, 0};
#pragma unroll 1
	for (int j = 0; j < dim; j++) {
		real m = 0;
#pragma unroll 1
		for (int i = 0; i < n; i++) m += data[i][j];
		mean[j] = m / (real) n;
	}
	real err_o = A7_HUGE;
	int maxTry = 1, done;
	do {
		const int Mi = collapse_indices(index, n);
		if (Mi == 0) {
			int e0[2][4];
			const real t = shake_single_index(T, data, n, alls, mean, clog, bits, type, dim, index, e0);
			if (t < err_o) {
#pragma unroll 1
				for (int k = 0; k < n; k++) index_io[k] = index[k];
				err_o = t;
			}
			return err_o;
		}
		int p0 = -1, q0 = -1;
		real err_2 = A7_HUGE;
		int idx_2[kMaxEntries];
#pragma unroll 1
		for (int k = 0; k < n; k++) idx_2[k] = 0;
#pragma unroll 1
		for (int q = 1; q * Mi <= Mi_; q++)
#pragma unroll 1
			for (int p = 0; p <= Mi_ - q * Mi; p++) {
				int cidx[kMaxEntries];
#pragma unroll 1
				for (int k = 0; k < n; k++) cidx[k] = index[k] * q + p;
				real epa[2][4];
				fit_endpoints(data, n, cidx, Mi_, dim, epa);
				real err_1 = A7_HUGE;
				int idx_1[kMaxEntries];
#pragma unroll 1
				for (int k = 0; k < n; k++) idx_1[k] = 0;
#pragma unroll 1
				for (int odd = 0; odd <= use_par; odd++)
#pragma unroll 1
					for (int flip = 0; flip <= bcc; flip++) shake_cube_lattice(data, n, clog, bits, epa, use_par, odd, flip, err_1, idx_1);
				if (err_1 < err_2) {
#pragma unroll 1
					for (int i = 0; i < n; i++) idx_2[i] = idx_1[i];
					err_2 = err_1;
					p0 = p;
					q0 = q;
				}
			}
		int change = 0;
#pragma unroll 1
		for (int k = 0; k < n; k++) change = change || (index[k] * q0 + p0 != idx_2[k]);
		const int better = err_2 < err_o;
		if (better) {
#pragma unroll 1
			for (int k = 0; k < n; k++) index_io[k] = index[k] = idx_2[k];
			err_o = err_2;
		}
		done = !(change && better);
	} while (!done && maxTry--);
	return err_o;
}

// ---- bit packing -----------------------------------------------------------------------------------------
struct Bits128 {
	uint64_t w[2];
	int pos;
};
A7_HD void put_bits(Bits128 &b, uint32_t v, int n) {
	if (n == 0) return;
	v &= (n >= 32) ? 0xffffffffu : ((1u << n) - 1u);
	const int p = b.pos;
	if (p < 64) {
		b.w[0] |= (uint64_t) v << p;
		if (p + n > 64) b.w[1] |= (uint64_t) v >> (64 - p);
	} else {
		b.w[1] |= (uint64_t) v << (p - 64);
	}
	b.pos = p + n;
}

A7_HD int subset_of(int subsets, int partition, int texel) {
	if (subsets == 2) return (kBc7Part2[partition] >> texel) & 1;
	if (subsets == 3) return (kBc7Part3[partition] >> (2 * texel)) & 3;
	return 0;
}

// Result of one single-index mode: endpoints as integer codes INCLUDING the parity bit in bit 0 (as the shakers
// produce them), indices per subset in subset-local order.
struct SingleIndexResult {
	int partition;
	int ep[3][2][4];
	int idx[3][kMaxEntries];
};

// EncodeSingleIndexBlock (:333-538) + the endpoint packing of :846-881.
// Reference quirk kept: for ONE_PBIT (mode 1) BOTH p-bits are taken from endpoint 1 of the subset (:443-448), after
// the anchor flip; the p-bit that endpoint 0 was searched with is dropped.
A7_HDN void pack_single_index(int mode, const SingleIndexResult &r, uint64_t out[2]) {
	const ModeInfo mi = mode_info(mode);
	const int dim = mi.alpha == 0 ? 3 : 4;
	const int cbits = mi.alpha == 0 ? mi.vector_bits / 3 : mi.vector_bits / 4;
	const int ib = mi.index_bits0;
	int blk[16], cnt[3] = {0, 0, 0};
	int fix[3] = {0, 0, 0};
	if (mi.subsets == 3) { fix[1] = kBc7Anchor3a[r.partition]; fix[2] = kBc7Anchor3b[r.partition]; }
	else if (mi.subsets == 2) fix[1] = kBc7Anchor2[r.partition];
	bool flip[3] = {false, false, false};
#pragma unroll 1
	for (int i = 0; i < 16; i++) {
		const int p = subset_of(mi.subsets, r.partition, i);
		blk[i] = r.idx[p][cnt[p]++];
#pragma unroll 1
		for (int j = 0; j < mi.subsets; j++)
			if (i == fix[j] && (blk[i] & (1 << (ib - 1)))) flip[j] = true;
	}
#pragma unroll 1
	for (int i = 0; i < 16; i++)
		if (flip[subset_of(mi.subsets, r.partition, i)]) blk[i] = ((1 << ib) - 1) - blk[i];
	Bits128 b = {{0, 0}, 0};
	put_bits(b, 1u << mode, mode + 1);
	put_bits(b, (uint32_t) r.partition, mi.partition_bits);
	int col[3][2][4], par[3][2];
#pragma unroll 1
	for (int s = 0; s < mi.subsets; s++) {
		const int a = flip[s] ? 1 : 0;
#pragma unroll 1
		for (int e = 0; e < 2; e++) {
			const int *src = r.ep[s][e ^ a];
			par[s][e] = 0;
#pragma unroll 1
			for (int k = 0; k < 4; k++) col[s][e][k] = (mi.parity != CART) ? (src[k] >> 1) : src[k];
			if (mi.parity != CART) par[s][e] = src[0] & 1;
		}
		if (mi.parity == SAME_PAR) par[s][0] = par[s][1]; // ONE_PBIT quirk
	}
#pragma unroll 1
	for (int k = 0; k < dim; k++)
#pragma unroll 1
		for (int s = 0; s < mi.subsets; s++)
#pragma unroll 1
			for (int e = 0; e < 2; e++) put_bits(b, (uint32_t) col[s][e][k], cbits);
	if (mi.parity == SAME_PAR)
#pragma unroll 1
		for (int s = 0; s < mi.subsets; s++) put_bits(b, (uint32_t) par[s][0], 1);
	else if (mi.parity == BCC)
#pragma unroll 1
		for (int s = 0; s < mi.subsets; s++) { put_bits(b, (uint32_t) par[s][0], 1); put_bits(b, (uint32_t) par[s][1], 1); }
#pragma unroll 1
	for (int i = 0; i < 16; i++) {
		const int p = subset_of(mi.subsets, r.partition, i);
		put_bits(b, (uint32_t) blk[i], (i == fix[p]) ? ib - 1 : ib);
	}
	out[0] = b.w[0];
	out[1] = b.w[1];
}

// EncodeDualIndexBlock (:902-1056)
A7_HDN void pack_dual_index(int mode, int index_selection, int rotation, int ep[2][2][4], int idx[2][16], uint64_t out[2]) {
	const ModeInfo mi = mode_info(mode);
	Bits128 b = {{0, 0}, 0};
	put_bits(b, 1u << mode, mode + 1);
	put_bits(b, (uint32_t) rotation, mi.rotation_bits);
	put_bits(b, index_selection ? 1u : 0u, mi.index_mode_bits);
	int ibits[2];
	ibits[0] = index_selection ? mi.index_bits1 : mi.index_bits0;
	ibits[1] = index_selection ? mi.index_bits0 : mi.index_bits1;
#pragma unroll 1
	for (int i = 0; i < 2; i++)
		if (idx[i][0] & (1 << (ibits[i] - 1))) {
#pragma unroll 1
			for (int j = 0; j < 16; j++) idx[i][j] = ((1 << ibits[i]) - 1) - idx[i][j];
#pragma unroll 1
			for (int k = 0; k < 4; k++) { const int t = ep[i][0][k]; ep[i][0][k] = ep[i][1][k]; ep[i][1][k] = t; }
		}
	const int vbits = mi.vector_bits / 3;
#pragma unroll 1
	for (int c = 0; c < 4; c++)
#pragma unroll 1
		for (int e = 0; e < 2; e++) {
			if (c != 3) put_bits(b, (uint32_t) ep[0][e][c], vbits);
			else put_bits(b, (uint32_t) ep[1][e][0], mi.scalar_bits);
		}
#pragma unroll 1
	for (int i = 0; i < 2; i++) {
		const int sel = index_selection ? (i ^ 1) : i;
#pragma unroll 1
		for (int j = 0; j < 16; j++) put_bits(b, (uint32_t) idx[sel][j], j == 0 ? ibits[sel] - 1 : ibits[sel]);
	}
	out[0] = b.w[0];
	out[1] = b.w[1];
}

// ---- block level (serial orchestration; the CUDA kernel spreads the same tasks over lanes) ---------------
struct BlockInput {
	real px[16][4]; // 0..255
	real pxc[4][16]; // the same, channel-major (quantiser input, see QuantIO)
	uint32_t mode_mask; // after the filter of :1340-1380
};

// the mode filter of CompressBlock (:1340-1380)
A7_HD uint32_t filter_modes(uint32_t valid_mode_mask, bool needs_alpha, bool zero_one, bool solid) {
	uint32_t mask = valid_mode_mask ? valid_mode_mask : 0xCFu;
#pragma unroll 1
	for (int m = 0; m < 8; m++) {
		if (!(mask & (1u << m))) continue;
		const int at = mode_info(m).alpha;
		if (needs_alpha && at == 0) mask &= ~(1u << m);
		if (!solid && !needs_alpha && at == 1) mask &= ~(1u << m);
		if (needs_alpha && zero_one && at == 1) mask &= ~(1u << m);
	}
	return mask;
}

// CompressBlock's set-up (:1296-1380): scale to 0..255, alpha classification, mode filter
A7_HDN void prepare_block(const float in[64], uint32_t valid_mode_mask, BlockInput &B) {
	bool needs_alpha = false, zero_one = false;
	real mn[4] = {A7_HUGE, A7_HUGE, A7_HUGE, A7_HUGE}, mx[4] = {0, 0, 0, 0};
#pragma unroll 1
	for (int i = 0; i < 16; i++) {
		const float a = in[i * 4 + 3];
		if (a < 1.0) needs_alpha = true;
		else if (((double) a >= 0.99999) || ((double) a < 0.00001)) zero_one = true;
#pragma unroll 1
		for (int j = 0; j < 4; j++) {
			const real v = (real) (in[i * 4 + j] * 255.0f);
			B.px[i][j] = v;
			B.pxc[j][i] = v;
			mn[j] = v < mn[j] ? v : mn[j];
			mx[j] = v > mx[j] ? v : mx[j];
		}
	}
	real range = mx[0] - mn[0];
#pragma unroll 1
	for (int j = 1; j < 4; j++) range = range > (mx[j] - mn[j]) ? range : (mx[j] - mn[j]);
	B.mode_mask = filter_modes(valid_mode_mask, needs_alpha, zero_one, range < 1e-10);
}

A7_HD void gather_subset(const BlockInput &B, int subsets, int partition, int subset, int dim, real out[][4], int &n) {
	n = 0;
#pragma unroll 1
	for (int i = 0; i < 16; i++)
		if (subset_of(subsets, partition, i) == subset) {
#pragma unroll 1
			for (int j = 0; j < 4; j++) out[n][j] = j < dim ? B.px[i][j] : 0;
			n++;
		}
}

struct ShakeParams {
	int bits[4]; // per channel incl. parity (0..2), total for both endpoints (3)
	int shake_size, clusters, dim, parity;
};
A7_HD ShakeParams single_index_shake_params(int mode) { // :651-707 at quality 1
	const ModeInfo mi = mode_info(mode);
	ShakeParams s;
	s.dim = mi.alpha == 0 ? 3 : 4;
	const int cb = mi.alpha == 0 ? mi.vector_bits / 3 : mi.vector_bits / 4;
	const int par = mi.parity != CART ? 1 : 0;
	s.bits[0] = s.bits[1] = s.bits[2] = cb + par;
	s.bits[3] = 2 * cb * s.dim + (mi.parity == BCC ? 2 : (mi.parity == SAME_PAR ? 1 : 0));
	int ss = 8 - (int) floor(1.5 * mi.index_bits0);
	ss = ss < 2 ? 2 : (ss > 6 ? 6 : ss);
	if (par) ss += 2;
	s.shake_size = ss;
	s.clusters = 1 << mi.index_bits0;
	s.parity = mi.parity;
	return s;
}

} // namespace amd7
} // namespace b200ic
