// bc7amd_int.cuh -- exact integer forms of the AMD BC7 endpoint shakers for 8-bit sources.
//
// For R8/RG8/RGB8/RGBA8 sources every texel component the encoder sees is an exact integer 0..255
// ((x / 255.0f) * 255.0f == x for all 256 x), and every ramp value is an integer, so the error sums of
// ep_shaker_d / ep_shaker_2_d (reference src/amd_shake.cpp:703-1404) are integers below 2^22: INT32 arithmetic gives
// the reference's FP64 results EXACTLY, in any summation order.  That turns the reference's hottest loop (82 % of its
// CPU time) into four native sm_100a instructions per (texel, ramp entry) for all channels at once:
//     VABSDIFF4.U8 (|palette - texel| per byte), IDP.4A.U8.U8 (sum of squares), LEA (error<<4 | entry), VIMNMX.U32
// and "first strict minimum in scan order" becomes a plain unsigned min of (error, scan position) keys.
// The least-squares endpoint fit, the floor search on the endpoint lattice and the single-colour path keep the
// reference's FP64 operations (their inputs are exact integers, so they reproduce it bit for bit as well).
// The functions are drop-in equivalents of shake_cube / shake_window in bc7amd_core.cuh (tests/hostbuild checks
// them against those and against the compiled reference).
#pragma once
#include "bc7amd_core.cuh"

namespace b200ic {
namespace amd7 {

A7_HD uint32_t sq_dist4(uint32_t a, uint32_t b) { // sum over the 4 bytes of (a_k - b_k)^2
#if defined(__CUDA_ARCH__)
	const uint32_t ad = __vabsdiffu4(a, b);
	return __dp4a(ad, ad, 0u);
#else
	uint32_t t = 0;
#pragma unroll 1
	for (int k = 0; k < 4; k++) {
		const int d = (int) ((a >> (8 * k)) & 255u) - (int) ((b >> (8 * k)) & 255u);
		t += (uint32_t) (d * d);
	}
	return t;
#endif
}
A7_HD uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }
// byte `pos` (0..3) of `into` replaced by byte c (0..7) of the ramp-table word `tab`: one PRMT on the GPU
template <int POS> A7_HD uint32_t put_ramp_byte(uint32_t into, uint64_t tab, int c) {
#if defined(__CUDA_ARCH__)
	const uint32_t w = c < 4 ? (uint32_t) tab : (uint32_t) (tab >> 32);
	const uint32_t keep = 0x3210u & ~(0xfu << (4 * POS));
	return __byte_perm(into, w, keep | ((4u + (uint32_t) (c & 3)) << (4 * POS)));
#else
	return (into & ~(0xffu << (8 * POS))) | ((uint32_t) ((tab >> (8 * c)) & 255u) << (8 * POS));
#endif
}
A7_HD uint32_t byte_of(uint64_t v, int i) { return (uint32_t) (v >> (8 * i)) & 255u; }

// ramp values of all C entries between two EXPANDED endpoints, one byte each (C <= 8 fits a u64; C = 16 uses two)
template <int CLOG> A7_HD void ramp_bytes(int e1, int e2, uint64_t out[(1 << CLOG) > 8 ? 2 : 1]) {
	constexpr int C = 1 << CLOG, D = C - 1;
	int num = 2 * D * e1 + D;
	const int step = 2 * (e2 - e1);
	out[0] = 0;
	if (C > 8) out[1] = 0;
#pragma unroll
	for (int c = 0; c < C; c++) {
		const uint32_t v = (uint32_t) (num / (2 * D));
		if (c < 8) out[0] |= (uint64_t) v << (8 * c);
		else out[1] |= (uint64_t) v << (8 * (c - 8));
		num += step;
	}
}

// ---- per-cluster statistics in registers --------------------------------------------------------------------------
// One 64-bit word per cluster: count in bits 0..7, channel sums (<= 16 * 255) in 14-bit fields at bits 8 + 14 j.  The
// words are indexed statically (loops over the cluster are unrolled), so they live in registers: no local-memory
// traffic in the innermost shaker loops.
template <int CLOG> struct ClusterAcc {
	uint64_t a[1 << CLOG];
};
A7_HD uint64_t acc_entry(uint32_t texel) { // one texel as an accumulator increment
	const uint32_t lo = 1u | ((texel & 255u) << 8) | (((texel >> 8) & 255u) << 22);
	const uint32_t hi = (((texel >> 16) & 255u) << 4) | ((texel >> 24) << 18);
	return (uint64_t) lo | ((uint64_t) hi << 32);
}
A7_HD int acc_count(uint64_t a) { return (int) (a & 255u); }
A7_HD int acc_sum(uint64_t a, int j) { return (int) ((a >> (8 + 14 * j)) & 0x3fffu); }
template <int CLOG> A7_HD void cluster_acc(const uint32_t *d, int n, uint64_t collapsed, int q, int p, ClusterAcc<CLOG> &cs) {
	constexpr int C = 1 << CLOG;
#pragma unroll
	for (int c = 0; c < C; c++) cs.a[c] = 0;
#pragma unroll 1
	for (int k = 0; k < n; k++) {
		const int ck = (int) ((collapsed >> (4 * k)) & 15u) * q + p;
		const uint64_t e = acc_entry(d[k]);
#pragma unroll
		for (int c = 0; c < C; c++) cs.a[c] += (ck == c) ? e : 0ull;
	}
}
// floor(x / d) for 1 <= d <= 16, 0 <= x < 32768: (x * (2^20 / d + 1)) >> 20
static B7T_QUAL uint32_t kInv20[17] = {0,     1048577, 524289, 349526, 262145, 209716, 174763, 149797, 131073,
																116509, 104858,  95326,  87382,  80660,  74899,  69906,  65537};
// Least-squares endpoints (:1167-1199 / :852-884) from the cluster statistics.  The cluster means
// floor(sum / cnt + 0.5) are exact small integers (the rounding cannot flip: the exact value is a multiple of
// 1 / (2 cnt) >= 1/32 away from any other integer, and m - 0.5 is exactly representable), and every sum the reference
// forms over the texels is a sum of integers, so it is done per cluster in INT32; only the 2x2 solve keeps the
// reference's FP64 operations.
template <int CLOG> A7_HD void fit_endpoints_acc(const ClusterAcc<CLOG> &cs, int dim, real epa[2][4]) {
	constexpr int C = 1 << CLOG, Mi_ = C - 1;
	int im00 = 0, im01 = 0, im11 = 0, rp0[4] = {0, 0, 0, 0}, rp1[4] = {0, 0, 0, 0};
#pragma unroll
	for (int c = 0; c < C; c++) {
		const int cnt = acc_count(cs.a[c]);
		if (cnt) {
			const int a = Mi_ - c, b = c;
			im00 += cnt * a * a;
			im01 += cnt * a * b;
			im11 += cnt * b * b;
			const uint32_t inv = kInv20[cnt];
#pragma unroll
			for (int j = 0; j < 4; j++)
				if (j < dim) {
					const int cc = (int) (((uint64_t) (uint32_t) (2 * acc_sum(cs.a[c], j) + cnt) * inv) >> 21); // floor(x / cnt) / 2 = floor(x / (2 cnt)); 64-bit product: up to 2^33
					rp0[j] += cnt * a * cc;
					rp1[j] += cnt * b * cc;
				}
		}
	}
	const real d00 = (real) im00, d01 = (real) im01, d11 = (real) im11;
	const real dd = d00 * d11 - d01 * d01;
	const real i00 = d11 / dd, i11 = d00 / dd, i01 = -d01 / dd;
#pragma unroll
	for (int j = 0; j < 4; j++)
		if (j < dim) {
			epa[0][j] = (i00 * (real) rp0[j] + i01 * (real) rp1[j]) * (real) Mi_;
			epa[1][j] = (i01 * (real) rp0[j] + i11 * (real) rp1[j]) * (real) Mi_;
		}
}

struct U8Subset {
	uint32_t d[kMaxEntries]; // packed texels, channel j in byte j (unused channels 0)
	int n;
	bool all_same;
	real mean[4];
};
A7_HD void make_u8_subset(const real data[][4], int n, int dim, U8Subset &S) {
	S.n = n;
	int sum[4] = {0, 0, 0, 0};
	bool same = true;
#pragma unroll 1
	for (int i = 0; i < n; i++) {
		uint32_t v = 0;
#pragma unroll 1
		for (int j = 0; j < dim; j++) {
			const int b = (int) data[i][j];
			v |= (uint32_t) b << (8 * j);
			sum[j] += b;
		}
		S.d[i] = v;
		same = same && (v == S.d[0]);
	}
	S.all_same = same;
#pragma unroll 1
	for (int j = 0; j < 4; j++) S.mean[j] = j < dim ? (real) sum[j] / (real) n : 0;
}

// the Mi == 0 path of both shakers on packed data (see shake_single_index)
A7_HDN real shake_single_index_u8(const Tables &T, const U8Subset &S, int clog, const int *bits, int type, int dim, int *index, int epo[2][4]) {
	int bi;
	real t;
	if (S.all_same) {
		real pt[4];
#pragma unroll 1
		for (int j = 0; j < 4; j++) pt[j] = (real) ((S.d[0] >> (8 * j)) & 255u);
		t = single_point(T, pt, clog, bits, type, dim, epo, bi) * (real) S.n;
	} else {
		single_point(T, S.mean, clog, bits, type, dim, epo, bi);
		uint32_t o = 0;
#pragma unroll 1
		for (int j = 0; j < dim; j++) o |= (uint32_t) ramp_int(expand_bits(bits[j], epo[0][j]), expand_bits(bits[j], epo[1][j]), bi, clog) << (8 * j);
		uint32_t e = 0;
#pragma unroll 1
		for (int i = 0; i < S.n; i++) e += sq_dist4(S.d[i], o);
		t = (real) e;
	}
#pragma unroll 1
	for (int i = 0; i < S.n; i++) index[i] = bi;
	return t;
}

A7_HD int gray_position(int s) { // p1 with gray(p1) == s, 6 bits
	s ^= s >> 1;
	s ^= s >> 2;
	s ^= s >> 4;
	return s & 63;
}

// All lattices of ep_shaker_d for one (q, p): returns min over (lattice, corner) of err << 8 | lattice << 6 | gray position,
// and that corner's index assignment (4 bits per texel).
// i0..i1 = the texels this caller sums over.  On the GPU an item with many texels is shared by two adjacent lanes
// (halves of its texel list): `pair_mask` != 0 names the lanes that meet at the exchange after every corner, `paired`
// says whether this lane's partner (lane ^ 1) holds the other half; the two then see the same keys and keep the index
// nibbles of their own texels (merged by the caller).
template <int CLOG>
A7_HD void cube_search_u8(const uint32_t *d, int n, const int *bits, const real epa[2][4], int use_par, int bcc, int z0, int z1, uint32_t &best_key,
													 uint64_t &best_idx, int i0 = 0, int i1 = -1, unsigned pair_mask = 0, bool paired = false) {
	constexpr int C = 1 << CLOG;
	if (i1 < 0) i1 = n;
	(void) pair_mask;
	(void) paired;
	// floor of each ideal endpoint on the parity-0 and parity-1 lattice
	int fl[2][3][2];
#pragma unroll 1
	for (int e = 0; e < 2; e++)
#pragma unroll 1
		for (int k = 0; k < 3; k++)
#pragma unroll 1
			for (int par = 0; par <= use_par; par++) fl[e][k][par] = endpoint_floor(epa[e][k], bits[k], use_par, par);
	// The walk only needs every corner's ERROR (3 instructions per texel and ramp entry: |palette - texel| per byte,
	// sum of squares, running minimum); the index vector is wanted for the winning corner alone, whose palette is kept
	// and re-scanned once at the end with the (distance << 4 | entry) keys that make the lowest entry win ties.
	uint32_t win_pal[C];
#pragma unroll
	for (int c = 0; c < C; c++) win_pal[c] = 0;
	bool improved = false;
	int lattice = 0;
#pragma unroll 1
	for (int odd = 0; odd <= use_par; odd++)
#pragma unroll 1
		for (int flip = 0; flip <= bcc; flip++, lattice++) {
			uint64_t tab[3][4]; // [channel][ei0 + 2*ei1] -> C ramp bytes
#pragma unroll 1
			for (int k = 0; k < 3; k++) {
				int ep[2][2];
#pragma unroll 1
				for (int e = 0; e < 2; e++) {
					const int f = fl[e][k][(odd ^ (flip & e)) & 1];
					const int top = (1 << bits[k]) - 1;
					ep[e][0] = expand_bits(bits[k], f);
					ep[e][1] = expand_bits(bits[k], f + ((top - f < (1 << use_par) ? top - f : (1 << use_par)) & ~use_par));
				}
#pragma unroll 1
				for (int x = 0; x < 4; x++) ramp_bytes<CLOG>(ep[0][x & 1], ep[1][x >> 1], &tab[k][x]);
			}
#pragma unroll 1
			for (int z = z0; z < z1; z++)
#pragma unroll 1
				for (int y = 0; y < 4; y++) {
					uint32_t pzy[C];
					const uint64_t tz = tab[2][z], ty = tab[1][y];
#pragma unroll
					for (int c = 0; c < C; c++) pzy[c] = put_ramp_byte<2>(put_ramp_byte<1>(0u, ty, c), tz, c);
#pragma unroll
					for (int x = 0; x < 4; x++) {
						uint32_t pal[C];
						const uint64_t tx = tab[0][x];
#pragma unroll
						for (int c = 0; c < C; c++) pal[c] = put_ramp_byte<0>(pzy[c], tx, c);
						uint32_t err = 0;
#pragma unroll 1
						for (int i = i0; i < i1; i++) {
							const uint32_t di = d[i];
							uint32_t m = 0xffffffffu;
#pragma unroll
							for (int c = 0; c < C; c++) m = umin32(m, sq_dist4(pal[c], di));
							err += m;
						}
#if defined(__CUDA_ARCH__)
						if (pair_mask) {
							const uint32_t other = __shfl_xor_sync(pair_mask, err, 1);
							if (paired) err += other;
						}
#endif
						const uint32_t key = (err << 8) | ((uint32_t) lattice << 6) | (uint32_t) gray_position(x | (y << 2) | (z << 4));
						if (key < best_key) {
							best_key = key;
							improved = true;
#pragma unroll
							for (int c = 0; c < C; c++) win_pal[c] = pal[c];
						}
					}
				}
		}
	if (improved) {
		uint64_t idx = 0;
#pragma unroll 1
		for (int i = i0; i < i1; i++) {
			const uint32_t di = d[i];
			uint32_t m = 0xffffffffu;
#pragma unroll
			for (int c = 0; c < C; c++) m = umin32(m, (sq_dist4(win_pal[c], di) << 4) | (uint32_t) c);
			idx |= (uint64_t) (m & 15u) << (4 * i);
		}
		best_idx = idx;
	}
}

// The same search as cube_search_u8 with an exact branch-and-bound.  For a corner (x, y, z) of a lattice the error is
// sum_i min_c [ r(c;x) + g(c;y) + b(c;z) ] >= sum_i min_c r + sum_i min_c g + sum_i min_c b = LB0[x] + LB1[y] + LB2[z]:
// twelve per-channel minima per lattice bound all 64 corners.  A corner whose bound is strictly above the best error
// found so far (over all lattices of the item) can neither win nor tie, so it is skipped; the survivors are evaluated
// exactly as before and the result -- min over corners of err << 8 | lattice << 6 | gray position -- is identical.
// The corner with the smallest bound goes first; the survivors of a lane are popped from a 64-bit mask, so the lanes
// of a warp stay in the same code while working on different corners.
#ifndef A7_STATS_CORNERS
#define A7_STATS_CORNERS(evaluated, total)
#endif
template <int CLOG>
A7_HD void cube_corner_u8(const uint32_t *d, int n, const uint64_t tab[3][4], int x, int y, int z, int lattice, uint32_t &best_key, uint32_t *win_pal) {
	constexpr int C = 1 << CLOG;
	uint32_t pal[C];
	const uint64_t t0 = tab[0][x], t1 = tab[1][y], t2 = tab[2][z];
#pragma unroll
	for (int c = 0; c < C; c++) pal[c] = byte_of(t0, c) | (byte_of(t1, c) << 8) | (byte_of(t2, c) << 16);
	uint32_t err = 0;
#pragma unroll 1
	for (int i = 0; i < n; i++) {
		const uint32_t di = d[i];
		uint32_t m = 0xffffffffu;
#pragma unroll
		for (int c = 0; c < C; c++) m = umin32(m, sq_dist4(pal[c], di));
		err += m;
	}
	const uint32_t key = (err << 8) | ((uint32_t) lattice << 6) | (uint32_t) gray_position(x | (y << 2) | (z << 4));
	if (key < best_key) {
		best_key = key;
#pragma unroll
		for (int c = 0; c < C; c++) win_pal[c] = pal[c];
	}
}
// index vector of the texels against a palette: the lowest entry wins ties (keys distance << 4 | entry)
template <int CLOG> A7_HD uint64_t palette_indices_u8(const uint32_t *d, int n, const uint32_t *pal) {
	constexpr int C = 1 << CLOG;
	uint64_t idx = 0;
#pragma unroll 1
	for (int i = 0; i < n; i++) {
		const uint32_t di = d[i];
		uint32_t m = 0xffffffffu;
#pragma unroll
		for (int c = 0; c < C; c++) m = umin32(m, (sq_dist4(pal[c], di) << 4) | (uint32_t) c);
		idx |= (uint64_t) (m & 15u) << (4 * i);
	}
	return idx;
}
template <int CLOG>
A7_HD void cube_search_pruned_u8(const uint32_t *d, int n, const int *bits, const real epa[2][4], int use_par, int bcc, int z0, int z1,
																 uint32_t &best_key, uint64_t &best_idx) {
	constexpr int C = 1 << CLOG;
	int fl[2][3][2];
#pragma unroll 1
	for (int e = 0; e < 2; e++)
#pragma unroll 1
		for (int k = 0; k < 3; k++)
#pragma unroll 1
			for (int par = 0; par <= use_par; par++) fl[e][k][par] = endpoint_floor(epa[e][k], bits[k], use_par, par);
	uint32_t win_pal[C];
#pragma unroll
	for (int c = 0; c < C; c++) win_pal[c] = 0;
	const uint32_t key_in = best_key;
	int lattice = 0;
#pragma unroll 1
	for (int odd = 0; odd <= use_par; odd++)
#pragma unroll 1
		for (int flip = 0; flip <= bcc; flip++, lattice++) {
			uint64_t tab[3][4]; // [channel][ei0 + 2*ei1] -> C ramp bytes
			uint32_t lb[3][4];  // [channel][combo] -> sum over the texels of the smallest squared distance to the ramp
#pragma unroll 1
			for (int k = 0; k < 3; k++) {
				int ep[2][2];
#pragma unroll 1
				for (int e = 0; e < 2; e++) {
					const int f = fl[e][k][(odd ^ (flip & e)) & 1];
					const int top = (1 << bits[k]) - 1;
					ep[e][0] = expand_bits(bits[k], f);
					ep[e][1] = expand_bits(bits[k], f + ((top - f < (1 << use_par) ? top - f : (1 << use_par)) & ~use_par));
				}
#pragma unroll 1
				for (int x = 0; x < 4; x++) {
					ramp_bytes<CLOG>(ep[0][x & 1], ep[1][x >> 1], &tab[k][x]);
					const uint64_t rv = tab[k][x];
					uint32_t sum = 0;
#pragma unroll 1
					for (int i = 0; i < n; i++) {
						const int v = (int) ((d[i] >> (8 * k)) & 255u);
						int m = 255;
#pragma unroll
						for (int c = 0; c < C; c++) {
							int a = (int) byte_of(rv, c) - v;
							a = a < 0 ? -a : a;
							m = a < m ? a : m;
						}
						sum += (uint32_t) (m * m);
					}
					lb[k][x] = sum;
				}
			}
			// seed: the corner with the smallest bound (first such in x, y, z order)
			int sx = 0, sy = 0, sz = z0;
#pragma unroll 1
			for (int x = 1; x < 4; x++) {
				if (lb[0][x] < lb[0][sx]) sx = x;
				if (lb[1][x] < lb[1][sy]) sy = x;
			}
#pragma unroll 1
			for (int z = z0 + 1; z < z1; z++)
				if (lb[2][z] < lb[2][sz]) sz = z;
			int evaluated = 0;
			if (lb[0][sx] + lb[1][sy] + lb[2][sz] <= (best_key >> 8)) {
				cube_corner_u8<CLOG>(d, n, tab, sx, sy, sz, lattice, best_key, win_pal);
				evaluated++;
			}
			uint64_t mask = 0;
#pragma unroll 1
			for (int z = z0; z < z1; z++)
#pragma unroll 1
				for (int y = 0; y < 4; y++) {
					const uint32_t t = lb[2][z] + lb[1][y];
#pragma unroll
					for (int x = 0; x < 4; x++) mask |= (uint64_t) (t + lb[0][x] <= (best_key >> 8) ? 1 : 0) << (x | (y << 2) | (z << 4));
				}
			mask &= ~(1ull << (sx | (sy << 2) | (sz << 4)));
#pragma unroll 1
			while (mask) {
#if defined(__CUDA_ARCH__)
				const int cnr = __ffsll((long long) mask) - 1;
#else
				const int cnr = __builtin_ctzll(mask);
#endif
				mask &= mask - 1;
				const int x = cnr & 3, y = (cnr >> 2) & 3, z = cnr >> 4;
				if (lb[0][x] + lb[1][y] + lb[2][z] <= (best_key >> 8)) {
					cube_corner_u8<CLOG>(d, n, tab, x, y, z, lattice, best_key, win_pal);
					evaluated++;
				}
			}
			A7_STATS_CORNERS(evaluated, 16 * (z1 - z0));
		}
	if (best_key != key_in) best_idx = palette_indices_u8<CLOG>(d, n, win_pal);
}

// Building blocks of the two-phase form of the pruned search used by the CUDA kernel (bc7amd.cu, cube_batch): phase A
// sets up every lattice of an item (ramp tables, per-channel bounds, the seed corner), phase B evaluates the surviving
// corners of ALL items of a batch as one evenly divided list of units.
A7_HD void cube_floors(const real epa[2][4], const int *bits, int use_par, int fl[2][3][2]) {
#pragma unroll 1
	for (int e = 0; e < 2; e++)
#pragma unroll 1
		for (int k = 0; k < 3; k++)
#pragma unroll 1
			for (int par = 0; par <= use_par; par++) fl[e][k][par] = endpoint_floor(epa[e][k], bits[k], use_par, par);
}
// tab[k * 4 + x] = C ramp bytes of channel k for endpoint combination x; lb[k * 4 + x] = sum over the texels of the
// smallest squared distance of channel k to that ramp
template <int CLOG>
A7_HD void cube_lattice_setup(const uint32_t *d, int n, const int *bits, const int fl[2][3][2], int use_par, int odd, int flip, uint64_t *tab,
															uint32_t *lb, uint32_t *ep_packed) {
	constexpr int C = 1 << CLOG;
#pragma unroll 1
	for (int k = 0; k < 3; k++) {
		int ep[2][2];
#pragma unroll
		for (int e = 0; e < 2; e++) {
			const int f = fl[e][k][(odd ^ (flip & e)) & 1];
			const int top = (1 << bits[k]) - 1;
			ep[e][0] = expand_bits(bits[k], f);
			ep[e][1] = expand_bits(bits[k], f + ((top - f < (1 << use_par) ? top - f : (1 << use_par)) & ~use_par));
		}
		ep_packed[k] = (uint32_t) ep[0][0] | ((uint32_t) ep[0][1] << 8) | ((uint32_t) ep[1][0] << 16) | ((uint32_t) ep[1][1] << 24);
#pragma unroll
		for (int x = 0; x < 4; x++) {
			uint64_t rv[1];
			ramp_bytes<CLOG>(ep[0][x & 1], ep[1][x >> 1], rv);
			tab[k * 4 + x] = rv[0];
			uint32_t sum = 0;
#pragma unroll 1
			for (int i = 0; i < n; i++) {
				const int v = (int) ((d[i] >> (8 * k)) & 255u);
				int m = 255;
#pragma unroll
				for (int c = 0; c < C; c++) {
					int a = (int) byte_of(rv[0], c) - v;
					a = a < 0 ? -a : a;
					m = a < m ? a : m;
				}
				sum += (uint32_t) (m * m);
			}
			lb[k * 4 + x] = sum;
		}
	}
}
// the ramp tables of a lattice from its packed expanded endpoints (see cube_lattice_setup)
template <int CLOG> A7_HD void cube_tab_from_ep(const uint32_t *ep_packed, uint64_t *tab) {
#pragma unroll 1
	for (int k = 0; k < 3; k++) {
		const uint32_t e = ep_packed[k];
#pragma unroll
		for (int x = 0; x < 4; x++) {
			uint64_t rv[1];
			ramp_bytes<CLOG>((int) ((e >> (8 * (x & 1))) & 255u), (int) ((e >> (16 + 8 * (x >> 1))) & 255u), rv);
			tab[k * 4 + x] = rv[0];
		}
	}
}
A7_HD uint32_t cube_corner_bound(const uint32_t *lb, int corner) { return lb[corner & 3] + lb[4 + ((corner >> 2) & 3)] + lb[8 + (corner >> 4)]; }
A7_HD int cube_seed_corner(const uint32_t *lb) { // smallest bound, first in (x, y, z) order
	int sx = 0, sy = 0, sz = 0;
#pragma unroll
	for (int x = 1; x < 4; x++) {
		if (lb[x] < lb[sx]) sx = x;
		if (lb[4 + x] < lb[4 + sy]) sy = x;
		if (lb[8 + x] < lb[8 + sz]) sz = x;
	}
	return sx | (sy << 2) | (sz << 4);
}
A7_HD uint64_t cube_survivors(const uint32_t *lb, uint32_t best_err) {
	uint64_t mask = 0;
#pragma unroll 1
	for (int z = 0; z < 4; z++)
#pragma unroll
		for (int y = 0; y < 4; y++) {
			const uint32_t t = lb[8 + z] + lb[4 + y];
#pragma unroll
			for (int x = 0; x < 4; x++) mask |= (uint64_t) (t + lb[x] <= best_err ? 1 : 0) << (x | (y << 2) | (z << 4));
		}
	return mask;
}

// ---- (q, p) re-indexings of a collapsed index set (the double loop of :1144-1146 / :835-836), as an ordered list
A7_HD int qp_count(int Mi, int Mi_) {
	int c = 0;
#pragma unroll 1
	for (int q = 1; q * Mi <= Mi_; q++) c += Mi_ - q * Mi + 1;
	return c;
}
A7_HD void qp_decode(int ord, int Mi, int Mi_, int &q, int &p) { // (ord < qp_count(Mi, Mi_); anything else ends at the last q)
	p = ord;
#pragma unroll 1
	for (q = 1; (q + 1) * Mi <= Mi_; q++) {
		const int cnt = Mi_ - q * Mi + 1;
		if (p < cnt) return;
		p -= cnt;
	}
}
// One work item of ep_shaker_d: texels d[], collapsed indices (4 bits each), one (q, p), z-slices [z0, z1) of every
// lattice. key = err << 8 | lattice << 6 | gray position (minimum = first strict minimum in the reference's order).
template <int CLOG>
A7_HDN void cube_item_u8(const uint32_t *d, int n, uint64_t collapsed, int q, int p, const int *bits, int type, int z0, int z1, uint32_t &key,
												uint64_t &idx, bool prune = false, int half = -1, unsigned pair_mask = 0) {
	ClusterAcc<CLOG> cs;
	cluster_acc<CLOG>(d, n, collapsed, q, p, cs);
	real epa[2][4];
	fit_endpoints_acc<CLOG>(cs, 3, epa);
	key = 0xffffffffu;
	idx = 0;
	if (prune) {
		cube_search_pruned_u8<CLOG>(d, n, bits, epa, (type == BCC || type == SAME_PAR), (type == BCC), z0, z1, key, idx);
		return;
	}
	// half = -1: all texels; 0 / 1: the first ceil(n / 2) texels / the rest, the partner lane holding the other half
	const int mid = (n + 1) >> 1;
	const int i0 = half == 1 ? mid : 0, i1 = half == 0 ? mid : n;
	cube_search_u8<CLOG>(d, n, bits, epa, (type == BCC || type == SAME_PAR), (type == BCC), z0, z1, key, idx, i0, i1, pair_mask, half >= 0);
#if defined(__CUDA_ARCH__)
	if (pair_mask) {
		const uint32_t lo = __shfl_xor_sync(pair_mask, (uint32_t) idx, 1), hi = __shfl_xor_sync(pair_mask, (uint32_t) (idx >> 32), 1);
		if (half >= 0) idx |= (uint64_t) lo | ((uint64_t) hi << 32);
	}
#endif
}

// ep_find_floor (:351-367) in closed form.  The reference bisects for the largest lattice index j >= 1 with
// expand(code(j)) <= v, code(j) = (j << use_par) + odd, and answers j = 0 when there is none.  v >= x <=> floor(v) >= x for
// integer x, expand_bits is increasing and expand(x) >= x << (8 - bits), so the largest code X with expand(X) <= floor(v) is
// floor(v) >> (8 - bits) or one less, and j = (X - odd) >> use_par (0 when X < odd).  bits >= 4.  (The bisection was as
// expensive as the window search it feeds: ~50 instructions, two calls per (item, channel, parity combination).)
A7_HD int endpoint_floor_int(real v, int bits, int use_par, int odd) {
	const int iv = (v >= 0) ? ((v >= 255.) ? 255 : (int) v) : -1; // NaN -> -1
	odd = use_par ? odd : 0;
	int X = -1;
	if (iv >= 0) {
		const int xh = iv >> (8 - bits);
		X = expand_bits(bits, xh) <= iv ? xh : xh - 1;
	}
	const int j = X >= odd ? (X - odd) >> use_par : 0;
	return (j << use_par) + odd;
}
// ---- lane = corner form of the cube walk (the CUDA cube kernel of bc7amd.cu) ----------------------------------------
// One work item = one (q, p) re-indexing of one task.  Its set-up runs on one lane per item (32 items at a time); the
// ramp tables of all its lattices are then built by the 32 lanes together, and the (lattices x 64) corners are dealt
// one per lane, so that every lane of the warp shares the texels, the trip count n and the tables.
//   cube_item_setup_u8 : cluster statistics, least-squares endpoints, floors on both parity lattices; result = the
//                        EXPANDED endpoint candidates of every (endpoint e, channel k), ep[e * 3 + k] =
//                        bytes { parity 0: floor, neighbour ; parity 1: floor, neighbour }
//   cube_tab_word      : 4 consecutive entries of one ramp of one lattice (table slot [lattice][k * 4 + x], x = ei0 + 2 ei1)
//   cube_lane_corners  : the corners of lane `lane` (2 .. 8 of them), min key and its (x, y)
//   cube_lane_palette  : palette of the lane's corner (x, y)
template <int CLOG> A7_HD void cube_item_setup_u8(const uint32_t *d, int n, uint64_t collapsed, int q, int p, int bits, int use_par, uint32_t ep_out[6]) {
	ClusterAcc<CLOG> cs;
	cluster_acc<CLOG>(d, n, collapsed, q, p, cs);
	real epa[2][4];
	fit_endpoints_acc<CLOG>(cs, 3, epa);
	const int top = (1 << bits) - 1, reach = 1 << use_par;
#pragma unroll 1
	for (int e = 0; e < 2; e++)
#pragma unroll 1
		for (int k = 0; k < 3; k++) {
			uint32_t w = 0;
#pragma unroll 1
			for (int par = 0; par <= use_par; par++) {
				const int f = endpoint_floor_int(epa[e][k], bits, use_par, par);
				const int up = f + ((top - f < reach ? top - f : reach) & ~use_par);
				w |= ((uint32_t) expand_bits(bits, f) | ((uint32_t) expand_bits(bits, up) << 8)) << (16 * par);
			}
			ep_out[e * 3 + k] = w;
		}
}
template <int CLOG> A7_HD uint32_t ramp_word(int e1, int e2, int c0) { // entries c0 .. c0 + 3 of ramp_bytes
	constexpr int D = (1 << CLOG) - 1;
	const int step = 2 * (e2 - e1);
	int num = 2 * D * e1 + D + c0 * step; // = 2 (D - c) e1 + 2 c e2 + D >= 0
	uint32_t w = 0;
#pragma unroll
	for (int c = 0; c < 4; c++) {
		w |= ((uint32_t) num / (uint32_t) (2 * D)) << (8 * c);
		num += step;
	}
	return w;
}
// word `id` of the item's tables: id = (lattice * 12 + k * 4 + x) * H + half, H = C / 4 words per ramp
template <int CLOG> A7_HD uint32_t cube_tab_word(const uint32_t ep[6], int bcc, int id) {
	constexpr int H = (1 << CLOG) / 4;
	const int hf = id % H, rid = id / H;
	const int x = rid & 3, lk = rid >> 2, l = lk / 3, k = lk - 3 * l;
	const int odd = bcc ? (l >> 1) : l, flip = bcc ? (l & 1) : 0;
	const int e1 = (int) ((ep[k] >> (16 * odd + 8 * (x & 1))) & 255u), e2 = (int) ((ep[3 + k] >> (16 * (odd ^ flip) + 8 * (x >> 1))) & 255u);
	return ramp_word<CLOG>(e1, e2, 4 * hf);
}
// sum over the texels of the smallest squared distance to the palette.  d: 16 words, 16-byte aligned (shared memory on the
// GPU: every lane of the warp reads the same address).  The texel loop is ROLLED, four texels per trip: with the 16 texels
// of the dual-index modes unrolled, a corner was 340 instructions of straight-line code run twice per work item, and the
// cube kernel stalled on instruction fetch more than on anything else (profiles/r2_notes.md).
template <int CLOG> A7_HD uint32_t palette_error_u8(const uint32_t *pal, const uint32_t *d, int n) {
	constexpr int C = 1 << CLOG;
	uint32_t err = 0;
#pragma unroll 1
	for (int t0 = 0; t0 < n; t0 += 4) {
#if defined(__CUDA_ARCH__)
		const uint4 dv = *reinterpret_cast<const uint4 *>(d + t0);
		const uint32_t dd[4] = {dv.x, dv.y, dv.z, dv.w};
#else
		const uint32_t dd[4] = {d[t0], d[t0 + 1], d[t0 + 2], d[t0 + 3]};
#endif
#pragma unroll
		for (int k = 0; k < 4; k++) {
			if (t0 + k >= n) break; // n is warp-uniform on the GPU
			uint32_t m = sq_dist4(pal[0], dd[k]);
#pragma unroll
			for (int c = 1; c < C; c++) m = umin32(m, sq_dist4(pal[c], dd[k]));
			err += m;
		}
	}
	return err;
}
// Corner -> lane map, nlb = log2(lattices): lane bits (LSB first) = lattice (nlb bits), z (2), then
//   nlb 2: y bit 0          ; the lane walks y bit 1 (outer) and x (inner) : 8 corners
//   nlb 1: y                ; the lane walks x                             : 4 corners
//   nlb 0: y, x bit 0       ; the lane walks x bit 1                       : 2 corners
// key = err << 8 | lattice << 6 | gray position (gray_position is linear over GF(2): one XOR per corner)
template <int CLOG>
A7_HD void cube_lane_corners(const uint64_t *tab, const uint32_t *d, int n, int nlb, unsigned lane, uint32_t &best_key, uint32_t &best_xy) {
	constexpr int C = 1 << CLOG;
	const int l = (int) lane & ((1 << nlb) - 1), rest = (int) lane >> nlb;
	const int z = rest & 3;
	const int yb = nlb == 2 ? ((rest >> 2) & 1) : ((rest >> 2) & 3), xb = nlb == 0 ? ((rest >> 4) & 1) : 0;
	const int outer = nlb == 2 ? 2 : 1, inner = nlb == 0 ? 2 : 4, xstep = nlb == 0 ? 2 : 1;
	const uint64_t *tl = tab + l * 12;
	const uint64_t tz = tl[8 + z];
	const uint32_t gz = (uint32_t) gray_position(z << 4) | ((uint32_t) l << 6);
	best_key = 0xffffffffu;
	best_xy = 0;
#pragma unroll 1
	for (int o = 0; o < outer; o++) {
		const int y = yb | (o << 1);
		const uint64_t ty = tl[4 + y];
		uint32_t pzy[C];
#pragma unroll
		for (int c = 0; c < C; c++) pzy[c] = put_ramp_byte<2>(put_ramp_byte<1>(0u, ty, c), tz, c);
		const uint32_t gzy = gz ^ (uint32_t) gray_position(y << 2);
#pragma unroll 1
		for (int i = 0; i < inner; i++) {
			const int x = xb + i * xstep;
			const uint64_t tx = tl[x];
			uint32_t pal[C];
#pragma unroll
			for (int c = 0; c < C; c++) pal[c] = put_ramp_byte<0>(pzy[c], tx, c);
			const uint32_t err = palette_error_u8<CLOG>(pal, d, n);
			const uint32_t key = (err << 8) | (gzy ^ (uint32_t) gray_position(x));
			if (key < best_key) {
				best_key = key;
				best_xy = (uint32_t) (x | (y << 2));
			}
		}
	}
}
template <int CLOG> A7_HD void cube_lane_palette(const uint64_t *tab, int nlb, unsigned lane, uint32_t xy, uint32_t *pal) {
	constexpr int C = 1 << CLOG;
	const int l = (int) lane & ((1 << nlb) - 1), z = ((int) lane >> nlb) & 3;
	const uint64_t *tl = tab + l * 12;
	const uint64_t tx = tl[xy & 3u], ty = tl[4 + ((xy >> 2) & 3u)], tz = tl[8 + z];
#pragma unroll
	for (int c = 0; c < C; c++) pal[c] = put_ramp_byte<2>(put_ramp_byte<1>(put_ramp_byte<0>(0u, tx, c), ty, c), tz, c);
}

// ---- byte-parallel helpers (native on the GPU) ---------------------------------------------------------------------
A7_HD uint32_t vabsdiff4_u8(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
	return __vabsdiffu4(a, b);
#else
	uint32_t r = 0;
	for (int k = 0; k < 4; k++) {
		const int d = (int) ((a >> (8 * k)) & 255u) - (int) ((b >> (8 * k)) & 255u);
		r |= (uint32_t) (d < 0 ? -d : d) << (8 * k);
	}
	return r;
#endif
}
A7_HD uint32_t vmin4_u8(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
	return __vminu4(a, b);
#else
	uint32_t r = 0;
	for (int k = 0; k < 4; k++) {
		const uint32_t x = (a >> (8 * k)) & 255u, y = (b >> (8 * k)) & 255u;
		r |= (x < y ? x : y) << (8 * k);
	}
	return r;
#endif
}
A7_HD uint32_t perm_bytes(uint32_t lo, uint32_t hi, uint32_t sel) { // byte b of the result = byte (sel >> 4b & 7) of hi:lo
#if defined(__CUDA_ARCH__)
	return __byte_perm(lo, hi, sel);
#else
	const uint64_t v = (uint64_t) lo | ((uint64_t) hi << 32);
	uint32_t r = 0;
	for (int b = 0; b < 4; b++) r |= (uint32_t) ((v >> (8 * ((sel >> (4 * b)) & 7u))) & 255u) << (8 * b);
	return r;
#endif
}
A7_HD uint32_t dot4_u8(uint32_t a, uint32_t b, uint32_t acc) {
#if defined(__CUDA_ARCH__)
	return __dp4a(a, b, acc);
#else
	for (int k = 0; k < 4; k++) acc += ((a >> (8 * k)) & 255u) * ((b >> (8 * k)) & 255u);
	return acc;
#endif
}
// ---- exact pruning of the SECOND pass of ep_shaker_d ----------------------------------------------------------------
// The second pass of a task only matters where it is STRICTLY better than the first (:1372-1400), so a corner whose lower
// bound reaches the first pass's error can be skipped.  Bound (cube_search_pruned_u8): the error of a corner is at least the
// sum over the three channels of sum_i min_c (ramp[c] - d_i,k)^2 -- twelve per-channel sums per lattice (cube_bound_u8)
// bound all its 64 corners.  A corner is named cid = lattice << 6 | z << 4 | y << 2 | x.
// plane[k * 4 + w] = channel k of texels 4w .. 4w+3 (pads 0): see window_planes_u8; pl = the four words of one channel
// (sm_100a has no byte-wise minimum -- __vminu4 is a dozen instructions -- but a native 16x2 one: the four absolute
// differences of a word are spread over two half-word pairs by byte permutes and reduced with VIMNMX3.U16x2)
A7_HD uint32_t vmin3_u16x2(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
	return __vimin3_u16x2(a, b, c);
#else
	uint32_t r = 0;
	for (int k = 0; k < 2; k++) {
		uint32_t x = (a >> (16 * k)) & 0xffffu;
		const uint32_t y = (b >> (16 * k)) & 0xffffu, z = (c >> (16 * k)) & 0xffffu;
		x = x < y ? x : y;
		x = x < z ? x : z;
		r |= x << (16 * k);
	}
	return r;
#endif
}
template <int CLOG> A7_HD uint32_t cube_bound_u8(uint64_t ramp, const uint32_t *pl, int n) {
	constexpr int C = 1 << CLOG;
	uint32_t rc[C];
#pragma unroll
	for (int c = 0; c < C; c++) rc[c] = byte_of(ramp, c) * 0x01010101u;
	const int nw = (n + 3) >> 2;
	const uint32_t last = (n & 3) ? ((1u << (8 * (n & 3))) - 1u) : 0xffffffffu;
	uint32_t acc = 0;
#pragma unroll
	for (int w = 0; w < 4; w++) {
		if (w >= nw) break;
		const uint32_t dw = pl[w];
		uint32_t mlo = 0xffffffffu, mhi = 0xffffffffu; // running minima of texels (0, 1) / (2, 3) of the word, one per half-word
#pragma unroll
		for (int c = 0; c < C; c += 2) {
			const uint32_t a0 = vabsdiff4_u8(rc[c], dw), a1 = vabsdiff4_u8(rc[c + 1], dw);
			mlo = vmin3_u16x2(mlo, perm_bytes(a0, 0u, 0x4140u), perm_bytes(a1, 0u, 0x4140u));
			mhi = vmin3_u16x2(mhi, perm_bytes(a0, 0u, 0x4342u), perm_bytes(a1, 0u, 0x4342u));
		}
		uint32_t m = perm_bytes(mlo, mhi, 0x6420u);
		if (w == nw - 1) m &= last;
		acc = dot4_u8(m, m, acc);
	}
	return acc;
}
A7_HD uint32_t cube_cid_bound(const uint32_t *lbs, int cid) {
	const uint32_t *l = lbs + 12 * (cid >> 6);
	return l[cid & 3] + l[4 + ((cid >> 2) & 3)] + l[8 + ((cid >> 4) & 3)];
}
template <int CLOG> A7_HD void cube_cid_palette(const uint64_t *tab, int cid, uint32_t *pal) {
	constexpr int C = 1 << CLOG;
	const uint64_t *tl = tab + 12 * (cid >> 6);
	const uint64_t tx = tl[cid & 3], ty = tl[4 + ((cid >> 2) & 3)], tz = tl[8 + ((cid >> 4) & 3)];
#pragma unroll
	for (int c = 0; c < C; c++) pal[c] = put_ramp_byte<2>(put_ramp_byte<1>(put_ramp_byte<0>(0u, tx, c), ty, c), tz, c);
}
A7_HD uint32_t cube_cid_key(uint32_t err, int cid) { return (err << 8) | (uint32_t) (cid & 0xc0) | (uint32_t) gray_position(cid & 63); }
template <int CLOG> A7_HD uint32_t cube_corner_error_u8(const uint32_t *pal, const uint32_t *d, int n) { return palette_error_u8<CLOG>(pal, d, n); }

// ---- ramps from a difference table ---------------------------------------------------------------------------------
// ramp entry c between expanded endpoints e1, e2 = floor((2 D e1 + D + 2 c (e2 - e1)) / (2 D)) = e1 + off(e2 - e1, c), and
// every entry stays within 0 .. 255, so a whole ramp is a byte-parallel add (e2 >= e1) or subtract of a table word: no
// carries between the bytes.  Table of C = 2^CLOG: words [a * H + h] for |e2 - e1| = a, H = C / 4 words per ramp; the
// non-negative differences first, the negative ones kRampLutNeg<CLOG> words later.
template <int CLOG> struct RampLutShape {
	static constexpr int H = (1 << CLOG) / 4, kNeg = 256 * H, kWords = 512 * H;
};
template <int CLOG> A7_HD void ramp_lut_fill(uint32_t *lut, int first, int step) { // entries first, first + step, ... of the table
	constexpr int C = 1 << CLOG, D = C - 1, H = C / 4;
#pragma unroll 1
	for (int id = first; id < 512 * H; id += step) {
		const int neg = id >= 256 * H, a = (id - (neg ? 256 * H : 0)) / H, h = id % H;
		uint32_t w = 0;
#pragma unroll 1
		for (int b = 0; b < 4; b++) {
			const int c = 4 * h + b;
			const int num = neg ? 2 * c * a - D : D + 2 * c * a;
			const int off = neg ? (num <= 0 ? 0 : (num + 2 * D - 1) / (2 * D)) : num / (2 * D);
			w |= (uint32_t) off << (8 * b);
		}
		lut[id] = w;
	}
}
template <int CLOG> A7_HD uint32_t ramp_lut_word(const uint32_t *lut, int e1, int e2, int h) {
	constexpr int H = (1 << CLOG) / 4;
	const int dl = e2 - e1;
	const uint32_t rep = (uint32_t) e1 * 0x01010101u;
	const uint32_t v = lut[(dl < 0 ? (256 - dl) * H : dl * H) + h]; // (one load, no branch: the lanes of a warp mix both signs)
	return dl < 0 ? rep - v : rep + v;
}
// the whole ramp (8 entries: both words with one 64-bit load; 4 entries: the high word is 0)
template <int CLOG> A7_HD void ramp_lut_pair(const uint32_t *lut, int e1, int e2, uint32_t &r0, uint32_t &r1) {
	constexpr int H = (1 << CLOG) / 4;
	const int dl = e2 - e1;
	const uint32_t rep = (uint32_t) e1 * 0x01010101u;
	const int at = (dl < 0 ? (256 - dl) * H : dl * H);
	if (H == 2) {
		const uint64_t v = *reinterpret_cast<const uint64_t *>(lut + at); // (entries are pairs of words: 8-byte aligned)
		r0 = dl < 0 ? rep - (uint32_t) v : rep + (uint32_t) v;
		r1 = dl < 0 ? rep - (uint32_t) (v >> 32) : rep + (uint32_t) (v >> 32);
	} else {
		const uint32_t v = lut[at];
		r0 = dl < 0 ? rep - v : rep + v;
		r1 = 0;
	}
}
// table-lookup form of cube_tab_word
template <int CLOG> A7_HD uint32_t cube_tab_word_lut(const uint32_t *lut, const uint32_t ep[6], int bcc, int id) {
	constexpr int H = (1 << CLOG) / 4;
	const int hf = id % H, rid = id / H;
	const int x = rid & 3, lk = rid >> 2, l = lk / 3, k = lk - 3 * l;
	const int odd = bcc ? (l >> 1) : l, flip = bcc ? (l & 1) : 0;
	const int e1 = (int) ((ep[k] >> (16 * odd + 8 * (x & 1))) & 255u), e2 = (int) ((ep[3 + k] >> (16 * (odd ^ flip) + 8 * (x >> 1))) & 255u);
	return ramp_lut_word<CLOG>(lut, e1, e2, hf);
}

// ep_shaker_d on packed 8-bit data (dimension 3). index_io in/out; returns the SSE (exact integer as real).
template <int CLOG>
A7_HDN real shake_cube_u8(const Tables &T, const U8Subset &S, int *index_io, const int *bits, int type) {
	constexpr int Mi_ = (1 << CLOG) - 1;
	const int n = S.n;
	const int use_par = (type == BCC || type == SAME_PAR), bcc = (type == BCC);
	int index[kMaxEntries];
#pragma unroll 1
	for (int k = 0; k < n; k++) index[k] = index_io[k];
	real err_o = A7_HUGE;
	int maxTry = 1, done;
	do {
		const int Mi = collapse_indices(index, n);
		if (Mi == 0) {
			int e0[2][4];
			const real t = shake_single_index_u8(T, S, CLOG, bits, type, 3, index, e0);
			if (t < err_o) {
#pragma unroll 1
				for (int k = 0; k < n; k++) index_io[k] = index[k];
				err_o = t;
			}
			return err_o;
		}
		int p0 = -1, q0 = -1;
		uint32_t err_2 = 0xffffffffu;
		uint64_t idx_2 = 0;
#pragma unroll 1
		for (int q = 1; q * Mi <= Mi_; q++)
#pragma unroll 1
			for (int p = 0; p <= Mi_ - q * Mi; p++) {
				uint64_t collapsed = 0;
#pragma unroll 1
				for (int k = 0; k < n; k++) collapsed |= (uint64_t) (index[k] & 15) << (4 * k);
				ClusterAcc<CLOG> cs;
				cluster_acc<CLOG>(S.d, n, collapsed, q, p, cs);
				real epa[2][4];
				fit_endpoints_acc<CLOG>(cs, 3, epa);
				uint32_t key = 0xffffffffu;
				uint64_t idx_1 = 0;
				cube_search_pruned_u8<CLOG>(S.d, n, bits, epa, use_par, bcc, 0, 4, key, idx_1);
				const uint32_t err_1 = key >> 8;
				if (err_1 < err_2) {
					err_2 = err_1;
					idx_2 = idx_1;
					p0 = p;
					q0 = q;
				}
			}
		int change = 0;
#pragma unroll 1
		for (int k = 0; k < n; k++) change = change || (index[k] * q0 + p0 != (int) ((idx_2 >> (4 * k)) & 15u));
		const int better = (real) err_2 < err_o;
		if (better) {
#pragma unroll 1
			for (int k = 0; k < n; k++) index_io[k] = index[k] = (int) ((idx_2 >> (4 * k)) & 15u);
			err_o = (real) err_2;
		}
		done = !(change && better);
	} while (!done && maxTry--);
	return err_o;
}

// ---- ep_shaker_2_d (:703-1053) on packed 8-bit data, cut into its two kinds of independent work -------------
A7_HD uint64_t pack_ep8(const int ep[2][4]) { // 8 endpoint codes, one byte each: ep[0][0..3] | ep[1][0..3] << 32
	uint64_t v = 0;
#pragma unroll
	for (int j = 0; j < 4; j++) v |= ((uint64_t) (ep[0][j] & 255) << (8 * j)) | ((uint64_t) (ep[1][j] & 255) << (32 + 8 * j));
	return v;
}
A7_HD void unpack_ep8(uint64_t v, int ep[2][4]) {
#pragma unroll
	for (int j = 0; j < 4; j++) {
		ep[0][j] = (int) ((v >> (8 * j)) & 255u);
		ep[1][j] = (int) ((v >> (32 + 8 * j)) & 255u);
	}
}

// One (q, p) re-indexing (:836-1000): least-squares endpoints, per-channel window search with fixed indices, best
// parity vector. Returns err_1 (exact integer) and epo_1 (packed).
template <int CLOG>
A7_HDN uint32_t window_item_u8(const uint32_t *d, int n, uint64_t collapsed, int q, int p, int size, int bits_total, int dim, uint64_t &epo_out) {
	constexpr int C = 1 << CLOG, Mi_ = C - 1;
	constexpr int W = C > 8 ? 2 : 1;
	const int type = bits_total % (2 * dim);
	const int use_par = type != 0;
	const int mb = (bits_total + 2 * dim - 1) / (2 * dim);
	int sq_total[4] = {0, 0, 0, 0}; // sum of squares of the data per channel
#pragma unroll 1
	for (int k = 0; k < n; k++) {
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const int b = (int) ((d[k] >> (8 * j)) & 255u);
			sq_total[j] += b * b;
		}
	}
	ClusterAcc<CLOG> cs;
	cluster_acc<CLOG>(d, n, collapsed, q, p, cs);
	real epa[2][4];
	fit_endpoints_acc<CLOG>(cs, dim, epa);
	int ed[2][2][4], ep2[2][2][2][4];
	const int rr = use_par ? 2 : 1, step = 1 << use_par, top = (1 << mb) - 1;
#pragma unroll 1
	for (int j = 0; j < dim; j++) {
		int cnt[C], sum2[C]; // this channel's cluster counts and doubled sums (static indices: registers)
#pragma unroll
		for (int c = 0; c < C; c++) {
			cnt[c] = acc_count(cs.a[c]);
			sum2[c] = 2 * acc_sum(cs.a[c], j);
		}
		const int sqj = j == 0 ? sq_total[0] : (j == 1 ? sq_total[1] : (j == 2 ? sq_total[2] : sq_total[3]));
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
				int best = INT32_MAX, b1 = 0, b2 = 0;
#pragma unroll 1
				for (int p1 = lo[0]; p1 <= hi[0]; p1 += step) {
					const int e1 = expand_bits(mb, p1);
#pragma unroll 1
					for (int p2 = lo[1]; p2 <= hi[1]; p2 += step) {
						// ramp entry c = (2 D e1 + D + 2 c (e2 - e1)) / (2 D) (ramp_bytes), used straight from the quotient;
						// sum_m (r[cidx[m]] - d[m])^2 == sum d^2 + sum_c r_c * (cnt_c * r_c - 2 * S1_c)
						int num = 2 * Mi_ * e1 + Mi_;
						const int dnum = 2 * (expand_bits(mb, p2) - e1);
						int t = sqj;
#pragma unroll
						for (int c = 0; c < C; c++) {
							const int r = (int) ((uint32_t) num / (uint32_t) (2 * Mi_));
							num += dnum;
							t += r * (cnt[c] * r - sum2[c]);
						}
						if (t < best) { best = t; b1 = p1; b2 = p2; }
					}
				}
				ed[pp0][pp1][j] = best;
				ep2[pp0][pp1][0][j] = b1;
				ep2[pp0][pp1][1][j] = b2;
			}
	}
	int64_t err_1 = INT64_MAX;
	int epo_1[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
#pragma unroll 1
	for (int pn = 0; pn < (1 << type); pn++) {
		const int v0 = type == SAME_PAR ? pn : (pn >> 1), v1 = type == SAME_PAR ? pn : (pn & 1);
		int64_t e2 = 0;
#pragma unroll 1
		for (int j = 0; j < dim; j++) e2 += ed[v0][v1][j];
		if (e2 < err_1) {
			err_1 = e2;
#pragma unroll 1
			for (int j = 0; j < dim; j++) { epo_1[0][j] = ep2[v0][v1][0][j]; epo_1[1][j] = ep2[v0][v1][1][j]; }
		}
	}
	epo_out = pack_ep8(epo_1);
	return (uint32_t) err_1;
}

// The same work item with the candidate ramps taken from the difference table and the error summed per TEXEL instead of
// per cluster: the ramp of a candidate (p1, p2) is a 64-bit word, texel m's entry r[cidx[m]] is picked out of it by a
// byte permute (4 texels per PRMT, selectors built once per item), and |r - d| / its square are 4-wide byte
// instructions against the channel plane of the data -- 3 instructions per 4 texels and candidate.
// sum_m (r[cidx[m]] - d[m])^2 is the same integer the cluster form computes, so the search is identical.
// plane[j * 4 + w] = channel j of texels 4w .. 4w+3 (pads 0), all four channels
A7_HD void window_planes_u8(const uint32_t *d, int n, uint32_t plane[16]) {
#pragma unroll 1
	for (int j = 0; j < 4; j++)
#pragma unroll 1
		for (int w = 0; w < 4; w++) {
			uint32_t v = 0;
#pragma unroll 1
			for (int b = 0; b < 4; b++)
				if (4 * w + b < n) v |= ((d[4 * w + b] >> (8 * j)) & 255u) << (8 * b);
			plane[j * 4 + w] = v;
		}
}
// The work item in three steps, so that the CUDA window kernel can run the middle one with a lane per (item, channel,
// parity combination) instead of a lane per item:
//   window_item_fit_u8     least-squares endpoints, byte-permute selectors
//   window_sub_search_u8   one channel, one (pp0, pp1): lattice floors of the two endpoints, the (p1, p2) window around them,
//                          first minimum in scan order
//   window_item_combine    sum over the channels per parity vector, first minimum (:975-1000)
// epa[i * 4 + j] = least-squares endpoint i of channel j
template <int CLOG>
A7_HD void window_item_fit_u8(const uint32_t *d, int n, uint64_t collapsed, int q, int p, int dim, real epa_out[8], uint32_t sel[4]) {
	ClusterAcc<CLOG> cs;
	cluster_acc<CLOG>(d, n, collapsed, q, p, cs);
	real epa[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
	fit_endpoints_acc<CLOG>(cs, dim, epa);
#pragma unroll
	for (int i = 0; i < 2; i++)
#pragma unroll
		for (int j = 0; j < 4; j++) epa_out[i * 4 + j] = epa[i][j];
#pragma unroll
	for (int w = 0; w < 4; w++) {
		uint32_t sw = 0;
#pragma unroll
		for (int b = 0; b < 4; b++) {
			const int k = 4 * w + b;
			const uint32_t ck = k < n ? (uint32_t) ((collapsed >> (4 * k)) & 15u) * (uint32_t) q + (uint32_t) p : 0u;
			sw |= ck << (4 * b);
		}
		sel[w] = sw;
	}
}
// returns best error << 16 | p1 << 8 | p2
template <int CLOG>
A7_HD uint64_t window_sub_search_u8(const uint32_t *lut, real ep0, real ep1, const uint32_t sel[4], const uint32_t plw[4], int n, int mb, int use_par,
																	int size, int pp0, int pp1) {
	constexpr int C = 1 << CLOG;
	static_assert(C <= 8, "one 64-bit ramp");
	const int nw = (n + 3) >> 2;
	const uint32_t last = (n & 3) ? ((1u << (8 * (n & 3))) - 1u) : 0xffffffffu;
	const int step = 1 << use_par, top = (1 << mb) - 1;
	int lo[2], hi[2];
#pragma unroll
	for (int i = 0; i < 2; i++) {
		const int f = endpoint_floor_int(i ? ep1 : ep0, mb, use_par, i ? pp1 : pp0);
		lo[i] = f - ((f < (size >> 1) - 1 ? f : (size >> 1) - 1) & ~use_par);
		hi[i] = f + ((top - f < (size >> 1) ? top - f : (size >> 1)) & ~use_par);
	}
	uint32_t best = 0x7fffffffu;
	int b1 = 0, b2 = 0;
#pragma unroll 1
	for (int p1 = lo[0]; p1 <= hi[0]; p1 += step) {
		const int e1 = expand_bits(mb, p1);
		// three candidates of the row at a time: their table loads are in flight together; compared in scan order
#pragma unroll 1
		for (int p2 = lo[1]; p2 <= hi[1]; p2 += 3 * step) {
			uint32_t t[3];
#pragma unroll
			for (int u = 0; u < 3; u++) {
				const int q2 = p2 + u * step;
				uint32_t r0, r1;
				ramp_lut_pair<CLOG>(lut, e1, expand_bits(mb, q2 <= hi[1] ? q2 : p2), r0, r1);
				uint32_t tu = 0;
#pragma unroll
				for (int w = 0; w < 4; w++) { // (static indices: the selectors and the plane words stay in registers)
					if (w >= nw) break;
					uint32_t ad = vabsdiff4_u8(perm_bytes(r0, r1, sel[w]), plw[w]);
					if (w == nw - 1) ad &= last;
					tu = dot4_u8(ad, ad, tu);
				}
				t[u] = q2 <= hi[1] ? tu : 0xffffffffu;
			}
#pragma unroll
			for (int u = 0; u < 3; u++)
				if (t[u] < best) { best = t[u]; b1 = p1; b2 = p2 + u * step; }
		}
	}
	return ((uint64_t) best << 16) | ((uint64_t) (b1 & 255) << 8) | (uint64_t) (b2 & 255);
}
// res[j * 4 + pp0 * 2 + pp1] from window_sub_search_u8 (SAME_PAR never reads the mixed-parity entries)
A7_HD uint32_t window_item_combine(const uint64_t *res, int type, int dim, uint64_t &epo_out, int stride = 4) { // res[j * stride + pp0 * 2 + pp1]
	int64_t err_1 = INT64_MAX;
	int epo_1[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
#pragma unroll 1
	for (int pn = 0; pn < (1 << type); pn++) {
		const int v0 = type == SAME_PAR ? pn : (pn >> 1), v1 = type == SAME_PAR ? pn : (pn & 1);
		int64_t e2 = 0;
#pragma unroll 1
		for (int j = 0; j < dim; j++) e2 += (int64_t) (res[j * stride + v0 * 2 + v1] >> 16);
		if (e2 < err_1) {
			err_1 = e2;
#pragma unroll 1
			for (int j = 0; j < dim; j++) {
				const uint64_t r = res[j * stride + v0 * 2 + v1];
				epo_1[0][j] = (int) ((r >> 8) & 255u);
				epo_1[1][j] = (int) (r & 255u);
			}
		}
	}
	epo_out = pack_ep8(epo_1);
	return (uint32_t) err_1;
}
// the three steps run by one caller (host build / test)
template <int CLOG>
A7_HDN uint32_t window_item_lut_u8(const uint32_t *lut, const uint32_t *d, const uint32_t *plane, int n, uint64_t collapsed, int q, int p, int size,
																	 int bits_total, int dim, uint64_t &epo_out) {
	const int type = bits_total % (2 * dim);
	const int use_par = type != 0;
	const int mb = (bits_total + 2 * dim - 1) / (2 * dim);
	real epa[8];
	uint32_t sel[4];
	window_item_fit_u8<CLOG>(d, n, collapsed, q, p, dim, epa, sel);
	uint64_t res[16];
#pragma unroll 1
	for (int j = 0; j < dim; j++)
#pragma unroll 1
		for (int pp0 = 0; pp0 <= use_par; pp0++)
#pragma unroll 1
			for (int pp1 = 0; pp1 <= use_par; pp1++) {
				if (type == SAME_PAR && pp0 != pp1) continue; // (the reference searches these too, nobody reads them)
				const uint32_t plw[4] = {plane[4 * j], plane[4 * j + 1], plane[4 * j + 2], plane[4 * j + 3]};
				res[j * 4 + pp0 * 2 + pp1] = window_sub_search_u8<CLOG>(lut, epa[j], epa[4 + j], sel, plw, n, mb, use_par, size, pp0, pp1);
			}
	return window_item_combine(res, type, dim, epo_out);
}

// Re-clustering against chosen endpoints (:1003-1030): packed palette, 4 native instructions per (texel, entry)
template <int CLOG>
A7_HDN uint32_t recluster_u8(const uint32_t *d, int n, uint64_t epo, int mb, int dim, uint64_t &idg_out) {
	constexpr int C = 1 << CLOG;
	constexpr int W = C > 8 ? 2 : 1;
	int ep[2][4];
	unpack_ep8(epo, ep);
	uint32_t pal[C];
	{
		uint64_t rb[4][W];
#pragma unroll
		for (int j = 0; j < 4; j++) {
			rb[j][0] = 0;
			if (W > 1) rb[j][W - 1] = 0;
			if (j < dim) ramp_bytes<CLOG>(expand_bits(mb, ep[0][j]), expand_bits(mb, ep[1][j]), rb[j]);
		}
#pragma unroll
		for (int c = 0; c < C; c++)
			pal[c] = byte_of(rb[0][c >> 3], c & 7) | (byte_of(rb[1][c >> 3], c & 7) << 8) | (byte_of(rb[2][c >> 3], c & 7) << 16) |
							 (byte_of(rb[3][c >> 3], c & 7) << 24);
	}
	uint32_t err_r = 0;
	uint64_t idg = 0;
#pragma unroll 1
	for (int i = 0; i < n; i++) {
		uint32_t m = 0xffffffffu;
#pragma unroll
		for (int c = 0; c < C; c++) m = umin32(m, (sq_dist4(pal[c], d[i]) << 4) | (uint32_t) c);
		err_r += m >> 4;
		idg |= (uint64_t) (m & 15u) << (4 * i);
	}
	idg_out = idg;
	return err_r;
}

// ep_shaker_2_d on packed 8-bit data (dimension 3 or 4). index_io in/out, epo_code out; returns the SSE.
template <int CLOG>
A7_HDN real shake_window_u8(const Tables &T, const U8Subset &S, int *index_io, int epo_code[2][4], int size, int bits_total, int dim) {
	constexpr int Mi_ = (1 << CLOG) - 1;
	const int n = S.n;
	const int type = bits_total % (2 * dim);
	const int mb = (bits_total + 2 * dim - 1) / (2 * dim);
	const int max_bits[4] = {mb, mb, mb, mb};
	int index[kMaxEntries];
#pragma unroll 1
	for (int k = 0; k < n; k++) index[k] = index_io[k];
	real err_o = A7_HUGE;
	int maxTry = 8, done;
	do {
		const int Mi = collapse_indices(index, n);
		if (Mi == 0) {
			int e0[2][4];
			const real t = shake_single_index_u8(T, S, CLOG, max_bits, type, dim, index, e0);
			if (t < err_o) {
#pragma unroll 1
				for (int k = 0; k < n; k++) index_io[k] = index[k];
#pragma unroll 1
				for (int j = 0; j < dim; j++) { epo_code[0][j] = e0[0][j]; epo_code[1][j] = e0[1][j]; }
				err_o = t;
			}
			return err_o;
		}
		uint64_t collapsed = 0;
#pragma unroll 1
		for (int k = 0; k < n; k++) collapsed |= (uint64_t) (index[k] & 15) << (4 * k);
		int p0 = -1, q0 = -1;
		uint32_t err_0 = 0xffffffffu;
		uint64_t epo_0 = 0;
#pragma unroll 1
		for (int q = 1; q * Mi <= Mi_; q++)
#pragma unroll 1
			for (int p = 0; p <= Mi_ - q * Mi; p++) {
				uint64_t epo_1;
				const uint32_t err_1 = window_item_u8<CLOG>(S.d, n, collapsed, q, p, size, bits_total, dim, epo_1);
				if (err_1 <= err_0) { // `<=`: the LAST minimum wins (:994)
					err_0 = err_1;
					p0 = p;
					q0 = q;
					epo_0 = epo_1;
				}
			}
		uint64_t idg;
		const uint32_t err_r = recluster_u8<CLOG>(S.d, n, epo_0, mb, dim, idg);
		int change = 0;
#pragma unroll 1
		for (int k = 0; k < n; k++) change = change || (index[k] * q0 + p0 != (int) ((idg >> (4 * k)) & 15u));
		const int better = (real) err_r < err_o;
		if (better) {
#pragma unroll 1
			for (int k = 0; k < n; k++) index_io[k] = index[k] = (int) ((idg >> (4 * k)) & 15u);
			unpack_ep8(epo_0, epo_code);
			err_o = (real) err_r;
		}
		done = !(change && better);
	} while (!done && maxTry--);
	return err_o;
}

// runtime-CLOG front ends
A7_HD real shake_cube_u8_any(const Tables &T, const U8Subset &S, int *idx, int clog, const int *bits, int type) {
	return clog == 2 ? shake_cube_u8<2>(T, S, idx, bits, type) : shake_cube_u8<3>(T, S, idx, bits, type);
}
A7_HD real shake_window_u8_any(const Tables &T, const U8Subset &S, int *idx, int ep[2][4], int size, int clog, int bits_total, int dim) {
	if (clog == 2) return shake_window_u8<2>(T, S, idx, ep, size, bits_total, dim);
	if (clog == 3) return shake_window_u8<3>(T, S, idx, ep, size, bits_total, dim);
	return shake_window_u8<4>(T, S, idx, ep, size, bits_total, dim);
}

// shake_subset (bc7amd_core.cuh) on packed data
A7_HD real shake_subset_u8(const Tables &T, const ShakeParams &sp, const U8Subset &S, int *idx, int ep[2][4]) {
	const int clog = ilog2(sp.clusters);
	if (sp.dim != 3) return shake_window_u8_any(T, S, idx, ep, sp.shake_size, clog, sp.bits[3], sp.dim);
	int tmp[kMaxEntries];
#pragma unroll 1
	for (int k = 0; k < S.n; k++) tmp[k] = idx[k];
	const real e0 = shake_cube_u8_any(T, S, tmp, clog, sp.bits, sp.parity);
	real e1 = shake_window_u8_any(T, S, idx, ep, sp.shake_size, clog, sp.bits[3], sp.dim);
	if (e0 < e1) {
		e1 = shake_window_u8_any(T, S, tmp, ep, sp.shake_size, clog, sp.bits[3], sp.dim);
#pragma unroll 1
		for (int k = 0; k < S.n; k++) idx[k] = tmp[k];
	}
	return e1;
}

} // namespace amd7
} // namespace b200ic
