// bc7rg_core.cuh -- bc7enc16-compatible BC7 encoder (modes 1 and 6), one 4x4 block per thread.
//
// Bit-exact target: reference src/richgel999_bc7enc16.cpp:99-1547 (bc7enc16_compress_block and callees) as
// driven by Image_CompressRichGel999BC7enc16 (:73-97): perceptual -> YCbCr weights {128,64,16,32} else {1,1,1,1};
// fast -> uber level 0 else 4; 64 mode-1 partitions; least squares on; partition filterbank on.
//
// Written from the algorithm, not the reference's data structures:
//   * a colour cell is a compacted list of <=16 RGBA8 texels kept as packed u32 + its selectors as 4-bit fields
//     of one u64 (the reference keeps byte arrays and copies them around);
//   * the running best of a cell is a value type `Best`; "evaluate" returns a candidate that the caller merges
//     with the reference's first-strict-minimum rule;
//   * the partition estimator evaluates the error of every shape completely and replays the reference's
//     scan (early-outs, filterbank, checkerboard break) on the finished numbers -- the reference's partial sums
//     are monotone, so a partial sum that broke out can never win (`<`) and the replay is output-identical.
// All floating point is FP32 with no contraction (--fmad=false / -ffp-contract=off) and in reference order.
//
// The same source builds for the host (tests/hostbuild, a debugging aid that is NOT part of the product library)
// and for sm_100a.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define B7_HD __host__ __device__ __forceinline__
#define B7_HDN __host__ __device__ __noinline__
#else
#define B7_HD inline
#define B7_HDN
#endif

namespace b200ic {
namespace rg {

struct OptimalEndpoint { uint16_t err; uint8_t lo, hi; }; // single-colour table entry  (:160)

struct Params {
	uint32_t w[4];       // channel / YCbCrA weights after the x4 scaling of :1524-1535
	int perceptual;
	int uber;            // 0 or 4
	const OptimalEndpoint *opt1; // [256][2] mode-1 single-colour table (:166-195)
};

// Host-side construction of the single-colour table (same search as :166-195, own loop structure).
inline void build_mode1_single_colour_table(OptimalEndpoint table[512]) {
	for (int c = 0; c < 256; c++)
		for (int p = 0; p < 2; p++) {
			OptimalEndpoint best = {0xffff, 0, 0};
			for (int l = 0; l < 64; l++) {
				int lo = ((l << 1) | p) << 1;
				lo |= lo >> 7;
				for (int h = 0; h < 64; h++) {
					int hi = ((h << 1) | p) << 1;
					hi |= hi >> 7;
					const int k = (lo * (64 - 18) + hi * 18 + 32) >> 6; // weight3[2] = 18
					const int e = (k - c) * (k - c);
					if (e < best.err) { best.err = (uint16_t) e; best.lo = (uint8_t) l; best.hi = (uint8_t) h; }
				}
			}
			table[c * 2 + p] = best;
		}
}

inline void make_params(Params &P, bool perceptual, bool fast, const OptimalEndpoint *table) {
	if (perceptual) { // :1524-1533 with m_weights = {128,64,16,32}
		const float pr = (.5f / (1.0f - .2126f)) * (.5f / (1.0f - .2126f));
		const float pb = (.5f / (1.0f - .0722f)) * (.5f / (1.0f - .0722f));
		P.w[0] = (uint32_t) (int) (128 * 4.0f);
		P.w[1] = (uint32_t) (int) (64 * 4.0f * pr);
		P.w[2] = (uint32_t) (int) (16 * 4.0f * pb);
		P.w[3] = 32 * 4;
	} else {
		P.w[0] = P.w[1] = P.w[2] = P.w[3] = 1;
	}
	P.perceptual = perceptual ? 1 : 0;
	P.uber = fast ? 0 : 4;
	P.opt1 = table;
}

// ---- small helpers --------------------------------------------------------------------------------------
struct F4 { float v[4]; };
B7_HD float sat(float x) { return x < 0.f ? 0.f : (x > 1.0f ? 1.0f : x); }
B7_HD int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
B7_HD uint32_t ch(uint32_t px, int c) { return (px >> (8 * c)) & 255u; }
B7_HD uint32_t pack4(uint32_t r, uint32_t g, uint32_t b, uint32_t a) { return r | (g << 8) | (b << 16) | (a << 24); }
B7_HD float dot4(const F4 &a, const F4 &b) { return a.v[0] * b.v[0] + a.v[1] * b.v[1] + a.v[2] * b.v[2] + a.v[3] * b.v[3]; }
B7_HD void normalise(F4 &a) {
	float s = a.v[0] * a.v[0] + a.v[1] * a.v[1] + a.v[2] * a.v[2] + a.v[3] * a.v[3];
	if (s != 0.0f) {
		s = 1.0f / sqrtf(s);
		a.v[0] *= s; a.v[1] *= s; a.v[2] *= s; a.v[3] *= s;
	}
}
B7_HD uint32_t sel_get(uint64_t s, int i) { return (uint32_t) (s >> (4 * i)) & 15u; }
B7_HD uint64_t sel_put(uint32_t v, int i) { return (uint64_t) v << (4 * i); }

B7_HD uint32_t weight3(int i) { return (0x40372e251b120900ull >> (8 * i)) & 255u; }          // 0,9,18,27,37,46,55,64
B7_HD uint32_t weight4(int i) {                                                               // 0,4,9,13,...,60,64
	const uint64_t lo = 0x1e1a15110d090400ull, hi = 0x403c37332f2b2622ull;
	return (uint32_t) ((i < 8 ? lo >> (8 * i) : hi >> (8 * (i - 8))) & 255u);
}

// The reference's precomputed least-squares weight constants are DECIMAL literals (6 digits), not the exact
// products, so they are data of the algorithm (:133-137): {w*w, (1-w)*w, (1-w)*(1-w), w} per selector.
#if defined(__CUDA_ARCH__)
#define B7_CONST __constant__
#else
#define B7_CONST static const
#endif
B7_CONST float kLsq3[8][4] = {
	{0.000000f, 0.000000f, 1.000000f, 0.000000f}, {0.019775f, 0.120850f, 0.738525f, 0.140625f},
	{0.079102f, 0.202148f, 0.516602f, 0.281250f}, {0.177979f, 0.243896f, 0.334229f, 0.421875f},
	{0.334229f, 0.243896f, 0.177979f, 0.578125f}, {0.516602f, 0.202148f, 0.079102f, 0.718750f},
	{0.738525f, 0.120850f, 0.019775f, 0.859375f}, {1.000000f, 0.000000f, 0.000000f, 1.000000f}};
B7_CONST float kLsq4[16][4] = {
	{0.000000f, 0.000000f, 1.000000f, 0.000000f}, {0.003906f, 0.058594f, 0.878906f, 0.062500f},
	{0.019775f, 0.120850f, 0.738525f, 0.140625f}, {0.041260f, 0.161865f, 0.635010f, 0.203125f},
	{0.070557f, 0.195068f, 0.539307f, 0.265625f}, {0.107666f, 0.220459f, 0.451416f, 0.328125f},
	{0.165039f, 0.241211f, 0.352539f, 0.406250f}, {0.219727f, 0.249023f, 0.282227f, 0.468750f},
	{0.282227f, 0.249023f, 0.219727f, 0.531250f}, {0.352539f, 0.241211f, 0.165039f, 0.593750f},
	{0.451416f, 0.220459f, 0.107666f, 0.671875f}, {0.539307f, 0.195068f, 0.070557f, 0.734375f},
	{0.635010f, 0.161865f, 0.041260f, 0.796875f}, {0.738525f, 0.120850f, 0.019775f, 0.859375f},
	{0.878906f, 0.058594f, 0.003906f, 0.937500f}, {1.000000f, 0.000000f, 0.000000f, 1.000000f}};

// The partition scan order and the filterbank predictor masks are tuning data of the estimator (:1167-1228).
B7_CONST uint8_t kScanOrder[64] = {0, 13, 1, 2, 15, 14, 10, 16, 3, 23, 26, 6, 7, 21, 19, 29, 8, 4, 9, 20, 5, 31,
																	 22, 17, 18, 11, 12, 30, 24, 25, 28, 27, 32, 33, 34, 45, 46, 51, 49, 50, 48, 38,
																	 39, 37, 53, 52, 54, 36, 57, 58, 55, 41, 40, 42, 43, 59, 44, 56, 47, 35, 60, 63, 62, 61};
#define B7_BITS(...) b7_bits_of(__VA_ARGS__)
B7_CONST uint32_t kPredictor[35] = {
	0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0x00000106u, 0x0000008au, 0xffffffffu,
	0xffffffffu, 0x00010104u, 0x00008088u, 0xffffffffu, 0x00014100u, 0x0000c080u, 0xffffffffu, 0xffffffffu,
	0xffffffffu, 0xffffffffu, 0x0000c000u, 0x00414000u, 0x01024000u, 0x0000c006u, 0xffffffffu, 0x0041400au,
	0xffffffffu, 0x01028006u, 0x0040000au, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0x0003c000u, 0xffffffffu,
	0xffffffffu, 0x0900001eu, 0x0803c800u};
#undef B7_BITS

// 2-subset partition shapes / anchors: BPTC specification tables (generated, see tools/gen_bc7_tables.py)
B7_CONST uint16_t kPart2[64] = {
	0xcccc, 0x8888, 0xeeee, 0xecc8, 0xc880, 0xfeec, 0xfec8, 0xec80, 0xc800, 0xffec, 0xfe80, 0xe800, 0xffe8, 0xff00, 0xfff0, 0xf000,
	0xf710, 0x008e, 0x7100, 0x08ce, 0x008c, 0x7310, 0x3100, 0x8cce, 0x088c, 0x3110, 0x6666, 0x366c, 0x17e8, 0x0ff0, 0x718e, 0x399c,
	0xaaaa, 0xf0f0, 0x5a5a, 0x33cc, 0x3c3c, 0x55aa, 0x9696, 0xa55a, 0x73ce, 0x13c8, 0x324c, 0x3bdc, 0x6996, 0xc33c, 0x9966, 0x0660,
	0x0272, 0x04e4, 0x4e40, 0x2720, 0xc936, 0x936c, 0x39c6, 0x639c, 0x9336, 0x9cc6, 0x817e, 0xe718, 0xccf0, 0x0fcc, 0x7744, 0xee22};
B7_CONST uint8_t kAnchor2[64] = {15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 2, 8, 2, 2, 8,
																 8, 15, 2, 8, 2, 2, 8, 8, 2, 2, 15, 15, 6, 8, 2, 8, 15, 15, 2, 8, 2, 2, 2, 15,
																 15, 6, 6, 2, 6, 8, 15, 15, 2, 2, 15, 15, 15, 15, 15, 2, 2, 15};

// ---- colour cell --------------------------------------------------------------------------------------
struct Cell {
	uint32_t px[16];
	int n;
};

struct Best {
	uint64_t err;     // UINT64_MAX = nothing evaluated yet
	uint32_t lo, hi;  // packed endpoints at the mode's endpoint precision
	uint32_t pbit[2];
	uint64_t sel;
};

template <int MODE> struct ModeTraits;
template <> struct ModeTraits<6> { enum { kSelectors = 16, kCompBits = 7, kSharedPbit = 0 }; };
template <> struct ModeTraits<1> { enum { kSelectors = 8, kCompBits = 6, kSharedPbit = 1 }; };

B7_HD uint64_t dist_rgb(uint32_t a, uint32_t b, const Params &P, bool perceptual) { // :325-349
	int dr, dg, db;
	const int ar = (int) ch(a, 0), ag = (int) ch(a, 1), ab = (int) ch(a, 2);
	const int br = (int) ch(b, 0), bg = (int) ch(b, 1), bb = (int) ch(b, 2);
	if (perceptual) {
		const int l1 = ar * 109 + ag * 366 + ab * 37, l2 = br * 109 + bg * 366 + bb * 37;
		dr = (l1 - l2) >> 8;
		dg = (((ar << 9) - l1) - ((br << 9) - l2)) >> 8;
		db = (((ab << 9) - l1) - ((bb << 9) - l2)) >> 8;
	} else {
		dr = ar - br;
		dg = ag - bg;
		db = ab - bb;
	}
	// the reference sums three u32 products in u32, then widens
	return (uint64_t) (uint32_t) (P.w[0] * (uint32_t) (dr * dr) + P.w[1] * (uint32_t) (dg * dg) + P.w[2] * (uint32_t) (db * db));
}
B7_HD uint64_t dist_rgba(uint32_t a, uint32_t b, const Params &P, bool perceptual) { // :351-355
	const int da = (int) ch(a, 3) - (int) ch(b, 3);
	return dist_rgb(a, b, P, perceptual) + (uint64_t) (uint32_t) (P.w[3] * (uint32_t) (da * da));
}

template <int MODE> B7_HD uint32_t expand_endpoint(uint32_t q) { // scale_color (:307-323), all four channels
	const int n = ModeTraits<MODE>::kCompBits + 1;
	uint32_t out = 0;
#pragma unroll
	for (int c = 0; c < 4; c++) {
		uint32_t v = (ch(q, c) << (8 - n)) & 0xffffffffu;
		v |= v >> n;
		out |= (v & 255u) << (8 * c);
	}
	return out;
}

// evaluate_solution (:405-572): palette from (lo, hi, pbits), nearest selector per texel, total error.
template <int MODE, bool ALPHA>
B7_HDN void evaluate(const Cell &cell, uint32_t lo, uint32_t hi, const uint32_t pbit[2], const Params &P, Best &best) {
	constexpr int N = ModeTraits<MODE>::kSelectors;
	const uint32_t pl = pbit[0], ph = ModeTraits<MODE>::kSharedPbit ? pbit[0] : pbit[1];
	uint32_t qlo = 0, qhi = 0;
#pragma unroll
	for (int c = 0; c < 4; c++) {
		qlo |= (((ch(lo, c) << 1) | pl) & 255u) << (8 * c);
		qhi |= (((ch(hi, c) << 1) | ph) & 255u) << (8 * c);
	}
	const uint32_t alo = expand_endpoint<MODE>(qlo), ahi = expand_endpoint<MODE>(qhi);
	constexpr int NC = ALPHA ? 4 : 3;
	uint32_t pal[N];
	pal[0] = alo;
	pal[N - 1] = ahi;
#pragma unroll
	for (int i = 1; i < N - 1; i++) {
		const uint32_t w = (N == 16) ? weight4(i) : weight3(i);
		uint32_t v = 0;
#pragma unroll
		for (int c = 0; c < NC; c++) v |= (((ch(alo, c) * (64 - w) + ch(ahi, c) * w + 32) >> 6) & 255u) << (8 * c);
		pal[i] = v; // the alpha byte of interior entries is never read when !ALPHA (the reference leaves it unset)
	}
	uint64_t total = 0, sels = 0;
	if (!P.perceptual) {
		const int lr = (int) ch(alo, 0), lg = (int) ch(alo, 1), lb = (int) ch(alo, 2), la = (int) ch(alo, 3);
		const int dr = (int) ch(ahi, 0) - lr, dg = (int) ch(ahi, 1) - lg, db = (int) ch(ahi, 2) - lb, da = (int) ch(ahi, 3) - la;
		const float f = ALPHA ? (float) N / (float) ((float) (dr * dr + dg * dg + db * db + da * da) + .00000125f)
													: (float) N / (float) ((float) (dr * dr + dg * dg + db * db) + .00000125f);
		#pragma unroll 1
		for (int i = 0; i < cell.n; i++) {
			const uint32_t c = cell.px[i];
			int proj = ((int) ch(c, 0) - lr) * dr + ((int) ch(c, 1) - lg) * dg + ((int) ch(c, 2) - lb) * db;
			if (ALPHA) proj += ((int) ch(c, 3) - la) * da;
			int s = (int) ((float) proj * f + .5f);
			s = clampi(s, 1, N - 1);
			const uint64_t e0 = ALPHA ? dist_rgba(pal[s - 1], c, P, false) : dist_rgb(pal[s - 1], c, P, false);
			uint64_t e1 = ALPHA ? dist_rgba(pal[s], c, P, false) : dist_rgb(pal[s], c, P, false);
			// RGBA keeps the upper selector on ties (err1 > err0), RGB too (err0 < err1): same rule
			if (e0 < e1) { e1 = e0; --s; }
			total += e1;
			sels |= sel_put((uint32_t) s, i);
		}
	} else {
		// YCbCr transform of the palette once; per texel an exhaustive scan keeping the first minimum
		int pl_[N], pcr[N], pcb[N], pa[N];
#pragma unroll
		for (int j = 0; j < N; j++) {
			const int r = (int) ch(pal[j], 0), g = (int) ch(pal[j], 1), b = (int) ch(pal[j], 2);
			pl_[j] = r * 109 + g * 366 + b * 37;
			pcr[j] = (r << 9) - pl_[j];
			pcb[j] = (b << 9) - pl_[j];
			pa[j] = (int) ch(pal[j], 3);
		}
		#pragma unroll 1
		for (int i = 0; i < cell.n; i++) {
			const uint32_t c = cell.px[i];
			const int r = (int) ch(c, 0), g = (int) ch(c, 1), b = (int) ch(c, 2), a = (int) ch(c, 3);
			const int l2 = r * 109 + g * 366 + b * 37, cr2 = (r << 9) - l2, cb2 = (b << 9) - l2;
			uint64_t be = ~0ull;
			uint32_t bs = 0;
#pragma unroll
			for (int j = 0; j < N; j++) {
				const int dl = (pl_[j] - l2) >> 8, dcr = (pcr[j] - cr2) >> 8, dcb = (pcb[j] - cb2) >> 8;
				uint64_t e = (uint64_t) (uint32_t) (P.w[0] * (uint32_t) (dl * dl) + P.w[1] * (uint32_t) (dcr * dcr) + P.w[2] * (uint32_t) (dcb * dcb));
				if (ALPHA) {
					const int da = pa[j] - a;
					e += (uint64_t) (uint32_t) (P.w[3] * (uint32_t) (da * da));
				}
				if (e < be) { be = e; bs = (uint32_t) j; }
			}
			total += be;
			sels |= sel_put(bs, i);
		}
	}
	if (total < best.err) {
		best.err = total;
		best.lo = lo;
		best.hi = hi;
		best.pbit[0] = pbit[0];
		best.pbit[1] = pbit[1];
		best.sel = sels;
	}
}

// find_optimal_solution (:606-729) for the p-bit modes (1 and 6 both carry p-bits): round the float endpoints
// to each p-bit lattice, keep the closer, fix degenerate mode-1 channels, evaluate unless it is the current best.
template <int MODE, bool ALPHA>
B7_HDN uint64_t try_endpoints(const Cell &cell, F4 xl, F4 xh, const Params &P, Best &best) {
#pragma unroll
	for (int c = 0; c < 4; c++) { xl.v[c] = sat(xl.v[c]); xh.v[c] = sat(xh.v[c]); }
	constexpr int iscalep = (1 << (ModeTraits<MODE>::kCompBits + 1)) - 1;
	const float scalep = (float) iscalep;
	constexpr int NC = ALPHA ? 4 : 3;
	uint32_t bp[2] = {0, 0}, blo = 0, bhi = 0;
	if (!ModeTraits<MODE>::kSharedPbit) {
		float be0 = 1e+9f, be1 = 1e+9f;
		for (int p = 0; p < 2; p++) {
			uint32_t mn = 0, mx = 0;
#pragma unroll
			for (int c = 0; c < 4; c++) {
				mn |= (uint32_t) clampi(((int) ((xl.v[c] * scalep - (float) p) / 2.0f + .5f)) * 2 + p, p, iscalep - 1 + p) << (8 * c);
				mx |= (uint32_t) clampi(((int) ((xh.v[c] * scalep - (float) p) / 2.0f + .5f)) * 2 + p, p, iscalep - 1 + p) << (8 * c);
			}
			const uint32_t sl = expand_endpoint<MODE>(mn), sh = expand_endpoint<MODE>(mx);
			float e0 = 0, e1 = 0;
#pragma unroll
			for (int c = 0; c < NC; c++) {
				const float a = (float) (int) ch(sl, c) - xl.v[c] * 255.0f, b = (float) (int) ch(sh, c) - xh.v[c] * 255.0f;
				e0 += a * a;
				e1 += b * b;
			}
			if (e0 < be0) { be0 = e0; bp[0] = (uint32_t) p; blo = (mn >> 1) & 0x7f7f7f7fu; }
			if (e1 < be1) { be1 = e1; bp[1] = (uint32_t) p; bhi = (mx >> 1) & 0x7f7f7f7fu; }
		}
	} else {
		float be = 1e+9f;
		for (int p = 0; p < 2; p++) {
			uint32_t mn = 0, mx = 0;
#pragma unroll
			for (int c = 0; c < 4; c++) {
				mn |= (uint32_t) clampi(((int) ((xl.v[c] * scalep - (float) p) / 2.0f + .5f)) * 2 + p, p, iscalep - 1 + p) << (8 * c);
				mx |= (uint32_t) clampi(((int) ((xh.v[c] * scalep - (float) p) / 2.0f + .5f)) * 2 + p, p, iscalep - 1 + p) << (8 * c);
			}
			const uint32_t sl = expand_endpoint<MODE>(mn), sh = expand_endpoint<MODE>(mx);
			float e = 0;
#pragma unroll
			for (int c = 0; c < NC; c++) {
				const float a = ((float) (int) ch(sl, c) / 255.0f) - xl.v[c], b = ((float) (int) ch(sh, c) / 255.0f) - xh.v[c];
				e += a * a + b * b;
			}
			if (e < be) { be = e; bp[0] = bp[1] = (uint32_t) p; blo = (mn >> 1) & 0x7f7f7f7fu; bhi = (mx >> 1) & 0x7f7f7f7fu; }
		}
	}
	if (MODE == 1) { // fixDegenerateEndpoints (:574-604); iscale = iscalep >> 1
		constexpr uint32_t iscale = (uint32_t) (iscalep >> 1);
#pragma unroll
		for (int c = 0; c < 3; c++) {
			uint32_t a = ch(blo, c), b = ch(bhi, c);
			if (a == b && fabsf(xl.v[c] - xh.v[c]) > 0.0f) {
				if (a > (iscale >> 1)) {
					if (a > 0) a--;
					else if (b < iscale) b++;
				} else {
					if (b < iscale) b++;
					else if (a > 0) a--;
				}
				blo = (blo & ~(255u << (8 * c))) | (a << (8 * c));
				bhi = (bhi & ~(255u << (8 * c))) | (b << (8 * c));
			}
		}
	}
	if (best.err == ~0ull || blo != best.lo || bhi != best.hi || bp[0] != best.pbit[0] || bp[1] != best.pbit[1])
		evaluate<MODE, ALPHA>(cell, blo, bhi, bp, P, best);
	return best.err;
}

// compute_least_squares_endpoints_rgb[a] (:197-280), result already scaled by 1/255 (:889-890)
template <int MODE, bool ALPHA>
B7_HDN void least_squares(const Cell &cell, uint64_t sels, F4 &xl, F4 &xh) {
	float z00 = 0.f, z10 = 0.f, z11 = 0.f;
	float q00[4] = {0.f, 0.f, 0.f, 0.f}, t[4] = {0.f, 0.f, 0.f, 0.f};
	constexpr int NC = ALPHA ? 4 : 3;
	#pragma unroll 1
	for (int i = 0; i < cell.n; i++) {
		const uint32_t s = sel_get(sels, i);
		const float *wx = (ModeTraits<MODE>::kSelectors == 16) ? kLsq4[s] : kLsq3[s];
		z00 += wx[0];
		z10 += wx[1];
		z11 += wx[2];
		const float w = wx[3];
#pragma unroll
		for (int c = 0; c < NC; c++) {
			const float v = (float) (int) ch(cell.px[i], c);
			q00[c] += w * v;
			t[c] += v;
		}
	}
	const float z01 = z10;
	float det = z00 * z11 - z01 * z10;
	if (det != 0.0f) det = 1.0f / det;
	const float iz00 = z11 * det, iz01 = -z01 * det, iz10 = -z10 * det, iz11 = z00 * det;
#pragma unroll
	for (int c = 0; c < NC; c++) {
		const float q10 = t[c] - q00[c];
		xl.v[c] = iz00 * q00[c] + iz01 * q10;
		xh.v[c] = iz10 * q00[c] + iz11 * q10;
	}
	if (!ALPHA) xl.v[3] = xh.v[3] = 255.0f;
#pragma unroll
	for (int c = 0; c < 4; c++) {
		xl.v[c] = xl.v[c] * (1.0f / 255.0f);
		xh.v[c] = xh.v[c] * (1.0f / 255.0f);
	}
}

// pack_mode1_to_one_color (:357-403)
B7_HD uint64_t single_colour_mode1(const Cell &cell, uint32_t r, uint32_t g, uint32_t b, const Params &P, Best &out) {
	uint32_t be = 0xffffffffu, bp = 0;
	for (uint32_t p = 0; p < 2; p++) {
		const uint32_t e = (uint32_t) P.opt1[r * 2 + p].err + P.opt1[g * 2 + p].err + P.opt1[b * 2 + p].err;
		if (e < be) { be = e; bp = p; }
	}
	const OptimalEndpoint er = P.opt1[r * 2 + bp], eg = P.opt1[g * 2 + bp], eb = P.opt1[b * 2 + bp];
	out.lo = pack4(er.lo, eg.lo, eb.lo, 0);
	out.hi = pack4(er.hi, eg.hi, eb.hi, 0);
	out.pbit[0] = bp;
	out.pbit[1] = 0;
	uint64_t sels = 0;
	#pragma unroll 1
	for (int i = 0; i < cell.n; i++) sels |= sel_put(2, i);
	out.sel = sels;
	uint32_t pc = 255u << 24;
#pragma unroll
	for (int c = 0; c < 3; c++) {
		uint32_t lo = ((ch(out.lo, c) << 1) | bp) << 1;
		lo |= lo >> 7;
		uint32_t hi = ((ch(out.hi, c) << 1) | bp) << 1;
		hi |= hi >> 7;
		pc |= (((lo * (64 - 18) + hi * 18 + 32) >> 6) & 255u) << (8 * c);
	}
	uint64_t total = 0;
	#pragma unroll 1
	for (int i = 0; i < cell.n; i++) total += dist_rgb(pc, cell.px[i], P, P.perceptual != 0);
	out.err = total;
	return total;
}

// color_cell_compression (:731-1024)
template <int MODE, bool ALPHA>
B7_HDN uint64_t compress_cell(const Cell &cell, const Params &P, Best &best) {
	constexpr int N = ModeTraits<MODE>::kSelectors;
	best.err = ~0ull;
	best.lo = best.hi = 0;
	best.pbit[0] = best.pbit[1] = 0;
	best.sel = 0;
	const int n = cell.n;
	if (MODE == 1) {
		const uint32_t rgb0 = cell.px[0] & 0xffffffu;
		bool same = true;
		#pragma unroll 1
		for (int i = 1; i < n; i++) same = same && ((cell.px[i] & 0xffffffu) == rgb0);
		if (same) return single_colour_mode1(cell, ch(rgb0, 0), ch(rgb0, 1), ch(rgb0, 2), P, best);
	}
	F4 sum = {{0.f, 0.f, 0.f, 0.f}};
	#pragma unroll 1
	for (int i = 0; i < n; i++) {
#pragma unroll
		for (int c = 0; c < 4; c++) sum.v[c] = sum.v[c] + (float) (int) ch(cell.px[i], c);
	}
	F4 mean_s, mean;
	const float inv_n = 1.0f / (float) (uint32_t) n, inv_n255 = 1.0f / (float) ((float) (uint32_t) n * 255.0f);
#pragma unroll
	for (int c = 0; c < 4; c++) {
		mean_s.v[c] = sum.v[c] * inv_n;
		mean.v[c] = sat(sum.v[c] * inv_n255);
	}
	F4 axis = {{0.f, 0.f, 0.f, 0.f}};
	if (ALPHA) { // incremental PCA (:771-791)
		#pragma unroll 1
		for (int i = 0; i < n; i++) {
			F4 col;
#pragma unroll
			for (int c = 0; c < 4; c++) col.v[c] = (float) (int) ch(cell.px[i], c) - mean_s.v[c];
			F4 nrm = i ? axis : col;
			normalise(nrm);
#pragma unroll
			for (int c = 0; c < 4; c++) {
				F4 a;
#pragma unroll
				for (int k = 0; k < 4; k++) a.v[k] = col.v[k] * col.v[c];
				axis.v[c] += dot4(a, nrm);
			}
		}
		normalise(axis);
	} else { // covariance + 3 power iterations (:795-832)
		float cov[6] = {0, 0, 0, 0, 0, 0};
		#pragma unroll 1
		for (int i = 0; i < n; i++) {
			const float r = (float) (int) ch(cell.px[i], 0) - mean_s.v[0];
			const float g = (float) (int) ch(cell.px[i], 1) - mean_s.v[1];
			const float b = (float) (int) ch(cell.px[i], 2) - mean_s.v[2];
			cov[0] += r * r; cov[1] += r * g; cov[2] += r * b; cov[3] += g * g; cov[4] += g * b; cov[5] += b * b;
		}
		float vr = .9f, vg = 1.0f, vb = .7f;
		for (int it = 0; it < 3; it++) {
			float r = vr * cov[0] + vg * cov[1] + vb * cov[2];
			float g = vr * cov[1] + vg * cov[3] + vb * cov[4];
			float b = vr * cov[2] + vg * cov[4] + vb * cov[5];
			float m = fabsf(r) > fabsf(g) ? fabsf(r) : fabsf(g);
			m = m > fabsf(b) ? m : fabsf(b);
			if (m > 1e-10f) {
				m = 1.0f / m;
				r *= m; g *= m; b *= m;
			}
			vr = r; vg = g; vb = b;
		}
		float len = vr * vr + vg * vg + vb * vb;
		if (!(len < 1e-10f)) {
			len = 1.0f / sqrtf(len);
			axis.v[0] = vr * len; axis.v[1] = vg * len; axis.v[2] = vb * len; axis.v[3] = 0.f;
		}
	}
	if (dot4(axis, axis) < .5f) {
		if (P.perceptual) { axis.v[0] = .213f; axis.v[1] = .715f; axis.v[2] = .072f; axis.v[3] = ALPHA ? .715f : 0.f; }
		else { axis.v[0] = axis.v[1] = axis.v[2] = 1.0f; axis.v[3] = ALPHA ? 1.0f : 0.f; }
		normalise(axis);
	}
	float l = 1e+9f, h = -1e+9f;
	#pragma unroll 1
	for (int i = 0; i < n; i++) {
		F4 q;
#pragma unroll
		for (int c = 0; c < 4; c++) q.v[c] = (float) (int) ch(cell.px[i], c) - mean_s.v[c];
		const float d = dot4(q, axis);
		l = l < d ? l : d;
		h = h > d ? h : d;
	}
	l *= (1.0f / 255.0f);
	h *= (1.0f / 255.0f);
	F4 mn, mx;
#pragma unroll
	for (int c = 0; c < 4; c++) {
		mn.v[c] = sat(mean.v[c] + axis.v[c] * l);
		mx.v[c] = sat(mean.v[c] + axis.v[c] * h);
	}
	const F4 ones = {{1.0f, 1.0f, 1.0f, 1.0f}};
	if (dot4(mn, ones) > dot4(mx, ones)) { const F4 t = mn; mn = mx; mx = t; }

	if (!try_endpoints<MODE, ALPHA>(cell, mn, mx, P, best)) return 0;
	F4 xl, xh;
	least_squares<MODE, ALPHA>(cell, best.sel, xl, xh); // m_try_least_squares is always on
	if (!try_endpoints<MODE, ALPHA>(cell, xl, xh, P, best)) return 0;

	if (P.uber > 0) { // :896-1007
		const uint64_t base = best.sel;
		uint32_t smin = 16, smax = 0;
		#pragma unroll 1
		for (int i = 0; i < n; i++) {
			const uint32_t s = sel_get(base, i);
			smin = s < smin ? s : smin;
			smax = s > smax ? s : smax;
		}
		for (int variant = 0; variant < 3; variant++) { // raise the lowest, lower the highest, both
			uint64_t t = 0;
			#pragma unroll 1
			for (int i = 0; i < n; i++) {
				uint32_t s = sel_get(base, i);
				if (variant != 1 && s == smin && s < (uint32_t) (N - 1)) s++;
				else if (variant != 0 && s == smax && s > 0) s--;
				t |= sel_put(s, i);
			}
			least_squares<MODE, ALPHA>(cell, t, xl, xh);
			if (!try_endpoints<MODE, ALPHA>(cell, xl, xh, P, best)) return 0;
		}
		const uint32_t thresh = ((uint32_t) n * 56) >> 4;
		if (P.uber >= 2 && best.err > thresh) {
			const int Q = P.uber >= 4 ? P.uber - 2 : 1;
			constexpr int top = N - 1;
			for (int ly = -Q; ly <= 1; ly++)
				for (int hy = top - 1; hy <= top + Q; hy++) {
					if (ly == 0 && hy == top) continue;
					uint64_t t = 0;
					#pragma unroll 1
					for (int i = 0; i < n; i++) {
						float v = floorf((float) top * ((float) sel_get(base, i) - (float) ly) / ((float) hy - (float) ly) + .5f);
						v = v < 0.f ? 0.f : (v > (float) top ? (float) top : v);
						t |= sel_put((uint32_t) (uint8_t) v, i);
					}
					least_squares<MODE, ALPHA>(cell, t, xl, xh);
					if (!try_endpoints<MODE, ALPHA>(cell, xl, xh, P, best)) return 0;
				}
		}
	}
	if (MODE == 1) { // single colour at the mean (:1009-1021)
		Best avg = best;
		const uint32_t r = (uint32_t) (int) (.5f + mean.v[0] * 255.0f), g = (uint32_t) (int) (.5f + mean.v[1] * 255.0f),
									 b = (uint32_t) (int) (.5f + mean.v[2] * 255.0f);
		const uint64_t e = single_colour_mode1(cell, r, g, b, P, avg);
		if (e < best.err) best = avg;
	}
	return best.err;
}

// color_cell_compression_est (:1026-1162) evaluated completely (no early-out): bounding-box diagonal, 8-point ramp
B7_HDN uint64_t estimate_subset(const uint32_t *px, uint32_t mask, const Params &P) {
	uint32_t lo[3] = {255, 255, 255}, hi[3] = {0, 0, 0};
#pragma unroll 1
	for (int i = 0; i < 16; i++)
		if ((mask >> i) & 1u) {
#pragma unroll
			for (int c = 0; c < 3; c++) {
				const uint32_t v = ch(px[i], c);
				lo[c] = v < lo[c] ? v : lo[c];
				hi[c] = v > hi[c] ? v : hi[c];
			}
		}
	int pal[8][3], dots[8], thr[7];
	const int ar = (int) hi[0] - (int) lo[0], ag = (int) hi[1] - (int) lo[1], ab = (int) hi[2] - (int) lo[2];
#pragma unroll
	for (int j = 0; j < 8; j++) {
		const uint32_t w = weight3(j);
#pragma unroll
		for (int c = 0; c < 3; c++)
			pal[j][c] = (j == 0) ? (int) lo[c] : (j == 7 ? (int) hi[c] : (int) (((lo[c] * (64 - w) + hi[c] * w + 32) >> 6) & 255u));
		dots[j] = pal[j][0] * ar + pal[j][1] * ag + pal[j][2] * ab;
	}
#pragma unroll
	for (int j = 0; j < 7; j++) thr[j] = (dots[j] + dots[j + 1] + 1) >> 1;
	uint64_t total = 0;
#pragma unroll 1
	for (int i = 0; i < 16; i++)
		if ((mask >> i) & 1u) {
			const int r = (int) ch(px[i], 0), g = (int) ch(px[i], 1), b = (int) ch(px[i], 2);
			const int d = ar * r + ag * g + ab * b;
			int s = 0;
#pragma unroll
			for (int j = 0; j < 7; j++) s += (d >= thr[j]) ? 1 : 0; // thresholds are non-decreasing: count == cascade
			int pr = pal[0][0], pg = pal[0][1], pb = pal[0][2];
#pragma unroll
			for (int j = 1; j < 8; j++)
				if (s == j) { pr = pal[j][0]; pg = pal[j][1]; pb = pal[j][2]; }
			if (P.perceptual) {
				const int l1 = pr * 109 + pg * 366 + pb * 37, l2 = r * 109 + g * 366 + b * 37;
				const int dl = (l1 - l2) >> 8, dcr = (((pr << 9) - l1) - ((r << 9) - l2)) >> 8, dcb = (((pb << 9) - l1) - ((b << 9) - l2)) >> 8;
				const int ie = (int) (P.w[0] * (uint32_t) (dl * dl) + P.w[1] * (uint32_t) (dcr * dcr) + P.w[2] * (uint32_t) (dcb * dcb));
				total += (uint64_t) (int64_t) ie;
			} else {
				const int dr = pr - r, dg = pg - g, db = pb - b;
				total += (uint64_t) (uint32_t) (P.w[0] * (uint32_t) (dr * dr) + P.w[1] * (uint32_t) (dg * dg) + P.w[2] * (uint32_t) (db * db));
			}
		}
	return total;
}

// estimate_partition (:1207-1281)
B7_HDN uint32_t estimate_partition(const uint32_t *px, const Params &P) {
	uint64_t best_err = ~0ull;
	uint32_t best = 0;
	int key = 0;
	for (int it = 0; it < 64 && best_err > 0; it++) {
		const uint32_t part = kScanOrder[it];
		if (it >= 14 && it <= 34) {
			if ((kPredictor[part] & (1u << (key + 1))) == 0) {
				if (it == 34) break;
				continue;
			}
		}
		const uint32_t m1 = kPart2[part];
		const uint64_t e0 = estimate_subset(px, ~m1 & 0xffffu, P);
		uint64_t tot = e0;
		if (tot < best_err) tot += estimate_subset(px, m1, P);
		if (tot < best_err) { best_err = tot; best = part; }
		if (part == 34 && best != 34) break;
		if (it == 13) key = (int) best;
	}
	return best;
}

// encode_bc7_block (:1307-1388), packing into two 64-bit words
struct BitSink {
	uint64_t w[2];
	uint32_t pos;
};
B7_HD void put(BitSink &s, uint32_t val, uint32_t bits) {
	const uint32_t p = s.pos;
	if (p < 64) {
		s.w[0] |= (uint64_t) val << p;
		if (p + bits > 64) s.w[1] |= (uint64_t) val >> (64 - p);
	} else {
		s.w[1] |= (uint64_t) val << (p - 64);
	}
	s.pos = p + bits;
}

struct BlockResult {
	uint32_t mode, partition;
	uint64_t sel;        // per texel (block order), 4 bits each
	uint32_t lo[2], hi[2];
	uint32_t pbit[2][2];
};

B7_HD void pack_block(const BlockResult &r, uint64_t out[2]) {
	const bool m1 = r.mode == 1;
	const int subsets = m1 ? 2 : 1;
	const uint32_t ibits = m1 ? 3 : 4, nsel = 1u << ibits;
	const uint32_t pmask = m1 ? kPart2[r.partition] : 0u;
	uint64_t sel = r.sel;
	uint32_t lo[2] = {r.lo[0], r.lo[1]}, hi[2] = {r.hi[0], r.hi[1]};
	uint32_t pb[2][2] = {{r.pbit[0][0], r.pbit[0][1]}, {r.pbit[1][0], r.pbit[1][1]}};
	int anchor[2] = {0, -1};
	for (int k = 0; k < subsets; k++) {
		const int a = k ? (int) kAnchor2[r.partition] : 0;
		anchor[k] = a;
		if (sel_get(sel, a) & (nsel >> 1)) {
			for (int i = 0; i < 16; i++)
				if (((pmask >> i) & 1u) == (uint32_t) k) {
					const uint32_t s = sel_get(sel, i);
					sel = (sel & ~(15ull << (4 * i))) | sel_put((nsel - 1) - s, i);
				}
			const uint32_t t = lo[k];
			lo[k] = hi[k];
			hi[k] = t;
			if (!m1) { const uint32_t u = pb[k][0]; pb[k][0] = pb[k][1]; pb[k][1] = u; }
		}
	}
	BitSink s = {{0, 0}, 0};
	put(s, 1u << r.mode, r.mode + 1);
	if (m1) put(s, r.partition, 6);
	const int comps = m1 ? 3 : 4;
	const uint32_t cbits = m1 ? 6 : 7;
	for (int c = 0; c < comps; c++)
		for (int k = 0; k < subsets; k++) {
			put(s, ch(lo[k], c), cbits);
			put(s, ch(hi[k], c), cbits);
		}
	for (int k = 0; k < subsets; k++) {
		put(s, pb[k][0], 1);
		if (!m1) put(s, pb[k][1], 1);
	}
	for (int i = 0; i < 16; i++) {
		const uint32_t nb = ibits - ((i == anchor[0] || i == anchor[1]) ? 1u : 0u);
		put(s, sel_get(sel, i), nb);
	}
	out[0] = s.w[0];
	out[1] = s.w[1];
}

// bc7enc16_compress_block (:1517-1547) with handle_alpha_block (:1390-1420) / handle_opaque_block (:1422-1515).
// The reference leaves m_endpoints_share_pbit uninitialised on the alpha path; the contract value is `false`
// (SURVEY.md 3.5), which is what ModeTraits<6> encodes.
B7_HD void encode_block(const uint32_t px[16], const Params &P, uint64_t out[2]) {
	bool alpha = false;
#pragma unroll
	for (int i = 0; i < 16; i++) alpha = alpha || ((px[i] >> 24) < 255u);
	Cell cell;
#pragma unroll
	for (int i = 0; i < 16; i++) cell.px[i] = px[i];
	cell.n = 16;
	BlockResult r;
	r.mode = 6;
	r.partition = 0;
	r.lo[1] = r.hi[1] = 0;
	r.pbit[1][0] = r.pbit[1][1] = 0;
	Best b6;
	bool mode1 = false;
	if (alpha) {
		compress_cell<6, true>(cell, P, b6);
	} else {
		const uint64_t err6 = compress_cell<6, false>(cell, P, b6);
		if (err6 > 0) {
			const uint32_t part = estimate_partition(px, P);
			const uint32_t mask = kPart2[part];
			Best bs[2];
			uint64_t trial = 0;
#pragma unroll 1
			for (int k = 0; k < 2; k++) {
				Cell sub;
				sub.n = 0;
#pragma unroll 1
				for (int i = 0; i < 16; i++)
					if (((mask >> i) & 1u) == (uint32_t) k) sub.px[sub.n++] = px[i];
				trial += compress_cell<1, false>(sub, P, bs[k]);
				if (trial > err6) break;
			}
			if (trial < err6) {
				mode1 = true;
				r.mode = 1;
				r.partition = part;
				uint64_t sel = 0;
				int cnt[2] = {0, 0};
#pragma unroll 1
				for (int i = 0; i < 16; i++) {
					const int k = (int) ((mask >> i) & 1u);
					sel |= sel_put(sel_get(bs[k].sel, cnt[k]++), i);
				}
				r.sel = sel;
#pragma unroll
				for (int k = 0; k < 2; k++) {
					r.lo[k] = bs[k].lo;
					r.hi[k] = bs[k].hi;
					r.pbit[k][0] = bs[k].pbit[0];
					r.pbit[k][1] = 0;
				}
			}
		}
	}
	if (!mode1) {
		r.sel = b6.sel;
		r.lo[0] = b6.lo;
		r.hi[0] = b6.hi;
		r.pbit[0][0] = b6.pbit[0];
		r.pbit[0][1] = b6.pbit[1];
	}
	pack_block(r, out);
}

} // namespace rg
} // namespace b200ic
