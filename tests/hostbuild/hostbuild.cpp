// TEST/DEBUG AID ONLY -- compiles the host-portable per-block encoder cores (csrc/*_core.cuh) with g++ so that
// their logic can be checked against the oracle in the CPU-only container. Never linked into the product library
// and never used by the package: the product path is the CUDA kernels.
#include <stdint.h>
#include <string.h>
#ifdef HB_BC7RG
#include "bc7rg_core.cuh"
#endif

#ifdef HB_BC7AMD
#include "bc7amd_block.cuh"
#include <thread>
#include <vector>
#include <atomic>
#include <cmath>
#endif

#ifdef HB_BC1
#include "bc1_core.cuh"
#endif
#ifdef HB_BC6H
#include "bc6h_core.cuh"
#endif

extern "C" {
#ifdef HB_BC6H
// in: nblocks x 64 floats (RGBA, linear HDR); out: nblocks x 16 bytes
void hb_bc6h_blocks(const float *in, uint64_t nblocks, int is_signed, uint8_t *out) {
	for (uint64_t b = 0; b < nblocks; b++) {
		uint64_t w[2];
		b200ic::bc6::encode_block_serial(in + b * 64, is_signed != 0, w);
		memcpy(out + b * 16, w, 16);
	}
}
#endif
#ifdef HB_BC1
// in: nblocks x 64 floats RGBA 0..1; out: nblocks x 8 bytes
void hb_bc1_blocks(const float *in, uint64_t nblocks, float alpha_threshold, int steps, uint8_t *out) {
	for (uint64_t b = 0; b < nblocks; b++) {
		uint32_t w[2];
		b200ic::bc1::encode_block(in + b * 64, alpha_threshold, steps, w);
		memcpy(out + b * 8, w, 8);
	}
}
#endif
#ifdef HB_BC1
// colour half of BC2 / BC3 as this repo defines it: the BC1 4-point fit without punch-through; in: nblocks x 64 floats
void hb_bc23_colour_blocks(const float *in, uint64_t nblocks, int steps, uint8_t *out) {
	for (uint64_t b = 0; b < nblocks; b++) {
		uint8_t ep[3][2], idx[16];
		uint32_t w[2];
		b200ic::bc1::compress(in + b * 64, 4, false, 0.0f, steps, ep, idx);
		b200ic::bc1::pack_fit(1, ep, idx, w);
		memcpy(out + b * 8, w, 8);
	}
}
#endif
#ifdef HB_BC7AMD
static uint32_t *hb_sp_table() {
	static std::vector<uint32_t> sp;
	if (sp.empty()) { sp.resize(b200ic::amd7::kSpEntries); b200ic::amd7::build_single_point_table(sp.data()); }
	return sp.data();
}
// in: nblocks x 64 floats (RGBA 0..1, texel order); out: nblocks x 16 bytes; err (may be NULL): encoder's SSE per block
void hb_bc7amd_blocks(const float *in, uint64_t nblocks, uint32_t mode_mask, uint8_t *out, double *err, int nthreads, int u8_path) {
	b200ic::amd7::Tables T{hb_sp_table()};
	std::atomic<uint64_t> next{0};
	auto work = [&]() {
		for (;;) {
			const uint64_t b = next.fetch_add(1);
			if (b >= nblocks) break;
			uint64_t w[2];
			const double e = u8_path ? b200ic::amd7::encode_block_serial<true>(T, in + b * 64, mode_mask, w)
			                         : b200ic::amd7::encode_block_serial<false>(T, in + b * 64, mode_mask, w);
			memcpy(out + b * 16, w, 16);
			if (err) err[b] = e;
		}
	};
	if (nthreads <= 1) { work(); return; }
	std::vector<std::thread> pool;
	for (int t = 0; t < nthreads; t++) pool.emplace_back(work);
	for (auto &t : pool) t.join();
}
#endif
#ifdef HB_BC7AMD
}
// The lane = corner form of the cube walk (bc7amd_int.cuh, what the CUDA cube kernel runs) emulated lane by lane
// against the serial cube_search_u8 on random items. Returns the number of items whose (key, indices) differ.
template <int CLOG> static int hb_cube_lane_trial(uint64_t &rng, int bits, int type) {
	using namespace b200ic::amd7;
	auto next = [&]() { rng = rng * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t) (rng >> 33); };
	constexpr int C = 1 << CLOG;
	const int n = 1 + (int) (next() % 16);
	uint32_t d[16] = {};
	const uint32_t base = next(), spread = 1u << (next() % 8);
	for (int i = 0; i < n; i++) {
		uint32_t v = 0;
		for (int j = 0; j < 3; j++) v |= ((((base >> (8 * j)) & 255u) + next() % spread) & 255u) << (8 * j);
		d[i] = v;
	}
	// a collapsed index vector with maximum Mi >= 1, and one of its (q, p) re-indexings
	int idx[16], Mi = 0;
	for (int i = 0; i < n; i++) idx[i] = (int) (next() % C);
	idx[0] = 0;
	if (n > 1) idx[n - 1] = 1 + (int) (next() % (C - 1));
	else return 0;
	Mi = collapse_indices(idx, n);
	if (Mi == 0) return 0;
	uint64_t cur = 0;
	for (int i = 0; i < n; i++) cur |= (uint64_t) idx[i] << (4 * i);
	const int count = qp_count(Mi, C - 1);
	int q, p;
	qp_decode((int) (next() % count), Mi, C - 1, q, p);
	const int use_par = (type == BCC || type == SAME_PAR) ? 1 : 0, bcc = type == BCC ? 1 : 0;
	const int b3[3] = {bits, bits, bits};
	uint32_t want_key;
	uint64_t want_idx;
	cube_item_u8<CLOG>(d, n, cur, q, p, b3, type, 0, 4, want_key, want_idx);
	// lane form
	uint32_t ep[6];
	cube_item_setup_u8<CLOG>(d, n, cur, q, p, bits, use_par, ep);
	const int nl = (use_par + 1) * (bcc + 1), nlb = nl == 4 ? 2 : (nl == 2 ? 1 : 0);
	uint64_t tab[4 * 12] = {};
	constexpr int H = C / 4;
	static std::vector<uint32_t> lut; // the difference-table ramps the kernel uses must equal the computed ones
	if (lut.empty()) { lut.resize(RampLutShape<CLOG>::kWords); ramp_lut_fill<CLOG>(lut.data(), 0, 1); }
	int bad_tab = 0;
	for (int id = 0; id < nl * 12 * H; id++) {
		const uint32_t w = cube_tab_word_lut<CLOG>(lut.data(), ep, bcc, id);
		if (w != cube_tab_word<CLOG>(ep, bcc, id)) bad_tab = 1;
		tab[id / H] |= (uint64_t) w << (32 * (id % H));
	}
	uint32_t best = 0xffffffffu, best_xy = 0;
	unsigned best_lane = 0;
	for (unsigned lane = 0; lane < 32; lane++) {
		uint32_t k, xy;
		cube_lane_corners<CLOG>(tab, d, n, nlb, lane, k, xy);
		if (k < best) { best = k; best_xy = xy; best_lane = lane; }
	}
	uint32_t pal[C];
	cube_lane_palette<CLOG>(tab, nlb, best_lane, best_xy, pal);
	const uint64_t got_idx = palette_indices_u8<CLOG>(d, n, pal);
	// the per-channel bounds of the second-pass pruning: equal to the plain definition, and never above a corner's error
	uint32_t plane[16], lbs[48];
	window_planes_u8(d, n, plane);
	int bad_bound = 0;
	for (int b = 0; b < nl * 12; b++) {
		const int k = (b >> 2) % 3;
		lbs[b] = cube_bound_u8<CLOG>(tab[b], plane + 4 * k, n);
		uint32_t sum = 0;
		for (int i = 0; i < n; i++) {
			int m = 255;
			for (int c = 0; c < C; c++) {
				int a = (int) ((tab[b] >> (8 * c)) & 255u) - (int) ((d[i] >> (8 * k)) & 255u);
				a = a < 0 ? -a : a;
				m = a < m ? a : m;
			}
			sum += (uint32_t) (m * m);
		}
		if (sum != lbs[b]) bad_bound = 1;
	}
	for (int cid = 0; cid < nl * 64; cid++) {
		uint32_t cp[C];
		cube_cid_palette<CLOG>(tab, cid, cp);
		const uint32_t e = cube_corner_error_u8<CLOG>(cp, d, n);
		if (cube_cid_bound(lbs, cid) > e) bad_bound = 1;
		if (cube_cid_key(e, cid) < best) bad_bound = 1; // (the lane walk's minimum is the minimum over the corner ids too)
	}
	return (bad_tab || bad_bound || best != want_key || got_idx != want_idx) ? 1 : 0;
}
// window_item_lut_u8 (table ramps, per-texel sums) against window_item_u8 (cluster sums) on random items
template <int CLOG> static int hb_window_trial(uint64_t &rng, int size, int bits_total, int dim) {
	using namespace b200ic::amd7;
	auto next = [&]() { rng = rng * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t) (rng >> 33); };
	constexpr int C = 1 << CLOG;
	static std::vector<uint32_t> lut;
	if (lut.empty()) { lut.resize(RampLutShape<CLOG>::kWords); ramp_lut_fill<CLOG>(lut.data(), 0, 1); }
	const int n = 2 + (int) (next() % 15);
	uint32_t d[16] = {};
	const uint32_t base = next(), spread = 1u << (next() % 9);
	for (int i = 0; i < n; i++) {
		uint32_t v = 0;
		for (int j = 0; j < dim; j++) v |= ((((base >> (8 * j)) & 255u) + next() % spread) & 255u) << (8 * j);
		d[i] = v;
	}
	int idx[16];
	for (int i = 0; i < n; i++) idx[i] = (int) (next() % C);
	idx[0] = 0;
	idx[n - 1] = 1 + (int) (next() % (C - 1));
	const int Mi = collapse_indices(idx, n);
	if (Mi == 0) return 0;
	uint64_t cur = 0;
	for (int i = 0; i < n; i++) cur |= (uint64_t) idx[i] << (4 * i);
	int q, p;
	qp_decode((int) (next() % qp_count(Mi, C - 1)), Mi, C - 1, q, p);
	uint32_t plane[16];
	window_planes_u8(d, n, plane);
	uint64_t ea, eb;
	const uint32_t a = window_item_u8<CLOG>(d, n, cur, q, p, size, bits_total, dim, ea);
	const uint32_t b = window_item_lut_u8<CLOG>(lut.data(), d, plane, n, cur, q, p, size, bits_total, dim, eb);
	return (a != b || ea != eb) ? 1 : 0;
}
// endpoint_floor_int (closed form) against endpoint_floor (the reference's bisection): every endpoint width, parity class and
// a dense set of values around every integer of -2 .. 258 (and NaN). Returns the number of mismatches.
extern "C" int hb_floor_check() {
	using namespace b200ic::amd7;
	int bad = 0;
	for (int bits = 4; bits <= 8; bits++)
		for (int use_par = 0; use_par <= 1; use_par++)
			for (int odd = 0; odd <= 1; odd++) {
				for (int i = -2 * 8; i <= 258 * 8; i++) {
					const double v = (double) i / 8.0;
					for (double dv : {0.0, 1e-9, -1e-9})
						if (endpoint_floor_int(v + dv, bits, use_par, odd) != endpoint_floor(v + dv, bits, use_par, odd)) bad++;
				}
				const double nan = std::nan("");
				if (endpoint_floor_int(nan, bits, use_par, odd) != endpoint_floor(nan, bits, use_par, odd)) bad++;
			}
	return bad;
}
extern "C" int hb_window_check(uint64_t seed, int trials) {
	using namespace b200ic::amd7;
	uint64_t rng = seed * 2 + 1;
	int bad = 0;
	for (int t = 0; t < trials; t++) {
		for (int mode = 0; mode < 8; mode++) {
			if (mode == 6) continue;
			const ModeInfo mi = mode_info(mode);
			if (mi.alpha == 2) {
				bad += hb_window_trial<2>(rng, 6, 6 * (mi.vector_bits / 3), 3);
				if (mode == 4) bad += hb_window_trial<3>(rng, 6, 6 * mi.scalar_bits, 3);
				else bad += hb_window_trial<2>(rng, 6, 6 * mi.scalar_bits, 3);
			} else {
				const ShakeParams sp = single_index_shake_params(mode);
				if (sp.clusters == 8) bad += hb_window_trial<3>(rng, sp.shake_size, sp.bits[3], sp.dim);
				else bad += hb_window_trial<2>(rng, sp.shake_size, sp.bits[3], sp.dim);
			}
		}
	}
	return bad;
}
extern "C" {
int hb_cube_lane_check(uint64_t seed, int trials) {
	using namespace b200ic::amd7;
	uint64_t rng = seed * 2 + 1;
	int bad = 0;
	for (int t = 0; t < trials; t++) {
		bad += hb_cube_lane_trial<3>(rng, 5, BCC);      // mode 0
		bad += hb_cube_lane_trial<3>(rng, 7, SAME_PAR); // mode 1
		bad += hb_cube_lane_trial<2>(rng, 5, CART);     // mode 2
		bad += hb_cube_lane_trial<2>(rng, 8, BCC);      // mode 3
		bad += hb_cube_lane_trial<2>(rng, 5, CART);     // mode 4 vector
		bad += hb_cube_lane_trial<3>(rng, 6, CART);     // mode 4 scalar (3-bit indices)
		bad += hb_cube_lane_trial<2>(rng, 7, CART);     // mode 5 vector
		bad += hb_cube_lane_trial<2>(rng, 8, CART);     // mode 5 scalar
	}
	return bad;
}
#endif
#ifdef HB_BC7RG
void hb_bc7rg_blocks(const uint32_t *px, uint64_t nblocks, int perceptual, int fast, uint8_t *out) {
	static b200ic::rg::OptimalEndpoint table[512];
	static bool inited = false;
	if (!inited) { b200ic::rg::build_mode1_single_colour_table(table); inited = true; }
	b200ic::rg::Params P;
	b200ic::rg::make_params(P, perceptual != 0, fast != 0, table);
	for (uint64_t b = 0; b < nblocks; b++) {
		uint64_t w[2];
		b200ic::rg::encode_block(px + b * 16, P, w);
		memcpy(out + b * 16, w, 16);
	}
}
#endif
}
