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
