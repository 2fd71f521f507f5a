// TEST/DEBUG AID ONLY -- compiles the host-portable per-block encoder cores (csrc/*_core.cuh) with g++ so that
// their logic can be checked against the oracle in the CPU-only container. Never linked into the product library
// and never used by the package: the product path is the CUDA kernels.
#include <stdint.h>
#include <string.h>
#ifdef HB_BC7RG
#include "bc7rg_core.cuh"
#endif

extern "C" {
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
