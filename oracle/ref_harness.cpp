// TEST INFRASTRUCTURE ONLY (see oracle/README.md): thin driver around the UNMODIFIED reference
// sources, which oracle/Makefile compiles from /root/reference/src where they lie into
// oracle/_ref/libref_oracle.so. Nothing here is on the product path.
//
// The reference is single-threaded; this harness splits an image into tiles of whole 4x4 blocks (block-row chunks,
// cut into block-column segments when block-rows alone would leave threads idle) and calls the reference's own
// image-level API (Image_CompressAMDBC1 ... ,
// include/gfx_imagecompress/imagecompress.h:69-100) on each tile from a pool of std::threads after one
// Image_CompressInit(). Blocks are independent, so the tiled output is byte-identical
// to a serial call.
#include "al2o3_platform/platform.h"
#include "gfx_image/image.h"
#include "gfx_imagecompress/imagecompress.h"
#include "amd_bc6h_body.hpp"
#include <thread>
#include <vector>
#include <atomic>
#include <algorithm>

extern "C" {

struct ref_opts {
	int32_t bc1_use_alpha;        // Image_CompressBC1Options.UseAlpha
	int32_t bc1_alpha_threshold;  // Image_CompressBC1Options.AlphaThreshold (0..255)
	int32_t amd_3d_refinement;
	int32_t amd_adaptive_weights;
	int32_t amd_refinement_steps;
	int32_t amd_mode_mask;
	int32_t rg_perceptual;
	int32_t rg_fast;
	int32_t use_defaults;         // !=0: pass nullptr option structs (the reference's own defaults)
};

enum { REF_BC1 = 1, REF_BC2 = 2, REF_BC3 = 3, REF_BC4 = 4, REF_BC5 = 5, REF_BC6H = 6, REF_BC7 = 7, REF_BC7_RG = 8 };

static Image_ImageHeader const *encode_one(int codec, Image_ImageHeader const *src, ref_opts const *o) {
	Image_CompressAMDBackendOptions amd;
	Image_CompressBC1Options bc1;
	Image_CompressRichGel999BackendOptions rg;
	bool const def = (o == nullptr) || o->use_defaults;
	if (!def) {
		amd.b3DRefinement = o->amd_3d_refinement != 0;
		amd.AdaptiveColourWeights = o->amd_adaptive_weights != 0;
		amd.RefinementSteps = (uint8_t) o->amd_refinement_steps;
		amd.ModeMask = (uint8_t) o->amd_mode_mask;
		bc1.UseAlpha = o->bc1_use_alpha != 0;
		bc1.AlphaThreshold = (uint8_t) o->bc1_alpha_threshold;
		rg.perceptual = o->rg_perceptual != 0;
		rg.fast = o->rg_fast != 0;
	}
	switch (codec) {
	case REF_BC1: return Image_CompressAMDBC1(src, def ? nullptr : &amd, def ? nullptr : &bc1, nullptr, nullptr);
	case REF_BC2: return Image_CompressAMDBC2(src, def ? nullptr : &amd, nullptr, nullptr);
	case REF_BC3: return Image_CompressAMDBC3(src, def ? nullptr : &amd, nullptr, nullptr);
	case REF_BC4: return Image_CompressAMDBC4(src, nullptr, nullptr);
	case REF_BC5: return Image_CompressAMDBC5(src, nullptr, nullptr);
	case REF_BC6H: return Image_CompressAMDBC6H(src, def ? nullptr : &amd, nullptr, nullptr);
	case REF_BC7: return Image_CompressAMDBC7(src, def ? nullptr : &amd, nullptr, nullptr);
	case REF_BC7_RG: return Image_CompressRichGel999BC7(src, def ? nullptr : &rg, nullptr, nullptr);
	default: return nullptr;
	}
}

// Threads that encoded at least one work unit in the last ref_encode_rows call (bench.py reports it beside `cores`).
static std::atomic<int> g_threads_used{0};
int ref_last_threads_used() { return g_threads_used.load(); }

// Encode block-rows [by0, by1) of a tightly packed w x h image (one slice) in format `fmt`
// (a TinyImageFormat value of the compat shim). `dst` receives (by1-by0)*blocksX blocks.
// Work units are tiles of whole 4x4 blocks: chunks of block-rows, cut further into block-column segments when there
// are fewer block-rows than 8 x nthreads, so that EVERY thread has work even for a sample of a few block-rows (only the
// last segment of a row can end off a multiple of 4 -- at the true image edge -- so the replicate-edge gather of
// src/block_utils.cpp:19,22 sees the same texels as in a whole-image call). Returns 0 on success.
int ref_encode_rows(int codec, void const *pixels, uint32_t w, uint32_t h, int fmt, uint32_t by0, uint32_t by1,
										void *dst, int nthreads, ref_opts const *opts) {
	TinyImageFormat const tf = (TinyImageFormat) fmt;
	uint32_t const bpp = TinyImageFormat_BytesPerPixel(tf);
	if (bpp == 0 || w == 0 || h == 0) return -1;
	uint32_t const blocksY = (h + 3) / 4, blocksX = (w + 3) / 4;
	if (by1 > blocksY) by1 = blocksY;
	if (by0 >= by1) return 0;
	uint32_t const nrows = by1 - by0;
	if (nthreads < 1) nthreads = 1;
	uint32_t const blockBytes = (codec == REF_BC1 || codec == REF_BC4) ? 8 : 16;
	uint32_t const want_units = (uint32_t) nthreads * 8;
	uint32_t const row_chunk = std::max<uint32_t>(1, nrows / want_units);
	uint32_t const row_units = (nrows + row_chunk - 1) / row_chunk;
	uint32_t segs = row_units >= want_units ? 1 : (want_units + row_units - 1) / row_units;
	segs = std::min(segs, blocksX);
	uint32_t const seg_blocks = (blocksX + segs - 1) / segs;
	segs = (blocksX + seg_blocks - 1) / seg_blocks;
	uint32_t const units = row_units * segs;
	nthreads = (int) std::min<uint32_t>((uint32_t) nthreads, units);

	Image_CompressInit();
	std::atomic<int> failed{0}, used{0};
	std::atomic<uint32_t> next{0};
	auto worker = [&]() {
		bool worked = false;
		for (;;) {
			uint32_t const u = next.fetch_add(1);
			if (u >= units) break;
			worked = true;
			uint32_t const ru = u / segs, su = u - ru * segs;
			uint32_t const c0 = ru * row_chunk;                                  // first block-row of the unit, relative to by0
			uint32_t const r0 = by0 + c0, r1 = std::min(by1, r0 + row_chunk);
			uint32_t const y0 = r0 * 4, y1 = std::min(h, r1 * 4);
			uint32_t const bx0 = su * seg_blocks, bx1 = std::min(blocksX, bx0 + seg_blocks);
			uint32_t const x0 = bx0 * 4, x1 = std::min(w, bx1 * 4);
			Image_ImageHeader const *tile = Image_CreateNoClear(x1 - x0, y1 - y0, 1, 1, tf);
			if (!tile) { failed = 1; break; }
			for (uint32_t y = y0; y < y1; ++y)
				memcpy((uint8_t *) Image_RawDataPtr(tile) + (size_t) (y - y0) * (x1 - x0) * bpp,
							 (uint8_t const *) pixels + ((size_t) y * w + x0) * bpp, (size_t) (x1 - x0) * bpp);
			Image_ImageHeader const *out = encode_one(codec, tile, opts);
			if (!out) { failed = 1; Image_Destroy(tile); break; }
			for (uint32_t r = r0; r < r1; ++r)
				memcpy((uint8_t *) dst + ((size_t) (r - by0) * blocksX + bx0) * blockBytes,
							 (uint8_t const *) Image_RawDataPtr(out) + (size_t) (r - r0) * (bx1 - bx0) * blockBytes, (size_t) (bx1 - bx0) * blockBytes);
			Image_Destroy(out);
			Image_Destroy(tile);
		}
		if (worked) used.fetch_add(1);
	};
	if (nthreads == 1) worker();
	else {
		std::vector<std::thread> pool;
		for (int t = 0; t < nthreads; ++t) pool.emplace_back(worker);
		for (auto &t : pool) t.join();
	}
	g_threads_used = used.load();
	return failed.load();
}

int ref_encode(int codec, void const *pixels, uint32_t w, uint32_t h, int fmt, void *dst, int nthreads,
							 ref_opts const *opts) {
	return ref_encode_rows(codec, pixels, w, h, fmt, 0, (h + 3) / 4, dst, nthreads, opts);
}

// The reference exposes no C block API for BC6H; this wraps
// BC6HBlockEncoder::CompressBlock (src/amd_bc6h_body.hpp:303-314) with the constructor arguments
// Image_CompressAMDBC6H uses (src/amd_bc6h_compressor.cpp:28).
void ref_bc6h_block(float const in[16][4], int isSigned, uint8_t modeMask, void *out) {
	BC6HBlockEncoder encoder(1.0f, false, isSigned != 0, modeMask, 1.0f);
	float tmp[16][4];
	memcpy(tmp, in, sizeof(tmp));
	encoder.CompressBlock(tmp, (uint8_t *) out);
}

int ref_hw_threads() { return (int) std::thread::hardware_concurrency(); }

} // extern "C"
