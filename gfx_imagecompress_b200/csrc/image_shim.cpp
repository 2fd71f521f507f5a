// image_shim.cpp -- the reference's public `Image_Compress*` C API (include/gfx_imagecompress/imagecompress.h)
// re-pointed at the B200 engine. This is the host side that replaces reference src/imagecompress.cpp and the
// image loops of src/amd_bc*_compressor.cpp / src/richgel999_bc7enc16.cpp:21-71: it validates, allocates the
// destination with Image_CreateNoClear exactly as the reference does, and hands the texels to
// b200ic_encode_host (H2D -> sm_100a kernels -> D2H). No CPU encode path exists here.
#pragma GCC visibility push(default)
#include "gfx_imagecompress/imagecompress.h"
#pragma GCC visibility pop
#include "b200ic.h"
#include <stdio.h>
#include <string.h>

namespace {

struct ProgressAdaptor {
	Image_CompressProgressFunc fn;
	void *user;
};
int progress_thunk(void *p, float pct) {
	auto *a = static_cast<ProgressAdaptor *>(p);
	return a->fn(a->user, pct) ? 1 : 0;
}

int to_b200_format(TinyImageFormat f) {
	switch (f) {
	case TinyImageFormat_R8_UNORM: return B200IC_FMT_R8;
	case TinyImageFormat_R8G8_UNORM: return B200IC_FMT_RG8;
	case TinyImageFormat_R8G8B8_UNORM: return B200IC_FMT_RGB8;
	case TinyImageFormat_R8G8B8_SRGB: return B200IC_FMT_RGB8_SRGB;
	case TinyImageFormat_R8G8B8A8_UNORM: return B200IC_FMT_RGBA8;
	case TinyImageFormat_R8G8B8A8_SRGB: return B200IC_FMT_RGBA8_SRGB;
	case TinyImageFormat_R16G16B16A16_SFLOAT: return B200IC_FMT_RGBA16F;
#ifdef B200IC_COMPAT_HAS_R16G16B16A16_UFLOAT // tag private to compat/tiny_imageformat (unsigned-half sources, BASELINE config 3); absent upstream
	case TinyImageFormat_R16G16B16A16_UFLOAT: return B200IC_FMT_RGBA16UF;
#endif
	case TinyImageFormat_R32G32B32A32_SFLOAT: return B200IC_FMT_RGBA32F;
	default: return 0;
	}
}

Image_ImageHeader const *run(int codec, Image_ImageHeader const *src, TinyImageFormat dstFmt, b200ic_opts const &opts,
														 Image_CompressProgressFunc cb, void *user) {
	if (!src || src->depth > 1) return nullptr; // e.g. src/amd_bc1_compressor.cpp:16
	int const fmt = to_b200_format(src->format);
	if (fmt == 0) return nullptr;
	Image_ImageHeader const *dst = Image_CreateNoClear(src->width, src->height, 1, src->slices, dstFmt);
	if (!dst) return nullptr;
	ProgressAdaptor ad{cb, user};
	// Block-rows are sharded over the GPUs of the box (b200ic_set_devices / B200IC_DEVICES; default: every visible device)
	// when each of them gets at least 64 Ki blocks -- smaller images stay on the current device.
	uint64_t const blocks = (uint64_t) ((src->width + 3) / 4) * ((src->height + 3) / 4) * src->slices;
	int nd = b200ic_get_devices();
	while (nd > 1 && blocks / (uint64_t) nd < (1u << 16)) nd--;
	int const rc = b200ic_encode_host_sharded(codec, Image_RawDataPtr(src), fmt, src->width, src->height, 0, src->slices, &opts,
																						Image_RawDataPtr(dst), cb ? progress_thunk : nullptr, &ad, nd);
	if (rc != 0) { // error or cancelled
		Image_Destroy(dst);
		return nullptr;
	}
	return dst;
}

b200ic_opts amd_opts(Image_CompressAMDBackendOptions const *a) {
	b200ic_opts o;
	b200ic_default_opts(&o);
	if (a) {
		o.amd_3d_refinement = a->b3DRefinement;
		o.amd_adaptive_weights = a->AdaptiveColourWeights;
		o.amd_refinement_steps = a->RefinementSteps;
		o.amd_mode_mask = a->ModeMask;
	}
	return o;
}

} // namespace

extern "C" {

void Image_CompressInit(void) {}
void Image_CompressDeinit(void) {}

Image_ImageHeader const *Image_CompressAMDBC1(Image_ImageHeader const *src, Image_CompressAMDBackendOptions const *amdOptions,
																							Image_CompressBC1Options const *options, Image_CompressProgressFunc cb,
																							void *user) {
	if (!src) return nullptr;
	b200ic_opts o = amd_opts(amdOptions);
	bool const useAlpha = options ? options->UseAlpha : false;
	o.bc1_alpha_threshold = (options ? options->AlphaThreshold : 128) / 255.0f;
	bool const sRGB = TinyImageFormat_IsSRGB(src->format);
	// format mapping incl. the sRGB && !UseAlpha -> RGB_UNORM oddity (src/amd_bc1_compressor.cpp:33-35)
	TinyImageFormat const dstFmt = sRGB ? (useAlpha ? TinyImageFormat_DXBC1_RGBA_SRGB : TinyImageFormat_DXBC1_RGB_UNORM)
																			: (useAlpha ? TinyImageFormat_DXBC1_RGBA_UNORM : TinyImageFormat_DXBC1_RGB_UNORM);
	return run(B200IC_BC1, src, dstFmt, o, cb, user);
}

// BC2 / BC3 (src/amd_bc2_compressor.cpp:11-60, src/amd_bc3_compressor.cpp:11-60): alpha half bit-exact with the reference,
// colour half = the BC1 4-point fit (see bc23_colour_kernel for why the reference's own colour bytes are not a target)
Image_ImageHeader const *Image_CompressAMDBC2(Image_ImageHeader const *src, Image_CompressAMDBackendOptions const *amdOptions,
																							Image_CompressProgressFunc cb, void *user) {
	if (!src) return nullptr;
	b200ic_opts o = amd_opts(amdOptions);
	return run(B200IC_BC2, src, TinyImageFormat_IsSRGB(src->format) ? TinyImageFormat_DXBC2_SRGB : TinyImageFormat_DXBC2_UNORM, o, cb, user);
}
Image_ImageHeader const *Image_CompressAMDBC3(Image_ImageHeader const *src, Image_CompressAMDBackendOptions const *amdOptions,
																							Image_CompressProgressFunc cb, void *user) {
	if (!src) return nullptr;
	b200ic_opts o = amd_opts(amdOptions);
	return run(B200IC_BC3, src, TinyImageFormat_IsSRGB(src->format) ? TinyImageFormat_DXBC3_SRGB : TinyImageFormat_DXBC3_UNORM, o, cb, user);
}

Image_ImageHeader const *Image_CompressAMDBC4(Image_ImageHeader const *src, Image_CompressProgressFunc cb, void *user) {
	if (!src) return nullptr;
	b200ic_opts o = amd_opts(nullptr);
	TinyImageFormat const dstFmt = TinyImageFormat_IsSigned(src->format) ? TinyImageFormat_DXBC4_SNORM : TinyImageFormat_DXBC4_UNORM;
	return run(B200IC_BC4, src, dstFmt, o, cb, user);
}

Image_ImageHeader const *Image_CompressAMDBC5(Image_ImageHeader const *src, Image_CompressProgressFunc cb, void *user) {
	if (!src) return nullptr;
	b200ic_opts o = amd_opts(nullptr);
	TinyImageFormat const dstFmt = TinyImageFormat_IsSigned(src->format) ? TinyImageFormat_DXBC5_SNORM : TinyImageFormat_DXBC5_UNORM;
	return run(B200IC_BC5, src, dstFmt, o, cb, user);
}

Image_ImageHeader const *Image_CompressAMDBC6H(Image_ImageHeader const *src, Image_CompressAMDBackendOptions const *amdOptions,
																							 Image_CompressProgressFunc cb, void *user) {
	if (!src) return nullptr;
	b200ic_opts o = amd_opts(amdOptions);
	bool const isSigned = TinyImageFormat_IsSigned(src->format);
	o.bc6h_signed = isSigned;
	return run(B200IC_BC6H, src, isSigned ? TinyImageFormat_DXBC6H_SFLOAT : TinyImageFormat_DXBC6H_UFLOAT, o, cb, user);
}

Image_ImageHeader const *Image_CompressAMDBC7(Image_ImageHeader const *src, Image_CompressAMDBackendOptions const *amdOptions,
																							Image_CompressProgressFunc cb, void *user) {
	if (!src) return nullptr;
	b200ic_opts o = amd_opts(amdOptions);
	o.src_has_alpha = TinyImageFormat_ChannelCount(src->format) > 3;
	return run(B200IC_BC7_AMD, src, TinyImageFormat_IsSRGB(src->format) ? TinyImageFormat_DXBC7_SRGB : TinyImageFormat_DXBC7_UNORM,
						 o, cb, user);
}

Image_ImageHeader const *Image_CompressRichGel999BC7(Image_ImageHeader const *src,
																										 Image_CompressRichGel999BackendOptions const *richOptions,
																										 Image_CompressProgressFunc cb, void *user) {
	if (!src) return nullptr;
	b200ic_opts o = amd_opts(nullptr);
	o.rg_perceptual = richOptions ? richOptions->perceptual : true;
	o.rg_fast = richOptions ? richOptions->fast : false;
	return run(B200IC_BC7_RG, src, TinyImageFormat_IsSRGB(src->format) ? TinyImageFormat_DXBC7_SRGB : TinyImageFormat_DXBC7_UNORM,
						 o, cb, user);
}

// type -> codec dispatch; same table as reference src/imagecompress.cpp:20-50
Image_ImageHeader const *ImageCompress_Compress(Image_CompressType type, bool fast, Image_ImageHeader const *src) {
	switch (type) {
	case Image_CT_None: return src;
	case Image_CT_DXBC1: return Image_CompressAMDBC1(src, nullptr, nullptr, nullptr, nullptr);
	case Image_CT_DXBC2: return Image_CompressAMDBC2(src, nullptr, nullptr, nullptr);
	case Image_CT_DXBC3: return Image_CompressAMDBC3(src, nullptr, nullptr, nullptr);
	case Image_CT_DXBC4: return Image_CompressAMDBC4(src, nullptr, nullptr);
	case Image_CT_DXBC5: return Image_CompressAMDBC5(src, nullptr, nullptr);
	case Image_CT_DXBC6H: return Image_CompressAMDBC6H(src, nullptr, nullptr, nullptr);
	case Image_CT_DXBC7:
		return fast ? Image_CompressRichGel999BC7(src, nullptr, nullptr, nullptr) : Image_CompressAMDBC7(src, nullptr, nullptr, nullptr);
	default: return nullptr; // ETC / ASTC are enum slots only in the reference too (:40-46)
	}
}

// Same decision tree as reference src/imagecompress.cpp:52-116
Image_CompressType ImageCompress_PickCompressionType(Image_CompressPickFlags flags, Image_ImageHeader const *src) {
	if (TinyImageFormat_IsFloat(src->format)) {
		if (!(flags & Image_CPF_AllowDXBC6and7)) return Image_CT_None;
	} else if (!TinyImageFormat_IsNormalised(src->format)) {
		return Image_CT_None;
	}
	uint32_t const channels = TinyImageFormat_ChannelCount(src->format);
	if (channels == 1 && (flags & Image_CPF_AllowDXBC1to5)) return Image_CT_DXBC4;
	if (channels == 2 && (flags & Image_CPF_AllowDXBC1to5)) return Image_CT_DXBC5;
	if (flags & Image_CPF_AllowDXBC6and7) return Image_CT_DXBC7;
	if (flags & Image_CPF_AllowASTC) return Image_CT_ASTC;
	if (flags & Image_CPF_AllowDXBC1to5) return channels == 4 ? Image_CT_DXBC3 : Image_CT_DXBC1;
	return Image_CT_None;
}

// ---- block-level API: one pre-gathered block through the batched C-ABI ---------------------------------
// These functions return void, so a failure (no CUDA device, an option combination that is not built) cannot be
// reported to the caller: the block is filled with 0xFF bytes -- never left as uninitialised memory -- and the
// reason is printed to stderr once per call site and kept in b200ic_last_error().
namespace {
void block_call(int codec, void const *input, int fmt, b200ic_opts const &o, void *out, unsigned out_bytes, char const *who) {
	if (b200ic_encode_blocks(codec, input, fmt, 1, &o, out) != 0) {
		memset(out, 0xFF, out_bytes);
		fprintf(stderr, "gfx_imagecompress_b200: %s failed: %s\n", who, b200ic_last_error());
	}
}
void block_unsupported(void *out, unsigned out_bytes, char const *who, char const *why) {
	memset(out, 0xFF, out_bytes);
	fprintf(stderr, "gfx_imagecompress_b200: %s: %s\n", who, why);
}
} // namespace

void Image_CompressAMDAlphaSingleModeBlock(float const input[16], void *out) {
	b200ic_opts o;
	b200ic_default_opts(&o);
	block_call(B200IC_BC4, input, B200IC_FMT_BLOCKS_F32X1, o, out, 8, "Image_CompressAMDAlphaSingleModeBlock");
}

void Image_CompressAMDBC1Block(float const input[64], bool adaptiveColourWeight, bool b3DRefinement, uint8_t refinementSteps,
															 float alphaThreshold, void *out) {
	b200ic_opts o;
	b200ic_default_opts(&o);
	o.amd_adaptive_weights = adaptiveColourWeight;
	o.amd_3d_refinement = b3DRefinement;
	o.amd_refinement_steps = refinementSteps;
	o.bc1_alpha_threshold = alphaThreshold;
	block_call(B200IC_BC1, input, B200IC_FMT_BLOCKS_F32X4, o, out, 8, "Image_CompressAMDBC1Block");
}

void Image_CompressAMDMultiModeLDRBlock(float const input[64], uint8_t modeMask, bool srcHasAlpha, float quality,
																				bool colourRestrict, bool alphaRestrict, float performance, void *out) {
	// the image path hard-wires quality = performance = 1, both restricts true (src/amd_bc7_compressor.cpp:58-65);
	// only that configuration is built.
	if (quality != 1.0f || performance != 1.0f || !colourRestrict || !alphaRestrict) {
		block_unsupported(out, 16, "Image_CompressAMDMultiModeLDRBlock", "only quality = performance = 1 with colourRestrict = alphaRestrict = true is built");
		return;
	}
	b200ic_opts o;
	b200ic_default_opts(&o);
	o.amd_mode_mask = modeMask;
	o.src_has_alpha = srcHasAlpha;
	block_call(B200IC_BC7_AMD, input, B200IC_FMT_BLOCKS_F32X4, o, out, 16, "Image_CompressAMDMultiModeLDRBlock");
}

void Image_CompressRichGel999BC7enc16(uint32_t const input[16], bool fast, bool perceptual, void *out) {
	b200ic_opts o;
	b200ic_default_opts(&o);
	o.rg_fast = fast;
	o.rg_perceptual = perceptual;
	if (reinterpret_cast<uintptr_t>(input) % 16 == 0) {
		block_call(B200IC_BC7_RG, input, B200IC_FMT_BLOCKS_RGBA8, o, out, 16, "Image_CompressRichGel999BC7enc16");
	} else { // the engine reads RGBA8 blocks with 128-bit loads
		alignas(16) uint32_t tmp[16];
		memcpy(tmp, input, sizeof(tmp));
		block_call(B200IC_BC7_RG, tmp, B200IC_FMT_BLOCKS_RGBA8, o, out, 16, "Image_CompressRichGel999BC7enc16");
	}
}

// colour half of BC2 / BC3 from a stride-3 RGB block (src/amd_bcx_helpers.cpp:142-179)
void Image_CompressAMDRGBSingleModeBlock(float const *rgbBlock, bool adaptiveColourWeights, bool threeDRefinement, uint8_t refinementSteps,
																				 void *out) {
	b200ic_opts o;
	b200ic_default_opts(&o);
	o.amd_adaptive_weights = adaptiveColourWeights;
	o.amd_3d_refinement = threeDRefinement;
	o.amd_refinement_steps = refinementSteps;
	block_call(B200IC_BC23_COLOUR_HALF, rgbBlock, B200IC_FMT_BLOCKS_F32X3, o, out, 8, "Image_CompressAMDRGBSingleModeBlock");
}
// BC2's 4-bit explicit alpha (src/amd_bcx_helpers.cpp:107-123)
void Image_CompressAMDExplictAlphaSingleModeBlock(float const *input, void *out) {
	b200ic_opts o;
	b200ic_default_opts(&o);
	block_call(B200IC_BC2_ALPHA_HALF, input, B200IC_FMT_BLOCKS_F32X1, o, out, 8, "Image_CompressAMDExplictAlphaSingleModeBlock");
}

} // extern "C"
