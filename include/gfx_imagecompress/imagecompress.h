/* gfx_imagecompress/imagecompress.h -- drop-in public C API of the B200 BCn engine.
 *
 * Same symbols, enums, option structs, argument meaning, ownership and error behaviour as the reference's
 * include/gfx_imagecompress/imagecompress.h (enums :7-33, option structs :35-50, image API :57-100,
 * block API :111-142), so that an application links this library instead of the reference's static
 * library without source changes.  Every entry point runs hand-written sm_100a kernels through the
 * C-ABI in include/b200ic.h; none falls back to the CPU (a missing / unusable GPU yields NULL from the
 * image functions; block functions leave `out` untouched).
 *
 * Behaviour kept from the reference on purpose (see DESIGN.md "Reference quirks"):
 *   - Image_CT_None returns `src` itself; ETC/ASTC types return NULL; depth > 1 returns NULL
 *   - BC1 drops texels with alpha < AlphaThreshold/255 from the fit even when UseAlpha is false
 *   - BC4 encodes channel 1 (green); BC5 channels 0 and 1
 *   - ImageCompress_Compress(DXBC7, fast=true) selects bc7enc16 at its default (perceptual, uber 4) options
 *   - a progress callback returning true cancels: NULL is returned (unlike the reference we free the
 *     partially written image instead of leaking it)
 */
#pragma once

#include "gfx_image/image.h"

#ifdef __cplusplus
extern "C" {
#endif

/* return true to cancel */
typedef bool (*Image_CompressProgressFunc)(void *user, float percentage);

typedef enum Image_CompressType {
	Image_CT_None = 0,
	Image_CT_DXBC1, Image_CT_DXBC2, Image_CT_DXBC3, Image_CT_DXBC4, Image_CT_DXBC5, Image_CT_DXBC6H, Image_CT_DXBC7,
	Image_CT_ETC_RGB, Image_CT_ETC2_RGB, Image_CT_ETC_RGBA_Explicit, Image_CT_ETC_RGBA_Interpolated,
	Image_CT_ASTC,
	Image_CT_MAX
} Image_CompressType;

typedef enum Image_CompressPickFlags {
	Image_CPF_AllowDXBC1to5 = 0x1,
	Image_CPF_AllowASTC = 0x2,
	Image_CPF_AllowETC = 0x8,
	Image_CPF_AllowDXBC6and7 = 0x10
} Image_CompressPickFlags;

typedef struct Image_CompressBC1Options {
	bool UseAlpha;          /* default false */
	uint8_t AlphaThreshold; /* default 128   */
} Image_CompressBC1Options;

typedef struct Image_CompressAMDBackendOptions {
	bool b3DRefinement;         /* default false; true is unsupported by the B200 engine (returns NULL) */
	bool AdaptiveColourWeights; /* default false; true is unsupported (reads uninitialised memory in the reference) */
	uint8_t RefinementSteps;    /* default 1 */
	uint8_t ModeMask;           /* default 0xFF; BC6H / BC7 */
} Image_CompressAMDBackendOptions;

typedef struct Image_CompressRichGel99BackendOptions {
	bool perceptual; /* default true  */
	bool fast;       /* default false */
} Image_CompressRichGel999BackendOptions;

/* Ref-counted global set-up. The B200 engine uploads its constant tables on first use; these remain for
 * source compatibility with callers of the block-level API. */
void Image_CompressInit(void);
void Image_CompressDeinit(void);

/* Generic entry: picks the encoder from `type` (and `fast` for BC7). NULL option structs = defaults. */
Image_ImageHeader const *ImageCompress_Compress(Image_CompressType type, bool fast, Image_ImageHeader const *src);
Image_CompressType ImageCompress_PickCompressionType(Image_CompressPickFlags flags, Image_ImageHeader const *src);

/* Per-codec image entry points; option pointers may be NULL. The result is a new image owned by the caller. */
Image_ImageHeader const *Image_CompressAMDBC1(Image_ImageHeader const *src, Image_CompressAMDBackendOptions const *amdOptions,
																							Image_CompressBC1Options const *options,
																							Image_CompressProgressFunc progressCallback, void *userCallbackData);
Image_ImageHeader const *Image_CompressAMDBC2(Image_ImageHeader const *src, Image_CompressAMDBackendOptions const *amdOptions,
																							Image_CompressProgressFunc progressCallback, void *userCallbackData);
Image_ImageHeader const *Image_CompressAMDBC3(Image_ImageHeader const *src, Image_CompressAMDBackendOptions const *amdOptions,
																							Image_CompressProgressFunc progressCallback, void *userCallbackData);
Image_ImageHeader const *Image_CompressAMDBC4(Image_ImageHeader const *src, Image_CompressProgressFunc progressCallback,
																							void *userCallbackData);
Image_ImageHeader const *Image_CompressAMDBC5(Image_ImageHeader const *src, Image_CompressProgressFunc progressCallback,
																							void *userCallbackData);
Image_ImageHeader const *Image_CompressAMDBC6H(Image_ImageHeader const *src, Image_CompressAMDBackendOptions const *amdOptions,
																							 Image_CompressProgressFunc progressCallback, void *userCallbackData);
Image_ImageHeader const *Image_CompressAMDBC7(Image_ImageHeader const *src, Image_CompressAMDBackendOptions const *amdOptions,
																							Image_CompressProgressFunc progressCallback, void *userCallbackData);
Image_ImageHeader const *Image_CompressRichGel999BC7(Image_ImageHeader const *src,
																										 Image_CompressRichGel999BackendOptions const *richOptions,
																										 Image_CompressProgressFunc progressCallback, void *userCallbackData);

/* Block-level API. Inputs are normalised floats (0..1) in texel order, row-major within the 4x4 block. */
void Image_CompressAMDRGBSingleModeBlock(float const input[4 * 4 * 3], bool adaptiveColourWeights, bool b3DRefinement,
																				 uint8_t refinementSteps, void *out);             /* -> 8 B colour block   */
void Image_CompressAMDAlphaSingleModeBlock(float const input[4 * 4], void *out);          /* -> 8 B BC4-style block */
void Image_CompressAMDExplictAlphaSingleModeBlock(float const input[4 * 4], void *out);   /* -> 8 B 4-bit alpha     */
void Image_CompressAMDBC1Block(float const input[4 * 4 * 4], bool adaptiveColourWeight, bool b3DRefinement,
															 uint8_t refinementSteps, float alphaThreshold, void *out); /* -> 8 B BC1 block       */
void Image_CompressAMDMultiModeLDRBlock(float const input[4 * 4 * 4], uint8_t modeMask, bool srcHasAlpha, float quality,
																				bool colourRestrict, bool alphaRestrict, float performance, void *out); /* -> 16 B BC7 */
void Image_CompressRichGel999BC7enc16(uint32_t const input[4 * 4], bool fast, bool perceptual, void *out); /* -> 16 B BC7 */

#ifdef __cplusplus
}
#endif
