// Compatibility shim: TinyImageFormat_EncodeLogicalPixelsF for R8G8B8A8_UNORM only
// (single call site: reference src/richgel999_bc7enc16.cpp:52-55).
// Semantics (ours): clamp to [0,1], (uint8)(v*255.0f + 0.5f), byte order R,G,B,A.
#pragma once
#include "tiny_imageformat/tinyimageformat_base.h"
#include "tiny_imageformat/tinyimageformat_query.h"

typedef struct TinyImageFormat_EncodeOutput {
	union { void *pixel; void *pixelPlane0; };
	void *pixelPlane1;
} TinyImageFormat_EncodeOutput;

#ifdef __cplusplus
extern "C" {
#endif
static inline uint8_t TinyImageFormat_Shim_F2U8(float v) {
	if (!(v > 0.0f)) v = 0.0f;
	if (v > 1.0f) v = 1.0f;
	return (uint8_t) (v * 255.0f + 0.5f);
}
static inline bool TinyImageFormat_EncodeLogicalPixelsF(TinyImageFormat fmt, float const *in, uint32_t width,
																												TinyImageFormat_EncodeOutput *out) {
	if (fmt != TinyImageFormat_R8G8B8A8_UNORM && fmt != TinyImageFormat_R8G8B8A8_SRGB) return false;
	uint8_t *o = (uint8_t *) out->pixel;
	for (uint32_t i = 0; i < width * 4; ++i) o[i] = TinyImageFormat_Shim_F2U8(in[i]);
	return true;
}
#ifdef __cplusplus
}
#endif
