// container.cu -- the steps either side of the encode path: box-filtered mip generation on the device
// (what a caller of the reference does on the CPU before compressing every level as its own image) and a DDS writer for
// the compressed chain (the reference's tests write DDS through the external gfx_imageio, tests/test_imagecompress.cpp:9-12).
#include "kernels.h"
#include <cstdio>
#include <cstring>

namespace b200ic {

namespace {

// One thread per destination texel: 2x2 box filter of an RGBA8 level, odd sizes drop the last row / column like
// floor(size / 2) mip chains do; a dimension that is already 1 is kept (2x1 / 1x2 averages).  Rounding: floor(x + 0.5).
__global__ void __launch_bounds__(256) box_mip_rgba8_kernel(const uint8_t *src, uint32_t sw, uint32_t sh, uint64_t spitch, uint8_t *dst, uint32_t dw,
																														uint32_t dh, uint64_t dpitch) {
	const uint32_t x = blockIdx.x * 16 + (threadIdx.x & 15), y = blockIdx.y * 16 + (threadIdx.x >> 4);
	if (x >= dw || y >= dh) return;
	const uint32_t x0 = sw > 1 ? 2 * x : 0, x1 = sw > 1 ? 2 * x + 1 : 0, y0 = sh > 1 ? 2 * y : 0, y1 = sh > 1 ? 2 * y + 1 : 0;
	const uint32_t a = *reinterpret_cast<const uint32_t *>(src + y0 * spitch + 4ull * x0), b = *reinterpret_cast<const uint32_t *>(src + y0 * spitch + 4ull * x1),
								 c = *reinterpret_cast<const uint32_t *>(src + y1 * spitch + 4ull * x0), d = *reinterpret_cast<const uint32_t *>(src + y1 * spitch + 4ull * x1);
	uint32_t out = 0;
#pragma unroll
	for (int k = 0; k < 4; k++) {
		const uint32_t s = ((a >> (8 * k)) & 255u) + ((b >> (8 * k)) & 255u) + ((c >> (8 * k)) & 255u) + ((d >> (8 * k)) & 255u);
		// vertical then horizontal average of exact halves == the sum / 4; floor(s / 4 + 0.5) = (s + 2) >> 2
		out |= ((s + 2u) >> 2) << (8 * k);
	}
	*reinterpret_cast<uint32_t *>(dst + y * dpitch + 4ull * x) = out;
}

} // namespace

cudaError_t launch_box_mip_rgba8(const void *src, uint32_t sw, uint32_t sh, uint64_t spitch, void *dst, uint64_t dpitch, cudaStream_t stream) {
	const uint32_t dw = sw > 1 ? sw / 2 : 1, dh = sh > 1 ? sh / 2 : 1;
	const dim3 grid((dw + 15) / 16, (dh + 15) / 16);
	box_mip_rgba8_kernel<<<grid, 256, 0, stream>>>(static_cast<const uint8_t *>(src), sw, sh, spitch ? spitch : 4ull * sw, static_cast<uint8_t *>(dst), dw, dh,
																								 dpitch ? dpitch : 4ull * dw);
	return cudaGetLastError();
}

// ---- DDS (DX10 header) ---------------------------------------------------------------------------------------------------
static uint32_t dxgi_format(int codec, int srgb, int is_signed) {
	switch (codec) {
	case B200IC_BC1: return srgb ? 72 : 71;   // BC1_UNORM(_SRGB)
	case B200IC_BC2: return srgb ? 75 : 74;
	case B200IC_BC3: return srgb ? 78 : 77;
	case B200IC_BC4: return is_signed ? 81 : 80;
	case B200IC_BC5: return is_signed ? 84 : 83;
	case B200IC_BC6H: return is_signed ? 96 : 95; // BC6H_SF16 / UF16
	case B200IC_BC7_AMD:
	case B200IC_BC7_RG: return srgb ? 99 : 98;
	default: return 0;
	}
}

int write_dds(const char *path, int codec, int srgb, int is_signed, uint32_t width, uint32_t height, uint32_t levels, const void *const *level_blocks) {
	const uint32_t fmt = dxgi_format(codec, srgb, is_signed);
	if (!path || !level_blocks || fmt == 0 || width == 0 || height == 0 || levels == 0) return -1;
	const uint32_t bb = (codec == B200IC_BC1 || codec == B200IC_BC4) ? 8 : 16;
	FILE *f = fopen(path, "wb");
	if (!f) return -2;
	uint32_t h[32] = {};
	h[0] = 0x20534444u; // "DDS "
	h[1] = 124;
	h[2] = 0x1 | 0x2 | 0x4 | 0x1000 | 0x80000 | (levels > 1 ? 0x20000 : 0); // CAPS | HEIGHT | WIDTH | PIXELFORMAT | LINEARSIZE | MIPMAPCOUNT
	h[3] = height;
	h[4] = width;
	h[5] = ((width + 3) / 4) * ((height + 3) / 4) * bb; // linear size of the top level
	h[7] = levels;
	h[19] = 32;          // DDS_PIXELFORMAT.size
	h[20] = 0x4;         // DDPF_FOURCC
	h[21] = 0x30315844u; // "DX10"
	h[27] = 0x1000 | (levels > 1 ? 0x400008 : 0); // TEXTURE (| MIPMAP | COMPLEX)
	const uint32_t dx10[5] = {fmt, 3 /* TEXTURE2D */, 0, 1, 0};
	bool ok = fwrite(h, 4, 32, f) == 32 && fwrite(dx10, 4, 5, f) == 5;
	uint32_t w = width, hh = height;
	for (uint32_t l = 0; l < levels && ok; l++) {
		const size_t n = (size_t) ((w + 3) / 4) * ((hh + 3) / 4) * bb;
		ok = level_blocks[l] && fwrite(level_blocks[l], 1, n, f) == n;
		w = w > 1 ? w / 2 : 1;
		hh = hh > 1 ? hh / 2 : 1;
	}
	ok = (fclose(f) == 0) && ok;
	return ok ? 0 : -3;
}

} // namespace b200ic
