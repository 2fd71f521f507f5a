/* TEST INFRASTRUCTURE ONLY -- format-specification block DECODERS (BC1, BC4, BC5, BC7; BC6H in bcdec6h.c).
 *
 * The reference ships no decoder, yet the AMD BC7 / BC6H paths are gated on decoded-texel
 * PSNR, so these follow the D3D11 / Khronos BPTC + S3TC/RGTC specifications, not any reference file.
 * The BC7 interpolation weights are the same constants the reference encoders use
 * (src/amd_bc7_body.cpp:123-141, src/richgel999_bc7enc16.cpp:130-131).
 */
#include <stdint.h>
#include <string.h>
#include "bc7_spec_tables.h"

/* ---------------- BC1 ---------------- */
static void bc1_block(const uint8_t *b, uint8_t out[16][4]) {
	uint32_t c0 = b[0] | (b[1] << 8), c1 = b[2] | (b[3] << 8);
	uint8_t pal[4][4];
	uint32_t c[2] = {c0, c1};
	for (int k = 0; k < 2; k++) {
		uint32_t r = (c[k] >> 11) & 31, g = (c[k] >> 5) & 63, bl = c[k] & 31;
		pal[k][0] = (uint8_t) ((r << 3) | (r >> 2));
		pal[k][1] = (uint8_t) ((g << 2) | (g >> 4));
		pal[k][2] = (uint8_t) ((bl << 3) | (bl >> 2));
		pal[k][3] = 255;
	}
	for (int ch = 0; ch < 3; ch++) {
		if (c0 > c1) {
			pal[2][ch] = (uint8_t) ((2 * pal[0][ch] + pal[1][ch] + 1) / 3);
			pal[3][ch] = (uint8_t) ((pal[0][ch] + 2 * pal[1][ch] + 1) / 3);
		} else {
			pal[2][ch] = (uint8_t) ((pal[0][ch] + pal[1][ch]) / 2);
			pal[3][ch] = 0;
		}
	}
	pal[2][3] = 255;
	pal[3][3] = (c0 > c1) ? 255 : 0;
	uint32_t idx = b[4] | (b[5] << 8) | (b[6] << 16) | ((uint32_t) b[7] << 24);
	for (int i = 0; i < 16; i++) memcpy(out[i], pal[(idx >> (2 * i)) & 3], 4);
}

/* ---------------- BC4 (one channel) ---------------- */
static void bc4_block(const uint8_t *b, uint8_t out[16]) {
	uint32_t a0 = b[0], a1 = b[1];
	uint8_t pal[8];
	pal[0] = (uint8_t) a0;
	pal[1] = (uint8_t) a1;
	if (a0 > a1) {
		for (int i = 1; i < 7; i++) pal[i + 1] = (uint8_t) (((7 - i) * a0 + i * a1 + 3) / 7);
	} else {
		for (int i = 1; i < 5; i++) pal[i + 1] = (uint8_t) (((5 - i) * a0 + i * a1 + 2) / 5);
		pal[6] = 0;
		pal[7] = 255;
	}
	uint64_t bits = 0;
	for (int i = 0; i < 6; i++) bits |= (uint64_t) b[2 + i] << (8 * i);
	for (int i = 0; i < 16; i++) out[i] = pal[(bits >> (3 * i)) & 7];
}

/* ---------------- BC7 ---------------- */
typedef struct { const uint8_t *p; int pos; } bitrd;
static uint32_t rd(bitrd *r, int n) {
	uint32_t v = 0;
	for (int i = 0; i < n; i++, r->pos++) v |= (uint32_t) ((r->p[r->pos >> 3] >> (r->pos & 7)) & 1) << i;
	return v;
}
static const uint8_t W2[4] = {0, 21, 43, 64};
static const uint8_t W3[8] = {0, 9, 18, 27, 37, 46, 55, 64};
static const uint8_t W4[16] = {0, 4, 9, 13, 17, 21, 26, 30, 34, 38, 43, 47, 51, 55, 60, 64};
static const uint8_t *weights(int bits) { return bits == 2 ? W2 : (bits == 3 ? W3 : W4); }
static uint32_t lerp7(uint32_t a, uint32_t b, uint32_t w) { return ((64 - w) * a + w * b + 32) >> 6; }

/*                         NS PB RB ISB CB AB EPB SPB IB IB2 */
static const uint8_t MODES[8][10] = {
	{3, 4, 0, 0, 4, 0, 1, 0, 3, 0}, {2, 6, 0, 0, 6, 0, 0, 1, 3, 0}, {3, 6, 0, 0, 5, 0, 0, 0, 2, 0}, {2, 6, 0, 0, 7, 0, 1, 0, 2, 0},
	{1, 0, 2, 1, 5, 6, 0, 0, 2, 3}, {1, 0, 2, 0, 7, 8, 0, 0, 2, 2}, {1, 0, 0, 0, 7, 7, 1, 0, 4, 0}, {2, 6, 0, 0, 5, 5, 1, 0, 2, 0}};

/* returns the mode (0..7) or -1 for a reserved block */
static int bc7_block(const uint8_t *b, uint8_t out[16][4], int *partition_out) {
	int mode = 0;
	while (mode < 8 && !((b[0] >> mode) & 1)) mode++;
	if (mode >= 8) { memset(out, 0, 64); return -1; }
	const uint8_t *M = MODES[mode];
	const int ns = M[0], cb = M[4], ab = M[5], ib = M[8], ib2 = M[9];
	bitrd r = {b, mode + 1};
	const uint32_t part = rd(&r, M[1]), rot = rd(&r, M[2]), isb = rd(&r, M[3]);
	if (partition_out) *partition_out = (int) part;
	uint32_t ep[6][4];
	for (int ch = 0; ch < 3; ch++) for (int e = 0; e < 2 * ns; e++) ep[e][ch] = rd(&r, cb);
	for (int e = 0; e < 2 * ns; e++) ep[e][3] = ab ? rd(&r, ab) : 255;
	int cbits = cb, abits = ab;
	if (M[6]) {
		for (int e = 0; e < 2 * ns; e++) {
			uint32_t p = rd(&r, 1);
			for (int ch = 0; ch < (ab ? 4 : 3); ch++) ep[e][ch] = (ep[e][ch] << 1) | p;
		}
		cbits++; if (ab) abits++;
	} else if (M[7]) {
		for (int s = 0; s < ns; s++) {
			uint32_t p = rd(&r, 1);
			for (int e = 2 * s; e < 2 * s + 2; e++) for (int ch = 0; ch < 3; ch++) ep[e][ch] = (ep[e][ch] << 1) | p;
		}
		cbits++;
	}
	for (int e = 0; e < 2 * ns; e++) {
		for (int ch = 0; ch < 3; ch++) { uint32_t v = ep[e][ch] << (8 - cbits); ep[e][ch] = v | (v >> cbits); }
		if (ab) { uint32_t v = ep[e][3] << (8 - abits); ep[e][3] = v | (v >> abits); }
	}
	int subset[16], anchor[3] = {0, -1, -1};
	for (int i = 0; i < 16; i++)
		subset[i] = ns == 1 ? 0 : (ns == 2 ? (kBc7Part2[part] >> i) & 1 : (kBc7Part3[part] >> (2 * i)) & 3);
	if (ns == 2) anchor[1] = kBc7Anchor2[part];
	if (ns == 3) { anchor[1] = kBc7Anchor3a[part]; anchor[2] = kBc7Anchor3b[part]; }
	uint32_t i1[16], i2[16];
	for (int i = 0; i < 16; i++) {
		int is_anchor = (i == anchor[0]) || (i == anchor[1]) || (i == anchor[2]);
		i1[i] = rd(&r, ib - is_anchor);
	}
	for (int i = 0; i < 16; i++) i2[i] = ib2 ? rd(&r, ib2 - (i == 0)) : 0;
	for (int i = 0; i < 16; i++) {
		const uint32_t *e0 = ep[2 * subset[i]], *e1 = ep[2 * subset[i] + 1];
		uint32_t ci = i1[i], ai = i1[i];
		int cib = ib, aib = ib;
		if (ib2) {
			if (isb) { ci = i2[i]; cib = ib2; ai = i1[i]; aib = ib; }
			else { ci = i1[i]; cib = ib; ai = i2[i]; aib = ib2; }
		}
		uint32_t px[4];
		for (int ch = 0; ch < 3; ch++) px[ch] = lerp7(e0[ch], e1[ch], weights(cib)[ci]);
		px[3] = ab ? lerp7(e0[3], e1[3], weights(aib)[ai]) : 255;
		if (rot) { uint32_t t = px[3]; px[3] = px[rot - 1]; px[rot - 1] = t; }
		for (int ch = 0; ch < 4; ch++) out[i][ch] = (uint8_t) px[ch];
	}
	return mode;
}

/* ---------------- image-level drivers: blocks (row-major) -> tightly packed texels ---------------- */
static void store_rgba(uint8_t *img, uint32_t w, uint32_t h, uint32_t bx, uint32_t by, uint8_t px[16][4]) {
	for (int i = 0; i < 16; i++) {
		uint32_t x = bx * 4 + (i & 3), y = by * 4 + (i >> 2);
		if (x < w && y < h) memcpy(img + ((size_t) y * w + x) * 4, px[i], 4);
	}
}
void bcdec_bc1(const uint8_t *blocks, uint32_t w, uint32_t h, uint8_t *rgba) {
	uint32_t nbx = (w + 3) / 4, nby = (h + 3) / 4;
	for (uint32_t by = 0; by < nby; by++) for (uint32_t bx = 0; bx < nbx; bx++) {
		uint8_t px[16][4];
		bc1_block(blocks + ((size_t) by * nbx + bx) * 8, px);
		store_rgba(rgba, w, h, bx, by, px);
	}
}
/* modes_hist (may be NULL): 9 counters, index 8 = reserved blocks */
void bcdec_bc7(const uint8_t *blocks, uint32_t w, uint32_t h, uint8_t *rgba, uint32_t *modes_hist) {
	uint32_t nbx = (w + 3) / 4, nby = (h + 3) / 4;
	for (uint32_t by = 0; by < nby; by++) for (uint32_t bx = 0; bx < nbx; bx++) {
		uint8_t px[16][4];
		int m = bc7_block(blocks + ((size_t) by * nbx + bx) * 16, px, 0);
		if (modes_hist) modes_hist[m < 0 ? 8 : m]++;
		store_rgba(rgba, w, h, bx, by, px);
	}
}
/* BC4 -> R8 (nch = 1), BC5 -> RG8 (nch = 2) */
void bcdec_bc45(const uint8_t *blocks, uint32_t w, uint32_t h, int nch, uint8_t *out) {
	uint32_t nbx = (w + 3) / 4, nby = (h + 3) / 4;
	for (uint32_t by = 0; by < nby; by++) for (uint32_t bx = 0; bx < nbx; bx++) for (int c = 0; c < nch; c++) {
		uint8_t px[16];
		bc4_block(blocks + (((size_t) by * nbx + bx) * nch + c) * 8, px);
		for (int i = 0; i < 16; i++) {
			uint32_t x = bx * 4 + (i & 3), y = by * 4 + (i >> 2);
			if (x < w && y < h) out[((size_t) y * w + x) * nch + c] = px[i];
		}
	}
}
