/* TEST INFRASTRUCTURE ONLY -- format-specification block DECODERS (BC1, BC2, BC3, BC4, BC5, BC6H, BC7).
 *
 * The reference ships no decoder, yet the AMD BC7 / BC6H paths are gated on decoded-texel
 * PSNR, so these follow the D3D11 / Khronos BPTC + S3TC/RGTC specifications, not any reference file.
 * The BC7 interpolation weights are the same constants the reference encoders use
 * (src/amd_bc7_body.cpp:123-141, src/richgel999_bc7enc16.cpp:130-131).  The BC6H decoder restates the specification's
 * unquantise / interpolate / finish steps, which are also what the reference's encoder evaluates candidates with
 * (src/amd_hdr_encode.cpp:117-150 Unquantize, src/amd_bc6h_body.cpp:1039-1049 finish_unquantizeF16).
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

/* ---------------- BC2 / BC3: 8 bytes of alpha (4-bit explicit / BC4 block), then a colour block that is ALWAYS decoded
 * in 4-colour mode, whatever the order of its two 565 end points ---------------- */
static void bc23_colour(const uint8_t *b, uint8_t out[16][4]) {
	uint32_t c[2] = {(uint32_t) (b[0] | (b[1] << 8)), (uint32_t) (b[2] | (b[3] << 8))};
	uint8_t pal[4][3];
	for (int k = 0; k < 2; k++) {
		uint32_t r = (c[k] >> 11) & 31, g = (c[k] >> 5) & 63, bl = c[k] & 31;
		pal[k][0] = (uint8_t) ((r << 3) | (r >> 2));
		pal[k][1] = (uint8_t) ((g << 2) | (g >> 4));
		pal[k][2] = (uint8_t) ((bl << 3) | (bl >> 2));
	}
	for (int ch = 0; ch < 3; ch++) {
		pal[2][ch] = (uint8_t) ((2 * pal[0][ch] + pal[1][ch] + 1) / 3);
		pal[3][ch] = (uint8_t) ((pal[0][ch] + 2 * pal[1][ch] + 1) / 3);
	}
	uint32_t idx = b[4] | (b[5] << 8) | (b[6] << 16) | ((uint32_t) b[7] << 24);
	for (int i = 0; i < 16; i++) memcpy(out[i], pal[(idx >> (2 * i)) & 3], 3);
}
/* explicit_alpha != 0: BC2, else BC3 */
void bcdec_bc23(const uint8_t *blocks, uint32_t w, uint32_t h, int explicit_alpha, uint8_t *rgba) {
	uint32_t nbx = (w + 3) / 4, nby = (h + 3) / 4;
	for (uint32_t by = 0; by < nby; by++) for (uint32_t bx = 0; bx < nbx; bx++) {
		const uint8_t *b = blocks + ((size_t) by * nbx + bx) * 16;
		uint8_t px[16][4], a[16];
		bc23_colour(b + 8, px);
		if (explicit_alpha) for (int i = 0; i < 16; i++) a[i] = (uint8_t) (((b[i >> 1] >> (4 * (i & 1))) & 15) * 17);
		else bc4_block(b, a);
		for (int i = 0; i < 16; i++) px[i][3] = a[i];
		store_rgba(rgba, w, h, bx, by, px);
	}
}

/* ---------------- BC6H (D3D11 / Khronos BPTC float specification) ---------------- */
/* field order: rw gw bw rx gx bx ry gy by rz gz bz; every mode as the specification's "name[hi:lo]" runs, LSB first */
typedef struct { int code, code_bits, regions, transformed, epb, db[3]; const char *seq; } bc6_mode;
static const bc6_mode BC6_MODES[14] = {
	{0x00, 2, 2, 1, 10, {5, 5, 5}, "gy4 by4 bz4 rw0-9 gw0-9 bw0-9 rx0-4 gz4 gy0-3 gx0-4 bz0 gz0-3 bx0-4 bz1 by0-3 ry0-4 bz2 rz0-4 bz3"},
	{0x01, 2, 2, 1, 7, {6, 6, 6}, "gy5 gz4 gz5 rw0-6 bz0 bz1 by4 gw0-6 by5 bz2 gy4 bw0-6 bz3 bz5 bz4 rx0-5 gy0-3 gx0-5 gz0-3 bx0-5 by0-3 ry0-5 rz0-5"},
	{0x02, 5, 2, 1, 11, {5, 4, 4}, "rw0-9 gw0-9 bw0-9 rx0-4 rw10 gy0-3 gx0-3 gw10 bz0 gz0-3 bx0-3 bw10 bz1 by0-3 ry0-4 bz2 rz0-4 bz3"},
	{0x06, 5, 2, 1, 11, {4, 5, 4}, "rw0-9 gw0-9 bw0-9 rx0-3 rw10 gz4 gy0-3 gx0-4 gw10 gz0-3 bx0-3 bw10 bz1 by0-3 ry0-3 bz0 bz2 rz0-3 gy4 bz3"},
	{0x0a, 5, 2, 1, 11, {4, 4, 5}, "rw0-9 gw0-9 bw0-9 rx0-3 rw10 by4 gy0-3 gx0-3 gw10 bz0 gz0-3 bx0-4 bw10 by0-3 ry0-3 bz1 bz2 rz0-3 bz4 bz3"},
	{0x0e, 5, 2, 1, 9, {5, 5, 5}, "rw0-8 by4 gw0-8 gy4 bw0-8 bz4 rx0-4 gz4 gy0-3 gx0-4 bz0 gz0-3 bx0-4 bz1 by0-3 ry0-4 bz2 rz0-4 bz3"},
	{0x12, 5, 2, 1, 8, {6, 5, 5}, "rw0-7 gz4 by4 gw0-7 bz2 gy4 bw0-7 bz3 bz4 rx0-5 gy0-3 gx0-4 bz0 gz0-3 bx0-4 bz1 by0-3 ry0-5 rz0-5"},
	{0x16, 5, 2, 1, 8, {5, 6, 5}, "rw0-7 bz0 by4 gw0-7 gy5 gy4 bw0-7 gz5 bz4 rx0-4 gz4 gy0-3 gx0-5 gz0-3 bx0-4 bz1 by0-3 ry0-4 bz2 rz0-4 bz3"},
	{0x1a, 5, 2, 1, 8, {5, 5, 6}, "rw0-7 bz1 by4 gw0-7 by5 gy4 bw0-7 bz5 bz4 rx0-4 gz4 gy0-3 gx0-4 bz0 gz0-3 bx0-5 by0-3 ry0-4 bz2 rz0-4 bz3"},
	{0x1e, 5, 2, 0, 6, {6, 6, 6}, "rw0-5 gz4 bz0 bz1 by4 gw0-5 gy5 by5 bz2 gy4 bw0-5 gz5 bz3 bz5 bz4 rx0-5 gy0-3 gx0-5 gz0-3 bx0-5 by0-3 ry0-5 rz0-5"},
	{0x03, 5, 1, 0, 10, {10, 10, 10}, "rw0-9 gw0-9 bw0-9 rx0-9 gx0-9 bx0-9"},
	{0x07, 5, 1, 1, 11, {9, 9, 9}, "rw0-9 gw0-9 bw0-9 rx0-8 rw10 gx0-8 gw10 bx0-8 bw10"},
	{0x0b, 5, 1, 1, 12, {8, 8, 8}, "rw0-9 gw0-9 bw0-9 rx0-7 rw11 rw10 gx0-7 gw11 gw10 bx0-7 bw11 bw10"},
	{0x0f, 5, 1, 1, 16, {4, 4, 4}, "rw0-9 gw0-9 bw0-9 rx0-3 rw15 rw14 rw13 rw12 rw11 rw10 gx0-3 gw15 gw14 gw13 gw12 gw11 gw10 bx0-3 bw15 bw14 bw13 bw12 bw11 bw10"},
};
static int bc6_bit(const uint8_t *b, int pos) { return (b[pos >> 3] >> (pos & 7)) & 1; }
static int bc6_sext(int v, int bits) { return (v & (1 << (bits - 1))) ? v - (1 << bits) : v; }
static int bc6_unquantize(int comp, int epb, int is_signed) {
	if (!is_signed) {
		if (epb >= 15) return comp;
		if (comp == 0) return 0;
		if (comp == (1 << epb) - 1) return 0xFFFF;
		return ((comp << 15) + 0x4000) >> (epb - 1);
	}
	if (epb >= 16) return comp;
	int s = comp < 0, unq;
	if (s) comp = -comp;
	if (comp == 0) unq = 0;
	else if (comp >= (1 << (epb - 1)) - 1) unq = 0x7FFF;
	else unq = ((comp << 15) + 0x4000) >> (epb - 1);
	return s ? -unq : unq;
}
static uint16_t bc6_finish(int v, int is_signed) {
	if (!is_signed) return (uint16_t) ((v * 31) >> 6);
	if (v < 0) return (uint16_t) (0x8000 | (((-v) * 31) >> 5));
	return (uint16_t) ((v * 31) >> 5);
}
/* one block -> 16 x RGB half bit patterns; returns the mode number 1..14 or 0 for a reserved encoding (black) */
static int bc6h_block(const uint8_t *b, int is_signed, uint16_t out[16][3]) {
	int two = b[0] & 3, five = b[0] & 31, m = -1;
	for (int i = 0; i < 14; i++)
		if ((BC6_MODES[i].code_bits == 2 && BC6_MODES[i].code == two) || (BC6_MODES[i].code_bits == 5 && two >= 2 && BC6_MODES[i].code == five)) { m = i; break; }
	if (m < 0) { memset(out, 0, 16 * 3 * sizeof(uint16_t)); return 0; }
	const bc6_mode *M = &BC6_MODES[m];
	int f[12] = {0};
	int pos = M->code_bits;
	for (const char *p = M->seq; *p;) {
		while (*p == ' ') p++;
		if (!*p) break;
		int ch = p[0] == 'r' ? 0 : (p[0] == 'g' ? 1 : 2), e = p[1] - 'w';
		p += 2;
		int lo = 0, hi;
		while (*p >= '0' && *p <= '9') lo = lo * 10 + (*p++ - '0');
		hi = lo;
		if (*p == '-') { p++; hi = 0; while (*p >= '0' && *p <= '9') hi = hi * 10 + (*p++ - '0'); }
		for (int k = lo; k <= hi; k++) f[e * 3 + ch] |= bc6_bit(b, pos++) << k;
	}
	int part = 0;
	if (M->regions == 2) for (int k = 0; k < 5; k++) part |= bc6_bit(b, pos++) << k;
	int ep[4][3]; /* w x y z */
	const int mask = (1 << M->epb) - 1;
	for (int ch = 0; ch < 3; ch++) {
		int w = f[ch];
		if (is_signed) w = bc6_sext(w, M->epb);
		ep[0][ch] = w;
		for (int e = 1; e < 2 * M->regions; e++) {
			int v = f[e * 3 + ch];
			if (M->transformed) {
				v = (w + bc6_sext(v, M->db[ch])) & mask;
				if (is_signed) v = bc6_sext(v, M->epb);
			} else if (is_signed) v = bc6_sext(v, M->epb);
			ep[e][ch] = v;
		}
	}
	for (int e = 0; e < 2 * M->regions; e++) for (int ch = 0; ch < 3; ch++) ep[e][ch] = bc6_unquantize(ep[e][ch], M->epb, is_signed);
	const int ib = M->regions == 2 ? 3 : 4;
	const int anchor2 = M->regions == 2 ? kBc7Anchor2[part] : -1;
	for (int i = 0; i < 16; i++) {
		int nb = ib - ((i == 0 || i == anchor2) ? 1 : 0), idx = 0;
		for (int k = 0; k < nb; k++) idx |= bc6_bit(b, pos++) << k;
		int s = M->regions == 2 ? (kBc7Part2[part] >> i) & 1 : 0;
		int wgt = ib == 3 ? W3[idx] : W4[idx];
		for (int ch = 0; ch < 3; ch++) out[i][ch] = bc6_finish((ep[2 * s][ch] * (64 - wgt) + ep[2 * s + 1][ch] * wgt + 32) >> 6, is_signed);
	}
	return m + 1;
}
/* blocks -> tightly packed RGB half bit patterns (3 x uint16 per texel); modes_hist (may be NULL): 15 counters */
void bcdec_bc6h(const uint8_t *blocks, uint32_t w, uint32_t h, int is_signed, uint16_t *rgb, uint32_t *modes_hist) {
	uint32_t nbx = (w + 3) / 4, nby = (h + 3) / 4;
	for (uint32_t by = 0; by < nby; by++) for (uint32_t bx = 0; bx < nbx; bx++) {
		uint16_t px[16][3];
		int m = bc6h_block(blocks + ((size_t) by * nbx + bx) * 16, is_signed, px);
		if (modes_hist) modes_hist[m]++;
		for (int i = 0; i < 16; i++) {
			uint32_t x = bx * 4 + (i & 3), y = by * 4 + (i >> 2);
			if (x < w && y < h) memcpy(rgb + ((size_t) y * w + x) * 3, px[i], 6);
		}
	}
}
