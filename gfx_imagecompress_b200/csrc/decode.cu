// decode.cu -- block DECODERS for every codec of the engine ("GPU decoders exposed publicly").
// The reference has no decoder; these follow the D3D11 / Khronos S3TC, RGTC and BPTC format specifications (the BC7
// interpolation weights are the constants the reference encoders use too, src/amd_bc7_body.cpp:123-141).
// One thread per 4x4 block: 8 / 16 bytes in, 16 texels out (clipped at the image edge), blocks row-major per slice.
#define B7T_QUAL __constant__ const
#include "kernels.h"
#include "bc7_tables.h"

namespace b200ic {

namespace {

struct DecParams {
	const uint8_t *blocks;
	uint8_t *dst;
	uint64_t row_pitch, slice_pitch, n_blocks;
	uint32_t width, height, blocks_x, blocks_y;
	int codec, is_signed;
};

__device__ __forceinline__ uint32_t exp565(uint32_t c, int ch) { // ch 0 r, 1 g, 2 b
	if (ch == 0) { const uint32_t r = (c >> 11) & 31u; return (r << 3) | (r >> 2); }
	if (ch == 1) { const uint32_t g = (c >> 5) & 63u; return (g << 2) | (g >> 4); }
	const uint32_t b = c & 31u;
	return (b << 3) | (b >> 2);
}
// colour block; four_only: BC2 / BC3 always decode in 4-colour mode
__device__ void colour_block(const uint2 w, bool four_only, uint32_t out[16]) {
	const uint32_t c0 = w.x & 0xffffu, c1 = w.x >> 16;
	uint32_t pal[4];
	const bool four = four_only || c0 > c1;
	uint32_t p0 = 0, p1 = 0, p2 = 0, p3 = 0;
#pragma unroll
	for (int ch = 0; ch < 3; ch++) {
		const uint32_t a = exp565(c0, ch), b = exp565(c1, ch);
		p0 |= a << (8 * ch);
		p1 |= b << (8 * ch);
		p2 |= (four ? (2 * a + b + 1) / 3 : (a + b) / 2) << (8 * ch);
		p3 |= (four ? (a + 2 * b + 1) / 3 : 0u) << (8 * ch);
	}
	pal[0] = p0 | 0xff000000u;
	pal[1] = p1 | 0xff000000u;
	pal[2] = p2 | 0xff000000u;
	pal[3] = four ? (p3 | 0xff000000u) : 0u;
#pragma unroll
	for (int i = 0; i < 16; i++) out[i] = pal[(w.y >> (2 * i)) & 3u];
}
__device__ void alpha_block(const uint2 w, uint32_t out[16]) { // BC4 / BC3 alpha: 8 values, 3-bit indices
	const uint32_t a0 = w.x & 255u, a1 = (w.x >> 8) & 255u;
	uint32_t pal[8];
	pal[0] = a0;
	pal[1] = a1;
	if (a0 > a1) {
#pragma unroll
		for (int i = 1; i < 7; i++) pal[i + 1] = ((7 - i) * a0 + i * a1 + 3) / 7;
	} else {
#pragma unroll
		for (int i = 1; i < 5; i++) pal[i + 1] = ((5 - i) * a0 + i * a1 + 2) / 5;
		pal[6] = 0;
		pal[7] = 255;
	}
	const uint64_t bits = ((uint64_t) w.y << 16) | (w.x >> 16);
#pragma unroll
	for (int i = 0; i < 16; i++) out[i] = pal[(bits >> (3 * i)) & 7u];
}

struct BitReader {
	uint64_t lo, hi;
	int pos;
	__device__ uint32_t get(int n) {
		uint32_t v = 0;
		for (int i = 0; i < n; i++, pos++) v |= (uint32_t) (((pos < 64 ? lo >> pos : hi >> (pos - 64)) & 1ull)) << i;
		return v;
	}
	__device__ uint32_t bit(int p) const { return (uint32_t) ((p < 64 ? lo >> p : hi >> (p - 64)) & 1ull); }
};
__device__ __forceinline__ uint32_t bptc_weight(int bits, uint32_t i) {
	if (bits == 2) return (0x402b1500u >> (8 * i)) & 255u;                       // 0 21 43 64
	if (bits == 3) return i < 4 ? (0x1b120900u >> (8 * i)) & 255u : (0x40372e25u >> (8 * (i - 4))) & 255u; // 0 9 18 27 37 46 55 64
	const uint32_t w4[4] = {0x0d090400u, 0x1e1a1511u, 0x2f2b2622u, 0x403c3733u}; // 0 4 9 13 17 21 26 30 34 38 43 47 51 55 60 64
	return (w4[i >> 2] >> (8 * (i & 3))) & 255u;
}
//                                    NS PB RB ISB CB AB EPB SPB IB IB2
__constant__ uint8_t kBc7Modes[8][10] = {{3, 4, 0, 0, 4, 0, 1, 0, 3, 0}, {2, 6, 0, 0, 6, 0, 0, 1, 3, 0}, {3, 6, 0, 0, 5, 0, 0, 0, 2, 0}, {2, 6, 0, 0, 7, 0, 1, 0, 2, 0},
																				 {1, 0, 2, 1, 5, 6, 0, 0, 2, 3}, {1, 0, 2, 0, 7, 8, 0, 0, 2, 2}, {1, 0, 0, 0, 7, 7, 1, 0, 4, 0}, {2, 6, 0, 0, 5, 5, 1, 0, 2, 0}};
__device__ void bc7_block(const uint4 w, uint32_t out[16]) {
	BitReader r{(uint64_t) w.x | ((uint64_t) w.y << 32), (uint64_t) w.z | ((uint64_t) w.w << 32), 0};
	int mode = 0;
	while (mode < 8 && !((w.x >> mode) & 1u)) mode++;
	if (mode >= 8) {
		for (int i = 0; i < 16; i++) out[i] = 0;
		return;
	}
	const uint8_t *M = kBc7Modes[mode];
	const int ns = M[0], cb = M[4], ab = M[5], ib = M[8], ib2 = M[9];
	r.pos = mode + 1;
	const uint32_t part = r.get(M[1]), rot = r.get(M[2]), isb = r.get(M[3]);
	uint32_t ep[6][4];
	for (int ch = 0; ch < 3; ch++)
		for (int e = 0; e < 2 * ns; e++) ep[e][ch] = r.get(cb);
	for (int e = 0; e < 2 * ns; e++) ep[e][3] = ab ? r.get(ab) : 255u;
	int cbits = cb, abits = ab;
	if (M[6]) {
		for (int e = 0; e < 2 * ns; e++) {
			const uint32_t p = r.get(1);
			for (int ch = 0; ch < (ab ? 4 : 3); ch++) ep[e][ch] = (ep[e][ch] << 1) | p;
		}
		cbits++;
		if (ab) abits++;
	} else if (M[7]) {
		for (int s = 0; s < ns; s++) {
			const uint32_t p = r.get(1);
			for (int e = 2 * s; e < 2 * s + 2; e++)
				for (int ch = 0; ch < 3; ch++) ep[e][ch] = (ep[e][ch] << 1) | p;
		}
		cbits++;
	}
	for (int e = 0; e < 2 * ns; e++) {
		for (int ch = 0; ch < 3; ch++) {
			const uint32_t v = ep[e][ch] << (8 - cbits);
			ep[e][ch] = v | (v >> cbits);
		}
		if (ab) {
			const uint32_t v = ep[e][3] << (8 - abits);
			ep[e][3] = v | (v >> abits);
		}
	}
	const int anchor1 = ns == 2 ? kBc7Anchor2[part] : (ns == 3 ? kBc7Anchor3a[part] : -1), anchor2 = ns == 3 ? kBc7Anchor3b[part] : -1;
	uint32_t i1[16], i2[16];
	for (int i = 0; i < 16; i++) i1[i] = r.get(ib - ((i == 0 || i == anchor1 || i == anchor2) ? 1 : 0));
	for (int i = 0; i < 16; i++) i2[i] = ib2 ? r.get(ib2 - (i == 0 ? 1 : 0)) : 0u;
	for (int i = 0; i < 16; i++) {
		const int s = ns == 1 ? 0 : (ns == 2 ? (kBc7Part2[part] >> i) & 1 : (kBc7Part3[part] >> (2 * i)) & 3);
		const uint32_t *e0 = ep[2 * s], *e1 = ep[2 * s + 1];
		uint32_t ci = i1[i], ai = i1[i];
		int cib = ib, aib = ib;
		if (ib2) {
			if (isb) { ci = i2[i]; cib = ib2; }
			else { ai = i2[i]; aib = ib2; }
		}
		uint32_t px[4];
		const uint32_t wc = bptc_weight(cib, ci), wa = bptc_weight(aib, ai);
		for (int ch = 0; ch < 3; ch++) px[ch] = ((64 - wc) * e0[ch] + wc * e1[ch] + 32) >> 6;
		px[3] = ab ? ((64 - wa) * e0[3] + wa * e1[3] + 32) >> 6 : 255u;
		if (rot) {
			const uint32_t t = px[3];
			px[3] = px[rot - 1];
			px[rot - 1] = t;
		}
		out[i] = px[0] | (px[1] << 8) | (px[2] << 16) | (px[3] << 24);
	}
}

// BC6H: field bit layouts of the 14 modes as (field * 16 + bit) per header bit, packed by the host at first use
struct Bc6Mode {
	uint8_t code, code_bits, regions, transformed, epb, db[3];
	uint8_t layout[82]; // 0xff = unused
};
__constant__ Bc6Mode c_bc6_modes[14];
__device__ __forceinline__ int sext(int v, int bits) { return (v & (1 << (bits - 1))) ? v - (1 << bits) : v; }
__device__ int bc6_unquantize(int comp, int epb, bool sgn) {
	if (!sgn) {
		if (epb >= 15) return comp;
		if (comp == 0) return 0;
		if (comp == (1 << epb) - 1) return 0xFFFF;
		return ((comp << 15) + 0x4000) >> (epb - 1);
	}
	if (epb >= 16) return comp;
	const bool s = comp < 0;
	if (s) comp = -comp;
	int unq;
	if (comp == 0) unq = 0;
	else if (comp >= (1 << (epb - 1)) - 1) unq = 0x7FFF;
	else unq = ((comp << 15) + 0x4000) >> (epb - 1);
	return s ? -unq : unq;
}
__device__ uint32_t bc6_finish(int v, bool sgn) {
	if (!sgn) return (uint32_t) ((v * 31) >> 6);
	if (v < 0) return 0x8000u | (uint32_t) (((-v) * 31) >> 5);
	return (uint32_t) ((v * 31) >> 5);
}
__device__ void bc6h_block(const uint4 w, bool sgn, uint2 out[16]) { // RGBA16F bit patterns, A = 1.0
	const BitReader r{(uint64_t) w.x | ((uint64_t) w.y << 32), (uint64_t) w.z | ((uint64_t) w.w << 32), 0};
	const uint32_t two = w.x & 3u, five = w.x & 31u;
	int m = -1;
	for (int i = 0; i < 14; i++)
		if ((c_bc6_modes[i].code_bits == 2 && c_bc6_modes[i].code == two) || (c_bc6_modes[i].code_bits == 5 && two >= 2 && c_bc6_modes[i].code == five)) { m = i; break; }
	if (m < 0) {
		for (int i = 0; i < 16; i++) out[i] = make_uint2(0u, 0x3C000000u);
		return;
	}
	const Bc6Mode &M = c_bc6_modes[m];
	int f[12];
	for (int i = 0; i < 12; i++) f[i] = 0;
	const int header = M.regions == 2 ? 77 : 65;
	for (int p = M.code_bits; p < header; p++) {
		const uint8_t d = M.layout[p];
		if (d != 0xff) f[d >> 4] |= (int) r.bit(p) << (d & 15);
	}
	int pos = header, part = 0;
	if (M.regions == 2)
		for (int k = 0; k < 5; k++) part |= (int) r.bit(pos++) << k;
	int ep[4][3];
	const int mask = (1 << M.epb) - 1;
	for (int ch = 0; ch < 3; ch++) {
		int base = f[ch];
		if (sgn) base = sext(base, M.epb);
		ep[0][ch] = base;
		for (int e = 1; e < 2 * M.regions; e++) {
			int v = f[e * 3 + ch];
			if (M.transformed) {
				v = (base + sext(v, M.db[ch])) & mask;
				if (sgn) v = sext(v, M.epb);
			} else if (sgn) v = sext(v, M.epb);
			ep[e][ch] = v;
		}
	}
	for (int e = 0; e < 2 * M.regions; e++)
		for (int ch = 0; ch < 3; ch++) ep[e][ch] = bc6_unquantize(ep[e][ch], M.epb, sgn);
	const int ib = M.regions == 2 ? 3 : 4, anchor = M.regions == 2 ? kBc7Anchor2[part] : -1;
	for (int i = 0; i < 16; i++) {
		const int nb = ib - ((i == 0 || i == anchor) ? 1 : 0);
		uint32_t idx = 0;
		for (int k = 0; k < nb; k++) idx |= r.bit(pos++) << k;
		const int s = M.regions == 2 ? (kBc7Part2[part] >> i) & 1 : 0;
		const int wgt = (int) bptc_weight(ib, idx);
		uint32_t h[3];
		for (int ch = 0; ch < 3; ch++) h[ch] = bc6_finish((ep[2 * s][ch] * (64 - wgt) + ep[2 * s + 1][ch] * wgt + 32) >> 6, sgn);
		out[i] = make_uint2(h[0] | (h[1] << 16), h[2] | 0x3C000000u);
	}
}

__global__ void __launch_bounds__(128) decode_kernel(const DecParams p) {
	const uint64_t block = (uint64_t) blockIdx.x * 128 + threadIdx.x;
	if (block >= p.n_blocks) return;
	const uint64_t per_slice = (uint64_t) p.blocks_x * p.blocks_y;
	const uint32_t slice = (uint32_t) (block / per_slice);
	const uint32_t rem = (uint32_t) (block - (uint64_t) slice * per_slice);
	const uint32_t by = rem / p.blocks_x, bx = rem - by * p.blocks_x;
	uint8_t *base = p.dst + (uint64_t) slice * p.slice_pitch;
	const int codec = p.codec;
	if (codec == B200IC_BC6H) {
		uint2 px[16];
		bc6h_block(reinterpret_cast<const uint4 *>(p.blocks)[block], p.is_signed != 0, px);
		for (int i = 0; i < 16; i++) {
			const uint32_t x = bx * 4 + (i & 3), y = by * 4 + (i >> 2);
			if (x < p.width && y < p.height) *reinterpret_cast<uint2 *>(base + (uint64_t) y * p.row_pitch + 8ull * x) = px[i];
		}
		return;
	}
	if (codec == B200IC_BC4 || codec == B200IC_BC5) {
		const int nch = codec == B200IC_BC4 ? 1 : 2;
		for (int c = 0; c < nch; c++) {
			uint32_t a[16];
			alpha_block(reinterpret_cast<const uint2 *>(p.blocks)[block * nch + c], a);
			for (int i = 0; i < 16; i++) {
				const uint32_t x = bx * 4 + (i & 3), y = by * 4 + (i >> 2);
				if (x < p.width && y < p.height) base[(uint64_t) y * p.row_pitch + (uint64_t) x * nch + c] = (uint8_t) a[i];
			}
		}
		return;
	}
	uint32_t px[16];
	if (codec == B200IC_BC1) {
		colour_block(reinterpret_cast<const uint2 *>(p.blocks)[block], false, px);
	} else if (codec == B200IC_BC2 || codec == B200IC_BC3) {
		const uint2 aw = reinterpret_cast<const uint2 *>(p.blocks)[block * 2];
		colour_block(reinterpret_cast<const uint2 *>(p.blocks)[block * 2 + 1], true, px);
		uint32_t a[16];
		if (codec == B200IC_BC3) alpha_block(aw, a);
		else
			for (int i = 0; i < 16; i++) a[i] = (((i < 8 ? aw.x >> (4 * i) : aw.y >> (4 * (i - 8))) & 15u) * 17u);
		for (int i = 0; i < 16; i++) px[i] = (px[i] & 0x00ffffffu) | (a[i] << 24);
	} else {
		bc7_block(reinterpret_cast<const uint4 *>(p.blocks)[block], px);
	}
	for (int i = 0; i < 16; i++) {
		const uint32_t x = bx * 4 + (i & 3), y = by * 4 + (i >> 2);
		if (x < p.width && y < p.height) *reinterpret_cast<uint32_t *>(base + (uint64_t) y * p.row_pitch + 4ull * x) = px[i];
	}
}

// the specification's "name[hi:lo]" field runs per mode (LSB first), the same text oracle/bcdec.c parses
struct Bc6Spec {
	int code, code_bits, regions, transformed, epb, db[3];
	const char *seq;
};
const Bc6Spec kBc6Spec[14] = {
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

cudaError_t upload_bc6_modes() {
	static bool done[16] = {};
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev >= 0 && dev < 16 && done[dev]) return cudaSuccess;
	Bc6Mode modes[14];
	for (int m = 0; m < 14; m++) {
		const Bc6Spec &S = kBc6Spec[m];
		Bc6Mode &M = modes[m];
		M.code = (uint8_t) S.code;
		M.code_bits = (uint8_t) S.code_bits;
		M.regions = (uint8_t) S.regions;
		M.transformed = (uint8_t) S.transformed;
		M.epb = (uint8_t) S.epb;
		for (int k = 0; k < 3; k++) M.db[k] = (uint8_t) S.db[k];
		for (int i = 0; i < 82; i++) M.layout[i] = 0xff;
		int pos = S.code_bits;
		for (const char *p = S.seq; *p;) {
			while (*p == ' ') p++;
			if (!*p) break;
			const int ch = p[0] == 'r' ? 0 : (p[0] == 'g' ? 1 : 2), e = p[1] - 'w';
			p += 2;
			int lo = 0, hi;
			while (*p >= '0' && *p <= '9') lo = lo * 10 + (*p++ - '0');
			hi = lo;
			if (*p == '-') {
				p++;
				hi = 0;
				while (*p >= '0' && *p <= '9') hi = hi * 10 + (*p++ - '0');
			}
			for (int k = lo; k <= hi; k++) M.layout[pos++] = (uint8_t) (((e * 3 + ch) << 4) | k);
		}
	}
	const cudaError_t e = cudaMemcpyToSymbol(c_bc6_modes, modes, sizeof(modes));
	if (e == cudaSuccess && dev >= 0 && dev < 16) done[dev] = true;
	return e;
}

} // namespace

cudaError_t launch_decode(int codec, const void *blocks, uint32_t width, uint32_t height, uint32_t slices, int is_signed, void *dst, uint64_t row_pitch,
													cudaStream_t stream) {
	DecParams p;
	p.blocks = static_cast<const uint8_t *>(blocks);
	p.dst = static_cast<uint8_t *>(dst);
	p.width = width;
	p.height = height;
	p.blocks_x = (width + 3) / 4;
	p.blocks_y = (height + 3) / 4;
	p.n_blocks = (uint64_t) p.blocks_x * p.blocks_y * slices;
	p.codec = codec == B200IC_BC7_RG ? B200IC_BC7_AMD : codec;
	p.is_signed = is_signed;
	const uint32_t tb = codec == B200IC_BC4 ? 1 : (codec == B200IC_BC5 ? 2 : (codec == B200IC_BC6H ? 8 : 4));
	p.row_pitch = row_pitch ? row_pitch : (uint64_t) width * tb;
	p.slice_pitch = p.row_pitch * height;
	if (p.n_blocks == 0) return cudaSuccess;
	if (codec == B200IC_BC6H) {
		const cudaError_t e = upload_bc6_modes();
		if (e != cudaSuccess) return e;
	}
	decode_kernel<<<(unsigned) ((p.n_blocks + 127) / 128), 128, 0, stream>>>(p);
	return cudaGetLastError();
}

} // namespace b200ic
