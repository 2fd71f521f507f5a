/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's BC4/BC5 scalar-channel encoder.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this; the product
 * path (gfx_imagecompress_b200/csrc) never does.  Pinned against the compiled, unmodified reference
 * (oracle/_ref/libref_oracle.so) by tests/test_oracle_restatement.py and against tests/golden/.
 *
 * Every function cites the reference file:line (relative to the reference tree) it follows.
 * Build: gcc -std=c11 -O2 -ffp-contract=off (FMA contraction changes BC4/BC5 output).
 */
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <float.h>

#define R4_MAX_ERROR 128000.f      /* src/amd_bcx_body.cpp:43 */
#define R4_GBL_SCH_STEP 0.018f     /* :47,54 */
#define R4_GBL_SCH_EXT 0.1f        /* :48,55 */
#define R4_LCL_SCH_STEP 0.6f       /* :49,56 */

/* src/amd_bcx_body.cpp:1510-1548 RmpSrch1: error of snapping each unique value to the nearest of
 * npoints evenly spaced ramp points in [lo,hi]; bails out returning exactly maxerr once exceeded. */
static float r4_ramp_search(const float *v, const float *rpt, float maxerr, float lo, float hi, int n, int npoints) {
	float error = 0;
	const float step = (hi - lo) / (float) (npoints - 1);
	const float step_h = step * 0.5f;
	const float rstep = 1.0f / step;
	for (int i = 0; i < n; i++) {
		float q, del;
		if ((del = v[i] - lo) <= 0) q = lo;
		else if (v[i] - hi >= 0) q = hi;
		else q = (floorf((del + step_h) * rstep) * step) + lo;
		float d = v[i] - q;
		error += d * d * rpt[i];
		if (maxerr < error) { error = maxerr; break; }
	}
	return error;
}

/* src/amd_bcx_body.cpp:1555-1607 Refine1: 3x3 hill climb, moves {0,-1,+1} x m_step on each end. */
static float r4_refine(const float *v, const float *rpt, float maxerror, float *plo, float *phi, float m_step,
											 float lo_bnd, float hi_bnd, int n, int npoints) {
	static const float mv[3] = {0.f, -1.f, 1.f}; /* sMvF :580 */
	float lo = *plo, hi = *phi;
	int best;
	do {
		float lo0 = lo, hi0 = hi;
		best = -1;
		for (int mode = 0; mode < 9; mode++) {
			float cl = lo + m_step * mv[mode / 3];
			float ch = hi + m_step * mv[mode % 3];
			cl = cl > lo_bnd ? cl : lo_bnd; /* Math_MaxF */
			ch = ch < hi_bnd ? ch : hi_bnd; /* Math_MinF */
			float e = r4_ramp_search(v, rpt, maxerror, cl, ch, n, npoints);
			if (e < maxerror) { maxerror = e; best = mode; lo0 = cl; hi0 = ch; }
		}
		if (best != -1) { lo = lo0; hi = hi0; }
	} while (best != -1);
	*plo = lo; *phi = hi;
	return maxerror;
}

static void r4_sort16(float *a, int n) { /* qsort with QSortFCmp :1609-1618 -- a total order on non-NaN floats */
	for (int i = 1; i < n; i++) {
		float k = a[i]; int j = i - 1;
		while (j >= 0 && a[j] - k > 0.) { a[j + 1] = a[j]; j--; }
		a[j + 1] = k;
	}
}

/* src/amd_bcx_body.cpp:1633-1832 CompBlock1 with _IntPrc=8,_FracPrc=0,_bFixedRamp=true (the only
 * configuration Image_CompressAMDAlphaSingleModeBlock uses, src/amd_bcx_helpers.cpp:132-134). */
static float r4_fit(float ramp_out[2], const float *blk, int nblk, int npoints, int fixedpts) {
	float fMaxError = 0.f, Ramp[2];
	const float IntFctr = 256.f;
	float uv[64], rp[64], s[64];
	for (int i = 0; i < 64; i++) uv[i] = rp[i] = 0.f;
	memcpy(s, blk, nblk * sizeof(float));
	r4_sort16(s, nblk);
	float new_p = -2.f;
	int nu = 0, need = 1;
	if (fixedpts) { /* :1664-1709 */
		for (int i = 0; i < nblk; i++) {
			if (new_p != s[i]) {
				new_p = s[i];
				if (new_p <= 1.5 / 255.) { /* double compare */ }
				else if (new_p >= 253.5 / 255.) { }
				else { uv[nu] = s[i]; rp[nu] = 1.f; nu++; }
			} else if (nu > 0 && uv[nu - 1] == new_p) rp[nu - 1] += 1.f;
		}
		if (nu <= 2) {
			if (nu == 2) { Ramp[0] = floorf(uv[0] * (IntFctr - 1) + 0.5f); Ramp[1] = floorf(uv[1] * (IntFctr - 1) + 0.5f); }
			else if (nu == 1) { Ramp[0] = floorf(uv[0] * (IntFctr - 1) + 0.5f); Ramp[1] = Ramp[0] + 1.f; }
			else { Ramp[0] = 128.f; Ramp[1] = Ramp[0] + 1.f; }
			fMaxError = 0.f; need = 0;
		}
	} else { /* :1710-1735 */
		for (int i = 0; i < nblk; i++) {
			if (new_p != s[i]) { uv[nu] = new_p = s[i]; rp[nu] = 1.f; nu++; }
			else rp[nu - 1] += 1.f;
		}
		if (nu <= 2) {
			Ramp[0] = floorf(uv[0] * (IntFctr - 1) + 0.5f);
			if (nu == 1) Ramp[1] = Ramp[0] + 1.f;
			else Ramp[1] = floorf(uv[1] * (IntFctr - 1) + 0.5f);
			fMaxError = 0.f; need = 0;
		}
	}
	if (need) { /* :1737-1817 */
		float min_ex = uv[0], max_ex = uv[nu - 1];
		float min_bnd = 0, max_bnd = 1.f;
		float min_r = min_ex, max_r = max_ex;
		float gbl_l = 0, gbl_r = 0;
		float cntr = (min_r + max_r) / 2;
		float gbl_err = R4_MAX_ERROR;
		int wantsSearch = !(max_ex - min_ex <= 48.f / IntFctr); /* _INT_GRID is true */
		if (wantsSearch) {
			float llb = (min_bnd > min_r - R4_GBL_SCH_EXT) ? min_bnd : min_r - R4_GBL_SCH_EXT;
			float rrb = (max_bnd < max_r + R4_GBL_SCH_EXT) ? max_bnd : max_r + R4_GBL_SCH_EXT;
			float lrb = (cntr < min_r + R4_GBL_SCH_EXT) ? cntr : min_r + R4_GBL_SCH_EXT;
			float rlb = (cntr > max_r - R4_GBL_SCH_EXT) ? cntr : max_r - R4_GBL_SCH_EXT;
			for (float sl = llb; sl < lrb; sl += R4_GBL_SCH_STEP)
				for (float sr = rrb; rlb <= sr; sr -= R4_GBL_SCH_STEP) {
					float e = r4_ramp_search(uv, rp, gbl_err, sl, sr, nu, npoints);
					if (e < gbl_err) { gbl_err = e; gbl_l = sl; gbl_r = sr; }
				}
			min_r = gbl_l; max_r = gbl_r;
		}
		float m_step = R4_LCL_SCH_STEP / IntFctr;
		fMaxError = r4_refine(uv, rp, gbl_err, &min_r, &max_r, m_step, min_bnd, max_bnd, nu, npoints);
		min_ex = min_r; max_ex = max_r;
		max_ex *= (IntFctr - 1);
		min_ex *= (IntFctr - 1);
		/* :1801 re-refine branch is dead: _INT_GRID is true */
		Ramp[1] = floorf(max_ex + 0.5f);
		Ramp[0] = floorf(min_ex + 0.5f);
	}
	if (Ramp[0] == Ramp[1]) { /* :1821-1827 */
		if (Ramp[1] < 255.f) Ramp[1]++; else Ramp[1]--;
	}
	ramp_out[0] = Ramp[0]; ramp_out[1] = Ramp[1];
	return fMaxError;
}

/* src/amd_bcx_body.cpp:1452-1505 Clstr1 + :1409-1447 GetRmp1 + :1395-1405 BldRmp1 (8.0 fixed-point grid) */
static float r4_cluster(uint8_t *idx, const float *blk, float ramp[2], int n, int npoints, int fixedpts) {
	float Err = 0.f, alpha[16];
	for (int i = 0; i < n; i++) idx[i] = 0;
	if (ramp[0] == ramp[1]) return Err;
	if ((!fixedpts && ramp[0] <= ramp[1]) || (fixedpts && ramp[0] > ramp[1])) {
		float t = ramp[0]; ramp[0] = ramp[1]; ramp[1] = t;
	}
	for (int e = npoints; e < 16; e++) alpha[e] = 100000.f;
	alpha[0] = ramp[0]; alpha[1] = ramp[1];
	for (int e = 1; e < npoints - 1; e++)
		alpha[e + 1] = (alpha[0] * (npoints - 1 - e) + alpha[1] * e) / (float) (npoints - 1);
	if (fixedpts) { alpha[npoints] = 0.f; alpha[npoints + 1] = 1.f * 256.f - 1.f; }
	for (int i = 0; i < npoints; i++) { alpha[i] = floorf(alpha[i] + 0.5f); alpha[i] /= 1.f; }
	if (fixedpts) npoints += 2;
	const float OverIntFctr = 1.f / (256.f - 1.f);
	for (int i = 0; i < npoints; i++) alpha[i] *= OverIntFctr;
	for (int i = 0; i < n; i++) {
		float shortest = 10000000.f, a = blk[i];
		for (int j = 0; j < npoints; j++) {
			float d = a - alpha[j]; d *= d;
			if (d < shortest) { shortest = d; idx[i] = (uint8_t) j; }
		}
		Err += shortest;
	}
	return Err;
}

/* src/amd_bcx_body.cpp:1848-1868 CompBlock1X */
static float r4_comp(const float *blk, uint8_t ep[2], uint8_t *idx, int npoints, int fixedpts) {
	float Ramp[2];
	r4_fit(Ramp, blk, 16, npoints, fixedpts);
	float e = r4_cluster(idx, blk, Ramp, 16, npoints, fixedpts);
	ep[0] = (uint8_t) Ramp[0]; ep[1] = (uint8_t) Ramp[1];
	return e;
}

/* src/amd_bcx_helpers.cpp:32-46 EncodeAlphaBlock */
static void r4_pack(uint32_t out[2], const uint8_t ep[2], const uint8_t idx[16]) {
	out[0] = ((int) ep[0]) | (((int) ep[1]) << 8);
	out[1] = 0;
	for (int i = 0; i < 16; i++) {
		if (i < 5) out[0] |= (uint32_t) (idx[i] & 7) << (16 + i * 3);
		else if (i > 5) out[1] |= (uint32_t) (idx[i] & 7) << (2 + (i - 6) * 3);
		else { out[0] |= (uint32_t) (idx[i] & 1) << 31; out[1] |= (uint32_t) (idx[i] & 6) >> 1; }
	}
}

/* src/amd_bcx_helpers.cpp:125-140 Image_CompressAMDAlphaSingleModeBlock */
void restate_alpha_block(const float in[16], void *out) {
	uint8_t ep[2][2], idx[2][16];
	float e8 = r4_comp(in, ep[0], idx[0], 8, 0);
	float e6 = (e8 == 0.f) ? FLT_MAX : r4_comp(in, ep[1], idx[1], 6, 1);
	uint32_t blk[2];
	if (e8 <= e6) r4_pack(blk, ep[0], idx[0]); else r4_pack(blk, ep[1], idx[1]);
	memcpy(out, blk, 8);
}

/* Image loops: src/amd_bc4_compressor.cpp:27-44 (channel 1!), src/amd_bc5_compressor.cpp:27-48
 * (channels 0 then 1) with the replicate-edge gather of src/block_utils.cpp:116-144.
 * `pixels` is tightly packed u8 with `nch` channels; missing channels read 0 (g,b) / 1 (a) per the
 * compat shim's Image_GetPixelAtF (compat/gfx_image/image.h). */
static float r4_texel(const uint8_t *pixels, uint32_t w, uint32_t h, int nch, uint32_t x, uint32_t y, int ch) {
	if (x >= w) x = w - 1;
	if (y >= h) y = h - 1;
	if (ch >= nch) return ch == 3 ? 1.0f : 0.0f;
	return pixels[((size_t) y * w + x) * nch + ch] / 255.0f;
}
void restate_bc4_image(const uint8_t *pixels, uint32_t w, uint32_t h, int nch, uint8_t *dst) {
	uint32_t bx = (w + 3) / 4, by = (h + 3) / 4;
	for (uint32_t y = 0; y < by; y++) for (uint32_t x = 0; x < bx; x++) {
		float b[16];
		for (int i = 0; i < 16; i++) b[i] = r4_texel(pixels, w, h, nch, x * 4 + (i & 3), y * 4 + (i >> 2), 1);
		restate_alpha_block(b, dst + ((size_t) y * bx + x) * 8);
	}
}
void restate_bc5_image(const uint8_t *pixels, uint32_t w, uint32_t h, int nch, uint8_t *dst) {
	uint32_t bx = (w + 3) / 4, by = (h + 3) / 4;
	for (uint32_t y = 0; y < by; y++) for (uint32_t x = 0; x < bx; x++) {
		float r[16], g[16];
		for (int i = 0; i < 16; i++) {
			r[i] = r4_texel(pixels, w, h, nch, x * 4 + (i & 3), y * 4 + (i >> 2), 0);
			g[i] = r4_texel(pixels, w, h, nch, x * 4 + (i & 3), y * 4 + (i >> 2), 1);
		}
		restate_alpha_block(r, dst + ((size_t) y * bx + x) * 16);
		restate_alpha_block(g, dst + ((size_t) y * bx + x) * 16 + 8);
	}
}
