// bc1_core.cuh -- AMD-Compressonator-compatible BC1 endpoint search, one 4x4 block per thread (bit-exact target).
//
// Follows reference src/amd_bcx_helpers.cpp:51-105 (Image_CompressAMDBC1Block), src/block_utils.cpp:162-174 (fixed
// colour weights), src/amd_bcx_body.cpp:1209-1297 (CompRGBABlock), :937-1203 (CompressRGBBlockX), :442-570 (FindAxis),
// :398-435 (RampSrchW), :122-151 (MkRmpOnGrid), :157-197 (MkWkRmpPts / BldClrRmp), :203-246 (ClstrErr), :582-806 (Refine),
// :258-378 (ClstrIntnl / ClstrBas / Clstr).
//
// All arithmetic is FP32 in the reference's operation order with no contraction (--fmad=false / -ffp-contract=off),
// plus the three places where the reference promotes to double (`BlkIn *= 255.0` :1267, `Err + 0.001 < ErrG` :1118,
// the final error compare). Internal channel order is the reference's: 0 = blue, 1 = green, 2 = red.
// Not built: AdaptiveColourWeights (reads uninitialised memory in the reference) and b3DRefinement (Refine3D).
#pragma once
#include <stdint.h>
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define B1_HD __host__ __device__ __forceinline__
#define B1_HDN __host__ __device__ __noinline__
#else
#define B1_HD inline
#define B1_HDN
#endif

namespace b200ic {
namespace bc1 {

B1_HD float fmin_ref(float a, float b) { return a < b ? a : b; } // Math_MinF / Math_MaxF of the compat shim
B1_HD float fmax_ref(float a, float b) { return a > b ? a : b; }
B1_HD uint32_t fbits(float f) {
#if defined(__CUDA_ARCH__)
	return __float_as_uint(f);
#else
	union { float f; uint32_t u; } c; c.f = f; return c.u;
#endif
}

// channel bit depths on the 5-6-5 grid, indexed by the internal channel order (B, G, R)
B1_HD float grid_levels(int j) { return j == 1 ? 64.f : 32.f; } // 1 << bits
constexpr int kRefine3D = 0x100; // flag in the `steps` argument: Refine3D instead of Refine (src/amd_bcx_body.cpp:1199-1202)
B1_HD float grid_step(int j) { return j == 1 ? 4.f : 8.f; }     // 1 << (8 - bits)
// channel weights: reference passes {0.3086 (R), 0.6094 (G), 0.0820 (B)}; internal order B, G, R
B1_HD float chan_weight(int j) { return j == 0 ? 0.0820f : (j == 1 ? 0.6094f : 0.3086f); }

// MkWkRmpPts (:157-181): endpoints as the decompressor expands them; eq = ramp is flat
B1_HD void expand_endpoints(bool &eq, float out[3][2], const float in[3][2]) {
	eq = true;
	for (int j = 0; j < 3; j++) eq = eq && (in[j][0] == in[j][1]);
	for (int j = 0; j < 3; j++)
		for (int k = 0; k < 2; k++) {
			float v = in[j][k] + floorf(in[j][k] / grid_levels(j));
			v = fmax_ref(v, 0.f);
			v = fmin_ref(v, 255.f);
			out[j][k] = v;
		}
}
// BldClrRmp (:188-197)
B1_HD void build_ramp(float rmp[5], const float ep[2], int np) {
	rmp[0] = ep[0];
	rmp[np - 1] = ep[1];
	if (np % 2) rmp[np] = 1000000.f;
	const float rnd = np == 4 ? 1.f : 0.f; // dwRndAmount[3] = 0, [4] = 1
	for (int e = 1; e < np - 1; e++) rmp[e] = floorf((rmp[0] * (float) (np - 1 - e) + rmp[np - 1] * (float) e + rnd) / (float) (np - 1));
}

// Refine (:582-806) with the fixed channel weights
B1_HD float refine(float out[3][2], const float in[3][2], const float blk[16][3], const float *rpt, int n, int np, int steps) {
	float rmp[3][5];
	float inp0[3][2], inp[3][2], wk[3][2];
	for (int k = 0; k < 2; k++)
		for (int j = 0; j < 3; j++) inp0[j][k] = inp[j][k] = out[j][k] = in[j][k];
	bool eq;
	expand_endpoints(eq, wk, inp);
	for (int j = 0; j < 3; j++) build_ramp(rmp[j], wk[j], np);
	// ClstrErr (:203-246)
	float best = 0.f;
	{
		const int len = eq ? 1 : np;
		for (int i = 0; i < n; i++) {
			float shortest = 99999999999.f;
			for (int r = 0; r < len; r++) {
				const float d = (blk[i][2] - rmp[2][r]) * (blk[i][2] - rmp[2][r]) * chan_weight(2) +
												(blk[i][1] - rmp[1][r]) * (blk[i][1] - rmp[1][r]) * chan_weight(1) +
												(blk[i][0] - rmp[0][r]) * (blk[i][0] - rmp[0][r]) * chan_weight(0);
				if (d < shortest) shortest = d;
			}
			best += shortest * rpt[i];
		}
	}
	if (best == 0.f || !steps) return best;
	const int span = steps < 8 ? steps : 8;
	float other[4][16]; // error of the two fixed channels per (ramp point, colour)
	// channels are tweaked in the order R, G, B = internal 2, 1, 0
	for (int pass = 0; pass < 3; pass++) {
		const int ch = 2 - pass;
		const int a = ch == 2 ? 1 : 2, b = ch == 0 ? 1 : 0; // the two fixed channels, in the reference's summation order
		for (int i = 0; i < n; i++)
			for (int r = 0; r < np; r++) {
				const float da = rmp[a][r] - blk[i][a], db = rmp[b][r] - blk[i][b];
				other[r][i] = da * da * chan_weight(a) + db * db * chan_weight(b);
			}
		float b0 = inp0[ch][0], b1 = inp0[ch][1];
		for (int i = -span; i <= span; i++)
			for (int j = -span; j <= span; j++) {
				inp[ch][0] = fmin_ref(fmax_ref(inp0[ch][0] + (float) i * grid_step(ch), 0.f), 255.f);
				inp[ch][1] = fmin_ref(fmax_ref(inp0[ch][1] + (float) j * grid_step(ch), 0.f), 255.f);
				expand_endpoints(eq, wk, inp);
				build_ramp(rmp[ch], wk[ch], np);
				float mse = 0.f;
				const int len = eq ? 1 : np;
				for (int k = 0; k < n; k++) {
					float me = 10000000.f;
					for (int r = 0; r < len; r++) {
						const float d = rmp[ch][r] - blk[k][ch];
						const float e = other[r][k] + d * d * chan_weight(ch);
						me = fmin_ref(me, e);
					}
					mse += me * rpt[k];
				}
				if (mse < best) {
					b0 = inp[ch][0];
					b1 = inp[ch][1];
					best = mse;
				}
			}
		inp[ch][0] = b0;
		inp[ch][1] = b1;
		if (pass < 2) {
			expand_endpoints(eq, wk, inp);
			for (int j = 0; j < 3; j++) build_ramp(rmp[j], wk[j], np);
		}
	}
	for (int j = 0; j < 3; j++)
		for (int k = 0; k < 2; k++) out[j][k] = inp[j][k];
	return best;
}

// Refine3D (:808-932, the b3DRefinement option): all six endpoint components jittered together, +-min(steps, 8) grid
// steps each -- (2 steps + 1)^6 candidate ramps, G outermost, then B, then R (internal channels 1, 0, 2), each against
// decompressor-exact ramps; first strict minimum in scan order.
B1_HD float refine3d(float out[3][2], const float in[3][2], const float blk[16][3], const float *rpt, int n, int np, int steps) {
	float rmp[3][5];
	float inp0[3][2], inp[3][2], wk[3][2];
	for (int k = 0; k < 2; k++)
		for (int j = 0; j < 3; j++) inp0[j][k] = inp[j][k] = out[j][k] = in[j][k];
	bool eq;
	expand_endpoints(eq, wk, inp);
	for (int j = 0; j < 3; j++) build_ramp(rmp[j], wk[j], np);
	float best = 0.f; // ClstrErr (:203-246)
	{
		const int len = eq ? 1 : np;
		for (int i = 0; i < n; i++) {
			float shortest = 99999999999.f;
			for (int r = 0; r < len; r++) {
				const float d = (blk[i][2] - rmp[2][r]) * (blk[i][2] - rmp[2][r]) * chan_weight(2) +
												(blk[i][1] - rmp[1][r]) * (blk[i][1] - rmp[1][r]) * chan_weight(1) +
												(blk[i][0] - rmp[0][r]) * (blk[i][0] - rmp[0][r]) * chan_weight(0);
				if (d < shortest) shortest = d;
			}
			best += shortest * rpt[i];
		}
	}
	if (best == 0.f || !steps) return best;
	const int span = steps < 8 ? steps : 8;
	float err_g[4][16], err_gb[4][16];
	for (int g0 = -span; g0 <= span; g0++) {
		inp[1][0] = fmin_ref(fmax_ref(inp0[1][0] + (float) g0 * grid_step(1), 0.f), 255.f);
		for (int g1 = -span; g1 <= span; g1++) {
			inp[1][1] = fmin_ref(fmax_ref(inp0[1][1] + (float) g1 * grid_step(1), 0.f), 255.f);
			expand_endpoints(eq, wk, inp);
			build_ramp(rmp[1], wk[1], np);
			for (int i = 0; i < n; i++)
				for (int r = 0; r < np; r++) {
					const float d = rmp[1][r] - blk[i][1];
					err_g[r][i] = d * d * chan_weight(1);
				}
			for (int b0 = -span; b0 <= span; b0++) {
				inp[0][0] = fmin_ref(fmax_ref(inp0[0][0] + (float) b0 * grid_step(0), 0.f), 255.f);
				for (int b1 = -span; b1 <= span; b1++) {
					inp[0][1] = fmin_ref(fmax_ref(inp0[0][1] + (float) b1 * grid_step(0), 0.f), 255.f);
					expand_endpoints(eq, wk, inp);
					build_ramp(rmp[0], wk[0], np);
					for (int i = 0; i < n; i++)
						for (int r = 0; r < np; r++) {
							const float d = rmp[0][r] - blk[i][0];
							err_gb[r][i] = err_g[r][i] + d * d * chan_weight(0);
						}
					for (int r0 = -span; r0 <= span; r0++) {
						inp[2][0] = fmin_ref(fmax_ref(inp0[2][0] + (float) r0 * grid_step(2), 0.f), 255.f);
						for (int r1 = -span; r1 <= span; r1++) {
							inp[2][1] = fmin_ref(fmax_ref(inp0[2][1] + (float) r1 * grid_step(2), 0.f), 255.f);
							expand_endpoints(eq, wk, inp);
							build_ramp(rmp[2], wk[2], np);
							float mse = 0.f;
							const int len = eq ? 1 : np;
							for (int k = 0; k < n; k++) {
								float me = 10000000.f;
								for (int r = 0; r < len; r++) {
									const float d = rmp[2][r] - blk[k][2];
									me = fmin_ref(me, err_gb[r][k] + d * d * chan_weight(2));
								}
								mse += me * rpt[k];
							}
							if (mse < best) {
								best = mse;
								for (int k = 0; k < 2; k++)
									for (int j = 0; j < 3; j++) out[j][k] = inp[j][k];
							}
						}
					}
				}
			}
		}
	}
	return best;
}

// FindAxis (:442-570). Returns false when the colour set is too small in diameter (the reference's *_pbSmall).
B1_HD bool find_axis(float sh[16][3], float dir[3], float centre[3], const float blk[16][3], const float *rpt, int n) {
	float crrl[3] = {0.f, 0.f, 0.f}, rgb2[3] = {0.f, 0.f, 0.f};
	dir[0] = dir[1] = dir[2] = 0.f;
	centre[0] = centre[1] = centre[2] = 0.f;
	float np = 0.f;
	for (int i = 0; i < n; i++) {
		centre[0] += blk[i][0] * rpt[i];
		centre[1] += blk[i][1] * rpt[i];
		centre[2] += blk[i][2] * rpt[i];
		np += rpt[i];
	}
	centre[0] /= np;
	centre[1] /= np;
	centre[2] /= np;
	for (int i = 0; i < n; i++) {
		sh[i][0] = blk[i][0] - centre[0];
		sh[i][1] = blk[i][1] - centre[1];
		sh[i][2] = blk[i][2] - centre[2];
		for (int j = 0; j < 3; j++) {
			rgb2[j] += sh[i][j] * sh[i][j] * rpt[i];
			crrl[j] += sh[i][j] * sh[i][(j + 1) % 3] * rpt[i];
		}
	}
	int i0 = 0, k = 0;
	float mx = 0.f;
	const float c = 2.f / 255.f;
	const float eps = np * c * c; // the reference's EPS macro is unparenthesised: (N * c) * c
	for (int j = 0; j < 3; j++) {
		if (rgb2[j] >= eps) k++;
		else rgb2[j] = 0.f;
		if (mx < rgb2[j]) { mx = rgb2[j]; i0 = j; }
	}
	const float eps2 = np * 3.f * c * c;
	bool small = true;
	for (int j = 0; j < 3; j++) small = small && (rgb2[j] < eps2);
	if (small) return false;
	if (k == 1) {
		dir[i0] = 1.f;
	} else if (k == 2) {
		const int i1 = (rgb2[(i0 + 1) % 3] > 0.f) ? (i0 + 1) % 3 : (i0 + 2) % 3;
		const float crl = (i1 == (i0 + 1) % 3) ? crrl[i0] : crrl[(i0 + 2) % 3];
		dir[i1] = crl / rgb2[i0];
		dir[i0] = 1.f;
	} else {
		// the reference starts its "largest determinant" scan at 100000.f, which no determinant of [0,1] data
		// exceeds: i0 stays the channel of largest variance and the solution is divided by 100000.f
		float max_det = 100000.f;
		for (int j = 0; j < 3; j++) {
			const float det = rgb2[j] * rgb2[(j + 1) % 3] - crrl[j] * crrl[j];
			if (max_det < det) { max_det = det; i0 = j; }
		}
		const float v0 = crrl[(i0 + 2) % 3], v1 = crrl[(i0 + 1) % 3];
		const float m00 = rgb2[(i0 + 1) % 3], m11 = rgb2[i0], m01 = -crrl[i0];
		float s0 = m00 * v0 + m01 * v1;
		float s1 = m01 * v0 + m11 * v1;
		s0 /= max_det;
		s1 /= max_det;
		dir[i0] = 1.f;
		dir[(i0 + 1) % 3] = 1.f;
		dir[(i0 + 2) % 3] = s0 + s1;
	}
	float len = dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2];
	len = sqrtf(len);
	for (int j = 0; j < 3; j++) dir[j] = (len > 0.f) ? dir[j] / len : 0.f;
	return true;
}

// RampSrchW (:398-435) evaluated completely: the reference's early-out returns exactly `_maxerror`, which can never
// pass the caller's strict `<`, and the partial sums are monotone, so the decision is identical.
B1_HD float ramp_error(const float *prj, const float *prj_err, const float *rep, float lo, float hi, int n, int np) {
	float error = 0;
	const float step = (hi - lo) / (float) (np - 1);
	const float step_h = step * 0.5f;
	const float rstep = 1.0f / step;
	for (int i = 0; i < n; i++) {
		float v;
		const float del = prj[i] - lo;
		if (del <= 0) v = lo;
		else if (prj[i] - hi >= 0) v = hi;
		else v = floorf((del + step_h) * rstep) * step + lo;
		float d = prj[i] - v;
		d *= d;
		error += rep[i] * d + prj_err[i];
	}
	return error;
}

// CompressRGBBlockX (:937-1203); in255[] = unique colours in 0..255, rpt = repeat counts
B1_HDN void fit_endpoints(float result[3][2], const float in255[16][3], const float *rpt, int n, int np, int steps) {
	float blk[16][3], sh[16][3];
	for (int i = 0; i < n; i++)
		for (int j = 0; j < 3; j++) blk[i][j] = in255[i][j] / 255.f;
	float rslt[3][2];
	bool done = false;
	float dir0[3], mdl[3];
	if (n <= 2) done = true;
	else if (!find_axis(sh, dir0, mdl, blk, rpt, n)) done = true;
	if (done) {
		for (int j = 0; j < 3; j++) {
			rslt[j][0] = in255[0][j];
			rslt[j][1] = in255[n - 1][j];
		}
	} else {
		float err_g = 10000000.f;
		float dir[3] = {dir0[0], dir0[1], dir0[2]}, dir_g[3] = {0.f, 0.f, 0.f}, pos_g[2] = {0.f, 0.f};
		for (;;) {
			float prj0[16], prj[16], prj_err[16], rep[16];
			float bnd[2] = {1000.f, -1000.f};
			for (int i = 0; i < n; i++) {
				const float p = sh[i][0] * dir[0] + sh[i][1] * dir[1] + sh[i][2] * dir[2];
				prj0[i] = prj[i] = p;
				prj_err[i] = (sh[i][0] - dir[0] * p) * (sh[i][0] - dir[0] * p) + (sh[i][1] - dir[1] * p) * (sh[i][1] - dir[1] * p) +
										 (sh[i][2] - dir[2] * p) * (sh[i][2] - dir[2] * p);
				bnd[0] = fmin_ref(bnd[0], p);
				bnd[1] = fmax_ref(bnd[1], p);
			}
			const float scl0 = bnd[0] - (bnd[1] - bnd[0]) * 0.125f, scl1 = bnd[1] + (bnd[1] - bnd[0]) * 0.125f;
			const float scl2 = (scl1 - scl0) * (scl1 - scl0);
			const float over = 1.f / (scl1 - scl0);
			for (int i = 0; i < n; i++) {
				prj[i] = (prj[i] - scl0) * over;
				rep[i] = rpt[i] * scl2;
			}
			for (int k = 0; k < 2; k++) bnd[k] = (bnd[k] - scl0) * over;
			float err = 128000.f; // MAX_ERROR
			const float stp = 0.025f;
			const float ls = (bnd[0] - 2.f * stp > 0.f) ? bnd[0] - 2.f * stp : 0.f;
			const float he = (bnd[1] + 2.f * stp < 1.f) ? bnd[1] + 2.f * stp : 1.f;
			float pos[2] = {0.f, 0.f};
			float lp = ls;
			for (int l = 0; l < 8; l++, lp += stp) {
				// the 8 candidates of a row together: every texel is loaded once per row instead of once per candidate (the
				// three per-texel arrays live in local memory; the search was waiting on those loads).  Each candidate's sum
				// still runs over the texels in order, the candidates are compared in scan order.
				float hps[8], e8[8], step8[8], steph8[8], rstep8[8];
				{
					float hp = he;
#pragma unroll
					for (int h = 0; h < 8; h++, hp -= stp) {
						hps[h] = hp;
						e8[h] = 0;
						step8[h] = (hp - lp) / (float) (np - 1);
						steph8[h] = step8[h] * 0.5f;
						rstep8[h] = 1.0f / step8[h];
					}
				}
#pragma unroll 1
				for (int i = 0; i < n; i++) {
					const float pi = prj[i], ri = rep[i], pe = prj_err[i];
					const float del = pi - lp;
#pragma unroll
					for (int h = 0; h < 8; h++) {
						float v;
						if (del <= 0) v = lp;
						else if (pi - hps[h] >= 0) v = hps[h];
						else v = floorf((del + steph8[h]) * rstep8[h]) * step8[h] + lp;
						float d = pi - v;
						d *= d;
						e8[h] += ri * d + pe;
					}
				}
#pragma unroll
				for (int h = 0; h < 8; h++)
					if (e8[h] < err) {
						err = e8[h];
						pos[0] = lp;
						pos[1] = hps[h];
					}
			}
			for (int k = 0; k < 2; k++) pos[k] = pos[k] * (scl1 - scl0) + scl0;
			if ((double) err + 0.001 < (double) err_g) {
				err_g = err;
				dir_g[0] = dir[0]; dir_g[1] = dir[1]; dir_g[2] = dir[2];
				pos_g[0] = pos[0];
				pos_g[1] = pos[1];
				const float step = (pos[1] - pos[0]) / (float) (np - 1);
				const float step_h = step * 0.5f;
				const float rstep = 1.0f / step;
				const float over_np = 1.f / (float) (np - 1);
				const float avg = (float) (np - 1) / 2.f;
				float crs[3] = {0.f, 0.f, 0.f}, len = 0.f;
				for (int i = 0; i < n; i++) {
					float ri;
					const float del = prj0[i] - pos[0];
					if (del <= 0) ri = 0.f;
					else if (prj0[i] - pos[1] >= 0) ri = (float) (np - 1);
					else ri = floorf((del + step_h) * rstep);
					ri = (ri - avg) * over_np;
					const float pre = ri * rpt[i];
					len += ri * pre;
					for (int j = 0; j < 3; j++) crs[j] += sh[i][j] * pre;
				}
				dir[0] = dir[1] = dir[2] = 0.f;
				if (len > 0.f) {
					dir[0] = crs[0] / len;
					dir[1] = crs[1] / len;
					dir[2] = crs[2] / len;
					float len2 = dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2];
					len2 = sqrtf(len2);
					dir[0] /= len2;
					dir[1] /= len2;
					dir[2] /= len2;
				}
			} else {
				break;
			}
		}
		for (int k = 0; k < 2; k++)
			for (int j = 0; j < 3; j++) rslt[j][k] = (pos_g[k] * dir_g[j] + mdl[j]) * 255.f;
	}
	// MkRmpOnGrid (:122-151), min 0 / max 255
	float grid[3][2];
	for (int j = 0; j < 3; j++)
		for (int k = 0; k < 2; k++) {
			float r = floorf(rslt[j][k]);
			if (r <= 0.f) r = 0.f;
			else {
				r += floorf(128.f / grid_levels(j)) - floorf(r / grid_levels(j));
				r = fmin_ref(r, 255.f);
			}
			grid[j][k] = floorf(r / grid_step(j)) * grid_step(j);
		}
	// `steps` carries the b3DRefinement option in bit 8 (kRefine3D)
	if (steps & kRefine3D) refine3d(result, grid, in255, rpt, n, np, steps & 0xff);
	else refine(result, grid, in255, rpt, n, np, steps);
}

// CompRGBABlock (:1209-1297). in = 16 RGBA texels 0..1. Returns the clustering error; ep[channel B,G,R][2], idx[16].
B1_HDN float compress(const float in[64], int np, bool use_alpha, float thr, int steps, uint8_t ep[3][2], uint8_t idx[16]) {
	float col[16][3]; // B, G, R of the opaque texels
	int m = 0;
	for (int i = 0; i < 16; i++)
		if (!use_alpha || in[i * 4 + 3] >= thr) {
			col[m][0] = in[i * 4 + 2];
			col[m][1] = in[i * 4 + 1];
			col[m][2] = in[i * 4 + 0];
			m++;
		}
	if (m == 0) { // everything transparent
		for (int j = 0; j < 3; j++) { ep[j][0] = 0; ep[j][1] = 0xff; }
		for (int i = 0; i < 16; i++) idx[i] = 0xff;
		return 0.0f;
	}
	if (m != 16 && use_alpha && !(np & 1)) return FLT_MAX;
	// sort by the bit patterns of (R, G, B) -- QSortFloatCmp compares element [2], then [1], then [0] (:103-117)
	for (int i = 1; i < m; i++) {
		const float c0 = col[i][0], c1 = col[i][1], c2 = col[i][2];
		const uint32_t k2 = fbits(c2), k1 = fbits(c1), k0 = fbits(c0);
		int j = i;
		while (j > 0) {
			const uint32_t p2 = fbits(col[j - 1][2]), p1 = fbits(col[j - 1][1]), p0 = fbits(col[j - 1][0]);
			const bool greater = (p2 > k2) || (p2 == k2 && p1 > k1) || (p2 == k2 && p1 == k1 && p0 > k0);
			if (!greater) break;
			col[j][0] = col[j - 1][0]; col[j][1] = col[j - 1][1]; col[j][2] = col[j - 1][2];
			j--;
		}
		col[j][0] = c0; col[j][1] = c1; col[j][2] = c2;
	}
	float uniq[16][3], rpt[16];
	int n = 0;
	for (int i = 0; i < m; i++) {
		if (i > 0 && fbits(col[i][0]) == fbits(col[i - 1][0]) && fbits(col[i][1]) == fbits(col[i - 1][1]) && fbits(col[i][2]) == fbits(col[i - 1][2])) {
			rpt[n - 1] += 1.f;
		} else {
			for (int j = 0; j < 3; j++) uniq[n][j] = (float) ((double) col[i][j] * 255.0);
			rpt[n] = 1.f;
			n++;
		}
	}
	float r[3][2];
	fit_endpoints(r, uniq, rpt, n, np, steps);
	for (int j = 0; j < 3; j++)
		for (int k = 0; k < 2; k++) ep[j][k] = (uint8_t) r[j][k];
	// Clstr (:342-378) -> ClstrBas (:322-337) -> ClstrIntnl (:258-317)
	const uint32_t c0 = ((uint32_t) (ep[2][0] & 0xf8) << 8) | ((uint32_t) (ep[1][0] & 0xfc) << 3) | ((uint32_t) (ep[0][0] & 0xf8) >> 3);
	const uint32_t c1 = ((uint32_t) (ep[2][1] & 0xf8) << 8) | ((uint32_t) (ep[1][1] & 0xfc) << 3) | ((uint32_t) (ep[0][1] & 0xf8) >> 3);
	int e0 = 0, e1 = 1;
	if ((!(np & 1) && c0 <= c1) || ((np & 1) && c0 > c1)) { e0 = 1; e1 = 0; }
	float inp[3][2], wk[3][2], rmp[3][5];
	for (int j = 0; j < 3; j++) {
		inp[j][0] = (float) ep[j][e0];
		inp[j][1] = (float) ep[j][e1];
	}
	bool eq;
	expand_endpoints(eq, wk, inp);
	for (int j = 0; j < 3; j++) build_ramp(rmp[j], wk[j], np);
	const float thr255 = thr * 255.f;
	const int len = eq ? 1 : np;
	float err = 0.f;
	for (int i = 0; i < 16; i++) {
		const float a255 = in[i * 4 + 3] * 255.0f;
		if (use_alpha && !(a255 >= thr255)) {
			idx[i] = (uint8_t) np;
			continue;
		}
		const float b = in[i * 4 + 2] * 255.0f, g = in[i * 4 + 1] * 255.0f, rr = in[i * 4 + 0] * 255.0f;
		float shortest = 99999999999.f;
		int si = 0;
		for (int q = 0; q < len; q++) {
			const float d = (rr - rmp[2][q]) * (rr - rmp[2][q]) * chan_weight(2) + (g - rmp[1][q]) * (g - rmp[1][q]) * chan_weight(1) +
											(b - rmp[0][q]) * (b - rmp[0][q]) * chan_weight(0);
			if (d < shortest) { shortest = d; si = q; }
		}
		err += shortest;
		if (si == np - 1) si = 1;
		else if (si) si++;
		idx[i] = (uint8_t) si;
	}
	return err;
}

// One texel of Image_CompressAMDExplictAlphaSingleModeBlock (src/amd_bcx_helpers.cpp:107-123): 8 -> 4 bits with the
// reference's rounding (+7 or +8 minus the high nibble)
B1_HD uint32_t explicit_alpha4(float a) {
	uint8_t c = (uint8_t) (a * 255.0f);
	c = (uint8_t) ((c + ((c >> 4) < 0x8 ? 7 : 8) - (c >> 4)) >> 4);
	return c > 0xf ? 0xfu : (uint32_t) c;
}

// The two halves of Image_CompressAMDBC1Block for a lane pair: `which` 0 = the 3-point fit, 1 = the 4-point fit (the
// reference skips the latter when the former is exact; its result cannot win then, e3 = 0 <= e4, so running it anyway
// changes nothing).  pack_fit() is the selection and packing (:89-104) from the two errors.
B1_HD double fit_half(const float in[64], int which, float alpha_threshold, int steps, uint8_t ep[3][2], uint8_t idx[16]) {
	return (double) compress(in, which ? 4 : 3, alpha_threshold > 0.0f, alpha_threshold, steps, ep, idx);
}
B1_HD void pack_fit(int m, const uint8_t ep[3][2], const uint8_t idx[16], uint32_t out[2]) {
	const uint32_t c0 = ((uint32_t) (ep[2][0] >> 3) << 11) | ((uint32_t) (ep[1][0] >> 2) << 5) | (uint32_t) (ep[0][0] >> 3);
	const uint32_t c1 = ((uint32_t) (ep[2][1] >> 3) << 11) | ((uint32_t) (ep[1][1] >> 2) << 5) | (uint32_t) (ep[0][1] >> 3);
	if ((m == 1 && c0 <= c1) || (m == 0 && c0 > c1)) out[0] = c1 | (c0 << 16);
	else out[0] = c0 | (c1 << 16);
	uint32_t bits = 0;
	for (int i = 0; i < 16; i++) bits |= (uint32_t) idx[i] << (2 * i);
	out[1] = bits;
}

// Image_CompressAMDBC1Block (:51-105) with adaptiveColourWeights = threeDRefinement = false
B1_HD void encode_block(const float in[64], float alpha_threshold, int steps, uint32_t out[2]) {
	uint8_t ep[2][3][2], idx[2][16];
	const bool use_alpha = alpha_threshold > 0.0f;
	const double e3 = (double) compress(in, 3, use_alpha, alpha_threshold, steps, ep[0], idx[0]);
	const double e4 = (e3 == 0.0) ? (double) FLT_MAX : (double) compress(in, 4, use_alpha, alpha_threshold, steps, ep[1], idx[1]);
	const int m = (e3 <= e4) ? 0 : 1;
	const uint32_t c0 = ((uint32_t) (ep[m][2][0] >> 3) << 11) | ((uint32_t) (ep[m][1][0] >> 2) << 5) | (uint32_t) (ep[m][0][0] >> 3);
	const uint32_t c1 = ((uint32_t) (ep[m][2][1] >> 3) << 11) | ((uint32_t) (ep[m][1][1] >> 2) << 5) | (uint32_t) (ep[m][0][1] >> 3);
	if ((m == 1 && c0 <= c1) || (m == 0 && c0 > c1)) out[0] = c1 | (c0 << 16);
	else out[0] = c0 | (c1 << 16);
	uint32_t bits = 0;
	for (int i = 0; i < 16; i++) bits |= (uint32_t) idx[m][i] << (2 * i);
	out[1] = bits;
}

} // namespace bc1
} // namespace b200ic
