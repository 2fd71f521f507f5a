// bc45.cu -- BC4 / BC5 scalar-channel encoder for sm_100a (bit-exact with the reference).
//
// Replaces: Image_CompressAMDAlphaSingleModeBlock (reference src/amd_bcx_helpers.cpp:125-140),
// CompBlock1X / CompBlock1 / RmpSrch1 / Refine1 / Clstr1 / GetRmp1 / BldRmp1
// (src/amd_bcx_body.cpp:1395-1868), EncodeAlphaBlock (src/amd_bcx_helpers.cpp:32-46) and the image loops
// of src/amd_bc4_compressor.cpp:27-50 / src/amd_bc5_compressor.cpp:27-54 with the gather of
// src/block_utils.cpp:116-144.
//
// Mapping (candidates -> lanes):
//   work item  = one 8-byte channel block (BC4: 1 per 4x4 block, BC5: 2)
//   half-warp  = one work item; its two 8-lane groups run the reference's two independent fits in lockstep:
//                group 0 = 8-point ramp, group 1 = 6-point ramp with fixed 0/255 (the reference's
//                `fError8 == 0 ? FLT_MAX` skip only saves time, it never changes the winner).
//   8 lanes    = the 8 non-trivial moves of Refine1's 3x3 hill climb, or 8 (step_l, step_r) candidates per
//                round of the global grid search; winner = lexicographic (error, scan index) arg-min by
//                warp shuffles, which reproduces the reference's "first strict minimum in scan order".
//   RmpSrch1's early-out returns exactly `_maxerror`, which can never win a strict `<`, and its partial
//   sums are monotone (non-negative terms), so evaluating every candidate fully is output-identical.
// Bit-exactness: compiled with --fmad=false, IEEE div, every accumulation kept in reference order.
#include "common.cuh"
#include <float.h>

namespace b200ic {

namespace {

constexpr float kMaxError = 128000.f; // MAX_ERROR        src/amd_bcx_body.cpp:43
constexpr float kGblStep = 0.018f;    // GBL_SCH_STEP_MXS :47
constexpr float kGblExt = 0.1f;       // GBL_SCH_EXT_MXS  :48
constexpr float kLclStep = 0.6f;      // LCL_SCH_STEP_MXS :49
constexpr int kWarpsPerCta = 4;

struct Bc45Params {
	SrcImage img;
	uint8_t *dst;
	uint64_t n_items;  // channel blocks
	int32_t channels;  // 1 (BC4) or 2 (BC5)
	int32_t first_channel;
	int32_t dst_stride; // bytes from one channel block to the next (8; 16 when the blocks are the alpha halves of BC3)
};

// RmpSrch1 (src/amd_bcx_body.cpp:1510-1548) without the early-out; ur[i] = (unique value, repeat count)
__device__ __forceinline__ float ramp_error(const float2 *ur, int n, float lo, float hi, int npoints) {
	const float step = (hi - lo) / (float) (npoints - 1);
	const float step_h = step * 0.5f;
	const float rstep = 1.0f / step;
	float error = 0.f;
	for (int i = 0; i < n; i++) {
		const float2 u = ur[i];
		const float del = u.x - lo;
		float q;
		if (del <= 0.f) q = lo;
		else if (u.x - hi >= 0.f) q = hi;
		else q = floorf((del + step_h) * rstep) * step + lo;
		const float d = u.x - q;
		error += d * d * u.y;
	}
	return error;
}

__device__ __forceinline__ float refine_move(int k) { return k == 0 ? 0.f : (k == 1 ? -1.f : 1.f); } // sMvF :580

__global__ void __launch_bounds__(kWarpsPerCta * 32) bc45_kernel(const Bc45Params p) {
	__shared__ float2 s_ur[kWarpsPerCta][4][17]; // [warp][half*2+pass][unique] (+1 pad: the 4 groups hit distinct banks)
	__shared__ float s_sorted[kWarpsPerCta][2][16];

	const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
	const unsigned hw = lane >> 4, l16 = lane & 15u, grp = lane >> 3, pass = grp & 1u, l8 = lane & 7u;
	const unsigned hw_base = lane & 16u, grp_base = lane & 24u;

	uint64_t item = ((uint64_t) blockIdx.x * kWarpsPerCta + warp) * 2 + hw;
	const bool valid = item < p.n_items;
	if (!valid) item = p.n_items - 1;
	const uint64_t block = item / (uint32_t) p.channels;
	const int csel = (int) (item - block * (uint32_t) p.channels);
	const uint64_t per_slice = (uint64_t) p.img.blocks_x * p.img.blocks_y;
	const uint32_t slice = (uint32_t) (block / per_slice);
	const uint32_t rem = (uint32_t) (block - (uint64_t) slice * per_slice);
	const uint32_t by = rem / p.img.blocks_x, bx = rem - by * p.img.blocks_x;

	// ---- gather: lane l16 owns texel l16 of this half-warp's channel block
	const float v = fetch_channel(p.img, block, bx, by, slice, (int) l16, p.first_channel + csel);

	// ---- sort (qsort + QSortFCmp, src/amd_bcx_body.cpp:1609-1618,1654): rank sort over 16 lanes
	int rank = 0;
#pragma unroll
	for (int j = 0; j < 16; j++) {
		const float vj = __shfl_sync(FULL, v, hw_base | j);
		rank += (vj < v || (vj == v && j < (int) l16)) ? 1 : 0;
	}
	s_sorted[warp][hw][rank] = v;
	__syncwarp();
	const float s = s_sorted[warp][hw][l16];
	const float sprev = __shfl_up_sync(FULL, s, 1);
	const bool first = (l16 == 0) || (s != sprev);
	// fixed-ramp pass drops values at the 0 / 1 ends (double-precision compares, :1671-1674)
	const bool dropped = ((double) s <= 1.5 / 255.) || ((double) s >= 253.5 / 255.);
	const bool keep = first && !dropped;
	const unsigned fm = (__ballot_sync(FULL, first) >> hw_base) & 0xffffu;
	const unsigned km = (__ballot_sync(FULL, keep) >> hw_base) & 0xffffu;
	{
		const unsigned above = fm >> (l16 + 1);
		const int run = above ? __ffs(above) : (int) (16 - l16);
		const unsigned below = (1u << l16) - 1u;
		if (first) s_ur[warp][hw * 2 + 0][__popc(fm & below)] = make_float2(s, (float) run);
		if (keep) s_ur[warp][hw * 2 + 1][__popc(km & below)] = make_float2(s, (float) run);
	}
	__syncwarp();

	const int n = __popc(pass ? km : fm);
	const int npoints = pass ? 6 : 8;
	const float2 *ur = s_ur[warp][grp];

	// ---- CompBlock1 (:1633-1832) with _IntPrc=8, _FracPrc=0, _bFixedRamp=true
	float ramp0, ramp1;
	const bool need = n > 2;
	const float uv0 = n > 0 ? ur[0].x : 0.f;
	if (!need) {
		if (n == 0) { // only reachable in the fixed pass (:1700-1704)
			ramp0 = 128.f;
			ramp1 = 129.f;
		} else {
			ramp0 = floorf(uv0 * 255.f + 0.5f);
			ramp1 = (n == 1) ? ramp0 + 1.f : floorf(ur[1].x * 255.f + 0.5f);
		}
	}
	float lo = uv0, hi = n > 0 ? ur[n - 1].x : 0.f;
	float maxerr = kMaxError;

	// global (step_l, step_r) grid search, only when the value range exceeds 48/256 (:1750-1779)
	const bool wants = need && !(hi - lo <= 48.f / 256.f);
	if (__any_sync(FULL, wants)) {
		float llb = 0.f, rrb = 0.f;
		int nl = 0, nr = 0;
		if (wants) {
			const float cntr = (lo + hi) / 2;
			llb = (0.f > lo - kGblExt) ? 0.f : lo - kGblExt;
			rrb = (1.f < hi + kGblExt) ? 1.f : hi + kGblExt;
			const float lrb = (cntr < lo + kGblExt) ? cntr : lo + kGblExt;
			const float rlb = (cntr > hi - kGblExt) ? cntr : hi - kGblExt;
			for (float sl = llb; sl < lrb && nl < 64; sl += kGblStep) nl++;
			for (float sr = rrb; rlb <= sr && nr < 64; sr -= kGblStep) nr++;
		}
		const int ncand = nl * nr;
		const int maxc = __reduce_max_sync(FULL, ncand);
		float gl = 0.f, gr = 0.f;
		for (int c0 = 0; c0 < maxc; c0 += 8) {
			const int c = c0 + (int) l8;
			float e = INFINITY;
			if (c < ncand) {
				const int il = c / nr, ir = c - il * nr;
				float sl = llb, sr = rrb;
				for (int k = 0; k < il; k++) sl += kGblStep; // the reference's loop variables accumulate
				for (int k = 0; k < ir; k++) sr -= kGblStep;
				e = ramp_error(ur, n, sl, sr, npoints);
			}
			int ci = c;
			group_argmin<8>(e, ci);
			if (ci < ncand && e < maxerr) {
				maxerr = e;
				const int il = ci / nr, ir = ci - il * nr;
				float sl = llb, sr = rrb;
				for (int k = 0; k < il; k++) sl += kGblStep;
				for (int k = 0; k < ir; k++) sr -= kGblStep;
				gl = sl;
				gr = sr;
			}
		}
		if (wants) {
			lo = gl;
			hi = gr;
		}
	}

	// Refine1 (:1555-1607): 3x3 hill climb; lane l8 evaluates move (l8+1); move 0 can only win in round 1
	{
		const float m_step = kLclStep / 256.f;
		bool active = need;
		bool first_round = true;
		while (__any_sync(FULL, active)) {
			bool improved = false;
			float nlo = lo, nhi = hi;
			if (first_round && active) {
				const float cl = Math_MaxF_dev(lo + m_step * 0.f, 0.f);
				const float ch = Math_MinF_dev(hi + m_step * 0.f, 1.f);
				const float e0 = ramp_error(ur, n, cl, ch, npoints);
				if (e0 < maxerr) {
					maxerr = e0;
					nlo = cl;
					nhi = ch;
					improved = true;
				}
			}
			int mode = (int) l8 + 1;
			float e = INFINITY;
			if (active) {
				const float cl = Math_MaxF_dev(lo + m_step * refine_move(mode / 3), 0.f);
				const float ch = Math_MinF_dev(hi + m_step * refine_move(mode % 3), 1.f);
				e = ramp_error(ur, n, cl, ch, npoints);
			}
			group_argmin<8>(e, mode);
			if (active && e < maxerr) {
				maxerr = e;
				nlo = Math_MaxF_dev(lo + m_step * refine_move(mode / 3), 0.f);
				nhi = Math_MinF_dev(hi + m_step * refine_move(mode % 3), 1.f);
				improved = true;
			}
			lo = nlo;
			hi = nhi;
			active = active && improved;
			first_round = false;
		}
	}
	if (need) {
		ramp1 = floorf(hi * 255.f + 0.5f);
		ramp0 = floorf(lo * 255.f + 0.5f);
	}
	if (ramp0 == ramp1) { // :1821-1827
		if (ramp1 < 255.f) ramp1 += 1.f;
		else ramp1 -= 1.f;
	}

	// ---- Clstr1 / GetRmp1 / BldRmp1 (:1395-1505): final index assignment on the 8.0 integer grid
	if ((!pass && ramp0 <= ramp1) || (pass && ramp0 > ramp1)) {
		const float t = ramp0;
		ramp0 = ramp1;
		ramp1 = t;
	}
	float alpha[8];
	{
		// 8 table entries either way: an 8-point ramp, or a 6-point ramp + the fixed {0, 255}
		alpha[0] = ramp0;
		alpha[1] = ramp1;
#pragma unroll
		for (int e = 1; e < 7; e++)
			if (e < npoints - 1) alpha[e + 1] = (ramp0 * (float) (npoints - 1 - e) + ramp1 * (float) e) / (float) (npoints - 1);
		if (pass) {
			alpha[6] = 0.f;
			alpha[7] = 255.f;
		}
		const float over = 1.f / (256.f - 1.f);
#pragma unroll
		for (int i = 0; i < 8; i++) {
			if (i < npoints) alpha[i] = floorf(alpha[i] + 0.5f);
			alpha[i] *= over;
		}
	}
	const float va = __shfl_sync(FULL, v, hw_base | l8);
	const float vb = __shfl_sync(FULL, v, hw_base | (l8 + 8));
	float da = 10000000.f, db = 10000000.f;
	uint32_t ia = 0, ib = 0;
#pragma unroll
	for (int j = 0; j < 8; j++) {
		float d = va - alpha[j];
		d *= d;
		if (d < da) { da = d; ia = j; }
		d = vb - alpha[j];
		d *= d;
		if (d < db) { db = d; ib = j; }
	}
	float err = 0.f;
#pragma unroll
	for (int i = 0; i < 16; i++) {
		const float t = __shfl_sync(FULL, i < 8 ? da : db, grp_base | (i & 7));
		err += t;
	}

	// ---- Image_CompressAMDAlphaSingleModeBlock's choice + EncodeAlphaBlock
	const float e8 = __shfl_sync(FULL, err, hw_base);
	const float e6 = __shfl_sync(FULL, err, hw_base | 8u);
	const bool use8 = (e8 == 0.f) || (e8 <= e6);
	uint64_t bits = 0;
	if ((pass == 0) == use8) {
		bits = ((uint64_t) ia << (16 + 3 * l8)) | ((uint64_t) ib << (16 + 3 * (l8 + 8)));
		if (l8 == 0) bits |= (uint64_t) ((uint32_t) ramp0 & 0xffu) | ((uint64_t) ((uint32_t) ramp1 & 0xffu) << 8);
	}
	uint32_t w0 = (uint32_t) bits, w1 = (uint32_t) (bits >> 32);
#pragma unroll
	for (int d = 8; d > 0; d >>= 1) {
		w0 |= __shfl_xor_sync(FULL, w0, d);
		w1 |= __shfl_xor_sync(FULL, w1, d);
	}
	if (valid && l16 == 0) *reinterpret_cast<uint2 *>(p.dst + item * (uint64_t) p.dst_stride) = make_uint2(w0, w1);
}

} // namespace

cudaError_t launch_bc45(const SrcImage &img, int channels, int first_channel, void *dst, cudaStream_t stream, int dst_stride) {
	Bc45Params p;
	p.img = img;
	p.dst = static_cast<uint8_t *>(dst);
	p.channels = channels;
	p.first_channel = first_channel;
	p.dst_stride = dst_stride;
	p.n_items = (uint64_t) img.blocks_x * img.blocks_y * img.slices * (uint64_t) channels;
	if (p.n_items == 0) return cudaSuccess;
	const uint64_t per_cta = (uint64_t) kWarpsPerCta * 2;
	const uint64_t grid = (p.n_items + per_cta - 1) / per_cta;
	bc45_kernel<<<(unsigned) grid, kWarpsPerCta * 32, 0, stream>>>(p);
	return cudaGetLastError();
}

} // namespace b200ic
