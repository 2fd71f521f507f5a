// bc7amd_block.cuh -- block-level orchestration of the AMD BC7 search in serial form (host build / restatement);
// the CUDA kernel (bc7amd.cu) spreads the same tasks over lanes. U8 = true selects the exact integer shakers of
// bc7amd_int.cuh (valid for 8-bit sources), U8 = false the generic FP64 ones of bc7amd_core.cuh.
#pragma once
#include "bc7amd_core.cuh"
#include "bc7amd_int.cuh"

namespace b200ic {
namespace amd7 {

// Shake one subset of one partition (:709-805). idx[] in (quantiser indices) / out. Returns the subset's error.
A7_HD real shake_subset(const Tables &T, const ShakeParams &sp, const real data[][4], int n, int *idx, int ep[2][4]) {
	if (sp.dim != 3) return shake_window(T, data, n, idx, ep, sp.shake_size, sp.clusters - 1, sp.bits[3], sp.dim);
	int tmp[kMaxEntries];
	for (int k = 0; k < n; k++) tmp[k] = idx[k];
	const real e0 = shake_cube(T, data, n, tmp, sp.clusters - 1, sp.bits, sp.parity);
	real e1 = shake_window(T, data, n, idx, ep, sp.shake_size, sp.clusters - 1, sp.bits[3], sp.dim);
	if (e0 < e1) {
		e1 = shake_window(T, data, n, tmp, ep, sp.shake_size, sp.clusters - 1, sp.bits[3], sp.dim);
		for (int k = 0; k < n; k++) idx[k] = tmp[k];
	}
	return e1;
}

// Serial-form helper: quantise subset `s` of `partition` (or, subsets == 0, the whole block with component j taken from
// channel (chan >> 2j) & 3) with plain local work arrays.
A7_HD real quantise_serial(const BlockInput &B, int subsets, int partition, int subset, uint32_t chan, int dim, int clusters, int *idx, int &n) {
	real proj[kMaxEntries], dev[kMaxEntries];
	uint32_t mask = 0;
	for (int i = 0; i < 16; i++)
		if (subsets == 0 || subset_of(subsets, partition, i) == subset) mask |= 1u << i;
	QuantIO io;
	io.px = &B.pxc[0][0];
	io.texels = texels_of_mask(mask, n);
	io.chan = chan;
	io.proj = proj;
	io.dev = dev;
	io.stride = 1;
	uint64_t packed = 0;
	const real e = n ? quantise_subset(io, n, clusters, dim, packed) : 0;
	for (int k = 0; k < n; k++) idx[k] = (int) ((packed >> (4 * k)) & 15u);
	return e;
}
constexpr uint32_t kChanRGBA = 0xE4u; // component j <- channel j

// CompressSingleIndexBlock (:548-890)
template <bool U8> A7_HDN real compress_single_index(const Tables &T, const BlockInput &B, int mode, uint64_t out[2]) {
	const ModeInfo mi = mode_info(mode);
	const ShakeParams sp = single_index_shake_params(mode);
	const int nparts = 1 << mi.partition_bits;
	real perr[64];
	for (int part = 0; part < nparts; part++) {
		real e = 0;
		for (int s = 0; s < mi.subsets; s++) {
			int n, idx[kMaxEntries];
			const real es = quantise_serial(B, mi.subsets, part, s, kChanRGBA, sp.dim, sp.clusters, idx, n);
			if (n) e += es;
		}
		perr[part] = e;
	}
	int order[64];
	sort_order(perr, order, nparts);
	const int attempts = nparts < 8 ? nparts : 8;
	real best = A7_HUGE;
	SingleIndexResult res, cur;
	res.partition = 0;
	for (int a = 0; a < attempts; a++) {
		const int part = order[a];
		real e = 0;
		cur.partition = part;
		for (int s = 0; s < mi.subsets; s++) {
			real sub[kMaxEntries][4];
			int n;
			gather_subset(B, mi.subsets, part, s, sp.dim, sub, n);
			if (!n) continue;
			quantise_serial(B, mi.subsets, part, s, kChanRGBA, sp.dim, sp.clusters, cur.idx[s], n); // the reference stored these in the first pass
			for (int k = 0; k < 4; k++) cur.ep[s][0][k] = cur.ep[s][1][k] = 0;
			if (U8) {
				U8Subset S;
				make_u8_subset(sub, n, sp.dim, S);
				e += shake_subset_u8(T, sp, S, cur.idx[s], cur.ep[s]);
			} else {
				e += shake_subset(T, sp, sub, n, cur.idx[s], cur.ep[s]);
			}
		}
		if (e < best) { best = e; res = cur; }
	}
	pack_single_index(mode, res, out);
	return best;
}

// CompressDualIndexBlock (:1059-1278) at quality 1: every rotation x index selection is quantised and shaken
A7_HD int rotation_channel(int rotation, int slot) { // componentRotations (:894-900)
	const int t[4][4] = {{3, 0, 1, 2}, {0, 3, 1, 2}, {1, 0, 3, 2}, {2, 0, 1, 3}};
	return t[rotation][slot];
}
struct DualCombo {
	real err;
	int ep[2][2][4];
	int idx[2][16];
};
template <bool U8> A7_HD void dual_index_combo(const Tables &T, const BlockInput &B, int mode, int rotation, int isel, DualCombo &r) {
	const ModeInfo mi = mode_info(mode);
	real cb[16][4], ab[16][4];
	for (int i = 0; i < 16; i++) {
		cb[i][0] = B.px[i][rotation_channel(rotation, 1)];
		cb[i][1] = B.px[i][rotation_channel(rotation, 2)];
		cb[i][2] = B.px[i][rotation_channel(rotation, 3)];
		cb[i][3] = 0;
		ab[i][0] = ab[i][1] = ab[i][2] = B.px[i][rotation_channel(rotation, 0)];
		ab[i][3] = 0;
	}
	const int ib[2] = {mi.index_bits0, mi.index_bits1};
	const int vb = mi.vector_bits / 3, sb = mi.scalar_bits;
	int n16;
	const uint32_t c0 = (uint32_t) rotation_channel(rotation, 0), c1 = (uint32_t) rotation_channel(rotation, 1),
								 c2 = (uint32_t) rotation_channel(rotation, 2), c3 = (uint32_t) rotation_channel(rotation, 3);
	quantise_serial(B, 0, 0, 0, c1 | (c2 << 2) | (c3 << 4), 3, 1 << ib[isel], r.idx[0], n16);
	quantise_serial(B, 0, 0, 0, c0 | (c0 << 2) | (c0 << 4), 3, 1 << ib[1 ^ isel], r.idx[1], n16);
	const int bits0[4] = {vb, vb, vb, 6 * vb}, bits1[4] = {sb, sb, sb, 6 * sb};
	for (int i = 0; i < 2; i++)
		for (int e = 0; e < 2; e++)
			for (int k = 0; k < 4; k++) r.ep[i][e][k] = 0;
	real e = 0;
	if (U8) {
		U8Subset S;
		make_u8_subset(cb, 16, 3, S);
		shake_cube_u8_any(T, S, r.idx[0], ib[isel], bits0, CART);
		e += shake_window_u8_any(T, S, r.idx[0], r.ep[0], 6, ib[isel], bits0[3], 3);
		make_u8_subset(ab, 16, 3, S);
		shake_cube_u8_any(T, S, r.idx[1], ib[1 ^ isel], bits1, CART);
		e += shake_window_u8_any(T, S, r.idx[1], r.ep[1], 6, ib[1 ^ isel], bits1[3], 3) / 3.;
	} else {
		shake_cube(T, cb, 16, r.idx[0], (1 << ib[isel]) - 1, bits0, CART);
		e += shake_window(T, cb, 16, r.idx[0], r.ep[0], 6, (1 << ib[isel]) - 1, bits0[3], 3);
		shake_cube(T, ab, 16, r.idx[1], (1 << ib[1 ^ isel]) - 1, bits1, CART);
		e += shake_window(T, ab, 16, r.idx[1], r.ep[1], 6, (1 << ib[1 ^ isel]) - 1, bits1[3], 3) / 3.;
	}
	r.err = e;
}
template <bool U8> A7_HDN real compress_dual_index(const Tables &T, const BlockInput &B, int mode, uint64_t out[2]) {
	const ModeInfo mi = mode_info(mode);
	real best = A7_HUGE;
	for (int rot = 0; rot < (1 << mi.rotation_bits); rot++)
		for (int isel = 0; isel < (1 << mi.index_mode_bits); isel++) {
			DualCombo c;
			dual_index_combo<U8>(T, B, mode, rot, isel, c);
			if (c.err < best) {
				pack_dual_index(mode, isel, rot, c.ep, c.idx, out);
				best = c.err;
			}
		}
	return best;
}

// CompressBlock (:1289-1465): modes in the order {6,4,3,1,2,0,7,5}, first strict minimum wins
A7_HD int mode_visit_order(int i) { return (int) ((0x57021346u >> (4 * i)) & 15u); }
template <bool U8> A7_HD real encode_block_serial(const Tables &T, const float in[64], uint32_t valid_mode_mask, uint64_t out[2]) {
	BlockInput B;
	prepare_block(in, valid_mode_mask, B);
	real best = A7_HUGE;
	out[0] = out[1] = 0;
	for (int i = 0; i < 8; i++) {
		const int m = mode_visit_order(i);
		if (!(B.mode_mask & (1u << m))) continue;
		uint64_t tmp[2];
		const real e = (mode_info(m).alpha != 2) ? compress_single_index<U8>(T, B, m, tmp) : compress_dual_index<U8>(T, B, m, tmp);
		if (e < best) { best = e; out[0] = tmp[0]; out[1] = tmp[1]; }
	}
	return best;
}

} // namespace amd7
} // namespace b200ic
