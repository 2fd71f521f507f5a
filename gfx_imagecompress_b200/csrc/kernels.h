// kernels.h -- launchers of the per-codec sm_100a kernels (one translation unit per codec so that the
// bit-exact ones can be built with --fmad=false independently of the tolerance-gated ones).
#pragma once
#include "common.cuh"

namespace b200ic {

cudaError_t launch_bc45(const SrcImage &img, int channels, int first_channel, void *dst, cudaStream_t stream, int dst_stride = 8);
// BC2 / BC3: kBc3Colour = BC3 (alpha half by the BC4 kernel, colour half here), kBc2Both = BC2 (4-bit alpha + colour);
// block API: kColourOnly / kAlphaOnly write one 8-byte half per block
enum Bc23Part { kBc3Colour = 0, kBc2Both = 1, kColourOnly = 2, kAlphaOnly = 3 };
cudaError_t launch_bc23(const SrcImage &img, const b200ic_opts &opts, int part, void *dst, cudaStream_t stream);
cudaError_t launch_bc1(const SrcImage &img, const b200ic_opts &opts, void *dst, cudaStream_t stream);
cudaError_t launch_bc7rg(const SrcImage &img, const b200ic_opts &opts, void *dst, cudaStream_t stream);
cudaError_t launch_bc7amd(const SrcImage &img, const b200ic_opts &opts, void *dst, cudaStream_t stream);
cudaError_t launch_bc6h(const SrcImage &img, const b200ic_opts &opts, void *dst, cudaStream_t stream);
cudaError_t launch_box_mip_rgba8(const void *src, uint32_t sw, uint32_t sh, uint64_t spitch, void *dst, uint64_t dpitch, cudaStream_t stream);
int write_dds(const char *path, int codec, int srgb, int is_signed, uint32_t width, uint32_t height, uint32_t levels, const void *const *level_blocks);
cudaError_t launch_decode(int codec, const void *blocks, uint32_t width, uint32_t height, uint32_t slices, int is_signed, void *dst, uint64_t row_pitch,
													cudaStream_t stream);
void count_launches(int extra); // launchers issuing more than one kernel per encode report the extra ones
cudaError_t init_bc7rg_tables();
cudaError_t init_bc7amd_tables();
cudaError_t init_bc6h_tables();

} // namespace b200ic
