// stubs.cu -- launchers of codecs that are not built yet fail loudly (never a CPU fallback).
#include "kernels.h"
namespace b200ic {
#ifndef B200IC_HAVE_BC1
cudaError_t launch_bc1(const SrcImage &, const b200ic_opts &, void *, cudaStream_t) { return cudaErrorNotSupported; }
#endif
#ifndef B200IC_HAVE_BC7RG
cudaError_t launch_bc7rg(const SrcImage &, const b200ic_opts &, void *, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t init_bc7rg_tables() { return cudaSuccess; }
#endif
#ifndef B200IC_HAVE_BC7AMD
cudaError_t launch_bc7amd(const SrcImage &, const b200ic_opts &, void *, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t init_bc7amd_tables() { return cudaSuccess; }
#endif
#ifndef B200IC_HAVE_BC6H
cudaError_t launch_bc6h(const SrcImage &, const b200ic_opts &, void *, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t init_bc6h_tables() { return cudaSuccess; }
#endif
} // namespace b200ic
