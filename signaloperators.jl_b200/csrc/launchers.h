// Kernel launchers compiled in their own translation units (the IIR kernels are
// instantiated for 1..8 sections x several modes; separate objects build in parallel).
#pragma once
#include <cuda_runtime.h>

#include "k_iir.cuh"
#include "k_iir_tma.cuh"

namespace sigops {

enum { LAUNCH_MAIN = 0, LAUNCH_FIX = 1, LAUNCH_WARM = 2 };

// generic program-carrying kernel (k_iir)
void launch_iir_generic(int mode, int M, dim3 grid, size_t smem, cudaStream_t st, const IirParams& P);
// cp.async fast path for unaligned Float64 buffers (k_iir_fast)
void launch_iir_cpasync(int mode, int M, bool unitb, dim3 grid, cudaStream_t st, const IirParams& P);
// TMA bulk-copy kernel (k_iir_tma); `prog` selects the variant with fused programs (WARM only)
void launch_iir_tma_any(int mode, bool prog, int M, bool unitb, dim3 grid, cudaStream_t st, const IirTmaParams& Q);

}  // namespace sigops
