// Instantiations of k_iir_tma<..., PROG = true> (fused input / epilogue programs, WARM mode).
// Compiled once per pair of section counts (-DSIGOPS_PROG_GROUP=g covers M = 2g+1, 2g+2): the
// embedded interpreter makes each instantiation slow to build.  No b0 = b2 = 1 specialisation:
// the programs dominate these kernels.
#include "common.h"
#include "launchers.h"

#ifndef SIGOPS_PROG_GROUP
#error "compile with -DSIGOPS_PROG_GROUP=0..3"
#endif

#define SIGOPS_CAT2(a, b) a##b
#define SIGOPS_CAT(a, b) SIGOPS_CAT2(a, b)

namespace sigops {
namespace {
constexpr size_t kTmaSmemBytes = (size_t)kTmaThreads * 2 * kStagePitch * sizeof(double) + (size_t)kTmaThreads * 2 * sizeof(uint64_t);

template <int M>
void launch_one(dim3 grid, cudaStream_t st, const IirTmaParams& Q) {
    ensure_dyn_smem(k_iir_tma<M, IIR_WARM, false, true>, kTmaSmemBytes);
    k_iir_tma<M, IIR_WARM, false, true><<<grid, kTmaThreads, kTmaSmemBytes, st>>>(Q);
    CUDA_OK(cudaGetLastError());
}
}  // namespace

void SIGOPS_CAT(launch_iir_tma_prog_g, SIGOPS_PROG_GROUP)(int M, dim3 grid, cudaStream_t st, const IirTmaParams& Q) {
    if (M == 2 * SIGOPS_PROG_GROUP + 1) launch_one<2 * SIGOPS_PROG_GROUP + 1>(grid, st, Q);
    else launch_one<2 * SIGOPS_PROG_GROUP + 2>(grid, st, Q);
}

}  // namespace sigops
