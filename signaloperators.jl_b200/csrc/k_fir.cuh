// K4/K5 — polyphase FIR resampling (ToFramerate on data signals).
//
// Replaces DSP.jl's `filt!(out, FIRFilter{FIRArbitrary|FIRRational|FIRDecimator|
// FIRInterpolator}, x)` reached from src/filters.jl:252-255 with the filter
// built at src/reformatting.jl:92-99 (kernels restated in SURVEY.md App. B.4).
// One-shot semantics: the input is the child followed by
// zeros and the first n_out outputs of the stream are kept.
//
// For every output m the host replayed the kernel's own index recurrence once
// (bit-exact Float64 phase accumulator) into two tables shared by every channel
// and instance of the batch:
//     xi0[m]  0-based index of the newest input sample in the window
//     phi[m]  FIRArbitrary phase accumulator (phiIdx = floor, alpha = frac) or
//             the integer phase of the rational kernels
// so  y[m] = sum_t (pfb[phiIdx][t] + alpha*dpfb[phiIdx][t]) * x[xi0 - tapsPerPhi + 1 + t].
//
// Block = tile of T outputs x RB rows (a row = one channel of one instance).
// The taps of the tile's outputs are merged once into shared memory and reused
// by all RB rows; the input window of the tile is staged transposed
// ([position][row]) so lanes = rows read it conflict-free; each thread keeps an
// 8-output x 4-row accumulator tile in registers, so one broadcast tap load
// feeds 4 FMAs and one sample load feeds 8.
//
// Roofline (44.1k -> 48k): 15.35 B per output sample of HBM traffic, 38 FP64
// FMAs per sample after merging.
#pragma once
#include "interp.cuh"

namespace sigops {

constexpr int kFirWarps = 8;
constexpr int kFirThreads = kFirWarps * 32;
constexpr int kFirR = 8;                    // outputs per thread
constexpr int kFirT = kFirWarps * kFirR;    // outputs per tile (64)
// rows per thread G in {1,2,4}: rows per block RB = 32*G, smem row pitch RB+1

struct FirParams {
    const sigops_instr* instrs;
    const BufRef* bufrefs;
    double* scalars;
    int nbuf, nscalars;
    int out_buf, sumsq_slot;
    int in_buf;                // input is always a materialised buffer, zero padded past in_len
    int64_t in_len;
    int epi_prog_start, epi_prog_len;
    int nch;
    int64_t nrows;
    int64_t n_out;
    int tapsper, dpad, tpad;   // tpad = tapsper + 2*dpad
    int pmax;                  // max window positions of any tile
    const double* pfb;         // [nphases][tapsper]
    const double* dpfb;        // or nullptr
    const int64_t* xi0;        // [ceil(n_out/T)*T], tail repeats the last entry
    const double* phi;
};

template <int kFirG>
__global__ void __launch_bounds__(kFirThreads)
k_fir(const __grid_constant__ FirParams P) {
    constexpr int kFirRB = 32 * kFirG;
    constexpr int kFirRowPitch = kFirRB + 1;
    __shared__ sigops_instr sprog_epi[SIGOPS_MAX_PROG];
    __shared__ double lc_epi[kFirWarps][SIGOPS_MAX_PROG];
    __shared__ int64_t s_xi0[kFirT];
    extern __shared__ double smem[];
    double* hm = smem;                                   // [T][tpad]
    double* xs = smem + (size_t)kFirT * P.tpad;          // [pmax][RB+1], later ys[T][RB+1]
    double* stack = xs + (size_t)(P.pmax > kFirT ? P.pmax : kFirT) * kFirRowPitch;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t m0 = (int64_t)blockIdx.x * kFirT;
    const int64_t row0 = (int64_t)blockIdx.y * kFirRB;

    for (int i = threadIdx.x; i < P.epi_prog_len; i += blockDim.x) sprog_epi[i] = P.instrs[P.epi_prog_start + i];
    if (threadIdx.x < kFirT) s_xi0[threadIdx.x] = P.xi0[m0 + threadIdx.x];
    // merged taps (zero outside [0,tapsper) so shifted windows need no predicate)
    for (int i = threadIdx.x; i < kFirT * P.tpad; i += blockDim.x) {
        const int m = i / P.tpad, t = i % P.tpad - P.dpad;
        double h = 0.0;
        if (t >= 0 && t < P.tapsper && m0 + m < P.n_out) {
            const double acc = P.phi[m0 + m];
            const double fl = floor(acc);
            const int64_t ph = (int64_t)fl - 1;
            h = P.pfb[ph * P.tapsper + t];
            if (P.dpfb) h = fma(acc - fl, P.dpfb[ph * P.tapsper + t], h);
        }
        hm[i] = h;
    }
    __syncthreads();

    // ---- stage the input window, transposed to [position][row]
    const int64_t p0 = s_xi0[0] - P.tapsper + 1;
    const int npos = (int)(s_xi0[kFirT - 1] - p0 + 1);
    for (int r = warp; r < kFirRB; r += kFirWarps) {
        const int64_t row = row0 + r;
        const bool live = row < P.nrows;
        BufRef ib{};
        int c = 0;
        if (live) {
            const int64_t inst = row / P.nch;
            c = (int)(row % P.nch);
            ib = P.bufrefs[(size_t)inst * P.nbuf + P.in_buf];
        }
        for (int j = lane; j < npos; j += 32) {
            const int64_t p = p0 + j;
            double v = 0.0;
            if (live && p >= 0 && p < P.in_len) v = load_elem(ib.ptr, ib.dtype, (int64_t)c * ib.ld + p);
            xs[(size_t)j * kFirRowPitch + r] = v;
        }
    }
    __syncthreads();

    // ---- register-tiled dot products: 8 outputs x 4 rows per thread
    double acc[kFirR][kFirG];
#pragma unroll
    for (int r = 0; r < kFirR; ++r)
#pragma unroll
        for (int g = 0; g < kFirG; ++g) acc[r][g] = 0.0;
    {
        const int mw = warp * kFirR;
        const int64_t xw0 = s_xi0[mw];
        const double* hp[kFirR];
#pragma unroll
        for (int r = 0; r < kFirR; ++r)
            hp[r] = hm + (size_t)(mw + r) * P.tpad + P.dpad - (int)(s_xi0[mw + r] - xw0);
        const int jn = P.tapsper + (int)(s_xi0[mw + kFirR - 1] - xw0);
        const double* xp = xs + (size_t)(xw0 - P.tapsper + 1 - p0) * kFirRowPitch + lane;
        for (int j = 0; j < jn; ++j) {
            double xv[kFirG];
#pragma unroll
            for (int g = 0; g < kFirG; ++g) xv[g] = xp[(size_t)j * kFirRowPitch + 32 * g];
#pragma unroll
            for (int r = 0; r < kFirR; ++r) {
                const double hv = hp[r][j];
#pragma unroll
                for (int g = 0; g < kFirG; ++g) acc[r][g] = fma(hv, xv[g], acc[r][g]);
            }
        }
    }
    __syncthreads();

    // ---- stage results as ys[output][row], then store rows coalesced along time
    double* ys = xs;
#pragma unroll
    for (int r = 0; r < kFirR; ++r)
#pragma unroll
        for (int g = 0; g < kFirG; ++g)
            ys[(size_t)(warp * kFirR + r) * kFirRowPitch + lane + 32 * g] = acc[r][g];
    __syncthreads();

    for (int r = warp; r < kFirRB; r += kFirWarps) {
        const int64_t row = row0 + r;
        if (row >= P.nrows) break;
        const int64_t inst = row / P.nch;
        const int c = (int)(row % P.nch);
        const BufRef* bufs = P.bufrefs + (size_t)inst * P.nbuf;
        Env env{bufs, P.scalars + (size_t)inst * P.nscalars};
        const BufRef ob = bufs[P.out_buf];
        double y[2], o[2];
        y[0] = ys[(size_t)lane * kFirRowPitch + r];
        y[1] = ys[(size_t)(lane + 32) * kFirRowPitch + r];
        if (P.epi_prog_len > 0) {
            for (int i = lane; i < P.epi_prog_len; i += 32) {
                const sigops_instr& I = sprog_epi[i];
                double v = 0.0;
                if (I.leaf == SIGOPS_LEAF_CONST) v = I.d0;
                else if (I.leaf == SIGOPS_LEAF_RMS) v = sqrt(env.scalars[I.buf] / I.d0);
                lc_epi[warp][i] = v;
            }
            __syncwarp();
            eval_program<2>(sprog_epi, lc_epi[warp], P.epi_prog_len, env, m0 + lane, 32, c, y, o,
                            stack + threadIdx.x, kFirThreads);
            __syncwarp();
        } else {
            o[0] = y[0]; o[1] = y[1];
        }
        double ss = 0.0;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int64_t m = m0 + lane + 32 * j;
            if (m < P.n_out) {
                const double w = store_elem(ob.ptr, ob.dtype, (int64_t)c * ob.ld + m, o[j]);
                ss += w * w;
            }
        }
        if (P.sumsq_slot >= 0) {
            ss = warp_sum(ss);
            if (lane == 0) atomicAdd(P.scalars + (size_t)inst * P.nscalars + P.sumsq_slot, ss);
        }
    }
}

}  // namespace sigops
