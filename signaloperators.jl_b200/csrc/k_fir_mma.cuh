// K4/K5 (fast path) — polyphase FIR resampling on the FP64 tensor cores.
//
// Same arithmetic as k_fir.cuh (DSP.jl `filt!(out, FIRFilter{...}, x)` reached from
// src/filters.jl:252-255, filter built at src/reformatting.jl:92-99; the host replays the
// kernel's index recurrence into xi0[m], phi[m]):
//     y[m] = sum_t (pfb[phi][t] + alpha*dpfb[phi][t]) * x[xi0[m] - tapsPerPhi + 1 + t]
// restated for a group of 8 consecutive outputs as one small matrix product
//     Y[row][n] = sum_k X[row][q + k] * H[k][n],        k = 0 .. KS-1
// where q is the first window position of the group and H is the dense band of the eight
// outputs' merged taps (output n's taps start at row xi0[n] - xi0[0] of its column; zeros
// elsewhere).  `mma.sync.m8n8k4.f64` (DMMA) does 256 FMAs per warp instruction with one
// 64-bit operand load per lane, so the shared-memory pipe that bounds the FMA formulation
// (one broadcast tap load per G FMAs) is no longer the limit: measured on B200 the tensor
// pipe sustains 18.5 T FP64 FMA/s against 17.0 T for DFMA.
//
// The kernel is persistent along the time axis and warp-specialised:
//   * a block owns RB = 8*MF rows (a row = one channel of one instance) and a contiguous
//     range of 32-output tiles; the rows' input windows live in a shared-memory RING indexed
//     by position mod `ring`, so every input sample is fetched from L2/HBM once per block
//   * warps 4-7 (aux) prepare tile t+1 while tile t is multiplied.  Aux warp x owns a quarter
//     of the rows (lane = row) and extends their ring by the positions the next two tiles
//     need: one TMA bulk copy per row, two where the ring wraps (cp.async.bulk ...
//     mbarrier::complete_tx), zero fill by hand outside [0, in_len) (history before the
//     signal / the reference's zero padding).  It also builds the tap band of group x
//     (merging pfb + alpha*dpfb, both banks staged in shared memory when they fit)
//   * warps 0-3 (one per SM sub-partition, one 8-output group each) run the DMMA loop on
//     the current tile and store their fragments straight to global memory: a lane holds
//     two consecutive outputs of one row, so every 32-byte sector is written whole
//   * `full[t&1]` (every aux thread arrives once per tile, copies announced with expect_tx)
//     releases the compute warps, `done[t&1]` (compute warps) releases ring slots and the
//     tap buffer two tiles later.
//
// Eligibility (checked by the host): Float64 in/out, 16-byte aligned rows, no epilogue
// program (the sum of squares for a following Normpower is supported).  Everything else
// takes k_fir.cuh.
#pragma once
#include "interp.cuh"
#include "k_iir_tma.cuh"   // mbarrier / bulk-copy helpers

namespace sigops {

constexpr int kFmT = 32;           // outputs per tile: 4 compute warps x 8
constexpr int kFmThreads = 256;
constexpr int kFmHbPitch = 12;     // doubles per position row of a tap band (8 used): the k-major
                                   // B-fragment load (k = lane&3, n = lane>>2) is then conflict free
constexpr int kFmAuxThreads = 128;  // warps 4-7: ring loads + tap bands

struct FirMmaParams {
    const BufRef* bufrefs;
    double* scalars;
    int nbuf, nscalars;
    int in_buf, out_buf, sumsq_slot;
    int64_t in_len;
    int nch;
    int64_t nrows;
    int64_t n_out;
    int tapsper;
    int ks;                 // positions per 8-output group (multiple of 4)
    int ring;               // ring capacity in positions (even)
    int pitch;              // doubles per ring row: >= ring, = 4 mod 16 (conflict-free A-fragment loads)
    int64_t ntiles;         // ceil(n_out / 32)
    int64_t tiles_per_seg;  // tiles per block along x
    const double* pfb;      // [nphases][tapsper]
    const double* dpfb;     // or nullptr
    const int64_t* xi0;     // padded to a multiple of 64 entries
    const double* phi;
    int tab_doubles;        // nphases*tapsper when the banks are copied to shared memory, else 0
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MF>
__global__ void __launch_bounds__(kFmThreads, 1)
k_fir_mma(const __grid_constant__ FirMmaParams P) {
    constexpr int RB = 8 * MF;
    extern __shared__ __align__(128) unsigned char fm_smem[];
    double* ring = reinterpret_cast<double*>(fm_smem);                       // [RB][pitch]
    double* hb = ring + (size_t)RB * P.pitch;                                // [2][4][ks][kFmHbPitch]
    double* tabs = hb + (size_t)2 * 4 * P.ks * kFmHbPitch;                   // pfb, dpfb copies when they fit
    __shared__ uint64_t bar_full[2], bar_done[2];
    __shared__ const double* s_src[RB];
    __shared__ double* s_dst[RB];
    __shared__ int s_inst[RB];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t row0 = (int64_t)blockIdx.y * RB;
    const int64_t t0 = (int64_t)blockIdx.x * P.tiles_per_seg;
    const int64_t t1 = (t0 + P.tiles_per_seg < P.ntiles) ? t0 + P.tiles_per_seg : P.ntiles;
    if (t0 >= t1) return;

    if (tid < RB) {
        const int64_t row = row0 + tid;
        const double* src = nullptr;
        double* dst = nullptr;
        int inst = 0;
        if (row < P.nrows) {
            inst = (int)(row / P.nch);
            const int c = (int)(row - (int64_t)inst * P.nch);
            const BufRef ib = P.bufrefs[(size_t)inst * P.nbuf + P.in_buf];
            const BufRef ob = P.bufrefs[(size_t)inst * P.nbuf + P.out_buf];
            src = reinterpret_cast<const double*>(ib.ptr) + (int64_t)c * ib.ld;
            dst = reinterpret_cast<double*>(ob.ptr) + (int64_t)c * ob.ld;
        }
        s_src[tid] = src;
        s_dst[tid] = dst;
        s_inst[tid] = inst;
    }
    // polyphase banks: a shared-memory copy when the host found room for it
    const double* pf_tab = P.pfb;
    const double* dpf_tab = P.dpfb;
    if (P.tab_doubles > 0) {
        for (int i = tid; i < P.tab_doubles; i += kFmThreads) {
            tabs[i] = __ldg(P.pfb + i);
            if (P.dpfb) tabs[P.tab_doubles + i] = __ldg(P.dpfb + i);
        }
        pf_tab = tabs;
        if (P.dpfb) dpf_tab = tabs + P.tab_doubles;
    }
    if (tid == 0) {
        mbar_init(&bar_full[0], kFmAuxThreads);
        mbar_init(&bar_full[1], kFmAuxThreads);
        mbar_init(&bar_done[0], 4);
        mbar_init(&bar_done[1], 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_async_smem();
    __syncthreads();

    // position of ring slot 0: the (even) first window position of the block's first tile
    const int64_t pos_base = (__ldg(P.xi0 + t0 * kFmT) - P.tapsper + 1) & ~int64_t(1);
    const int hb_tile = 4 * P.ks * kFmHbPitch;          // doubles per tile buffer

    if (warp >= 4) {
        // ---------------- loader + tap bands: aux warp x owns rows [x*RB/4, (x+1)*RB/4) and group x ----------------
        const int aux = warp - 4;
        constexpr int RPA = RB / 4;                     // rows per aux warp (lane = row)
        const int myrow = aux * RPA + lane;
        const double* const src = lane < RPA ? s_src[myrow] : nullptr;
        double* const rrow = ring + (size_t)(lane < RPA ? myrow : 0) * P.pitch;
        int64_t have = pos_base;                        // positions below `have` are in the ring
        for (int64_t t = t0; t < t1; ++t) {
            const int64_t u = t - t0;
            const int s = (int)(u & 1);
            // everything that comes from global tables is requested before the wait
            // (a) ring extension: even tiles fetch what this tile and the next one need
            const int64_t tl = (u & 1) ? t : (t + 1 < t1 ? t + 1 : t);
            const int64_t need = (__ldg(P.xi0 + tl * kFmT + kFmT - 1) + 2) & ~int64_t(1);
            // (b) lane l holds shift / phase of output l of this tile
            const int64_t m_l = t * kFmT + lane;
            const int64_t xi_l = __ldg(P.xi0 + m_l);
            const double phi_l = __ldg(P.phi + m_l);
            const int64_t xg = __shfl_sync(0xffffffffu, xi_l, lane & ~7);
            const int sh_l = (int)(xi_l - xg);
            const double fl_l = floor(phi_l);
            const double alpha_l = phi_l - fl_l;
            const int off_l = ((int)fl_l - 1) * P.tapsper - sh_l;
            const int lo_l = m_l < P.n_out ? sh_l : P.ks;                 // band rows [lo, sh + tapsper) hold taps
            if (u >= 2) mbar_wait(&bar_done[s], (unsigned)((u >> 1) - 1) & 1u);

            // ---- ring: [a, b) comes from the signal by TMA, the rest of [have, need) by hand
            const int64_t a = have < 0 ? 0 : have;
            int64_t b = need < P.in_len ? need : (P.in_len & ~int64_t(1));
            if (b < a) b = a;
            const int cnt = (int)(b - a);
            if (lane < RPA && cnt < need - have) {
                for (int64_t p = have; p < need; ++p)
                    if (p < a || p >= b) rrow[(int)((p - pos_base) % P.ring)] = (src && p >= 0 && p < P.in_len) ? src[p] : 0.0;
            }
            // ---- tap band of group `aux`: lanes along the band rows
            double* band = hb + (size_t)s * hb_tile + (size_t)aux * P.ks * kFmHbPitch;
#pragma unroll 1
            for (int n = 0; n < 8; ++n) {
                const int o = aux * 8 + n;
                const int off = __shfl_sync(0xffffffffu, off_l, o);
                const int lo = __shfl_sync(0xffffffffu, lo_l, o);
                const int hi = __shfl_sync(0xffffffffu, sh_l, o) + P.tapsper;
                const double alpha = __shfl_sync(0xffffffffu, alpha_l, o);
                for (int k = lane; k < P.ks; k += 32) {
                    double h = 0.0;
                    if (k >= lo && k < hi) {
                        h = pf_tab[off + k];
                        if (dpf_tab) h = fma(alpha, dpf_tab[off + k], h);
                    }
                    band[k * kFmHbPitch + n] = h;
                }
            }
            // ---- every aux thread arrives once per tile; lanes with a copy announce its bytes first
            if (cnt && src) {
                mbar_expect_tx(&bar_full[s], (unsigned)cnt * 8u);
                const int ia = (int)((a - pos_base) % P.ring);
                const int first = cnt < P.ring - ia ? cnt : P.ring - ia;
                bulk_load(rrow + ia, src + a, (unsigned)first * 8u, &bar_full[s]);
                if (cnt > first) bulk_load(rrow, src + a + first, (unsigned)(cnt - first) * 8u, &bar_full[s]);
            } else
                mbar_arrive(&bar_full[s]);
            have = need > have ? need : have;
        }
    } else if (warp < 4) {
        // ---------------- DMMA ----------------
        const int g = warp;
        const int kk = lane & 3, rr = lane >> 2;
        double ssq[MF];
#pragma unroll
        for (int i = 0; i < MF; ++i) ssq[i] = 0.0;
        const double* arow = ring + (size_t)rr * P.pitch;
        const int nks = P.ks >> 2;
        int64_t qprev = pos_base;
        int qidx = 0;                                   // ring index of position qprev
        for (int64_t t = t0; t < t1; ++t) {
            const int s = (int)((t - t0) & 1);
            // first window position of my group, as a ring index (requested before the wait)
            const int64_t q = __ldg(P.xi0 + t * kFmT + 8 * g) - P.tapsper + 1;
            qidx += (int)(q - qprev);
            qprev = q;
            while (qidx >= P.ring) qidx -= P.ring;
            int idx = qidx + kk;
            if (idx >= P.ring) idx -= P.ring;
            const double* bp = hb + (size_t)s * hb_tile + (size_t)g * P.ks * kFmHbPitch + kk * kFmHbPitch + rr;
            double acc[MF][2];
#pragma unroll
            for (int i = 0; i < MF; ++i) acc[i][0] = acc[i][1] = 0.0;
            mbar_wait(&bar_full[s], (unsigned)((t - t0) >> 1) & 1u);
#pragma unroll 2
            for (int ks = 0; ks < nks; ++ks) {
                const double b = bp[ks * 4 * kFmHbPitch];
                const double* ap = arow + idx;
#pragma unroll
                for (int i = 0; i < MF; ++i) dmma884(acc[i][0], acc[i][1], ap[(size_t)i * 8 * P.pitch], b);
                idx += 4;
                if (idx >= P.ring) idx -= P.ring;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_done[s]);
            // fragment (row = 8i + lane/4, outputs 2(lane%4), +1) -> global
            const int64_t m = t * kFmT + 8 * g + 2 * kk;
#pragma unroll
            for (int i = 0; i < MF; ++i) {
                double* dst = s_dst[8 * i + rr];
                if (!dst) continue;
                if (m + 1 < P.n_out) {
                    *reinterpret_cast<double2*>(dst + m) = make_double2(acc[i][0], acc[i][1]);
                    ssq[i] = fma(acc[i][0], acc[i][0], fma(acc[i][1], acc[i][1], ssq[i]));
                } else if (m < P.n_out) {
                    dst[m] = acc[i][0];
                    ssq[i] = fma(acc[i][0], acc[i][0], ssq[i]);
                }
            }
        }
        if (P.sumsq_slot >= 0) {
#pragma unroll
            for (int i = 0; i < MF; ++i) {
                double v = ssq[i];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                if (kk == 0 && s_dst[8 * i + rr]) atomicAdd(P.scalars + (size_t)s_inst[8 * i + rr] * P.nscalars + P.sumsq_slot, v);
            }
        }
    }
}

}  // namespace sigops
