// K4/K5 (fastest path) — polyphase FIR resampling on the FP64 tensor cores, fed and drained by
// tensor-map TMA.
//
// Same arithmetic as k_fir_mma.cuh / k_fir.cuh (DSP.jl `filt!(out, FIRFilter{...}, x)` reached from
// src/filters.jl:252-255, filter built at src/reformatting.jl:92-99; the host replays the kernel's
// index recurrence into xi0[m], phi[m]):
//     y[m] = sum_t (pfb[phi][t] + alpha*dpfb[phi][t]) * x[xi0[m] - tapsPerPhi + 1 + t]
// as 8-output groups Y[row][n] = sum_k X[row][q + k] * H[k][n] on `mma.sync.m8n8k4.f64` (DMMA.8x8x4).
// What changes against k_fir_mma is who moves the data.  There, 128 lanes issue one bulk copy per
// row and tile (bound by the TMA unit's issue rate: ~1850 cycles per tile on the helper path) and
// the compute warps push their fragments to global memory themselves (~950 cycles per tile in the
// LSU).  Here the [rows][frames] matrices are described by tensor maps, so
//   * the input ring is made of SLOTS of 16 consecutive positions x 128 rows (a 16 KB box, 128-byte
//     swizzle), each filled by ONE `cp.async.bulk.tensor.2d` issued by a producer thread that runs
//     as far ahead as the ring allows (per-slot `full` barriers; a slot is reused once the last
//     tile that read it is done).  Slots start on 128-byte lines of the rows; positions before
//     the signal, past its end, and rows past the last one are zero-filled by the TMA unit —
//     the history before the first sample and the reference's zero padding (src/filters.jl:240);
//   * finished 32-output tiles are staged in shared memory (two 16-output swizzled boxes) and
//     written with two `cp.async.bulk.tensor.2d` stores by a store thread (whole 128-byte lines;
//     columns past n_out and rows past the last are clipped by the tensor bounds);
//   * the compute warps do nothing but LDS + DMMA + 8 shared-memory stores per tile.
// A fragment rows are taken in bit-reversed order (fragment row r -> signal row {0,4,2,6,1,5,3,7}[r])
// which makes both the swizzled A-fragment loads (any position alignment) and the staging stores
// bank-conflict free.  The tap bands live at pitch 8 with an XOR on the column index (conflict-free
// B-fragment loads) and are merged by four helper warps ON THE TENSOR PIPE (see below).  A fused constant gain (`ToFramerate |> Amplify(c)` is one launch) is folded into
// the staged banks: g*(sum h x) becomes sum (g h) x, a rounding-level difference.  Groups whose eight outputs span fewer positions skip
// the last k-step (44.1 -> 48 kHz: 11.4 instead of 12 on average).
//
// Block = 16 warps in four warpgroups that re-balance their registers with setmaxnreg: 8 compute (4 output
// groups x 2 row halves, 176 registers), 4 helpers (tap bands, one per group), 1 producer (ring loads),
// 1 store.  Barriers: full[slot], done[tile&3] (compute warps), taps[tile&1],
// stg_full / stg_free (staging hand-over).
//
// Eligibility (host): Float64 in/out, every row of the wave at base + row*stride with 16-byte
// aligned base and stride (one batch tensor, or the library's own staging), epilogue = none or
// constant gain, band <= 64 positions, a tile's window within the ring.  Everything else takes
// k_fir_mma / k_fir.  SIGOPS_NO_FIR_TMAP=1 switches this kernel off.
#pragma once
#include <cuda.h>

#include "k_fir_mma.cuh"

namespace sigops {

constexpr int kFtRows = 128;                    // rows per block
constexpr int kFtSlotPos = 16;                  // positions per ring slot (one 128-byte row piece)
constexpr int kFtSlotBytes = kFtRows * 128;     // 16 KB
constexpr int kFtMaxSlots = 10;
constexpr int kFtNCW = 8, kFtNAW = 4;           // compute / helper warps
constexpr int kFtThreads = 16 * 32;              // four warpgroups: compute x2, helpers, producer + store (+ 2 idle warps)
// registers per thread after the warpgroups re-balance the 128 they are launched with (setmaxnreg)
constexpr int kFtRegsCompute = 176, kFtRegsHelper = 112, kFtRegsMisc = 40;
constexpr int kFtMaxKs = 64;

struct FirTmParams {
    double* scalars;
    int nscalars, sumsq_slot, nch;
    int64_t nrows, n_out;
    int tapsper;
    int ks;                 // positions per 8-output band (multiple of 4, <= 64)
    int nslot;              // ring slots
    int64_t ntiles, tiles_per_seg;
    const double* pfb;      // [nphases][tapsper]
    const double* dpfb;     // or nullptr
    const int64_t* xi0;     // padded to a multiple of 64 entries
    const int32_t* poff;    // [m] (phase index - 1) * tapsper
    const double* alpha;    // [m] fractional phase
    int tab_doubles;
    double gain;            // folded into the taps
    long long* dbg;         // optional [blocks][8] cycle counters (tuning aid, SIGOPS_FIR_DBG=1), or nullptr
    int exp;                // tuning experiments (SIGOPS_FIR_EXP bit mask; results are wrong when set): 1 no staging
                            // stores, 2 no tensor stores, 4 no A-operand loads in the k-loop, 8 no DMMAs
};

inline size_t fir_tm_smem_bytes(int nslot, int ks, int tab_doubles, bool has_dpfb) {
    return (size_t)nslot * kFtSlotBytes + 2 * 16384 + (size_t)2 * 4 * ks * 8 * sizeof(double) +
           (size_t)tab_doubles * sizeof(double) * (has_dpfb ? 2 : 1) + 1024;
}

// The dynamic shared memory of this kernel is addressed through 32-bit shared-window addresses and
// explicit ld.shared / st.shared: the 1024-byte alignment of the swizzled boxes is computed on the
// address, and a pointer rebuilt from it would make every access a generic LD/ST.
__device__ __forceinline__ void tmap_load_2d(unsigned smem_dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_dst),
                 "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmap_store_2d(const CUtensorMap* tm, int c0, int c1, unsigned smem_src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm), "r"(c0), "r"(c1), "r"(smem_src)
                 : "memory");
}
__device__ __forceinline__ double lds_f64(unsigned a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f64(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts_v2f64(unsigned a, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}

template <bool SSQ>
__global__ void __launch_bounds__(kFtThreads, 1)
k_fir_tmap(const __grid_constant__ FirTmParams P, const __grid_constant__ CUtensorMap tm_in,
           const __grid_constant__ CUtensorMap tm_out) {
    extern __shared__ unsigned char ft_smem_raw[];
    __shared__ uint64_t bar_full[kFtMaxSlots], bar_done[4], bar_taps[2], bar_stg_full, bar_stg_free;
    // tiles finished, counted once per compute warp: what the producer polls.  (A plain counter, not the
    // `done` barriers: with strong up-sampling the ring holds many tiles, the producer may trail the compute
    // warps by more than the two phases a parity wait can tell apart.)
    __shared__ unsigned done_count;

    const unsigned ring = (smem_u32(ft_smem_raw) + 1023u) & ~1023u;               // [nslot][128 rows][128 B], swizzled
    const unsigned stg = ring + (unsigned)P.nslot * kFtSlotBytes;                  // [2][128 rows][128 B], swizzled
    const unsigned bands = stg + 2 * 16384;                                        // [2][4][ks][8] doubles
    const unsigned tabs = bands + 2u * 4u * (unsigned)P.ks * 64u;                  // pfb, dpfb

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = (int)blockIdx.y * kFtRows;
    const int64_t t0 = (int64_t)blockIdx.x * P.tiles_per_seg;
    const int64_t t1 = (t0 + P.tiles_per_seg < P.ntiles) ? t0 + P.tiles_per_seg : P.ntiles;
    if (t0 >= t1) return;

    // both banks ride along in shared memory, with the constant-gain epilogue folded in once per block
    for (int i = tid; i < P.tab_doubles; i += kFtThreads) {
        sts_f64(tabs + 8u * i, __ldg(P.pfb + i) * P.gain);
        if (P.dpfb) sts_f64(tabs + 8u * (P.tab_doubles + i), __ldg(P.dpfb + i) * P.gain);
    }
    if (tid == 0) {
        done_count = 0u;
        for (int i = 0; i < P.nslot; ++i) mbar_init(&bar_full[i], 1);
        for (int i = 0; i < 4; ++i) mbar_init(&bar_done[i], kFtNCW);
        mbar_init(&bar_taps[0], kFtNAW * 32);
        mbar_init(&bar_taps[1], kFtNAW * 32);
        mbar_init(&bar_stg_full, kFtNCW);
        mbar_init(&bar_stg_free, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_async_smem();
    __syncthreads();

    // position of ring slot 0: the first window position of the block's first tile, rounded down to a
    // 128-byte line of the rows
    const int64_t pos_base = (__ldg(P.xi0 + t0 * kFmT) - P.tapsper + 1) & ~int64_t(kFtSlotPos - 1);
    const unsigned band_tile = 4u * (unsigned)P.ks * 64u;      // bytes per band buffer

    if (warp < kFtNCW) {
        // ---------------- DMMA ----------------
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kFtRegsCompute));
        const int g = warp & 3, half = warp >> 2;
        const int kk = lane & 3, rr = lane >> 2;
        const int rmap = ((rr & 1) << 2) | (rr & 2) | (rr >> 2);        // 3-bit reversal
        const int rloc = half * 64 + rmap;                               // fragment i holds row rloc + 8i
        const unsigned key0 = (unsigned)rmap << 4;                       // swizzle: 16-byte chunk ^= row & 7
        const unsigned arow0 = ring + rloc * 128;
        const unsigned srow0 = stg + (g >> 1) * 16384 + rloc * 128 + ((((g & 1) * 4 + kk) << 4) ^ key0);
        const int nn = rr ^ ((kk >> 1) << 2);                            // band column after the XOR
        const int nks_max = P.ks >> 2;
        // (opaque to the optimiser: otherwise ptxas re-derives these from the thread index inside the k-loop —
        //  S2R + a dozen integer instructions per k-step — instead of keeping three registers)
        unsigned arow = arow0, key = key0, srow = srow0;
        asm volatile("" : "+r"(arow), "+r"(key), "+r"(srow));
        double ssq[SSQ ? 8 : 1];
#pragma unroll
        for (int i = 0; i < (SSQ ? 8 : 1); ++i) ssq[i] = 0.0;

        int64_t xq_nx = __ldg(P.xi0 + t0 * kFmT + 8 * g);
        int64_t x7_nx = __ldg(P.xi0 + t0 * kFmT + 8 * g + 7);
        int64_t xe_nx = __ldg(P.xi0 + t0 * kFmT + kFmT - 1);
        int64_t jwait = 0;                              // slots [0, jwait) have been waited for
        int jw_slot = 0;
        unsigned jw_par = 0;
        int64_t jslot0 = 0;                             // global slot index behind `slot_t`
        int slot_t = 0;
        long long dbg_acc[3] = {0, 0, 0};

        // operand stream of the tile in flight: ring slot / byte offset in the 128-byte row of the k-step
        // loaded last, band pointer, k-steps
        int slot = 0, up = 0, nks = 0;
        unsigned bp = 0;
        // two operand sets: one feeds the DMMAs of a k-step while the other is being loaded for the next
        double a0[8], b0, a1[8], b1;

        // Set tile t up (index arithmetic, barrier waits) and load the operands of its first k-step into
        // (a, b).  Runs underneath the last k-step of the tile before it.
        auto setup = [&](int64_t t, double (&a)[8], double& b) {
            const int64_t u = t - t0;
            const int s = (int)(u & 1);
            const int64_t q = xq_nx - P.tapsper + 1;
            int n = (int)((x7_nx - xq_nx + P.tapsper + 3) >> 2);
            nks = n < nks_max ? n : nks_max;
            const int64_t hi = (xe_nx + 1 - pos_base + (kFtSlotPos - 1)) >> 4;     // slots [0, hi) hold the tile's window
            if (t + 1 < t1) {
                xq_nx = __ldg(P.xi0 + (t + 1) * kFmT + 8 * g);
                x7_nx = __ldg(P.xi0 + (t + 1) * kFmT + 8 * g + 7);
                xe_nx = __ldg(P.xi0 + (t + 1) * kFmT + kFmT - 1);
            }
            const int64_t rel = q - pos_base + kk;                                   // >= 0
            const int64_t j0 = rel >> 4;
            slot_t += (int)(j0 - jslot0);                                            // windows only move forward, a few slots per tile
            jslot0 = j0;
            while (slot_t >= P.nslot) slot_t -= P.nslot;
            slot = slot_t;
            up = (int)(rel & 15) << 3;
            bp = bands + s * band_tile + (unsigned)(g * P.ks + kk) * 64u + nn * 8u;
            const long long c0 = P.dbg ? clock64() : 0;
            mbar_wait(&bar_taps[s], (unsigned)(u >> 1) & 1u);
            const long long c1 = P.dbg ? clock64() : 0;
            while (jwait < hi) {
                mbar_wait(&bar_full[jw_slot], jw_par);
                ++jwait;
                if (++jw_slot == P.nslot) { jw_slot = 0; jw_par ^= 1u; }
            }
            if (P.dbg) {
                const long long c2 = clock64();
                dbg_acc[0] += c1 - c0; dbg_acc[1] += c2 - c1;
            }
            const unsigned ap = arow + slot * kFtSlotBytes + ((unsigned)up ^ key);
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = lds_f64(ap + i * 1024);
            b = lds_f64(bp);
        };
        // operands of k-step `ksn` of the tile in flight (the next one along its stream)
        auto load_next = [&](double (&a)[8], double& b, unsigned bp_t, int ksn) {
            up += 32;
            if (up >= 128) {
                up -= 128;
                if (++slot == P.nslot) slot = 0;
            }
            const unsigned ap = arow + slot * kFtSlotBytes + ((unsigned)up ^ key);
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = lds_f64(ap + i * 1024);
            b = lds_f64(bp_t + ksn * 256);
        };
        auto mma8 = [&](double (&acc)[8][2], const double (&a)[8], double b) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma884(acc[i][0], acc[i][1], a[i], b);
        };
        // fragments of a finished tile -> staging (the store thread has read the tile before it out of it)
        auto stage_out = [&](double (&o)[8][2], int64_t u_old) {
            const long long c0 = P.dbg ? clock64() : 0;
            if (u_old > 0) mbar_wait(&bar_stg_free, (unsigned)(u_old - 1) & 1u);
            if (P.dbg) dbg_acc[2] += clock64() - c0;
            if (!(P.exp & 1)) {
#pragma unroll
                for (int i = 0; i < 8; ++i) sts_v2f64(srow + i * 1024, o[i][0], o[i][1]);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_stg_full);
            if (SSQ) {
                // outputs past n_out have all-zero taps, rows past the last read zeros: exactly 0
#pragma unroll
                for (int i = 0; i < 8; ++i) ssq[i] = fma(o[i][0], o[i][0], fma(o[i][1], o[i][1], ssq[i]));
            }
        };
        // One tile: its DMMAs accumulate into `acc` while the fragments of the tile before it (`old`) drain to
        // the staging boxes underneath them and, during the last k-step, the next tile is set up and its
        // first operands are fetched: the tensor pipe is never left waiting for a tile boundary.  On entry
        // set 0 holds the operands of k-step 0; on exit it holds those of the next tile's.
        auto tile = [&](double (&acc)[8][2], double (&old)[8][2], int64_t t) {
            const int64_t u = t - t0;
            const bool more = t + 1 < t1;
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = 0.0;
            const int n = nks;
            const unsigned bp_t = bp;
            if (n >= 2) {
                load_next(a1, b1, bp_t, 1);
                mma8(acc, a0, b0);
                if (u > 0) stage_out(old, u - 1);
                int ks = 1;                                   // set 1 holds k-step ks
                for (; ks + 2 < n; ks += 2) {
                    load_next(a0, b0, bp_t, ks + 1);
                    mma8(acc, a1, b1);
                    load_next(a1, b1, bp_t, ks + 2);
                    mma8(acc, a0, b0);
                }
                if (n - ks == 2) {
                    load_next(a0, b0, bp_t, ks + 1);
                    mma8(acc, a1, b1);
                    if (more) setup(t + 1, a1, b1);
                    mma8(acc, a0, b0);
#pragma unroll
                    for (int i = 0; i < 8; ++i) a0[i] = a1[i];
                    b0 = b1;
                } else {
                    if (more) setup(t + 1, a0, b0);
                    mma8(acc, a1, b1);
                }
            } else {
                if (more) setup(t + 1, a1, b1);
                mma8(acc, a0, b0);
                if (u > 0) stage_out(old, u - 1);
#pragma unroll
                for (int i = 0; i < 8; ++i) a0[i] = a1[i];
                b0 = b1;
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&bar_done[u & 3]);
                asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" ::"r"(smem_u32(&done_count)) : "memory");
            }
        };

        double accA[8][2], accB[8][2];
        setup(t0, a0, b0);
        const long long cstart = P.dbg ? clock64() : 0;
        for (int64_t t = t0; t < t1; t += 2) {
            tile(accA, accB, t);
            if (t + 1 < t1) tile(accB, accA, t + 1);
        }
        if (P.dbg && warp == 0 && lane == 0) {
            long long* d = P.dbg + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8;
            d[0] = dbg_acc[0]; d[1] = dbg_acc[1]; d[2] = dbg_acc[2]; d[3] = clock64() - cstart;
        }
        if ((t1 - t0) & 1) stage_out(accA, t1 - t0 - 1);
        else stage_out(accB, t1 - t0 - 1);
        if (SSQ) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                double v = ssq[i];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                const int64_t row = (int64_t)row0 + rloc + 8 * i;
                if (kk == 0 && row < P.nrows) atomicAdd(P.scalars + (size_t)(row / P.nch) * P.nscalars + P.sumsq_slot, v);
            }
        }
    } else if (warp < kFtNCW + kFtNAW) {
        // ---------------- tap bands: helper warp g merges the taps of output group g — on the tensor pipe ----------------
        // h[t][n] = pfb[phi_n][t] + alpha_n * dpfb[phi_n][t] for the group's 8 outputs n is itself a small matrix
        // product: C (8 taps x 8 outputs, preloaded with the pfb values) += A (8 taps x 8: dpfb values) * B
        // (diag(alpha)), two DMMA.8x8x4 per 8 taps.  Scalar FP64 instructions of a helper warp queue behind the
        // DMMAs of the compute warps on the same sub-partition (measured here: ~200 cycles per DFMA, 4500 cycles
        // per tile for one FMA per merged tap); DMMAs of another warp simply interleave, at 5.5 % more
        // tensor work.  Operands are table look-ups and selects: no FP64 ALU instruction in this branch.
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kFtRegsHelper));
        const int grp = warp - kFtNCW;
        const int kk = lane & 3, rr = lane >> 2;
        const unsigned pf_tab = tabs, dpf_tab = tabs + 8u * P.tab_doubles;
        const bool has_d = P.dpfb != nullptr;
        const int nblk = (P.tapsper + 7) >> 3;            // 8-tap row blocks
        long long hdbg[2] = {0, 0};
        // lane l < 8 carries column l of the group (others mirror it)
        int64_t xi_nx = __ldg(P.xi0 + t0 * kFmT + 8 * grp + (lane & 7));
        int po_nx = __ldg(P.poff + t0 * kFmT + 8 * grp + (lane & 7));
        double al_nx = __ldg(P.alpha + t0 * kFmT + 8 * grp + (lane & 7));
        for (int64_t t = t0; t < t1; ++t) {
            const int64_t u = t - t0;
            const int s = (int)(u & 1);
            const int64_t m = t * kFmT + 8 * grp + (lane & 7);
            const int64_t xi = xi_nx;
            const bool live = m < P.n_out;
            const int po = po_nx;
            const double al = live ? al_nx : 0.0;
            if (t + 1 < t1) {
                xi_nx = __ldg(P.xi0 + m + kFmT);
                po_nx = __ldg(P.poff + m + kFmT);
                al_nx = __ldg(P.alpha + m + kFmT);
            }
            const int64_t xg = __shfl_sync(0xffffffffu, xi, 0);
            const int sh = (int)(xi - xg);                                       // band rows [sh, sh + tapsper) hold column (lane & 7)'s taps
            // what my fragment elements need: C columns 2kk, 2kk+1; A columns kk, 4+kk; B column rr
            const int sh_c0 = __shfl_sync(0xffffffffu, sh, 2 * kk), sh_c1 = __shfl_sync(0xffffffffu, sh, 2 * kk + 1);
            const int po_c0 = __shfl_sync(0xffffffffu, po, 2 * kk), po_c1 = __shfl_sync(0xffffffffu, po, 2 * kk + 1);
            const int po_a0 = __shfl_sync(0xffffffffu, po, kk), po_a1 = __shfl_sync(0xffffffffu, po, 4 + kk);
            const unsigned lv = __ballot_sync(0xffffffffu, live);
            const bool lv_c0 = (lv >> (2 * kk)) & 1u, lv_c1 = (lv >> (2 * kk + 1)) & 1u;
            const bool lv_a0 = (lv >> kk) & 1u, lv_a1 = (lv >> (4 + kk)) & 1u;
            const double al_b = __shfl_sync(0xffffffffu, al, rr);
            const double b0 = (kk == rr) ? al_b : 0.0, b1 = (4 + kk == rr) ? al_b : 0.0;
            const unsigned band = bands + s * band_tile + (unsigned)(grp * P.ks) * 64u;
            auto cell = [&](int k, int n) { return band + 8u * (unsigned)(k * 8 + (n ^ (((k >> 1) & 1) << 2))); };
            // tile t-2 must be finished: its band buffer is about to be overwritten
            const long long c0 = P.dbg ? clock64() : 0;
            if (u >= 2) mbar_wait(&bar_done[(u - 2) & 3], (unsigned)((u - 2) >> 2) & 1u);
            const long long c1 = P.dbg ? clock64() : 0;
            hdbg[0] += c1 - c0;
            // rows outside a column's taps are zero (the buffer still holds tile t-2's band): lane = (column, row mod 4)
            {
                const int n = lane & 7, part = lane >> 3;
                const int top = live ? sh : 0, bot = live ? sh + P.tapsper : 0;
                for (int k = part; k < P.ks; k += 4)
                    if (k < top || k >= bot) sts_f64(cell(k, n), 0.0);
            }
            for (int b = 0; b < nblk; ++b) {
                const int tt = 8 * b + rr;                                       // my tap row
                const bool in = tt < P.tapsper;
                double cv0 = (in && lv_c0) ? lds_f64(pf_tab + 8u * (unsigned)(po_c0 + tt)) : 0.0;
                double cv1 = (in && lv_c1) ? lds_f64(pf_tab + 8u * (unsigned)(po_c1 + tt)) : 0.0;
                if (has_d) {
                    const double a0 = (in && lv_a0) ? lds_f64(dpf_tab + 8u * (unsigned)(po_a0 + tt)) : 0.0;
                    const double a1 = (in && lv_a1) ? lds_f64(dpf_tab + 8u * (unsigned)(po_a1 + tt)) : 0.0;
                    dmma884(cv0, cv1, a0, b0);
                    dmma884(cv0, cv1, a1, b1);
                }
                if (in && lv_c0) sts_f64(cell(sh_c0 + tt, 2 * kk), cv0);
                if (in && lv_c1) sts_f64(cell(sh_c1 + tt, 2 * kk + 1), cv1);
            }
            mbar_arrive(&bar_taps[s]);
            if (P.dbg) hdbg[1] += clock64() - c1;
        }
        if (P.dbg && grp == 0 && lane == 0) {
            long long* d = P.dbg + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8;
            d[4] = hdbg[0]; d[5] = hdbg[1];
        }
    } else {
      // (one setmaxnreg for the whole fourth warpgroup: producer, store and two warps that only give registers away)
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kFtRegsMisc));
      if (warp == kFtNCW + kFtNAW) {
        // ---------------- producer: one tensor copy per ring slot, as far ahead as the ring allows ----------------
        if (lane == 0) {
            const int64_t last_need = __ldg(P.xi0 + (t1 - 1) * kFmT + kFmT - 1) + 1;
            const int64_t nslots_total = (last_need - pos_base + (kFtSlotPos - 1)) >> 4;
            auto lo_of = [&](int64_t t) { return (__ldg(P.xi0 + t * kFmT) - P.tapsper + 1 - pos_base) >> 4; };
            int64_t tp = t0;
            int64_t lo_next = t0 + 1 < t1 ? lo_of(t0 + 1) : (int64_t(1) << 62);
            int s = 0;
            for (int64_t j = 0; j < nslots_total; ++j) {
                if (j >= P.nslot) {
                    // the last tile that reads slot j - nslot must be done before the slot is refilled
                    const int64_t k = j - P.nslot;
                    while (lo_next <= k) {
                        ++tp;
                        lo_next = tp + 1 < t1 ? lo_of(tp + 1) : (int64_t(1) << 62);
                    }
                    const unsigned target = (unsigned)(tp - t0 + 1) * kFtNCW;
                    unsigned seen;
                    do {
                        asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(seen) : "r"(smem_u32(&done_count)) : "memory");
                    } while (seen < target);
                }
                mbar_expect_tx(&bar_full[s], kFtSlotBytes);
                tmap_load_2d(ring + (unsigned)s * kFtSlotBytes, &tm_in, (int)(pos_base + j * kFtSlotPos), row0, &bar_full[s]);
                if (++s == P.nslot) s = 0;
            }
        }
      } else if (warp == kFtNCW + kFtNAW + 1) {
        // ---------------- store: two tensor stores per tile out of the staging boxes ----------------
        if (lane == 0) {
            for (int64_t t = t0; t < t1; ++t) {
                const int64_t u = t - t0;
                mbar_wait(&bar_stg_full, (unsigned)u & 1u);
                const int64_t m0 = t * kFmT;
                if (!(P.exp & 2)) {
                    tmap_store_2d(&tm_out, (int)m0, row0, stg);
                    if (m0 + 16 < P.n_out) tmap_store_2d(&tm_out, (int)(m0 + 16), row0, stg + 16384);
                }
                bulk_commit();
                bulk_wait_read_all();
                mbar_arrive(&bar_stg_free);
            }
            bulk_wait_all();
        }
      }
    }
}

}  // namespace sigops
