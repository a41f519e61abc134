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
//   * the compute warps do nothing but LDS + DMMA + 16 shared-memory stores per tile.  Shared-memory
//     bandwidth is the co-critical resource (ncu on the first version: LSU wavefronts + TMA traffic busy 70 %
//     of the time, tensor pipe 66 %), so the k axis is walked in ring-aligned blocks of 4 positions whose A
//     fragments are shared by all four output groups of the tile.
// A fragment rows are taken in bit-reversed order (fragment row r -> signal row {0,4,2,6,1,5,3,7}[r])
// which makes both the swizzled A-fragment loads (any position alignment) and the staging stores
// bank-conflict free.  The tap bands live at pitch 8 with an XOR on the column index (conflict-free
// B-fragment loads) and are merged by four helper warps ON THE TENSOR PIPE (see below).  A fused constant gain (`ToFramerate |> Amplify(c)` is one launch) is folded into
// the staged banks: g*(sum h x) becomes sum (g h) x, a rounding-level difference.  Groups whose eight outputs span fewer positions skip
// the last k-step (44.1 -> 48 kHz: 11.4 instead of 12 on average).
//
// Block = 16 warps in four warpgroups that re-balance their registers with setmaxnreg: 8 compute (two per
// sub-partition; each 8/G row fragments x G output groups, 176 registers), 4 helpers (tap bands, one per
// group), 1 producer (ring loads), 1 store.  Barriers: full[slot], done[tile&3] (compute warps), taps[tile&1],
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
    int aligned;            // 1: band row 0 = the group's first window position rounded down to a multiple of 4 ring
                            //    positions (compute warps share A blocks between groups, G > 1); 0: exactly that position
    long long* dbg;         // optional [blocks][8] cycle counters (tuning aid, SIGOPS_FIR_DBG=1), or nullptr
    int exp;                // tuning experiments (SIGOPS_FIR_EXP bit mask; wrong results): 1 no staging stores, 2 no tensor
                            // stores, 4 no tap-band building
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

template <bool SSQ, int G>
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
        // ---------------- DMMA: a warp owns F = 8/G row fragments (8F rows) x G consecutive output groups of the tile ----------------
        // The k axis is walked in blocks of 4 positions aligned to the ring (position mod 4 == 0): the A
        // fragments of a block are loaded once and feed every group of the warp whose band covers the block.
        // Two compute warps per sub-partition: a single one cannot hide its own index arithmetic and operand
        // loads behind its DMMAs (measured with 4 warps x 16 fragments: 39 cycles per DMMA instead of 16).
        constexpr int F = 8 / G;
        const int gw = warp % (4 / G), rw = warp / (4 / G);
        const int g0 = G * gw;                                           // my first output group
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kFtRegsCompute));
        const int kk = lane & 3, rr = lane >> 2;
        const int rmap = ((rr & 1) << 2) | (rr & 2) | (rr >> 2);        // 3-bit reversal
        const int rloc = rw * (8 * F) + rmap;                            // fragment i holds row rloc + 8i
        const unsigned key0 = (unsigned)rmap << 4;                       // swizzle: 16-byte chunk ^= row & 7
        const int nn = rr ^ ((kk >> 1) << 2);                            // band column after the XOR
        // (opaque to the optimiser: otherwise ptxas re-derives these from the thread index inside the k-loop —
        //  S2R + a dozen integer instructions per k-step — instead of keeping a few registers)
        unsigned arow = ring + rloc * 128 + ((unsigned)(kk * 8) ^ key0);  // my element of block 0 of slot 0, swizzled
        unsigned key = key0;
        unsigned srow = stg + rloc * 128;
        unsigned bcol = (unsigned)(kk * 64 + nn * 8);                    // my B element inside a 4-row band block
        unsigned arow1 = ring + rloc * 128;                              // (G == 1) my row of slot 0
        asm volatile("" : "+r"(arow), "+r"(key), "+r"(srow), "+r"(bcol), "+r"(arow1));
        double ssq[SSQ ? 8 : 1];
#pragma unroll
        for (int i = 0; i < (SSQ ? 8 : 1); ++i) ssq[i] = 0.0;

        // index tables, read one tile ahead: first / last output of every group, last output of the tile
        int64_t xq_nx[G], x7_nx[G];
        auto fetch_idx = [&](int64_t t) {
#pragma unroll
            for (int g = 0; g < G; ++g) {
                xq_nx[g] = __ldg(P.xi0 + t * kFmT + 8 * (g0 + g));
                x7_nx[g] = __ldg(P.xi0 + t * kFmT + 8 * (g0 + g) + 7);
            }
        };
        fetch_idx(t0);
        long long dbg_acc[3] = {0, 0, 0};

        // the tile in flight: first aligned block (global index and ring slot / quarter), blocks in all, and per
        // group the first block and the number of blocks of its band
        int64_t blk0 = 0;                               // global block index (position / 4 relative to pos_base) of the block loaded last
        int slot = 0, sub = 0;                          // its ring slot and quarter of the slot (G == 1: per lane, byte offset in the row)
        int64_t jcur = 0;                               // (G == 1) global slot index behind `slot`
        int nbt = 0, ob[G], nb[G];
        unsigned bandbase = 0;
        struct Ops { double a[F], b[G]; };
        Ops o0, o1;

        auto load_ops = [&](Ops& o, int b, unsigned bb, const int (&obx)[G], const int (&nbx)[G]) {
            // G > 1: blocks are ring-aligned, (slot, sub) are warp-uniform; low 7 address bits = (kk*8 ^ swizzle key)
            // from `arow`, quarter of the slot in bits 5-6.  G == 1: the band starts at the group's own first
            // window position, every lane tracks its own position (slot, byte offset `sub` in the 128-byte row).
            const unsigned ap = G > 1 ? (arow + (unsigned)slot * kFtSlotBytes) ^ (unsigned)(sub << 5)
                                      : arow1 + (unsigned)slot * kFtSlotBytes + ((unsigned)sub ^ key);
#pragma unroll
            for (int i = 0; i < F; ++i) o.a[i] = lds_f64(ap + i * 1024);
#pragma unroll
            for (int g = 0; g < G; ++g) {
                int kb = b - obx[g];                    // clamped: groups that do not cover the block load a row they own anyway
                kb = kb < 0 ? 0 : (kb >= nbx[g] ? nbx[g] - 1 : kb);
                o.b[g] = lds_f64(bb + (unsigned)((g0 + g) * P.ks + 4 * kb) * 64u + bcol);
            }
        };
        auto advance = [&]() {                           // next block of 4 positions along the ring
            ++blk0;
            if (G > 1) {
                if (++sub == 4) {
                    sub = 0;
                    if (++slot == P.nslot) slot = 0;
                }
            } else {
                sub += 32;
                if (sub >= 128) {
                    sub -= 128;
                    ++jcur;
                    if (++slot == P.nslot) slot = 0;
                }
            }
        };
        // Set tile t up (index arithmetic, barrier waits) and load the operands of its first block into `o`.
        // Runs underneath the last block of the tile before it.
        auto setup = [&](int64_t t, Ops& o) {
            const int64_t u = t - t0;
            const int s = (int)(u & 1);
            int64_t ab[G];
            int last = 0;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const int64_t r = xq_nx[g] - P.tapsper + 1 - pos_base;               // >= 0: first window position of the group
                ab[g] = r >> 2;
                nb[g] = (int)(((G > 1 ? (r & 3) : 0) + (x7_nx[g] - xq_nx[g]) + P.tapsper + 3) >> 2);
                nb[g] = nb[g] < (P.ks >> 2) ? nb[g] : (P.ks >> 2);
            }
            const int64_t rlane = xq_nx[0] - P.tapsper + 1 - pos_base + kk;          // (G == 1) my position of block 0
#pragma unroll
            for (int g = 0; g < G; ++g) {
                ob[g] = (int)(ab[g] - ab[0]);
                last = last > ob[g] + nb[g] ? last : ob[g] + nb[g];
            }
            nbt = last;
            if (t + 1 < t1) fetch_idx(t + 1);
            // ring position of the tile's first block, relative to the block loaded last (the end of the tile in
            // flight): a few blocks back, windows overlap
            if (G > 1) {
                const int adv = (int)(ab[0] - blk0);
                blk0 = ab[0];
                sub += adv;
                slot += sub >> 2;                       // (arithmetic shift: floor for negative steps)
                sub &= 3;
            } else {
                const int64_t j_new = rlane >> 4;       // global slot index of my position of block 0
                slot += (int)(j_new - jcur);
                jcur = j_new;
                sub = (int)(rlane & 15) << 3;
            }
            while (slot >= P.nslot) slot -= P.nslot;
            while (slot < 0) slot += P.nslot;
            bandbase = bands + s * band_tile;
            // one barrier per tile: the helpers arrive on it once the tile's band is built AND its ring slots
            // have landed (they watch the `full` barriers, off the compute warps' path)
            const long long c0 = P.dbg ? clock64() : 0;
            mbar_wait(&bar_taps[s], (unsigned)(u >> 1) & 1u);
            if (P.dbg) dbg_acc[0] += clock64() - c0;
            load_ops(o, 0, bandbase, ob, nb);
        };
        auto mma_block = [&](double (&acc)[F][G][2], const Ops& o, int b, const int (&obx)[G], const int (&nbx)[G]) {
#pragma unroll
            for (int g = 0; g < G; ++g) {
                if (G == 1 || (b >= obx[g] && b < obx[g] + nbx[g])) {
#pragma unroll
                    for (int i = 0; i < F; ++i) dmma884(acc[i][g][0], acc[i][g][1], o.a[i], o.b[g]);
                }
            }
        };
        // fragments of a finished tile -> staging (the store thread has read the tile before it out of it)
        auto stage_out = [&](double (&o)[F][G][2], int64_t u_old) {
            const long long c0 = P.dbg ? clock64() : 0;
            if (u_old > 0) mbar_wait(&bar_stg_free, (unsigned)(u_old - 1) & 1u);
            if (P.dbg) dbg_acc[2] += clock64() - c0;
            if (!(P.exp & 1)) {
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const int gg = g0 + g;
                    const unsigned sp = srow + (gg >> 1) * 16384 + ((unsigned)(((gg & 1) * 4 + kk) << 4) ^ key);
#pragma unroll
                    for (int i = 0; i < F; ++i) sts_v2f64(sp + i * 1024, o[i][g][0], o[i][g][1]);
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_stg_full);
            if (SSQ) {
                // outputs past n_out have all-zero taps, rows past the last read zeros: exactly 0
#pragma unroll
                for (int i = 0; i < F; ++i)
#pragma unroll
                    for (int g = 0; g < G; ++g)
                        ssq[G * i + g] = fma(o[i][g][0], o[i][g][0], fma(o[i][g][1], o[i][g][1], ssq[G * i + g]));
            }
        };
        // One tile: its DMMAs accumulate into `acc` while the fragments of the tile before it (`old`) drain to
        // the staging boxes underneath them and, during the last block, the next tile is set up and its
        // first operands are fetched: the tensor pipe is never left waiting for a tile boundary.  On entry
        // o0 holds the operands of block 0; on exit it holds those of the next tile's.
        auto tile = [&](double (&acc)[F][G][2], double (&old)[F][G][2], int64_t t) {
            const int64_t u = t - t0;
            const bool more = t + 1 < t1;
#pragma unroll
            for (int i = 0; i < F; ++i)
#pragma unroll
                for (int g = 0; g < G; ++g) acc[i][g][0] = acc[i][g][1] = 0.0;
            // this tile's schedule (setup of the next tile overwrites the shared copies)
            const int n = nbt;
            const unsigned bb = bandbase;
            int obt[G], nbx[G];
#pragma unroll
            for (int g = 0; g < G; ++g) { obt[g] = ob[g]; nbx[g] = nb[g]; }
            if (n >= 2) {
                advance();
                load_ops(o1, 1, bb, obt, nbx);
                mma_block(acc, o0, 0, obt, nbx);
                if (u > 0) stage_out(old, u - 1);
                int b = 1;                                    // o1 holds block b
                for (; b + 2 < n; b += 2) {
                    advance();
                    load_ops(o0, b + 1, bb, obt, nbx);
                    mma_block(acc, o1, b, obt, nbx);
                    advance();
                    load_ops(o1, b + 2, bb, obt, nbx);
                    mma_block(acc, o0, b + 1, obt, nbx);
                }
                if (n - b == 2) {
                    advance();
                    load_ops(o0, b + 1, bb, obt, nbx);
                    mma_block(acc, o1, b, obt, nbx);
                    if (more) setup(t + 1, o1);
                    mma_block(acc, o0, b + 1, obt, nbx);
                    o0 = o1;
                } else {
                    if (more) setup(t + 1, o0);
                    mma_block(acc, o1, b, obt, nbx);
                }
            } else {
                if (more) setup(t + 1, o1);
                mma_block(acc, o0, 0, obt, nbx);
                if (u > 0) stage_out(old, u - 1);
                o0 = o1;
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&bar_done[u & 3]);
                asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" ::"r"(smem_u32(&done_count)) : "memory");
            }
        };

        double accA[F][G][2], accB[F][G][2];
        setup(t0, o0);
        const long long cstart = P.dbg ? clock64() : 0;
        for (int64_t t = t0; t < t1; t += 2) {
            tile(accA, accB, t);
            if (t + 1 < t1) tile(accB, accA, t + 1);
        }
        if (P.dbg && warp == 0 && lane == 0) {
            long long* d = P.dbg + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8;
            d[0] = dbg_acc[0]; d[1] = 0; d[2] = dbg_acc[2]; d[3] = clock64() - cstart;
        }
        if ((t1 - t0) & 1) stage_out(accA, t1 - t0 - 1);
        else stage_out(accB, t1 - t0 - 1);
        if (SSQ) {
#pragma unroll
            for (int i = 0; i < F; ++i) {
                double v = 0.0;
#pragma unroll
                for (int g = 0; g < G; ++g) v += ssq[G * i + g];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                const int64_t row = (int64_t)row0 + rloc + 8 * i;
                if (kk == 0 && row < P.nrows) atomicAdd(P.scalars + (size_t)(row / P.nch) * P.nscalars + P.sumsq_slot, v);
            }
        }
    } else if (warp < kFtNCW + kFtNAW) {
        // ---------------- tap bands: helper warp g merges the taps of output group g — on the tensor pipe ----------------
        // h[k][n] = pfb[phi_n][k - s_n] + alpha_n * dpfb[phi_n][k - s_n] for band rows k and the group's 8 outputs n
        // (s_n = where output n's taps start in the band, zero outside) is itself a small matrix product:
        // C (8 rows x 8 outputs, preloaded with the pfb values) += A (8 rows x 8: dpfb values) * B (diag(alpha)),
        // two DMMA.8x8x4 per 8 band rows.  Scalar FP64 instructions of a helper warp queue behind the DMMAs of
        // the compute warp on the same sub-partition (measured here: ~200 cycles per DFMA, 4500 cycles per tile
        // for one FMA per merged tap); DMMAs of another warp simply interleave, at 6 % more tensor work.
        // Operands are table look-ups and selects: no FP64 ALU instruction in this branch.  Whole band rows are
        // written (zeros included) as 16-byte pairs.  Band row 0 sits at the group's first window position
        // rounded down to a multiple of 4 ring positions, so that the compute warps can share A blocks.
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kFtRegsHelper));
        const int grp = warp - kFtNCW;
        const int kk = lane & 3, rr = lane >> 2;
        const unsigned pf_tab = tabs, dpf_tab = tabs + 8u * P.tab_doubles;
        const bool has_d = P.dpfb != nullptr;
        const int nblk = P.ks >> 3;                        // 8-row band blocks (ks is a multiple of 8 here)
        long long hdbg[2] = {0, 0};
        int64_t jwait = 0;                              // ring slots [0, jwait) have been seen to land
        int jw_slot = 0;
        unsigned jw_par = 0;
        int64_t xe_nx = __ldg(P.xi0 + t0 * kFmT + kFmT - 1);
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
            const int64_t hi = (xe_nx + 1 - pos_base + (kFtSlotPos - 1)) >> 4;     // slots [0, hi) hold the tile's window
            if (t + 1 < t1) {
                xi_nx = __ldg(P.xi0 + m + kFmT);
                po_nx = __ldg(P.poff + m + kFmT);
                al_nx = __ldg(P.alpha + m + kFmT);
                xe_nx = __ldg(P.xi0 + (t + 1) * kFmT + kFmT - 1);
            }
            const int64_t xg = __shfl_sync(0xffffffffu, xi, 0);
            const int og = P.aligned ? (int)((xg - P.tapsper + 1 - pos_base) & 3) : 0;   // the group's window start inside its first aligned block
            // band rows [st, st + tapsper) hold column (lane & 7)'s taps; a dead column (past n_out) holds none
            const int st = live ? (int)(xi - xg) + og : (1 << 20);
            // what my fragment elements need: C columns 2kk, 2kk+1; A columns kk, 4+kk; B column rr
            const int st_c0 = __shfl_sync(0xffffffffu, st, 2 * kk), st_c1 = __shfl_sync(0xffffffffu, st, 2 * kk + 1);
            const int po_c0 = __shfl_sync(0xffffffffu, po, 2 * kk), po_c1 = __shfl_sync(0xffffffffu, po, 2 * kk + 1);
            const int st_a0 = __shfl_sync(0xffffffffu, st, kk), st_a1 = __shfl_sync(0xffffffffu, st, 4 + kk);
            const int po_a0 = __shfl_sync(0xffffffffu, po, kk), po_a1 = __shfl_sync(0xffffffffu, po, 4 + kk);
            const double al_b = __shfl_sync(0xffffffffu, al, rr);
            const double b0 = (kk == rr) ? al_b : 0.0, b1 = (4 + kk == rr) ? al_b : 0.0;
            const unsigned band = bands + s * band_tile + (unsigned)(grp * P.ks) * 64u;
            // tile t-2 must be finished: its band buffer is about to be overwritten
            const long long c0 = P.dbg ? clock64() : 0;
            if (u >= 2) mbar_wait(&bar_done[(u - 2) & 3], (unsigned)((u - 2) >> 2) & 1u);
            const long long c1 = P.dbg ? clock64() : 0;
            hdbg[0] += c1 - c0;
            for (int b = 0; b < ((P.exp & 4) ? 0 : nblk); ++b) {
                const int k = 8 * b + rr;                                        // my band row
                const unsigned tc0 = (unsigned)(k - st_c0), tc1 = (unsigned)(k - st_c1);   // tap indices (unsigned: one range test)
                const unsigned ta0 = (unsigned)(k - st_a0), ta1 = (unsigned)(k - st_a1);
                double cv0 = tc0 < (unsigned)P.tapsper ? lds_f64(pf_tab + 8u * ((unsigned)po_c0 + tc0)) : 0.0;
                double cv1 = tc1 < (unsigned)P.tapsper ? lds_f64(pf_tab + 8u * ((unsigned)po_c1 + tc1)) : 0.0;
                if (has_d) {
                    const double a0 = ta0 < (unsigned)P.tapsper ? lds_f64(dpf_tab + 8u * ((unsigned)po_a0 + ta0)) : 0.0;
                    const double a1 = ta1 < (unsigned)P.tapsper ? lds_f64(dpf_tab + 8u * ((unsigned)po_a1 + ta1)) : 0.0;
                    dmma884(cv0, cv1, a0, b0);
                    dmma884(cv0, cv1, a1, b1);
                }
                // columns 2kk, 2kk+1 stay adjacent under the XOR on bit 2 of the column index
                sts_v2f64(band + 8u * (unsigned)(k * 8 + ((2 * kk) ^ (((k >> 1) & 1) << 2))), cv0, cv1);
            }
            // ... and the tile's window has landed in the ring
            while (jwait < hi) {
                mbar_wait(&bar_full[jw_slot], jw_par);
                ++jwait;
                if (++jw_slot == P.nslot) { jw_slot = 0; jw_par ^= 1u; }
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
