// Micro-benchmark: HBM copy bandwidth with the IIR kernel's access pattern
// (each warp streams 32 chunks `L` frames apart, COLS frames of each per stage).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void cp_async8(double* s, const double* g) {
    unsigned d = (unsigned)__cvta_generic_to_shared(s);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async16(double* s, const double* g) {
    unsigned d = (unsigned)__cvta_generic_to_shared(s);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g) : "memory");
}

template <int COLS, int NBUF>
__global__ void __launch_bounds__(128) k_pattern(const double* __restrict__ x, double* __restrict__ y, int64_t L, int64_t nrow_chunks) {
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int PITCH = COLS + 1;
    double* tile = smem + (size_t)warp * NBUF * 32 * PITCH;
    const int64_t chunk0 = ((int64_t)blockIdx.x * 4 + warp) * 32;
    if (chunk0 >= nrow_chunks) return;
    constexpr int RPI = 32 / COLS > 0 ? 32 / COLS : 1;       // rows per instruction
    constexpr int IPR = COLS / 32 > 0 ? COLS / 32 : 1;       // instructions per row
    const int crow = (COLS < 32) ? lane / COLS : 0, ccol = (COLS < 32) ? lane % COLS : lane;
    const int64_t nst = L / COLS;
    auto issue = [&](int64_t s, double* buf) {
#pragma unroll
        for (int r = 0; r < 32; r += RPI)
#pragma unroll
            for (int q = 0; q < IPR; ++q)
                cp_async8(buf + (r + crow) * PITCH + ccol + 32 * q, x + (chunk0 + r + crow) * L + s * COLS + ccol + 32 * q);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue(0, tile);
    for (int64_t s = 0; s < nst; ++s) {
        double* buf = tile + (s % NBUF) * 32 * PITCH;
        if (NBUF > 1 && s + 1 < nst) { issue(s + 1, tile + ((s + 1) % NBUF) * 32 * PITCH); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 32; r += RPI)
#pragma unroll
            for (int q = 0; q < IPR; ++q)
                y[(chunk0 + r + crow) * L + s * COLS + ccol + 32 * q] = buf[(r + crow) * PITCH + ccol + 32 * q] * 1.5;
        __syncwarp();
        if (NBUF == 1 && s + 1 < nst) issue(s + 1, tile);
    }
}

// per-lane streams moved by TMA bulk copies (the k_iir_tma data path, no compute)
__device__ __forceinline__ unsigned su32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int COLS>
__global__ void __launch_bounds__(192, 1) k_bulk(const double* __restrict__ x, double* __restrict__ y, int64_t L, int64_t nchunks) {
    extern __shared__ __align__(128) unsigned char raw[];
    constexpr int PITCH = COLS + 2;
    const int tid = threadIdx.x;
    double* st0 = reinterpret_cast<double*>(raw) + (size_t)tid * PITCH;
    double* st1 = st0 + (size_t)192 * PITCH;
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<double*>(raw) + (size_t)192 * 2 * PITCH) + tid * 2;
    const int64_t g = (int64_t)blockIdx.x * 192 + tid;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&bars[0])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&bars[1])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (g >= nchunks) return;
    const double* src = x + g * L;
    double* dst = y + g * L;
    const int64_t nst = L / COLS;
    unsigned par = 0;
    auto load = [&](int64_t h) {
        const unsigned b = h & 1;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(&bars[b])), "r"(COLS * 8) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(su32(b ? st1 : st0)),
                     "l"(src + h * COLS), "r"(COLS * 8), "r"(su32(&bars[b])) : "memory");
    };
    load(0);
    for (int64_t h = 0; h < nst; ++h) {
        const unsigned b = h & 1;
        double* buf = b ? st1 : st0;
        asm volatile("{\n.reg .pred p;\nLW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra LD;\nbra LW;\nLD:\n}\n" ::"r"(su32(&bars[b])), "r"((par >> b) & 1u) : "memory");
        par ^= 1u << b;
        if (h + 1 < nst) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            load(h + 1);
        }
        for (int i = 0; i < COLS; i += 2) {
            double2 v = *reinterpret_cast<double2*>(buf + i);
            v.x *= 1.5; v.y *= 1.5;
            *reinterpret_cast<double2*>(buf + i) = v;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + h * COLS), "r"(su32(buf)), "r"(COLS * 8) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int COLS>
void run_bulk(const double* x, double* y, int64_t n, int64_t L) {
    const int64_t chunks = n / L;
    const int blocks = (int)((chunks + 191) / 192);
    const size_t smem = (size_t)192 * 2 * (COLS + 2) * 8 + 192 * 16;
    cudaFuncSetAttribute(k_bulk<COLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k_bulk<COLS><<<blocks, 192, smem>>>(x, y, L, chunks);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    printf("per-lane TMA bulk %4d B stages   L=%6lld blocks=%5d smem/blk=%6zu : %.3f ms  %.0f GB/s  (%s)\n", COLS * 8, (long long)L, blocks, smem, best,
           2.0 * n * 8 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}

template <int COLS, int NBUF>
void run(const double* x, double* y, int64_t n, int64_t L, const char* name) {
    const int64_t chunks = n / L;
    const int blocks = (int)((chunks / 32 + 3) / 4);
    const size_t smem = (size_t)4 * NBUF * 32 * (COLS + 1) * sizeof(double);
    cudaFuncSetAttribute(k_pattern<COLS, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k_pattern<COLS, NBUF><<<blocks, 128, smem>>>(x, y, L, chunks);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    printf("%-28s L=%6lld blocks=%5d smem/blk=%6zu : %.3f ms  %.0f GB/s  (%s)\n", name, (long long)L, blocks, smem, best,
           2.0 * n * 8 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}

__global__ void k_plain(const double2* __restrict__ x, double2* __restrict__ y, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { double2 v = x[i]; v.x *= 1.5; v.y *= 1.5; y[i] = v; }
}

int main() {
    const int64_t n = 245760000;
    double *x, *y;
    cudaMalloc(&x, n * 8); cudaMalloc(&y, n * 8);
    cudaMemset(x, 0, n * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0); k_plain<<<148 * 16, 512>>>((const double2*)x, (double2*)y, n / 2); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep == 2) printf("plain grid-stride copy: %.3f ms %.0f GB/s\n", ms, 2.0 * n * 8 / ms / 1e6);
    }
    for (int64_t L : {8704LL, 4352LL, 2176LL}) {
        run_bulk<32>(x, y, n, L);
        run_bulk<64>(x, y, n, L);
        run_bulk<128>(x, y, n, L);
    }
    for (int64_t L : {3776LL}) {
        run<16, 2>(x, y, n, L, "16 cols (128B), 2 buffers");
        run<32, 1>(x, y, n, L, "32 cols (256B), 1 buffer");
        run<32, 2>(x, y, n, L, "32 cols (256B), 2 buffers");
        run<64, 1>(x, y, n, L, "64 cols (512B), 1 buffer");
        run<64, 2>(x, y, n, L, "64 cols (512B), 2 buffers");
        run<128, 1>(x, y, n, L, "128 cols (1KB), 1 buffer");
    }
    return 0;
}
