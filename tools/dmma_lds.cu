// Micro-benchmark: how fast does one SM issue DMMA.8x8x4 when the A/B operands of every instruction are
// fresh shared-memory loads (the k_fir_tmap inner loop), for 1..4 warps per sub-partition and NL loads per
// 8 DMMAs?  build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/dmma_lds tools/dmma_lds.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ double lds(unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NL>   // A loads per k-step (8 DMMAs); B is always one load
__global__ void k(double* out, int iters) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = 1e-3 * i;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned base = (unsigned)__cvta_generic_to_shared(sm) + lane * 8 + warp * 2048;
    double c[8][2];
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
    double a0[8], a1[8], b0, b1;
    for (int i = 0; i < 8; ++i) a0[i] = lds(base + (i % (NL ? NL : 1)) * 1024);
    b0 = lds(base + 512);
    unsigned off = 0;
    for (int it = 0; it < iters; it += 2) {
        off = (off + 256) & 0x3fff;
        for (int i = 0; i < 8; ++i) a1[i] = i < NL ? lds(base + ((off + i * 1024) & 0xffff)) : a0[i];
        b1 = lds(base + ((off + 512) & 0xffff));
        for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a0[i], b0);
        off = (off + 256) & 0x3fff;
        for (int i = 0; i < 8; ++i) a0[i] = i < NL ? lds(base + ((off + i * 1024) & 0xffff)) : a1[i];
        b0 = lds(base + ((off + 512) & 0xffff));
        for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a1[i], b1);
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NL>
void run(int sms, double* out) {
    const int iters = 20000;
    cudaFuncSetAttribute(k<NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    for (int threads : {128, 256, 384, 512}) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<NL><<<sms, threads, 160 * 1024>>>(out, iters);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        k<NL><<<sms, threads, 160 * 1024>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double dm = (double)sms * (threads / 32) * iters * 8;
        printf("A loads/k-step %d, %d warps/SMSP: %.2f T FMA/s, %.1f cycles per DMMA per SMSP (at 1.965 GHz)  %s\n", NL, threads / 128,
               dm * 256 / (ms * 1e-3) / 1e12, ms * 1e-3 * 1.965e9 / (dm / (sms * 4)), cudaGetErrorString(cudaGetLastError()));
    }
}

// the same loop held for `seconds`: the sustained rate under the board's power cap (clocks drop)
void sustained(int sms, double* out, double seconds) {
    const int iters = 20000, threads = 256;
    cudaFuncSetAttribute(k<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double dm = (double)sms * (threads / 32) * iters * 8;
    float ms1;
    cudaEventRecord(e0); k<8><<<sms, threads, 160 * 1024>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms1, e0, e1);
    const int n = (int)(seconds * 1e3 / ms1) + 1;
    const int tail = n / 4 > 0 ? n / 4 : 1;
    for (int i = 0; i < n - tail; ++i) k<8><<<sms, threads, 160 * 1024>>>(out, iters);
    cudaEventRecord(e0);
    for (int i = 0; i < tail; ++i) k<8><<<sms, threads, 160 * 1024>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("sustained %.1f s, 2 warps/SMSP, 9 loads per 8 DMMAs: %.2f T FMA/s over the last quarter (burst launch: %.2f)\n", seconds,
           dm * tail * 256 / (ms * 1e-3) / 1e12, dm * 256 / (ms1 * 1e-3) / 1e12);
}

int main(int argc, char** argv) {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* out; cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 1024);
    if (argc > 1) { sustained(p.multiProcessorCount, out, atof(argv[1])); return 0; }
    run<0>(p.multiProcessorCount, out);
    run<2>(p.multiProcessorCount, out);
    run<4>(p.multiProcessorCount, out);
    run<8>(p.multiProcessorCount, out);
    return 0;
}
