// Micro-benchmark: FP64 tensor-core (mma.sync f64) versus FP64 FMA throughput on this GPU.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/dmma_peak tools/dmma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dmma884(double* out, int iters, double a0, double b0) {
    double c[8][2];
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma16816(double* out, int iters, double a0, double b0) {
    double c[4][4];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x * 1e-3 + i + j;
    double a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = a0 + threadIdx.x * 1e-9 + i;
    for (int i = 0; i < 4; ++i) b[i] = b0 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                           "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
    double s = 0;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


__global__ void k_dmma1684(double* out, int iters, double a0, double b0) {
    double c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x * 1e-3 + i + j;
    double a[2] = {a0 + threadIdx.x * 1e-9, a0 + 1}, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a[0]), "d"(a[1]), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma1688(double* out, int iters, double a0, double b0) {
    double c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x * 1e-3 + i + j;
    double a[4], b[2] = {b0, b0 + 1};
    for (int i = 0; i < 4; ++i) a[i] = a0 + threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dfma(double* out, int iters, double a, double b) {
    double x[8];
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
    double s = 0;
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float timeit(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    const int iters = 20000;
    for (int threads : {128, 256, 512, 1024}) {
        for (int bps : {1, 2}) {
            const int blocks = sms * bps;
            const double warps = (double)blocks * threads / 32;
            float t1 = timeit([&] { k_dmma884<<<blocks, threads>>>(out, iters, 1.0000001, 0.5); });
            float t2 = timeit([&] { k_dmma16816<<<blocks, threads>>>(out, iters, 1.0000001, 0.5); });
            float t3 = timeit([&] { k_dfma<<<blocks, threads>>>(out, iters, 1.0000001, 0.5); });
            float t4 = timeit([&] { k_dmma1684<<<blocks, threads>>>(out, iters, 1.0000001, 0.5); });
            float t5 = timeit([&] { k_dmma1688<<<blocks, threads>>>(out, iters, 1.0000001, 0.5); });
            printf("threads %4d x %d blocks/SM: m8n8k4 %.2f  m16n8k4 %.2f  m16n8k8 %.2f  m16n8k16 %.2f  dfma %.2f  (T FMA/s)\n", threads, bps,
                   warps * iters * 8 * 256 / (t1 * 1e-3) / 1e12, warps * iters * 8 * 512 / (t4 * 1e-3) / 1e12,
                   warps * iters * 8 * 1024 / (t5 * 1e-3) / 1e12, warps * iters * 4 * 2048 / (t2 * 1e-3) / 1e12,
                   warps * iters * 8 * 32 / (t3 * 1e-3) / 1e12);
        }
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
