// Micro-benchmark: does a DMMA stream on one warp leave the sub-partition's issue port free for
// another warp's integer / shared-memory work?  (warps 0-3: DMMA, warps 4-7: ALU + LDS)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/dmma_overlap tools/dmma_overlap.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k(double* out, int it_mma, int it_alu, int mode) {
    __shared__ float sm[4096];
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
    __syncthreads();
    if (warp < 4) {
        if (!(mode & 1)) return;
        double c[16][2];
        for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = threadIdx.x + i;
        double a = 1.0 + threadIdx.x * 1e-9, b = 0.5;
        for (int it = 0; it < it_mma; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
        double s = 0;
        for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else {
        if (!(mode & 2)) return;
        unsigned x = threadIdx.x, y = 12345u;
        float acc = 0;
        for (int it = 0; it < it_alu; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                x = x * 1664525u + 1013904223u;
                y ^= x >> 7;
                acc += sm[(x >> 20) & 4095];
            }
        }
        out[blockIdx.x * blockDim.x + threadIdx.x] = acc + y;
    }
}

float run(double* out, int a, int b, int mode) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<148, 256>>>(out, a, b, mode); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<<<148, 256>>>(out, a, b, mode); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    double* out; cudaMalloc(&out, sizeof(double) * 148 * 256);
    const int it_mma = 20000;
    for (int it_alu : {5000, 10000, 20000, 40000}) {
        float t1 = run(out, it_mma, it_alu, 1), t2 = run(out, it_mma, it_alu, 2), t3 = run(out, it_mma, it_alu, 3);
        printf("alu iters %6d: dmma only %.3f ms, alu only %.3f ms, both %.3f ms (sum %.3f, max %.3f)\n", it_alu, t1, t2, t3, t1 + t2,
               t1 > t2 ? t1 : t2);
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
