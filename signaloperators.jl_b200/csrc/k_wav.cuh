// The step either side of the path in the reference's README pipeline: `sink(x, "file.wav")` / `Signal("file.wav")`
// (src/sink.jl:139-142, src/WAV.jl:3-15 -> WAV.jl `wavwrite(data, file, Fs=round(Int,fs))` / `wavread`).
// A WAV data chunk is frame-interleaved ([frame][channel]) in the file's sample encoding; the engine works
// channel-planar (Julia column-major).  These two kernels do the transposition and the sample conversion
// on the device, so results leave the GPU already in file layout and files enter it as they are on disk:
//   k_wav_encode   planar Float64/Float32 stage output -> interleaved Float64 / Float32 / PCM16
//   k_wav_decode   interleaved Float64 / Float32 / PCM16 -> planar Float64/Float32 input buffer
// PCM16 follows WAV.jl: write round(clamp(x, -1, 1) * 32767) (ties to even, Julia `round`), read x / 32768.
// A 32 x 32 shared-memory tile makes both the planar side (lanes along frames) and the interleaved side
// (lanes along channels of consecutive frames) coalesced.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/signalops.h"

namespace sigops {

struct WavParams {
    const void* src;
    void* dst;
    int64_t nframes;
    int nch;
    int64_t ld;            // planar side: elements between channels
    int planar_dtype;      // SIGOPS_F32 / SIGOPS_F64
    int file_dtype;        // SIGOPS_F32 / SIGOPS_F64 / SIGOPS_I16
};

__device__ __forceinline__ double wav_load_planar(const void* p, int dt, int64_t i) {
    return dt == SIGOPS_F64 ? reinterpret_cast<const double*>(p)[i] : (double)reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void wav_store_planar(void* p, int dt, int64_t i, double v) {
    if (dt == SIGOPS_F64) reinterpret_cast<double*>(p)[i] = v;
    else reinterpret_cast<float*>(p)[i] = (float)v;
}
__device__ __forceinline__ void wav_store_file(void* p, int dt, int64_t i, double v) {
    if (dt == SIGOPS_F64) reinterpret_cast<double*>(p)[i] = v;
    else if (dt == SIGOPS_F32) reinterpret_cast<float*>(p)[i] = (float)v;
    else reinterpret_cast<int16_t*>(p)[i] = (int16_t)rint(fmin(fmax(v, -1.0), 1.0) * 32767.0);
}
__device__ __forceinline__ double wav_load_file(const void* p, int dt, int64_t i) {
    if (dt == SIGOPS_F64) return reinterpret_cast<const double*>(p)[i];
    if (dt == SIGOPS_F32) return (double)reinterpret_cast<const float*>(p)[i];
    return (double)reinterpret_cast<const int16_t*>(p)[i] / 32768.0;
}

// grid: (frame tiles, channel tiles); block: 32 x 8
template <bool ENCODE>
__global__ void k_wav(const WavParams P) {
    __shared__ double tile[32][33];
    const int64_t n0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    if (ENCODE) {
        for (int j = threadIdx.y; j < 32; j += 8) {            // planar read: lanes along frames
            const int64_t n = n0 + threadIdx.x;
            const int c = c0 + j;
            if (n < P.nframes && c < P.nch) tile[j][threadIdx.x] = wav_load_planar(P.src, P.planar_dtype, (int64_t)c * P.ld + n);
        }
        __syncthreads();
        for (int j = threadIdx.y; j < 32; j += 8) {            // interleaved write: lanes along channels
            const int64_t n = n0 + j;
            const int c = c0 + threadIdx.x;
            if (n < P.nframes && c < P.nch) wav_store_file(P.dst, P.file_dtype, n * P.nch + c, tile[threadIdx.x][j]);
        }
    } else {
        for (int j = threadIdx.y; j < 32; j += 8) {
            const int64_t n = n0 + j;
            const int c = c0 + threadIdx.x;
            if (n < P.nframes && c < P.nch) tile[threadIdx.x][j] = wav_load_file(P.src, P.file_dtype, n * P.nch + c);
        }
        __syncthreads();
        for (int j = threadIdx.y; j < 32; j += 8) {
            const int64_t n = n0 + threadIdx.x;
            const int c = c0 + j;
            if (n < P.nframes && c < P.nch) wav_store_planar(P.dst, P.planar_dtype, (int64_t)c * P.ld + n, tile[j][threadIdx.x]);
        }
    }
}

inline void launch_wav(bool encode, const WavParams& P, cudaStream_t st) {
    if (P.nframes <= 0 || P.nch <= 0) return;
    dim3 grid((unsigned)((P.nframes + 31) / 32), (unsigned)((P.nch + 31) / 32)), block(32, 8);
    if (encode) k_wav<true><<<grid, block, 0, st>>>(P);
    else k_wav<false><<<grid, block, 0, st>>>(P);
}

}  // namespace sigops
