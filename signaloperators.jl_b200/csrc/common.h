// Shared by every translation unit of libsignalops_cuda.so: the internal error type,
// CUDA status checking and the once-per-device dynamic shared memory opt-in.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <string>
#include <unordered_map>
#include <utility>

#include <cuda_runtime.h>

#include "../../include/signalops.h"

namespace sigops {

struct Failure {
    int code;
    std::string msg;
};

[[noreturn]] inline void fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw Failure{code, buf};
}

#define CUDA_OK(expr)                                                                       \
    do {                                                                                    \
        cudaError_t e__ = (expr);                                                           \
        if (e__ != cudaSuccess)                                                             \
            ::sigops::fail(e__ == cudaErrorMemoryAllocation ? SIGOPS_ERR_NOMEM : SIGOPS_ERR_CUDA, \
                           "CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, __LINE__,  \
                           cudaGetErrorString(e__));                                        \
    } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device, size)
inline std::mutex& attr_mutex() { static std::mutex m; return m; }
inline std::unordered_map<const void*, std::pair<uint64_t, size_t>>& attr_done() {
    static std::unordered_map<const void*, std::pair<uint64_t, size_t>> m;
    return m;
}
template <class K>
void ensure_dyn_smem(K kernel, size_t bytes) {
    int devno = 0;
    cudaGetDevice(&devno);
    std::lock_guard<std::mutex> lk(attr_mutex());
    auto& e = attr_done()[(const void*)kernel];
    if ((e.first >> (devno & 63) & 1) && e.second >= bytes) return;
    CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    e.first |= uint64_t(1) << (devno & 63);
    e.second = std::max(e.second, bytes);
}

}  // namespace sigops
