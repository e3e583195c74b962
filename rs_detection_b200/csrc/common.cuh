// common.cuh -- shared host/device helpers for librsdet (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rsdet.h"

namespace rsdet {

constexpr int kNumSMs = 148;  // B200

// number of kernel launches issued by the library (bench.py reports it as gpu_launches)
extern unsigned long long g_launches;
inline void count_launch(int n = 1) { g_launches += (unsigned long long)n; }

inline int cuda_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? RSDET_OK : (int)e;
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// bump allocator over the caller's workspace
struct Workspace {
    char* base;
    size_t size, used;
    Workspace(void* p, size_t s) : base((char*)p), size(s), used(0) {}
    template <typename T>
    T* take(size_t count) {
        size_t bytes = align256(count * sizeof(T));
        T* r = (T*)(base + used);
        used += bytes;
        return r;
    }
    bool ok() const { return used <= size && (base != nullptr || used == 0); }
};

template <typename T>
inline size_t ws_bytes(size_t count) { return align256(count * sizeof(T)); }

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

}  // namespace rsdet
