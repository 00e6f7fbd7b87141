// Shared helpers for libscb200 (sm_100a).  Not part of the public ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/scb200.h"

namespace scb {

constexpr int kWarp = 32;
constexpr int kNumSM = 148;  // B200: 2 dies x 74 SMs (grid sizing only)

void set_last_cuda_error(cudaError_t e, const char* file, int line);
void count_launches(int n);  // process-wide kernel launch counter (scb_launch_count)
// stream-ordered scratch from the library-private memory pool of the current device (pipeline.cu)
int pool_alloc(void** p, size_t bytes, cudaStream_t st);
void pool_free(void* p, cudaStream_t st);

#define SCB_CUDA(expr)                                              \
    do {                                                            \
        cudaError_t _e = (expr);                                    \
        if (_e != cudaSuccess) {                                    \
            ::scb::set_last_cuda_error(_e, __FILE__, __LINE__);     \
            return SCB_ERR_CUDA;                                    \
        }                                                           \
    } while (0)

// one kernel launch precedes every SCB_LAUNCH_CHECK(); extra launches are counted explicitly
#define SCB_LAUNCH_CHECK()             \
    do {                               \
        ::scb::count_launches(1);      \
        SCB_CUDA(cudaGetLastError());  \
    } while (0)

#define SCB_TRY(expr)                 \
    do {                              \
        int _s = (expr);              \
        if (_s != SCB_OK) return _s;  \
    } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// atomic max for non-negative doubles (bit pattern order == numeric order)
__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
    atomicMax(reinterpret_cast<unsigned long long*>(addr),
              static_cast<unsigned long long>(__double_as_longlong(v)));
}

// counter-based RNG: uniform in (-1, 1), independent of the launch geometry
__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ inline double uniform_pm1(uint64_t seed, uint64_t idx) {
    uint64_t r = splitmix64(seed ^ splitmix64(idx));
    return (double)(r >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}

// bump allocator over a caller-provided workspace
struct Arena {
    char* base;
    size_t size;
    size_t off;
    Arena(void* p, size_t s) : base(static_cast<char*>(p)), size(s), off(0) {}
    template <typename T>
    T* take(size_t count) {
        size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
        if (base != nullptr && off + bytes > size) { off += bytes; return nullptr; }
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += bytes;
        return p;
    }
    bool ok() const { return off <= size; }
};

}  // namespace scb
