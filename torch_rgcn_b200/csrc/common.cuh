// Shared host/device helpers for the rgcn_b200 CUDA library.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <atomic>
#include "../../include/rgcn_b200.h"

namespace rgcn {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

#define RGCN_CHECK_CUDA(expr)                                                              \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            rgcn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return RGCN_ERR_CUDA;                                                          \
        }                                                                                  \
    } while (0)

#define RGCN_REQUIRE(cond, code, ...)                                                      \
    do {                                                                                   \
        if (!(cond)) {                                                                     \
            rgcn::set_error(__VA_ARGS__);                                                  \
            return (code);                                                                 \
        }                                                                                  \
    } while (0)

// every kernel launch goes through this so rgcn_launch_count() is honest
#define RGCN_LAUNCH(kernel, grid, block, smem, stream, ...)                                \
    do {                                                                                   \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                        \
        rgcn::g_launches.fetch_add(1, std::memory_order_relaxed);                          \
        RGCN_CHECK_CUDA(cudaGetLastError());                                               \
    } while (0)

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// bump allocator over the caller's workspace
struct Carver {
    char* base;
    size_t off = 0;
    explicit Carver(void* p) : base(static_cast<char*>(p)) {}
    template <typename T>
    T* take(size_t n) {
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += align_up(n * sizeof(T));
        return p;
    }
};

static inline int grid_for(int64_t n, int block) { return (int)((n + block - 1) / block); }
constexpr int kNumSMs = 148;   // B200

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }

}  // namespace rgcn
