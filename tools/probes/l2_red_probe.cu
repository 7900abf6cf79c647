// Probe: how fast can B200's L2 absorb fp32 reductions of 256-byte rows (the per-edge scatter-add of the RGCN
// forward)?  Three ways of adding a 64-float row into out[row] for random rows inside a window of W MB:
//   0  st.global.v4.f32        (plain stores, reference for the write path)
//   1  red.global.add.v4.f32   (REDG.E.ADD.F32x4, two rows per warp instruction)
//   2  cp.reduce.async.bulk.global.shared::cta.add.f32 (TMA bulk reduction, one 256-byte row per instruction)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l2_red_probe l2_red_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int MODE>
__global__ void __launch_bounds__(256) k_probe(float* out, uint32_t rows_in_window, uint32_t ops_per_warp, uint32_t seed) {
    __shared__ __align__(128) float stage[8][2][64];        // per warp: two rows
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t gw = blockIdx.x * 8 + warp;
    for (int i = lane; i < 128; i += 32) (&stage[warp][0][0])[i] = 1.0f;
    __syncwarp();
    const float4 v = make_float4(1.f, 1.f, 1.f, 1.f);
    for (uint32_t i = 0; i < ops_per_warp; ++i) {
        const uint32_t r0 = hash32(seed + gw * 1000003u + 2 * i) % rows_in_window;
        const uint32_t r1 = hash32(seed + gw * 1000003u + 2 * i + 1) % rows_in_window;
        if (MODE == 0) {
            float4* p = reinterpret_cast<float4*>(out + (size_t)(lane < 16 ? r0 : r1) * 64) + (lane & 15);
            *p = v;
        } else if (MODE == 1) {
            float* p = out + (size_t)(lane < 16 ? r0 : r1) * 64 + (lane & 15) * 4;
            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        } else {
            if (lane < 2) {
                const uint32_t src = (uint32_t)__cvta_generic_to_shared(&stage[warp][lane][0]);
                float* dst = out + (size_t)(lane == 0 ? r0 : r1) * 64;
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 256;" ::"l"(dst), "r"(src) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if ((i & 7) == 7) asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
            }
        }
    }
    if (MODE == 2 && lane < 2) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int MODE>
float run(float* out, uint32_t rows, uint32_t ops, int grid) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k_probe<MODE><<<grid, 256>>>(out, rows, ops / 4, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k_probe<MODE><<<grid, 256>>>(out, rows, ops, 7);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
    return ms;
}

int main() {
    const size_t total_rows = 1700000;                       // 435 MB of fp32 rows
    float* out;
    cudaMalloc(&out, total_rows * 256);
    cudaMemset(out, 0, total_rows * 256);
    const int grid = 148 * 8;
    const uint32_t ops = 2048;                               // per warp: 2 rows each -> grid * 8 * ops * 2 rows
    const double rows_done = (double)grid * 8 * ops * 2;
    const uint32_t windows_mb[] = {16, 48, 96, 435};
    for (uint32_t w : windows_mb) {
        uint32_t rows = (uint32_t)((size_t)w * 1000000 / 256);
        if (rows > total_rows) rows = total_rows;
        const float t0 = run<0>(out, rows, ops, grid), t1 = run<1>(out, rows, ops, grid), t2 = run<2>(out, rows, ops, grid);
        printf("window %4u MB: st.v4 %.3f ms (%.0f GB/s)  red.v4.f32 %.3f ms (%.0f GB/s)  bulk reduce %.3f ms (%.0f GB/s)\n", w, t0,
               rows_done * 256 / t0 / 1e6, t1, rows_done * 256 / t1 / 1e6, t2, rows_done * 256 / t2 / 1e6);
    }
    return 0;
}
