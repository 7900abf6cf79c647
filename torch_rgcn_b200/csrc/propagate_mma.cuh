// Tensor-core relation-batched kernels for bf16 features and 16x16 weight blocks (AM / SYN shapes).
//
// A CTA owns one 1024-edge chunk of one relation.  Each warp keeps the bf16 fragments of "its" four 16x16
// weight blocks in registers for the whole chunk and streams 16-edge tiles through a private 3-stage
// cp.async ring: the 16 gathered feature-row slices (16 x 128 B) land in shared memory with a 16-byte XOR
// swizzle, ldmatrix feeds them to mma.sync.m16n8k16 (bf16 x bf16 -> fp32), the result is scaled by the
// per-edge weight, packed to bf16, bounced through the same tile buffer and written to the message rows
// with coalesced 16-byte stores.  No CTA-wide barrier after the index prologue: warps are independent.
//
// Each block is a true dense 16x16 GEMM tile shared by all edges of the relation, which is the one place the
// path is GEMM-shaped; everything else stays a gather/scatter.  mma.sync (HMMA) rather than tcgen05: the
// tiles are 16 edges x 16 x 16, far below the 128-row UMMA atom, and the kernel is gather-bound, not
// tensor-bound.
#pragma once
#include "common.cuh"
#include "propagate_fast.cuh"

namespace rgcn {

constexpr int kMmaStages = 3;
constexpr int kTileBytes = 16 * 128;          // 16 edges x 128-byte row slice (4 blocks of 16 bf16)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// byte offset of 16-byte chunk `chunk` of row `row` inside a swizzled 16 x 128 B tile
__device__ __forceinline__ int tile_off(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

// msg[slot(e), bg*64 .. bg*64+63] = bf16( val_e * X[src_e, bg*64 ..] @ blockdiag(W_p[4bg .. 4bg+3]) )
__global__ void __launch_bounds__(256) k_rel_mma_fwd(RelArgs A, const __nv_bfloat16* __restrict__ X,
                                                     __nv_bfloat16* __restrict__ msg) {
    extern __shared__ __align__(128) unsigned char smem_mma[];
    unsigned char* smem = smem_mma;
    const int c = blockIdx.x;
    if (c >= A.chunkptr[A.num_rels]) return;
    int p, e0, e1;
    chunk_lookup(A, c, p, e0, e1);
    const int n = e1 - e0;
    int32_t* s_src = reinterpret_cast<int32_t*>(smem);
    int32_t* s_slot = s_src + RGCN_CHUNK_EDGES;
    float* s_val = reinterpret_cast<float*>(s_slot + RGCN_CHUNK_EDGES);
    unsigned char* rings = reinterpret_cast<unsigned char*>(s_val + RGCN_CHUNK_EDGES);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        s_src[i] = A.gather[e0 + i]; s_slot[i] = A.slot[e0 + i]; s_val[i] = A.val[e0 + i];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int NG = A.nb >> 2;                       // block groups of four 16x16 blocks (128 B of every row)
    const int bg = warp % NG, wsub = warp / NG, nsub = 8 / NG;
    const size_t row_bytes = (size_t)A.nb * 32;     // I == O == nb * 16 bf16

    // weight fragments (col-major B operand): b0 = W[2t..2t+1][n], b1 = W[2t+8..2t+9][n], n = 8h + g
    uint32_t bfrag[4][2][2];
    {
        const float* wp = A.W + ((size_t)p * A.nb + (size_t)bg * 4) * 256;
#pragma unroll
        for (int kb = 0; kb < 4; ++kb)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float* wb = wp + kb * 256 + h * 8 + g;
                bfrag[kb][h][0] = pack_bf16x2(__ldg(wb + (2 * t) * 16), __ldg(wb + (2 * t + 1) * 16));
                bfrag[kb][h][1] = pack_bf16x2(__ldg(wb + (2 * t + 8) * 16), __ldg(wb + (2 * t + 9) * 16));
            }
    }
    unsigned char* ring = rings + (size_t)warp * kMmaStages * kTileBytes;
    const int ntiles = (n + 15) >> 4;
    const unsigned char* Xb = reinterpret_cast<const unsigned char*>(X) + (size_t)bg * 128;
    unsigned char* Mb = reinterpret_cast<unsigned char*>(msg) + (size_t)bg * 128;

    auto issue = [&](int k) {
        const int tile = wsub + k * nsub;
        if (tile < ntiles) {
            unsigned char* st = ring + (k % kMmaStages) * kTileBytes;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int row = (lane >> 3) + 4 * it, chunk = lane & 7, le = tile * 16 + row;
                const bool ok = le < n;
                cp_async16(st + tile_off(row, chunk), Xb + (size_t)s_src[ok ? le : 0] * row_bytes + chunk * 16, ok);
            }
        }
        cp_async_commit();
    };
    issue(0);
    issue(1);
    for (int k = 0; wsub + k * nsub < ntiles; ++k) {
        issue(k + 2);
        cp_async_wait<2>();
        __syncwarp();
        unsigned char* st = ring + (k % kMmaStages) * kTileBytes;
        const int tile = wsub + k * nsub;
        float acc[4][2][4];
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[kb][h][q] = 0.f;
            uint32_t a[4];
            const int row = (lane & 7) + ((lane >> 3) & 1) * 8, chunk = kb * 2 + (lane >> 4);
            ldmatrix_x4(a, smem_u32(st + tile_off(row, chunk)));
            mma_bf16_16816(acc[kb][0], a, bfrag[kb][0][0], bfrag[kb][0][1]);
            mma_bf16_16816(acc[kb][1], a, bfrag[kb][1][0], bfrag[kb][1][1]);
        }
        __syncwarp();                                   // tile fully consumed; reuse it to transpose the result
        const int le0 = tile * 16 + g, le1 = le0 + 8;
        const float v0 = le0 < n ? s_val[le0] : 0.f, v1 = le1 < n ? s_val[le1] : 0.f;
#pragma unroll
        for (int kb = 0; kb < 4; ++kb)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int chunk = kb * 2 + h;
                *reinterpret_cast<uint32_t*>(st + tile_off(g, chunk) + t * 4) =
                    pack_bf16x2(acc[kb][h][0] * v0, acc[kb][h][1] * v0);
                *reinterpret_cast<uint32_t*>(st + tile_off(g + 8, chunk) + t * 4) =
                    pack_bf16x2(acc[kb][h][2] * v1, acc[kb][h][3] * v1);
            }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int row = (lane >> 3) + 4 * it, chunk = lane & 7, le = tile * 16 + row;
            if (le < n) {
                const uint4 v = *reinterpret_cast<const uint4*>(st + tile_off(row, chunk));
                *reinterpret_cast<uint4*>(Mb + (size_t)s_slot[le] * row_bytes + chunk * 16) = v;
            }
        }
        __syncwarp();                                   // before a later issue() refills this stage
    }
    cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------------
// Fused backward for the same shapes.  One relation-major pass reads X[src] (bf16) and G[dst] (fp32) once and
// produces BOTH
//   msg'[sslot(e)] = bf16( (val_e G[dst_e]) @ blockdiag(W_p)^T )      -> summed per source row by k_row_sum
//   gW_p          += X[src]^T (val_e G[dst_e])                         -> fp32 fragments kept in registers for
//                                                                         the whole chunk, one atomic flush
// G rows are staged as fp32, scaled by val and rounded to bf16 in shared memory (fp32 accumulation in the MMA).
// ------------------------------------------------------------------------------------------------------
constexpr int kBwdStages = 3;
constexpr int kGTileBytes = 16 * 256;              // 16 edges x 64 fp32 outputs
constexpr int kBwdStageBytes = kTileBytes + kGTileBytes;
constexpr int kBwdWarpBytes = kBwdStages * kBwdStageBytes + kTileBytes;   // ring + bf16 G tile

__global__ void __launch_bounds__(256, 1) k_rel_mma_bwd(RelArgs A, const __nv_bfloat16* __restrict__ X,
                                                        const float* __restrict__ G, __nv_bfloat16* __restrict__ msg,
                                                        float* __restrict__ gW) {
    extern __shared__ __align__(128) unsigned char smem_bwd[];
    unsigned char* smem = smem_bwd;
    const int c = blockIdx.x;
    if (c >= A.chunkptr[A.num_rels]) return;
    int p, e0, e1;
    chunk_lookup(A, c, p, e0, e1);
    const int n = e1 - e0;
    int32_t* s_src = reinterpret_cast<int32_t*>(smem);
    int32_t* s_dst = s_src + RGCN_CHUNK_EDGES;
    int32_t* s_slot = s_dst + RGCN_CHUNK_EDGES;
    float* s_val = reinterpret_cast<float*>(s_slot + RGCN_CHUNK_EDGES);
    unsigned char* rings = reinterpret_cast<unsigned char*>(s_val + RGCN_CHUNK_EDGES);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        s_src[i] = A.gather[e0 + i]; s_dst[i] = A.other[e0 + i]; s_val[i] = A.val[e0 + i];
        if (msg) s_slot[i] = A.slot[e0 + i];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int NG = A.nb >> 2;
    const int bg = warp % NG, wsub = warp / NG, nsub = 8 / NG;
    const size_t xrow_bytes = (size_t)A.nb * 32;    // bf16 rows of X and msg'
    const size_t grow_bytes = (size_t)A.nb * 64;    // fp32 rows of G

    // W^T fragments for msg' = Gb @ W^T:  B[k = j][n = i] = W[i][j];  b0 = W[n][2t..2t+1], b1 = W[n][2t+8..2t+9]
    uint32_t wt[4][2][2];
    if (msg) {
        const float* wp = A.W + ((size_t)p * A.nb + (size_t)bg * 4) * 256;
#pragma unroll
        for (int kb = 0; kb < 4; ++kb)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float* wr = wp + kb * 256 + (h * 8 + g) * 16;
                wt[kb][h][0] = pack_bf16x2(__ldg(wr + 2 * t), __ldg(wr + 2 * t + 1));
                wt[kb][h][1] = pack_bf16x2(__ldg(wr + 2 * t + 8), __ldg(wr + 2 * t + 9));
            }
    }
    float gacc[4][2][4];
#pragma unroll
    for (int kb = 0; kb < 4; ++kb)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int q = 0; q < 4; ++q) gacc[kb][h][q] = 0.f;

    unsigned char* ring = rings + (size_t)warp * kBwdWarpBytes;
    unsigned char* gb = ring + kBwdStages * kBwdStageBytes;          // bf16 (val * G) tile, swizzled
    const int ntiles = (n + 15) >> 4;
    const unsigned char* Xb = reinterpret_cast<const unsigned char*>(X) + (size_t)bg * 128;
    const unsigned char* Gp = reinterpret_cast<const unsigned char*>(G) + (size_t)bg * 256;
    unsigned char* Mb = reinterpret_cast<unsigned char*>(msg) + (size_t)bg * 128;

    auto issue = [&](int k) {
        const int tile = wsub + k * nsub;
        if (tile < ntiles) {
            unsigned char* st = ring + (k % kBwdStages) * kBwdStageBytes;
            if (gW) {
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int row = (lane >> 3) + 4 * it, chunk = lane & 7, le = tile * 16 + row;
                    const bool ok = le < n;
                    cp_async16(st + tile_off(row, chunk), Xb + (size_t)s_src[ok ? le : 0] * xrow_bytes + chunk * 16, ok);
                }
            }
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int row = (lane >> 4) + 2 * it, c16 = lane & 15, le = tile * 16 + row;
                const bool ok = le < n;
                cp_async16(st + kTileBytes + row * 256 + c16 * 16, Gp + (size_t)s_dst[ok ? le : 0] * grow_bytes + c16 * 16, ok);
            }
        }
        cp_async_commit();
    };
    issue(0);
    issue(1);
    for (int k = 0; wsub + k * nsub < ntiles; ++k) {
        issue(k + 2);
        cp_async_wait<2>();
        __syncwarp();
        unsigned char* st = ring + (k % kBwdStages) * kBwdStageBytes;
        const int tile = wsub + k * nsub;
        // ---- fp32 G rows -> val-scaled bf16 tile (each lane: 4 floats of one row per step)
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int row = (lane >> 4) + 2 * it, c4 = lane & 15, le = tile * 16 + row;
            const float v = le < n ? s_val[le] : 0.f;
            const float4 f = *reinterpret_cast<const float4*>(st + kTileBytes + row * 256 + c4 * 16);
            *reinterpret_cast<uint2*>(gb + tile_off(row, c4 >> 1) + (c4 & 1) * 8) =
                make_uint2(pack_bf16x2(f.x * v, f.y * v), pack_bf16x2(f.z * v, f.w * v));
        }
        __syncwarp();
        float macc[4][2][4];
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
            const int row = (lane & 7) + ((lane >> 3) & 1) * 8, chunk = kb * 2 + (lane >> 4);
            const uint32_t gaddr = smem_u32(gb + tile_off(row, chunk));
            if (msg) {
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int q = 0; q < 4; ++q) macc[kb][h][q] = 0.f;
                uint32_t a[4];
                ldmatrix_x4(a, gaddr);                                  // A = Gb[edges][outputs of block kb]
                mma_bf16_16816(macc[kb][0], a, wt[kb][0][0], wt[kb][0][1]);
                mma_bf16_16816(macc[kb][1], a, wt[kb][1][0], wt[kb][1][1]);
            }
            if (gW) {
                uint32_t bt[4], xa[4];
                ldmatrix_x4_trans(bt, gaddr);                           // B[k = edge][n = output]: {h0:b0,b1, h1:b0,b1}
                const int xrow = (lane & 7) + ((lane >> 4) & 1) * 8, xchunk = kb * 2 + ((lane >> 3) & 1);
                ldmatrix_x4_trans(xa, smem_u32(st + tile_off(xrow, xchunk)));   // A[m = input][k = edge]
                mma_bf16_16816(gacc[kb][0], xa, bt[0], bt[1]);
                mma_bf16_16816(gacc[kb][1], xa, bt[2], bt[3]);
            }
        }
        __syncwarp();                                   // X tile consumed: reuse it to transpose msg'
        if (msg) {
#pragma unroll
            for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int chunk = kb * 2 + h;
                    *reinterpret_cast<uint32_t*>(st + tile_off(g, chunk) + t * 4) = pack_bf16x2(macc[kb][h][0], macc[kb][h][1]);
                    *reinterpret_cast<uint32_t*>(st + tile_off(g + 8, chunk) + t * 4) = pack_bf16x2(macc[kb][h][2], macc[kb][h][3]);
                }
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int row = (lane >> 3) + 4 * it, chunk = lane & 7, le = tile * 16 + row;
                if (le < n) {
                    const uint4 v = *reinterpret_cast<const uint4*>(st + tile_off(row, chunk));
                    *reinterpret_cast<uint4*>(Mb + (size_t)s_slot[le] * xrow_bytes + chunk * 16) = v;
                }
            }
        }
        __syncwarp();
    }
    cp_async_wait<0>();
    if (!gW) return;
    // ---- reduce the per-warp weight-gradient fragments over the warps that share a block group, then flush
    __syncthreads();
    float* red = reinterpret_cast<float*>(rings);      // [8 warps][4 blocks][16][16]
    {
        float* mine = red + (size_t)warp * 1024;
#pragma unroll
        for (int kb = 0; kb < 4; ++kb)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float* b = mine + kb * 256 + h * 8 + 2 * t;
                b[g * 16] = gacc[kb][h][0]; b[g * 16 + 1] = gacc[kb][h][1];
                b[(g + 8) * 16] = gacc[kb][h][2]; b[(g + 8) * 16 + 1] = gacc[kb][h][3];
            }
    }
    __syncthreads();
    float* dst = gW + (size_t)p * A.nb * 256;
    for (int i = threadIdx.x; i < NG * 1024; i += blockDim.x) {
        const int grp = i >> 10, el = i & 1023;
        float s = 0.f;
        for (int w = 0; w < nsub; ++w) s += red[(size_t)(grp + w * NG) * 1024 + el];
        if (s != 0.f) atomicAdd(dst + (size_t)grp * 1024 + el, s);
    }
}

inline bool mma_shape_supported(int nb, int bi, int bo) {
    if (bi != 16 || bo != 16 || nb % 4 != 0) return false;
    int ng = nb / 4;
    return ng == 1 || ng == 2 || ng == 4 || ng == 8;
}

inline int launch_rel_mma_fwd(const RelArgs& A, const __nv_bfloat16* X, __nv_bfloat16* msg, int max_chunks,
                              cudaStream_t st) {
    const size_t smem = 3 * RGCN_CHUNK_EDGES * sizeof(int32_t) + (size_t)8 * kMmaStages * kTileBytes;
    static bool attr_set = false;
    if (!attr_set) {
        RGCN_CHECK_CUDA(cudaFuncSetAttribute(k_rel_mma_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    RGCN_LAUNCH(k_rel_mma_fwd, max_chunks, 256, smem, st, A, X, msg);
    return RGCN_OK;
}

}  // namespace rgcn

namespace rgcn {
// msg == nullptr: weight gradient only; gW == nullptr: feature-gradient messages only
inline int launch_rel_mma_bwd(const RelArgs& A, const __nv_bfloat16* X, const float* G, __nv_bfloat16* msg, float* gW,
                              int max_chunks, cudaStream_t st) {
    const size_t smem = 4 * RGCN_CHUNK_EDGES * sizeof(int32_t) + (size_t)8 * kBwdWarpBytes;
    static bool attr_set = false;
    if (!attr_set) {
        RGCN_CHECK_CUDA(cudaFuncSetAttribute(k_rel_mma_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    RGCN_LAUNCH(k_rel_mma_bwd, max_chunks, 256, smem, st, A, X, G, msg, gW);
    return RGCN_OK;
}
}  // namespace rgcn
