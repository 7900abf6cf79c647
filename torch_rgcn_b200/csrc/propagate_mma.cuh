// Tensor-core relation-batched kernels for bf16 features and 16x16 weight blocks (AM / SYN shapes).
//
// Work unit: one <=1024-edge chunk of one relation (or of one (tile, relation) group).  Each warp keeps the bf16
// fragments of "its" four 16x16 weight blocks in registers for the whole chunk and streams 16-edge tiles through a
// private 3-stage cp.async ring: the 16 gathered row slices (16 x 128 B) land in shared memory with a 16-byte XOR
// swizzle, ldmatrix feeds them to mma.sync.m16n8k16 (bf16 x bf16 -> fp32), the result is scaled by the per-edge
// weight, packed to bf16, bounced through the same tile buffer and written to the message rows with coalesced
// 16-byte stores.  After the index prologue the warps of a CTA are independent (no CTA-wide barriers).
//
// Two drivers use the chunk bodies:
//   k_rel_mma_fwd / k_rel_mma_bwd      one CTA per chunk, messages go to an nnz-sized buffer in HBM
//   k_tiled_mma_fwd / k_tiled_mma_bwd  persistent CTAs pull work from an in-order queue over row super-tiles:
//                                      messages of a tile live in a small ring that stays in L2, and the row sums of
//                                      tile k overlap the transform of tile k+1
//
// Each block is a true dense 16x16 GEMM tile shared by all edges of the relation, the one place the path is
// GEMM-shaped.  mma.sync (HMMA) rather than tcgen05: tiles are 16 edges x 16 x 16, far below the 128-row UMMA atom,
// and the kernels are gather-bound, not tensor-bound.
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

// One chunk of edges of relation p, described by global index arrays already offset to the chunk's first edge.
struct Chunk {
    int p, n;                 // relation, edges in the chunk (<= RGCN_CHUNK_EDGES)
    const int32_t* gather;    // rows of the bf16 matrix (X)
    const int32_t* other;     // rows of the fp32 matrix (G), backward only
    const int32_t* slot;      // message row ids
    const float* val;
    int slot_bias;            // subtracted from slot (first slot of the tile when messages go to a ring)
};

constexpr size_t kFwdSmemBytes = 3 * RGCN_CHUNK_EDGES * sizeof(int32_t) + (size_t)8 * kMmaStages * kTileBytes;

// msg[slot(e) - bias, bg*64 .. bg*64+63] = bf16( val_e * X[src_e, bg*64 ..] @ blockdiag(W_p[4bg .. 4bg+3]) )
__device__ __forceinline__ void mma_fwd_chunk(const Chunk& C, const float* __restrict__ W, int nb,
                                              const __nv_bfloat16* __restrict__ X, __nv_bfloat16* __restrict__ msg,
                                              unsigned char* smem) {
    const int n = C.n;
    int32_t* s_src = reinterpret_cast<int32_t*>(smem);
    int32_t* s_slot = s_src + RGCN_CHUNK_EDGES;
    float* s_val = reinterpret_cast<float*>(s_slot + RGCN_CHUNK_EDGES);
    unsigned char* rings = reinterpret_cast<unsigned char*>(s_val + RGCN_CHUNK_EDGES);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        s_src[i] = C.gather[i]; s_slot[i] = C.slot[i] - C.slot_bias; s_val[i] = C.val[i];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int NG = nb >> 2;                         // block groups of four 16x16 blocks (128 B of every row)
    const int bg = warp % NG, wsub = warp / NG, nsub = 8 / NG;
    const size_t row_bytes = (size_t)nb * 32;       // I == O == nb * 16 bf16

    // weight fragments (col-major B operand): b0 = W[2t..2t+1][n], b1 = W[2t+8..2t+9][n], n = 8h + g
    uint32_t bfrag[4][2][2];
    {
        const float* wp = W + ((size_t)C.p * nb + (size_t)bg * 4) * 256;
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
                __stcs(reinterpret_cast<uint4*>(Mb + (size_t)s_slot[le] * row_bytes + chunk * 16), v);   // streaming: read once, later
            }
        }
        __syncwarp();                                   // before a later issue() refills this stage
    }
    cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------------
// Fused backward chunk.  One relation-major pass reads X[gather] (bf16) and G[other] (fp32) once and produces BOTH
//   msg'[slot(e) - bias] = bf16( (val_e G[other_e]) @ blockdiag(W_p)^T )   -> summed per source row afterwards
//   gW_p                += X[gather]^T (val_e G[other_e])                   -> fp32 fragments kept in registers for the
//                                                                              whole chunk, one atomic flush
// G rows are staged as fp32, scaled by val and rounded to bf16 in shared memory (fp32 accumulation in the MMA).
// ------------------------------------------------------------------------------------------------------
constexpr int kBwdStages = 3;
constexpr int kGTileBytes = 16 * 256;              // 16 edges x 64 fp32 outputs
// fp32 G: per stage an X tile + an fp32 G tile, plus one bf16 G tile per warp; bf16 G: X tile + bf16 G tile
template <bool GBF16> struct BwdSmem {
    static constexpr int kStageBytes = GBF16 ? 2 * kTileBytes : kTileBytes + kGTileBytes;
    static constexpr int kWarpBytes = kBwdStages * kStageBytes + (GBF16 ? 0 : kTileBytes);
    static constexpr size_t kBytes = 4 * RGCN_CHUNK_EDGES * sizeof(int32_t) + (size_t)8 * kWarpBytes;
};

// GBF16 = true: G was pre-rounded to bf16 by k_cast_colsum, so its rows are gathered straight into a swizzled
// tile (128 B per edge instead of 256 B); val is applied to the message in fp32 and to the X tile in place.
template <bool GBF16>
__device__ __forceinline__ void mma_bwd_chunk(const Chunk& C, const float* __restrict__ W, int nb,
                                              const __nv_bfloat16* __restrict__ X, const void* __restrict__ Gv,
                                              __nv_bfloat16* __restrict__ msg, float* __restrict__ gW,
                                              unsigned char* smem) {
    using L = BwdSmem<GBF16>;
    const int n = C.n;
    int32_t* s_src = reinterpret_cast<int32_t*>(smem);
    int32_t* s_dst = s_src + RGCN_CHUNK_EDGES;
    int32_t* s_slot = s_dst + RGCN_CHUNK_EDGES;
    float* s_val = reinterpret_cast<float*>(s_slot + RGCN_CHUNK_EDGES);
    unsigned char* rings = reinterpret_cast<unsigned char*>(s_val + RGCN_CHUNK_EDGES);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        s_src[i] = C.gather[i]; s_dst[i] = C.other[i]; s_val[i] = C.val[i];
        if (msg) s_slot[i] = C.slot[i] - C.slot_bias;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int NG = nb >> 2;
    const int bg = warp % NG, wsub = warp / NG, nsub = 8 / NG;
    const size_t xrow_bytes = (size_t)nb * 32;                      // bf16 rows of X, msg' (and bf16 G)
    const size_t grow_bytes = (size_t)nb * (GBF16 ? 32 : 64);

    // W^T fragments for msg' = Gb @ W^T:  B[k = j][n = i] = W[i][j];  b0 = W[n][2t..2t+1], b1 = W[n][2t+8..2t+9]
    uint32_t wt[4][2][2];
    if (msg) {
        const float* wp = W + ((size_t)C.p * nb + (size_t)bg * 4) * 256;
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

    unsigned char* ring = rings + (size_t)warp * L::kWarpBytes;
    unsigned char* gb_shared = ring + kBwdStages * L::kStageBytes;  // fp32-G variant: single bf16 (val * G) tile
    const int ntiles = (n + 15) >> 4;
    const unsigned char* Xb = reinterpret_cast<const unsigned char*>(X) + (size_t)bg * 128;
    const unsigned char* Gp = reinterpret_cast<const unsigned char*>(Gv) + (size_t)bg * (GBF16 ? 128 : 256);
    unsigned char* Mb = reinterpret_cast<unsigned char*>(msg) + (size_t)bg * 128;

    auto issue = [&](int k) {
        const int tile = wsub + k * nsub;
        if (tile < ntiles) {
            unsigned char* st = ring + (k % kBwdStages) * L::kStageBytes;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int row = (lane >> 3) + 4 * it, chunk = lane & 7, le = tile * 16 + row;
                const bool ok = le < n;
                if (gW) cp_async16(st + tile_off(row, chunk), Xb + (size_t)s_src[ok ? le : 0] * xrow_bytes + chunk * 16, ok);
                if constexpr (GBF16)
                    cp_async16(st + kTileBytes + tile_off(row, chunk),
                               Gp + (size_t)s_dst[ok ? le : 0] * grow_bytes + chunk * 16, ok);
            }
            if constexpr (!GBF16) {
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int row = (lane >> 4) + 2 * it, c16 = lane & 15, le = tile * 16 + row;
                    const bool ok = le < n;
                    cp_async16(st + kTileBytes + row * 256 + c16 * 16, Gp + (size_t)s_dst[ok ? le : 0] * grow_bytes + c16 * 16, ok);
                }
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
        unsigned char* st = ring + (k % kBwdStages) * L::kStageBytes;
        const int tile = wsub + k * nsub;
        unsigned char* gb;
        if constexpr (GBF16) {
            gb = st + kTileBytes;                       // already bf16 and swizzled
            if (gW) {                                   // X'[e] = val_e * X[e] in place (operand of the weight gradient)
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int row = (lane >> 3) + 4 * it, chunk = lane & 7, le = tile * 16 + row;
                    const float v = le < n ? s_val[le] : 0.f;
                    uint4* px = reinterpret_cast<uint4*>(st + tile_off(row, chunk));
                    uint4 x = *px;
                    float f[8];
                    unpack_bf16x2(x.x, f[0], f[1]); unpack_bf16x2(x.y, f[2], f[3]);
                    unpack_bf16x2(x.z, f[4], f[5]); unpack_bf16x2(x.w, f[6], f[7]);
                    *px = make_uint4(pack_bf16x2(f[0] * v, f[1] * v), pack_bf16x2(f[2] * v, f[3] * v),
                                     pack_bf16x2(f[4] * v, f[5] * v), pack_bf16x2(f[6] * v, f[7] * v));
                }
            }
        } else {
            gb = gb_shared;                             // fp32 G rows -> val-scaled bf16 tile
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int row = (lane >> 4) + 2 * it, c4 = lane & 15, le = tile * 16 + row;
                const float v = le < n ? s_val[le] : 0.f;
                const float4 f = *reinterpret_cast<const float4*>(st + kTileBytes + row * 256 + c4 * 16);
                *reinterpret_cast<uint2*>(gb + tile_off(row, c4 >> 1) + (c4 & 1) * 8) =
                    make_uint2(pack_bf16x2(f.x * v, f.y * v), pack_bf16x2(f.z * v, f.w * v));
            }
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
            float v0 = 1.f, v1 = 1.f;                   // fp32-G variant folded val into the G tile already
            if constexpr (GBF16) {
                const int le0 = tile * 16 + g, le1 = le0 + 8;
                v0 = le0 < n ? s_val[le0] : 0.f; v1 = le1 < n ? s_val[le1] : 0.f;
            }
#pragma unroll
            for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int chunk = kb * 2 + h;
                    *reinterpret_cast<uint32_t*>(st + tile_off(g, chunk) + t * 4) =
                        pack_bf16x2(macc[kb][h][0] * v0, macc[kb][h][1] * v0);
                    *reinterpret_cast<uint32_t*>(st + tile_off(g + 8, chunk) + t * 4) =
                        pack_bf16x2(macc[kb][h][2] * v1, macc[kb][h][3] * v1);
                }
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int row = (lane >> 3) + 4 * it, chunk = lane & 7, le = tile * 16 + row;
                if (le < n) {
                    const uint4 v = *reinterpret_cast<const uint4*>(st + tile_off(row, chunk));
                    __stcs(reinterpret_cast<uint4*>(Mb + (size_t)s_slot[le] * xrow_bytes + chunk * 16), v);
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
    float* dst = gW + (size_t)C.p * nb * 256;
    for (int i = threadIdx.x; i < NG * 1024; i += blockDim.x) {
        const int grp = i >> 10, el = i & 1023;
        float s = 0.f;
        for (int w = 0; w < nsub; ++w) s += red[(size_t)(grp + w * NG) * 1024 + el];
        if (s != 0.f) atomicAdd(dst + (size_t)grp * 1024 + el, s);
    }
}

// G (N, O) fp32 -> bf16 copy, fused with the bias gradient (column sums): one streaming pass over G.
// Thread = 8 consecutive columns of a strided set of rows; block-level smem reduction, then atomics.
__global__ void __launch_bounds__(256) k_cast_colsum(const float* __restrict__ G, long long N, int O,
                                                     long long rows_per_block, __nv_bfloat16* __restrict__ Gb,
                                                     float* __restrict__ gbias) {
    extern __shared__ float sums[];
    const int cg = O >> 3;                               // 8-column groups per row
    for (int j = threadIdx.x; j < O; j += blockDim.x) sums[j] = 0.f;
    __syncthreads();
    const int lanes = blockDim.x / cg;                   // rows handled concurrently
    const int q = threadIdx.x % cg, rl = threadIdx.x / cg;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (rl < lanes) {
        const long long r0 = blockIdx.x * rows_per_block, r1 = min(N, r0 + rows_per_block);
        for (long long r = r0 + rl; r < r1; r += lanes) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(G + (size_t)r * O + 8 * q));
            const float4 b = __ldg(reinterpret_cast<const float4*>(G + (size_t)r * O + 8 * q) + 1);
            *reinterpret_cast<uint4*>(Gb + (size_t)r * O + 8 * q) =
                make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
            acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
            acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
        }
        if (gbias) {
#pragma unroll
            for (int j = 0; j < 8; ++j) atomicAdd(&sums[8 * q + j], acc[j]);
        }
    }
    if (!gbias) return;
    __syncthreads();
    for (int j = threadIdx.x; j < O; j += blockDim.x) atomicAdd(gbias + j, sums[j]);
}

// ---- one CTA per chunk --------------------------------------------------------------------------------
__device__ __forceinline__ Chunk rel_chunk(const RelArgs& A, int c) {
    int p, e0, e1;
    chunk_lookup(A, c, p, e0, e1);
    Chunk C;
    C.p = p; C.n = e1 - e0;
    C.gather = A.gather + e0; C.other = A.other ? A.other + e0 : nullptr;
    C.slot = A.slot ? A.slot + e0 : nullptr; C.val = A.val + e0; C.slot_bias = 0;
    return C;
}

__global__ void __launch_bounds__(256) k_rel_mma_fwd(RelArgs A, const __nv_bfloat16* __restrict__ X,
                                                     __nv_bfloat16* __restrict__ msg) {
    extern __shared__ __align__(128) unsigned char smem_mma_fwd[];
    if ((int)blockIdx.x >= A.chunkptr[A.num_rels]) return;
    mma_fwd_chunk(rel_chunk(A, blockIdx.x), A.W, A.nb, X, msg, smem_mma_fwd);
}

template <bool GBF16>
__global__ void __launch_bounds__(256, GBF16 ? 2 : 1) k_rel_mma_bwd(RelArgs A, const __nv_bfloat16* __restrict__ X,
                                                                    const void* __restrict__ G,
                                                                    __nv_bfloat16* __restrict__ msg,
                                                                    float* __restrict__ gW) {
    extern __shared__ __align__(128) unsigned char smem_mma_bwd[];
    if ((int)blockIdx.x >= A.chunkptr[A.num_rels]) return;
    mma_bwd_chunk<GBF16>(rel_chunk(A, blockIdx.x), A.W, A.nb, X, G, msg, gW, smem_mma_bwd);
}

// ------------------------------------------------------------------------------------------------------
// Persistent tiled driver: in-order work queue over row super-tiles, messages in a small ring that stays in L2.
//   step j of the queue = [transform spans of tile j] then [row-sum blocks of tile j - lag]
//   a row-sum block of tile k waits until all spans of tile k are done (done1[k]);
//   a span of tile k waits until every earlier tile that used ring slot k % depth has been summed: per-slot
//   counter slot_done[k % depth] >= slotneed[k] (a per-tile "tile k-depth is done" test is NOT enough: empty
//   tiles in between break the chain).
// Items are claimed in order, so a waiting CTA only ever waits for items held by running CTAs: no deadlock.
//
// A span is up to 1024 consecutive edges of a tile in (relation, row) order and may cross relation boundaries:
// each warp walks a contiguous run of 16-edge MMA tiles and reloads its (pre-packed, bf16) weight fragments when
// the relation changes; a 16-edge tile that straddles relations is multiplied once per relation run and every
// row keeps the result of its own relation.  The same kernel serves the forward (X, W) and the feature-gradient
// messages (bf16 copy of grad_out, W^T).
// ------------------------------------------------------------------------------------------------------
struct TiledArgs {
    rgcn_tiling tl;
    const int32_t* rowptr;     // CSR of the tile side (d_rowptr forward, s_rowptr backward)
    int T, nb, depth;
    long long capacity;        // message rows per ring slot
    int32_t* queue;            // [0]: next item
    int32_t* slot_done;        // [depth] finished row blocks per ring slot
    int32_t* done1;            // [T] finished spans per tile
    int32_t* status;           // [3] set if a wait gave up (watchdog)
    const uint4* wfrag;        // packed bf16 weight fragments, see k_pack_wfrag
    const float* bias;
};

// frag[((p * NG + bg) * 32 + lane) * 16 + kb * 4 + h * 2 + r]: the two B-operand registers (r) of block 4bg+kb,
// output half h, for lane (g = lane / 4, t = lane % 4).  transpose = 0: B[k][n] = W[k][n] (forward);
// transpose = 1: B[k][n] = W[n][k] (messages of the feature gradient).
__global__ void k_pack_wfrag(const float* __restrict__ W, int Rp, int nb, int transpose, uint32_t* __restrict__ frag) {
    const int NG = nb >> 2;
    const long long total = (long long)Rp * NG * 32 * 16;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int reg = (int)(i & 15), lane = (int)((i >> 4) & 31);
    const long long pg = i >> 9;                   // p * NG + bg
    const int bg = (int)(pg % NG);
    const long long p = pg / NG;
    const int kb = reg >> 2, h = (reg >> 1) & 1, r = reg & 1;
    const int g = lane >> 2, t = lane & 3;
    const float* wb = W + ((size_t)p * nb + (size_t)bg * 4 + kb) * 256;
    const int n = h * 8 + g, k0 = 2 * t + 8 * r;
    const float lo = transpose ? wb[n * 16 + k0] : wb[k0 * 16 + n];
    const float hi = transpose ? wb[n * 16 + k0 + 1] : wb[(k0 + 1) * 16 + n];
    frag[i] = pack_bf16x2(lo, hi);
}

__device__ __forceinline__ int ld_acquire(const int32_t* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void wait_count(const int32_t* counter, int need, int32_t* status) {
    if (threadIdx.x == 0) {
        unsigned spins = 0;
        while (ld_acquire(counter) < need) {
            __nanosleep(64);
            // ~1 s without progress (a pre-empted or time-sliced device, a profiler replay): record it and abort the
            // launch.  Running on with an incomplete message ring would return silently wrong sums; the trap turns
            // it into a CUDA error the caller sees at its next synchronisation.
            if (++spins > (1u << 24)) { atomicExch(status + 3, 1); __threadfence_system(); __trap(); }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void signal_done(int32_t* counter) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1);
    }
}

constexpr size_t kSpanSmemBytes = 4 * RGCN_SPAN_EDGES * sizeof(int32_t) + (size_t)8 * kMmaStages * kTileBytes;

struct Span {
    int n;                     // edges (<= RGCN_SPAN_EDGES)
    const int32_t* gather;     // rows of the bf16 source matrix
    const int32_t* rel;
    const int32_t* slot;
    const float* val;
    int slot_bias;
};

__device__ __forceinline__ void mma_span(const Span& S, const uint4* __restrict__ wfrag, int nb,
                                         const __nv_bfloat16* __restrict__ X, __nv_bfloat16* __restrict__ msg,
                                         unsigned char* smem) {
    const int n = S.n;
    int32_t* s_src = reinterpret_cast<int32_t*>(smem);
    int32_t* s_slot = s_src + RGCN_SPAN_EDGES;
    int32_t* s_rel = s_slot + RGCN_SPAN_EDGES;
    float* s_val = reinterpret_cast<float*>(s_rel + RGCN_SPAN_EDGES);
    unsigned char* rings = reinterpret_cast<unsigned char*>(s_val + RGCN_SPAN_EDGES);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        s_src[i] = S.gather[i]; s_slot[i] = S.slot[i] - S.slot_bias; s_rel[i] = S.rel[i]; s_val[i] = S.val[i];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int NG = nb >> 2;
    const int bg = warp % NG, wsub = warp / NG, nsub = 8 / NG;
    const size_t row_bytes = (size_t)nb * 32;
    unsigned char* ring = rings + (size_t)warp * kMmaStages * kTileBytes;
    const int ntiles = (n + 15) >> 4;
    const int per = (ntiles + nsub - 1) / nsub;      // this warp: contiguous MMA tiles [t0, t1)
    const int t0 = wsub * per, t1 = min(ntiles, t0 + per);
    const unsigned char* Xb = reinterpret_cast<const unsigned char*>(X) + (size_t)bg * 128;
    unsigned char* Mb = reinterpret_cast<unsigned char*>(msg) + (size_t)bg * 128;

    uint32_t bf[16];
    int cur_rel = -1;
    auto load_frags = [&](int p) {
        if (p == cur_rel) return;
        const uint4* f = wfrag + ((size_t)p * NG + bg) * 32 * 4 + lane * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint4 v = __ldg(f + q);
            bf[4 * q] = v.x; bf[4 * q + 1] = v.y; bf[4 * q + 2] = v.z; bf[4 * q + 3] = v.w;
        }
        cur_rel = p;
    };
    auto issue = [&](int k) {
        const int tile = t0 + k;
        if (tile < t1) {
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
    for (int k = 0; t0 + k < t1; ++k) {
        issue(k + 2);
        cp_async_wait<2>();
        __syncwarp();
        unsigned char* st = ring + (k % kMmaStages) * kTileBytes;
        const int tile = t0 + k;
        const int rows = min(16, n - tile * 16);
        uint32_t a[4][4];
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
            const int row = (lane & 7) + ((lane >> 3) & 1) * 8, chunk = kb * 2 + (lane >> 4);
            ldmatrix_x4(a[kb], smem_u32(st + tile_off(row, chunk)));
        }
        float acc[4][2][4];
#pragma unroll
        for (int kb = 0; kb < 4; ++kb)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[kb][h][q] = 0.f;
        const int r_first = s_rel[tile * 16], r_last = s_rel[tile * 16 + rows - 1];
        if (r_first == r_last) {
            load_frags(r_first);
#pragma unroll
            for (int kb = 0; kb < 4; ++kb) {
                mma_bf16_16816(acc[kb][0], a[kb], bf[kb * 4], bf[kb * 4 + 1]);
                mma_bf16_16816(acc[kb][1], a[kb], bf[kb * 4 + 2], bf[kb * 4 + 3]);
            }
        } else {                                        // tile straddles relations: one pass per relation run
            int row0 = 0;
            while (row0 < rows) {
                const int p = s_rel[tile * 16 + row0];
                int row1 = row0 + 1;
                while (row1 < rows && s_rel[tile * 16 + row1] == p) ++row1;
                load_frags(p);
                const bool top = g >= row0 && g < row1, bot = g + 8 >= row0 && g + 8 < row1;
#pragma unroll
                for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float c[4] = {0.f, 0.f, 0.f, 0.f};
                        mma_bf16_16816(c, a[kb], bf[kb * 4 + 2 * h], bf[kb * 4 + 2 * h + 1]);
                        if (top) { acc[kb][h][0] = c[0]; acc[kb][h][1] = c[1]; }
                        if (bot) { acc[kb][h][2] = c[2]; acc[kb][h][3] = c[3]; }
                    }
                row0 = row1;
            }
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
                __stcs(reinterpret_cast<uint4*>(Mb + (size_t)s_slot[le] * row_bytes + chunk * 16), v);   // streaming: read once, later
            }
        }
        __syncwarp();
    }
    cp_async_wait<0>();
}

// L2-only 16-byte load (the ring is rewritten inside the kernel, so L1 must not serve it)
__device__ __forceinline__ void ldcg8(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 v = __ldcg(reinterpret_cast<const uint4*>(p));
    unpack_bf16x2(v.x, f[0], f[1]); unpack_bf16x2(v.y, f[2], f[3]);
    unpack_bf16x2(v.z, f[4], f[5]); unpack_bf16x2(v.w, f[6], f[7]);
}

// rows [r0, r1): out[row, :] = bias + sum of the row's messages (contiguous in the ring slot).
// Thread = 8 columns of one row; 4 independent 16-byte loads in flight per thread.
__device__ __forceinline__ void row_sum_block(const int32_t* __restrict__ rowptr, int r0, int r1, int width,
                                              const __nv_bfloat16* __restrict__ ring_slot, int slot_bias,
                                              const float* __restrict__ bias, float* __restrict__ out,
                                              float* scratch) {
    const int cg = width >> 3;
    const int total = (r1 - r0) * cg;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int row = r0 + idx / cg, q = idx % cg;
        const int e0 = rowptr[row] - slot_bias, e1 = rowptr[row + 1] - slot_bias;
        if (e1 - e0 > RGCN_LONG_ROW) continue;                  // hub rows: cooperative pass below
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const __nv_bfloat16* m = ring_slot + (size_t)e0 * width + 8 * q;
        int e = e0;
        for (; e + 3 < e1; e += 4, m += 4 * (size_t)width) {
            float a[8], b[8], c[8], d[8];
            ldcg8(m, a); ldcg8(m + width, b); ldcg8(m + 2 * (size_t)width, c); ldcg8(m + 3 * (size_t)width, d);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += (a[j] + b[j]) + (c[j] + d[j]);
        }
        for (; e < e1; ++e, m += width) {
            float a[8];
            ldcg8(m, a);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += a[j];
        }
        if (bias) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias) + 2 * q);
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias) + 2 * q + 1);
            acc[0] += b0.x; acc[1] += b0.y; acc[2] += b0.z; acc[3] += b0.w;
            acc[4] += b1.x; acc[5] += b1.y; acc[6] += b1.z; acc[7] += b1.w;
        }
        float4* o = reinterpret_cast<float4*>(out + (size_t)row * width + 8 * q);
        o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    // hub rows: the whole CTA sums one row (edge-parallel), partials reduced through `scratch` (256 x 8 floats)
    const int lanes = 256 / cg, q = threadIdx.x % cg, lane = threadIdx.x / cg;      // cg <= 64 (width <= 512)
    for (int row = r0; row < r1; ++row) {
        const int e0 = rowptr[row] - slot_bias, e1 = rowptr[row + 1] - slot_bias;
        if (e1 - e0 <= RGCN_LONG_ROW) continue;                 // uniform across the CTA
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (lane < lanes)
            for (int e = e0 + lane; e < e1; e += lanes) {
                float a[8];
                ldcg8(ring_slot + (size_t)e * width + 8 * q, a);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += a[j];
            }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j) scratch[threadIdx.x * 8 + j] = acc[j];
        __syncthreads();
        if (lane == 0) {
            for (int l = 1; l < lanes; ++l)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += scratch[(threadIdx.x + l * cg) * 8 + j];
            if (bias)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += bias[8 * q + j];
            float4* o = reinterpret_cast<float4*>(out + (size_t)row * width + 8 * q);
            o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
            o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
        }
    }
}

// claim the next queue item (all threads get the same record); false when the queue is exhausted
__device__ __forceinline__ bool next_item(const TiledArgs& A, int* s_item, rgcn_tile_item& it) {
    if (threadIdx.x == 0) *s_item = atomicAdd(A.queue, 1);
    __syncthreads();
    const int item = *s_item;
    __syncthreads();
    if (item >= __ldg(A.tl.stepptr + A.T + A.depth / 2)) return false;
    const int4* rec = reinterpret_cast<const int4*>(A.tl.items) + 2 * (size_t)item;
    const int4 lo = __ldg(rec), hi = __ldg(rec + 1);
    it.kind = lo.x; it.tile = lo.y; it.a = lo.z; it.b = lo.w;
    it.c = hi.x; it.slot_bias = hi.y; it.need = hi.z; it.pad = 0;
    return true;
}

// rows = tile side; gathers src[tl.col]; writes out[row] = bias + sum of messages
__global__ void __launch_bounds__(256, 2) k_tiled_span(TiledArgs A, const __nv_bfloat16* __restrict__ src,
                                                       __nv_bfloat16* __restrict__ ring, float* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem_tiled[];
    __shared__ int s_item;
    const size_t width = (size_t)A.nb * 16;
    rgcn_tile_item it;
    while (next_item(A, &s_item, it)) {
        const int k = it.tile;
        __nv_bfloat16* slot_base = ring + (size_t)(k % A.depth) * A.capacity * width;
        if (it.kind == 0) {
            Span S;
            S.n = it.c; S.gather = A.tl.col + it.b; S.rel = A.tl.rel + it.b; S.slot = A.tl.slot + it.b;
            S.val = A.tl.val + it.b; S.slot_bias = it.slot_bias;
            wait_count(A.slot_done + (k % A.depth), it.need, A.status);
            mma_span(S, A.wfrag, A.nb, src, slot_base, smem_tiled);
            signal_done(A.done1 + k);
        } else {
            wait_count(A.done1 + k, it.need, A.status);
            row_sum_block(A.rowptr, it.a, it.b, (int)width, slot_base, it.slot_bias, A.bias, out,
                          reinterpret_cast<float*>(smem_tiled));
            signal_done(A.slot_done + (k % A.depth));
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
inline bool mma_shape_supported(int nb, int bi, int bo) {
    if (bi != 16 || bo != 16 || nb % 4 != 0) return false;
    int ng = nb / 4;
    return ng == 1 || ng == 2 || ng == 4 || ng == 8;
}

inline int launch_rel_mma_fwd(const RelArgs& A, const __nv_bfloat16* X, __nv_bfloat16* msg, int max_chunks,
                              cudaStream_t st) {
    // per device / context, so set on every launch (cheap) rather than once per process
    RGCN_CHECK_CUDA(cudaFuncSetAttribute(k_rel_mma_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmemBytes));
    RGCN_LAUNCH(k_rel_mma_fwd, max_chunks, 256, kFwdSmemBytes, st, A, X, msg);
    return RGCN_OK;
}

// msg == nullptr: weight gradient only; gW == nullptr: feature-gradient messages only.  G is the bf16 copy.
inline int launch_rel_mma_bwd(const RelArgs& A, const __nv_bfloat16* X, const __nv_bfloat16* Gb, __nv_bfloat16* msg,
                              float* gW, int max_chunks, cudaStream_t st) {
    constexpr size_t smem = BwdSmem<true>::kBytes;
    RGCN_CHECK_CUDA(cudaFuncSetAttribute(k_rel_mma_bwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RGCN_LAUNCH(k_rel_mma_bwd<true>, max_chunks, 256, smem, st, A, X, static_cast<const void*>(Gb), msg, gW);
    return RGCN_OK;
}

inline int launch_cast_colsum(const float* G, int64_t N, int O, __nv_bfloat16* Gb, float* gbias, cudaStream_t st) {
    int64_t rows_per_block = (N + kNumSMs * 8 - 1) / (kNumSMs * 8);
    if (rows_per_block < 64) rows_per_block = 64;
    const int grid = (int)((N + rows_per_block - 1) / rows_per_block);
    RGCN_LAUNCH(k_cast_colsum, grid, 256, (size_t)O * sizeof(float), st, G, (long long)N, O, (long long)rows_per_block, Gb, gbias);
    return RGCN_OK;
}


inline size_t tiled_counter_bytes(int64_t T) { return align_up((size_t)(T + 4 + RGCN_MAX_RING_DEPTH) * sizeof(int32_t)); }
inline size_t wfrag_bytes(int64_t Rp, int nb) { return align_up((size_t)Rp * (nb / 4) * 32 * 16 * sizeof(uint32_t)); }

inline int launch_pack_wfrag(const float* W, int64_t Rp, int nb, bool transpose, uint32_t* frag, cudaStream_t st) {
    const long long total = (long long)Rp * (nb / 4) * 32 * 16;
    RGCN_LAUNCH(k_pack_wfrag, grid_for(total, 256), 256, 0, st, W, (int)Rp, nb, transpose ? 1 : 0, frag);
    return RGCN_OK;
}

inline TiledArgs make_tiled_args(const rgcn_graph* g, bool backward, int nb, const uint32_t* wfrag, const float* bias,
                                 int32_t* counters) {
    TiledArgs A{};
    A.tl = backward ? g->bt : g->ft;
    A.rowptr = backward ? g->s_rowptr : g->d_rowptr;
    A.T = (int)g->num_tiles; A.nb = nb; A.depth = (int)g->ring_depth;
    A.capacity = (long long)g->tile_capacity;
    A.queue = counters; A.slot_done = counters + 4; A.done1 = counters + 4 + RGCN_MAX_RING_DEPTH;
    A.status = g->status; A.wfrag = reinterpret_cast<const uint4*>(wfrag); A.bias = bias;
    return A;
}

inline int launch_tiled_span(const TiledArgs& A, const __nv_bfloat16* src, __nv_bfloat16* ring, float* out,
                             cudaStream_t st) {
    RGCN_CHECK_CUDA(cudaFuncSetAttribute(k_tiled_span, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSpanSmemBytes));
    int per_sm = 0, dev = 0, sms = kNumSMs;
    RGCN_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tiled_span, 256, kSpanSmemBytes));
    RGCN_CHECK_CUDA(cudaGetDevice(&dev));
    RGCN_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = sms * (per_sm > 0 ? per_sm : 1);
    RGCN_LAUNCH(k_tiled_span, grid, 256, kSpanSmemBytes, st, A, src, ring, out);
    return RGCN_OK;
}

}  // namespace rgcn
