// Relation-batched propagation kernels for aligned block-diagonal / small dense weights.
//
// Why relation-major: with one (s, p) segment per edge (typical of sparse multi-relational graphs) a
// destination-major walk has to fetch a different W_p for every edge, and that weight traffic (L1/L2),
// not the feature gather, bounds the kernel.  Here edges are walked in relation order, so a CTA stages
// W_p once per 1024-edge chunk and every weight element is reused across the whole chunk:
//
//   k_rel_transform : msg[slot(e)] = val_e * X[src_e] @ W_p        (gather rows, per-block FMA, scatter rows)
//   k_row_sum       : out[row] = bias + sum of the row's messages  (slot order == row-major order, so this is
//                                                                   a contiguous streaming segmented sum)
//   k_row_sum_long  : the same for hub rows (> RGCN_LONG_ROW messages): one CTA per row, edge-parallel
//   k_rel_wgrad     : gW_p += sum_e val_e X[src_e]^T G[dst_e]      (cp.async-staged rows, 8x8 register tiles,
//                                                                   compile-time shapes, shuffle split-K reduction)
//
// The same two kernels serve the feature gradient: walk with (gather = dst, slot = source-major position),
// transposed weights and the upstream gradient as the feature matrix.  These are the exact-fp32 kernels; bf16
// features with 16x16 blocks use the tensor-core kernels of propagate_mma.cuh.
//
// Return convention of the launch_* dispatchers: 0 = launched, < 0 = error, > 0 = shape not instantiated here.
#pragma once
#include "common.cuh"
#include "propagate_generic.cuh"

namespace rgcn {

struct RelArgs {
    const int32_t* relptr; const int32_t* chunkptr; int num_rels;
    const int32_t* gather;    // per edge: row of the feature matrix to read
    const int32_t* other;     // per edge: the other endpoint (weight gradient only)
    const int32_t* slot;      // per edge: message row to write
    const float* val;
    const float* W;           // per relation: nb contiguous (BI, BO) blocks
    int nb;
};

__device__ __forceinline__ void chunk_lookup(const RelArgs& A, int c, int& p, int& e0, int& e1) {
    int lo = 0, hi = A.num_rels;              // chunkptr[lo] <= c < chunkptr[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (A.chunkptr[mid] <= c) lo = mid; else hi = mid;
    }
    p = lo;
    e0 = A.relptr[p] + (c - A.chunkptr[p]) * RGCN_CHUNK_EDGES;
    e1 = min(A.relptr[p + 1], e0 + RGCN_CHUNK_EDGES);
}

// ---- vector row-slice loads / stores -----------------------------------------------------------------
template <int n>
__device__ __forceinline__ void load_slice(const float* __restrict__ p, float (&x)[n], float scale) {
    static_assert(n % 4 == 0, "slice must be a multiple of 4 floats");
#pragma unroll
    for (int k = 0; k < n / 4; ++k) {
        float4 v = __ldg(reinterpret_cast<const float4*>(p) + k);
        x[4 * k] = v.x * scale; x[4 * k + 1] = v.y * scale; x[4 * k + 2] = v.z * scale; x[4 * k + 3] = v.w * scale;
    }
}
__device__ __forceinline__ void unpack_bf16x2(uint32_t w, float& lo, float& hi) {
    lo = __uint_as_float(w << 16);
    hi = __uint_as_float(w & 0xffff0000u);
}
template <int n>
__device__ __forceinline__ void load_slice(const __nv_bfloat16* __restrict__ p, float (&x)[n], float scale) {
    static_assert(n % 4 == 0, "slice must be a multiple of 4 elements");
    if constexpr (n % 8 == 0) {
#pragma unroll
        for (int k = 0; k < n / 8; ++k) {
            uint4 v = __ldg(reinterpret_cast<const uint4*>(p) + k);
            unpack_bf16x2(v.x, x[8 * k], x[8 * k + 1]); unpack_bf16x2(v.y, x[8 * k + 2], x[8 * k + 3]);
            unpack_bf16x2(v.z, x[8 * k + 4], x[8 * k + 5]); unpack_bf16x2(v.w, x[8 * k + 6], x[8 * k + 7]);
        }
    } else {
#pragma unroll
        for (int k = 0; k < n / 4; ++k) {
            uint2 v = __ldg(reinterpret_cast<const uint2*>(p) + k);
            unpack_bf16x2(v.x, x[4 * k], x[4 * k + 1]); unpack_bf16x2(v.y, x[4 * k + 2], x[4 * k + 3]);
        }
    }
#pragma unroll
    for (int k = 0; k < n; ++k) x[k] *= scale;
}
template <int n>
__device__ __forceinline__ void store_slice(float* __restrict__ p, const float (&y)[n]) {
#pragma unroll
    for (int k = 0; k < n / 4; ++k)
        reinterpret_cast<float4*>(p)[k] = make_float4(y[4 * k], y[4 * k + 1], y[4 * k + 2], y[4 * k + 3]);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
template <int n>
__device__ __forceinline__ void store_slice(__nv_bfloat16* __restrict__ p, const float (&y)[n]) {
#pragma unroll
    for (int k = 0; k < n / 4; ++k)
        reinterpret_cast<uint2*>(p)[k] = make_uint2(pack_bf16x2(y[4 * k], y[4 * k + 1]), pack_bf16x2(y[4 * k + 2], y[4 * k + 3]));
}

// ------------------------------------------------------------------------------------------------------
// phase 1: one thread per (edge, block); W_p broadcast from shared memory
// ------------------------------------------------------------------------------------------------------
template <typename XT, typename MT, int BI, int BO, int EPT>
__global__ void __launch_bounds__(256) k_rel_transform(RelArgs A, const XT* __restrict__ X, MT* __restrict__ msg) {
    extern __shared__ float Ws[];
    const int c = blockIdx.x;
    if (c >= A.chunkptr[A.num_rels]) return;
    int p, e0, e1;
    chunk_lookup(A, c, p, e0, e1);
    constexpr int WB = BI * BO + 4;            // padded block stride: blocks start 4 banks apart
    const int nb = A.nb;
    const float* wp = A.W + (size_t)p * nb * BI * BO;
    for (int i = threadIdx.x; i < nb * BI * BO; i += blockDim.x) {
        int b = i / (BI * BO);
        Ws[b * WB + (i - b * BI * BO)] = wp[i];
    }
    __syncthreads();
    const int groups = blockDim.x / nb;
    const int blk = threadIdx.x % nb, grp = threadIdx.x / nb;
    if (grp >= groups) return;
    const size_t I = (size_t)nb * BI, O = (size_t)nb * BO;
    const float* w = Ws + blk * WB;
    for (int base = e0; base < e1; base += groups * EPT) {
        float x[EPT][BI], y[EPT][BO];
        int slot[EPT];
#pragma unroll
        for (int u = 0; u < EPT; ++u) {
            const int e = base + u * groups + grp;
            slot[u] = -1;
            if (e < e1) {
                slot[u] = A.slot[e];
                load_slice<BI>(X + (size_t)A.gather[e] * I + blk * BI, x[u], A.val[e]);
            } else {
#pragma unroll
                for (int i = 0; i < BI; ++i) x[u][i] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < BO; ++j) y[u][j] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < BI; ++i) {
#pragma unroll
            for (int j4 = 0; j4 < BO / 4; ++j4) {
                const float4 w4 = *reinterpret_cast<const float4*>(w + i * BO + 4 * j4);
#pragma unroll
                for (int u = 0; u < EPT; ++u) {
                    y[u][4 * j4] += x[u][i] * w4.x; y[u][4 * j4 + 1] += x[u][i] * w4.y;
                    y[u][4 * j4 + 2] += x[u][i] * w4.z; y[u][4 * j4 + 3] += x[u][i] * w4.w;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < EPT; ++u)
            if (slot[u] >= 0) store_slice<BO>(msg + (size_t)slot[u] * O + blk * BO, y[u]);
    }
}

// ------------------------------------------------------------------------------------------------------
// phase 2: out[row, V*q .. V*q+V-1] = bias + sum_{e in row} msg[e, ...]   (thread per (row, V columns), V = 4 or 8)
// ------------------------------------------------------------------------------------------------------
template <int V> struct VecF { float v[V]; };

template <int V>
__device__ __forceinline__ VecF<V> loadv(const float* p) {
    VecF<V> r;
#pragma unroll
    for (int k = 0; k < V / 4; ++k) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p) + k);
        r.v[4 * k] = a.x; r.v[4 * k + 1] = a.y; r.v[4 * k + 2] = a.z; r.v[4 * k + 3] = a.w;
    }
    return r;
}
template <int V>
__device__ __forceinline__ VecF<V> loadv(const __nv_bfloat16* p) {
    VecF<V> r;
    if constexpr (V == 8) {
        const uint4 a = __ldcs(reinterpret_cast<const uint4*>(p));      // messages are read exactly once
        unpack_bf16x2(a.x, r.v[0], r.v[1]); unpack_bf16x2(a.y, r.v[2], r.v[3]);
        unpack_bf16x2(a.z, r.v[4], r.v[5]); unpack_bf16x2(a.w, r.v[6], r.v[7]);
    } else {
        const uint2 a = __ldg(reinterpret_cast<const uint2*>(p));
        unpack_bf16x2(a.x, r.v[0], r.v[1]); unpack_bf16x2(a.y, r.v[2], r.v[3]);
    }
    return r;
}
template <int V>
__device__ __forceinline__ void storev(float* p, const VecF<V>& a) {
#pragma unroll
    for (int k = 0; k < V / 4; ++k)
        reinterpret_cast<float4*>(p)[k] = make_float4(a.v[4 * k], a.v[4 * k + 1], a.v[4 * k + 2], a.v[4 * k + 3]);
}
template <int V>
__device__ __forceinline__ void storev(__nv_bfloat16* p, const VecF<V>& a) {
#pragma unroll
    for (int k = 0; k < V / 4; ++k)
        reinterpret_cast<uint2*>(p)[k] = make_uint2(pack_bf16x2(a.v[4 * k], a.v[4 * k + 1]), pack_bf16x2(a.v[4 * k + 2], a.v[4 * k + 3]));
}
__device__ __forceinline__ float4 load4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

template <typename MT, int V, typename OT>
__global__ void __launch_bounds__(256) k_row_sum(const int32_t* __restrict__ rowptr, int64_t nrows, int O,
                                                 const MT* __restrict__ msg, const float* __restrict__ bias,
                                                 OT* __restrict__ out) {
    const int cg = O / V;
    const int64_t total = nrows * cg;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = idx / cg;
        const int q = (int)(idx - row * cg);
        const int e0 = rowptr[row], e1 = rowptr[row + 1];
        if (e1 - e0 > RGCN_LONG_ROW) continue;                 // hub rows: k_row_sum_long
        VecF<V> acc;
#pragma unroll
        for (int j = 0; j < V; ++j) acc.v[j] = 0.f;
        const MT* m = msg + (size_t)e0 * O + V * q;
        int e = e0;
        for (; e + 1 < e1; e += 2, m += 2 * (size_t)O) {
            const VecF<V> a = loadv<V>(m), b = loadv<V>(m + O);
#pragma unroll
            for (int j = 0; j < V; ++j) acc.v[j] += a.v[j] + b.v[j];
        }
        if (e < e1) {
            const VecF<V> a = loadv<V>(m);
#pragma unroll
            for (int j = 0; j < V; ++j) acc.v[j] += a.v[j];
        }
        if (bias) {
            const VecF<V> b = loadv<V>(bias + V * q);
#pragma unroll
            for (int j = 0; j < V; ++j) acc.v[j] += b.v[j];
        }
        storev<V>(out + (size_t)row * O + V * q, acc);
    }
}

// hub rows (more than RGCN_LONG_ROW messages): one CTA per row, edge-parallel partial sums (4 independent loads in
// flight per thread), shared-memory reduction
template <typename MT, int V, typename OT>
__global__ void __launch_bounds__(256) k_row_sum_long(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ list,
                                                      const int32_t* __restrict__ count, int O, const MT* __restrict__ msg,
                                                      const float* __restrict__ bias, OT* __restrict__ out) {
    __shared__ float red[256 * V];
    if ((int)blockIdx.x >= *count) return;
    const int row = list[blockIdx.x];
    const int e0 = rowptr[row], e1 = rowptr[row + 1];
    const int cg = O / V;
    const int cgb = cg < 256 ? cg : 256;                       // column groups handled per pass
    const int lanes = 256 / cgb;                               // edge lanes per column group
    for (int q0 = 0; q0 < cg; q0 += 256) {
        const int q = q0 + (int)threadIdx.x % cgb, lane = threadIdx.x / cgb;
        VecF<V> acc;
#pragma unroll
        for (int j = 0; j < V; ++j) acc.v[j] = 0.f;
        if (q < cg && lane < lanes) {
            const MT* m = msg + V * q;
            int e = e0 + lane;
            for (; e + 3 * lanes < e1; e += 4 * lanes) {
                const VecF<V> a = loadv<V>(m + (size_t)e * O), b = loadv<V>(m + (size_t)(e + lanes) * O);
                const VecF<V> c = loadv<V>(m + (size_t)(e + 2 * lanes) * O), d = loadv<V>(m + (size_t)(e + 3 * lanes) * O);
#pragma unroll
                for (int j = 0; j < V; ++j) acc.v[j] += (a.v[j] + b.v[j]) + (c.v[j] + d.v[j]);
            }
            for (; e < e1; e += lanes) {
                const VecF<V> a = loadv<V>(m + (size_t)e * O);
#pragma unroll
                for (int j = 0; j < V; ++j) acc.v[j] += a.v[j];
            }
        }
#pragma unroll
        for (int j = 0; j < V; ++j) red[threadIdx.x * V + j] = acc.v[j];
        __syncthreads();
        if (lane == 0 && q < cg) {
            for (int l = 1; l < lanes; ++l)
#pragma unroll
                for (int j = 0; j < V; ++j) acc.v[j] += red[(threadIdx.x + l * cgb) * V + j];
            if (bias) {
                const VecF<V> b = loadv<V>(bias + V * q);
#pragma unroll
                for (int j = 0; j < V; ++j) acc.v[j] += b.v[j];
            }
            storev<V>(out + (size_t)row * O + V * q, acc);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------------
// weight gradient: rows of X and G staged through shared memory with cp.async (3-stage ring), then
// TI x TJ register tiles of gW_p accumulated with FMAs; one atomic flush per chunk
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    int sz = valid ? 16 : 0;                        // src-size 0 -> the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gmem_src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

template <int n>
__device__ __forceinline__ void lds_slice(const float* p, float (&x)[n]) {
#pragma unroll
    for (int k = 0; k < n / 4; ++k) {
        float4 v = reinterpret_cast<const float4*>(p)[k];
        x[4 * k] = v.x; x[4 * k + 1] = v.y; x[4 * k + 2] = v.z; x[4 * k + 3] = v.w;
    }
}
template <int n>
__device__ __forceinline__ void lds_slice(const __nv_bfloat16* p, float (&x)[n]) {
    if constexpr (n % 8 == 0) {
#pragma unroll
        for (int k = 0; k < n / 8; ++k) {
            uint4 v = reinterpret_cast<const uint4*>(p)[k];
            unpack_bf16x2(v.x, x[8 * k], x[8 * k + 1]); unpack_bf16x2(v.y, x[8 * k + 2], x[8 * k + 3]);
            unpack_bf16x2(v.z, x[8 * k + 4], x[8 * k + 5]); unpack_bf16x2(v.w, x[8 * k + 6], x[8 * k + 7]);
        }
    } else {
#pragma unroll
        for (int k = 0; k < n / 4; ++k) {
            uint2 v = reinterpret_cast<const uint2*>(p)[k];
            unpack_bf16x2(v.x, x[4 * k], x[4 * k + 1]); unpack_bf16x2(v.y, x[4 * k + 2], x[4 * k + 3]);
        }
    }
}

constexpr int kWgStages = 3;

// All shape parameters are compile-time so that the staging and tile addressing reduce to constants
// (the first version of this kernel spent 3/4 of its instructions on index arithmetic).
template <typename XT, int BI, int BO, int NB>
__global__ void __launch_bounds__(256) k_rel_wgrad(RelArgs A, const XT* __restrict__ X, const float* __restrict__ G,
                                                   float* __restrict__ gW) {
    constexpr int TI = BI < 8 ? BI : 8, TJ = BO < 8 ? BO : 8;
    constexpr int TPB = (BI / TI) * (BO / TJ);        // threads covering one (BI, BO) block
    constexpr int TC = NB * TPB;                       // threads covering the whole relation weight
    constexpr int SETS = 256 / TC;                     // split-K groups
    static_assert(TC <= 256 && SETS >= 1, "weight does not fit one CTA");
    constexpr int I = NB * BI, O = NB * BO;
    constexpr int RX = I * (int)sizeof(XT), RG = O * 4;   // row bytes
    constexpr int NZ = NB * BI * BO;
    constexpr int OPS = (RX + RG) / 16;                // 16-byte copies per edge
    constexpr int TE_RAW = (36 * 1024) / (kWgStages * (RX + RG));
    constexpr int TE_CAP = TE_RAW > 64 ? 64 : TE_RAW;
    constexpr int TE = TE_CAP < SETS ? SETS : (TE_CAP / SETS) * SETS;                     // multiple of SETS
    constexpr int STAGE = TE * (RX + RG);
    static_assert(kWgStages * STAGE >= NZ * 4, "ring too small for the split-K reduction");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int c = blockIdx.x;
    if (c >= A.chunkptr[A.num_rels]) return;
    int p, e0, e1;
    chunk_lookup(A, c, p, e0, e1);
    int32_t* s_src = reinterpret_cast<int32_t*>(smem_raw);
    int32_t* s_dst = s_src + RGCN_CHUNK_EDGES;
    float* s_val = reinterpret_cast<float*>(s_dst + RGCN_CHUNK_EDGES);
    unsigned char* ring = reinterpret_cast<unsigned char*>(s_val + RGCN_CHUNK_EDGES);
    const int n = e1 - e0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        s_src[i] = A.gather[e0 + i]; s_dst[i] = A.other[e0 + i]; s_val[i] = A.val[e0 + i];
    }
    __syncthreads();
    const int ntiles = (n + TE - 1) / TE;
    const unsigned char* Xb = reinterpret_cast<const unsigned char*>(X);
    const unsigned char* Gb = reinterpret_cast<const unsigned char*>(G);
    auto issue = [&](int tile) {
        if (tile < ntiles) {
            unsigned char* st = ring + (tile % kWgStages) * STAGE;
#pragma unroll
            for (int k = threadIdx.x; k < TE * OPS; k += 256) {
                const int t = k / OPS, piece = k - t * OPS, le = tile * TE + t;      // OPS is a constant
                const bool ok = le < n;
                const int li = ok ? le : 0;
                if (piece * 16 < RX)
                    cp_async16(st + t * RX + piece * 16, Xb + (size_t)s_src[li] * RX + piece * 16, ok);
                else
                    cp_async16(st + TE * RX + t * RG + (piece * 16 - RX), Gb + (size_t)s_dst[li] * RG + (piece * 16 - RX), ok);
            }
        }
        cp_async_commit();
    };
    const int set = threadIdx.x / TC, r = threadIdx.x % TC;
    const int blk = r / TPB, q = r % TPB, ti = q / (BO / TJ), tj = q % (BO / TJ);
    const bool active = set < SETS;
    float acc[TI][TJ];
#pragma unroll
    for (int i = 0; i < TI; ++i)
#pragma unroll
        for (int j = 0; j < TJ; ++j) acc[i][j] = 0.f;

    issue(0);
    issue(1);
    for (int tile = 0; tile < ntiles; ++tile) {
        issue(tile + 2);
        cp_async_wait<2>();                           // this tile has landed (two younger groups may be in flight)
        __syncthreads();
        const unsigned char* st = ring + (tile % kWgStages) * STAGE;
        if (active) {
            const XT* xs = reinterpret_cast<const XT*>(st) + blk * BI + ti * TI;
            const float* gs = reinterpret_cast<const float*>(st + TE * RX) + blk * BO + tj * TJ;
            const float* vs = s_val + tile * TE;
#pragma unroll 4
            for (int t = set; t < TE; t += SETS) {    // rows past the chunk end were zero-filled by cp.async
                float x[TI], g[TJ];
                lds_slice<TI>(xs + t * I, x);
                lds_slice<TJ>(gs + t * O, g);
                const float v = (tile * TE + t < n) ? vs[t] : 0.f;
#pragma unroll
                for (int i = 0; i < TI; ++i) {
                    const float xv = x[i] * v;
#pragma unroll
                    for (int j = 0; j < TJ; ++j) acc[i][j] += xv * g[j];
                }
            }
        }
        __syncthreads();                              // everyone is done with this stage before it is refilled
    }
    cp_async_wait<0>();
    // reduce the split-K partials: first across the sets that live in the same warp (shuffles), then across warps
    // through shared memory (reusing the ring) in at most 8 rounds
    if constexpr (TC < 32) {
#pragma unroll
        for (int off = TC; off < 32; off <<= 1)
#pragma unroll
            for (int i = 0; i < TI; ++i)
#pragma unroll
                for (int j = 0; j < TJ; ++j) acc[i][j] += __shfl_xor_sync(0xffffffffu, acc[i][j], off);
    }
    constexpr int ROUNDS = TC < 32 ? 8 : SETS;
    const int my_round = TC < 32 ? (int)(threadIdx.x >> 5) : set;
    const bool writer = TC < 32 ? ((threadIdx.x & 31) < TC) : active;
    float* red = reinterpret_cast<float*>(ring);
#pragma unroll 1
    for (int s = 0; s < ROUNDS; ++s) {
        if (writer && my_round == s) {
#pragma unroll
            for (int i = 0; i < TI; ++i)
#pragma unroll
                for (int j = 0; j < TJ; ++j) {
                    const int idx = blk * BI * BO + (ti * TI + i) * BO + tj * TJ + j;
                    red[idx] = (s == 0) ? acc[i][j] : red[idx] + acc[i][j];
                }
        }
        __syncthreads();
    }
    float* dst = gW + (size_t)p * NZ;
    for (int i = threadIdx.x; i < NZ; i += blockDim.x) {
        const float v = red[i];
        if (v != 0.f) atomicAdd(dst + i, v);
    }
}

// ------------------------------------------------------------------------------------------------------
// host dispatch
// ------------------------------------------------------------------------------------------------------
struct RelShape { int nb, bi, bo; };

// which (XT, BI, BO) the relation-batched kernels are instantiated for
inline bool rel_shape_supported(int nb, int bi, int bo, bool bf16) {
    if (nb < 1 || nb > 256) return false;
    bool pair = (bi == 8 && bo == 8) || (bi == 16 && bo == 16) || (bi == 16 && bo == 4) || (bi == 4 && bo == 16);
    if (!pair) return false;
    if (bf16 && bi % 8 != 0) return false;
    return true;
}

template <typename XT, typename MT, int BI, int BO>
int launch_rel_transform_t(const RelArgs& A, const XT* X, MT* msg, int max_chunks, cudaStream_t st) {
    constexpr int EPT = 2;                           // 4 edges/thread was measured slower (am16: 0.73 -> 0.91 ms)
    const size_t smem = (size_t)A.nb * (BI * BO + 4) * sizeof(float);
    auto kern = k_rel_transform<XT, MT, BI, BO, EPT>;
    if (smem > 48 * 1024) RGCN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RGCN_LAUNCH(kern, max_chunks, 256, smem, st, A, X, msg);
    return RGCN_OK;
}

template <typename XT, typename MT>
int launch_rel_transform(const RelArgs& A, int bi, int bo, const XT* X, MT* msg, int max_chunks, cudaStream_t st) {
    if (bi == 8 && bo == 8) return launch_rel_transform_t<XT, MT, 8, 8>(A, X, msg, max_chunks, st);
    if (bi == 16 && bo == 16) return launch_rel_transform_t<XT, MT, 16, 16>(A, X, msg, max_chunks, st);
    if (bi == 16 && bo == 4) return launch_rel_transform_t<XT, MT, 16, 4>(A, X, msg, max_chunks, st);
    if constexpr (sizeof(XT) == 4) {
        if (bi == 4 && bo == 16) return launch_rel_transform_t<XT, MT, 4, 16>(A, X, msg, max_chunks, st);
    }
    set_error("relation-batched transform: unsupported block %dx%d", bi, bo);
    return RGCN_ERR_UNSUPPORTED;
}

template <typename MT, int V, typename OT>
int launch_row_sum_v(const int32_t* rowptr, int64_t nrows, int O, const MT* msg, const float* bias, OT* out,
                     const int32_t* long_list, const int32_t* long_count, int64_t num_long, int64_t nnz, cudaStream_t st) {
    int64_t total = nrows * (O / V);
    int64_t want = (total + 255) / 256;
    int grid = (int)(want < (int64_t)kNumSMs * 32 ? want : (int64_t)kNumSMs * 32);
    if (grid < 1) grid = 1;
    RGCN_LAUNCH((k_row_sum<MT, V, OT>), grid, 256, 0, st, rowptr, nrows, O, msg, bias, out);
    // exact count when the plan read it back, else the bound: at most nnz / RGCN_LONG_ROW rows can be that long
    const int bound = num_long >= 0 ? (int)num_long : (int)(nnz / RGCN_LONG_ROW);
    if (bound > 0) RGCN_LAUNCH((k_row_sum_long<MT, V, OT>), bound, 256, 0, st, rowptr, long_list, long_count, O, msg, bias, out);
    return RGCN_OK;
}

template <typename MT, typename OT>
int launch_row_sum(const int32_t* rowptr, int64_t nrows, int O, const MT* msg, const float* bias, OT* out,
                   const int32_t* long_list, const int32_t* long_count, int64_t num_long, int64_t nnz, cudaStream_t st) {
    if (O % 8 == 0) return launch_row_sum_v<MT, 8, OT>(rowptr, nrows, O, msg, bias, out, long_list, long_count, num_long, nnz, st);
    return launch_row_sum_v<MT, 4, OT>(rowptr, nrows, O, msg, bias, out, long_list, long_count, num_long, nnz, st);
}

// fp32 -> bf16 copy (feature gradient of paths that produce fp32 when the caller asked for bf16)
__global__ void k_cast_bf16(const float* __restrict__ in, int64_t n, __nv_bfloat16* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = __float2bfloat16_rn(in[i]);
}

template <typename XT, int BI, int BO, int NB>
int launch_rel_wgrad_t(const RelArgs& A, const XT* X, const float* G, float* gW, int max_chunks, cudaStream_t st) {
    constexpr int row_bytes = NB * BI * (int)sizeof(XT) + NB * BO * 4;
    constexpr int TI = BI < 8 ? BI : 8, TJ = BO < 8 ? BO : 8;
    constexpr int SETS = 256 / (NB * (BI / TI) * (BO / TJ));
    constexpr int TE_RAW = (36 * 1024) / (kWgStages * row_bytes);
    constexpr int TE_CAP = TE_RAW > 64 ? 64 : TE_RAW;
    constexpr int TE = TE_CAP < SETS ? SETS : (TE_CAP / SETS) * SETS;
    const size_t smem = 3 * RGCN_CHUNK_EDGES * sizeof(int32_t) + (size_t)kWgStages * TE * row_bytes;
    auto kern = k_rel_wgrad<XT, BI, BO, NB>;
    if (smem > 48 * 1024) RGCN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RGCN_LAUNCH(kern, max_chunks, 256, smem, st, A, X, G, gW);
    return RGCN_OK;
}

template <typename XT, int BI, int BO>
int launch_rel_wgrad_b(const RelArgs& A, const XT* X, const float* G, float* gW, int max_chunks, cudaStream_t st) {
    switch (A.nb) {
        case 1: return launch_rel_wgrad_t<XT, BI, BO, 1>(A, X, G, gW, max_chunks, st);
        case 2: return launch_rel_wgrad_t<XT, BI, BO, 2>(A, X, G, gW, max_chunks, st);
        case 4: return launch_rel_wgrad_t<XT, BI, BO, 4>(A, X, G, gW, max_chunks, st);
        case 8: return launch_rel_wgrad_t<XT, BI, BO, 8>(A, X, G, gW, max_chunks, st);
        default: return 1;   // not instantiated: caller falls back to the generic kernel
    }
}

// returns > 0 when this (block shape, block count) is not instantiated
template <typename XT>
int launch_rel_wgrad(const RelArgs& A, int bi, int bo, const XT* X, const float* G, float* gW, int max_chunks,
                     cudaStream_t st) {
    if (bi == 8 && bo == 8) return launch_rel_wgrad_b<XT, 8, 8>(A, X, G, gW, max_chunks, st);
    if (bi == 16 && bo == 16) return launch_rel_wgrad_b<XT, 16, 16>(A, X, G, gW, max_chunks, st);
    if (bi == 16 && bo == 4) return launch_rel_wgrad_b<XT, 16, 4>(A, X, G, gW, max_chunks, st);
    if constexpr (sizeof(XT) == 4) {
        if (bi == 4 && bo == 16) return launch_rel_wgrad_b<XT, 4, 16>(A, X, G, gW, max_chunks, st);
    }
    return 1;
}

}  // namespace rgcn
