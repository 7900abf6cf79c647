// Generic (any dims, any weight form) propagation kernels.
//
//   out[row] = bias + sum_{e in row} val_e * T_{rel_e}(src[col_e])
//
// walked on a CSR whose rows are destination nodes (forward) or source nodes (backward to
// features, with transposed weights).  Edges of a row are sorted by relation, so the kernel first
// aggregates the (row, relation) segment and applies the relation's weight once per segment
// (aggregate-then-transform: the reference's "vertical" order, layers.py:293-297, without ever
// materialising the (R', N, I) temporary).
//
// These kernels favour generality over speed: one warp per row, lane-strided scalar loads.  The
// shapes the benchmark configs use are served by the specialised kernels in propagate_fast.cuh.
#pragma once
#include "common.cuh"

namespace rgcn {

struct PropArgs {
    // graph view
    const int32_t* rowptr; const int32_t* col; const int32_t* rel; const float* val;
    int64_t nrows;
    int64_t N;                // number of nodes (row stride of featureless weight tables)
    // weights
    int form; int featureless;
    int I, O, B, nb, bi, bo;
    int self_rel;             // relation whose weight is the dense `blocks_self`, or -1
    const float* W;           // DENSE (R', I, O) | DIAG (R', I)
    const float* bases; const float* comps; const float* blocks; const float* blocks_self;
    const float* bias;
    const float* out_mask;    // (nrows, O): multiplies the transformed segment of relation mask_rel
    const float* in_mask;     // (N, I): multiplies source rows of relation mask_rel before aggregation
    int mask_rel;
    int skip_rel_plus1;         // 0: none; else edges of relation skip_rel_plus1 - 1 are left to another kernel
    float* out;
    const int32_t* long_list;   // rows with more than RGCN_LONG_ROW edges (may be NULL: no special handling)
    const int32_t* long_count;
    int long_mode;              // 0: all rows, long rows only get bias; 1: long rows only, split over warps, atomics
    int64_t nnz_hint;           // nnz (host-side bound on the number of long rows)
    int64_t num_long;           // exact number of long rows if known on the host, else -1
};

template <typename XT, int K>
__global__ void __launch_bounds__(256) k_prop_generic(PropArgs A, const XT* __restrict__ X) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    float* a_s = smem + (size_t)warp * A.I;      // staged segment aggregate (featured forms only)
    const int I = A.I, O = A.O;

    // long_mode 1: blockIdx.x = index into the long-row list, the row's edges are split into 256-edge pieces that
    // the warps of gridDim.y CTAs take round-robin; partial results are added atomically (the sum is linear)
    const bool long_mode = A.long_mode == 1;
    if (long_mode && (int)blockIdx.x >= *A.long_count) return;
    const int64_t row_begin = long_mode ? A.long_list[blockIdx.x] : (int64_t)blockIdx.x * wpb + warp;
    const int64_t row_step = long_mode ? A.nrows : (int64_t)gridDim.x * wpb;
    for (int64_t row = row_begin; row < A.nrows; row += row_step) {
        const int e0 = A.rowptr[row], e1 = A.rowptr[row + 1];
        float acc[K], a[K];
#pragma unroll
        for (int k = 0; k < K; ++k) { acc[k] = 0.f; a[k] = 0.f; }
        int cur = -1;

        auto transform = [&](int p) {
            if (A.form == RGCN_W_DIAG) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    int i = lane + 32 * k;
                    if (i < I) acc[k] += a[k] * A.W[(size_t)p * I + i];
                }
                return;
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < K; ++k) {
                int i = lane + 32 * k;
                if (i < I) a_s[i] = a[k];
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < K; ++k) {
                int j = lane + 32 * k;
                if (j < O) {
                    float s = 0.f;
                    if (A.form == RGCN_W_BLOCK && p != A.self_rel) {
                        int kb = j / A.bo, jj = j - kb * A.bo;
                        const float* w = A.blocks + (((size_t)p * A.nb + kb) * A.bi) * A.bo + jj;
                        const float* as = a_s + kb * A.bi;
                        for (int ii = 0; ii < A.bi; ++ii) s += as[ii] * w[(size_t)ii * A.bo];
                    } else {
                        const float* w = (A.form == RGCN_W_BLOCK) ? A.blocks_self + j : A.W + (size_t)p * I * O + j;
                        for (int i = 0; i < I; ++i) s += a_s[i] * w[(size_t)i * O];
                    }
                    if (A.out_mask && p == A.mask_rel) s *= A.out_mask[(size_t)row * O + j];
                    acc[k] += s;
                }
            }
        };

        auto walk = [&](int eb, int ee) {
            for (int e = eb; e < ee; ++e) {
                const int r = A.rel[e], c = A.col[e];
                const float v = A.val[e];
                if (A.featureless) {
                    if (A.form == RGCN_W_DENSE) {
                        const float* w = A.W + ((size_t)r * A.N + c) * O;
#pragma unroll
                        for (int k = 0; k < K; ++k) { int j = lane + 32 * k; if (j < O) acc[k] += v * w[j]; }
                    } else if (A.form == RGCN_W_BASIS) {
                        for (int b = 0; b < A.B; ++b) {
                            const float cb = v * A.comps[(size_t)r * A.B + b];
                            const float* w = A.bases + ((size_t)b * A.N + c) * O;
#pragma unroll
                            for (int k = 0; k < K; ++k) { int j = lane + 32 * k; if (j < O) acc[k] += cb * w[j]; }
                        }
                    } else {   // BLOCK: row c of blockdiag(blocks[r]) lives in block c / bi, columns [kb*bo, kb*bo+bo)
                        const int kb = c / A.bi, ii = c - kb * A.bi;
                        const float* w = A.blocks + (((size_t)r * A.nb + kb) * A.bi + ii) * A.bo;
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            int j = lane + 32 * k, jj = j - kb * A.bo;
                            if (j < O && jj >= 0 && jj < A.bo) acc[k] += v * w[jj];
                        }
                    }
                    continue;
                }
                if (r + 1 == A.skip_rel_plus1) continue;
                if (r != cur) {
                    if (cur >= 0) transform(cur);
#pragma unroll
                    for (int k = 0; k < K; ++k) a[k] = 0.f;
                    cur = r;
                }
                const XT* xr = X + (size_t)c * I;
                const float* mr = (A.in_mask && r == A.mask_rel) ? A.in_mask + (size_t)c * I : nullptr;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    int i = lane + 32 * k;
                    if (i < I) {
                        float x = to_f32(xr[i]);
                        if (mr) x *= mr[i];
                        a[k] += v * x;
                    }
                }
            }
            if (cur >= 0) transform(cur);
            cur = -1;
        };

        if (long_mode) {
            const int nsplit = (e1 - e0 + 255) / 256, nwarps = gridDim.y * wpb;
            bool any = false;
            for (int sp = blockIdx.y * wpb + warp; sp < nsplit; sp += nwarps) {
                walk(e0 + sp * 256, min(e1, e0 + (sp + 1) * 256));
                any = true;
            }
            if (any) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    int j = lane + 32 * k;
                    if (j < O) atomicAdd(A.out + (size_t)row * O + j, acc[k]);
                }
            }
            continue;
        }
        const bool is_long = A.long_list && (e1 - e0 > RGCN_LONG_ROW);   // handled by the long_mode launch
        if (!is_long) walk(e0, e1);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            int j = lane + 32 * k;
            if (j < O) A.out[(size_t)row * O + j] = acc[k] + (A.bias ? A.bias[j] : 0.f);
        }
    }
}

// ------------------------------------------------------------------------------------------
// weight gradient, relation-major:  gW_p += sum_{e in p} val_e * X[o_e]^T (G[s_e] (*) mask)
// ------------------------------------------------------------------------------------------
struct WGradArgs {
    const int32_t* relptr; const int32_t* dst; const int32_t* src; const float* val;
    int form; int I, O, nb, bi, bo; int self_rel; int num_block_rels;
    const float* mask; int mask_rel;   // (N, O) on G rows of relation mask_rel
    float* gW;        // DENSE (R', I, O) | DIAG (R', I)
    float* gblocks;   // (Rb, nb, bi, bo)
    float* gself;     // (I, O)
    int tile;         // edges staged per smem tile
    int rel0;         // tiled dense kernel: first relation of the launch, and the relation gW[0] belongs to
};

template <typename XT, int KE>
__global__ void __launch_bounds__(256) k_wgrad_generic(WGradArgs A, const XT* __restrict__ X, const float* __restrict__ G) {
    extern __shared__ float smem[];
    const int p = blockIdx.x;
    const int e0 = A.relptr[p], e1 = A.relptr[p + 1];
    const int n = e1 - e0;
    if (n <= 0) return;
    const int per = (n + gridDim.y - 1) / gridDim.y;
    const int b0 = e0 + blockIdx.y * per;
    const int b1 = min(e1, b0 + per);
    if (b0 >= b1) return;
    const int I = A.I, O = A.O, TE = A.tile;
    float* Xs = smem;                       // [TE][I], pre-multiplied by val
    float* Gs = smem + (size_t)TE * I;      // [TE][O]

    int kind;                               // 0 dense, 1 block, 2 diag
    float* dest;
    int nel;
    if (A.form == RGCN_W_DIAG) {
        if (!A.gW) return;
        kind = 2; nel = I; dest = A.gW + (size_t)p * I;
    } else if (A.form == RGCN_W_BLOCK && p != A.self_rel) {
        if (!A.gblocks || p >= A.num_block_rels) return;
        kind = 1; nel = A.nb * A.bi * A.bo; dest = A.gblocks + (size_t)p * nel;
    } else if (A.form == RGCN_W_BLOCK) {
        if (!A.gself) return;
        kind = 0; nel = I * O; dest = A.gself;
    } else {
        if (!A.gW) return;
        kind = 0; nel = I * O; dest = A.gW + (size_t)p * nel;
    }
    const bool masked = A.mask && p == A.mask_rel;

    for (int slab = 0; slab < nel; slab += blockDim.x * KE) {
        float acc[KE];
        int ei[KE], ej[KE];
#pragma unroll
        for (int k = 0; k < KE; ++k) {
            acc[k] = 0.f;
            int el = slab + k * blockDim.x + threadIdx.x;
            if (el < nel) {
                if (kind == 0) { ei[k] = el / O; ej[k] = el - ei[k] * O; }
                else if (kind == 1) {
                    int kb = el / (A.bi * A.bo), rem = el - kb * A.bi * A.bo;
                    int ii = rem / A.bo, jj = rem - ii * A.bo;
                    ei[k] = kb * A.bi + ii; ej[k] = kb * A.bo + jj;
                } else { ei[k] = el; ej[k] = el; }
            } else { ei[k] = -1; ej[k] = 0; }
        }
        for (int t0 = b0; t0 < b1; t0 += TE) {
            __syncthreads();
            for (int idx = threadIdx.x; idx < TE * I; idx += blockDim.x) {
                int t = idx / I, i = idx - t * I, e = t0 + t;
                Xs[idx] = (e < b1) ? A.val[e] * to_f32(X[(size_t)A.src[e] * I + i]) : 0.f;
            }
            for (int idx = threadIdx.x; idx < TE * O; idx += blockDim.x) {
                int t = idx / O, j = idx - t * O, e = t0 + t;
                float g = 0.f;
                if (e < b1) {
                    size_t off = (size_t)A.dst[e] * O + j;
                    g = G[off];
                    if (masked) g *= A.mask[off];
                }
                Gs[idx] = g;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < KE; ++k) {
                if (ei[k] >= 0) {
                    float s = 0.f;
                    for (int t = 0; t < TE; ++t) s += Xs[t * I + ei[k]] * Gs[t * O + ej[k]];
                    acc[k] += s;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < KE; ++k) {
            int el = slab + k * blockDim.x + threadIdx.x;
            if (ei[k] >= 0 && acc[k] != 0.f) atomicAdd(dest + el, acc[k]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// featureless weight gradients, walked on the source-major CSR (rows = object o):
//   dense: gW[p, o, :]  = sum_{e=(s,p,o)} val_e G[s]
//   block: the same row lands in gblocks[p, o / bi, o % bi, :]
//   basis: gbases[b, o, :] = sum_e val_e comps[p,b] G[s];  gcomps[p,b] += <bases[b,o,:], sum_e val_e G[s]>
// ------------------------------------------------------------------------------------------
struct FeaturelessGradArgs {
    const int32_t* rowptr; const int32_t* col; const int32_t* rel; const float* val;
    int64_t N; int num_rels;
    int form; int O, B, nb, bi, bo;
    const float* comps; const float* bases;
    float* gW; float* gblocks; float* gbases; float* gcomps;
    int comps_in_smem;       // R'*B floats of block-shared accumulators follow the per-warp areas
};

template <int K>
__global__ void __launch_bounds__(256) k_wgrad_featureless(FeaturelessGradArgs A, const float* __restrict__ G) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int O = A.O, B = A.B;
    float* gb_s = smem + (size_t)warp * B * O;                 // per-warp (B, O) accumulator (basis only)
    float* gc_s = smem + (size_t)wpb * B * O;                  // block-shared (R', B) accumulator
    if (A.form == RGCN_W_BASIS && A.comps_in_smem) {
        for (int i = threadIdx.x; i < A.num_rels * B; i += blockDim.x) gc_s[i] = 0.f;
        __syncthreads();
    }
    for (int64_t row = (int64_t)blockIdx.x * wpb + warp; row < A.N; row += (int64_t)gridDim.x * wpb) {
        const int e0 = A.rowptr[row], e1 = A.rowptr[row + 1];
        if (A.form == RGCN_W_BASIS) {
            for (int i = lane; i < B * O; i += 32) gb_s[i] = 0.f;
            __syncwarp();
        }
        int e = e0;
        while (e < e1) {
            const int p = A.rel[e];
            float t[K];
#pragma unroll
            for (int k = 0; k < K; ++k) t[k] = 0.f;
            for (; e < e1 && A.rel[e] == p; ++e) {
                const float v = A.val[e];
                const float* g = G + (size_t)A.col[e] * O;
#pragma unroll
                for (int k = 0; k < K; ++k) { int j = lane + 32 * k; if (j < O) t[k] += v * g[j]; }
            }
            if (A.form == RGCN_W_DENSE) {
                float* d = A.gW + ((size_t)p * A.N + row) * O;
#pragma unroll
                for (int k = 0; k < K; ++k) { int j = lane + 32 * k; if (j < O) d[j] = t[k]; }
            } else if (A.form == RGCN_W_BLOCK) {
                const int kb = (int)(row / A.bi), ii = (int)(row - (int64_t)kb * A.bi);
                float* d = A.gblocks + (((size_t)p * A.nb + kb) * A.bi + ii) * A.bo;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    int j = lane + 32 * k, jj = j - kb * A.bo;
                    if (j < O && jj >= 0 && jj < A.bo) d[jj] = t[k];
                }
            } else if (K == 1 && B <= 32) {
                // basis, O <= 32: lanes own output columns for gbases and basis indices for gcomps, so the
                // <bases[b, o, :], t> products need no cross-lane reduction (one atomic per (segment, basis))
                const float tj = t[0];
                for (int b = 0; b < B; ++b)
                    if (lane < O) gb_s[b * O + lane] += A.comps[(size_t)p * B + b] * tj;
                const float* bs = A.bases + ((size_t)(lane < B ? lane : 0) * A.N + row) * O;
                float dot = 0.f;
                for (int j = 0; j < O; ++j) {
                    const float tv = __shfl_sync(0xffffffffu, tj, j);
                    if (lane < B) dot += bs[j] * tv;
                }
                if (lane < B) atomicAdd((A.comps_in_smem ? gc_s : A.gcomps) + (size_t)p * B + lane, dot);
            } else {
                for (int b = 0; b < B; ++b) {
                    const float c = A.comps[(size_t)p * B + b];
                    const float* bs = A.bases + ((size_t)b * A.N + row) * O;
                    float dot = 0.f;
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        int j = lane + 32 * k;
                        if (j < O) { gb_s[b * O + j] += c * t[k]; dot += bs[j] * t[k]; }
                    }
#pragma unroll
                    for (int off = 16; off; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
                    if (lane == 0) atomicAdd((A.comps_in_smem ? gc_s : A.gcomps) + (size_t)p * B + b, dot);
                }
            }
        }
        if (A.form == RGCN_W_BASIS) {
            __syncwarp();
            for (int i = lane; i < B * O; i += 32) {
                int b = i / O, j = i - b * O;
                A.gbases[((size_t)b * A.N + row) * O + j] = gb_s[i];
            }
            __syncwarp();
        }
    }
    if (A.form == RGCN_W_BASIS && A.comps_in_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < A.num_rels * B; i += blockDim.x)
            if (gc_s[i] != 0.f) atomicAdd(A.gcomps + i, gc_s[i]);
    }
}

// ------------------------------------------------------------------------------------------
// small dense helpers
// ------------------------------------------------------------------------------------------
// W[r, x] = sum_b comps[r, b] * bases[b, x]           (layers.py:242 / :469)
__global__ void k_basis_combine(const float* __restrict__ comps, const float* __restrict__ bases, int Rp, int B,
                                int64_t IO, float* __restrict__ W) {
    int64_t x = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int r = blockIdx.y;
    if (x >= IO) return;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += comps[(size_t)r * B + b] * bases[(size_t)b * IO + x];
    W[(size_t)r * IO + x] = s;
}

// gcomps[r, b] = <gW[r], bases[b]>
__global__ void k_basis_grad_comps(const float* __restrict__ gW, const float* __restrict__ bases, int B, int64_t IO,
                                   float* __restrict__ gcomps) {
    __shared__ float red[256];
    int r = blockIdx.x, b = blockIdx.y;
    float s = 0.f;
    for (int64_t x = threadIdx.x; x < IO; x += blockDim.x) s += gW[(size_t)r * IO + x] * bases[(size_t)b * IO + x];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = blockDim.x / 2; w; w >>= 1) {
        if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) gcomps[(size_t)r * B + b] = red[0];
}

// gbases[b, x] = sum_r comps[r, b] * gW[r, x]
__global__ void k_basis_grad_bases(const float* __restrict__ gW, const float* __restrict__ comps, int Rp, int B,
                                   int64_t IO, float* __restrict__ gbases) {
    int64_t x = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int b = blockIdx.y;
    if (x >= IO) return;
    float s = 0.f;
    for (int r = 0; r < Rp; ++r) s += comps[(size_t)r * B + b] * gW[(size_t)r * IO + x];
    gbases[(size_t)b * IO + x] = s;
}

// out[n, c, r] = in[n, r, c]
__global__ void k_transpose_batched(const float* __restrict__ in, int64_t n, int rows, int cols, float* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t per = (int64_t)rows * cols;
    if (i >= n * per) return;
    int64_t m = i / per, rem = i - m * per;
    int c = (int)(rem / rows), r = (int)(rem - (int64_t)c * rows);
    out[i] = in[m * per + (int64_t)r * cols + c];
}

// gbias[j] = sum_s G[s, j]
__global__ void k_colsum(const float* __restrict__ G, int64_t N, int O, int64_t rows_per_block, float* __restrict__ out) {
    extern __shared__ float sums[];
    for (int j = threadIdx.x; j < O; j += blockDim.x) sums[j] = 0.f;
    __syncthreads();
    int64_t r0 = blockIdx.x * rows_per_block, r1 = min(N, r0 + rows_per_block);
    if (O <= (int)blockDim.x) {
        int lanes = blockDim.x / O;                 // row lanes
        int j = threadIdx.x % O, rl = threadIdx.x / O;
        if (rl < lanes) {
            float s = 0.f;
            for (int64_t r = r0 + rl; r < r1; r += lanes) s += G[(size_t)r * O + j];
            atomicAdd(&sums[j], s);
        }
    } else {
        for (int j = threadIdx.x; j < O; j += blockDim.x) {
            float s = 0.f;
            for (int64_t r = r0; r < r1; ++r) s += G[(size_t)r * O + j];
            sums[j] = s;
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < O; j += blockDim.x) atomicAdd(out + j, sums[j]);
}

}  // namespace rgcn
