// rgcn_forward / rgcn_backward: argument checking, workspace layout and kernel dispatch.
//
// Forward  (reference layers.py:286-306 / :518-556):  out[s] = bias + sum_e val_e * T_p(X[o])
// Backward (reference: autograd only; closed forms in SURVEY a10):
//   gX[o]  = sum_e val_e * T_p^T(G[s])        -- same walk on the source-major CSR, transposed weights
//   gW_p   = sum_{e in p} val_e X[o]^T G[s]   -- relation-major walk, projected onto the decomposition
//   gbias  = sum_s G[s]
#include "common.cuh"
#include "propagate_generic.cuh"
#include "propagate_fast.cuh"
#include "propagate_mma.cuh"
#include "propagate_fused.cuh"
#include "propagate_umma.cuh"

using namespace rgcn;

namespace {

int pow2_slots(int dim) {   // lane-strided register slots needed to cover `dim` values with 32 lanes
    int k = (dim + 31) / 32;
    int p = 1;
    while (p < k) p <<= 1;
    return p;
}

template <typename XT>
int launch_prop_generic(const PropArgs& A, const XT* X, cudaStream_t st) {
    const int maxdim = A.featureless ? A.O : (A.I > A.O ? A.I : A.O);
    const int K = pow2_slots(maxdim);
    RGCN_REQUIRE(K <= 16, RGCN_ERR_UNSUPPORTED, "generic propagate: feature width %d > 512 not supported yet", maxdim);
    const int wpb = 8, block = wpb * 32;
    const size_t smem = (A.featureless || A.form == RGCN_W_DIAG) ? 0 : (size_t)wpb * A.I * sizeof(float);
    int64_t want = (A.nrows + wpb - 1) / wpb;
    int grid = (int)(want < (int64_t)kNumSMs * 64 ? want : (int64_t)kNumSMs * 64);
    if (grid < 1) grid = 1;
    // pass 0: every row (hub rows only get their bias); pass 1: hub rows, split over the warps of 16 CTAs each
    const int long_bound = !A.long_list ? 0 : (A.num_long >= 0 ? (int)A.num_long : (int)(A.nnz_hint / RGCN_LONG_ROW));
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1 && long_bound == 0) break;
        PropArgs P = A;
        P.long_mode = pass;
        dim3 g = pass ? dim3(long_bound, 16) : dim3(grid);
        switch (K) {
            case 1: RGCN_LAUNCH((k_prop_generic<XT, 1>), g, block, smem, st, P, X); break;
            case 2: RGCN_LAUNCH((k_prop_generic<XT, 2>), g, block, smem, st, P, X); break;
            case 4: RGCN_LAUNCH((k_prop_generic<XT, 4>), g, block, smem, st, P, X); break;
            case 8: RGCN_LAUNCH((k_prop_generic<XT, 8>), g, block, smem, st, P, X); break;
            default: RGCN_LAUNCH((k_prop_generic<XT, 16>), g, block, smem, st, P, X); break;
        }
    }
    return RGCN_OK;
}

template <typename XT>
int launch_prop(const PropArgs& A, const XT* X, cudaStream_t st) { return launch_prop_generic(A, X, st); }

template <typename XT>
int launch_wgrad_generic(WGradArgs A, const XT* X, const float* G, int64_t nnz, int Rp, cudaStream_t st) {
    int nel = (A.form == RGCN_W_DIAG) ? A.I : (A.form == RGCN_W_BLOCK ? A.nb * A.bi * A.bo : A.I * A.O);
    if (A.form == RGCN_W_BLOCK && A.self_rel >= 0 && A.I * A.O > nel) nel = A.I * A.O;
    int KE = 1;
    while (KE < 32 && KE * 256 < nel) KE <<= 1;
    int tile = (int)(40960 / ((size_t)(A.I + A.O) * sizeof(float)));
    tile = tile < 1 ? 1 : (tile > 32 ? 32 : tile);
    A.tile = tile;
    const size_t smem = (size_t)tile * (A.I + A.O) * sizeof(float);
    RGCN_REQUIRE(smem <= 48 * 1024, RGCN_ERR_UNSUPPORTED, "generic weight-grad: I+O=%d too wide", A.I + A.O);
    // edges per CTA slice: small weights (few atomics per flush) get many short slices, large weights long ones
    int64_t slice = nel < 256 ? 256 : (nel > 4096 ? 4096 : nel);
    int64_t gy = nnz / ((int64_t)Rp * slice) + 1;
    if (gy > 1024) gy = 1024;
    dim3 grid(Rp, (unsigned)gy);
    switch (KE) {
        case 1: RGCN_LAUNCH((k_wgrad_generic<XT, 1>), grid, 256, smem, st, A, X, G); break;
        case 2: RGCN_LAUNCH((k_wgrad_generic<XT, 2>), grid, 256, smem, st, A, X, G); break;
        case 4: RGCN_LAUNCH((k_wgrad_generic<XT, 4>), grid, 256, smem, st, A, X, G); break;
        case 8: RGCN_LAUNCH((k_wgrad_generic<XT, 8>), grid, 256, smem, st, A, X, G); break;
        case 16: RGCN_LAUNCH((k_wgrad_generic<XT, 16>), grid, 256, smem, st, A, X, G); break;
        default: RGCN_LAUNCH((k_wgrad_generic<XT, 32>), grid, 256, smem, st, A, X, G); break;
    }
    return RGCN_OK;
}

// Dense weight gradient of mid / wide layers (the 200 x 200 basis layer of configs/rgcn/lp-WN18.yaml, 500-wide layers):
// gW[p] += (val X[src])^T G[dst] over the edges of relation p is a GEMM with gathered rows.  One CTA owns a 64 x 64 tile
// of gW[p] and a slice of the relation's edges: 16-edge slabs of the two gathered row tiles go through shared memory,
// each thread keeps a 4 x 4 register tile, slices are combined with one atomicAdd per element (split-K).
template <typename XT>
__global__ void __launch_bounds__(256) k_wgrad_dense_tiled(WGradArgs A, const XT* __restrict__ X, const float* __restrict__ G) {
    constexpr int TE = 16, T = 64;
    __shared__ __align__(16) float Xs[TE][T];
    __shared__ __align__(16) float Gs[TE][T];
    const int I = A.I, O = A.O;
    const int tiles_j = (O + T - 1) / T;
    const int ti = blockIdx.x / tiles_j, tj = blockIdx.x % tiles_j, p = blockIdx.y + A.rel0;
    const float* mask = (A.mask && p == A.mask_rel) ? A.mask : nullptr;
    const int e0 = A.relptr[p], e1 = A.relptr[p + 1];
    const int n = e1 - e0;
    if (n <= 0) return;
    // slices of at least 256 edges: short relations flush their tile once, the long (self-loop) one is spread wide
    const int per = max(256, ((n + (int)gridDim.z - 1) / (int)gridDim.z + TE - 1) / TE * TE);
    const int b0 = e0 + blockIdx.z * per, b1 = min(e1, b0 + per);
    if (b0 >= b1) return;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int le = tid >> 4, lc = (tid & 15) * 4;            // loader role: edge le of the slab, columns lc .. lc + 3
    float acc[4][4] = {};
    // loads of slab s + 1 are in flight while slab s is multiplied (software pipeline through registers)
    auto load_slab = [&](int s0, float4& xv, float4& gv) {
        const int e = s0 + le;
        xv = make_float4(0.f, 0.f, 0.f, 0.f);
        gv = xv;
        if (e < b1) {
            const float v = A.val[e];
            const XT* xr = X + (size_t)A.src[e] * I + ti * T + lc;
            const float* gr = G + (size_t)A.dst[e] * O + tj * T + lc;
            const int ci = ti * T + lc, cj = tj * T + lc;
            if (sizeof(XT) == 4 && (I & 3) == 0 && ci + 3 < I) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(xr));
                xv = make_float4(v * t.x, v * t.y, v * t.z, v * t.w);
            } else {
                if (ci + 0 < I) xv.x = v * to_f32(xr[0]);
                if (ci + 1 < I) xv.y = v * to_f32(xr[1]);
                if (ci + 2 < I) xv.z = v * to_f32(xr[2]);
                if (ci + 3 < I) xv.w = v * to_f32(xr[3]);
            }
            if ((O & 3) == 0 && cj + 3 < O) {
                gv = __ldg(reinterpret_cast<const float4*>(gr));
            } else {
                if (cj + 0 < O) gv.x = gr[0];
                if (cj + 1 < O) gv.y = gr[1];
                if (cj + 2 < O) gv.z = gr[2];
                if (cj + 3 < O) gv.w = gr[3];
            }
            if (mask) {
                const float* mr = mask + (size_t)A.dst[e] * O + tj * T + lc;
                if (cj + 0 < O) gv.x *= mr[0];
                if (cj + 1 < O) gv.y *= mr[1];
                if (cj + 2 < O) gv.z *= mr[2];
                if (cj + 3 < O) gv.w *= mr[3];
            }
        }
    };
    float4 xv, gv;
    load_slab(b0, xv, gv);
    for (int s0 = b0; s0 < b1; s0 += TE) {
        __syncthreads();                                     // the previous slab has been consumed
        *reinterpret_cast<float4*>(&Xs[le][lc]) = xv;
        *reinterpret_cast<float4*>(&Gs[le][lc]) = gv;
        __syncthreads();
        if (s0 + TE < b1) load_slab(s0 + TE, xv, gv);
#pragma unroll
        for (int k = 0; k < TE; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&Xs[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Gs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
    }
    float* gw = A.gW + (size_t)(p - A.rel0) * I * O;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = ti * T + ty * 4 + i;
        if (r >= I) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = tj * T + tx * 4 + j;
            if (c < O && acc[i][j] != 0.f) atomicAdd(gw + (size_t)r * O + c, acc[i][j]);
        }
    }
}

// Dense propagation of mid / wide layers (200 -> 200 basis layer of configs/rgcn/lp-WN18.yaml, 500-wide dense self-loop
// weights): out[scatter_e] += val_e * X[gather_e] W_p is, per relation, a GEMM whose A rows are gathered.  One CTA owns a
// chunk of <= RGCN_CHUNK_EDGES edges of one relation and a 64-column tile of the output; 64-edge sub-tiles go through
// the inner dimension in 16-wide slabs (gathered rows and the weight slab staged in shared memory, 4 x 4 register tile
// per thread), and each finished sub-tile is added to its destination rows with vector reductions.  Serves the forward
// (gather sources, W) and the feature gradient (gather destinations, W^T) — reference layers.py:293-301 / :534-548
// without the (R'N, I) temporary.  `out` holds the bias / zeros before the launch.
struct GemmArgs {
    const int32_t* relptr; const int32_t* chunkptr; int num_rels;
    const int32_t* gather; const int32_t* scatter; const float* val;
    const float* W;               // (R', I, O)
    int I, O;
    const float* out_mask;        // (N, O): multiplies the messages of relation mask_rel
    const float* in_mask;         // (N, I): multiplies the gathered rows of relation mask_rel
    int mask_rel;
    float* out;
    int only_rel_plus1;           // 0: all relations; else only relation only_rel_plus1 - 1, whose weight is W[0]
};

template <typename XT>
__global__ void __launch_bounds__(256) k_prop_dense_tiled(GemmArgs A, const XT* __restrict__ X) {
    constexpr int TK = 16, T = 64;
    __shared__ __align__(16) float Xs[TK][T + 4];
    __shared__ __align__(16) float Ws[TK][T];
    const int tiles_o = (A.O + 63) / 64;
    const int c = (int)(blockIdx.x / tiles_o);                // column tiles of a chunk are adjacent CTAs
    if (c >= A.chunkptr[A.num_rels]) return;
    int lo = 0, hi = A.num_rels;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (A.chunkptr[mid] <= c) lo = mid; else hi = mid;
    }
    const int p = lo;
    if (A.only_rel_plus1 && p + 1 != A.only_rel_plus1) return;
    const int e0 = A.relptr[p] + (c - A.chunkptr[p]) * RGCN_CHUNK_EDGES;
    const int e1 = min(A.relptr[p + 1], e0 + RGCN_CHUNK_EDGES);
    const int I = A.I, O = A.O, tj = (int)(blockIdx.x % tiles_o);
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int le = tid >> 2, lk = (tid & 3) * 4;             // X loader: edge le of the sub-tile, inner offsets lk .. lk + 3
    const int wk = tid >> 4, wc = (tid & 15) * 4;            // W loader: inner row wk of the slab, columns wc .. wc + 3
    const float* Wp = A.W + (A.only_rel_plus1 ? 0 : (size_t)p * I * O);
    const float* imask = (A.in_mask && p == A.mask_rel) ? A.in_mask : nullptr;
    const float* omask = (A.out_mask && p == A.mask_rel) ? A.out_mask : nullptr;
    const bool vec_w = (O & 3) == 0, vec_x = (I & 3) == 0 && sizeof(XT) == 4;
    for (int sub = e0 + T * (int)blockIdx.y; sub < e1; sub += T * (int)gridDim.y) {
        const int e = sub + le;
        const int64_t grow = e < e1 ? A.gather[e] : -1;
        float acc[4][4] = {};
        // loads of slab k0 + TK are in flight while slab k0 is multiplied (software pipeline through registers)
        auto load_slab = [&](int k0, float (&xv)[4], float4& wv) {
            xv[0] = xv[1] = xv[2] = xv[3] = 0.f;
            wv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (grow >= 0) {
                const int k = k0 + lk;
                const XT* xr = X + (size_t)grow * I + k;
                if (vec_x && k + 3 < I) {
                    const float4 t = __ldg(reinterpret_cast<const float4*>(xr));
                    xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) if (k + i < I) xv[i] = to_f32(xr[i]);
                }
                if (imask) {
                    const float* mr = imask + (size_t)grow * I + k;
#pragma unroll
                    for (int i = 0; i < 4; ++i) if (k + i < I) xv[i] *= mr[i];
                }
            }
            if (k0 + wk < I) {
                const int col = tj * T + wc;
                const float* wr = Wp + (size_t)(k0 + wk) * O + col;
                if (vec_w && col + 3 < O) wv = __ldg(reinterpret_cast<const float4*>(wr));
                else {
                    if (col + 0 < O) wv.x = wr[0];
                    if (col + 1 < O) wv.y = wr[1];
                    if (col + 2 < O) wv.z = wr[2];
                    if (col + 3 < O) wv.w = wr[3];
                }
            }
        };
        float xv[4];
        float4 wv;
        load_slab(0, xv, wv);
        for (int k0 = 0; k0 < I; k0 += TK) {
            __syncthreads();                                 // the previous slab has been consumed
#pragma unroll
            for (int i = 0; i < 4; ++i) Xs[lk + i][le] = xv[i];
            *reinterpret_cast<float4*>(&Ws[wk][wc]) = wv;
            __syncthreads();
            if (k0 + TK < I) load_slab(k0 + TK, xv, wv);
#pragma unroll
            for (int k = 0; k < TK; ++k) {
                const float4 a = *reinterpret_cast<const float4*>(&Xs[k][ty * 4]);
                const float4 b = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
                const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
        }
        const int col = tj * T + tx * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int ee = sub + ty * 4 + i;
            if (ee >= e1 || col >= O) continue;
            const float v = A.val[ee];
            const size_t base = (size_t)A.scatter[ee] * O + col;
            float r[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) r[j] = v * acc[i][j] * ((omask && col + j < O) ? omask[base + j] : 1.f);
            if (vec_w) {
                atomicAdd(reinterpret_cast<float4*>(A.out + base), make_float4(r[0], r[1], r[2], r[3]));
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (col + j < O) atomicAdd(A.out + base + j, r[j]);
            }
        }
    }
}

// out[row, :] = bias (or 0): the starting value of the reductions above
__global__ void k_init_rows(float* __restrict__ out, int64_t n, int O, const float* __restrict__ bias) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = bias ? bias[i % O] : 0.f;
}

// Stage a gathered row in shared memory, `scale` applied: the loads of a batch of 16 x 32 elements are issued together
// (a plain strided loop makes one round trip to L2 / DRAM per few elements).
template <typename T>
__device__ __forceinline__ void stage_row(float* __restrict__ dst, const T* __restrict__ src, int n, int lane, float scale) {
    for (int base = 0; base < n; base += 512) {
        float r[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int i = base + lane + 32 * k;
            r[k] = i < n ? to_f32(src[i]) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int i = base + lane + 32 * k;
            if (i < n) dst[i] = scale * r[k];
        }
    }
}

// Block-diagonal weights of any block size (the 100 blocks of 5 x 5 of configs/rgcn/lp-FB-toy.yaml, 10 of 5 x 5, ...) that
// the templated relation-batched kernels do not cover.  Relation-major and edge-parallel instead of one warp per
// destination row: a CTA owns a chunk of <= RGCN_CHUNK_EDGES edges of one relation, so the relation's blocks (a few KB)
// stay in L1; a warp takes one edge at a time, stages the gathered row in shared memory and adds
// val * x blockdiag(W_p) to the destination row (out holds the bias beforehand).  The self-loop relation of an LP layer
// (dense blocks_self) is skipped here and served by the tiled GEMM kernel.
struct BlockEdgeArgs {
    const int32_t* relptr; const int32_t* chunkptr; int num_rels;
    const int32_t* gather; const int32_t* scatter; const float* val;
    const float* blocks;          // (Rb, nb, bi, bo)
    int num_block_rels, nb, bi, bo, I, O;
    float* out;
};

template <typename XT>
__global__ void __launch_bounds__(256) k_block_edges(BlockEdgeArgs A, const XT* __restrict__ X) {
    extern __shared__ __align__(16) float be_smem[];
    const int c = blockIdx.x;
    if (c >= A.chunkptr[A.num_rels]) return;
    int lo = 0, hi = A.num_rels;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (A.chunkptr[mid] <= c) lo = mid; else hi = mid;
    }
    const int p = lo;
    if (p >= A.num_block_rels) return;
    const int e0 = A.relptr[p] + (c - A.chunkptr[p]) * RGCN_CHUNK_EDGES;
    const int e1 = min(A.relptr[p + 1], e0 + RGCN_CHUNK_EDGES);
    const int I = A.I, O = A.O, bi = A.bi, bo = A.bo;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int* wofs = reinterpret_cast<int*>(be_smem);              // per output column: offset of W[kb][0][jj] in the relation's blocks
    int* xofs = wofs + O;                                     // per output column: first input of its block
    float* xs = be_smem + 2 * O + (size_t)warp * I;
    for (int j = threadIdx.x; j < O; j += blockDim.x) {
        const int kb = j / bo, jj = j - kb * bo;
        wofs[j] = kb * bi * bo + jj;
        xofs[j] = kb * bi;
    }
    __syncthreads();
    const float* Wp = A.blocks + (size_t)p * A.nb * bi * bo;
    for (int e = e0 + warp; e < e1; e += 8) {
        const XT* xr = X + (size_t)A.gather[e] * I;
        stage_row(xs, xr, I, lane, 1.f);
        __syncwarp();
        const float v = A.val[e];
        float* orow = A.out + (size_t)A.scatter[e] * O;
        for (int j = lane; j < O; j += 32) {                  // (16-byte reductions of 4 outputs per lane measured slower:
            const float* w = Wp + wofs[j];                     //  strided weight reads and shared-memory bank conflicts)
            const float* x = xs + xofs[j];
            float sum = 0.f;
            for (int ii = 0; ii < bi; ++ii) sum = fmaf(x[ii], __ldg(w + ii * bo), sum);
            atomicAdd(orow + j, v * sum);
        }
        __syncwarp();
    }
}

// gblocks[p, kb, ii, jj] += sum_e val_e X[src_e, kb bi + ii] G[dst_e, kb bo + jj]: a CTA per relation chunk; a warp stages
// the two rows of an edge, lane l owns the elements l, l + 32, ... of the relation's blocks in registers across its
// edges and adds them to the gradient once per chunk.
template <typename XT, int KE>
__global__ void __launch_bounds__(256) k_block_wgrad(BlockEdgeArgs A, const XT* __restrict__ X, const float* __restrict__ G,
                                                     float* __restrict__ gblocks) {
    extern __shared__ __align__(16) float be_smem[];
    const int c = blockIdx.x;
    if (c >= A.chunkptr[A.num_rels]) return;
    int lo = 0, hi = A.num_rels;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (A.chunkptr[mid] <= c) lo = mid; else hi = mid;
    }
    const int p = lo;
    if (p >= A.num_block_rels) return;
    const int e0 = A.relptr[p] + (c - A.chunkptr[p]) * RGCN_CHUNK_EDGES;
    const int e1 = min(A.relptr[p + 1], e0 + RGCN_CHUNK_EDGES);
    const int I = A.I, O = A.O, bi = A.bi, bo = A.bo, nel = A.nb * bi * bo;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* xs = be_smem + (size_t)warp * (I + O);
    float* gs = xs + I;
    uint32_t xg[KE];                                          // input index | output index << 16 of the lane's elements
    float acc[KE];
    const int el0 = blockIdx.y * 32 * KE;                     // this CTA's part of the relation's block elements
#pragma unroll
    for (int k = 0; k < KE; ++k) {
        const int el = el0 + lane + 32 * k;
        acc[k] = 0.f;
        xg[k] = 0;
        if (el < nel) {
            const int kb = el / (bi * bo), rem = el - kb * bi * bo, ii = rem / bo;
            xg[k] = (uint32_t)(kb * bi + ii) | ((uint32_t)(kb * bo + (rem - ii * bo)) << 16);
        }
    }
    for (int e = e0 + warp; e < e1; e += 8) {
        const float v = A.val[e];
        const XT* xr = X + (size_t)A.gather[e] * I;
        const float* gr = G + (size_t)A.scatter[e] * O;
        stage_row(xs, xr, I, lane, v);
        stage_row(gs, gr, O, lane, 1.f);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < KE; ++k) acc[k] = fmaf(xs[xg[k] & 0xffffu], gs[xg[k] >> 16], acc[k]);
        __syncwarp();
    }
    float* dest = gblocks + (size_t)p * nel;
#pragma unroll
    for (int k = 0; k < KE; ++k) {
        const int el = el0 + lane + 32 * k;
        if (el < nel && acc[k] != 0.f) atomicAdd(dest + el, acc[k]);
    }
}

// Small square blocks with a compile-time size (2 x 2, 4 x 4, 5 x 5 — what `num_blocks` = width / 5 of the shipped LP
// configs gives): lane = block.  The relation's blocks and the staged row sit in shared memory (lane strides of B*B and B
// floats: conflict-free for odd B), the B*B products of a block are independent FMAs in registers, and the weight
// gradient keeps a lane's blocks (KB = nb / 32 of them) in registers across the chunk.
template <typename XT, int B>
__global__ void __launch_bounds__(256) k_block_edges_sq(BlockEdgeArgs A, const XT* __restrict__ X) {
    extern __shared__ __align__(16) float be_smem[];
    const int c = blockIdx.x;
    if (c >= A.chunkptr[A.num_rels]) return;
    int lo = 0, hi = A.num_rels;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (A.chunkptr[mid] <= c) lo = mid; else hi = mid;
    }
    const int p = lo;
    if (p >= A.num_block_rels) return;
    const int e0 = A.relptr[p] + (c - A.chunkptr[p]) * RGCN_CHUNK_EDGES;
    const int e1 = min(A.relptr[p + 1], e0 + RGCN_CHUNK_EDGES);
    const int I = A.I, O = A.O, nb = A.nb;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* ws = be_smem;                                      // the relation's blocks: nb * B * B floats
    float* xs = ws + (nb * B * B + 3) / 4 * 4 + (size_t)warp * (I + O);
    float* os = xs + I;                                       // the edge's message, turned into coalesced 16-byte reductions
    const float* Wp = A.blocks + (size_t)p * nb * B * B;
    for (int i = threadIdx.x; i < nb * B * B; i += blockDim.x) ws[i] = Wp[i];
    __syncthreads();
    for (int e = e0 + warp; e < e1; e += 8) {
        const XT* xr = X + (size_t)A.gather[e] * I;
        stage_row(xs, xr, I, lane, 1.f);
        __syncwarp();
        const float v = A.val[e];
        float* orow = A.out + (size_t)A.scatter[e] * O;
        for (int kb = lane; kb < nb; kb += 32) {
            const float* w = ws + kb * B * B;
            const float* x = xs + kb * B;
            float acc[B];
#pragma unroll
            for (int jj = 0; jj < B; ++jj) acc[jj] = 0.f;
#pragma unroll
            for (int ii = 0; ii < B; ++ii) {
                const float xv = x[ii];
#pragma unroll
                for (int jj = 0; jj < B; ++jj) acc[jj] = fmaf(xv, w[ii * B + jj], acc[jj]);
            }
#pragma unroll
            for (int jj = 0; jj < B; ++jj) os[kb * B + jj] = v * acc[jj];
        }
        __syncwarp();
        if (((O | I) & 3) == 0) {
            for (int j = 4 * lane; j < O; j += 128)
                atomicAdd(reinterpret_cast<float4*>(orow + j), *reinterpret_cast<const float4*>(os + j));
        } else {
            for (int j = lane; j < O; j += 32) atomicAdd(orow + j, os[j]);
        }
        __syncwarp();
    }
}

template <typename XT, int B, int KB>
__global__ void __launch_bounds__(256) k_block_wgrad_sq(BlockEdgeArgs A, const XT* __restrict__ X, const float* __restrict__ G,
                                                        float* __restrict__ gblocks) {
    extern __shared__ __align__(16) float be_smem[];
    const int c = blockIdx.x;
    if (c >= A.chunkptr[A.num_rels]) return;
    int lo = 0, hi = A.num_rels;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (A.chunkptr[mid] <= c) lo = mid; else hi = mid;
    }
    const int p = lo;
    if (p >= A.num_block_rels) return;
    const int e0 = A.relptr[p] + (c - A.chunkptr[p]) * RGCN_CHUNK_EDGES;
    const int e1 = min(A.relptr[p + 1], e0 + RGCN_CHUNK_EDGES);
    const int I = A.I, O = A.O, nb = A.nb;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* xs = be_smem + (size_t)warp * (I + O);
    float* gs = xs + I;
    float acc[KB][B * B];
#pragma unroll
    for (int k = 0; k < KB; ++k)
#pragma unroll
        for (int i = 0; i < B * B; ++i) acc[k][i] = 0.f;
    for (int e = e0 + warp; e < e1; e += 8) {
        const float v = A.val[e];
        const XT* xr = X + (size_t)A.gather[e] * I;
        const float* gr = G + (size_t)A.scatter[e] * O;
        stage_row(xs, xr, I, lane, v);
        stage_row(gs, gr, O, lane, 1.f);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            const int kb = lane + 32 * k;
            if (kb < nb) {
                float gv[B];
#pragma unroll
                for (int jj = 0; jj < B; ++jj) gv[jj] = gs[kb * B + jj];
#pragma unroll
                for (int ii = 0; ii < B; ++ii) {
                    const float xv = xs[kb * B + ii];
#pragma unroll
                    for (int jj = 0; jj < B; ++jj) acc[k][ii * B + jj] = fmaf(xv, gv[jj], acc[k][ii * B + jj]);
                }
            }
        }
        __syncwarp();
    }
    float* dest = gblocks + (size_t)p * nb * B * B;
#pragma unroll
    for (int k = 0; k < KB; ++k) {
        const int kb = lane + 32 * k;
        if (kb < nb) {
#pragma unroll
            for (int i = 0; i < B * B; ++i)
                if (acc[k][i] != 0.f) atomicAdd(dest + kb * B * B + i, acc[k][i]);
        }
    }
}

inline bool block_sq_shape(int nb, int bi, int bo, int I, int O) {
    const char* e = getenv("RGCN_BLOCK_SQ");
    if (e && e[0] == '0') return false;
    return bi == bo && (bi == 2 || bi == 4 || bi == 5) && nb <= 128 &&
           (size_t)(nb * bi * bo + 8 * (I + O)) * 4 <= 96 * 1024;
}

template <typename XT, int B>
int launch_block_edges_sq(const BlockEdgeArgs& A, const XT* X, int chunks, cudaStream_t st) {
    const size_t smem = (size_t)((A.nb * B * B + 3) / 4 * 4 + 8 * (A.I + A.O)) * sizeof(float);
    RGCN_CHECK_CUDA(cudaFuncSetAttribute((k_block_edges_sq<XT, B>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RGCN_LAUNCH((k_block_edges_sq<XT, B>), chunks, 256, smem, st, A, X);
    return RGCN_OK;
}

template <typename XT, int B>
int launch_block_wgrad_sq(const BlockEdgeArgs& A, const XT* X, const float* G, float* gblocks, int chunks, cudaStream_t st) {
    const size_t smem = (size_t)8 * (A.I + A.O) * sizeof(float);
    auto go = [&](auto kernel) -> int {
        RGCN_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RGCN_LAUNCH(kernel, chunks, 256, smem, st, A, X, G, gblocks);
        return RGCN_OK;
    };
    if (A.nb <= 32) return go(k_block_wgrad_sq<XT, B, 1>);
    if (A.nb <= 64) return go(k_block_wgrad_sq<XT, B, 2>);
    return go(k_block_wgrad_sq<XT, B, 4>);
}

// shapes for the two kernels above: featured block layers the templated kernels do not take, blocks of one relation
// small enough for the per-lane register tile of the weight gradient
bool block_edges_shape(int nb, int bi, int bo, int I, int O) {
    const char* e = getenv("RGCN_BLOCK_EDGES");
    if (e && e[0] == '0') return false;
    return nb * bi * bo <= 32 * 80 * 4 && I >= 32 && O >= 32 && I < 65536 && O < 65536 &&
           (size_t)(2 * O + 8 * (I + O)) * 4 <= 96 * 1024;
}

template <typename XT>
int launch_block_edges(const BlockEdgeArgs& A, const XT* X, int chunks, cudaStream_t st) {
    if (block_sq_shape(A.nb, A.bi, A.bo, A.I, A.O)) {
        if (A.bi == 2) return launch_block_edges_sq<XT, 2>(A, X, chunks, st);
        if (A.bi == 4) return launch_block_edges_sq<XT, 4>(A, X, chunks, st);
        return launch_block_edges_sq<XT, 5>(A, X, chunks, st);
    }
    const size_t smem = (size_t)(2 * A.O + 8 * A.I) * sizeof(float);
    RGCN_CHECK_CUDA(cudaFuncSetAttribute(k_block_edges<XT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RGCN_LAUNCH((k_block_edges<XT>), chunks, 256, smem, st, A, X);
    return RGCN_OK;
}

template <typename XT>
int launch_block_wgrad(const BlockEdgeArgs& A, const XT* X, const float* G, float* gblocks, int chunks, cudaStream_t st) {
    if (block_sq_shape(A.nb, A.bi, A.bo, A.I, A.O)) {
        if (A.bi == 2) return launch_block_wgrad_sq<XT, 2>(A, X, G, gblocks, chunks, st);
        if (A.bi == 4) return launch_block_wgrad_sq<XT, 4>(A, X, G, gblocks, chunks, st);
        return launch_block_wgrad_sq<XT, 5>(A, X, G, gblocks, chunks, st);
    }
    const size_t smem = (size_t)8 * (A.I + A.O) * sizeof(float);
    const int nel = A.nb * A.bi * A.bo;
    auto go = [&](auto kernel, int ke) -> int {
        RGCN_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((unsigned)chunks, (unsigned)((nel + 32 * ke - 1) / (32 * ke)));   // y: parts of 32 * KE block elements
        RGCN_LAUNCH(kernel, grid, 256, smem, st, A, X, G, gblocks);
        return RGCN_OK;
    };
    // one part per relation chunk whenever the elements fit the register tile (splitting 2,500 elements over three
    // CTAs of 32 per lane measured slower than one CTA of 80 per lane: every part re-stages the same rows)
    if (nel <= 32 * 8) return go(k_block_wgrad<XT, 8>, 8);
    const char* ke = getenv("RGCN_BLOCK_WGRAD_KE");
    if (nel <= 32 * 32 || (ke && atoi(ke) == 32)) return go(k_block_wgrad<XT, 32>, 32);
    return go(k_block_wgrad<XT, 80>, 80);
}

// RGCN_SPLIT_SELF=0 keeps the dense self-loop weight of LP block layers on the generic kernels (A/B measurements)
bool split_self_enabled() {
    const char* e = getenv("RGCN_SPLIT_SELF");
    return !(e && e[0] == '0');
}

bool dense_tiled_shape(int form, int featureless, int I, int O, int64_t nnz) {
    return form == RGCN_W_DENSE && !featureless && (int64_t)I * O >= 1024 && nnz > 0;
}

template <typename XT>
int launch_prop_dense_tiled(GemmArgs A, const XT* X, int64_t N, const float* bias, int chunks, cudaStream_t st,
                            bool init = true, int64_t nnz_hint = 1 << 30) {
    const int64_t n = N * (int64_t)A.O;
    if (init) RGCN_LAUNCH(k_init_rows, grid_for(n, 256), 256, 0, st, A.out, n, A.O, bias);
    // small graphs: the 64-edge sub-tiles of a chunk are dealt over up to 16 CTAs so that the grid fills the GPU
    const int tiles = (A.O + 63) / 64;
    const int64_t edges = A.only_rel_plus1 ? N : nnz_hint;
    int64_t z = (4 * kNumSMs) / (tiles * (edges / RGCN_CHUNK_EDGES + 1)) + 1;
    z = z > 16 ? 16 : z;
    RGCN_REQUIRE((int64_t)tiles * chunks < (1ll << 31), RGCN_ERR_UNSUPPORTED, "dense tiled propagation: grid too large");
    dim3 grid((unsigned)((int64_t)tiles * chunks), (unsigned)z);
    RGCN_LAUNCH((k_prop_dense_tiled<XT>), grid, 256, 0, st, A, X);
    return RGCN_OK;
}

template <typename XT>
int launch_wgrad(const WGradArgs& A, const XT* X, const float* G, int64_t nnz, int Rp, cudaStream_t st) {
    if (A.form == RGCN_W_DENSE && (int64_t)A.I * A.O >= 1024 && Rp - A.rel0 <= 65535) {
        Rp -= A.rel0;                                  // relations rel0 .. R' - 1 (the self relation alone, or all)
        const int tiles = ((A.I + 63) / 64) * ((A.O + 63) / 64);
        // up to 64 slices per relation, sized for the longest one a graph of nnz edges can hold (the self-loop
        // relation of an LP step has more edges than all others together); CTAs of slices past a relation's end exit
        int64_t z = nnz / 512 + 1;
        if (z > 64) z = 64;
        while (z > 1 && (int64_t)tiles * Rp * z > 65535LL * 4) --z;
        dim3 grid((unsigned)tiles, (unsigned)Rp, (unsigned)z);
        RGCN_LAUNCH((k_wgrad_dense_tiled<XT>), grid, 256, 0, st, A, X, G);
        return RGCN_OK;
    }
    return launch_wgrad_generic(A, X, G, nnz, Rp, st);
}

int launch_featureless_grad(FeaturelessGradArgs A, const float* G, cudaStream_t st) {
    const int K = pow2_slots(A.O);
    RGCN_REQUIRE(K <= 16, RGCN_ERR_UNSUPPORTED, "featureless weight-grad: out_dim %d > 512 not supported yet", A.O);
    int wpb = 8;
    size_t per_warp = (A.form == RGCN_W_BASIS) ? (size_t)A.B * A.O * sizeof(float) : 0;
    while (wpb > 1 && per_warp * wpb > 32 * 1024) wpb >>= 1;
    RGCN_REQUIRE(per_warp * wpb <= 40 * 1024, RGCN_ERR_UNSUPPORTED, "featureless basis grad: B*O=%d too large", A.B * A.O);
    size_t comps_bytes = (A.form == RGCN_W_BASIS) ? (size_t)A.num_rels * A.B * sizeof(float) : 0;
    A.comps_in_smem = (comps_bytes > 0 && per_warp * wpb + comps_bytes <= 48 * 1024) ? 1 : 0;
    const size_t smem = per_warp * wpb + (A.comps_in_smem ? comps_bytes : 0);
    int64_t want = (A.N + wpb - 1) / wpb;
    int grid = (int)(want < (int64_t)kNumSMs * 8 ? want : (int64_t)kNumSMs * 8);
    if (grid < 1) grid = 1;
    const int block = wpb * 32;
    switch (K) {
        case 1: RGCN_LAUNCH((k_wgrad_featureless<1>), grid, block, smem, st, A, G); break;
        case 2: RGCN_LAUNCH((k_wgrad_featureless<2>), grid, block, smem, st, A, G); break;
        case 4: RGCN_LAUNCH((k_wgrad_featureless<4>), grid, block, smem, st, A, G); break;
        case 8: RGCN_LAUNCH((k_wgrad_featureless<8>), grid, block, smem, st, A, G); break;
        default: RGCN_LAUNCH((k_wgrad_featureless<16>), grid, block, smem, st, A, G); break;
    }
    return RGCN_OK;
}

int launch_transpose(const float* in, int64_t n, int rows, int cols, float* out, cudaStream_t st) {
    int64_t total = n * rows * cols;
    if (total == 0) return RGCN_OK;
    RGCN_LAUNCH(k_transpose_batched, grid_for(total, 256), 256, 0, st, in, n, rows, cols, out);
    return RGCN_OK;
}

struct Shape {
    int64_t N, Rp, nnz;
    int I, O, B, nb, bi, bo, Rb;
    size_t w_elems;          // elements of the effective dense (R', I, O) weight (featured forms)
    size_t blocks_elems;
};

int check_common(const rgcn_graph* g, const rgcn_params* p, Shape* s, const char* who) {
    RGCN_REQUIRE(g && p, RGCN_ERR_ARG, "%s: NULL graph or params", who);
    s->N = g->num_nodes; s->Rp = g->num_rels; s->nnz = g->nnz;
    RGCN_REQUIRE(p->out_dim > 0 && p->in_dim > 0, RGCN_ERR_ARG, "%s: bad dims in=%lld out=%lld", who,
                 (long long)p->in_dim, (long long)p->out_dim);
    RGCN_REQUIRE(p->out_dim < (1 << 30) && (p->featureless || p->in_dim < (1 << 30)), RGCN_ERR_ARG, "%s: dims too large", who);
    s->I = (int)p->in_dim; s->O = (int)p->out_dim;
    s->B = (int)p->num_bases; s->nb = (int)p->num_blocks; s->bi = s->bo = 0; s->Rb = (int)p->num_block_rels;
    if (p->featureless) {
        RGCN_REQUIRE(p->in_dim == g->num_nodes, RGCN_ERR_ARG, "%s: featureless layer needs in_dim == num_nodes", who);
        RGCN_REQUIRE(p->form != RGCN_W_DIAG, RGCN_ERR_UNSUPPORTED, "%s: featureless diagonal weights are not defined", who);
        RGCN_REQUIRE(!p->blocks_self && !p->self_mask, RGCN_ERR_UNSUPPORTED,
                     "%s: featureless layers with blocks_self / self_mask are unreachable in the reference", who);
    }
    switch (p->form) {
        case RGCN_W_DENSE: RGCN_REQUIRE(p->weights, RGCN_ERR_ARG, "%s: weights is NULL", who); break;
        case RGCN_W_DIAG:
            RGCN_REQUIRE(p->weights, RGCN_ERR_ARG, "%s: weights is NULL", who);
            RGCN_REQUIRE(p->in_dim == p->out_dim, RGCN_ERR_ARG, "%s: diagonal weights need in_dim == out_dim", who);
            break;
        case RGCN_W_BASIS:
            RGCN_REQUIRE(p->bases && p->comps && p->num_bases > 0, RGCN_ERR_ARG, "%s: bases/comps missing", who);
            break;
        case RGCN_W_BLOCK:
            RGCN_REQUIRE(p->blocks && p->num_blocks > 0, RGCN_ERR_ARG, "%s: blocks missing", who);
            RGCN_REQUIRE(p->in_dim % p->num_blocks == 0 && p->out_dim % p->num_blocks == 0, RGCN_ERR_ARG,
                         "%s: dims (%lld, %lld) not divisible by num_blocks %lld", who, (long long)p->in_dim,
                         (long long)p->out_dim, (long long)p->num_blocks);
            s->bi = (int)(p->in_dim / p->num_blocks); s->bo = (int)(p->out_dim / p->num_blocks);
            RGCN_REQUIRE(p->num_block_rels == s->Rp - (p->blocks_self ? 1 : 0), RGCN_ERR_ARG,
                         "%s: num_block_rels %lld inconsistent with R'=%lld", who, (long long)p->num_block_rels,
                         (long long)s->Rp);
            break;
        default: RGCN_REQUIRE(false, RGCN_ERR_ARG, "%s: unknown weight form %d", who, p->form);
    }
    s->w_elems = p->featureless ? 0 : (size_t)s->Rp * s->I * s->O;
    s->blocks_elems = (p->form == RGCN_W_BLOCK) ? (size_t)s->Rb * s->nb * s->bi * s->bo : 0;
    return RGCN_OK;
}

void fill_weights(PropArgs& A, const rgcn_params* p, const Shape& s) {
    A.form = p->form; A.featureless = p->featureless;
    A.I = s.I; A.O = s.O; A.B = s.B; A.nb = s.nb; A.bi = s.bi; A.bo = s.bo;
    A.self_rel = p->blocks_self ? (int)s.Rp - 1 : -1;
    A.W = p->weights; A.bases = p->bases; A.comps = p->comps; A.blocks = p->blocks; A.blocks_self = p->blocks_self;
    A.bias = nullptr; A.out_mask = nullptr; A.in_mask = nullptr; A.mask_rel = (int)s.Rp - 1;
}

// Relation-batched path (propagate_fast.cuh): featured, dense or pure block-diagonal weights with an
// instantiated block shape, no dropout mask, and a message buffer that stays within kMaxMsgBytes.
constexpr size_t kMaxMsgBytes = (size_t)24 << 30;

bool rel_path(const rgcn_params* p, const Shape& s, bool bf16, bool transposed, RelShape* rs, size_t* msg_bytes) {
    if (p->featureless || p->self_mask || p->blocks_self || s.nnz == 0) return false;
    if (p->form == RGCN_W_DIAG) return false;
    int nb = 1, bi = s.I, bo = s.O;
    if (p->form == RGCN_W_BLOCK) { nb = s.nb; bi = s.bi; bo = s.bo; }
    if (transposed) { int t = bi; bi = bo; bo = t; }
    if (!rel_shape_supported(nb, bi, bo, bf16)) return false;
    const int ti = bi < 8 ? bi : 8, tj = bo < 8 ? bo : 8;
    if (nb * (bi / ti) * (bo / tj) > 256) return false;
    size_t bytes = (size_t)s.nnz * nb * bo * (bf16 ? 2 : 4);
    if (bytes > kMaxMsgBytes) return false;
    rs->nb = nb; rs->bi = bi; rs->bo = bo;
    *msg_bytes = align_up(bytes);
    return true;
}

// Tiled tensor-core path: bf16 features, 16x16 blocks, and a plan built with tile_edges > 0.
bool tiled_path(const rgcn_graph* g, const rgcn_params* p, const Shape& s, bool bf16) {
    return bf16 && g->tile_edges > 0 && g->num_tiles > 0 && g->tile_capacity > 0 && !p->featureless &&
           p->form == RGCN_W_BLOCK && !p->blocks_self && !p->self_mask && s.nnz > 0 &&
           mma_shape_supported(s.nb, s.bi, s.bo);
}

// Fused row-block path (propagate_fused.cuh): bf16 features, four 16x16 blocks (64 -> 64), and a plan whose fused
// list for this direction was built and fits (fuse_items > 0).
bool fused_path(const rgcn_graph* g, const rgcn_params* p, const Shape& s, bool bf16, bool backward) {
    return bf16 && g->fuse_rows > 0 && g->fuse_items[backward ? 1 : 0] > 0 && !p->featureless &&
           p->form == RGCN_W_BLOCK && !p->blocks_self && !p->self_mask && s.nnz > 0 && s.nb >= 4 && s.nb % 4 == 0 &&
           s.bi == 16 && s.bo == 16;
}

// Tensor-core gathered GEMM (propagate_umma.cuh): bf16 features, dense or basis weights of at least 64 x 64, no mask.
bool umma_path(const rgcn_params* p, const Shape& s, bool bf16) {
    return bf16 && !p->featureless && !p->self_mask && (p->form == RGCN_W_DENSE || p->form == RGCN_W_BASIS) && s.nnz > 0 &&
           umma_shape_supported(s.I, s.O) && umma_shape_supported(s.O, s.I);
}

size_t tiled_ring_bytes(const rgcn_graph* g, int width) {
    return align_up((size_t)g->ring_depth * (size_t)g->tile_capacity * width * 2);
}

int max_chunks(const Shape& s) { return (int)(s.nnz / RGCN_CHUNK_EDGES + s.Rp); }

}  // namespace

namespace {
__global__ void k_widen_bf16(const uint4* __restrict__ in, int64_t n8, float4* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        uint4 v;
        asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(in + i));
        const float4 a = make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u),
                                     __uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u));
        const float4 b = make_float4(__uint_as_float(v.z << 16), __uint_as_float(v.z & 0xffff0000u),
                                     __uint_as_float(v.w << 16), __uint_as_float(v.w & 0xffff0000u));
        __stcs(out + 2 * i, a);
        __stcs(out + 2 * i + 1, b);
    }
}
}  // namespace

extern "C" int rgcn_widen_rows(const void* in, int64_t n, float* out, rgcn_stream_t stream) {
    RGCN_REQUIRE(n >= 0 && n % 8 == 0 && (n == 0 || (in && out)), RGCN_ERR_ARG, "rgcn_widen_rows: n must be a multiple of 8");
    RGCN_REQUIRE(((uintptr_t)in | (uintptr_t)out) % 16 == 0, RGCN_ERR_ARG, "rgcn_widen_rows: pointers must be 16-byte aligned");
    if (n == 0) return RGCN_OK;
    const int64_t n8 = n / 8;
    const int grid = (int)(n8 / 256 + 1 < (int64_t)kNumSMs * 16 ? n8 / 256 + 1 : (int64_t)kNumSMs * 16);
    RGCN_LAUNCH(k_widen_bf16, grid, 256, 0, (cudaStream_t)stream, static_cast<const uint4*>(in), n8, reinterpret_cast<float4*>(out));
    return RGCN_OK;
}

extern "C" size_t rgcn_forward_workspace_bytes(const rgcn_graph* g, const rgcn_params* p, int x_dtype) {
    Shape s;
    if (check_common(g, p, &s, "rgcn_forward_workspace_bytes")) return 0;
    size_t bytes = 0;
    if (p->form == RGCN_W_BASIS && !p->featureless) bytes += align_up(s.w_elems * sizeof(float));
    if (umma_path(p, s, x_dtype == RGCN_BF16)) return bytes + umma_wt_bytes(s.Rp, s.I, s.O);
    if (fused_path(g, p, s, x_dtype == RGCN_BF16, false)) return bytes + fused_ws_bytes(s.Rp, s.nb);
    if (tiled_path(g, p, s, x_dtype == RGCN_BF16))
        return bytes + tiled_counter_bytes(g->num_tiles) + tiled_ring_bytes(g, s.O) + wfrag_bytes(s.Rp, s.nb);
    RelShape rs; size_t msg = 0;
    if (rel_path(p, s, x_dtype == RGCN_BF16, false, &rs, &msg)) bytes += msg;
    return bytes;
}

extern "C" int rgcn_forward(const rgcn_graph* g, const rgcn_params* p, const void* X, int x_dtype, float* out,
                            void* ws, size_t ws_bytes, rgcn_stream_t stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    Shape s;
    int rc = check_common(g, p, &s, "rgcn_forward");
    if (rc) return rc;
    RGCN_REQUIRE(out, RGCN_ERR_ARG, "rgcn_forward: out is NULL");
    RGCN_REQUIRE(p->featureless || X, RGCN_ERR_ARG, "rgcn_forward: features is NULL");
    RGCN_REQUIRE(x_dtype == RGCN_F32 || x_dtype == RGCN_BF16, RGCN_ERR_ARG, "rgcn_forward: unknown dtype %d", x_dtype);
    size_t need = rgcn_forward_workspace_bytes(g, p, x_dtype);
    RGCN_REQUIRE(ws_bytes >= need && (ws || need == 0), RGCN_ERR_WORKSPACE, "rgcn_forward: workspace %zu < %zu", ws_bytes, need);
    Carver carve(ws);

    PropArgs A{};
    A.rowptr = g->d_rowptr; A.col = g->d_src; A.rel = g->d_rel; A.val = g->d_val;
    A.nrows = s.N; A.N = s.N;
    fill_weights(A, p, s);
    A.bias = p->bias; A.out_mask = p->self_mask; A.out = out;
    A.long_list = g->d_long; A.long_count = g->status + 4; A.nnz_hint = s.nnz; A.num_long = g->num_long_dst;
    if (p->form == RGCN_W_BASIS && !p->featureless) {      // materialise the small (R', I, O) table only
        float* weff = carve.take<float>(s.w_elems);
        int64_t IO = (int64_t)s.I * s.O;
        dim3 grid(grid_for(IO, 256), (unsigned)s.Rp);
        RGCN_LAUNCH(k_basis_combine, grid, 256, 0, st, p->comps, p->bases, (int)s.Rp, s.B, IO, weff);
        A.form = RGCN_W_DENSE; A.W = weff;
    }
    const bool bf16 = x_dtype == RGCN_BF16 && !p->featureless;
    RGCN_REQUIRE(p->out_dtype == RGCN_F32 || p->out_dtype == RGCN_BF16, RGCN_ERR_ARG, "rgcn_forward: unknown out_dtype %d",
                 p->out_dtype);
    const bool ranged = p->row_lo != 0 || p->row_hi != 0;
    if (fused_path(g, p, s, bf16, false)) {
        void* fws = carve.take<char>(fused_ws_bytes(s.Rp, s.nb));
        if (ranged)
            RGCN_REQUIRE(p->row_lo >= 0 && p->row_lo < p->row_hi && p->row_lo % g->fuse_rows == 0 &&
                             (p->row_hi % g->fuse_rows == 0 || p->row_hi >= s.N), RGCN_ERR_ARG,
                         "rgcn_forward: row range [%lld, %lld) must be cut at multiples of fuse_rows %lld",
                         (long long)p->row_lo, (long long)p->row_hi, (long long)g->fuse_rows);
        const int64_t lo = ranged ? p->row_lo : 0, hi = ranged ? p->row_hi : s.N;
        RGCN_REQUIRE(p->num_peer_out == 0 || p->out_dtype == RGCN_BF16, RGCN_ERR_ARG, "rgcn_forward: peer_out needs a bf16 output");
        if (p->out_dtype == RGCN_BF16)
            return launch_fused_rows<__nv_bfloat16>(g, false, p->blocks, p->bias, static_cast<const __nv_bfloat16*>(X),
                                                    reinterpret_cast<__nv_bfloat16*>(out), fws, st, lo, hi, p->peer_out,
                                                    p->num_peer_out, s.nb);
        return launch_fused_rows<float>(g, false, p->blocks, p->bias, static_cast<const __nv_bfloat16*>(X), out, fws, st, lo, hi,
                                        nullptr, 0, s.nb);
    }
    RGCN_REQUIRE(p->out_dtype == RGCN_F32 && !ranged && p->num_peer_out == 0, RGCN_ERR_UNSUPPORTED,
                 "rgcn_forward: a bf16 output / an output row range / peer stores need the fused row-block path");
    if (tiled_path(g, p, s, bf16)) {
        int32_t* counters = reinterpret_cast<int32_t*>(carve.take<char>(tiled_counter_bytes(g->num_tiles)));
        __nv_bfloat16* ring = reinterpret_cast<__nv_bfloat16*>(carve.take<char>(tiled_ring_bytes(g, s.O)));
        uint32_t* frag = reinterpret_cast<uint32_t*>(carve.take<char>(wfrag_bytes(s.Rp, s.nb)));
        RGCN_CHECK_CUDA(cudaMemsetAsync(counters, 0, tiled_counter_bytes(g->num_tiles), st));
        rc = launch_pack_wfrag(p->blocks, s.Rp, s.nb, false, frag, st);
        if (rc) return rc;
        TiledArgs T = make_tiled_args(g, false, s.nb, frag, p->bias, counters);
        return launch_tiled_span(T, static_cast<const __nv_bfloat16*>(X), ring, out, st);
    }
    RelShape rs; size_t msg_bytes = 0;
    if (rel_path(p, s, bf16, false, &rs, &msg_bytes)) {
        RelArgs R{g->r_relptr, g->r_chunkptr, (int)s.Rp, g->r_src, nullptr, g->r_dslot, g->r_val,
                  A.form == RGCN_W_BLOCK ? A.blocks : A.W, rs.nb};
        void* msg = carve.take<char>(msg_bytes);
        if (bf16) {
            if (mma_shape_supported(rs.nb, rs.bi, rs.bo))
                rc = launch_rel_mma_fwd(R, static_cast<const __nv_bfloat16*>(X), static_cast<__nv_bfloat16*>(msg),
                                        max_chunks(s), st);
            else
                rc = launch_rel_transform(R, rs.bi, rs.bo, static_cast<const __nv_bfloat16*>(X),
                                          static_cast<__nv_bfloat16*>(msg), max_chunks(s), st);
            if (rc) return rc;
            return launch_row_sum(g->d_rowptr, s.N, s.O, static_cast<const __nv_bfloat16*>(msg), p->bias, out, g->d_long,
                                  g->status + 4, g->num_long_dst, s.nnz, st);
        }
        rc = launch_rel_transform(R, rs.bi, rs.bo, static_cast<const float*>(X), static_cast<float*>(msg),
                                  max_chunks(s), st);
        if (rc) return rc;
        return launch_row_sum(g->d_rowptr, s.N, s.O, static_cast<const float*>(msg), p->bias, out, g->d_long, g->status + 4,
                              g->num_long_dst, s.nnz, st);
    }
    if (umma_path(p, s, bf16)) {
        __nv_bfloat16* wt = reinterpret_cast<__nv_bfloat16*>(carve.take<char>(umma_wt_bytes(s.Rp, s.I, s.O)));
        const int64_t cnt = (int64_t)s.Rp * s.I * s.O;
        RGCN_LAUNCH(k_pack_wt_bf16, grid_for(cnt, 256), 256, 0, st, A.W, cnt, s.I, s.O, 1, wt);
        const int64_t n = s.N * (int64_t)s.O;
        RGCN_LAUNCH(k_init_rows, grid_for(n, 256), 256, 0, st, out, n, s.O, p->bias);
        UmmaArgs U{g->r_relptr, g->r_chunkptr, (int)s.Rp, g->r_src, g->r_dst, g->r_val, s.I, s.O, 0, out};
        return launch_gemm_umma(U, static_cast<const __nv_bfloat16*>(X), s.N, wt, max_chunks(s), st, g->max_rel_edges);
    }
    if (dense_tiled_shape(A.form, p->featureless, s.I, s.O, s.nnz)) {
        GemmArgs Gm{g->r_relptr, g->r_chunkptr, (int)s.Rp, g->r_src, g->r_dst, g->r_val, A.W, s.I, s.O,
                    p->self_mask, nullptr, (int)s.Rp - 1, out};
        if (bf16) return launch_prop_dense_tiled(Gm, static_cast<const __nv_bfloat16*>(X), s.N, p->bias, max_chunks(s), st, true, s.nnz);
        return launch_prop_dense_tiled(Gm, static_cast<const float*>(X), s.N, p->bias, max_chunks(s), st, true, s.nnz);
    }
    // LP block decomposition with a dense self-loop weight (layers.py:534-548): the block relations take the generic
    // kernel, the self-loop relation — N edges through one (I, O) matrix, a plain GEMM — the tiled one
    const bool split_self = p->form == RGCN_W_BLOCK && p->blocks_self && !p->featureless && (int64_t)s.I * s.O >= 1024 &&
                            s.nnz > 0 && split_self_enabled();
    if (split_self) A.skip_rel_plus1 = (int)s.Rp;
    // blocks of any size: edge-parallel relation-major kernel (needs the self-loop relation, if dense, served separately)
    const bool block_edges = p->form == RGCN_W_BLOCK && !p->featureless && (!p->blocks_self || split_self) &&
                             (!p->self_mask || split_self) && s.nnz > 0 && block_edges_shape(s.nb, s.bi, s.bo, s.I, s.O);
    if (block_edges) {
        const int64_t n = s.N * (int64_t)s.O;
        RGCN_LAUNCH(k_init_rows, grid_for(n, 256), 256, 0, st, out, n, s.O, p->bias);
        BlockEdgeArgs B{g->r_relptr, g->r_chunkptr, (int)s.Rp, g->r_src, g->r_dst, g->r_val, p->blocks, s.Rb, s.nb, s.bi, s.bo,
                        s.I, s.O, out};
        if (bf16) rc = launch_block_edges(B, static_cast<const __nv_bfloat16*>(X), max_chunks(s), st);
        else rc = launch_block_edges(B, static_cast<const float*>(X), max_chunks(s), st);
    } else if (bf16) rc = launch_prop(A, static_cast<const __nv_bfloat16*>(X), st);
    else rc = launch_prop(A, static_cast<const float*>(X), st);
    if (rc || !split_self) return rc;
    GemmArgs Gs{g->r_relptr, g->r_chunkptr, (int)s.Rp, g->r_src, g->r_dst, g->r_val, p->blocks_self, s.I, s.O,
                p->self_mask, nullptr, (int)s.Rp - 1, out, (int)s.Rp};
    if (bf16) return launch_prop_dense_tiled(Gs, static_cast<const __nv_bfloat16*>(X), s.N, nullptr, max_chunks(s), st, false);
    return launch_prop_dense_tiled(Gs, static_cast<const float*>(X), s.N, nullptr, max_chunks(s), st, false);
}

extern "C" size_t rgcn_backward_workspace_bytes(const rgcn_graph* g, const rgcn_params* p, int x_dtype) {
    (void)x_dtype;
    Shape s;
    if (check_common(g, p, &s, "rgcn_backward_workspace_bytes") || p->featureless) return 0;
    size_t w = align_up(s.w_elems * sizeof(float));
    size_t bytes = 0;
    switch (p->form) {
        case RGCN_W_DENSE: bytes = w; break;                       // W^T
        case RGCN_W_BASIS: bytes = 3 * w; break;                   // W_eff, W_eff^T, gW_eff
        case RGCN_W_BLOCK:
            bytes = align_up(s.blocks_elems * sizeof(float)) +
                    (p->blocks_self ? align_up((size_t)s.I * s.O * sizeof(float)) : 0);
            break;
        default: break;
    }
    if (x_dtype == RGCN_BF16) bytes += align_up((size_t)s.N * s.O * 2);   // bf16 copy of grad_out (tensor-core path)
    if (x_dtype == RGCN_BF16) bytes += align_up((size_t)s.N * s.I * 4);   // fp32 staging of a bf16 feature gradient
    if (umma_path(p, s, x_dtype == RGCN_BF16)) bytes += umma_wt_bytes(s.Rp, s.I, s.O);   // bf16 weights of the tensor-core GEMM
    if (fused_path(g, p, s, x_dtype == RGCN_BF16, true)) return bytes + fused_ws_bytes(s.Rp, s.nb);
    if (tiled_path(g, p, s, x_dtype == RGCN_BF16))
        return bytes + tiled_counter_bytes(g->num_tiles) + tiled_ring_bytes(g, s.I) + wfrag_bytes(s.Rp, s.nb);
    // bf16 16x16-block tensor-core backward: bf16 feature-gradient messages, sized by the predicate rgcn_backward uses
    // (rel_path() below would give up above kMaxMsgBytes of fp32 messages and under-report this path)
    if (x_dtype == RGCN_BF16 && p->form == RGCN_W_BLOCK && !p->blocks_self && !p->self_mask && s.nnz > 0 &&
        mma_shape_supported(s.nb, s.bi, s.bo))
        return bytes + align_up((size_t)s.nnz * s.I * 2);
    RelShape rs; size_t msg = 0;
    if (rel_path(p, s, false, true, &rs, &msg)) bytes += msg;     // feature-gradient messages (fp32)
    return bytes;
}

extern "C" int rgcn_backward(const rgcn_graph* g, const rgcn_params* p, const void* X, int x_dtype,
                             const float* G, const rgcn_grads* gr, void* ws, size_t ws_bytes, rgcn_stream_t stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    Shape s;
    int rc = check_common(g, p, &s, "rgcn_backward");
    if (rc) return rc;
    RGCN_REQUIRE(G && gr, RGCN_ERR_ARG, "rgcn_backward: grad_out or grads is NULL");
    RGCN_REQUIRE(p->featureless || X, RGCN_ERR_ARG, "rgcn_backward: features is NULL");
    RGCN_REQUIRE(x_dtype == RGCN_F32 || x_dtype == RGCN_BF16, RGCN_ERR_ARG, "rgcn_backward: unknown dtype %d", x_dtype);
    size_t need = rgcn_backward_workspace_bytes(g, p, x_dtype);
    RGCN_REQUIRE(ws_bytes >= need && (ws || need == 0), RGCN_ERR_WORKSPACE, "rgcn_backward: workspace %zu < %zu", ws_bytes, need);
    Carver carve(ws);
    const int64_t IO = (int64_t)s.I * s.O;
    const bool gx_bf16 = gr->features && gr->features_dtype == RGCN_BF16;
    RGCN_REQUIRE(!gx_bf16 || x_dtype == RGCN_BF16, RGCN_ERR_ARG, "rgcn_backward: bf16 feature gradient needs bf16 features");
    // paths that produce an fp32 feature gradient write it here first when the caller wants bf16
    float* gx_f32 = gx_bf16 ? carve.take<float>((size_t)s.N * s.I) : static_cast<float*>(gr->features);
    auto finish_gx = [&]() -> int {
        if (!gx_bf16) return RGCN_OK;
        const int64_t n = s.N * (int64_t)s.I;
        RGCN_LAUNCH(k_cast_bf16, grid_for(n, 256), 256, 0, st, gx_f32, n, static_cast<__nv_bfloat16*>(gr->features));
        return RGCN_OK;
    };

    // ---- bf16 features + 16x16 blocks (untiled): G is rounded to bf16 once, fused with the bias gradient, then ONE
    //      tensor-core pass produces the feature-gradient messages and the weight gradient
    const bool mma_bwd = x_dtype == RGCN_BF16 && !p->featureless && p->form == RGCN_W_BLOCK && !p->blocks_self &&
                         !p->self_mask && s.nnz > 0 && mma_shape_supported(s.nb, s.bi, s.bo) &&
                         (gr->features || gr->blocks);
    const bool fused_bwd = mma_bwd && gr->features && fused_path(g, p, s, true, true);
    const bool tiled_bwd = mma_bwd && !fused_bwd && gr->features && tiled_path(g, p, s, true);
    RGCN_REQUIRE(!mma_bwd || fused_bwd || tiled_bwd || !gr->features || align_up((size_t)s.nnz * s.I * 2) <= kMaxMsgBytes,
                 RGCN_ERR_UNSUPPORTED, "rgcn_backward: message buffer too large; build the plan with tile_edges > 0");
    if (mma_bwd) {
        __nv_bfloat16* gb16 = reinterpret_cast<__nv_bfloat16*>(carve.take<char>(align_up((size_t)s.N * s.O * 2)));
        if (gr->bias) RGCN_CHECK_CUDA(cudaMemsetAsync(gr->bias, 0, (size_t)s.O * sizeof(float), st));
        rc = launch_cast_colsum(G, s.N, s.O, gb16, gr->bias, st);
        if (rc) return rc;
        if (gr->blocks) RGCN_CHECK_CUDA(cudaMemsetAsync(gr->blocks, 0, s.blocks_elems * sizeof(float), st));
        if (fused_bwd) {
            // weight gradient: relation-major pass without messages; feature gradient: fused row-block kernel on the
            // source-row lists with W^T slices, gathering the bf16 copy of grad_out
            if (gr->blocks) {
                RelArgs Rw{g->r_relptr, g->r_chunkptr, (int)s.Rp, g->r_src, g->r_dst, nullptr, g->r_val, p->blocks, s.nb};
                rc = launch_rel_mma_bwd(Rw, static_cast<const __nv_bfloat16*>(X), gb16, nullptr, gr->blocks, max_chunks(s), st);
                if (rc) return rc;
            }
            void* fws = carve.take<char>(fused_ws_bytes(s.Rp, s.nb));
            if (gx_bf16 && g->fuse_split[1] == 0)
                return launch_fused_rows<__nv_bfloat16>(g, true, p->blocks, nullptr, gb16,
                                                        static_cast<__nv_bfloat16*>(gr->features), fws, st, 0, -1, nullptr, 0, s.nb);
            rc = launch_fused_rows<float>(g, true, p->blocks, nullptr, gb16, gx_f32, fws, st, 0, -1, nullptr, 0, s.nb);
            if (rc) return rc;
            return finish_gx();
        }
        if (tiled_bwd) {
            // weight gradient: untiled relation-major pass (needs full relation batches); feature gradient: span kernel
            // on the source tiling with W^T fragments, messages stay in the ring
            if (gr->blocks) {
                RelArgs Rw{g->r_relptr, g->r_chunkptr, (int)s.Rp, g->r_src, g->r_dst, nullptr, g->r_val, p->blocks, s.nb};
                rc = launch_rel_mma_bwd(Rw, static_cast<const __nv_bfloat16*>(X), gb16, nullptr, gr->blocks, max_chunks(s), st);
                if (rc) return rc;
            }
            int32_t* counters = reinterpret_cast<int32_t*>(carve.take<char>(tiled_counter_bytes(g->num_tiles)));
            __nv_bfloat16* ring = reinterpret_cast<__nv_bfloat16*>(carve.take<char>(tiled_ring_bytes(g, s.I)));
            uint32_t* frag = reinterpret_cast<uint32_t*>(carve.take<char>(wfrag_bytes(s.Rp, s.nb)));
            RGCN_CHECK_CUDA(cudaMemsetAsync(counters, 0, tiled_counter_bytes(g->num_tiles), st));
            rc = launch_pack_wfrag(p->blocks, s.Rp, s.nb, true, frag, st);
            if (rc) return rc;
            TiledArgs T = make_tiled_args(g, true, s.nb, frag, nullptr, counters);
            rc = launch_tiled_span(T, gb16, ring, gx_f32, st);
            if (rc) return rc;
            return finish_gx();
        }
        RelArgs R{g->r_relptr, g->r_chunkptr, (int)s.Rp, g->r_src, g->r_dst, g->r_sslot, g->r_val, p->blocks, s.nb};
        __nv_bfloat16* msg = nullptr;
        if (gr->features) msg = reinterpret_cast<__nv_bfloat16*>(carve.take<char>(align_up((size_t)s.nnz * s.I * 2)));
        rc = launch_rel_mma_bwd(R, static_cast<const __nv_bfloat16*>(X), gb16, msg, gr->blocks, max_chunks(s), st);
        if (rc) return rc;
        if (gr->features && gx_bf16) {
            // a row-sharded caller's plan holds only the edges out of rows [row_lo, row_hi): sum (and write) those rows only
            const bool ranged = p->row_hi > p->row_lo && p->row_lo >= 0 && p->row_hi <= s.N;
            const int64_t lo = ranged ? p->row_lo : 0, n = ranged ? p->row_hi - p->row_lo : s.N;
            return launch_row_sum(g->s_rowptr + lo, n, s.I, msg, (const float*)nullptr,
                                  static_cast<__nv_bfloat16*>(gr->features) + lo * s.I, g->s_long, g->status + 5,
                                  g->num_long_src, s.nnz, st);
        }
        if (gr->features)
            return launch_row_sum(g->s_rowptr, s.N, s.I, msg, (const float*)nullptr, static_cast<float*>(gr->features),
                                  g->s_long, g->status + 5, g->num_long_src, s.nnz, st);
        return RGCN_OK;
    }

    // ---- gbias
    if (gr->bias) {
        RGCN_CHECK_CUDA(cudaMemsetAsync(gr->bias, 0, (size_t)s.O * sizeof(float), st));
        int64_t rows_per_block = (s.N + kNumSMs * 4 - 1) / (kNumSMs * 4);
        if (rows_per_block < 64) rows_per_block = 64;
        int grid = (int)((s.N + rows_per_block - 1) / rows_per_block);
        RGCN_LAUNCH(k_colsum, grid, 256, (size_t)s.O * sizeof(float), st, G, s.N, s.O, rows_per_block, gr->bias);
    }

    // ---- featureless: weight rows are the messages
    if (p->featureless) {
        FeaturelessGradArgs F{};
        F.rowptr = g->s_rowptr; F.col = g->s_dst; F.rel = g->s_rel; F.val = g->s_val;
        F.N = s.N; F.num_rels = (int)s.Rp; F.form = p->form; F.O = s.O; F.B = s.B; F.nb = s.nb; F.bi = s.bi; F.bo = s.bo;
        F.comps = p->comps; F.bases = p->bases;
        F.gW = gr->weights; F.gblocks = gr->blocks; F.gbases = gr->bases; F.gcomps = gr->comps;
        if (p->form == RGCN_W_DENSE) {
            if (!gr->weights) return RGCN_OK;
            RGCN_CHECK_CUDA(cudaMemsetAsync(gr->weights, 0, (size_t)s.Rp * s.N * s.O * sizeof(float), st));
        } else if (p->form == RGCN_W_BLOCK) {
            if (!gr->blocks) return RGCN_OK;
            RGCN_CHECK_CUDA(cudaMemsetAsync(gr->blocks, 0, s.blocks_elems * sizeof(float), st));
        } else {
            if (!gr->bases && !gr->comps) return RGCN_OK;
            RGCN_REQUIRE(gr->bases && gr->comps, RGCN_ERR_ARG, "rgcn_backward: basis gradients come as a pair");
            RGCN_CHECK_CUDA(cudaMemsetAsync(gr->comps, 0, (size_t)s.Rp * s.B * sizeof(float), st));
        }
        return launch_featureless_grad(F, G, st);
    }

    // ---- effective / transposed weights for the feature gradient
    __nv_bfloat16* gb16_dense = nullptr;     // bf16 copy of grad_out shared by the tensor-core GEMMs of dense layers
    float* weff = nullptr;
    if (p->form == RGCN_W_BASIS) {
        weff = carve.take<float>((size_t)s.Rp * IO);
        dim3 grid(grid_for(IO, 256), (unsigned)s.Rp);
        RGCN_LAUNCH(k_basis_combine, grid, 256, 0, st, p->comps, p->bases, (int)s.Rp, s.B, IO, weff);
    }
    if (gr->features) {
        PropArgs A{};
        A.rowptr = g->s_rowptr; A.col = g->s_dst; A.rel = g->s_rel; A.val = g->s_val;
        A.nrows = s.N; A.N = s.N;
        fill_weights(A, p, s);
        A.I = s.O; A.O = s.I; A.bi = s.bo; A.bo = s.bi;
        A.in_mask = p->self_mask; A.out = gx_f32;
        A.long_list = g->s_long; A.long_count = g->status + 5; A.nnz_hint = s.nnz; A.num_long = g->num_long_src;
        if (p->form == RGCN_W_DENSE || p->form == RGCN_W_BASIS) {
            float* wt = carve.take<float>((size_t)s.Rp * IO);
            rc = launch_transpose(p->form == RGCN_W_BASIS ? weff : p->weights, s.Rp, s.I, s.O, wt, st);
            if (rc) return rc;
            A.form = RGCN_W_DENSE; A.W = wt;
        } else if (p->form == RGCN_W_BLOCK) {
            float* bt = carve.take<float>(s.blocks_elems);
            rc = launch_transpose(p->blocks, (int64_t)s.Rb * s.nb, s.bi, s.bo, bt, st);
            if (rc) return rc;
            A.blocks = bt;
            if (p->blocks_self) {
                float* stp = carve.take<float>((size_t)IO);
                rc = launch_transpose(p->blocks_self, 1, s.I, s.O, stp, st);
                if (rc) return rc;
                A.blocks_self = stp;
            }
        }
        RelShape rs; size_t msg_bytes = 0;
        if (rel_path(p, s, false, true, &rs, &msg_bytes)) {
            // same two kernels as the forward: gather G[dst], transposed weights, slots in source-major order
            RelArgs R{g->r_relptr, g->r_chunkptr, (int)s.Rp, g->r_dst, nullptr, g->r_sslot, g->r_val,
                      A.form == RGCN_W_BLOCK ? A.blocks : A.W, rs.nb};
            float* msg = reinterpret_cast<float*>(carve.take<char>(msg_bytes));
            rc = launch_rel_transform(R, rs.bi, rs.bo, G, msg, max_chunks(s), st);
            if (rc) return rc;
            rc = launch_row_sum(g->s_rowptr, s.N, s.I, msg, (const float*)nullptr, gx_f32, g->s_long, g->status + 5,
                                g->num_long_src, s.nnz, st);
        } else if (umma_path(p, s, x_dtype == RGCN_BF16)) {
            // gX[o_e] += val_e * bf16(G)[s_e] W_p^T: the inner dimension is the layer's output, so W_p is K-major as stored
            __nv_bfloat16* gb16 = reinterpret_cast<__nv_bfloat16*>(carve.take<char>(align_up((size_t)s.N * s.O * 2)));
            rc = launch_cast_colsum(G, s.N, s.O, gb16, nullptr, st);
            if (rc) return rc;
            gb16_dense = gb16;
            __nv_bfloat16* wb = reinterpret_cast<__nv_bfloat16*>(carve.take<char>(umma_wt_bytes(s.Rp, s.I, s.O)));
            const int64_t cnt = (int64_t)s.Rp * IO;
            RGCN_LAUNCH(k_pack_wt_bf16, grid_for(cnt, 256), 256, 0, st, p->form == RGCN_W_BASIS ? weff : p->weights, cnt,
                        s.I, s.O, 0, wb);
            const int64_t n = s.N * (int64_t)s.I;
            RGCN_LAUNCH(k_init_rows, grid_for(n, 256), 256, 0, st, gx_f32, n, s.I, (const float*)nullptr);
            UmmaArgs U{g->r_relptr, g->r_chunkptr, (int)s.Rp, g->r_dst, g->r_src, g->r_val, s.O, s.I, 0, gx_f32};
            rc = launch_gemm_umma(U, gb16, s.N, wb, max_chunks(s), st);
        } else if (dense_tiled_shape(A.form, 0, s.O, s.I, s.nnz)) {
            GemmArgs Gm{g->r_relptr, g->r_chunkptr, (int)s.Rp, g->r_dst, g->r_src, g->r_val, A.W, s.O, s.I,
                        nullptr, p->self_mask, (int)s.Rp - 1, gx_f32};
            rc = launch_prop_dense_tiled(Gm, G, s.N, (const float*)nullptr, max_chunks(s), st, true, s.nnz);
        } else {
            const bool split_self = p->form == RGCN_W_BLOCK && p->blocks_self && IO >= 1024 && s.nnz > 0 && split_self_enabled();
            if (split_self) A.skip_rel_plus1 = (int)s.Rp;
            const bool block_edges = p->form == RGCN_W_BLOCK && (!p->blocks_self || split_self) &&
                                     (!p->self_mask || split_self) && s.nnz > 0 && block_edges_shape(s.nb, s.bo, s.bi, s.O, s.I);
            if (block_edges) {
                const int64_t n = s.N * (int64_t)s.I;
                RGCN_LAUNCH(k_init_rows, grid_for(n, 256), 256, 0, st, gx_f32, n, s.I, (const float*)nullptr);
                BlockEdgeArgs B{g->r_relptr, g->r_chunkptr, (int)s.Rp, g->r_dst, g->r_src, g->r_val, A.blocks, s.Rb, s.nb,
                                s.bo, s.bi, s.O, s.I, gx_f32};
                rc = launch_block_edges(B, G, max_chunks(s), st);
            } else {
                rc = launch_prop(A, G, st);
            }
            if (!rc && split_self) {
                GemmArgs Gs{g->r_relptr, g->r_chunkptr, (int)s.Rp, g->r_dst, g->r_src, g->r_val, A.blocks_self, s.O, s.I,
                            nullptr, p->self_mask, (int)s.Rp - 1, gx_f32, (int)s.Rp};
                rc = launch_prop_dense_tiled(Gs, G, s.N, (const float*)nullptr, max_chunks(s), st, false);
            }
        }
        if (rc) return rc;
        rc = finish_gx();
        if (rc) return rc;
    } else if (p->form == RGCN_W_DENSE || p->form == RGCN_W_BASIS) {
        carve.take<float>((size_t)s.Rp * IO);            // keep the layout identical to the query
    }

    // ---- weight gradients
    WGradArgs Wg{};
    Wg.relptr = g->r_relptr; Wg.dst = g->r_dst; Wg.src = g->r_src; Wg.val = g->r_val;
    Wg.form = p->form; Wg.I = s.I; Wg.O = s.O; Wg.nb = s.nb; Wg.bi = s.bi; Wg.bo = s.bo;
    Wg.self_rel = p->blocks_self ? (int)s.Rp - 1 : -1; Wg.num_block_rels = s.Rb;
    Wg.mask = p->self_mask; Wg.mask_rel = (int)s.Rp - 1;
    bool want = false;
    if (p->form == RGCN_W_DENSE) {
        want = gr->weights != nullptr;
        if (want) { RGCN_CHECK_CUDA(cudaMemsetAsync(gr->weights, 0, (size_t)s.Rp * IO * sizeof(float), st)); Wg.gW = gr->weights; }
    } else if (p->form == RGCN_W_DIAG) {
        want = gr->weights != nullptr;
        if (want) { RGCN_CHECK_CUDA(cudaMemsetAsync(gr->weights, 0, (size_t)s.Rp * s.I * sizeof(float), st)); Wg.gW = gr->weights; }
    } else if (p->form == RGCN_W_BLOCK) {
        want = gr->blocks || gr->blocks_self;
        if (gr->blocks) RGCN_CHECK_CUDA(cudaMemsetAsync(gr->blocks, 0, s.blocks_elems * sizeof(float), st));
        if (gr->blocks_self) RGCN_CHECK_CUDA(cudaMemsetAsync(gr->blocks_self, 0, (size_t)IO * sizeof(float), st));
        Wg.gblocks = gr->blocks; Wg.gself = gr->blocks_self;
    } else {
        want = gr->bases || gr->comps;
        if (want) {
            Wg.form = RGCN_W_DENSE;
            Wg.gW = carve.take<float>((size_t)s.Rp * IO);
            RGCN_CHECK_CUDA(cudaMemsetAsync(Wg.gW, 0, (size_t)s.Rp * IO * sizeof(float), st));
        }
    }
    if (!want || s.nnz == 0) {
        if (want && p->form == RGCN_W_BASIS) {
            if (gr->comps) RGCN_CHECK_CUDA(cudaMemsetAsync(gr->comps, 0, (size_t)s.Rp * s.B * sizeof(float), st));
            if (gr->bases) RGCN_CHECK_CUDA(cudaMemsetAsync(gr->bases, 0, (size_t)s.B * IO * sizeof(float), st));
        }
        return RGCN_OK;
    }
    RelShape ws_shape; size_t unused = 0;
    rc = 1;                                   // > 0: not served by the relation-batched kernel
    if (rel_path(p, s, x_dtype == RGCN_BF16, false, &ws_shape, &unused)) {
        RelArgs R{g->r_relptr, g->r_chunkptr, (int)s.Rp, g->r_src, g->r_dst, nullptr, g->r_val, nullptr, ws_shape.nb};
        float* target = (p->form == RGCN_W_BLOCK) ? gr->blocks : Wg.gW;
        if (x_dtype == RGCN_BF16)
            rc = launch_rel_wgrad(R, ws_shape.bi, ws_shape.bo, static_cast<const __nv_bfloat16*>(X), G, target,
                                  max_chunks(s), st);
        else
            rc = launch_rel_wgrad(R, ws_shape.bi, ws_shape.bo, static_cast<const float*>(X), G, target, max_chunks(s), st);
    }
    if (rc > 0 && umma_path(p, s, x_dtype == RGCN_BF16) && umma_wgrad_shape_supported(s.I, s.O)) {
        if (!gb16_dense) {
            gb16_dense = reinterpret_cast<__nv_bfloat16*>(carve.take<char>(align_up((size_t)s.N * s.O * 2)));
            rc = launch_cast_colsum(G, s.N, s.O, gb16_dense, nullptr, st);
            if (rc) return rc;
        }
        UmmaWgradArgs U{g->r_relptr, g->r_chunkptr, (int)s.Rp, g->r_src, g->r_dst, g->r_val, s.I, s.O, 0, Wg.gW};
        rc = launch_wgrad_umma(U, static_cast<const __nv_bfloat16*>(X), gb16_dense, s.N, max_chunks(s), st);
    }
    if (rc > 0) {
        // the dense self-loop weight of an LP block layer: its gradient is one gathered GEMM over the self-loop relation
        WGradArgs Ws = Wg;
        const bool split_self = p->form == RGCN_W_BLOCK && gr->blocks_self && IO >= 1024 && split_self_enabled();
        if (split_self) Wg.gself = nullptr;
        rc = RGCN_OK;
        const bool block_edges = p->form == RGCN_W_BLOCK && gr->blocks && (!gr->blocks_self || split_self) &&
                                 (!p->self_mask || split_self) && block_edges_shape(s.nb, s.bi, s.bo, s.I, s.O);
        if (block_edges) {
            BlockEdgeArgs B{g->r_relptr, g->r_chunkptr, (int)s.Rp, g->r_src, g->r_dst, g->r_val, nullptr, s.Rb, s.nb, s.bi,
                            s.bo, s.I, s.O, nullptr};
            if (x_dtype == RGCN_BF16) rc = launch_block_wgrad(B, static_cast<const __nv_bfloat16*>(X), G, gr->blocks, max_chunks(s), st);
            else rc = launch_block_wgrad(B, static_cast<const float*>(X), G, gr->blocks, max_chunks(s), st);
        } else if (!split_self || gr->blocks) {
            if (x_dtype == RGCN_BF16) rc = launch_wgrad(Wg, static_cast<const __nv_bfloat16*>(X), G, s.nnz, (int)s.Rp, st);
            else rc = launch_wgrad(Wg, static_cast<const float*>(X), G, s.nnz, (int)s.Rp, st);
        }
        if (!rc && split_self) {
            Ws.form = RGCN_W_DENSE; Ws.gW = gr->blocks_self; Ws.rel0 = (int)s.Rp - 1;
            if (x_dtype == RGCN_BF16) rc = launch_wgrad(Ws, static_cast<const __nv_bfloat16*>(X), G, s.nnz, (int)s.Rp, st);
            else rc = launch_wgrad(Ws, static_cast<const float*>(X), G, s.nnz, (int)s.Rp, st);
        }
    }
    if (rc) return rc;
    if (p->form == RGCN_W_BASIS) {
        if (gr->comps) {
            dim3 grid((unsigned)s.Rp, (unsigned)s.B);
            RGCN_LAUNCH(k_basis_grad_comps, grid, 256, 0, st, Wg.gW, p->bases, s.B, IO, gr->comps);
        }
        if (gr->bases) {
            dim3 grid(grid_for(IO, 256), (unsigned)s.B);
            RGCN_LAUNCH(k_basis_grad_bases, grid, 256, 0, st, Wg.gW, p->comps, (int)s.Rp, s.B, IO, gr->bases);
        }
    }
    return RGCN_OK;
}
