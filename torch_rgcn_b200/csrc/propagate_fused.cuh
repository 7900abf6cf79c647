// Fused row-block kernel for bf16 features and four 16x16 weight blocks (64 -> 64, the AM-shaped layer).
//
// The two-phase kernels of propagate_mma.cuh write one bf16 message per edge to HBM and read it back in the row
// sum: 256 B per edge of round trip on top of the 128 B gather.  Here a CTA owns a block of `fuse_rows`
// consecutive output rows and keeps their fp32 sums in shared memory, so an edge costs its gather and nothing
// else:
//
//   work item  = (row block, range of 16-entry tiles) from the plan's rgcn_fused list: the block's edges sorted by
//                (relation, row), every (block, relation) run padded to whole tiles -> one relation per MMA tile
//   pipeline   = 64-entry stages (4 tiles): the gathered rows (64 x 128 B, XOR-swizzled like propagate_mma.cuh) and
//                the {row, val} records land through cp.async, kFuAhead stages in flight; the gather indices and
//                the weight slices are requested equally early into register rings (see the note in the kernel)
//   ownership  = warp w owns output columns [8w, 8w + 8): per tile it runs ONE mma.sync.m16n8k16 (A = the 16
//                inputs of block w / 2 of the 16 gathered rows, B = its 16 x 8 weight slice, prefetched from the
//                packed table one stage ahead) and adds val * result into its private column slice of the
//                shared-memory tile.  No two warps ever touch the same address, so the accumulation needs no
//                atomics and its order is fixed by the plan: results are run-to-run deterministic.
//   duplicates = tiles in which two entries share a row (a (row, relation) segment longer than one edge) are
//                flagged by the plan (bit 31 of tile_rel) and accumulated one entry at a time
//   banks      = a row's 32-byte slice sits in bank group row % 4; with fuse_order 1 the plan places a run's
//                edges so that the four rows of one access phase come from different groups when possible
//   flush      = out[row] = bias + sum, 256-byte coalesced rows (fp32, or bf16 for a bf16 feature gradient); items
//                of a split (hub) block add into rows pre-set by k_fused_init_shared with fp32 atomics
//
// The same kernel serves the forward (gather X[o], W) and the feature gradient (gather the bf16 copy of
// grad_out[s], W^T) on the plan's ff / fb lists.
#pragma once
#include "common.cuh"
#include "propagate_fast.cuh"
#include "propagate_mma.cuh"

namespace rgcn {

constexpr int kFuAhead = 5;                            // stages in flight (gathers, indices, weight slices)
constexpr int kFuStages = kFuAhead + 1;
constexpr int kFuStageEntries = 64;
constexpr int kFuXBytes = kFuStageEntries * 128;       // gathered rows of one stage
constexpr int kFuRvBytes = kFuStageEntries * 8;        // {row, val} records of one stage
constexpr int kFuWidth = 64;                           // I == O == 4 blocks of 16

struct FusedArgs {
    rgcn_fused fl;
    int n_items;               // host copy of fl.meta[0]
    int fuse_rows;
    long long N;
    const uint2* wslice;       // [(p * 8 + w) * 32 + lane]: the two B-operand registers of warp w's 16 x 8 slice
    const float* bias;         // added at the flush of unshared items
    int32_t* counter;          // work-queue head, zeroed by the launcher
};

__host__ __device__ inline size_t fused_slice_stride(int fuse_rows) { return (size_t)fuse_rows * 32 + 16; }
__host__ __device__ inline size_t fused_tile_bytes(int fuse_rows) { return 8 * fused_slice_stride(fuse_rows); }   // multiple of 128
inline size_t fused_smem_bytes(int fuse_rows) {
    return fused_tile_bytes(fuse_rows) + (size_t)kFuStages * (kFuXBytes + kFuRvBytes) +
           RGCN_FUSE_MAX_ITEM_TILES * sizeof(int32_t) + 16;
}

// slices[(p * 8 + w) * 32 + lane] = {b0, b1} of lane (g = lane / 4, t = lane % 4) for the 16 x 8 slice of block
// w / 2, output columns 8 (w % 2) .. + 7.  transpose = 0: B[k][n] = W[k][n] (forward); 1: B[k][n] = W[n][k].
__global__ void k_pack_wslice(const float* __restrict__ W, int Rp, int transpose, uint2* __restrict__ slices) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)Rp * 256) return;
    const int lane = (int)(i & 31), w = (int)((i >> 5) & 7);
    const long long p = i >> 8;
    const int g = lane >> 2, t = lane & 3, n = (w & 1) * 8 + g;
    const float* wb = W + ((size_t)p * 4 + (w >> 1)) * 256;
    auto at = [&](int k) { return transpose ? wb[n * 16 + k] : wb[k * 16 + n]; };
    slices[i] = make_uint2(pack_bf16x2(at(2 * t), at(2 * t + 1)), pack_bf16x2(at(2 * t + 8), at(2 * t + 9)));
}

// rows of split blocks start from the bias (or zero): their items add partial sums with atomics
__global__ void k_fused_init_shared(const int32_t* __restrict__ items, int n_items, const int32_t* __restrict__ blk_tile,
                                    int fuse_rows, long long N, const float* __restrict__ bias,
                                    float* __restrict__ out) {
    const int q = blockIdx.x;
    if (q >= n_items) return;
    const int4 it = __ldg(reinterpret_cast<const int4*>(items) + q);
    if (!it.w || it.y != __ldg(blk_tile + it.x)) return;        // only the first item of a split block
    const long long row0 = (long long)it.x * fuse_rows;
    const int nrows = (int)min((long long)fuse_rows, N - row0);
    for (int i = threadIdx.x; i < nrows * (kFuWidth / 4); i += blockDim.x) {
        const int c4 = i % (kFuWidth / 4);
        const float4 b = bias ? __ldg(reinterpret_cast<const float4*>(bias) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        reinterpret_cast<float4*>(out + (size_t)row0 * kFuWidth)[i] = b;
    }
}

template <typename OT>
__global__ void __launch_bounds__(256, 1) k_fused_rows(FusedArgs A, const __nv_bfloat16* __restrict__ X,
                                                       OT* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem_fused[];
    const int FR = A.fuse_rows;
    const size_t slice_stride = fused_slice_stride(FR);
    const size_t tile_bytes = fused_tile_bytes(FR);
    unsigned char* tile = smem_fused;
    unsigned char* xst = tile + tile_bytes;
    unsigned char* rvst = xst + (size_t)kFuStages * kFuXBytes;
    int32_t* s_rel = reinterpret_cast<int32_t*>(rvst + (size_t)kFuStages * kFuRvBytes);
    int32_t* s_item = s_rel + RGCN_FUSE_MAX_ITEM_TILES;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int kb = warp >> 1;                                   // weight block of this warp's columns
    const int j0 = tid >> 3, ch = tid & 7;                      // gather role: entries j0, j0 + 32 of a stage, 16-byte piece ch
    const unsigned char* Xb = reinterpret_cast<const unsigned char*>(X);
    const uint2* wmine = A.wslice + (size_t)warp * 32 + lane;
    unsigned char* myslice = tile + (size_t)warp * slice_stride + t * 8;
    float4 bias_lo = make_float4(0.f, 0.f, 0.f, 0.f), bias_hi = bias_lo;   // flush role: columns 8 (tid % 8) .. + 7
    if (A.bias) {
        bias_lo = __ldg(reinterpret_cast<const float4*>(A.bias) + 2 * (tid & 7));
        bias_hi = __ldg(reinterpret_cast<const float4*>(A.bias) + 2 * (tid & 7) + 1);
    }

    while (true) {
        if (tid == 0) *s_item = atomicAdd(A.counter, 1);
        __syncthreads();
        const int q = *s_item;
        if (q >= A.n_items) break;
        const int4 item = __ldg(reinterpret_cast<const int4*>(A.fl.items) + q);
        const int blk = item.x, t0 = item.y, nt = item.z - item.y, shared_item = item.w;
        const int nst = (nt + 3) >> 2;
        const int n_ent = nt * RGCN_FUSE_TILE;
        const int32_t* colp = A.fl.col + (size_t)t0 * RGCN_FUSE_TILE;
        const unsigned char* rvp = reinterpret_cast<const unsigned char*>(A.fl.rv) + (size_t)t0 * RGCN_FUSE_TILE * 8;

        for (int i = tid; i < (int)(tile_bytes / 16); i += 256)
            reinterpret_cast<float4*>(tile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = tid; i < nt; i += 256) s_rel[i] = __ldg(A.fl.tile_rel + t0 + i);

        auto load_cols = [&](int k, int& c0, int& c1) {         // gather indices of stage k (-1: nothing to gather)
            const int e0 = k * kFuStageEntries + j0, e1 = e0 + 32;
            c0 = e0 < n_ent ? __ldg(colp + e0) : -1;
            c1 = e1 < n_ent ? __ldg(colp + e1) : -1;
        };
        auto issue = [&](int k, int c0, int c1) {
            if (k < nst) {
                unsigned char* xs = xst + (size_t)(k % kFuStages) * kFuXBytes;
                cp_async16(xs + (j0 >> 4) * kTileBytes + tile_off(j0 & 15, ch),
                           Xb + (size_t)(c0 >= 0 ? c0 : 0) * 128 + ch * 16, c0 >= 0);
                cp_async16(xs + ((j0 >> 4) + 2) * kTileBytes + tile_off(j0 & 15, ch),
                           Xb + (size_t)(c1 >= 0 ? c1 : 0) * 128 + ch * 16, c1 >= 0);
                if (tid < 32) {                                  // 64 records of 8 bytes = 32 pieces of 16 bytes
                    const int ent = k * kFuStageEntries + tid * 2;
                    const bool ok = ent < n_ent;
                    cp_async16(rvst + (size_t)(k % kFuStages) * kFuRvBytes + tid * 16,
                               ok ? rvp + (size_t)ent * 8 : reinterpret_cast<const unsigned char*>(A.fl.rv), ok);
                }
            }
            cp_async_commit();
        };
        auto load_w = [&](int k, uint2 (&w)[4]) {                // weight slices of the four tiles of stage k
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ti = 4 * k + j;
                const int rel = ti < nt ? (s_rel[ti] & 0x7fffffff) : 0;
                w[j] = __ldg(wmine + (size_t)rel * 256);
            }
        };

        // Everything the loop reads from global memory is requested kFuAhead stages early: the L1 returns data in
        // request order, so a plain load issued now completes only after every gather already in flight -- a load
        // consumed one stage later would stall for the whole depth of the gather pipeline.
        //   colr[u] : gather indices of stage k + kFuAhead   (k % kFuAhead == u), refilled for stage k + 2 kFuAhead
        //   wr[u]   : weight slices of stage k,               refilled for stage k + kFuAhead
        int colr[kFuAhead][2];
        uint2 wr[kFuAhead][4];
        {
            int pc[kFuAhead][2];
#pragma unroll
            for (int d = 0; d < kFuAhead; ++d) load_cols(d, pc[d][0], pc[d][1]);
#pragma unroll
            for (int d = 0; d < kFuAhead; ++d) load_cols(kFuAhead + d, colr[d][0], colr[d][1]);
            __syncthreads();                                    // tile zeroed, s_rel filled
#pragma unroll
            for (int d = 0; d < kFuAhead; ++d) load_w(d, wr[d]);
#pragma unroll
            for (int d = 0; d < kFuAhead; ++d) issue(d, pc[d][0], pc[d][1]);
        }

        for (int k0 = 0; k0 < nst; k0 += kFuAhead) {
#pragma unroll
            for (int u = 0; u < kFuAhead; ++u) {
                const int k = k0 + u;
                if (k >= nst) break;
                cp_async_wait<kFuAhead - 1>();                  // stage k has landed (this thread's copies)
                __syncthreads();                                // ... everyone's, and stage k - 1 is fully consumed
                issue(k + kFuAhead, colr[u][0], colr[u][1]);    // refills the buffer of stage k - 1
                load_cols(k + 2 * kFuAhead, colr[u][0], colr[u][1]);
                uint2 w[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) w[j] = wr[u][j];
                load_w(k + kFuAhead, wr[u]);
                const unsigned char* xs = xst + (size_t)(k % kFuStages) * kFuXBytes;
                const int2* rvs = reinterpret_cast<const int2*>(rvst + (size_t)(k % kFuStages) * kFuRvBytes);
                // phase a: the four tiles' products and {row, val} records (independent of each other)
                float c[4][4];
                int2 r[4][2];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
                    r[j][0] = r[j][1] = make_int2(0, 0);
                    if (4 * k + j < nt) {
                        uint32_t a[4];
                        const int row = (lane & 7) + ((lane >> 3) & 1) * 8, chunk = kb * 2 + (lane >> 4);
                        ldmatrix_x4(a, smem_u32(xs + j * kTileBytes + tile_off(row, chunk)));
                        mma_bf16_16816(c[j], a, w[j].x, w[j].y);
                        r[j][0] = rvs[j * 16 + g];
                        r[j][1] = rvs[j * 16 + g + 8];
                    }
                }
                // phase b: add val * product into this warp's column slice, slots 0-7 then 8-15 of every tile
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (4 * k + j >= nt) break;
                    const bool serial = s_rel[4 * k + j] < 0;   // plan flag: two entries of one step share a row
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const bool valid = r[j][half].y != 0;   // padding and zero-weight edges add nothing
                        const float v = __int_as_float(r[j][half].y), x0 = c[j][2 * half], x1 = c[j][2 * half + 1];
                        float2* p = reinterpret_cast<float2*>(myslice + (size_t)r[j][half].x * 32);
                        if (!serial) {
                            if (valid) {
                                float2 o = *p;
                                o.x = fmaf(x0, v, o.x); o.y = fmaf(x1, v, o.y);
                                *p = o;
                            }
                        } else {                                // one entry at a time
#pragma unroll 1
                            for (int i = 0; i < 8; ++i) {
                                if (valid && g == i) {
                                    float2 o = *p;
                                    o.x = fmaf(x0, v, o.x); o.y = fmaf(x1, v, o.y);
                                    *p = o;
                                }
                                __syncwarp();
                            }
                        }
                        __syncwarp();
                    }
                }
            }
        }
        cp_async_wait<0>();
        __syncthreads();

        // ---- flush: 8 consecutive threads write one 256-byte row
        const long long row0 = (long long)blk * FR;
        const int nrows = (int)min((long long)FR, A.N - row0);
        for (int idx = tid; idx < nrows * 8; idx += 256) {
            const int r = idx >> 3, w = idx & 7;
            const float4* src = reinterpret_cast<const float4*>(tile + (size_t)w * slice_stride + (size_t)r * 32);
            float4 lo = src[0], hi = src[1];
            const size_t o = (size_t)(row0 + r) * kFuWidth + 8 * w;
            if (!shared_item) {
                lo.x += bias_lo.x; lo.y += bias_lo.y; lo.z += bias_lo.z; lo.w += bias_lo.w;
                hi.x += bias_hi.x; hi.y += bias_hi.y; hi.z += bias_hi.z; hi.w += bias_hi.w;
                if constexpr (sizeof(OT) == 2) {
                    *reinterpret_cast<uint4*>(out + o) = make_uint4(pack_bf16x2(lo.x, lo.y), pack_bf16x2(lo.z, lo.w),
                                                                    pack_bf16x2(hi.x, hi.y), pack_bf16x2(hi.z, hi.w));
                } else {
                    float4* dst = reinterpret_cast<float4*>(out + o);
                    dst[0] = lo; dst[1] = hi;
                }
            } else {
                if constexpr (sizeof(OT) == 4) {                 // split blocks are never routed to a bf16 output
                    float* dst = reinterpret_cast<float*>(out) + o;
                    atomicAdd(dst, lo.x); atomicAdd(dst + 1, lo.y); atomicAdd(dst + 2, lo.z); atomicAdd(dst + 3, lo.w);
                    atomicAdd(dst + 4, hi.x); atomicAdd(dst + 5, hi.y); atomicAdd(dst + 6, hi.z); atomicAdd(dst + 7, hi.w);
                }
            }
        }
        // the barrier at the top of the loop separates this flush from the next item's zero fill
    }
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
inline size_t fused_ws_bytes(int64_t Rp) { return align_up((size_t)Rp * 256 * sizeof(uint2)) + align_up(sizeof(int32_t)); }

// W: (R', 4, 16, 16) blocks.  out: (N, 64) fp32 or bf16 (bf16 only when the list has no split blocks).
template <typename OT>
inline int launch_fused_rows(const rgcn_graph* g, bool backward, const float* W, const float* bias,
                             const __nv_bfloat16* src, OT* out, void* ws, cudaStream_t st) {
    const rgcn_fused& fl = backward ? g->fb : g->ff;
    const int n_items = (int)g->fuse_items[backward ? 1 : 0];
    const int n_split = (int)g->fuse_split[backward ? 1 : 0];
    const int FR = (int)g->fuse_rows;
    RGCN_REQUIRE(n_items > 0, RGCN_ERR_ARG, "fused rows: the plan has no usable fused list");
    RGCN_REQUIRE(n_split == 0 || sizeof(OT) == 4, RGCN_ERR_ARG, "fused rows: split blocks need an fp32 output");
    const size_t smem = fused_smem_bytes(FR);
    RGCN_REQUIRE(smem <= 227 * 1024, RGCN_ERR_UNSUPPORTED, "fused rows: fuse_rows %d needs %zu bytes of shared memory", FR, smem);
    Carver carve(ws);
    uint2* slices = carve.take<uint2>((size_t)g->num_rels * 256);
    int32_t* counter = carve.take<int32_t>(1);
    RGCN_CHECK_CUDA(cudaMemsetAsync(counter, 0, sizeof(int32_t), st));
    RGCN_LAUNCH(k_pack_wslice, grid_for(g->num_rels * 256, 256), 256, 0, st, W, (int)g->num_rels, backward ? 1 : 0, slices);
    if (n_split > 0) {
        if constexpr (sizeof(OT) == 4)
            RGCN_LAUNCH(k_fused_init_shared, n_items, 256, 0, st, fl.items, n_items, fl.blk_tile, FR, (long long)g->num_nodes,
                        bias, reinterpret_cast<float*>(out));
    }
    static size_t attr_smem[2] = {0, 0};
    size_t& cur = attr_smem[sizeof(OT) == 2 ? 1 : 0];
    if (cur < smem) {
        RGCN_CHECK_CUDA(cudaFuncSetAttribute(k_fused_rows<OT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cur = smem;
    }
    FusedArgs A{};
    A.fl = fl; A.n_items = n_items; A.fuse_rows = FR; A.N = (long long)g->num_nodes;
    A.wslice = slices; A.bias = bias; A.counter = counter;
    const int grid = n_items < kNumSMs ? n_items : kNumSMs;
    RGCN_LAUNCH(k_fused_rows<OT>, grid, 256, smem, st, A, src, out);
    return RGCN_OK;
}

}  // namespace rgcn
