// Gathered GEMM on the 5th-generation tensor cores for dense (and basis) weights over bf16 features:
//
//     out[scatter_e, :] += val_e * X[gather_e, :] W_p          for the edges e of relation p
//
// is, per relation, a real GEMM (edges_p x I x O — 105 TFLOP for the 200 M-edge 512 x 512 layer of SURVEY §8) whose A
// rows are gathered.  Reference: the dense / basis branch of torch_rgcn/layers.py:286-301 (NC) and :518-551 (LP), there an
// (R'N, I) sparse product followed by an einsum.
//
//   CTA        = one chunk of <= RGCN_CHUNK_EDGES edges of one relation (the plan's relation-major lists) x one tile of
//                NT <= 256 output columns; 128-edge M tiles, the inner dimension in 64-wide blocks.
//   producer   = warp 0.  Per k block: the 128 gathered rows come in through 32 TMA tile::gather4 copies (one per
//                lane, four rows each, 128-byte swizzle) — exactly the canonical K-major SWIZZLE_128B operand layout of
//                tcgen05.mma, so no thread touches the data; the NT x 64 weight tile (bf16, packed K-major by
//                k_pack_wt_bf16) through one 3-D TMA tile copy.  Both complete on the stage's "full" mbarrier.
//   MMA        = one elected lane of warp 1: four tcgen05.mma.cta_group::1.kind::f16 (M 128, N NT, K 16) per stage on
//                shared-memory descriptors, fp32 accumulators in tensor memory (NT columns x 128 lanes);
//                tcgen05.commit hands the stage back to the producer and, after the last k block, the accumulator
//                to the epilogue.
//   epilogue   = warps 2-5 (one TMEM lane quarter each): tcgen05.ld 32 columns at a time, scale by the edge's
//                normalisation weight, 16-byte fp32 reductions into the destination row (rows of different M tiles,
//                chunks and relations meet in `out`, which holds the bias beforehand).
//   occupancy  = 2 CTAs per SM (2 x 96 KB of stages, 2 x 256 TMEM columns): the epilogue of one overlaps the MMAs of
//                the other.
//
// The same kernel serves the forward (gather sources, W_p) and the feature gradient (gather the bf16 copy of
// grad_out by destination, W_p^T).
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "propagate_fused.cuh"

namespace rgcn {

constexpr int kUmM = 128;            // edges per M tile = TMEM lanes
constexpr int kUmK = 64;             // inner elements per stage = one 128-byte swizzle atom of bf16
constexpr int kUmMaxStages = 4;      // 2 stages: two CTAs per SM (default); 3-4: one CTA per SM with a deeper pipeline
constexpr int kUmThreads = 192;      // warp 0 producer, warp 1 MMA + TMEM owner, warps 2-5 epilogue
constexpr int kUmABytes = kUmM * kUmK * 2;

struct UmmaArgs {
    const int32_t* relptr; const int32_t* chunkptr; int num_rels;
    const int32_t* gather; const int32_t* scatter; const float* val;
    int I, O, NT;
    float* out;
    int maxc;                  // chunks of the largest relation: walk the chunks quantile by quantile; 0 = in list order
    int group;                 // relations per quantile sweep (their weight tiles stay L2-resident together)
};

// Which chunk a CTA works on.  `slot` counts chunk slots, the column / row tiles of one chunk being adjacent CTAs (their
// gathers meet in L2).  In list order, slot = chunk.  Quantile order (maxc > 0): relations are taken in groups of
// `group`; within a group, slot = qq * group + j is the chunk of the group's j-th relation that covers the qq-th of maxc
// quantiles of its (destination-sorted) edges.  The CTAs in flight at any time then scatter into one narrow range of
// destination rows, which stays L2-resident while all relations of the group add to it — the output streams through
// DRAM once per group instead of once per relation — and the group is small enough for its weight tiles to stay in L2
// as well.  Slots no chunk maps to exit.
__device__ __forceinline__ bool umma_chunk(const int32_t* __restrict__ relptr, const int32_t* __restrict__ chunkptr, int num_rels,
                                           int maxc, int group, int slot, int& p, int& e0, int& e1) {
    int c;
    if (maxc > 0) {
        const int per = maxc * group, g = slot / per, r = slot - g * per, qq = r / group;
        p = g * group + (r - qq * group);
        if (p >= num_rels) return false;
        const int first = chunkptr[p], np = chunkptr[p + 1] - first;
        if (np == 0) return false;
        const int q = (int)(((long long)qq * np + maxc - 1) / maxc);
        if (q >= np || (int)((long long)q * maxc / np) != qq) return false;
        c = first + q;
    } else {
        c = slot;
        if (c >= chunkptr[num_rels]) return false;
        int lo = 0, hi = num_rels;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (chunkptr[mid] <= c) lo = mid; else hi = mid;
        }
        p = lo;
    }
    e0 = relptr[p] + (c - chunkptr[p]) * RGCN_CHUNK_EDGES;
    e1 = min(relptr[p + 1], e0 + RGCN_CHUNK_EDGES);
    return true;
}

constexpr int kUmStgStride = 20;     // floats per staged row: 16 columns + 4 of padding (conflict-free 16-byte stores)
constexpr int kUmStgBytes = 32 * kUmStgStride * 4;

inline size_t umma_smem_bytes(int NT, int stages) {
    return 1024 + (size_t)stages * (kUmABytes + (size_t)NT * kUmK * 2) + 4 * kUmStgBytes;
}
inline size_t umma_wt_bytes(int64_t Rp, int I, int O) { return align_up((size_t)Rp * I * O * 2); }

// bf16 weights, K-major: wt[p][n][k] = W[p][k][n] (forward, transpose = 1) or W[p][n][k] (feature gradient: the inner
// dimension is the layer's output, transpose = 0; `rows` x `cols` is the shape of one W[p])
__global__ void k_pack_wt_bf16(const float* __restrict__ W, int64_t count, int rows, int cols, int transpose,
                               __nv_bfloat16* __restrict__ wt) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= count) return;
    if (!transpose) { wt[i] = __float2bfloat16(W[i]); return; }
    const int64_t per = (int64_t)rows * cols, p = i / per, rem = i - p * per;
    const int n = (int)(rem / rows), k = (int)(rem - (int64_t)n * rows);       // wt viewed as [p][cols][rows]
    wt[i] = __float2bfloat16(W[p * per + (int64_t)k * cols + n]);
}

__device__ __forceinline__ void tma_tile_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major operand tile with the 128-byte swizzle: 128-byte rows, 8-row groups 1024 bytes apart (cute UMMA::SmemDescriptor:
// start address and offsets in 16-byte units, version 1, layout type 2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int kUmStages>
__global__ void __launch_bounds__(kUmThreads, kUmStages == 2 ? 2 : 1)
k_gemm_umma(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, UmmaArgs A, int tmem_cols) {
    extern __shared__ unsigned char um_smem[];
    __shared__ __align__(8) unsigned long long bars[2 * kUmMaxStages + 2];
    __shared__ uint32_t tmem_slot;
    const int NT = A.NT, ntile = A.O / NT, n0 = ((int)blockIdx.x % ntile) * NT;
    int p, e0, e1;
    if (!umma_chunk(A.relptr, A.chunkptr, A.num_rels, A.maxc, A.group, (int)blockIdx.x / ntile, p, e0, e1)) return;
    const int mtiles = (e1 - e0 + kUmM - 1) / kUmM, kblocks = (A.I + kUmK - 1) / kUmK;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t base = ((uint32_t)__cvta_generic_to_shared(um_smem) + 1023u) & ~1023u;
    const uint32_t stage_bytes = kUmABytes + (uint32_t)NT * kUmK * 2;
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bars);
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (kUmStages + s); };
    const uint32_t tfull = bar0 + 8u * (2 * kUmStages), tempty = tfull + 8u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kUmStages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        mbar_init(tfull, 1);
        mbar_init(tempty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 0) {
        // ---- producer
        int it = 0;
        for (int mt = 0; mt < mtiles; ++mt) {
            int r[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int e = e0 + mt * kUmM + 4 * lane + j;
                r[j] = A.gather[e < e1 ? e : e0];                 // rows past the chunk's end: any valid row, never stored
            }
            for (int kb = 0; kb < kblocks; ++kb, ++it) {
                const int s = it % kUmStages;
                mbar_wait(empty(s), ((it / kUmStages) & 1) ^ 1);
                if (lane == 0) mbar_expect_tx(full(s), stage_bytes);
                __syncwarp();
                const uint32_t sa = base + (uint32_t)s * stage_bytes;
                tma_gather4(sa + (uint32_t)lane * 512u, &tmA, full(s), kb * kUmK, r[0], r[1], r[2], r[3]);
                if (lane == 0) tma_tile_3d(sa + kUmABytes, &tmB, full(s), kb * kUmK, n0, p);
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer
        // instruction descriptor (cute UMMA::InstrDescriptor): fp32 accumulate, bf16 x bf16, both operands K-major
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(kUmM >> 4) << 24);
        int it = 0;
        for (int mt = 0; mt < mtiles; ++mt) {
            mbar_wait(tempty, (mt & 1) ^ 1);                      // the epilogue has drained the previous M tile
            tc_fence_after();
            for (int kb = 0; kb < kblocks; ++kb, ++it) {
                const int s = it % kUmStages;
                mbar_wait(full(s), (it / kUmStages) & 1);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t sa = base + (uint32_t)s * stage_bytes;
                    const uint64_t ad = umma_desc_sw128(sa), bd = umma_desc_sw128(sa + kUmABytes);
#pragma unroll
                    for (int k = 0; k < kUmK / 16; ++k)           // 16 bf16 = 32 bytes = 2 descriptor units per step
                        tc_mma_bf16(tmem, ad + 2u * k, bd + 2u * k, idesc, (uint32_t)((kb | k) != 0));
                    tc_commit(empty(s));
                    if (kb == kblocks - 1) tc_commit(tfull);
                }
                __syncwarp();
            }
        }
    } else {
        // ---- epilogue: warp w reads the TMEM lane quarter w % 4 (lane = edge), scales by the edge weight, and turns the
        //      16-column slices through a padded staging tile so that a warp instruction adds 64 contiguous bytes to each of
        //      8 destination rows instead of 16 bytes to each of 32
        const int q = warp & 3;
        const uint32_t stg = base + (uint32_t)kUmStages * stage_bytes + (uint32_t)(warp - 2) * kUmStgBytes;
        for (int mt = 0; mt < mtiles; ++mt) {
            mbar_wait(tfull, mt & 1);
            tc_fence_after();
            const int e = e0 + mt * kUmM + q * 32 + lane;
            const bool valid = e < e1;
            const float v = valid ? A.val[e] : 0.f;
            const int drow = valid ? A.scatter[e] : -1;
            for (int c0 = 0; c0 < NT; c0 += 16) {
                uint32_t r[16];
                tc_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    sts128(stg + (uint32_t)(lane * kUmStgStride + 4 * j) * 4u,
                           make_float4(v * __uint_as_float(r[4 * j]), v * __uint_as_float(r[4 * j + 1]),
                                       v * __uint_as_float(r[4 * j + 2]), v * __uint_as_float(r[4 * j + 3])));
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = i * 8 + (lane >> 2), ch = lane & 3;
                    const int d = __shfl_sync(0xffffffffu, drow, row);
                    const float4 t = lds128(stg + (uint32_t)(row * kUmStgStride + 4 * ch) * 4u);
                    if (d >= 0) atomicAdd(reinterpret_cast<float4*>(A.out + (size_t)d * A.O + n0 + c0 + 4 * ch), t);
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// Weight gradient on the tensor cores:  gW_p[i, o] += sum_{e in p} (val_e X[src_e, i]) G[dst_e, o]
//
// Per relation a GEMM whose INNER dimension is the gathered one: the rows TMA brings in (one edge each, 128 bytes = 64
// consecutive inputs / outputs) are the k rows of MN-major operands — the canonical MN-major SWIZZLE_128B layout is
// exactly "8 k-rows of 128 contiguous MN bytes" per 1 KB group, 64-element MN atoms `LBO` apart.  A CTA owns a chunk of
// <= RGCN_CHUNK_EDGES edges of relation p, 128 rows (inputs) and NT <= 256 columns (outputs) of gW_p; stages of 64
// edges; warps 2-5 scale the X rows of a landed stage by the edge weights in place (generic-proxy writes, then
// fence.proxy.async) before the elected lane issues the stage's four MMAs; the 128 x NT accumulator leaves TMEM once per
// chunk and is added to gW_p with 64-byte row pieces.
// ---------------------------------------------------------------------------------------------------------------
struct UmmaWgradArgs {
    const int32_t* relptr; const int32_t* chunkptr; int num_rels;
    const int32_t* src; const int32_t* dst; const float* val;
    int I, O, NT;
    float* gW;                 // (R', I, O)
};

__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t atom_stride_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((atom_stride_bytes >> 4) & 0x3FFFu) << 16) | (64ull << 32) |
           (1ull << 46) | (2ull << 61);
}

// kMT = 128-row tiles of gW_p per CTA: 1 (two CTAs per SM, two stages) or 2 (a 256 x NT tile on two TMEM accumulators that
// share the G operand: a third less operand traffic per output; one CTA per SM, three stages)
// Producer warps: a stage is (2 kMT + NT / 64) atoms x 16 gather4 copies, issued one at a time per warp, so kMT = 2 runs
// two producer warps (warps 0 .. kProd - 1; MMA warp kProd; scale / epilogue warps kProd + 1 .. kProd + 4).
template <int kMT>
__global__ void __launch_bounds__((kMT + 5) * 32, kMT == 1 ? 2 : 1)
k_wgrad_umma(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG, UmmaWgradArgs A, int tmem_cols) {
    constexpr int kStages = kMT == 1 ? 2 : 3, kE = 64, kProd = kMT;                   // edges per stage
    extern __shared__ unsigned char um_smem[];
    __shared__ __align__(8) unsigned long long bars[3 * kStages + 1];
    __shared__ uint32_t tmem_slot;
    const int NT = A.NT, mtiles = A.I / (kUmM * kMT), ntiles = mtiles * (A.O / NT), tile = (int)blockIdx.x % ntiles;
    const int m0 = (tile % mtiles) * kUmM * kMT, n0 = (tile / mtiles) * NT;
    int p, e0, e1;
    if (!umma_chunk(A.relptr, A.chunkptr, A.num_rels, 0, 1, (int)blockIdx.x / ntiles, p, e0, e1)) return;
    const int nstages = (e1 - e0 + kE - 1) / kE, natoms_b = NT / 64;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t base = ((uint32_t)__cvta_generic_to_shared(um_smem) + 1023u) & ~1023u;
    constexpr uint32_t kXBytes = kUmABytes * kMT;          // X tile of a stage: 2 kMT atoms of 64 edges x 128 bytes
    const uint32_t stage_bytes = kXBytes + (uint32_t)NT * 128u;
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bars);
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (kStages + s); };
    auto scaled = [&](int s) { return bar0 + 8u * (2 * kStages + s); };
    const uint32_t tfull = bar0 + 8u * (3 * kStages);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); mbar_init(scaled(s), 128); }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kProd) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp < kProd) {
        // ---- producers: copy i of a stage = (atom i / 16, quad of four edges i % 16), dealt over the producer lanes
        const int quad = lane & 15, first = (int)threadIdx.x >> 4, natoms = 2 * kMT + natoms_b;
        for (int ks = 0; ks < nstages; ++ks) {
            const int s = ks % kStages;
            int rs[4], rd[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int e = e0 + ks * kE + 4 * quad + j, ee = e < e1 ? e : e0;   // past the end: a valid row, weight 0
                rs[j] = A.src[ee]; rd[j] = A.dst[ee];
            }
            mbar_wait(empty(s), ((ks / kStages) & 1) ^ 1);
            if (threadIdx.x == 0) mbar_expect_tx(full(s), stage_bytes);
            __syncwarp();
            const uint32_t sa = base + (uint32_t)s * stage_bytes;
            for (int a = first; a < natoms; a += 2 * kProd) {
                const uint32_t dst = sa + (uint32_t)a * 8192u + (uint32_t)quad * 512u;   // X atoms first, G atoms behind them
                if (a < 2 * kMT) tma_gather4(dst, &tmX, full(s), m0 + a * 64, rs[0], rs[1], rs[2], rs[3]);
                else tma_gather4(dst, &tmG, full(s), n0 + (a - 2 * kMT) * 64, rd[0], rd[1], rd[2], rd[3]);
            }
        }
    } else if (warp == kProd) {
        // ---- MMA issuer: both operands MN-major (bits 15 / 16 of the instruction descriptor)
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(NT >> 3) << 17) |
                               ((uint32_t)(kUmM >> 4) << 24);
        for (int ks = 0; ks < nstages; ++ks) {
            const int s = ks % kStages;
            mbar_wait(scaled(s), (ks / kStages) & 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t sa = base + (uint32_t)s * stage_bytes;
#pragma unroll
                for (int k = 0; k < kE / 16; ++k) {               // 16 edges = two 8-row groups = 2048 bytes per step
                    const uint64_t bd = umma_desc_mn_sw128(sa + kXBytes + 2048u * k, 8192u);
#pragma unroll
                    for (int mt = 0; mt < kMT; ++mt)
                        tc_mma_bf16(tmem + (uint32_t)(mt * NT), umma_desc_mn_sw128(sa + (uint32_t)mt * kUmABytes + 2048u * k, 8192u), bd,
                                    idesc, (uint32_t)((ks | k) != 0));
                }
                tc_commit(empty(s));
                if (ks == nstages - 1) tc_commit(tfull);
            }
            __syncwarp();
        }
    } else {
        // ---- warps 2-5: scale the X rows of each landed stage by the edge weights, then the epilogue
        const int t = threadIdx.x - (kProd + 1) * 32;              // 0 .. 127
        for (int ks = 0; ks < nstages; ++ks) {
            const int s = ks % kStages;
            mbar_wait(full(s), (ks / kStages) & 1);
            const uint32_t sa = base + (uint32_t)s * stage_bytes;
#pragma unroll
            for (int i = 0; i < 8 * kMT; ++i) {
                const int idx = t + 128 * i;                      // 16-byte piece of the X tile
                const int row = (idx >> 3) & 63;                  // edge within the stage (both atoms hold rows 0 .. 63)
                const int e = e0 + ks * kE + row;
                const float v = e < e1 ? A.val[e] : 0.f;
                uint4 w = lds128u(sa + (uint32_t)idx * 16u);
                uint32_t* wp = reinterpret_cast<uint32_t*>(&w);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float a, b;
                    unpack_bf16x2(wp[j], a, b);
                    wp[j] = pack_bf16x2(a * v, b * v);
                }
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sa + (uint32_t)idx * 16u), "r"(w.x), "r"(w.y), "r"(w.z),
                             "r"(w.w) : "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(scaled(s));
        }
        const int q = warp & 3;
        const uint32_t stg = base + (uint32_t)kStages * stage_bytes + (uint32_t)(warp - kProd - 1) * kUmStgBytes;
        mbar_wait(tfull, 0);
        tc_fence_after();
        for (int cc = 0; cc < kMT * NT; cc += 16) {
            const int mt = cc / NT, c0 = cc - mt * NT;
            float* gw = A.gW + ((size_t)p * A.I + m0 + mt * kUmM + q * 32) * A.O + n0;
            uint32_t r[16];
            tc_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)cc, r);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                sts128(stg + (uint32_t)(lane * kUmStgStride + 4 * j) * 4u,
                       make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                   __uint_as_float(r[4 * j + 3])));
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = i * 8 + (lane >> 2), ch = lane & 3;
                const float4 v4 = lds128(stg + (uint32_t)(row * kUmStgStride + 4 * ch) * 4u);
                atomicAdd(reinterpret_cast<float4*>(gw + (size_t)row * A.O + c0 + 4 * ch), v4);
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kProd)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
}

// column tile: the largest multiple of 16 that divides O and fits one MMA (N <= 256)
inline int umma_col_tile(int O) {
    for (int nt = 256; nt >= 16; nt -= 16)
        if (O % nt == 0) return nt;
    return 0;
}

inline bool umma_shape_supported(int I, int O) {
    const char* e = getenv("RGCN_UMMA");
    if (e && e[0] == '0') return false;
    return I >= 64 && O >= 64 && I % 64 == 0 && O % 64 == 0;   // whole 64-element k blocks in both directions
}

// `wt` = bf16 K-major weights from k_pack_wt_bf16: (R', O, I); X = gathered bf16 matrix (N, I); out (N, O) pre-set.
// max_rel_edges > 0 (and `scatter` sorted within a relation, i.e. the forward): quantile order of the chunks
inline int launch_gemm_umma(UmmaArgs A, const __nv_bfloat16* X, int64_t N, const __nv_bfloat16* wt, int chunks, cudaStream_t st,
                            int64_t max_rel_edges = 0) {
    rb_encode_fn enc = rb_encoder();
    RGCN_REQUIRE(enc, RGCN_ERR_CUDA, "tensor-core GEMM: cuTensorMapEncodeTiled is not available from this driver");
    const int NT = umma_col_tile(A.O);
    A.NT = NT;
    CUtensorMap ta, tb;
    memset(&ta, 0, sizeof(ta));
    memset(&tb, 0, sizeof(tb));
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)A.I, (cuuint64_t)N};
        const cuuint64_t gstr[1] = {(cuuint64_t)A.I * 2};
        const cuuint32_t box[2] = {(cuuint32_t)kUmK, 1};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = enc(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(X), gdim, gstr, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        RGCN_REQUIRE(r == CUDA_SUCCESS, RGCN_ERR_CUDA, "tensor-core GEMM: feature tensor map failed with %d", (int)r);
    }
    {
        const cuuint64_t gdim[3] = {(cuuint64_t)A.I, (cuuint64_t)A.O, (cuuint64_t)A.num_rels};
        const cuuint64_t gstr[2] = {(cuuint64_t)A.I * 2, (cuuint64_t)A.I * A.O * 2};
        const cuuint32_t box[3] = {(cuuint32_t)kUmK, (cuuint32_t)NT, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = enc(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16*>(wt), gdim, gstr, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        RGCN_REQUIRE(r == CUDA_SUCCESS, RGCN_ERR_CUDA, "tensor-core GEMM: weight tensor map failed with %d", (int)r);
    }
    int cols = 32;
    while (cols < NT) cols <<= 1;
    int stages = 2;
    if (const char* e = getenv("RGCN_UMMA_STAGES")) stages = atoi(e);
    stages = stages < 2 ? 2 : (stages > kUmMaxStages ? kUmMaxStages : stages);
    const size_t smem = umma_smem_bytes(NT, stages);
    // quantile order only while most slots hold a chunk (relation sizes within ~4x of the largest on average)
    int64_t slots = chunks;
    const int64_t maxc = (max_rel_edges + RGCN_CHUNK_EDGES - 1) / RGCN_CHUNK_EDGES;
    A.maxc = 0;
    A.group = 1;
    int64_t group = ((int64_t)32 << 20) / ((int64_t)A.I * A.O * 2);      // ~32 MB of bf16 weights per sweep
    if (const char* e = getenv("RGCN_UMMA_GROUP")) group = atoi(e);
    group = group < 1 ? 1 : (group > A.num_rels ? A.num_rels : group);
    const int64_t ngroups = (A.num_rels + group - 1) / group;
    // Quantile order is opt-in (RGCN_UMMA_ORDER=q): measured on the 200 M-edge 512 x 512 layer it LOSES to list order
    // (forward 273 ms vs 232 ms, any group size) — the sorted, streaming read-modify-write of list order is cheaper than
    // 150 concurrent chunks reducing into the same few thousand rows; kept for graphs whose output fits L2.
    const char* order = getenv("RGCN_UMMA_ORDER");
    if (order && order[0] == 'q' && maxc > 0 && maxc * ngroups * group <= 4 * (int64_t)chunks) {
        A.maxc = (int)maxc; A.group = (int)group; slots = maxc * ngroups * group;
    }
    const int64_t nblk = slots * (A.O / NT);
    RGCN_REQUIRE(nblk < (1ll << 31), RGCN_ERR_UNSUPPORTED, "tensor-core GEMM: %lld CTAs exceed the grid", (long long)nblk);
    dim3 grid((unsigned)nblk);
    auto go = [&](auto kernel) -> int {
        RGCN_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RGCN_LAUNCH(kernel, grid, kUmThreads, smem, st, ta, tb, A, cols);
        return RGCN_OK;
    };
    if (stages == 2) return go(k_gemm_umma<2>);
    if (stages == 3) return go(k_gemm_umma<3>);
    return go(k_gemm_umma<4>);
}

inline bool umma_wgrad_shape_supported(int I, int O) { return umma_shape_supported(I, O) && I % kUmM == 0; }

inline int umma_gather_map(CUtensorMap* tm, const __nv_bfloat16* M, int64_t rows, int width) {
    rb_encode_fn enc = rb_encoder();
    RGCN_REQUIRE(enc, RGCN_ERR_CUDA, "tensor-core GEMM: cuTensorMapEncodeTiled is not available from this driver");
    memset(tm, 0, sizeof(*tm));
    const cuuint64_t gdim[2] = {(cuuint64_t)width, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)width * 2};
    const cuuint32_t box[2] = {64, 1};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(M), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    RGCN_REQUIRE(r == CUDA_SUCCESS, RGCN_ERR_CUDA, "tensor-core GEMM: gather tensor map failed with %d", (int)r);
    return RGCN_OK;
}

// gW (R', I, O) must be zeroed by the caller; X (N, I) and Gb (N, O) are bf16
inline int launch_wgrad_umma(UmmaWgradArgs A, const __nv_bfloat16* X, const __nv_bfloat16* Gb, int64_t N, int chunks,
                             cudaStream_t st) {
    const int NT = umma_col_tile(A.O);
    A.NT = NT;
    CUtensorMap tx, tg;
    int rc = umma_gather_map(&tx, X, N, A.I);
    if (rc) return rc;
    rc = umma_gather_map(&tg, Gb, N, A.O);
    if (rc) return rc;
    int mt = (A.I % (2 * kUmM) == 0 && 2 * NT <= 512) ? 2 : 1;
    if (const char* e = getenv("RGCN_UMMA_WGRAD_MT")) mt = (atoi(e) == 2 && A.I % (2 * kUmM) == 0 && 2 * NT <= 512) ? 2 : 1;
    int cols = 32;
    while (cols < mt * NT) cols <<= 1;
    const size_t smem = 1024 + (size_t)(mt == 1 ? 2 : 3) * ((size_t)mt * kUmABytes + (size_t)NT * 128) + 4 * kUmStgBytes;
    const int64_t nblk = (int64_t)chunks * (A.I / (kUmM * mt)) * (A.O / NT);
    RGCN_REQUIRE(nblk < (1ll << 31), RGCN_ERR_UNSUPPORTED, "tensor-core weight gradient: %lld CTAs exceed the grid", (long long)nblk);
    dim3 grid((unsigned)nblk);
    auto go = [&](auto kernel) -> int {
        RGCN_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RGCN_LAUNCH(kernel, grid, (mt + 5) * 32, smem, st, tx, tg, A, cols);
        return RGCN_OK;
    };
    return mt == 2 ? go(k_wgrad_umma<2>) : go(k_wgrad_umma<1>);
}

}  // namespace rgcn
