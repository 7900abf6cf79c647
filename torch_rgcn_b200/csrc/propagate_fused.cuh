// Fused row-block kernel for bf16 features and four 16x16 weight blocks (64 -> 64, the AM-shaped layer):
// ONE pass over the edges, no per-edge message ever leaves the SM.
//
// The two-phase kernels of propagate_mma.cuh write one bf16 message per edge to HBM and read it back in the row
// sum: 256 B per edge of round trip on top of the 128 B gather.  Here a CTA owns a block of `fuse_rows`
// consecutive output rows and keeps their fp32 sums in shared memory, so an edge costs its gather and nothing else.
//
//   lists      = rgcn_fused of the plan: the block's edges sorted by (relation, row parity, row), every
//                (block, relation) run dealt over whole 16-entry tiles -> one relation per MMA tile.  Tiles are
//                numbered block by block, so a CTA streams ONE contiguous range of tiles (a static, tile-balanced
//                cut of the work items) and switches accumulators when the tile index crosses an item boundary.
//   producers  = warps 8-11.  Per stage of 4 tiles: the 64 gathered rows arrive through TMA
//                (cp.async.bulk.tensor.2d tile::gather4, 128-byte swizzle, one instruction per four rows) or, in the
//                fallback build, through one cp.async.bulk per row; the stage's 4 tile records (576 B, row offsets,
//                edge weights, relation of the tile 8 places ahead) through one cp.async.bulk.  Everything
//                completes on the stage's "full" mbarrier; a stage is refilled when the four consumer warps that
//                read it have arrived on its "empty" mbarrier.  No thread ever computes a per-lane gather address.
//   consumers  = warps 0-7 = 4 column blocks x 2.  The pair (b, 0), (b, 1) owns weight block b = output columns
//                [16 b, 16 b + 16) and takes the stages in turn.  A stage has two phases:
//                  front: per tile one ldmatrix.x4 (the 16 inputs of block b of the 16 gathered rows), two
//                         mma.sync.m16n8k16 against register-resident B fragments, the tile record; afterwards the
//                         stage's slot holds nothing the warp still needs and goes back to the producers;
//                  back : for the two rows a lane holds, a 16-byte read-modify-write of the fp32 accumulators:
//                         acc[row][4 t .. 4 t + 3] += val * product (the B fragments are packed with permuted columns
//                         so that a lane's four values are consecutive).
//                While one warp of the pair runs back(s), the other runs front(s + 1); a 64-thread named barrier per
//                stage hands the accumulators over, so the read-modify-write chains of a column block stay serial (a
//                row that two tiles share is updated correctly) while every load / MMA latency hides behind them.
//                Column blocks are disjoint, so the accumulation needs no atomics and its order is fixed by the
//                plan: results are run-to-run deterministic on unsplit blocks.
//   weights    = the fragments of the tile RGCN_FUSE_AHEAD places ahead (the warp's next stage) are requested
//                (16 B per lane, L1/L2 resident table packed by k_pack_wfrag4) while the current tile is processed;
//                plain loads are not queued behind the gathers because those do not pass through the LSU.
//   banks      = a row's 64-byte slice of block b sits in bank half (b + row) % 2; the plan pairs an even and an
//                odd row in the two entries a quarter-warp touches, so the 16-byte accesses are conflict-free
//                whenever a run has both parities.
//   duplicates = tiles in which entries share a row (a (row, relation) segment longer than its run's tile count)
//                carry per-entry ranks from the plan and are accumulated rank by rank.
//   flush      = out[row] = sums (the accumulators start from the bias), 256-byte coalesced rows (fp32, or bf16 for a
//                bf16 feature gradient); items of a split (hub) block start from zero and add into rows pre-set by
//                k_fused_init_shared with 16-byte fp32 reductions.
//
// The same kernel serves the forward (gather X[o], W) and the feature gradient (gather the bf16 copy of
// grad_out[s], W^T) on the plan's ff / fb lists.
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "propagate_fast.cuh"
#include "propagate_mma.cuh"

namespace rgcn {

constexpr int kRbStageTiles = 4;                        // tiles per pipeline stage
constexpr int kRbRecBytes = RGCN_FUSE_REC_WORDS * 4;    // 144
constexpr int kRbStageRecBytes = kRbStageTiles * kRbRecBytes;
constexpr int kRbBlocks = 4;                            // 16-column weight blocks = consumer warps that read one stage
constexpr int kRbTurns = 2;                             // consumer warps per column block, taking the stages in turn
constexpr int kRbConsumers = kRbBlocks * kRbTurns;      // warps 0-7: warp = turn * 4 + block
constexpr int kRbProducers = 4;                         // warps 8-11
constexpr int kRbThreads = (kRbConsumers + kRbProducers) * 32;
constexpr int kRbWidth = 64;                            // columns one launch computes: 4 blocks of 16
constexpr int kRbItemWeight = 48;                       // cost of a work item's flush in tiles (CTA partition)
static_assert(kRbStageTiles * kRbTurns == RGCN_FUSE_AHEAD, "a tile record names the relation of the warp's next stage");

// gathered rows in shared memory: 128-byte rows with the TMA 128-byte swizzle, or 144-byte rows (linear copies)
template <bool kGather4> struct RbX {
    static constexpr int row_bytes = kGather4 ? 128 : 144;
    static constexpr int tile_bytes = 16 * row_bytes;
    static constexpr int stage_bytes = kRbStageTiles * tile_bytes;
};

struct RbArgs {
    const int32_t* col;        // rgcn_fused.col
    const int32_t* rec;        // rgcn_fused.rec
    const int32_t* items;      // rgcn_fused.items
    int n_items;               // host copy of meta[0]
    int total_tiles;           // host copy of meta[1]
    int fuse_rows;
    int nstage;                // pipeline depth
    long long N;
    const uint4* wfrag;        // [(p * 4 + b) * 32 + lane]: B fragments of the two n8 halves of block b
    const float* bias;         // initial value of the accumulators of unshared items, or NULL
    const unsigned char* src;  // the gathered bf16 matrix (N, 64), used by the cp.async.bulk fallback
    int item_lo, item_hi;      // work items [item_lo, item_hi) are processed (a row range of a sharded layer)
    void* peers[RGCN_MAX_PEERS];   // bf16 output: the exchange buffers of all ranks (peer-to-peer stores), see rgcn_params
    int n_peers;
    int ld;                    // row pitch of the gathered matrix and of `out` in elements (the layer width, 64 * groups)
    int col0;                  // first column of the 64-column group this launch computes
    int rel_stride;            // uint4 entries between consecutive relations in wfrag (blocks per relation * 32)
};

__host__ __device__ inline size_t rb_smem_bytes(int fuse_rows, int nstage, bool gather4) {
    const size_t xs = gather4 ? RbX<true>::stage_bytes : RbX<false>::stage_bytes;
    return 1024 /* alignment slack */ + (size_t)nstage * xs + (size_t)fuse_rows * 256 + (size_t)nstage * kRbStageRecBytes +
           (size_t)nstage * 16;
}

// ---- PTX wrappers ---------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done;
}
// Waits for the phase with the given parity.  try_wait suspends the warp in hardware for a bounded time, so the loop
// turns over slowly; a wait that outlasts ~2^24 attempts (seconds: a lost copy, a plan that disagrees with the kernel)
// traps instead of hanging the device or running on with incomplete data.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_test(bar, parity)) return;
#pragma unroll 1
    for (uint32_t spin = 0; !mbar_test(bar, parity); ++spin) {
        __nanosleep(20);
        if (spin > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int r0, int r1, int r2,
                                            int r3) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes "
                 "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// named barriers: 1 + turn = the four warps of one turn (flush), 3 = all consumers, 4 + block = the pair of a column block
__device__ __forceinline__ void named_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 lds128u(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ int lds32(uint32_t a) {
    int v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void mma_bf16_16816_z(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};\n"
                 : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}

// wfrag[(p * nb + b) * 32 + lane] = {b0, b1 of half 0, b0, b1 of half 1} for lane (g = lane / 4, t = lane % 4):
// half h multiplies by the 16 x 8 matrix whose column n is column 4 (n / 2) + 2 h + (n % 2) of block b, so that after
// the two MMAs a lane holds columns 4 t .. 4 t + 3 of rows g and g + 8.  transpose = 0: B[k][c] = W[k][c] (forward);
// 1: B[k][c] = W[c][k] (feature gradient).  A launch of k_rowblock uses the four blocks of one 64-column group.
__global__ void k_pack_wfrag4(const float* __restrict__ W, long long n_blocks, int transpose, uint4* __restrict__ frag) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n_blocks * 32) return;
    const int lane = (int)(i & 31);
    const int g = lane >> 2, t = lane & 3;
    const float* wb = W + (size_t)(i >> 5) * 256;
    auto at = [&](int k, int c) { return transpose ? wb[c * 16 + k] : wb[k * 16 + c]; };
    uint32_t r[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int c = 4 * (g >> 1) + 2 * h + (g & 1);
        r[2 * h] = pack_bf16x2(at(2 * t, c), at(2 * t + 1, c));
        r[2 * h + 1] = pack_bf16x2(at(2 * t + 8, c), at(2 * t + 9, c));
    }
    frag[i] = make_uint4(r[0], r[1], r[2], r[3]);
}

// rows of split blocks start from the bias (or zero): their items add partial sums with atomics
__global__ void k_fused_init_shared(const int32_t* __restrict__ items, int n_items, const int32_t* __restrict__ blk_tile,
                                    int fuse_rows, long long N, int width, const float* __restrict__ bias,
                                    float* __restrict__ out) {
    const int q = blockIdx.x;
    if (q >= n_items) return;
    const int4 it = __ldg(reinterpret_cast<const int4*>(items) + q);
    if (!it.w || it.y != __ldg(blk_tile + it.x)) return;        // only the first item of a split block
    const long long row0 = (long long)it.x * fuse_rows;
    const int nrows = (int)min((long long)fuse_rows, N - row0);
    for (int i = threadIdx.x; i < nrows * (width / 4); i += blockDim.x) {
        const int c4 = i % (width / 4);
        const float4 b = bias ? __ldg(reinterpret_cast<const float4*>(bias) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        reinterpret_cast<float4*>(out + (size_t)row0 * width)[i] = b;
    }
}

// kNarrow: the layer IS one 64-column group (64 -> 64, the headline shape): pitch, column offset and the relation stride
// of the weight fragments are compile-time constants.
template <typename OT, bool kGather4, bool kNarrow>
__global__ void __launch_bounds__(kRbThreads, 1) k_rowblock(const __grid_constant__ CUtensorMap tmap, RbArgs A_,
                                                            OT* __restrict__ out) {
    struct Geo {
        const RbArgs& a;
        __device__ __forceinline__ int ld() const { return kNarrow ? kRbWidth : a.ld; }
        __device__ __forceinline__ int col0() const { return kNarrow ? 0 : a.col0; }
        __device__ __forceinline__ int rel_stride() const { return kNarrow ? kRbBlocks * 32 : a.rel_stride; }
    };
    const RbArgs& A = A_;
    const Geo geo{A_};
    using XL = RbX<kGather4>;
    extern __shared__ unsigned char smem_rb[];
    const uint32_t base = (smem_u32(smem_rb) + 1023u) & ~1023u;
    const int NS = A.nstage;
    const uint32_t xs0 = base;                                          // NS stages of gathered rows (1024-aligned)
    const uint32_t acc0 = xs0 + (uint32_t)NS * XL::stage_bytes;         // fuse_rows x 256 B accumulators (256-aligned)
    const uint32_t rec0 = acc0 + (uint32_t)A.fuse_rows * 256u;          // NS stages of 4 tile records
    const uint32_t full0 = rec0 + (uint32_t)NS * kRbStageRecBytes;      // NS "full" barriers, then NS "empty" barriers
    const uint32_t empty0 = full0 + (uint32_t)NS * 8u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- this CTA's share: a contiguous range of work items holding ~1/gridDim of the work
    const int4* items = reinterpret_cast<const int4*>(A.items);
    // Cut by weight = tiles + kRbItemWeight * items: item i starts at weight (first tile of i) + kRbItemWeight * i, so both
    // long items and runs of empty items spread over the CTAs.  An item costs its flush whether or not it has edges
    // (640 rows x 256 B written from shared memory: measured ~5 us, the time of ~48 tiles) — with a weight of 1 the one CTA
    // that inherited the ~1000 edge-less row blocks of a row-sharded 512-wide layer ran 5 ms per launch.
    const int IL = A.item_lo, IH = A.item_hi;
    auto first_item_at = [&](long long w) {                             // first item of [IL, IH) with start weight >= w
        int lo = IL, hi = IH;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((long long)__ldg(&items[mid].y) + (long long)kRbItemWeight * mid < w) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    const long long w_lo = (long long)(IL < A.n_items ? __ldg(&items[IL].y) : A.total_tiles) + (long long)kRbItemWeight * IL;
    const long long w_hi = (long long)(IH < A.n_items ? __ldg(&items[IH].y) : A.total_tiles) + (long long)kRbItemWeight * IH;
    const int G = gridDim.x, c = blockIdx.x;
    const int i0 = c == 0 ? IL : first_item_at(w_lo + (w_hi - w_lo) * c / G);
    const int i1 = c == G - 1 ? IH : first_item_at(w_lo + (w_hi - w_lo) * (c + 1) / G);
    const int tile_begin = i0 < A.n_items ? __ldg(&items[i0].y) : A.total_tiles;
    const int tile_end = i1 < A.n_items ? __ldg(&items[i1].y) : A.total_tiles;
    const int n_stages = (tile_end - tile_begin + kRbStageTiles - 1) / kRbStageTiles;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, kRbBlocks);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= kRbConsumers) {
        // =========================== producers ===========================
        // Producer warp pw owns the pipeline slots == pw (mod NP); NP divides the pipeline depth, so a warp's
        // successive visits to a slot are exactly one lap apart and the parity waits below cannot alias.
        const int NP = (NS % 4 == 0) ? 4 : (NS % 3 == 0) ? 3 : (NS % 2 == 0) ? 2 : 1;
        const int pw = warp - kRbConsumers;
        if (pw >= NP) return;
        const int q = lane & 15;                                        // quad of four rows: tile q / 4 of the stage
        auto load_idx = [&](int s) {
            const long long tile = (long long)tile_begin + (long long)s * kRbStageTiles + (q >> 2);
            int4 v = make_int4(-1, -1, -1, -1);
            if (lane < 16 && tile < tile_end) v = __ldg(reinterpret_cast<const int4*>(A.col + tile * RGCN_FUSE_TILE) + (q & 3));
            return v;
        };
        int4 idx = load_idx(pw);
        int slot = pw;                                                  // pw < NP <= NS
        uint32_t par = 0;
        for (int s = pw; s < n_stages; s += NP) {
            const int4 cur = idx;
            idx = load_idx(s + NP);                                     // lands while this stage's slot frees up
            const int tile0 = tile_begin + s * kRbStageTiles;
            const int nt = min(kRbStageTiles, tile_end - tile0);
            mbar_wait(empty0 + 8 * slot, par ^ 1u);
            const uint32_t fb = full0 + 8 * slot;
            const uint32_t xdst = xs0 + (uint32_t)slot * XL::stage_bytes + (uint32_t)(q >> 2) * XL::tile_bytes;
            if constexpr (kGather4) {
                const bool go = lane < 16 && cur.x >= 0;                // padding fills a tile from its tail
                const unsigned m = __ballot_sync(0xffffffffu, go);
                if (lane == 0) {
                    mbar_expect_tx(fb, (uint32_t)__popc(m) * 512u + (uint32_t)nt * kRbRecBytes);
                    bulk_g2s(rec0 + (uint32_t)slot * kRbStageRecBytes, A.rec + (size_t)tile0 * RGCN_FUSE_REC_WORDS,
                             (uint32_t)nt * kRbRecBytes, fb);
                }
                __syncwarp();
                if (go) tma_gather4(xdst + (uint32_t)(q & 3) * 512u, &tmap, fb, geo.col0(), cur.x, cur.y, cur.z, cur.w);
            } else {
                const int rows[4] = {cur.x, cur.y, cur.z, cur.w};
                int n = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) n += (lane < 16 && rows[k] >= 0) ? 1 : 0;
                const int total = __reduce_add_sync(0xffffffffu, n);
                if (lane == 0) {
                    mbar_expect_tx(fb, (uint32_t)total * 128u + (uint32_t)nt * kRbRecBytes);
                    bulk_g2s(rec0 + (uint32_t)slot * kRbStageRecBytes, A.rec + (size_t)tile0 * RGCN_FUSE_REC_WORDS,
                             (uint32_t)nt * kRbRecBytes, fb);
                }
                __syncwarp();
                if (lane < 16) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (rows[k] >= 0)
                            bulk_g2s(xdst + (uint32_t)((q & 3) * 4 + k) * XL::row_bytes,
                                     A.src + ((size_t)rows[k] * geo.ld() + geo.col0()) * 2, 128u, fb);
                }
            }
            slot += NP;
            if (slot >= NS) { slot -= NS; par ^= 1u; }
        }
        return;
    }

    // =========================== consumers ===========================
    const int b = warp & (kRbBlocks - 1), turn = warp / kRbBlocks, g = lane >> 2, t = lane & 3;
    const int ctid = tid & (kRbBlocks * 32 - 1);                        // 0 .. 127 within the four warps of a turn
    const uint32_t lane_c = (uint32_t)((b << 6) | (t << 4));
    const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lchunk = 2 * b + (lane >> 4);
    const uint32_t ldm_off = kGather4 ? (uint32_t)(lrow * 128 + ((lchunk ^ (lrow & 7)) << 4))
                                      : (uint32_t)(lrow * 144 + lchunk * 16);
    const uint4* wmine = A.wfrag + (size_t)b * 32 + lane;
    const int FR = A.fuse_rows;

    // accumulator (re)initialisation and flush: 16 consecutive threads handle one 256-byte row (fp32 output) or 8
    // threads one 128-byte row (bf16 output, 16-byte stores)
    const int fc4 = ctid & 15;                                          // float4 column of this thread in a row
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (A.bias) bias4 = __ldg(reinterpret_cast<const float4*>(A.bias + geo.col0()) + fc4);
    auto acc_addr_c = [&](int r, int c4) {                              // un-swizzled float4 c4 of local row r
        return acc0 + (uint32_t)r * 256u + (uint32_t)((((c4 >> 2) ^ (r & 1)) << 6) | ((c4 & 3) << 4));
    };
    auto acc_addr = [&](int r) { return acc_addr_c(r, fc4); };
    int cur = i0;
    int4 it = cur < i1 ? __ldg(&items[cur]) : make_int4(0, 0x7fffffff, 0x7fffffff, 0);
    if (turn == 0)
        for (int r = ctid >> 4; r < FR; r += 8) sts128(acc_addr(r), it.w ? make_float4(0.f, 0.f, 0.f, 0.f) : bias4);
    named_sync(3, kRbConsumers * 32);

    // writes item `item` out and prepares the accumulators for an item that is shared or not; called by the four
    // warps of one turn while the other four only run front phases
    auto flush = [&](const int4& item, int nxt_shared) {
        named_sync(1 + turn, kRbBlocks * 32);                           // every column block has added its last tile
        const long long row0 = (long long)item.x * FR;
        const int nrows = (int)min((long long)FR, A.N - row0);
        const float4 init = nxt_shared ? make_float4(0.f, 0.f, 0.f, 0.f) : bias4;
        if constexpr (sizeof(OT) == 2) {
            // bf16 rows: 8 threads per row, two float4 columns each -> one 16-byte store per thread and destination;
            // with peer buffers the row goes to every rank's exchange buffer (NVLink peer-to-peer stores)
            const int c8 = ctid & 7;
            for (int r = ctid >> 3; r < FR; r += 16) {
                const uint32_t a0 = acc_addr_c(r, 2 * c8), a1 = acc_addr_c(r, 2 * c8 + 1);
                if (r < nrows && !item.w) {
                    const float4 v = lds128(a0), w = lds128(a1);
                    const uint4 pk = make_uint4(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w), pack_bf16x2(w.x, w.y),
                                                pack_bf16x2(w.z, w.w));
                    const size_t o = (size_t)(row0 + r) * geo.ld() + geo.col0() + 8 * c8;
                    if (A.n_peers == 0) {
                        *reinterpret_cast<uint4*>(out + o) = pk;
                    } else {
#pragma unroll 1
                        for (int q = 0; q < A.n_peers; ++q)
                            *reinterpret_cast<uint4*>(static_cast<OT*>(A.peers[q]) + o) = pk;
                    }
                }
            }
            named_sync(1 + turn, kRbBlocks * 32);                       // all rows read before they are re-initialised
            for (int r = ctid >> 4; r < FR; r += 8) sts128(acc_addr(r), init);
        } else {
            for (int r = ctid >> 4; r < FR; r += 8) {
                const uint32_t a = acc_addr(r);
                if (r < nrows) {
                    const float4 v = lds128(a);
                    const size_t o = (size_t)(row0 + r) * geo.ld() + geo.col0() + 4 * fc4;
                    if (!item.w) *reinterpret_cast<float4*>(out + o) = v;
                    else atomicAdd(reinterpret_cast<float4*>(out + o), v);
                }
                sts128(a, init);
            }
        }
        named_sync(1 + turn, kRbBlocks * 32);
    };
    auto next_item = [&]() {
        ++cur;
        it = cur < i1 ? __ldg(&items[cur]) : make_int4(0, 0x7fffffff, 0x7fffffff, 0);
    };

    // weight fragments of this warp's first stage; afterwards every tile requests those of the tile 8 places ahead
    uint4 wr[kRbStageTiles];
#pragma unroll
    for (int j = 0; j < kRbStageTiles; ++j) {
        const long long tile = (long long)tile_begin + turn * kRbStageTiles + j;
        const int rel = tile < tile_end ? __ldg(A.rec + tile * RGCN_FUSE_REC_WORDS + 33) : 0;
        wr[j] = __ldg(wmine + (size_t)rel * geo.rel_stride());
    }

    uint4 rv[kRbStageTiles];            // per tile {offset of row g | rank, val, offset of row g + 8 | rank, val}
    float d[kRbStageTiles][8];          // products: row g columns 4 t .. 4 t + 3, then row g + 8
    int hd[kRbStageTiles];              // header word: relation 8 tiles ahead | (largest rank in the tile) << 24

    int fslot = turn;                   // pipeline slot and phase parity of this warp's next front stage (NS >= 2)
    uint32_t fpar = 0;
    const unsigned char* wbytes = reinterpret_cast<const unsigned char*>(wmine);
    const uint32_t wrel_bytes = (uint32_t)geo.rel_stride() * 16u;
    auto front_tile = [&](uint32_t xs, uint32_t rs, int j) {
        rv[j] = lds128u(rs + j * kRbRecBytes + g * 16);
        hd[j] = lds32(rs + j * kRbRecBytes + 128);
        uint32_t a[4];
        ldmatrix_x4(a, xs + j * XL::tile_bytes);
        float d0[4], d1[4];
        mma_bf16_16816_z(d0, a, wr[j].x, wr[j].y);
        mma_bf16_16816_z(d1, a, wr[j].z, wr[j].w);
        d[j][0] = d0[0]; d[j][1] = d0[1]; d[j][2] = d1[0]; d[j][3] = d1[1];
        d[j][4] = d0[2]; d[j][5] = d0[3]; d[j][6] = d1[2]; d[j][7] = d1[3];
        wr[j] = __ldg(reinterpret_cast<const uint4*>(wbytes + (size_t)(uint32_t)(hd[j] & 0xffffff) * wrel_bytes));
    };
    auto front = [&](int s) {
        const int slot = fslot;
        const uint32_t par = fpar;
        fslot += kRbTurns;
        if (fslot >= NS) { fslot -= NS; fpar ^= 1u; }
        const int nt = min(kRbStageTiles, tile_end - (tile_begin + s * kRbStageTiles));
        mbar_wait(full0 + 8 * slot, par);
        const uint32_t xs = xs0 + (uint32_t)slot * XL::stage_bytes + ldm_off;
        const uint32_t rs = rec0 + (uint32_t)slot * kRbStageRecBytes;
        if (nt == kRbStageTiles) {
#pragma unroll
            for (int j = 0; j < kRbStageTiles; ++j) front_tile(xs, rs, j);
        } else {
#pragma unroll
            for (int j = 0; j < kRbStageTiles; ++j) {
                if (j < nt) {
                    front_tile(xs, rs, j);
                } else {
                    rv[j] = make_uint4(0u, 0u, 0u, 0u);
                    hd[j] = 0;
                }
            }
        }
        __syncwarp();
        // lane 0 releases the slot; the operands tie the arrive to the completion of this stage's shared-memory loads
        asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %1, 0;\n\t@p mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}"
                     ::"r"(empty0 + 8 * slot), "r"(lane), "r"(rv[0].x), "r"(rv[1].x), "r"(rv[2].x), "r"(rv[3].x),
                       "r"(hd[0]), "r"(hd[1]), "r"(hd[2]), "r"(hd[3]), "f"(d[0][7]), "f"(d[1][7]), "f"(d[2][7]), "f"(d[3][7])
                     : "memory");
    };
    // acc[row g / g + 8][4 t .. 4 t + 3] += val * product of one tile whose rows are all different
    auto rmw = [&](const uint4& r, const float (&dd)[8]) {
        const float v0 = __uint_as_float(r.y), v1 = __uint_as_float(r.w);
        const uint32_t p0 = acc0 + (r.x ^ lane_c), p1 = acc0 + (r.z ^ lane_c);
        float4 x0, x1;
        if (r.y) x0 = lds128(p0);
        if (r.w) x1 = lds128(p1);
        if (r.y) {
            x0.x = fmaf(v0, dd[0], x0.x); x0.y = fmaf(v0, dd[1], x0.y);
            x0.z = fmaf(v0, dd[2], x0.z); x0.w = fmaf(v0, dd[3], x0.w);
            sts128(p0, x0);
        }
        if (r.w) {
            x1.x = fmaf(v1, dd[4], x1.x); x1.y = fmaf(v1, dd[5], x1.y);
            x1.z = fmaf(v1, dd[6], x1.z); x1.w = fmaf(v1, dd[7], x1.w);
            sts128(p1, x1);
        }
    };
    // the same for a tile in which entries share rows: the plan ranks the entries of one row 0, 1, ... (low 4 bits of
    // the offset word); pass q adds the entries of rank q, whose rows are all different
    auto rmw_ranked = [&](const uint4& r, const float (&dd)[8], int maxrank) {
        const float v0 = __uint_as_float(r.y), v1 = __uint_as_float(r.w);
        const uint32_t p0 = acc0 + ((r.x & ~15u) ^ lane_c), p1 = acc0 + ((r.z & ~15u) ^ lane_c);
        const uint32_t k0 = r.x & 15u, k1 = r.z & 15u;
#pragma unroll 1
        for (uint32_t q = 0; q <= (uint32_t)maxrank; ++q) {
            const bool a0 = r.y && k0 == q, a1 = r.w && k1 == q;
            float4 x0, x1;
            if (a0) x0 = lds128(p0);
            if (a1) x1 = lds128(p1);
            if (a0) {
                x0.x = fmaf(v0, dd[0], x0.x); x0.y = fmaf(v0, dd[1], x0.y);
                x0.z = fmaf(v0, dd[2], x0.z); x0.w = fmaf(v0, dd[3], x0.w);
                sts128(p0, x0);
            }
            if (a1) {
                x1.x = fmaf(v1, dd[4], x1.x); x1.y = fmaf(v1, dd[5], x1.y);
                x1.z = fmaf(v1, dd[6], x1.z); x1.w = fmaf(v1, dd[7], x1.w);
                sts128(p1, x1);
            }
        }
    };
    auto back = [&](int s) {
        const int tile0 = tile_begin + s * kRbStageTiles;
        const int nt = min(kRbStageTiles, tile_end - tile0);
        // item boundaries before this stage were handled by the other turn
        while (it.z < tile0 && cur < i1) next_item();
        if (nt == kRbStageTiles && tile0 + kRbStageTiles <= it.z) {         // whole stage inside the current item
            if (((hd[0] | hd[1] | hd[2] | hd[3]) >> 24) == 0) {             // ... and no shared rows
#pragma unroll
                for (int j = 0; j < kRbStageTiles; ++j) rmw(rv[j], d[j]);
            } else {
#pragma unroll
                for (int j = 0; j < kRbStageTiles; ++j) {
                    if ((hd[j] >> 24) == 0) rmw(rv[j], d[j]); else rmw_ranked(rv[j], d[j], hd[j] >> 24);
                }
            }
            return;
        }
#pragma unroll
        for (int j = 0; j < kRbStageTiles; ++j) {
            if (j < nt) {
                const int tile = tile0 + j;
                while (tile == it.z && cur < i1) {                      // the tile opens the next work item
                    const int4 done = it;
                    next_item();
                    flush(done, it.w);
                }
                if ((hd[j] >> 24) == 0) rmw(rv[j], d[j]); else rmw_ranked(rv[j], d[j], hd[j] >> 24);
            }
        }
    };

    // stage s belongs to turn s % 2: between two pair barriers one warp runs back(s) and the other front(s + 1)
    int pair_bar = 4 + b;
    asm volatile("" : "+r"(pair_bar));                                  // keep the barrier id in a register
    auto pair_sync = [&]() { named_sync(pair_bar, kRbTurns * 32); };
    if (turn == 0) {
        if (n_stages > 0) front(0);
        for (int s = 0; s < n_stages; s += 2) {
            back(s);
            pair_sync();
            if (s + 1 < n_stages) {
                if (s + 2 < n_stages) front(s + 2);
                pair_sync();
            }
        }
    } else {
        for (int s = 1; s < n_stages; s += 2) {
            front(s);
            pair_sync();
            back(s);
            pair_sync();
        }
        if (n_stages & 1) pair_sync();
    }
    // the last item(s), empty blocks included: flushed by the turn that ran the last back phase
    if (turn == (n_stages > 0 ? (n_stages - 1) & 1 : 0)) {
        while (it.z < tile_end && cur < i1) next_item();
        while (cur < i1) {
            const int4 done = it;
            next_item();
            flush(done, it.w);
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
inline size_t fused_ws_bytes(int64_t Rp, int nb = 4) { return align_up((size_t)Rp * nb * 32 * sizeof(uint4)); }

typedef CUresult (*rb_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point query: the library does not link libcuda
inline rb_encode_fn rb_encoder() {
    static rb_encode_fn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<rb_encode_fn>(p);
    }();
    return fn;
}

// tuning knobs (read at every launch): RGCN_FUSED_TMA = gather4 (default) | bulk, RGCN_FUSED_STAGES = pipeline depth
struct RbTuning { bool gather4; int stages; };
inline RbTuning rb_tuning() {
    RbTuning r{true, 0};
    const char* e = getenv("RGCN_FUSED_TMA");
    if (e && e[0] == 'b') r.gather4 = false;
    e = getenv("RGCN_FUSED_STAGES");
    if (e) r.stages = atoi(e);
    return r;
}

// W: (R', nb, 16, 16) blocks, nb a multiple of 4; src (N, 16 nb) bf16; out: (N, 16 nb) fp32 or bf16 (bf16 only when the
// list has no split blocks).  The layer is nb / 4 independent 64-column layers over the same graph (block-diagonal
// weights): one launch of k_rowblock per 64-column group, all on the same lists.
template <typename OT>
inline int launch_fused_rows(const rgcn_graph* g, bool backward, const float* W, const float* bias,
                             const __nv_bfloat16* src, OT* out, void* ws, cudaStream_t st, int64_t row_lo = 0,
                             int64_t row_hi = -1, void* const* peers = nullptr, int n_peers = 0, int nb = 4) {
    const rgcn_fused& fl = backward ? g->fb : g->ff;
    const int n_items = (int)g->fuse_items[backward ? 1 : 0];
    const int n_split = (int)g->fuse_split[backward ? 1 : 0];
    const int total_tiles = (int)g->fuse_tiles[backward ? 1 : 0];
    const int FR = (int)g->fuse_rows;
    RGCN_REQUIRE(n_items > 0, RGCN_ERR_ARG, "fused rows: the plan has no usable fused list");
    RGCN_REQUIRE(n_split == 0 || sizeof(OT) == 4, RGCN_ERR_ARG, "fused rows: split blocks need an fp32 output");
    const RbTuning tune = rb_tuning();
    const size_t xstage = tune.gather4 ? RbX<true>::stage_bytes : RbX<false>::stage_bytes;
    const size_t fixed = rb_smem_bytes(FR, 0, tune.gather4);
    const size_t per_sm = 228 * 1024, reserve = 1024, stage = xstage + kRbStageRecBytes + 16;   // 1 KB per CTA is the system's
    RGCN_REQUIRE(fixed + 2 * stage <= per_sm - reserve, RGCN_ERR_UNSUPPORTED,
                 "fused rows: fuse_rows %d leaves no room for the gather pipeline", FR);
    const int ctas = 1;                                   // 12 warps x 168 registers: one CTA owns the SM
    const size_t limit = per_sm / ctas - reserve;
    int nstage = tune.stages > 0 ? tune.stages : (limit > fixed ? (int)((limit - fixed) / stage) : 0);
    if (nstage > 12) nstage = 12;
    if (nstage < 2) nstage = 2;
    if (nstage == 5 || nstage == 7 || nstage == 10 || nstage == 11) --nstage;      // keep >= 2 producer warps busy
    const size_t smem = rb_smem_bytes(FR, nstage, tune.gather4);
    RGCN_REQUIRE(smem <= per_sm - reserve, RGCN_ERR_UNSUPPORTED,
                 "fused rows: %d stages with fuse_rows %d need %zu bytes of shared memory", nstage, FR, smem);
    Carver carve(ws);
    RGCN_REQUIRE(nb >= 4 && nb % 4 == 0, RGCN_ERR_ARG, "fused rows: the number of 16x16 blocks must be a multiple of 4");
    const int width = 16 * nb;
    uint4* frag = carve.take<uint4>((size_t)g->num_rels * nb * 32);
    RGCN_LAUNCH(k_pack_wfrag4, grid_for(g->num_rels * nb * 32, 256), 256, 0, st, W, (long long)g->num_rels * nb,
                backward ? 1 : 0, frag);
    if (n_split > 0) {
        if constexpr (sizeof(OT) == 4)
            RGCN_LAUNCH(k_fused_init_shared, n_items, 256, 0, st, fl.items, n_items, fl.blk_tile, FR, (long long)g->num_nodes,
                        width, bias, reinterpret_cast<float*>(out));
    }
    CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    if (tune.gather4) {
        rb_encode_fn enc = rb_encoder();
        RGCN_REQUIRE(enc, RGCN_ERR_CUDA, "fused rows: cuTensorMapEncodeTiled is not available from this driver");
        const cuuint64_t gdim[2] = {(cuuint64_t)width, (cuuint64_t)g->num_nodes};
        const cuuint64_t gstr[1] = {(cuuint64_t)width * 2};
        const cuuint32_t box[2] = {(cuuint32_t)kRbWidth, 1};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(src), gdim, gstr, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        RGCN_REQUIRE(r == CUDA_SUCCESS, RGCN_ERR_CUDA, "fused rows: cuTensorMapEncodeTiled failed with %d", (int)r);
    }
    RbArgs A{};
    A.col = fl.col; A.rec = fl.rec; A.items = fl.items; A.n_items = n_items; A.total_tiles = total_tiles;
    A.fuse_rows = FR; A.nstage = nstage; A.N = (long long)g->num_nodes;
    A.wfrag = frag; A.bias = bias; A.src = reinterpret_cast<const unsigned char*>(src);
    A.ld = width; A.rel_stride = nb * 32;
    // output rows [row_lo, row_hi): the work items of those row blocks.  Items are listed block by block; a range cut at
    // block boundaries is a contiguous item range only if no block is split, which is what a row-sharded caller has.
    const int64_t NB = (g->num_nodes + FR - 1) / FR;
    if (row_hi < 0 || row_hi > g->num_nodes) row_hi = g->num_nodes;
    A.item_lo = 0; A.item_hi = n_items;
    RGCN_REQUIRE(n_peers >= 0 && n_peers <= RGCN_MAX_PEERS && (n_peers == 0 || (sizeof(OT) == 2 && n_split == 0)), RGCN_ERR_ARG,
                 "fused rows: peer stores need a bf16 output, unsplit blocks and at most %d peers", RGCN_MAX_PEERS);
    A.n_peers = n_peers;
    for (int q = 0; q < n_peers; ++q) A.peers[q] = peers[q];
    if (row_lo > 0 || row_hi < g->num_nodes) {
        RGCN_REQUIRE(n_split == 0 && n_items == NB, RGCN_ERR_UNSUPPORTED, "fused rows: a row range needs unsplit row blocks");
        A.item_lo = (int)(row_lo / FR);
        A.item_hi = (int)((row_hi + FR - 1) / FR);
    }
    int grid = kNumSMs * ctas;
    if (grid > A.item_hi - A.item_lo) grid = A.item_hi - A.item_lo;
    if (grid < 1) return RGCN_OK;
    auto go = [&](auto kernel) -> int {
        RGCN_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int cg = 0; cg < nb / 4; ++cg) {                           // one 64-column group per launch
            A.col0 = 64 * cg;
            A.wfrag = frag + (size_t)cg * 4 * 32;                       // blocks 4 cg .. 4 cg + 3 of every relation
            RGCN_LAUNCH(kernel, grid, kRbThreads, smem, st, tm, A, out);
        }
        return RGCN_OK;
    };
    const char* nw = getenv("RGCN_FUSED_NARROW");
    if (nb == 4 && !(nw && nw[0] == '0')) return tune.gather4 ? go(k_rowblock<OT, true, true>) : go(k_rowblock<OT, false, true>);
    return tune.gather4 ? go(k_rowblock<OT, true, false>) : go(k_rowblock<OT, false, false>);
}

}  // namespace rgcn
