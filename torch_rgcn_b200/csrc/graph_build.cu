// Integer side of the path: triple augmentation, adjacency stacking, degree normalisation and the
// sorted edge lists ("graph plan") that the propagation kernels walk.
//
// Reference behaviour restated here (never its code): torch_rgcn/utils.py:71-97 (sum_sparse),
// :100-141 (inverse / self-loop triples), :143-166 (stack_matrices), :168-196 (block_diag) and the
// normalisation glue in torch_rgcn/layers.py:255-273 / :490-510.
//
// Sorting uses cub::DeviceRadixSort / DeviceScan from the CUDA toolkit (header-only, compiled into
// this library); everything else is hand-written.  All of this is O(nnz) integer work done once per
// graph (NC) or once per forward (LP).
#include <cub/cub.cuh>
#include "common.cuh"

using namespace rgcn;

namespace {

constexpr int kBlock = 256;

// ------------------------------------------------------------------------------------------
// helper kernels (utils.py equivalents)
// ------------------------------------------------------------------------------------------
__global__ void k_add_inverse_and_self(const int64_t* __restrict__ t, int64_t E, int64_t N, int64_t R,
                                       int64_t* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t total = 2 * E + N;
    if (i >= total) return;
    int64_t s, p, o;
    if (i < E) {
        s = t[3 * i]; p = t[3 * i + 1]; o = t[3 * i + 2];
    } else if (i < 2 * E) {
        int64_t j = i - E;
        s = t[3 * j + 2]; p = t[3 * j + 1] + R; o = t[3 * j];
    } else {
        s = o = i - 2 * E; p = 2 * R;
    }
    out[3 * i] = s; out[3 * i + 1] = p; out[3 * i + 2] = o;
}

__global__ void k_generate_inverses(const int64_t* __restrict__ t, int64_t E, int64_t R, int64_t* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= E) return;
    out[3 * i] = t[3 * i + 2]; out[3 * i + 1] = t[3 * i + 1] + R; out[3 * i + 2] = t[3 * i];
}

__global__ void k_lp_triples_plus(const int64_t* __restrict__ t, int64_t E, int64_t R,
                                  const int64_t* __restrict__ self_nodes, int64_t n_self, int64_t* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t total = 3 * E + n_self;
    if (i >= total) return;
    int64_t s, p, o;
    if (i < E) {
        s = t[3 * i]; p = t[3 * i + 1]; o = t[3 * i + 2];
    } else if (i < 2 * E) {
        int64_t j = i - E;
        s = t[3 * j + 2]; p = t[3 * j + 1] + R; o = t[3 * j];
    } else if (i < 3 * E) {                       // the reference appends the triples a second time (utils.py:124)
        int64_t j = i - 2 * E;
        s = t[3 * j]; p = t[3 * j + 1]; o = t[3 * j + 2];
    } else {
        s = o = self_nodes[i - 3 * E]; p = 2 * R;
    }
    out[3 * i] = s; out[3 * i + 1] = p; out[3 * i + 2] = o;
}

__global__ void k_init_bounds(int64_t* b) { b[0] = INT64_MIN; b[1] = INT64_MIN; }

__global__ void k_stack_matrices(const int64_t* __restrict__ t, int64_t nnz, int64_t N, int vertical,
                                 int64_t* __restrict__ out, int64_t* bounds) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    int64_t fr = t[3 * i], off = t[3 * i + 1] * N, to = t[3 * i + 2];
    if (vertical) fr += off; else to += off;
    out[2 * i] = fr; out[2 * i + 1] = to;
    if (bounds) {
        atomicMax(reinterpret_cast<long long*>(bounds), (long long)fr);
        atomicMax(reinterpret_cast<long long*>(bounds + 1), (long long)to);
    }
}

__global__ void k_table_add(const int64_t* __restrict__ idx, const float* __restrict__ v, int64_t nnz, int sel,
                            float* table) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < nnz) atomicAdd(table + idx[2 * i + sel], v[i]);
}

__global__ void k_table_gather(const int64_t* __restrict__ idx, int64_t nnz, int sel, const float* __restrict__ table,
                               float* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < nnz) out[i] = table[idx[2 * i + sel]];
}

__global__ void k_block_diag(const float* __restrict__ b, int64_t R, int64_t nb, int64_t bi, int64_t bo,
                             float* __restrict__ out) {
    int64_t I = nb * bi, O = nb * bo;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= R * I * O) return;
    int64_t y = i % O, x = (i / O) % I, r = i / (I * O);
    int64_t kb = x / bi;
    out[i] = (y / bo == kb) ? b[((r * nb + kb) * bi + x % bi) * bo + y % bo] : 0.f;
}

// ------------------------------------------------------------------------------------------
// plan build
// ------------------------------------------------------------------------------------------
enum Ordering { ORD_DST = 0, ORD_SRC = 1, ORD_REL = 2 };

// key layouts, bit-packed (nb = bits of a node id, rb = bits of a relation id):
//   DST  s | p | o      SRC  o | p | s      REL  p | o | s   (gathers of X[o] walk rows in order)
// Only the two upper fields are sorted on (radix passes over bits [nb, 2 nb + rb)): rows and their (row, relation)
// segments must be contiguous, the order inside a segment is free (the stable sort keeps the caller's edge order).
__global__ void k_make_keys(const int64_t* __restrict__ t, int64_t nnz, int64_t N, int64_t Rp, int ord, int nb, int rb,
                            uint64_t* __restrict__ keys, int32_t* __restrict__ idx, int32_t* status) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    int64_t s = t[3 * e], p = t[3 * e + 1], o = t[3 * e + 2];
    if (s < 0 || s >= N || o < 0 || o >= N || p < 0 || p >= Rp) {
        if (status) atomicAdd(status, 1);
        s = p = o = 0;                              // keep the walk in bounds; the caller raises on status != 0
    }
    uint64_t k;
    if (ord == ORD_DST) k = ((((uint64_t)s << rb) | (uint64_t)p) << nb) | (uint64_t)o;
    else if (ord == ORD_SRC) k = ((((uint64_t)o << rb) | (uint64_t)p) << nb) | (uint64_t)s;
    else k = ((((uint64_t)p << nb) | (uint64_t)o) << nb) | (uint64_t)s;
    keys[e] = k;
    idx[e] = (int32_t)e;
}

// sorted keys -> index arrays + row pointer.  `a` = row id (s / o / p), written only through rowptr.
__global__ void k_decode(const uint64_t* __restrict__ keys, int64_t nnz, int ord, int nb, int rb,
                         int64_t nrows, int32_t* __restrict__ rowptr, int32_t* __restrict__ c0,
                         int32_t* __restrict__ c1) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    const uint64_t k = keys[e], nmask = (1ull << nb) - 1, rmask = (1ull << rb) - 1;
    int64_t row, prev;
    c0[e] = (int32_t)(k & nmask);                // REL: dst s; DST / SRC: the other endpoint
    if (ord == ORD_REL) {
        c1[e] = (int32_t)((k >> nb) & nmask);    // src o
        row = (int64_t)(k >> (2 * nb));
        prev = e ? (int64_t)(keys[e - 1] >> (2 * nb)) : -1;
    } else {
        c1[e] = (int32_t)((k >> nb) & rmask);    // relation
        row = (int64_t)(k >> (nb + rb));
        prev = e ? (int64_t)(keys[e - 1] >> (nb + rb)) : -1;
    }
    for (int64_t r = prev + 1; r <= row; ++r) rowptr[r] = (int32_t)e;
    if (e == nnz - 1)
        for (int64_t r = row + 1; r <= nrows; ++r) rowptr[r] = (int32_t)nnz;
}

// segment = run of equal upper key part in the sorted list: key >> shift when shift >= 0 ((s,p) for DST, (o,p) for
// SRC of the bit-packed plan keys), else key / N (mixed-radix keys of the fused row-block lists)
__device__ __forceinline__ uint64_t seg_of(uint64_t k, int64_t N, int shift) { return shift >= 0 ? k >> shift : k / (uint64_t)N; }

__global__ void k_seg_flags(const uint64_t* __restrict__ keys, int64_t nnz, int64_t N, int shift, int32_t* __restrict__ flag) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    flag[e] = (e == 0 || seg_of(keys[e], N, shift) != seg_of(keys[e - 1], N, shift)) ? 1 : 0;
}

__global__ void k_seg_bounds(const uint64_t* __restrict__ keys, int64_t nnz, int64_t N, int shift,
                             const int32_t* __restrict__ segid, int32_t* __restrict__ starts,
                             int32_t* __restrict__ ends) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    uint64_t k = seg_of(keys[e], N, shift);
    int32_t sgm = segid[e] - 1;
    if (e == 0 || seg_of(keys[e - 1], N, shift) != k) starts[sgm] = (int32_t)e;
    if (e == nnz - 1 || seg_of(keys[e + 1], N, shift) != k) ends[sgm] = (int32_t)e + 1;
}

// count of the segment each edge belongs to, scattered back to the caller's edge order
__global__ void k_seg_count_scatter(int64_t nnz, const int32_t* __restrict__ segid, const int32_t* __restrict__ starts,
                                    const int32_t* __restrict__ ends, const int32_t* __restrict__ perm,
                                    int32_t* __restrict__ cnt_orig) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    int32_t sgm = segid[e] - 1;
    cnt_orig[perm[e]] = ends[sgm] - starts[sgm];
}

// val = 1 / sums, with the horizontal-mode permutation cat([sums[n:2n], sums[:n], sums[-i:]])
// (layers.py:271-273 / :509-510).  IEEE division, like torch's vals / sums.
__global__ void k_edge_values(int64_t nnz, const int32_t* __restrict__ cnt, int norm, int64_t n, int64_t i,
                              float* __restrict__ val) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    int64_t src = e;
    if (norm == RGCN_NORM_COL_SWAPPED) {
        if (e < n) src = e + n;
        else if (e < 2 * n) src = e - n;
        else src = nnz - i + (e - 2 * n);
    }
    val[e] = __fdiv_rn(1.0f, (float)cnt[src]);
}

// inv[perm[e]] = e : position of every caller-order edge in the sorted list
// out[e] = val[perm[e]] and inv[perm[e]] = e in one pass over the sorted order
__global__ void k_gather_val_invert(int64_t nnz, const int32_t* __restrict__ perm, const float* __restrict__ val,
                                    float* __restrict__ out, int32_t* __restrict__ inv) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    const int32_t o = perm[e];
    out[e] = val[o];
    inv[o] = (int32_t)e;
}

__global__ void k_gather_slots(int64_t nnz, const int32_t* __restrict__ perm, const int32_t* __restrict__ inv_d,
                               const int32_t* __restrict__ inv_s, int32_t* __restrict__ dslot,
                               int32_t* __restrict__ sslot, const float* __restrict__ val, float* __restrict__ oval) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    int32_t o = perm[e];
    dslot[e] = inv_d[o];
    sslot[e] = inv_s[o];
    oval[e] = val[o];
}

// chunkptr[p] = number of RGCN_CHUNK_EDGES-sized chunks of relations < p (R' is small: one thread scans);
// status[6] = the largest relation in edges
__global__ void k_chunkptr(const int32_t* __restrict__ relptr, int64_t Rp, int32_t* __restrict__ chunkptr,
                           int32_t* __restrict__ status) {
    if (blockIdx.x || threadIdx.x) return;
    int32_t acc = 0, longest = 0;
    for (int64_t p = 0; p < Rp; ++p) {
        chunkptr[p] = acc;
        const int32_t n = relptr[p + 1] - relptr[p];
        longest = n > longest ? n : longest;
        acc += (n + RGCN_CHUNK_EDGES - 1) / RGCN_CHUNK_EDGES;
    }
    chunkptr[Rp] = acc;
    status[6] = longest;
}



// ---- super-tiling (L2-resident message ring) ----------------------------------------------------------
// key = (tile * R' + p) * N + a,  a = tile-side endpoint, tile = rowptr[a] / tile_edges
__global__ void k_make_tile_keys(const int64_t* __restrict__ t, int64_t nnz, int64_t N, int64_t Rp, int backward,
                                 const int32_t* __restrict__ rowptr, int64_t tile_edges,
                                 uint64_t* __restrict__ keys, int32_t* __restrict__ idx) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    int64_t s = t[3 * e], p = t[3 * e + 1], o = t[3 * e + 2];
    if (s < 0 || s >= N || o < 0 || o >= N || p < 0 || p >= Rp) s = p = o = 0;
    const int64_t a = backward ? o : s;
    const uint64_t tile = (uint64_t)(rowptr[a] / tile_edges);
    keys[e] = (tile * Rp + p) * N + a;
    idx[e] = (int32_t)e;
}

__global__ void k_decode_tiles(const uint64_t* __restrict__ keys, const int32_t* __restrict__ perm,
                               const int64_t* __restrict__ t, int64_t nnz, int64_t N, int64_t Rp, int backward,
                               const int32_t* __restrict__ inv, const float* __restrict__ val,
                               int32_t* __restrict__ row, int32_t* __restrict__ col, int32_t* __restrict__ rel,
                               int32_t* __restrict__ slot, float* __restrict__ oval) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    const uint64_t k = keys[e];
    const int32_t orig = perm[e];
    int64_t s = t[3 * (int64_t)orig], p = t[3 * (int64_t)orig + 1], o = t[3 * (int64_t)orig + 2];
    if (s < 0 || s >= N || o < 0 || o >= N || p < 0 || p >= Rp) s = p = o = 0;
    row[e] = (int32_t)(k % N);
    col[e] = (int32_t)(backward ? s : o);
    rel[e] = (int32_t)((k / N) % Rp);
    slot[e] = inv[orig];
    oval[e] = val[orig];
}

// tilerow[k] = first row whose tile id is >= k; tilerow[T] = N
__global__ void k_tile_rows(const int32_t* __restrict__ rowptr, int64_t N, int64_t tile_edges, int64_t T,
                            int32_t* __restrict__ tilerow) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= N) return;
    auto tile_of = [&](int64_t x) { int64_t k = rowptr[x] / tile_edges; return k < T ? k : T - 1; };   // trailing empty rows
    const int64_t tile = tile_of(r);
    const int64_t prev = r ? tile_of(r - 1) : -1;
    for (int64_t k = prev + 1; k <= tile; ++k) tilerow[k] = (int32_t)r;
    if (r == N - 1)
        for (int64_t k = tile + 1; k <= T; ++k) tilerow[k] = (int32_t)N;
}

// per-tile capacity (atomicMax) and the work-queue prefix over steps:
// step j = spans of tile j (j < T) + row blocks of tile j - lag (lag <= j < T + lag)
__global__ void k_tile_steps(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ tilerow, int64_t T,
                             int64_t lag, int32_t* __restrict__ stepcnt, int32_t* cap) {
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j > T + lag) return;
    int32_t n = 0;
    if (j < T) {
        const int32_t edges = rowptr[tilerow[j + 1]] - rowptr[tilerow[j]];
        n += (edges + RGCN_SPAN_EDGES - 1) / RGCN_SPAN_EDGES;
        atomicMax(cap, edges);
    }
    if (j >= lag && j < T + lag) {
        const int32_t rows = tilerow[j - lag + 1] - tilerow[j - lag];
        n += (rows + RGCN_TILE_ROWS_PER_ITEM - 1) / RGCN_TILE_ROWS_PER_ITEM;
    }
    stepcnt[j] = (j < T + lag) ? n : 0;
}

// slotneed[k] = row blocks of all tiles k' < k with k' % depth == k % depth (one thread per ring slot)
__global__ void k_slot_need(const int32_t* __restrict__ tilerow, int64_t T, int depth, int32_t* __restrict__ slotneed) {
    const int s = threadIdx.x;
    if (blockIdx.x || s >= depth) return;
    int32_t acc = 0;
    for (int64_t k = s; k < T; k += depth) {
        slotneed[k] = acc;
        acc += (tilerow[k + 1] - tilerow[k] + RGCN_TILE_ROWS_PER_ITEM - 1) / RGCN_TILE_ROWS_PER_ITEM;
    }
}

// work-queue item table: item q of step j is a transform span of tile j (first) or a row block of tile j - lag
__global__ void k_fill_items(const int32_t* __restrict__ stepptr, const int32_t* __restrict__ tilerow,
                             const int32_t* __restrict__ rowptr, const int32_t* __restrict__ slotneed,
                             int64_t T, int64_t lag, int64_t bound, rgcn_tile_item* __restrict__ items) {
    int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= bound || q >= stepptr[T + lag]) return;
    int64_t lo = 0, hi = T + lag;                   // stepptr[lo] <= q < stepptr[hi]
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if (stepptr[mid] <= q) lo = mid; else hi = mid;
    }
    const int local = (int)(q - stepptr[lo]);
    int n1 = 0, t0 = 0, t1 = 0;
    if (lo < T) {
        t0 = rowptr[tilerow[lo]]; t1 = rowptr[tilerow[lo + 1]];
        n1 = (t1 - t0 + RGCN_SPAN_EDGES - 1) / RGCN_SPAN_EDGES;
    }
    rgcn_tile_item it;
    it.pad = 0; it.a = 0;
    if (local < n1) {
        const int e0 = t0 + local * RGCN_SPAN_EDGES;
        it.kind = 0; it.tile = (int)lo; it.b = e0; it.c = min(t1, e0 + RGCN_SPAN_EDGES) - e0;
        it.slot_bias = t0;
        it.need = slotneed[lo];
    } else {
        const int64_t k = lo - lag;
        const int r0 = tilerow[k] + (local - n1) * RGCN_TILE_ROWS_PER_ITEM;
        const int k0 = rowptr[tilerow[k]], k1 = rowptr[tilerow[k + 1]];
        it.kind = 1; it.tile = (int)k; it.a = r0; it.b = min(tilerow[k + 1], r0 + RGCN_TILE_ROWS_PER_ITEM); it.c = 0;
        it.slot_bias = k0;
        it.need = (k1 - k0 + RGCN_SPAN_EDGES - 1) / RGCN_SPAN_EDGES;
    }
    items[q] = it;
}

// ---- fused row-block lists (rgcn_fused) ----------------------------------------------------------------
// key = ((block * R' + p) * 2 + (a & 1)) * N + a,  a = block-side endpoint, block = a / fuse_rows: a (block, relation)
// run lists its even rows first, then its odd rows
__global__ void k_make_block_keys(const int64_t* __restrict__ t, int64_t nnz, int64_t N, int64_t Rp, int backward,
                                  int64_t fuse_rows, uint64_t* __restrict__ keys, int32_t* __restrict__ idx) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    int64_t s = t[3 * e], p = t[3 * e + 1], o = t[3 * e + 2];
    if (s < 0 || s >= N || o < 0 || o >= N || p < 0 || p >= Rp) s = p = o = 0;
    const int64_t a = backward ? o : s;
    keys[e] = (((uint64_t)(a / fuse_rows) * Rp + p) * 2 + (uint64_t)(a & 1)) * N + a;
    idx[e] = (int32_t)e;
}

// cnt[g] = 16-entry tiles of run g (runs = segments of equal key / (2 N)), 0 beyond the last run
__global__ void k_fused_run_tiles(int64_t nnz, const int32_t* __restrict__ segid, const int32_t* __restrict__ starts,
                                  const int32_t* __restrict__ ends, int32_t* __restrict__ cnt) {
    int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= nnz) return;
    cnt[g] = g < segid[nnz - 1] ? (ends[g] - starts[g] + RGCN_FUSE_TILE - 1) / RGCN_FUSE_TILE : 0;
}

// Sorted edge e is edge i = e - starts[g] of run g (n edges), which owns tiles tbase[g] .. tbase[g] + cnt[g] - 1.
// The run is dealt round-robin over its tiles: tile i % cnt[g], where it is the tile's w-th entry (w = i / cnt[g]) of
// m = ceil((n - tile) / cnt[g]).  Entries w < h = ceil(m / 2) take the even slots 2 w, the others the odd slots
// 2 (w - h) + 1: the slot pairs (2 q, 2 q + 1) that one shared-memory access phase of the kernel touches then hold
// an even and an odd row whenever the run has them (the row's parity picks the bank half of its accumulators), and
// the tile's entries fill slots 0 .. m - 1.  Record layout of a tile (RGCN_FUSE_REC_WORDS int32):
//   word 4 (s % 8) + 2 (s / 8)     = accumulator offset of slot s: 256 * local row + 64 * (local row & 1), plus the
//                                    entry's rank among the entries of the same row in this tile (0 .. 15)
//   word 4 (s % 8) + 2 (s / 8) + 1 = bits of the fp32 edge weight (0 = padding)
//   word 32 = relation of tile + RGCN_FUSE_AHEAD | (largest rank in the tile) << 24 (k_fused_headers)
//   word 33 = relation of the tile, word 34 = largest rank in the tile, word 35 = 0
// The first edge of a run also labels the run's tiles and, for the first run of a row block, the block's first tile.
__global__ void k_fused_scatter(const uint64_t* __restrict__ keys, const int32_t* __restrict__ perm,
                                const int64_t* __restrict__ t, int64_t nnz, int64_t N, int64_t Rp, int backward,
                                int64_t fuse_rows, int64_t NB, const int32_t* __restrict__ segid,
                                const int32_t* __restrict__ starts, const int32_t* __restrict__ ends,
                                const int32_t* __restrict__ tbase, const int32_t* __restrict__ cnt,
                                const float* __restrict__ val, int64_t cap, int32_t* __restrict__ col,
                                int32_t* __restrict__ rec, int32_t* __restrict__ blk_tile, int32_t* __restrict__ meta) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    const uint64_t k = keys[e];
    const int64_t a = (int64_t)(k % N);
    const uint64_t grp = k / N / 2;
    const int64_t p = (int64_t)(grp % Rp), blk = (int64_t)(grp / Rp);
    const int32_t g = segid[e] - 1;
    const int32_t orig = perm[e];
    int64_t s = t[3 * (int64_t)orig], pp = t[3 * (int64_t)orig + 1], o = t[3 * (int64_t)orig + 2];
    if (s < 0 || s >= N || o < 0 || o >= N || pp < 0 || pp >= Rp) s = o = 0;
    const int64_t first_tile = tbase[g], ntile = cnt[g], n = ends[g] - starts[g];
    const int64_t i = e - starts[g];
    const int64_t tile_i = i % ntile, w = i / ntile;
    const int64_t m = (n - tile_i + ntile - 1) / ntile, h = (m + 1) / 2;
    const int64_t slot = w < h ? 2 * w : 2 * (w - h) + 1;
    const int64_t tile = first_tile + tile_i;
    if ((tile + 1) * RGCN_FUSE_TILE <= cap) {
        col[tile * RGCN_FUSE_TILE + slot] = (int32_t)(backward ? s : o);
        const int64_t rloc = a - blk * fuse_rows;
        int32_t* r = rec + tile * RGCN_FUSE_REC_WORDS + 4 * (slot & 7) + 2 * (slot >> 3);
        // equal rows of a run are neighbours in the sorted list and land in the same tile every ntile places: the entry's
        // rank among the entries of its row in this tile (a tile has 16 slots, so at most 15)
        int rank = 0;
        while (rank < 15 && i >= (rank + 1) * ntile && (int64_t)(keys[e - (rank + 1) * ntile] % N) == a) ++rank;
        r[0] = (int32_t)(rloc * 256 + (rloc & 1) * 64 + rank);
        r[1] = __float_as_int(val[orig]);
        if (rank) atomicMax(rec + tile * RGCN_FUSE_REC_WORDS + 34, rank);
    }
    if (e == starts[g]) {
        for (int64_t q = first_tile; q < first_tile + ntile && (q + 1) * RGCN_FUSE_TILE <= cap; ++q)
            rec[q * RGCN_FUSE_REC_WORDS + 33] = (int32_t)p;
        const int64_t prev = e ? (int64_t)(keys[e - 1] / N / 2 / Rp) : -1;    // block of the previous run
        for (int64_t b = prev + 1; b <= blk; ++b) blk_tile[b] = (int32_t)first_tile;
    }
    if (e == nnz - 1) {
        const int64_t total = first_tile + ntile;
        for (int64_t b = blk + 1; b <= NB; ++b) blk_tile[b] = (int32_t)total;
        meta[1] = (int32_t)total;
        meta[2] = total * RGCN_FUSE_TILE > cap ? 1 : 0;
    }
}

// word 32 of every tile record: relation of the tile RGCN_FUSE_AHEAD places later (the kernel requests that tile's
// weight fragments while it works on this one) and the largest rank of the tile (0 = all rows different);
// meta[4] counts the tiles with shared rows
__global__ void k_fused_headers(int32_t* __restrict__ rec, int64_t cap, int32_t* __restrict__ meta) {
    int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t total = meta[1];
    if (q >= total || (q + 1) * RGCN_FUSE_TILE > cap) return;
    int64_t ahead = q + RGCN_FUSE_AHEAD;
    if (ahead >= total || (ahead + 1) * RGCN_FUSE_TILE > cap) ahead = q;
    const int32_t maxrank = rec[q * RGCN_FUSE_REC_WORDS + 34];
    rec[q * RGCN_FUSE_REC_WORDS + 32] = rec[ahead * RGCN_FUSE_REC_WORDS + 33] | (maxrank << 24);
    if (maxrank) atomicAdd(meta + 4, 1);
}

__global__ void k_fused_item_counts(const int32_t* __restrict__ blk_tile, int64_t NB, int item_tiles,
                                    int32_t* __restrict__ icnt) {
    int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b > NB) return;
    int32_t n = 0;
    if (b < NB) {
        n = (blk_tile[b + 1] - blk_tile[b] + item_tiles - 1) / item_tiles;
        if (n < 1) n = 1;                              // empty blocks still write their bias rows
    }
    icnt[b] = n;
}

__global__ void k_fused_fill_items(const int32_t* __restrict__ blk_tile, const int32_t* __restrict__ iptr, int64_t NB,
                                   int item_tiles, int64_t bound, int32_t* __restrict__ items, int32_t* __restrict__ meta) {
    int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= NB) return;
    const int32_t q0 = iptr[b], n = iptr[b + 1] - q0, t0 = blk_tile[b], t1 = blk_tile[b + 1];
    for (int32_t i = 0; i < n; ++i) {
        const int64_t q = (int64_t)q0 + i;
        if (q >= bound) break;
        const int32_t a = t0 + i * item_tiles;
        int32_t z = a + item_tiles;
        if (z > t1) z = t1;
        reinterpret_cast<int4*>(items)[q] = make_int4((int)b, a, z, n > 1 ? 1 : 0);
    }
    if (n > 1) atomicAdd(meta + 3, 1);
    if (b == NB - 1) meta[0] = iptr[NB];
}

// rows with more than RGCN_LONG_ROW edges (hubs) are listed so that kernels can process them cooperatively
// blockIdx.y picks the list: 0 destination-major (count[0]), 1 source-major (count[1])
__global__ void k_long_rows(const int32_t* __restrict__ d_rowptr, const int32_t* __restrict__ s_rowptr, int64_t N,
                            int32_t* __restrict__ d_list, int32_t* __restrict__ s_list, int32_t* count) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= N) return;
    const int32_t* rowptr = blockIdx.y ? s_rowptr : d_rowptr;
    int32_t* list = blockIdx.y ? s_list : d_list;
    if (rowptr[r + 1] - rowptr[r] > RGCN_LONG_ROW) list[atomicAdd(count + blockIdx.y, 1)] = (int32_t)r;
}

int bits_for(unsigned __int128 maxkey) {
    int b = 1;
    while (b < 64 && (maxkey >> b) != 0) ++b;
    return b;
}

struct BuildWs {
    uint64_t *k0, *k1;
    int32_t *i0, *i1, *flag, *segid, *starts, *ends, *cnt, *inv_d, *inv_s;
    void* cub;
    size_t cub_bytes;
    size_t total;
};

BuildWs carve_build(void* ws, int64_t nnz, int64_t ngroups = 0) {
    BuildWs b;
    size_t n = (size_t)(nnz > 0 ? nnz : 1);
    if ((size_t)ngroups + RGCN_MAX_RING_DEPTH + 2 > n) n = (size_t)ngroups + RGCN_MAX_RING_DEPTH + 2;   // scans over groups / steps reuse the int scratch
    size_t sort_bytes = 0, scan_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (uint64_t*)nullptr, (uint64_t*)nullptr, (int32_t*)nullptr,
                                    (int32_t*)nullptr, (int)n);
    cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, (int32_t*)nullptr, (int32_t*)nullptr, (int)n);
    Carver c(ws);
    b.k0 = c.take<uint64_t>(n); b.k1 = c.take<uint64_t>(n);
    b.i0 = c.take<int32_t>(n); b.i1 = c.take<int32_t>(n);
    b.flag = c.take<int32_t>(n); b.segid = c.take<int32_t>(n);
    b.starts = c.take<int32_t>(n); b.ends = c.take<int32_t>(n);
    b.cnt = c.take<int32_t>(n);
    b.inv_d = c.take<int32_t>(n); b.inv_s = c.take<int32_t>(n);
    b.cub_bytes = align_up(sort_bytes > scan_bytes ? sort_bytes : scan_bytes);
    b.cub = c.take<char>(b.cub_bytes);
    b.total = c.off;
    return b;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" int rgcn_add_inverse_and_self(const int64_t* triples, int64_t E, int64_t N, int64_t R, int64_t* out,
                                         rgcn_stream_t stream) {
    RGCN_REQUIRE(E >= 0 && N >= 0 && out && (triples || E == 0), RGCN_ERR_ARG, "rgcn_add_inverse_and_self: bad arguments");
    int64_t total = 2 * E + N;
    if (total == 0) return RGCN_OK;
    RGCN_LAUNCH(k_add_inverse_and_self, grid_for(total, kBlock), kBlock, 0, (cudaStream_t)stream, triples, E, N, R, out);
    return RGCN_OK;
}

extern "C" int rgcn_generate_inverses(const int64_t* triples, int64_t E, int64_t R, int64_t* out, rgcn_stream_t stream) {
    RGCN_REQUIRE(E >= 0 && (E == 0 || (triples && out)), RGCN_ERR_ARG, "rgcn_generate_inverses: bad arguments");
    if (E == 0) return RGCN_OK;
    RGCN_LAUNCH(k_generate_inverses, grid_for(E, kBlock), kBlock, 0, (cudaStream_t)stream, triples, E, R, out);
    return RGCN_OK;
}

extern "C" int rgcn_lp_triples_plus(const int64_t* triples, int64_t E, int64_t R, const int64_t* self_nodes,
                                    int64_t n_self, int64_t* out, rgcn_stream_t stream) {
    RGCN_REQUIRE(E >= 0 && n_self >= 0 && (E == 0 || triples) && (n_self == 0 || self_nodes), RGCN_ERR_ARG,
                 "rgcn_lp_triples_plus: bad arguments");
    int64_t total = 3 * E + n_self;
    if (total == 0) return RGCN_OK;
    RGCN_REQUIRE(out, RGCN_ERR_ARG, "rgcn_lp_triples_plus: out is NULL");
    RGCN_LAUNCH(k_lp_triples_plus, grid_for(total, kBlock), kBlock, 0, (cudaStream_t)stream, triples, E, R, self_nodes,
                n_self, out);
    return RGCN_OK;
}

extern "C" int rgcn_stack_matrices(const int64_t* triples, int64_t nnz, int64_t N, int64_t Rp, int vertical,
                                   int64_t* indices_out, int64_t* bounds_out, rgcn_stream_t stream) {
    (void)Rp;
    RGCN_REQUIRE(nnz >= 0 && (nnz == 0 || (triples && indices_out)), RGCN_ERR_ARG, "rgcn_stack_matrices: bad arguments");
    if (bounds_out) RGCN_LAUNCH(k_init_bounds, 1, 1, 0, (cudaStream_t)stream, bounds_out);
    if (nnz == 0) return RGCN_OK;
    RGCN_LAUNCH(k_stack_matrices, grid_for(nnz, kBlock), kBlock, 0, (cudaStream_t)stream, triples, nnz, N, vertical,
                indices_out, bounds_out);
    return RGCN_OK;
}

extern "C" int rgcn_sum_sparse(const int64_t* indices, const float* values, int64_t nnz, int64_t rows, int64_t cols,
                               int row_normalisation, float* table_ws, float* sums_out, rgcn_stream_t stream) {
    RGCN_REQUIRE(nnz >= 0 && rows >= 0 && cols >= 0, RGCN_ERR_ARG, "rgcn_sum_sparse: bad sizes");
    if (nnz == 0) return RGCN_OK;
    RGCN_REQUIRE(indices && values && table_ws && sums_out, RGCN_ERR_ARG, "rgcn_sum_sparse: NULL pointer");
    int sel = row_normalisation ? 0 : 1;
    int64_t len = row_normalisation ? rows : cols;
    RGCN_CHECK_CUDA(cudaMemsetAsync(table_ws, 0, (size_t)len * sizeof(float), (cudaStream_t)stream));
    RGCN_LAUNCH(k_table_add, grid_for(nnz, kBlock), kBlock, 0, (cudaStream_t)stream, indices, values, nnz, sel, table_ws);
    RGCN_LAUNCH(k_table_gather, grid_for(nnz, kBlock), kBlock, 0, (cudaStream_t)stream, indices, nnz, sel, table_ws,
                sums_out);
    return RGCN_OK;
}

extern "C" int rgcn_block_diag(const float* blocks, int64_t R, int64_t nb, int64_t bi, int64_t bo, float* out,
                               rgcn_stream_t stream) {
    RGCN_REQUIRE(R >= 0 && nb > 0 && bi > 0 && bo > 0, RGCN_ERR_ARG, "rgcn_block_diag: bad sizes");
    int64_t total = R * nb * bi * nb * bo;
    if (total == 0) return RGCN_OK;
    RGCN_REQUIRE(blocks && out, RGCN_ERR_ARG, "rgcn_block_diag: NULL pointer");
    RGCN_LAUNCH(k_block_diag, grid_for(total, kBlock), kBlock, 0, (cudaStream_t)stream, blocks, R, nb, bi, bo, out);
    return RGCN_OK;
}

static int64_t tile_groups(int64_t nnz, int64_t Rp, int64_t tile_edges) {
    if (tile_edges <= 0 || nnz <= 0) return 0;
    return ((nnz - 1) / tile_edges + 1) * Rp;
}

extern "C" int64_t rgcn_tile_items_bound(int64_t nnz, int64_t N, int64_t Rp, int64_t tile_edges) {
    if (tile_edges <= 0 || nnz <= 0) return 0;
    const int64_t T = (nnz - 1) / tile_edges + 1;
    (void)Rp;
    return nnz / RGCN_SPAN_EDGES + T + N / RGCN_TILE_ROWS_PER_ITEM + T + 2;
}

extern "C" int64_t rgcn_tile_steps_len(int64_t nnz, int64_t tile_edges, int64_t ring_depth) {
    if (tile_edges <= 0 || nnz <= 0) return 0;
    return (nnz - 1) / tile_edges + 1 + ring_depth / 2 + 1;
}

static int64_t fused_blocks(int64_t N, int64_t fuse_rows) { return fuse_rows > 0 ? (N + fuse_rows - 1) / fuse_rows : 0; }

// scans over (tile, relation) groups, queue steps and row blocks reuse the int scratch of the build
static int64_t scratch_items(int64_t nnz, int64_t N, int64_t Rp, int64_t tile_edges, int64_t fuse_rows) {
    const int64_t a = tile_groups(nnz, Rp, tile_edges), b = fused_blocks(N, fuse_rows) + 2;
    return a > b ? a : b;
}

extern "C" int64_t rgcn_fused_items_bound(int64_t N, int64_t fuse_rows, int64_t fuse_cap, int64_t item_tiles) {
    if (fuse_rows <= 0 || item_tiles <= 0) return 0;
    return fused_blocks(N, fuse_rows) + fuse_cap / RGCN_FUSE_TILE / item_tiles + 2;
}

extern "C" size_t rgcn_graph_workspace_bytes(int64_t nnz, int64_t N, int64_t Rp, int64_t tile_edges, int64_t fuse_rows) {
    return carve_build(nullptr, nnz, scratch_items(nnz, N, Rp, tile_edges, fuse_rows)).total;
}

extern "C" int rgcn_graph_build(const int64_t* triples, int64_t nnz, int64_t N, int64_t Rp, int norm,
                                int64_t n_general, int64_t n_self, const float* val_in, rgcn_graph* g, void* ws,
                                size_t ws_bytes, rgcn_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    RGCN_REQUIRE(g, RGCN_ERR_ARG, "rgcn_graph_build: graph is NULL");
    RGCN_REQUIRE(nnz >= 0 && N > 0 && Rp > 0, RGCN_ERR_ARG, "rgcn_graph_build: bad sizes nnz=%lld N=%lld R'=%lld",
                 (long long)nnz, (long long)N, (long long)Rp);
    RGCN_REQUIRE(nnz < (int64_t)INT32_MAX && N < (int64_t)INT32_MAX, RGCN_ERR_UNSUPPORTED,
                 "rgcn_graph_build: nnz and num_nodes must fit int32");
    const int nb = bits_for((unsigned __int128)(N > 1 ? N - 1 : 1)), rb = bits_for((unsigned __int128)(Rp > 1 ? Rp - 1 : 1));
    RGCN_REQUIRE(2 * nb + rb <= 63, RGCN_ERR_UNSUPPORTED, "rgcn_graph_build: (node, relation, node) does not fit a 64-bit sort key");
    RGCN_REQUIRE(norm == RGCN_NORM_ROW || norm == RGCN_NORM_COL_SWAPPED || norm == RGCN_NORM_EXPLICIT, RGCN_ERR_ARG,
                 "rgcn_graph_build: unknown normalisation %d", norm);
    if (norm == RGCN_NORM_COL_SWAPPED)
        RGCN_REQUIRE(n_general >= 0 && n_self >= 0 && 2 * n_general + n_self == nnz && (n_self > 0 || nnz == 0),
                     RGCN_ERR_ARG,
                     "rgcn_graph_build: horizontal permutation needs 2n+i == nnz (n=%lld i=%lld nnz=%lld)",
                     (long long)n_general, (long long)n_self, (long long)nnz);
    if (norm == RGCN_NORM_EXPLICIT) RGCN_REQUIRE(val_in || nnz == 0, RGCN_ERR_ARG, "rgcn_graph_build: val_in is NULL");
    RGCN_REQUIRE(g->d_rowptr && g->s_rowptr && g->r_relptr && g->r_chunkptr && g->status, RGCN_ERR_ARG,
                 "rgcn_graph_build: NULL plan array");
    if (nnz > 0)
        RGCN_REQUIRE(triples && g->d_src && g->d_rel && g->d_val && g->s_dst && g->s_rel && g->s_val && g->r_dst &&
                         g->r_src && g->r_val && g->r_dslot && g->r_sslot && g->val, RGCN_ERR_ARG, "rgcn_graph_build: NULL plan array");
    RGCN_REQUIRE(g->fuse_rows >= 0 && g->fuse_rows % RGCN_FUSE_TILE == 0 && g->fuse_rows < 65536, RGCN_ERR_ARG,
                 "rgcn_graph_build: fuse_rows must be a multiple of 16 below 65536");
    BuildWs b = carve_build(ws, nnz, scratch_items(nnz, N, Rp, g->tile_edges, g->fuse_rows));
    RGCN_REQUIRE(ws_bytes >= b.total && (ws || b.total == 0), RGCN_ERR_WORKSPACE,
                 "rgcn_graph_build: workspace %zu < %zu bytes", ws_bytes, b.total);
    g->num_nodes = N; g->num_rels = Rp; g->nnz = nnz;
    g->num_tiles = 0; g->tile_capacity = 0; g->num_long_dst = g->num_long_src = -1;
    g->fuse_items[0] = g->fuse_items[1] = 0; g->fuse_split[0] = g->fuse_split[1] = 0;
    g->fuse_tiles[0] = g->fuse_tiles[1] = 0;
    RGCN_REQUIRE(g->tile_edges >= 0, RGCN_ERR_ARG, "rgcn_graph_build: negative tile_edges");
    RGCN_REQUIRE(g->d_long && g->s_long, RGCN_ERR_ARG, "rgcn_graph_build: NULL long-row list");
    RGCN_CHECK_CUDA(cudaMemsetAsync(g->status, 0, 8 * sizeof(int32_t), stream));
    if (nnz == 0) {
        RGCN_CHECK_CUDA(cudaMemsetAsync(g->d_rowptr, 0, (size_t)(N + 1) * sizeof(int32_t), stream));
        RGCN_CHECK_CUDA(cudaMemsetAsync(g->s_rowptr, 0, (size_t)(N + 1) * sizeof(int32_t), stream));
        RGCN_CHECK_CUDA(cudaMemsetAsync(g->r_relptr, 0, (size_t)(Rp + 1) * sizeof(int32_t), stream));
        RGCN_CHECK_CUDA(cudaMemsetAsync(g->r_chunkptr, 0, (size_t)(Rp + 1) * sizeof(int32_t), stream));
        return RGCN_OK;
    }
    const int grid = grid_for(nnz, kBlock);
    if (norm == RGCN_NORM_EXPLICIT)
        RGCN_CHECK_CUDA(cudaMemcpyAsync(g->val, val_in, (size_t)nnz * sizeof(float), cudaMemcpyDeviceToDevice, stream));

    // the ordering whose segments give the normalisation counts goes first
    const int first = (norm == RGCN_NORM_COL_SWAPPED) ? ORD_SRC : ORD_DST;
    const int order[3] = {first, first == ORD_DST ? ORD_SRC : ORD_DST, ORD_REL};
    for (int step = 0; step < 3; ++step) {
        const int ord = order[step];
        RGCN_LAUNCH(k_make_keys, grid, kBlock, 0, stream, triples, nnz, N, Rp, ord, nb, rb, b.k0, b.i0,
                    step == 0 ? g->status : (int32_t*)nullptr);
        size_t cub_bytes = b.cub_bytes;
        const int key_bits = 2 * nb + rb, low = nb;                                    // the lowest field is not sorted on
        RGCN_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(b.cub, cub_bytes, b.k0, b.k1, b.i0, b.i1, (int)nnz, low, key_bits,
                                                        stream));
        rgcn::g_launches.fetch_add((key_bits - low + 7) / 8 + 1, std::memory_order_relaxed);
        int32_t *rowptr, *c0, *c1; float* oval; int64_t nrows;
        if (ord == ORD_DST) { rowptr = g->d_rowptr; c0 = g->d_src; c1 = g->d_rel; oval = g->d_val; nrows = N; }
        else if (ord == ORD_SRC) { rowptr = g->s_rowptr; c0 = g->s_dst; c1 = g->s_rel; oval = g->s_val; nrows = N; }
        else { rowptr = g->r_relptr; c0 = g->r_dst; c1 = g->r_src; oval = g->r_val; nrows = Rp; }
        RGCN_LAUNCH(k_decode, grid, kBlock, 0, stream, b.k1, nnz, ord, nb, rb, nrows, rowptr, c0, c1);
        if (step == 0 && norm != RGCN_NORM_EXPLICIT) {
            RGCN_LAUNCH(k_seg_flags, grid, kBlock, 0, stream, b.k1, nnz, N, nb, b.flag);
            cub_bytes = b.cub_bytes;
            RGCN_CHECK_CUDA(cub::DeviceScan::InclusiveSum(b.cub, cub_bytes, b.flag, b.segid, (int)nnz, stream));
            rgcn::g_launches.fetch_add(1, std::memory_order_relaxed);
            RGCN_LAUNCH(k_seg_bounds, grid, kBlock, 0, stream, b.k1, nnz, N, nb, b.segid, b.starts, b.ends);
            RGCN_LAUNCH(k_seg_count_scatter, grid, kBlock, 0, stream, nnz, b.segid, b.starts, b.ends, b.i1, b.cnt);
            RGCN_LAUNCH(k_edge_values, grid, kBlock, 0, stream, nnz, b.cnt, norm, n_general, n_self, g->val);
        }
        if (ord == ORD_DST) RGCN_LAUNCH(k_gather_val_invert, grid, kBlock, 0, stream, nnz, b.i1, g->val, oval, b.inv_d);
        else if (ord == ORD_SRC) RGCN_LAUNCH(k_gather_val_invert, grid, kBlock, 0, stream, nnz, b.i1, g->val, oval, b.inv_s);
        else {
            RGCN_LAUNCH(k_gather_slots, grid, kBlock, 0, stream, nnz, b.i1, b.inv_d, b.inv_s, g->r_dslot, g->r_sslot, g->val,
                        oval);
            RGCN_LAUNCH(k_chunkptr, 1, 32, 0, stream, g->r_relptr, Rp, g->r_chunkptr, g->status);
        }
    }
    RGCN_LAUNCH(k_long_rows, dim3(grid_for(N, kBlock), 2), kBlock, 0, stream, g->d_rowptr, g->s_rowptr, N, g->d_long, g->s_long,
                g->status + 4);
    if (g->tile_edges > 0) {
        const int64_t te = g->tile_edges;
        const int64_t T = (nnz - 1) / te + 1;
        const int64_t ngroups = T * Rp;
        const int64_t depth = g->ring_depth, lag = depth / 2;
        RGCN_REQUIRE(depth >= 2 && depth <= RGCN_MAX_RING_DEPTH, RGCN_ERR_ARG, "rgcn_graph_build: ring_depth %lld out of range",
                     (long long)depth);
        RGCN_REQUIRE(ngroups < (int64_t)INT32_MAX, RGCN_ERR_UNSUPPORTED, "rgcn_graph_build: too many (tile, relation) groups");
        g->num_tiles = T;
        const int tbits = bits_for((unsigned __int128)ngroups * (unsigned __int128)N);
        RGCN_REQUIRE(tbits <= 63, RGCN_ERR_UNSUPPORTED, "rgcn_graph_build: tile key does not fit 64 bits");
        for (int backward = 0; backward < 2; ++backward) {
            rgcn_tiling& tl = backward ? g->bt : g->ft;
            RGCN_REQUIRE(tl.tilerow && tl.row && tl.col && tl.rel && tl.slot && tl.val && tl.stepptr && tl.slotneed && tl.items,
                         RGCN_ERR_ARG, "rgcn_graph_build: NULL tiling array");
            const int32_t* rowptr = backward ? g->s_rowptr : g->d_rowptr;
            RGCN_LAUNCH(k_make_tile_keys, grid, kBlock, 0, stream, triples, nnz, N, Rp, backward, rowptr, te, b.k0, b.i0);
            size_t cub_bytes = b.cub_bytes;
            RGCN_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(b.cub, cub_bytes, b.k0, b.k1, b.i0, b.i1, (int)nnz, 0, tbits, stream));
            rgcn::g_launches.fetch_add((tbits + 7) / 8 + 1, std::memory_order_relaxed);
            RGCN_LAUNCH(k_decode_tiles, grid, kBlock, 0, stream, b.k1, b.i1, triples, nnz, N, Rp, backward,
                        backward ? b.inv_s : b.inv_d, g->val, tl.row, tl.col, tl.rel, tl.slot, tl.val);
            RGCN_LAUNCH(k_tile_rows, grid_for(N, kBlock), kBlock, 0, stream, rowptr, N, te, T, tl.tilerow);
            RGCN_LAUNCH(k_slot_need, 1, 64, 0, stream, tl.tilerow, T, (int)depth, tl.slotneed);
            RGCN_LAUNCH(k_tile_steps, grid_for(T + lag + 1, kBlock), kBlock, 0, stream, rowptr, tl.tilerow, T, lag, b.segid,
                        g->status + 1 + backward);
            cub_bytes = b.cub_bytes;
            RGCN_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(b.cub, cub_bytes, b.segid, tl.stepptr, (int)(T + lag + 1), stream));
            rgcn::g_launches.fetch_add(1, std::memory_order_relaxed);
            const int64_t bound = rgcn_tile_items_bound(nnz, N, Rp, te);
            RGCN_LAUNCH(k_fill_items, grid_for(bound, kBlock), kBlock, 0, stream, tl.stepptr, tl.tilerow, rowptr, tl.slotneed,
                        T, lag, bound, reinterpret_cast<rgcn_tile_item*>(tl.items));
        }
    }
    if (g->fuse_rows > 0) {
        const int64_t FR = g->fuse_rows, NB = fused_blocks(N, FR), cap = g->fuse_cap;
        const int item_tiles = (int)g->fuse_item_tiles;
        RGCN_REQUIRE(cap > 0 && cap % RGCN_FUSE_TILE == 0 && cap < (int64_t)INT32_MAX, RGCN_ERR_ARG,
                     "rgcn_graph_build: fuse_cap must be a positive multiple of 16 that fits int32");
        RGCN_REQUIRE(item_tiles >= 1 && item_tiles <= RGCN_FUSE_MAX_ITEM_TILES, RGCN_ERR_ARG,
                     "rgcn_graph_build: fuse_item_tiles %d out of range", item_tiles);
        RGCN_REQUIRE(Rp < (1 << 24), RGCN_ERR_UNSUPPORTED, "rgcn_graph_build: fused lists need fewer than 2^24 relations");
        const unsigned __int128 fkey = (unsigned __int128)NB * (unsigned __int128)Rp * 2 * (unsigned __int128)N;
        RGCN_REQUIRE((fkey >> 63) == 0, RGCN_ERR_UNSUPPORTED, "rgcn_graph_build: row-block key does not fit 64 bits");
        const int fbits = bits_for(fkey);
        const int64_t bound = rgcn_fused_items_bound(N, FR, cap, item_tiles);
        const int64_t cap_tiles = cap / RGCN_FUSE_TILE;
        RGCN_REQUIRE(g->fuse_dirs >= 1 && g->fuse_dirs <= 3, RGCN_ERR_ARG, "rgcn_graph_build: fuse_dirs must be 1, 2 or 3");
        for (int backward = 0; backward < 2; ++backward) {
            if (!((g->fuse_dirs >> backward) & 1)) continue;
            rgcn_fused& fl = backward ? g->fb : g->ff;
            RGCN_REQUIRE(fl.col && fl.rec && fl.blk_tile && fl.items && fl.meta, RGCN_ERR_ARG,
                         "rgcn_graph_build: NULL fused list array");
            RGCN_CHECK_CUDA(cudaMemsetAsync(fl.meta, 0, 8 * sizeof(int32_t), stream));
            RGCN_CHECK_CUDA(cudaMemsetAsync(fl.col, 0xFF, (size_t)cap * sizeof(int32_t), stream));
            RGCN_CHECK_CUDA(cudaMemsetAsync(fl.rec, 0, (size_t)cap_tiles * RGCN_FUSE_REC_WORDS * sizeof(int32_t), stream));
            RGCN_LAUNCH(k_make_block_keys, grid, kBlock, 0, stream, triples, nnz, N, Rp, backward, FR, b.k0, b.i0);
            size_t cub_bytes = b.cub_bytes;
            RGCN_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(b.cub, cub_bytes, b.k0, b.k1, b.i0, b.i1, (int)nnz, 0, fbits, stream));
            rgcn::g_launches.fetch_add((fbits + 7) / 8 + 1, std::memory_order_relaxed);
            // runs = (block, relation) segments of the sorted list: equal key / (2 N)
            RGCN_LAUNCH(k_seg_flags, grid, kBlock, 0, stream, b.k1, nnz, 2 * N, -1, b.flag);
            cub_bytes = b.cub_bytes;
            RGCN_CHECK_CUDA(cub::DeviceScan::InclusiveSum(b.cub, cub_bytes, b.flag, b.segid, (int)nnz, stream));
            RGCN_LAUNCH(k_seg_bounds, grid, kBlock, 0, stream, b.k1, nnz, 2 * N, -1, b.segid, b.starts, b.ends);
            RGCN_LAUNCH(k_fused_run_tiles, grid, kBlock, 0, stream, nnz, b.segid, b.starts, b.ends, b.cnt);
            cub_bytes = b.cub_bytes;
            RGCN_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(b.cub, cub_bytes, b.cnt, b.inv_d, (int)nnz, stream));
            rgcn::g_launches.fetch_add(2, std::memory_order_relaxed);
            RGCN_LAUNCH(k_fused_scatter, grid, kBlock, 0, stream, b.k1, b.i1, triples, nnz, N, Rp, backward, FR, NB,
                        b.segid, b.starts, b.ends, b.inv_d, b.cnt, g->val, cap, fl.col, fl.rec, fl.blk_tile, fl.meta);
            RGCN_LAUNCH(k_fused_headers, grid_for(cap_tiles, kBlock), kBlock, 0, stream, fl.rec, cap, fl.meta);
            // work items: every row block, split into pieces of at most item_tiles tiles
            RGCN_LAUNCH(k_fused_item_counts, grid_for(NB + 1, kBlock), kBlock, 0, stream, fl.blk_tile, NB, item_tiles, b.flag);
            cub_bytes = b.cub_bytes;
            RGCN_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(b.cub, cub_bytes, b.flag, b.segid, (int)(NB + 1), stream));
            rgcn::g_launches.fetch_add(1, std::memory_order_relaxed);
            RGCN_LAUNCH(k_fused_fill_items, grid_for(NB, kBlock), kBlock, 0, stream, fl.blk_tile, b.segid, NB, item_tiles,
                        bound, fl.items, fl.meta);
        }
    }
    return RGCN_OK;
}
