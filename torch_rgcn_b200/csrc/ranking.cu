// Filtered ranking evaluation of the link-prediction models on the device.
//
// Reference behaviour restated here (never its code): utils/misc.py:60-110 `evaluate` scores, for every test triple,
// all num_nodes head (or tail) completions with the decoder, sets the scores of the other known true completions to
// -inf (`filter_scores`, :39-58, from the dictionaries of `generate_true_dict`, :29-37), and ranks the target:
//     rank = #{scores > true} + (#{scores == true} - 1) // 2 + 1.
// Upstream this materialises a (batch, num_nodes, 3) triple tensor and a (batch, num_nodes) score matrix per batch and
// re-runs the whole encoder for every batch.  Here the node embeddings are computed once by the caller, and
//   k_rank_queries : q_i = relations[p_i] * nodes[other_i], the target's score t_i
//   k_rank_count   : a register-tiled (128 x 128 x 8, 8 x 8 per thread) fp32 product Q X^T whose epilogue only COUNTS scores > t_i and == t_i
//                    (the score matrix is never written)
//   k_rank_filter  : recomputes the scores of the known true completions of each query (sorted key lists instead of
//                    Python dictionaries) and takes them out of the counts again
//   k_rank_finish  : the tie rule above.
// Every score is the same left-to-right fma chain over the embedding dimension, so a candidate's score is bit-identical
// wherever it is computed and the > / == counts are consistent.
#include <cub/cub.cuh>
#include "common.cuh"

using namespace rgcn;

namespace {

constexpr int BM = 128, BN = 128, BK = 8;

// (sbias[s] + pbias[p]) + obias[o], the association of layers.py:96
__device__ __forceinline__ float with_bias(float acc, const float* sb, const float* pb, const float* ob, long long s,
                                           long long p, long long o) {
    return sb ? acc + ((__ldg(sb + s) + __ldg(pb + p)) + __ldg(ob + o)) : acc;
}

__device__ __forceinline__ float dot_seq(const float* __restrict__ a, const float* __restrict__ b, int dim) {
    float acc = 0.f;
    for (int k = 0; k < dim; ++k) acc = fmaf(a[k], b[k], acc);
    return acc;
}

// head != 0: candidates replace the subject, other = object;  head == 0: candidates replace the object
__global__ void k_rank_queries(const int64_t* __restrict__ q, long long T, int head, const float* __restrict__ nodes,
                               long long N, const float* __restrict__ rel, long long R, int dim,
                               const float* __restrict__ sb, const float* __restrict__ pb, const float* __restrict__ ob,
                               float* __restrict__ Q, float* __restrict__ tscore, int64_t* __restrict__ qclean,
                               int32_t* status) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= T) return;
    long long s = q[3 * i], p = q[3 * i + 1], o = q[3 * i + 2];
    if (s < 0 || s >= N || o < 0 || o >= N || p < 0 || p >= R) {
        if (status) atomicAdd(status, 1);
        s = p = o = 0;                                 // keep the later kernels in bounds; the caller raises on status
    }
    qclean[3 * i] = s; qclean[3 * i + 1] = p; qclean[3 * i + 2] = o;
    const long long other = head ? o : s, target = head ? s : o;
    const float* xr = rel + (size_t)p * dim;
    const float* xo = nodes + (size_t)other * dim;
    const float* xt = nodes + (size_t)target * dim;
    float acc = 0.f;
    for (int k = 0; k < dim; ++k) {
        const float v = xr[k] * xo[k];
        Q[(size_t)i * dim + k] = v;
        acc = fmaf(v, xt[k], acc);
    }
    tscore[i] = with_bias(acc, sb, pb, ob, s, p, o);
}

// counts[2 i] += #{c : score(i, c) > t_i}, counts[2 i + 1] += #{c : score(i, c) == t_i}
//
// 128 x 128 (query x candidate) tiles, 256 threads, 8 x 8 accumulators per thread (two 4-wide groups 64 apart in each
// direction, so the LDS.128 of a half-warp are conflict-free), k-slabs of 8 double-buffered through registers: one
// __syncthreads per slab, 4 LDS.128 per 64 FFMA.  Each accumulator runs k = 0 .. dim-1 in order (no split-K), which
// keeps the score of a (query, candidate) pair bit-identical to dot_seq.
template <bool VEC>
__global__ void __launch_bounds__(256, 2) k_rank_count(const float* __restrict__ Q, const int64_t* __restrict__ q, long long T,
                                                       int head, const float* __restrict__ X, long long N, int dim,
                                                       const float* __restrict__ sb, const float* __restrict__ pb,
                                                       const float* __restrict__ ob, const float* __restrict__ tscore,
                                                       int32_t* __restrict__ counts) {
    __shared__ __align__(16) float Qs[2][BK][BM], Xs[2][BK][BN];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const long long i0 = (long long)blockIdx.y * BM;
    const int lr = tid >> 1, lk = (tid & 1) * 4;              // loader role: row lr of the tile, 4 consecutive k
    const long long qi = i0 + lr;
    const int nslab = (dim + BK - 1) / BK;
    float tq[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const long long i = i0 + (a >> 2) * 64 + ty * 4 + (a & 3);
        tq[a] = i < T ? __ldg(tscore + i) : 0.f;
    }
    int gt[8], eq[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) gt[a] = eq[a] = 0;

    auto fetch = [&](const float* __restrict__ base, long long row, long long rows, int k, float (&v)[4]) {
        if (VEC) {                                              // dim % 4 == 0 and 16-byte aligned rows
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < rows && k < dim) t = __ldg(reinterpret_cast<const float4*>(base + (size_t)row * dim + k));
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = (row < rows && k + e < dim) ? __ldg(base + (size_t)row * dim + k + e) : 0.f;
        }
    };

    for (long long c0 = (long long)blockIdx.x * BN; c0 < N; c0 += (long long)gridDim.x * BN) {
        float acc[8][8];
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
        const long long ci = c0 + lr;
        float qv[4], xv[4];
        fetch(Q, qi, T, lk, qv);
        fetch(X, ci, N, lk, xv);
        __syncthreads();                                        // the previous tile's readers are done with buffer 0
#pragma unroll
        for (int e = 0; e < 4; ++e) { Qs[0][lk + e][lr] = qv[e]; Xs[0][lk + e][lr] = xv[e]; }
        __syncthreads();
        for (int sl = 0; sl < nslab; ++sl) {
            const int cur = sl & 1;
            if (sl + 1 < nslab) {
                fetch(Q, qi, T, (sl + 1) * BK + lk, qv);
                fetch(X, ci, N, (sl + 1) * BK + lk, xv);
            }
            const int kleft = dim - sl * BK;                    // the zero-padded tail adds no fma steps
            const float* qrow = &Qs[cur][0][ty * 4];
            const float* xrow = &Xs[cur][0][tx * 4];
            auto kstep = [&](int kk) {
                const float4 a0 = *reinterpret_cast<const float4*>(qrow + kk * BM);
                const float4 a1 = *reinterpret_cast<const float4*>(qrow + kk * BM + 64);
                const float4 b0 = *reinterpret_cast<const float4*>(xrow + kk * BN);
                const float4 b1 = *reinterpret_cast<const float4*>(xrow + kk * BN + 64);
                const float a8[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float b8[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int a = 0; a < 8; ++a)
#pragma unroll
                    for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(a8[a], b8[b], acc[a][b]);
            };
            if (kleft >= BK) {                                  // whole slab: straight-line code, no per-step tests
#pragma unroll
                for (int kk = 0; kk < BK; ++kk) kstep(kk);
            } else {
                for (int kk = 0; kk < kleft; ++kk) kstep(kk);
            }
            if (sl + 1 < nslab) {
#pragma unroll
                for (int e = 0; e < 4; ++e) { Qs[cur ^ 1][lk + e][lr] = qv[e]; Xs[cur ^ 1][lk + e][lr] = xv[e]; }
            }
            __syncthreads();
        }
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const long long i = i0 + (a >> 2) * 64 + ty * 4 + (a & 3);
            if (i >= T) continue;
            long long qs = 0, qp = 0, qo = 0;
            if (sb) { qs = q[3 * i]; qp = q[3 * i + 1]; qo = q[3 * i + 2]; }
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const long long c = c0 + (b >> 2) * 64 + tx * 4 + (b & 3);
                if (c >= N) continue;
                const float sc = head ? with_bias(acc[a][b], sb, pb, ob, c, qp, qo) : with_bias(acc[a][b], sb, pb, ob, qs, qp, c);
                gt[a] += sc > tq[a];
                eq[a] += sc == tq[a];
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 8; ++a) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {                        // the 16 threads (half a warp) that share these query rows
            gt[a] += __shfl_xor_sync(0xffffffffu, gt[a], o);
            eq[a] += __shfl_xor_sync(0xffffffffu, eq[a], o);
        }
        const long long i = i0 + (a >> 2) * 64 + ty * 4 + (a & 3);
        if (tx == 0 && i < T) {
            if (gt[a]) atomicAdd(counts + 2 * i, gt[a]);
            if (eq[a]) atomicAdd(counts + 2 * i + 1, eq[a]);
        }
    }
}

// Known true completions of query i = entries [lo, hi) of the list sorted by (key, value) with key = p * N + other.
// Each distinct completion except the target is taken out of the counts (reference: its score is set to -inf).
__global__ void __launch_bounds__(256) k_rank_filter(const float* __restrict__ Q, const int64_t* __restrict__ q, long long T,
                                                     int head, const float* __restrict__ X, long long N, int dim,
                                                     const float* __restrict__ sb, const float* __restrict__ pb,
                                                     const float* __restrict__ ob, const float* __restrict__ tscore,
                                                     const uint64_t* __restrict__ keys, const int32_t* __restrict__ vals,
                                                     long long M, int32_t* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= T) return;
    const long long s = q[3 * i], p = q[3 * i + 1], o = q[3 * i + 2];
    if (s < 0 || s >= N || o < 0 || o >= N || p < 0) return;
    const long long target = head ? s : o;
    const uint64_t key = (uint64_t)p * (uint64_t)N + (uint64_t)(head ? o : s);
    long long lo = 0, hi = M;                                    // first entry with keys >= key
    while (lo < hi) { const long long mid = (lo + hi) >> 1; if (keys[mid] < key) lo = mid + 1; else hi = mid; }
    long long end = lo, top = M;                                 // first entry with keys > key
    while (end < top) { const long long mid = (end + top) >> 1; if (keys[mid] <= key) end = mid + 1; else top = mid; }
    const float t = __ldg(tscore + i);
    int gt = 0, eq = 0;
    for (long long e = lo + lane; e < end; e += 32) {
        const long long c = vals[e];
        if (c == target || (e > lo && vals[e - 1] == c)) continue;          // the target stays; repeated triples count once
        const float acc = dot_seq(Q + (size_t)i * dim, X + (size_t)c * dim, dim);
        const float sc = head ? with_bias(acc, sb, pb, ob, c, p, o) : with_bias(acc, sb, pb, ob, s, p, c);
        gt += sc > t;
        eq += sc == t;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        gt += __shfl_xor_sync(0xffffffffu, gt, d);
        eq += __shfl_xor_sync(0xffffffffu, eq, d);
    }
    if (lane == 0) {
        if (gt) atomicSub(counts + 2 * i, gt);
        if (eq) atomicSub(counts + 2 * i + 1, eq);
    }
}

// utils/misc.py:95-101: rank = raw + (ties - 1) // 2 + 1, ties counting the target itself
__global__ void k_rank_finish(const int32_t* __restrict__ counts, long long T, int64_t* __restrict__ ranks) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= T) return;
    const long long raw = counts[2 * i], ties = counts[2 * i + 1];
    ranks[i] = raw + (ties - 1) / 2 + 1;                           // ties >= 1 unless the target score is NaN
}

// composite sort key ((p * N + other) * N + value); other = o, value = s for the head lists and vice versa
__global__ void k_filter_keys(const int64_t* __restrict__ t, long long M, long long N, long long R, int head,
                              uint64_t* __restrict__ keys, int32_t* status) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= M) return;
    long long s = t[3 * e], p = t[3 * e + 1], o = t[3 * e + 2];
    if (s < 0 || s >= N || o < 0 || o >= N || p < 0 || p >= R) {
        if (status) atomicAdd(status, 1);
        s = p = o = 0;
    }
    const uint64_t other = head ? o : s, value = head ? s : o;
    keys[e] = ((uint64_t)p * N + other) * N + value;
}

__global__ void k_filter_split(const uint64_t* __restrict__ sorted, long long M, long long N, uint64_t* __restrict__ keys,
                               int32_t* __restrict__ vals) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= M) return;
    keys[e] = sorted[e] / (uint64_t)N;
    vals[e] = (int32_t)(sorted[e] % (uint64_t)N);
}

size_t sort_temp_bytes(int64_t M) {
    size_t b = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, b, (uint64_t*)nullptr, (uint64_t*)nullptr, (int)(M > 0 ? M : 1));
    return align_up(b);
}

}  // namespace

extern "C" size_t rgcn_rank_filter_workspace_bytes(int64_t num_true) {
    const size_t n = (size_t)(num_true > 0 ? num_true : 1);
    return 2 * align_up(n * sizeof(uint64_t)) + sort_temp_bytes(num_true);
}

extern "C" int rgcn_rank_build_filter(const int64_t* all_triples, int64_t M, int64_t N, int64_t R, int head,
                                      uint64_t* keys, int32_t* vals, int32_t* status, void* ws, size_t ws_bytes,
                                      rgcn_stream_t stream) {
    RGCN_REQUIRE(M >= 0 && N > 0 && R > 0, RGCN_ERR_ARG, "rgcn_rank_build_filter: bad sizes");
    if (M == 0) return RGCN_OK;
    RGCN_REQUIRE(all_triples && keys && vals, RGCN_ERR_ARG, "rgcn_rank_build_filter: NULL pointer");
    RGCN_REQUIRE(M < (int64_t)INT32_MAX && N < (int64_t)INT32_MAX, RGCN_ERR_UNSUPPORTED, "rgcn_rank_build_filter: sizes must fit int32");
    const unsigned __int128 maxkey = (unsigned __int128)R * (unsigned __int128)N * (unsigned __int128)N;
    RGCN_REQUIRE((maxkey >> 63) == 0, RGCN_ERR_UNSUPPORTED, "rgcn_rank_build_filter: R*N*N does not fit a 64-bit sort key");
    RGCN_REQUIRE(ws && ws_bytes >= rgcn_rank_filter_workspace_bytes(M), RGCN_ERR_WORKSPACE, "rgcn_rank_build_filter: workspace too small");
    const cudaStream_t st = (cudaStream_t)stream;
    Carver c(ws);
    uint64_t* k0 = c.take<uint64_t>((size_t)M);
    uint64_t* k1 = c.take<uint64_t>((size_t)M);
    size_t temp = sort_temp_bytes(M);
    void* cubws = c.take<char>(temp);
    int bits = 1;
    while (bits < 64 && (maxkey >> bits) != 0) ++bits;
    RGCN_LAUNCH(k_filter_keys, grid_for(M, 256), 256, 0, st, all_triples, (long long)M, (long long)N, (long long)R, head, k0, status);
    RGCN_CHECK_CUDA(cub::DeviceRadixSort::SortKeys(cubws, temp, k0, k1, (int)M, 0, bits, st));
    rgcn::g_launches.fetch_add((bits + 7) / 8 + 1, std::memory_order_relaxed);
    RGCN_LAUNCH(k_filter_split, grid_for(M, 256), 256, 0, st, k1, (long long)M, (long long)N, keys, vals);
    return RGCN_OK;
}

extern "C" size_t rgcn_rank_workspace_bytes(int64_t num_queries, int64_t dim) {
    const size_t t = (size_t)(num_queries > 0 ? num_queries : 1);
    return align_up(t * (size_t)dim * sizeof(float)) + align_up(t * sizeof(float)) + align_up(2 * t * sizeof(int32_t)) +
           align_up(3 * t * sizeof(int64_t));
}

extern "C" int rgcn_rank_triples(const int64_t* queries, int64_t T, int head, const float* nodes, int64_t N,
                                 const float* relations, int64_t R, int64_t dim, const float* sbias, const float* pbias,
                                 const float* obias, const uint64_t* filter_keys, const int32_t* filter_vals,
                                 int64_t num_true, int64_t* ranks, int32_t* status, void* ws, size_t ws_bytes,
                                 rgcn_stream_t stream) {
    RGCN_REQUIRE(T >= 0 && N > 0 && R > 0 && dim > 0 && dim < (1 << 20), RGCN_ERR_ARG, "rgcn_rank_triples: bad sizes");
    RGCN_REQUIRE((sbias && pbias && obias) || (!sbias && !pbias && !obias), RGCN_ERR_ARG,
                 "rgcn_rank_triples: the three biases come together");
    if (T == 0) return RGCN_OK;
    RGCN_REQUIRE(queries && nodes && relations && ranks, RGCN_ERR_ARG, "rgcn_rank_triples: NULL pointer");
    RGCN_REQUIRE(num_true == 0 || (filter_keys && filter_vals), RGCN_ERR_ARG, "rgcn_rank_triples: filter lists missing");
    RGCN_REQUIRE(ws && ws_bytes >= rgcn_rank_workspace_bytes(T, dim), RGCN_ERR_WORKSPACE, "rgcn_rank_triples: workspace too small");
    RGCN_REQUIRE((T + BM - 1) / BM <= 65535, RGCN_ERR_UNSUPPORTED, "rgcn_rank_triples: at most %d queries per call", 65535 * BM);
    const cudaStream_t st = (cudaStream_t)stream;
    Carver c(ws);
    float* Q = c.take<float>((size_t)T * dim);
    float* tscore = c.take<float>((size_t)T);
    int32_t* counts = c.take<int32_t>(2 * (size_t)T);
    int64_t* qclean = c.take<int64_t>(3 * (size_t)T);
    RGCN_CHECK_CUDA(cudaMemsetAsync(counts, 0, 2 * (size_t)T * sizeof(int32_t), st));
    RGCN_LAUNCH(k_rank_queries, grid_for(T, 128), 128, 0, st, queries, (long long)T, head, nodes, (long long)N, relations,
                (long long)R, (int)dim, sbias, pbias, obias, Q, tscore, qclean, status);
    const int ytiles = (int)((T + BM - 1) / BM);
    int64_t xtiles = (N + BN - 1) / BN;
    // One balanced wave: at most 2 CTAs per SM in total (a 1.08-wave grid costs two waves), candidate tiles split
    // evenly over the grid's x dimension.  More query tiles than CTA slots: one column of CTAs, several waves.
    int64_t want = ((int64_t)kNumSMs * 2) / ytiles;
    if (want < 1) want = 1;
    if (xtiles > want) xtiles = want;
    const bool vec = dim % 4 == 0 && (reinterpret_cast<uintptr_t>(nodes) & 15) == 0;     // Q comes from the aligned workspace
    if (vec)
        RGCN_LAUNCH(k_rank_count<true>, dim3((unsigned)xtiles, (unsigned)ytiles), 256, 0, st, Q, qclean, (long long)T, head,
                    nodes, (long long)N, (int)dim, sbias, pbias, obias, tscore, counts);
    else
        RGCN_LAUNCH(k_rank_count<false>, dim3((unsigned)xtiles, (unsigned)ytiles), 256, 0, st, Q, qclean, (long long)T, head,
                    nodes, (long long)N, (int)dim, sbias, pbias, obias, tscore, counts);
    if (num_true > 0)
        RGCN_LAUNCH(k_rank_filter, grid_for(T, 8), 256, 0, st, Q, qclean, (long long)T, head, nodes, (long long)N, (int)dim,
                    sbias, pbias, obias, tscore, filter_keys, filter_vals, (long long)num_true, counts);
    RGCN_LAUNCH(k_rank_finish, grid_for(T, 256), 256, 0, st, counts, (long long)T, ranks);
    return RGCN_OK;
}
