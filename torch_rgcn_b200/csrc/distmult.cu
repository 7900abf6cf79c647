// DistMult decoder of the link-prediction models: scores, their gradients, the decoder's
// L2 penalty and the negative-sampling corruption step.
//
// Reference behaviour restated here (never its code):
//   torch_rgcn/layers.py:86-98   DistMult.forward    score_b = sum_d nodes[s_b, d] * relations[p_b, d] * nodes[o_b, d]
//                                                              (+ sbias[s_b] + pbias[p_b] + obias[o_b])
//   torch_rgcn/layers.py:77-84   DistMult.s_penalty  mean(nodes[s]^2) + mean(relations[p]^2) + mean(nodes[o]^2)
//   utils/misc.py:174-189        negative_sampling   batch[mask] = corruptions, mask = [head?, 0, !head?] per triple
// Backward is autograd upstream; closed forms: g_nodes[s] += g p o, g_nodes[o] += g s p, g_relations[p] += g s o,
// bias gradients are scatter-adds of g.
//
// One warp per triple, lanes stride over the embedding in 16-byte pieces; the three gathered rows come from a
// node table that is L2-resident at the reference's sizes (WN18: 40,943 x 128 fp32 = 21 MB), so these kernels are
// bound by L2 bandwidth / atomic throughput, not HBM.
#include <algorithm>
#include "common.cuh"

using namespace rgcn;

namespace {

constexpr int kWarps = 8;          // warps per CTA

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void red_add4(float* p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// triple b -> (s, p, o); false (and one count in status) if any index is out of range
__device__ __forceinline__ bool load_triple(const int64_t* __restrict__ t, long long b, long long N, long long R,
                                            int lane, int32_t* status, long long& s, long long& p, long long& o) {
    s = t[3 * b]; p = t[3 * b + 1]; o = t[3 * b + 2];
    const bool ok = s >= 0 && s < N && o >= 0 && o < N && p >= 0 && p < R;
    if (!ok && lane == 0 && status) atomicAdd(status, 1);
    return ok;
}

// Warp per triple.  U triples per warp iteration were measured on the WN18-shaped batch: U = 1 0.353 ms, U = 4
// 0.397 ms (the extra registers cost more occupancy than the extra loads in flight buy), so U stays 1.
template <bool VEC>
__global__ void __launch_bounds__(kWarps * 32) k_distmult_fwd(const int64_t* __restrict__ t, long long B,
                                                              const float* __restrict__ nodes, long long N,
                                                              const float* __restrict__ rel, long long R, int dim,
                                                              const float* __restrict__ sbias,
                                                              const float* __restrict__ pbias,
                                                              const float* __restrict__ obias,
                                                              float* __restrict__ scores, int32_t* status) {
    constexpr int U = 1;
    const int lane = threadIdx.x & 31;
    const long long nwarp = (long long)gridDim.x * kWarps;
    for (long long b0 = ((long long)blockIdx.x * kWarps + (threadIdx.x >> 5)) * U; b0 < B; b0 += nwarp * U) {
        long long s[U], p[U], o[U];
        bool ok[U];
        float acc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            acc[u] = 0.f;
            ok[u] = b0 + u < B && load_triple(t, b0 + u, N, R, lane, status, s[u], p[u], o[u]);
        }
        if constexpr (VEC) {
            for (int j = lane; j < dim / 4; j += 32) {
                float4 a[U], r[U], c[U];
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (ok[u]) {
                        a[u] = __ldg(reinterpret_cast<const float4*>(nodes + (size_t)s[u] * dim) + j);
                        r[u] = __ldg(reinterpret_cast<const float4*>(rel + (size_t)p[u] * dim) + j);
                        c[u] = __ldg(reinterpret_cast<const float4*>(nodes + (size_t)o[u] * dim) + j);
                    }
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (ok[u])
                        acc[u] += a[u].x * r[u].x * c[u].x + a[u].y * r[u].y * c[u].y + a[u].z * r[u].z * c[u].z +
                                  a[u].w * r[u].w * c[u].w;
            }
        } else {
            for (int j = lane; j < dim; j += 32)
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (ok[u])
                        acc[u] += __ldg(nodes + (size_t)s[u] * dim + j) * __ldg(rel + (size_t)p[u] * dim + j) *
                                  __ldg(nodes + (size_t)o[u] * dim + j);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float v = warp_sum(acc[u]);
            if (lane == 0 && b0 + u < B) {
                float out = 0.f;
                if (ok[u]) {
                    out = v;
                    if (sbias) out += __ldg(sbias + s[u]) + __ldg(pbias + p[u]) + __ldg(obias + o[u]);   // layers.py:95-96
                }
                scores[b0 + u] = out;
            }
        }
    }
}

// Each warp walks a contiguous range of triples and keeps the relation gradient of the current relation in
// registers (batches of negatives repeat the relation of their positive, utils/misc.py:174-189 keeps p), flushing
// with vector reductions when the relation changes.  Node gradients go straight to vector reductions.
template <bool VEC, int MAXV>
__global__ void __launch_bounds__(kWarps * 32) k_distmult_bwd(const int64_t* __restrict__ t, long long B,
                                                              const float* __restrict__ nodes, long long N,
                                                              const float* __restrict__ rel, long long R, int dim,
                                                              const float* __restrict__ g, float* __restrict__ g_nodes,
                                                              float* __restrict__ g_rel, float* __restrict__ g_sbias,
                                                              float* __restrict__ g_pbias, float* __restrict__ g_obias,
                                                              long long per_warp) {
    const int lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
    const long long b0 = w * per_warp, b1 = min(B, b0 + per_warp);
    long long cur_p = -1;
    float4 racc[MAXV];
#pragma unroll
    for (int k = 0; k < MAXV; ++k) racc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    auto flush = [&]() {
        if (cur_p < 0 || !g_rel) return;
        float* dst = g_rel + (size_t)cur_p * dim;
        if constexpr (VEC) {
#pragma unroll
            for (int k = 0; k < MAXV; ++k) {
                const int j = lane + 32 * k;
                if (j < dim / 4) red_add4(dst + 4 * j, racc[k]);
                racc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {
#pragma unroll
            for (int k = 0; k < MAXV; ++k) {
                const int j = lane + 32 * k;
                if (j < dim) atomicAdd(dst + j, racc[k].x);
                racc[k].x = 0.f;
            }
        }
    };
    for (long long b = b0; b < b1; ++b) {
        long long s, p, o;
        if (!load_triple(t, b, N, R, lane, nullptr, s, p, o)) continue;
        const float gb = __ldg(g + b);
        if (p != cur_p) { flush(); cur_p = p; }
        const float* xs = nodes + (size_t)s * dim;
        const float* xp = rel + (size_t)p * dim;
        const float* xo = nodes + (size_t)o * dim;
        if constexpr (VEC) {
#pragma unroll
            for (int k = 0; k < MAXV; ++k) {
                const int j = lane + 32 * k;
                if (j >= dim / 4) break;
                const float4 a = __ldg(reinterpret_cast<const float4*>(xs) + j);
                const float4 r = __ldg(reinterpret_cast<const float4*>(xp) + j);
                const float4 c = __ldg(reinterpret_cast<const float4*>(xo) + j);
                racc[k].x += gb * a.x * c.x; racc[k].y += gb * a.y * c.y;
                racc[k].z += gb * a.z * c.z; racc[k].w += gb * a.w * c.w;
                if (g_nodes) {
                    red_add4(g_nodes + (size_t)s * dim + 4 * j, make_float4(gb * r.x * c.x, gb * r.y * c.y, gb * r.z * c.z, gb * r.w * c.w));
                    red_add4(g_nodes + (size_t)o * dim + 4 * j, make_float4(gb * a.x * r.x, gb * a.y * r.y, gb * a.z * r.z, gb * a.w * r.w));
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < MAXV; ++k) {
                const int j = lane + 32 * k;
                if (j >= dim) break;
                const float a = __ldg(xs + j), r = __ldg(xp + j), c = __ldg(xo + j);
                racc[k].x += gb * a * c;
                if (g_nodes) {
                    atomicAdd(g_nodes + (size_t)s * dim + j, gb * r * c);
                    atomicAdd(g_nodes + (size_t)o * dim + j, gb * a * r);
                }
            }
        }
        if (lane == 0 && g_sbias) {
            atomicAdd(g_sbias + s, gb); atomicAdd(g_pbias + p, gb); atomicAdd(g_obias + o, gb);
        }
    }
    flush();
}

// Penalty through occurrence counts: sum_b |nodes[s_b]|^2 + |nodes[o_b]|^2 = sum_v cnt_n[v] |nodes[v]|^2 (likewise
// for relations), so the 3 B gathered rows of the reference collapse into one histogram over the triples and one
// pass over the (small) embedding tables.  The counts stay in the workspace for the backward.
__global__ void __launch_bounds__(256) k_distmult_count(const int64_t* __restrict__ t, long long B, long long N,
                                                        long long R, int32_t* __restrict__ cnt_n,
                                                        int32_t* __restrict__ cnt_r, int32_t* status) {
    extern __shared__ int32_t s_rel[];                   // relation histogram of this CTA (few, hot addresses)
    const bool smem_hist = R <= 8192;
    if (smem_hist) {
        for (int i = threadIdx.x; i < (int)R; i += blockDim.x) s_rel[i] = 0;
        __syncthreads();
    }
    for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
        const long long s = t[3 * b], p = t[3 * b + 1], o = t[3 * b + 2];
        if (s < 0 || s >= N || o < 0 || o >= N || p < 0 || p >= R) {
            if (status) atomicAdd(status, 1);
            continue;
        }
        atomicAdd(cnt_n + s, 1);
        atomicAdd(cnt_n + o, 1);
        if (smem_hist) atomicAdd(s_rel + p, 1); else atomicAdd(cnt_r + p, 1);
    }
    if (smem_hist) {
        __syncthreads();
        for (int i = threadIdx.x; i < (int)R; i += blockDim.x)
            if (s_rel[i]) atomicAdd(cnt_r + i, s_rel[i]);
    }
}

// sums[which] += sum_v cnt[v] |table[v]|^2 : warp per row, double accumulation across rows
__global__ void __launch_bounds__(kWarps * 32) k_distmult_weighted_sq(const float* __restrict__ table, long long rows, int dim,
                                                                      const int32_t* __restrict__ cnt,
                                                                      double* __restrict__ sum) {
    __shared__ double part[kWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc = 0.0;
    for (long long v = (long long)blockIdx.x * kWarps + warp; v < rows; v += (long long)gridDim.x * kWarps) {
        const int32_t c = __ldg(cnt + v);
        if (!c) continue;
        float a = 0.f;
        for (int j = lane; j < dim; j += 32) {
            const float x = __ldg(table + (size_t)v * dim + j);
            a += x * x;
        }
        acc += (double)a * (double)c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) part[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0.0;
        for (int k = 0; k < kWarps; ++k) v += part[k];
        if (v != 0.0) atomicAdd(sum, v);
    }
}

__global__ void k_distmult_penalty_finish(const double* __restrict__ sums, long long B, int dim, float* __restrict__ out) {
    if (blockIdx.x || threadIdx.x) return;
    const double n = (double)B * (double)dim;
    out[0] = (float)(sums[0] / n) + (float)(sums[1] / n);     // nodes (subjects + objects), relations
}

// g[v, :] = 2 grad cnt[v] table[v, :] / (B dim)
__global__ void k_distmult_penalty_bwd(const float* __restrict__ table, long long rows, int dim, const int32_t* __restrict__ cnt,
                                       const float* __restrict__ grad, long long B, float* __restrict__ g) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= rows * dim) return;
    const float scale = 2.f * __ldg(grad) / ((float)B * (float)dim);
    g[i] = scale * (float)__ldg(cnt + i / dim) * __ldg(table + i);
}

// batch[i, head[i] ? 0 : 2] = corruptions[i]  (the masked assignment of utils/misc.py:181-187, row-major order)
__global__ void k_corrupt_triples(int64_t* __restrict__ batch, const uint8_t* __restrict__ head,
                                  const int64_t* __restrict__ corruptions, long long count) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= count) return;
    batch[3 * i + (head[i] ? 0 : 2)] = corruptions[i];
}

int check_args(const char* who, const int64_t* t, int64_t B, const float* nodes, int64_t N, const float* rel, int64_t R,
               int64_t dim) {
    RGCN_REQUIRE(B >= 0 && N > 0 && R > 0 && dim > 0 && dim < (1 << 20), RGCN_ERR_ARG, "%s: bad sizes B=%lld N=%lld R=%lld dim=%lld",
                 who, (long long)B, (long long)N, (long long)R, (long long)dim);
    RGCN_REQUIRE(nodes && rel && (t || B == 0), RGCN_ERR_ARG, "%s: NULL pointer", who);
    return RGCN_OK;
}

int grid_warps(int64_t B, int per_sm) {
    int64_t want = (B + kWarps - 1) / kWarps;
    const int64_t cap = (int64_t)kNumSMs * per_sm;
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace

extern "C" int rgcn_distmult_forward(const int64_t* triples, int64_t B, const float* nodes, int64_t N,
                                     const float* relations, int64_t R, int64_t dim, const float* sbias,
                                     const float* pbias, const float* obias, float* scores, int32_t* status,
                                     rgcn_stream_t stream) {
    int rc = check_args("rgcn_distmult_forward", triples, B, nodes, N, relations, R, dim);
    if (rc) return rc;
    RGCN_REQUIRE((sbias && pbias && obias) || (!sbias && !pbias && !obias), RGCN_ERR_ARG,
                 "rgcn_distmult_forward: the three biases come together (layers.py:27-34)");
    if (B == 0) return RGCN_OK;
    RGCN_REQUIRE(scores, RGCN_ERR_ARG, "rgcn_distmult_forward: scores is NULL");
    const cudaStream_t st = (cudaStream_t)stream;
    const int grid = grid_warps(B, 16);
    if (dim % 4 == 0)
        RGCN_LAUNCH(k_distmult_fwd<true>, grid, kWarps * 32, 0, st, triples, (long long)B, nodes, (long long)N, relations,
                    (long long)R, (int)dim, sbias, pbias, obias, scores, status);
    else
        RGCN_LAUNCH(k_distmult_fwd<false>, grid, kWarps * 32, 0, st, triples, (long long)B, nodes, (long long)N, relations,
                    (long long)R, (int)dim, sbias, pbias, obias, scores, status);
    return RGCN_OK;
}

extern "C" int rgcn_distmult_backward(const int64_t* triples, int64_t B, const float* nodes, int64_t N,
                                      const float* relations, int64_t R, int64_t dim, const float* grad_scores,
                                      float* g_nodes, float* g_relations, float* g_sbias, float* g_pbias, float* g_obias,
                                      rgcn_stream_t stream) {
    int rc = check_args("rgcn_distmult_backward", triples, B, nodes, N, relations, R, dim);
    if (rc) return rc;
    RGCN_REQUIRE((g_sbias && g_pbias && g_obias) || (!g_sbias && !g_pbias && !g_obias), RGCN_ERR_ARG,
                 "rgcn_distmult_backward: the three bias gradients come together");
    RGCN_REQUIRE(dim <= 4096, RGCN_ERR_UNSUPPORTED, "rgcn_distmult_backward: dim %lld > 4096 not supported", (long long)dim);
    const cudaStream_t st = (cudaStream_t)stream;
    if (g_nodes) RGCN_CHECK_CUDA(cudaMemsetAsync(g_nodes, 0, (size_t)N * dim * sizeof(float), st));
    if (g_relations) RGCN_CHECK_CUDA(cudaMemsetAsync(g_relations, 0, (size_t)R * dim * sizeof(float), st));
    if (g_sbias) {
        RGCN_CHECK_CUDA(cudaMemsetAsync(g_sbias, 0, (size_t)N * sizeof(float), st));
        RGCN_CHECK_CUDA(cudaMemsetAsync(g_obias, 0, (size_t)N * sizeof(float), st));
        RGCN_CHECK_CUDA(cudaMemsetAsync(g_pbias, 0, (size_t)R * sizeof(float), st));
    }
    if (B == 0) return RGCN_OK;
    RGCN_REQUIRE(grad_scores, RGCN_ERR_ARG, "rgcn_distmult_backward: grad_scores is NULL");
    // contiguous triple ranges per warp: short enough to fill the GPU, long enough to amortise relation flushes
    int64_t per_warp = (B + (int64_t)kNumSMs * 16 * kWarps - 1) / ((int64_t)kNumSMs * 16 * kWarps);
    if (per_warp < 16) per_warp = 16;
    const int grid = (int)((B + per_warp * kWarps - 1) / (per_warp * kWarps));
#define RGCN_DM_BWD(VEC, MAXV)                                                                                          \
    RGCN_LAUNCH((k_distmult_bwd<VEC, MAXV>), grid, kWarps * 32, 0, st, triples, (long long)B, nodes, (long long)N,      \
                relations, (long long)R, (int)dim, grad_scores, g_nodes, g_relations, g_sbias, g_pbias, g_obias,        \
                (long long)per_warp)
    if (dim % 4 == 0) {
        const int64_t v = (dim / 4 + 31) / 32;
        if (v <= 1) RGCN_DM_BWD(true, 1); else if (v <= 2) RGCN_DM_BWD(true, 2); else if (v <= 4) RGCN_DM_BWD(true, 4);
        else if (v <= 8) RGCN_DM_BWD(true, 8); else if (v <= 16) RGCN_DM_BWD(true, 16); else RGCN_DM_BWD(true, 32);
    } else {
        const int64_t v = (dim + 31) / 32;
        RGCN_REQUIRE(v <= 32, RGCN_ERR_UNSUPPORTED, "rgcn_distmult_backward: dim %lld not a multiple of 4 and > 1024", (long long)dim);
        if (v <= 2) RGCN_DM_BWD(false, 2); else if (v <= 8) RGCN_DM_BWD(false, 8); else RGCN_DM_BWD(false, 32);
    }
#undef RGCN_DM_BWD
    return RGCN_OK;
}

extern "C" size_t rgcn_distmult_penalty_workspace_bytes(int64_t N, int64_t R) {
    return align_up(2 * sizeof(double)) + align_up((size_t)(N > 0 ? N : 0) * sizeof(int32_t)) +
           align_up((size_t)(R > 0 ? R : 0) * sizeof(int32_t));
}

namespace {
struct PenaltyWs { double* sums; int32_t* cnt_n; int32_t* cnt_r; };
PenaltyWs carve_penalty(void* ws, int64_t N, int64_t R) {
    Carver c(ws);
    PenaltyWs w;
    w.sums = c.take<double>(2); w.cnt_n = c.take<int32_t>((size_t)N); w.cnt_r = c.take<int32_t>((size_t)R);
    return w;
}
}  // namespace

extern "C" int rgcn_distmult_penalty(const int64_t* triples, int64_t B, const float* nodes, int64_t N,
                                     const float* relations, int64_t R, int64_t dim, float* out, int32_t* status,
                                     void* workspace, size_t workspace_bytes, rgcn_stream_t stream) {
    int rc = check_args("rgcn_distmult_penalty", triples, B, nodes, N, relations, R, dim);
    if (rc) return rc;
    const size_t need = rgcn_distmult_penalty_workspace_bytes(N, R);
    RGCN_REQUIRE(out && workspace && workspace_bytes >= need, RGCN_ERR_WORKSPACE,
                 "rgcn_distmult_penalty: out / workspace missing (%zu < %zu bytes)", workspace_bytes, need);
    RGCN_REQUIRE(B > 0, RGCN_ERR_ARG, "rgcn_distmult_penalty: the mean over an empty batch is undefined");
    const cudaStream_t st = (cudaStream_t)stream;
    PenaltyWs w = carve_penalty(workspace, N, R);
    RGCN_CHECK_CUDA(cudaMemsetAsync(workspace, 0, need, st));
    const int cgrid = (int)std::min<int64_t>((B + 255) / 256, (int64_t)kNumSMs * 8);
    RGCN_LAUNCH(k_distmult_count, cgrid, 256, R <= 8192 ? (size_t)R * sizeof(int32_t) : 0, st, triples, (long long)B,
                (long long)N, (long long)R, w.cnt_n, w.cnt_r, status);
    RGCN_LAUNCH(k_distmult_weighted_sq, grid_warps(N, 8), kWarps * 32, 0, st, nodes, (long long)N, (int)dim, w.cnt_n, w.sums);
    RGCN_LAUNCH(k_distmult_weighted_sq, grid_warps(R, 8), kWarps * 32, 0, st, relations, (long long)R, (int)dim, w.cnt_r,
                w.sums + 1);
    RGCN_LAUNCH(k_distmult_penalty_finish, 1, 32, 0, st, w.sums, (long long)B, (int)dim, out);
    return RGCN_OK;
}

extern "C" int rgcn_distmult_penalty_backward(const void* workspace, int64_t B, const float* nodes, int64_t N,
                                              const float* relations, int64_t R, int64_t dim, const float* grad,
                                              float* g_nodes, float* g_relations, rgcn_stream_t stream) {
    RGCN_REQUIRE(workspace && nodes && relations && grad && B > 0 && N > 0 && R > 0 && dim > 0, RGCN_ERR_ARG,
                 "rgcn_distmult_penalty_backward: bad arguments");
    const cudaStream_t st = (cudaStream_t)stream;
    PenaltyWs w = carve_penalty(const_cast<void*>(workspace), N, R);
    if (g_nodes)
        RGCN_LAUNCH(k_distmult_penalty_bwd, grid_for(N * dim, 256), 256, 0, st, nodes, (long long)N, (int)dim, w.cnt_n, grad,
                    (long long)B, g_nodes);
    if (g_relations)
        RGCN_LAUNCH(k_distmult_penalty_bwd, grid_for(R * dim, 256), 256, 0, st, relations, (long long)R, (int)dim, w.cnt_r,
                    grad, (long long)B, g_relations);
    return RGCN_OK;
}

extern "C" int rgcn_corrupt_triples(int64_t* batch, const uint8_t* head_mask, const int64_t* corruptions, int64_t count,
                                    rgcn_stream_t stream) {
    RGCN_REQUIRE(count >= 0 && (count == 0 || (batch && head_mask && corruptions)), RGCN_ERR_ARG,
                 "rgcn_corrupt_triples: bad arguments");
    if (count == 0) return RGCN_OK;
    RGCN_LAUNCH(k_corrupt_triples, grid_for(count, 256), 256, 0, (cudaStream_t)stream, batch, head_mask, corruptions,
                (long long)count);
    return RGCN_OK;
}
