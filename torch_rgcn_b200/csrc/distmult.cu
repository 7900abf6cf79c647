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

template <bool VEC>
__global__ void __launch_bounds__(kWarps * 32) k_distmult_fwd(const int64_t* __restrict__ t, long long B,
                                                              const float* __restrict__ nodes, long long N,
                                                              const float* __restrict__ rel, long long R, int dim,
                                                              const float* __restrict__ sbias,
                                                              const float* __restrict__ pbias,
                                                              const float* __restrict__ obias,
                                                              float* __restrict__ scores, int32_t* status) {
    const int lane = threadIdx.x & 31;
    const long long nwarp = (long long)gridDim.x * kWarps;
    for (long long b = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5); b < B; b += nwarp) {
        long long s, p, o;
        if (!load_triple(t, b, N, R, lane, status, s, p, o)) {
            if (lane == 0) scores[b] = 0.f;
            continue;
        }
        const float* xs = nodes + (size_t)s * dim;
        const float* xp = rel + (size_t)p * dim;
        const float* xo = nodes + (size_t)o * dim;
        float acc = 0.f;
        if constexpr (VEC) {
            for (int j = lane; j < dim / 4; j += 32) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(xs) + j);
                const float4 r = __ldg(reinterpret_cast<const float4*>(xp) + j);
                const float4 c = __ldg(reinterpret_cast<const float4*>(xo) + j);
                acc += a.x * r.x * c.x + a.y * r.y * c.y + a.z * r.z * c.z + a.w * r.w * c.w;
            }
        } else {
            for (int j = lane; j < dim; j += 32) acc += __ldg(xs + j) * __ldg(xp + j) * __ldg(xo + j);
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            if (sbias) acc += __ldg(sbias + s) + __ldg(pbias + p) + __ldg(obias + o);   // layers.py:95-96
            scores[b] = acc;
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

// sums[0..2] += sum over this CTA's triples of |nodes[s]|^2, |relations[p]|^2, |nodes[o]|^2 (double accumulators)
__global__ void __launch_bounds__(kWarps * 32) k_distmult_penalty(const int64_t* __restrict__ t, long long B,
                                                                  const float* __restrict__ nodes, long long N,
                                                                  const float* __restrict__ rel, long long R, int dim,
                                                                  double* __restrict__ sums, int32_t* status) {
    __shared__ double part[kWarps][3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long nwarp = (long long)gridDim.x * kWarps;
    float as = 0.f, ap = 0.f, ao = 0.f;
    double ds = 0.0, dp = 0.0, dob = 0.0;
    int since = 0;
    for (long long b = (long long)blockIdx.x * kWarps + warp; b < B; b += nwarp) {
        long long s, p, o;
        if (!load_triple(t, b, N, R, lane, status, s, p, o)) continue;
        for (int j = lane; j < dim; j += 32) {
            const float a = __ldg(nodes + (size_t)s * dim + j), r = __ldg(rel + (size_t)p * dim + j),
                        c = __ldg(nodes + (size_t)o * dim + j);
            as += a * a; ap += r * r; ao += c * c;
        }
        if (++since == 64) { ds += as; dp += ap; dob += ao; as = ap = ao = 0.f; since = 0; }   // bound the fp32 run length
    }
    ds += as; dp += ap; dob += ao;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ds += __shfl_xor_sync(0xffffffffu, ds, o);
        dp += __shfl_xor_sync(0xffffffffu, dp, o);
        dob += __shfl_xor_sync(0xffffffffu, dob, o);
    }
    if (lane == 0) { part[warp][0] = ds; part[warp][1] = dp; part[warp][2] = dob; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double v = 0.0;
        for (int k = 0; k < kWarps; ++k) v += part[k][threadIdx.x];
        atomicAdd(sums + threadIdx.x, v);
    }
}

__global__ void k_distmult_penalty_finish(const double* __restrict__ sums, long long B, int dim, float* __restrict__ out) {
    if (blockIdx.x || threadIdx.x) return;
    const double n = (double)B * (double)dim;
    out[0] = (float)(sums[0] / n) + (float)(sums[1] / n) + (float)(sums[2] / n);
}

// d penalty / d nodes[s_b] = 2 nodes[s_b] / (B dim) per occurrence, likewise for o and for relations[p]
__global__ void __launch_bounds__(kWarps * 32) k_distmult_penalty_bwd(const int64_t* __restrict__ t, long long B,
                                                                      const float* __restrict__ nodes, long long N,
                                                                      const float* __restrict__ rel, long long R, int dim,
                                                                      const float* __restrict__ grad,
                                                                      float* __restrict__ g_nodes, float* __restrict__ g_rel) {
    const int lane = threadIdx.x & 31;
    const long long nwarp = (long long)gridDim.x * kWarps;
    const float scale = 2.f * __ldg(grad) / ((float)B * (float)dim);
    for (long long b = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5); b < B; b += nwarp) {
        long long s, p, o;
        if (!load_triple(t, b, N, R, lane, nullptr, s, p, o)) continue;
        for (int j = lane; j < dim; j += 32) {
            if (g_nodes) {
                atomicAdd(g_nodes + (size_t)s * dim + j, scale * __ldg(nodes + (size_t)s * dim + j));
                atomicAdd(g_nodes + (size_t)o * dim + j, scale * __ldg(nodes + (size_t)o * dim + j));
            }
            if (g_rel) atomicAdd(g_rel + (size_t)p * dim + j, scale * __ldg(rel + (size_t)p * dim + j));
        }
    }
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

extern "C" size_t rgcn_distmult_penalty_workspace_bytes(void) { return align_up(3 * sizeof(double)); }

extern "C" int rgcn_distmult_penalty(const int64_t* triples, int64_t B, const float* nodes, int64_t N,
                                     const float* relations, int64_t R, int64_t dim, float* out, int32_t* status,
                                     void* workspace, size_t workspace_bytes, rgcn_stream_t stream) {
    int rc = check_args("rgcn_distmult_penalty", triples, B, nodes, N, relations, R, dim);
    if (rc) return rc;
    RGCN_REQUIRE(out && workspace && workspace_bytes >= rgcn_distmult_penalty_workspace_bytes(), RGCN_ERR_WORKSPACE,
                 "rgcn_distmult_penalty: out / workspace missing");
    RGCN_REQUIRE(B > 0, RGCN_ERR_ARG, "rgcn_distmult_penalty: the mean over an empty batch is undefined");
    const cudaStream_t st = (cudaStream_t)stream;
    double* sums = static_cast<double*>(workspace);
    RGCN_CHECK_CUDA(cudaMemsetAsync(sums, 0, 3 * sizeof(double), st));
    RGCN_LAUNCH(k_distmult_penalty, grid_warps(B, 8), kWarps * 32, 0, st, triples, (long long)B, nodes, (long long)N,
                relations, (long long)R, (int)dim, sums, status);
    RGCN_LAUNCH(k_distmult_penalty_finish, 1, 32, 0, st, sums, (long long)B, (int)dim, out);
    return RGCN_OK;
}

extern "C" int rgcn_distmult_penalty_backward(const int64_t* triples, int64_t B, const float* nodes, int64_t N,
                                              const float* relations, int64_t R, int64_t dim, const float* grad,
                                              float* g_nodes, float* g_relations, rgcn_stream_t stream) {
    int rc = check_args("rgcn_distmult_penalty_backward", triples, B, nodes, N, relations, R, dim);
    if (rc) return rc;
    RGCN_REQUIRE(grad && B > 0, RGCN_ERR_ARG, "rgcn_distmult_penalty_backward: grad is NULL or the batch is empty");
    const cudaStream_t st = (cudaStream_t)stream;
    if (g_nodes) RGCN_CHECK_CUDA(cudaMemsetAsync(g_nodes, 0, (size_t)N * dim * sizeof(float), st));
    if (g_relations) RGCN_CHECK_CUDA(cudaMemsetAsync(g_relations, 0, (size_t)R * dim * sizeof(float), st));
    RGCN_LAUNCH(k_distmult_penalty_bwd, grid_warps(B, 8), kWarps * 32, 0, st, triples, (long long)B, nodes, (long long)N,
                relations, (long long)R, (int)dim, grad, g_nodes, g_relations);
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
