// Per-step graph construction for link prediction on the device.
//
// Reference behaviour restated here (never its code): utils/misc.py:125-172 `edge_neighborhood` grows a sample of
// training edges from a frontier: every pick chooses a vertex with probability proportional to
// (remaining degree) x (seen before), or uniformly among vertices with remaining edges while nothing seen has edges
// left, then one of that vertex's not-yet-picked incident edges uniformly, and marks both ends seen.  Upstream this is
// a Python loop that rebuilds an O(N) probability vector with numpy for each of the 30,000 picks of a step.
//
// The process is sequential by definition (the distribution of pick i depends on all earlier picks), so the device
// version is ONE warp that keeps the whole sampler state on chip and makes each pick in O(log32 N):
//   * the vertex weights are the leaves of two implicit 32-ary sum trees (A: count x seen, B: count > 0) whose upper
//     levels live in shared memory; a descent is one shared-memory read + one warp scan per level;
//   * counts, seen / picked bitmaps sit in shared memory when they fit (WN18: 164 KB + 5 KB + 18 KB), else in the
//     caller's workspace;
//   * the adjacency (edge order of the reference: per vertex by edge index, subject entry first) is built once per
//     training set by rgcn_sampler_build (radix sort).
// Randomness arrives as two uniforms per pick, so the kernel is a pure function of its inputs and is checked pick for
// pick against oracle/sampling_oracle.py.
#include <cub/cub.cuh>
#include "common.cuh"

using namespace rgcn;

namespace {

constexpr int kMaxLevels = 6;                      // upper levels: 32^7 > 2^31 leaves
constexpr size_t kSmemBudget = 200 * 1024;         // dynamic shared memory the sampler may use

struct TreeShape {
    int levels;                                    // upper levels (level 1 = groups of 32 leaves ...); 0 when N <= 32
    int size[kMaxLevels];                          // entries of upper level l, padded to a multiple of 32
    int off[kMaxLevels];                           // offset of upper level l inside one tree's array
    int total;                                     // entries of one tree's upper levels
};

TreeShape tree_shape(int64_t N) {
    TreeShape t{};
    int64_t n = N;
    int off = 0;
    while (n > 32) {
        n = (n + 31) / 32;
        const int pad = (int)((n + 31) / 32 * 32);
        t.size[t.levels] = pad;
        t.off[t.levels] = off;
        off += pad;
        ++t.levels;
    }
    t.total = off;
    return t;
}

struct SamplerArgs {
    const int32_t* adj_ptr;                        // (N + 1)
    const int2* adj;                               // (2 E) {edge index, other end} of each adjacency entry
    const float* uniforms;                         // (S, 2)
    int32_t* out;                                  // (S) picked edge indices
    int32_t* g_counts;                             // global fallbacks (NULL when the array is in shared memory)
    uint32_t* g_seen;
    uint32_t* g_picked;
    int32_t* g_tree;                               // 2 * shape.total
    int32_t* status;                               // != 0 afterwards: ran out of edges (never with S <= E)
    long long N, E, S;
    TreeShape shape;
};

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

__device__ __forceinline__ bool get_bit(const uint32_t* b, int i) { return (b[i >> 5] >> (i & 31)) & 1u; }

// The S picks, by one warp.  Upper tree levels hold INCLUSIVE PREFIX SUMS inside each group of 32 siblings, so a
// descent step is one read + ballot (no warp scan) and an update adds the weight change to the siblings at and after
// the changed child.  LEVELS is a template parameter so every per-level loop unrolls and the level offsets stay in
// registers (a single warp is instruction-latency bound: ~250 instructions per pick).
template <int LEVELS>
__device__ void sampler_picks(const SamplerArgs& a, int32_t* counts, uint32_t* seen, uint32_t* picked, int32_t* treeA,
                              int32_t* treeB, const int lane) {       // no __restrict__: lanes exchange data through these
    const int N = (int)a.N;
    int off[LEVELS > 0 ? LEVELS : 1];
#pragma unroll
    for (int l = 0; l < LEVELS; ++l) off[l] = a.shape.off[l];
    unsigned WA = 0, WB;
    if (LEVELS == 0) WB = __reduce_add_sync(0xffffffffu, (lane < N && counts[lane] > 0) ? 1 : 0);
    else WB = (unsigned)treeB[off[LEVELS - 1] + 31];               // the top level is one padded group

    const float2* __restrict__ uni = reinterpret_cast<const float2*>(a.uniforms);
    float2 unext = __ldg(uni);
    const int S = (int)a.S;
    for (int it = 0; it < S; ++it) {
        const float u1 = unext.x, u2 = unext.y;
        if (it + 1 < S) unext = __ldg(uni + it + 1);                // off the critical path
        const bool useB = WA == 0;
        const unsigned W = useB ? WB : WA;
        if (W == 0) {                                               // out of edges: the reference divides by zero here
            if (lane == 0) atomicAdd(a.status, 1);
            return;
        }
        unsigned target = (unsigned)((double)u1 * (double)W);
        if (target > W - 1) target = W - 1;
        // ---- descent: the first index whose inclusive prefix sum exceeds target
        const int32_t* tr = useB ? treeB : treeA;
        int node = 0;
#pragma unroll
        for (int l = LEVELS - 1; l >= 0; --l) {
            const unsigned p = (unsigned)tr[off[l] + node * 32 + lane];
            const unsigned hit = __ballot_sync(0xffffffffu, p > target);
            const int child = hit ? __ffs(hit) - 1 : 31;            // hit != 0 by construction (target < group sum)
            const unsigned prev = __shfl_sync(0xffffffffu, p, child > 0 ? child - 1 : 0);
            target -= child > 0 ? prev : 0u;
            node = node * 32 + child;
        }
        int v, c, lo, hi;
        {
            const int idx = node * 32 + lane;
            int w = 0, cl = 0, pl = 0, ph = 0;
            if (idx < N) {
                pl = __ldg(a.adj_ptr + idx);                        // every candidate's list bounds: the L2 latency of the
                ph = __ldg(a.adj_ptr + idx + 1);                    // chosen one hides behind the scan below
                cl = counts[idx];
                w = useB ? (cl > 0) : (get_bit(seen, idx) ? cl : 0);
            }
            const int incl = warp_incl_scan(w, lane);
            const unsigned hit = __ballot_sync(0xffffffffu, (unsigned)incl > target);
            const int child = hit ? __ffs(hit) - 1 : 31;
            v = node * 32 + child;
            c = __shfl_sync(0xffffffffu, cl, child);
            lo = __shfl_sync(0xffffffffu, pl, child);
            hi = __shfl_sync(0xffffffffu, ph, child);
        }
        if (v >= N || c <= 0) {                                     // cannot happen (tree sums match the leaves)
            if (lane == 0) atomicAdd(a.status, 1 << 20);
            return;
        }
        // ---- the j-th not-yet-picked entry of adj[v]
        int j = (int)((double)u2 * (double)c);
        if (j > c - 1) j = c - 1;
        int e = -1, other = -1;
        for (int base = lo; base < hi; base += 32) {
            const int k = base + lane;
            int2 ent = make_int2(-1, -1);
            bool free_entry = false;
            if (k < hi) { ent = a.adj[k]; free_entry = !get_bit(picked, ent.x); }
            const unsigned m = __ballot_sync(0xffffffffu, free_entry);
            const int n = __popc(m);
            if (j < n) {
                const int rank = __popc(m & ((1u << lane) - 1u));   // this lane's position among the free entries
                const unsigned sel = __ballot_sync(0xffffffffu, free_entry && rank == j);
                const int src = __ffs(sel) - 1;
                e = __shfl_sync(0xffffffffu, ent.x, src);
                other = __shfl_sync(0xffffffffu, ent.y, src);
                break;
            }
            j -= n;
        }
        if (e < 0) {                                                // cannot happen (counts[v] == free entries)
            if (lane == 0) atomicAdd(a.status, 1 << 16);
            return;
        }
        // ---- state update
        const int cv = c, co = counts[other];
        const bool sv = get_bit(seen, v), so = get_bit(seen, other);
        const bool same = other == v;
        const int cv2 = same ? cv - 2 : cv - 1, co2 = same ? cv2 : co - 1;
        const int dAv = cv2 - (sv ? cv : 0);                        // v is seen afterwards
        const int dAo = same ? 0 : co2 - (so ? co : 0);
        const int dBv = (cv2 > 0) - (cv > 0);
        const int dBo = same ? 0 : (co2 > 0) - (co > 0);
        __syncwarp();
        if (lane == 0) {
            counts[v] = cv2;
            if (!same) counts[other] = co2;
            seen[v >> 5] |= 1u << (v & 31);
            seen[other >> 5] |= 1u << (other & 31);                 // same word as v's: read after the write above
            picked[e >> 5] |= 1u << (e & 31);
            a.out[it] = e;
        }
#pragma unroll
        for (int l = 0; l < LEVELS; ++l) {                          // prefix sums of the siblings at and after the child
            const int iv = v >> (5 * (l + 1)), io = other >> (5 * (l + 1));
            const int kv = (iv & ~31) + lane, ko = (io & ~31) + lane;
            if (kv >= iv) {
                if (dAv) treeA[off[l] + kv] += dAv;
                if (dBv) treeB[off[l] + kv] += dBv;
            }
            if (ko >= io) {                                         // the same lane as above when the groups coincide
                if (dAo) treeA[off[l] + ko] += dAo;
                if (dBo) treeB[off[l] + ko] += dBo;
            }
        }
        WA += (unsigned)(dAv + dAo);
        WB += (unsigned)(dBv + dBo);
        __syncwarp();
    }
}

// One CTA.  All threads initialise the state; warp 0 then makes the S picks.
__global__ void __launch_bounds__(1024, 1) k_sample_edge_neighborhood(SamplerArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const TreeShape& sh = a.shape;
    const int N = (int)a.N;
    // carve: whatever has no global fallback pointer lives in shared memory, in this order
    size_t off = 0;
    auto carve = [&](void* g, size_t bytes) -> void* {
        if (g) return g;
        void* p = smem + off;
        off += (bytes + 15) / 16 * 16;
        return p;
    };
    const int treeN = sh.total > 0 ? sh.total : 1;
    int32_t* tree = (int32_t*)carve(a.g_tree, 2 * (size_t)treeN * 4);
    uint32_t* seen = (uint32_t*)carve(a.g_seen, (size_t)((a.N + 31) / 32) * 4);
    uint32_t* picked = (uint32_t*)carve(a.g_picked, (size_t)((a.E + 31) / 32) * 4);
    int32_t* counts = (int32_t*)carve(a.g_counts, (size_t)a.N * 4);
    int32_t* treeA = tree;
    int32_t* treeB = tree + treeN;

    for (int i = tid; i < (N + 31) / 32; i += blockDim.x) seen[i] = 0u;
    for (long long i = tid; i < (a.E + 31) / 32; i += blockDim.x) picked[i] = 0u;
    for (int v = tid; v < N; v += blockDim.x) counts[v] = a.adj_ptr[v + 1] - a.adj_ptr[v];
    for (int i = tid; i < treeN; i += blockDim.x) treeA[i] = 0;      // nothing is seen yet
    // pull the adjacency into L2 while all 32 warps are still here: every pick then pays L2, not DRAM, latency
    for (long long k = (long long)tid * 16; k < 2 * a.E; k += (long long)blockDim.x * 16)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.adj + k));
    for (long long k = (long long)tid * 32; k <= a.N; k += (long long)blockDim.x * 32)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.adj_ptr + k));
    __syncthreads();
    // tree B bottom-up: raw weights of a level (children with edges left), then the prefix sums inside each group
    for (int l = 0; l < sh.levels; ++l) {
        for (int i = tid; i < sh.size[l]; i += blockDim.x) {
            int s = 0;
            if (l == 0) {
                for (int c = 0; c < 32; ++c) {
                    const long long ch = (long long)i * 32 + c;
                    if (ch < N) s += counts[ch] > 0;
                }
            } else if ((long long)i * 32 + 31 < sh.size[l - 1]) {
                s = treeB[sh.off[l - 1] + i * 32 + 31];              // the child group's total (its last prefix sum)
            }
            treeB[sh.off[l] + i] = s;
        }
        __syncthreads();
        for (int g = tid; g < sh.size[l] / 32; g += blockDim.x) {
            int run = 0;
            for (int c = 0; c < 32; ++c) {
                run += treeB[sh.off[l] + g * 32 + c];
                treeB[sh.off[l] + g * 32 + c] = run;
            }
        }
        __syncthreads();
    }
    if (tid >= 32) return;
#define RGCN_SAMPLER_DISPATCH(C, S, P, TA, TB)                                        \
    switch (sh.levels) {                                                                \
        case 0: sampler_picks<0>(a, C, S, P, TA, TB, lane); break;                      \
        case 1: sampler_picks<1>(a, C, S, P, TA, TB, lane); break;                      \
        case 2: sampler_picks<2>(a, C, S, P, TA, TB, lane); break;                      \
        case 3: sampler_picks<3>(a, C, S, P, TA, TB, lane); break;                      \
        case 4: sampler_picks<4>(a, C, S, P, TA, TB, lane); break;                      \
        case 5: sampler_picks<5>(a, C, S, P, TA, TB, lane); break;                      \
        default: sampler_picks<6>(a, C, S, P, TA, TB, lane); break;                     \
    }
    if (!a.g_tree && !a.g_seen && !a.g_picked && !a.g_counts) {
        // the whole state is in shared memory: re-derive the pointers from the shared array itself, so the compiler
        // sees the address space and this copy of the pick loop uses LDS / STS instead of generic loads
        size_t o = 0;
        int32_t* s_tree = reinterpret_cast<int32_t*>(smem + o);    o += (2 * (size_t)treeN * 4 + 15) / 16 * 16;
        uint32_t* s_seen = reinterpret_cast<uint32_t*>(smem + o);  o += ((size_t)((a.N + 31) / 32) * 4 + 15) / 16 * 16;
        uint32_t* s_picked = reinterpret_cast<uint32_t*>(smem + o); o += ((size_t)((a.E + 31) / 32) * 4 + 15) / 16 * 16;
        int32_t* s_counts = reinterpret_cast<int32_t*>(smem + o);
        RGCN_SAMPLER_DISPATCH(s_counts, s_seen, s_picked, s_tree, s_tree + treeN)
    } else {
        RGCN_SAMPLER_DISPATCH(counts, seen, picked, treeA, treeB)
    }
#undef RGCN_SAMPLER_DISPATCH
}

// adjacency entries before sorting: entry 2 i = (s_i, edge i, other o_i), entry 2 i + 1 = (o_i, edge i, other s_i)
__global__ void k_adj_entries(const int64_t* __restrict__ t, long long E, long long N, uint32_t* __restrict__ keys,
                              uint32_t* __restrict__ ids, int32_t* status) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= E) return;
    long long s = t[3 * i], o = t[3 * i + 2];
    if (s < 0 || s >= N || o < 0 || o >= N) {
        if (status) atomicAdd(status, 1);
        s = o = 0;
    }
    keys[2 * i] = (uint32_t)s; keys[2 * i + 1] = (uint32_t)o;
    ids[2 * i] = (uint32_t)(2 * i); ids[2 * i + 1] = (uint32_t)(2 * i + 1);
}

__global__ void k_adj_decode(const int64_t* __restrict__ t, long long E, long long N, const uint32_t* __restrict__ ids,
                             int2* __restrict__ adj) {
    const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= 2 * E) return;
    const uint32_t id = ids[k];
    const long long e = id >> 1;
    long long other = t[3 * e + ((id & 1) ? 0 : 2)];
    if (other < 0 || other >= N) other = 0;
    adj[k] = make_int2((int)e, (int)other);
}

// adj_ptr[v] = first sorted position with key >= v
__global__ void k_adj_ptr(const uint32_t* __restrict__ keys, long long M, long long N, int32_t* __restrict__ adj_ptr) {
    const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (v > N) return;
    long long lo = 0, hi = M;
    while (lo < hi) { const long long mid = (lo + hi) >> 1; if ((long long)keys[mid] < v) lo = mid + 1; else hi = mid; }
    adj_ptr[v] = (int32_t)lo;
}

size_t pair_sort_temp_bytes(int64_t M) {
    size_t b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)(M > 0 ? M : 1));
    return align_up(b);
}

struct Placement {                                   // which sampler arrays go to shared memory
    bool tree, seen, picked, counts;
    size_t smem;
};

Placement place(int64_t E, int64_t N, const TreeShape& sh) {
    Placement p{};
    size_t used = 0;
    auto fits = [&](size_t bytes) {
        bytes = (bytes + 15) / 16 * 16;
        if (used + bytes > kSmemBudget) return false;
        used += bytes;
        return true;
    };
    p.tree = fits(2 * (size_t)(sh.total > 0 ? sh.total : 1) * 4);
    p.seen = fits((size_t)((N + 31) / 32) * 4);
    p.picked = fits((size_t)((E + 31) / 32) * 4);
    p.counts = fits((size_t)N * 4);
    p.smem = used;
    return p;
}

}  // namespace

extern "C" size_t rgcn_sampler_build_workspace_bytes(int64_t num_edges) {
    const size_t m = 2 * (size_t)(num_edges > 0 ? num_edges : 1);
    return 4 * align_up(m * sizeof(uint32_t)) + pair_sort_temp_bytes(2 * num_edges);
}

extern "C" int rgcn_sampler_build(const int64_t* triples, int64_t E, int64_t N, int32_t* adj_ptr, int32_t* adj,
                                  int32_t* status, void* ws, size_t ws_bytes, rgcn_stream_t stream) {
    RGCN_REQUIRE(E >= 0 && N > 0, RGCN_ERR_ARG, "rgcn_sampler_build: bad sizes");
    RGCN_REQUIRE(adj_ptr, RGCN_ERR_ARG, "rgcn_sampler_build: NULL pointer");
    RGCN_REQUIRE(2 * E < (int64_t)INT32_MAX && N < (int64_t)INT32_MAX, RGCN_ERR_UNSUPPORTED, "rgcn_sampler_build: sizes must fit int32");
    const cudaStream_t st = (cudaStream_t)stream;
    if (E == 0) {
        RGCN_CHECK_CUDA(cudaMemsetAsync(adj_ptr, 0, (size_t)(N + 1) * sizeof(int32_t), st));
        return RGCN_OK;
    }
    RGCN_REQUIRE(triples && adj, RGCN_ERR_ARG, "rgcn_sampler_build: NULL pointer");
    RGCN_REQUIRE((reinterpret_cast<uintptr_t>(adj) & 7) == 0, RGCN_ERR_ARG, "rgcn_sampler_build: adj must be 8-byte aligned");
    RGCN_REQUIRE(ws && ws_bytes >= rgcn_sampler_build_workspace_bytes(E), RGCN_ERR_WORKSPACE, "rgcn_sampler_build: workspace too small");
    const int64_t M = 2 * E;
    Carver c(ws);
    uint32_t* k0 = c.take<uint32_t>((size_t)M);
    uint32_t* k1 = c.take<uint32_t>((size_t)M);
    uint32_t* v0 = c.take<uint32_t>((size_t)M);
    uint32_t* v1 = c.take<uint32_t>((size_t)M);
    size_t temp = pair_sort_temp_bytes(M);
    void* cubws = c.take<char>(temp);
    int bits = 1;
    while (bits < 32 && ((uint64_t)N >> bits) != 0) ++bits;
    RGCN_LAUNCH(k_adj_entries, grid_for(E, 256), 256, 0, st, triples, (long long)E, (long long)N, k0, v0, status);
    RGCN_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(cubws, temp, k0, k1, v0, v1, (int)M, 0, bits, st));   // stable
    rgcn::g_launches.fetch_add((bits + 7) / 8 + 1, std::memory_order_relaxed);
    RGCN_LAUNCH(k_adj_decode, grid_for(M, 256), 256, 0, st, triples, (long long)E, (long long)N, v1, reinterpret_cast<int2*>(adj));
    RGCN_LAUNCH(k_adj_ptr, grid_for(N + 1, 256), 256, 0, st, k1, (long long)M, (long long)N, adj_ptr);
    return RGCN_OK;
}

extern "C" size_t rgcn_sample_workspace_bytes(int64_t num_edges, int64_t num_nodes) {
    const TreeShape sh = tree_shape(num_nodes);
    const Placement p = place(num_edges, num_nodes, sh);
    size_t b = 256;
    if (!p.tree) b += align_up(2 * (size_t)(sh.total > 0 ? sh.total : 1) * 4);
    if (!p.seen) b += align_up((size_t)((num_nodes + 31) / 32) * 4);
    if (!p.picked) b += align_up((size_t)((num_edges + 31) / 32) * 4);
    if (!p.counts) b += align_up((size_t)num_nodes * 4);
    return b;
}

extern "C" int rgcn_sample_edge_neighborhood(const int32_t* adj_ptr, const int32_t* adj, int64_t E, int64_t N, const float* uniforms, int64_t S, int32_t* out_edges,
                                             int32_t* status, void* ws, size_t ws_bytes, rgcn_stream_t stream) {
    RGCN_REQUIRE(E >= 0 && N > 0 && S >= 0, RGCN_ERR_ARG, "rgcn_sample_edge_neighborhood: bad sizes");
    RGCN_REQUIRE(S <= E, RGCN_ERR_ARG, "rgcn_sample_edge_neighborhood: sample_size %lld exceeds the %lld edges", (long long)S, (long long)E);
    if (S == 0) return RGCN_OK;
    RGCN_REQUIRE(adj_ptr && adj && uniforms && out_edges && status, RGCN_ERR_ARG,
                 "rgcn_sample_edge_neighborhood: NULL pointer");
    RGCN_REQUIRE((reinterpret_cast<uintptr_t>(adj) & 7) == 0 && (reinterpret_cast<uintptr_t>(uniforms) & 7) == 0, RGCN_ERR_ARG,
                 "rgcn_sample_edge_neighborhood: adj and uniforms must be 8-byte aligned");
    RGCN_REQUIRE(2 * E < (int64_t)INT32_MAX && N < (int64_t)INT32_MAX, RGCN_ERR_UNSUPPORTED, "rgcn_sample_edge_neighborhood: sizes must fit int32");
    RGCN_REQUIRE(ws && ws_bytes >= rgcn_sample_workspace_bytes(E, N), RGCN_ERR_WORKSPACE, "rgcn_sample_edge_neighborhood: workspace too small");
    const cudaStream_t st = (cudaStream_t)stream;
    SamplerArgs a{};
    a.adj_ptr = adj_ptr; a.adj = reinterpret_cast<const int2*>(adj); a.uniforms = uniforms; a.out = out_edges;
    a.status = status; a.N = N; a.E = E; a.S = S;
    a.shape = tree_shape(N);
    const Placement p = place(E, N, a.shape);
    Carver c(ws);
    c.take<char>(256);
    if (!p.tree) a.g_tree = c.take<int32_t>(2 * (size_t)(a.shape.total > 0 ? a.shape.total : 1));
    if (!p.seen) a.g_seen = c.take<uint32_t>((size_t)((N + 31) / 32));
    if (!p.picked) a.g_picked = c.take<uint32_t>((size_t)((E + 31) / 32));
    if (!p.counts) a.g_counts = c.take<int32_t>((size_t)N);
    // per device / context, so set on every launch (cheap) rather than once per process
    RGCN_CHECK_CUDA(cudaFuncSetAttribute(k_sample_edge_neighborhood, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
    RGCN_LAUNCH(k_sample_edge_neighborhood, 1, 1024, p.smem, st, a);
    return RGCN_OK;
}

// rows[k] = triples[index[k]]  (index int32 or int64): turns picked edge numbers / a dropout permutation into a graph
template <typename I>
__global__ void k_take_rows(const int64_t* __restrict__ t, const I* __restrict__ index, long long n, long long rows,
                            int64_t* __restrict__ out, int32_t* status) {
    const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= n) return;
    long long e = (long long)index[k];
    if (e < 0 || e >= rows) {
        if (status) atomicAdd(status, 1);
        e = 0;
    }
    out[3 * k] = t[3 * e]; out[3 * k + 1] = t[3 * e + 1]; out[3 * k + 2] = t[3 * e + 2];
}

extern "C" int rgcn_take_triples(const int64_t* triples, int64_t num_rows, const void* index, int index_is_int64, int64_t n,
                                 int64_t* out, int32_t* status, rgcn_stream_t stream) {
    RGCN_REQUIRE(num_rows >= 0 && n >= 0, RGCN_ERR_ARG, "rgcn_take_triples: bad sizes");
    if (n == 0) return RGCN_OK;
    RGCN_REQUIRE(triples && index && out && num_rows > 0, RGCN_ERR_ARG, "rgcn_take_triples: NULL pointer or empty source");
    const cudaStream_t st = (cudaStream_t)stream;
    if (index_is_int64)
        RGCN_LAUNCH(k_take_rows<int64_t>, grid_for(n, 256), 256, 0, st, triples, (const int64_t*)index, (long long)n,
                    (long long)num_rows, out, status);
    else
        RGCN_LAUNCH(k_take_rows<int32_t>, grid_for(n, 256), 256, 0, st, triples, (const int32_t*)index, (long long)n,
                    (long long)num_rows, out, status);
    return RGCN_OK;
}
