// Error reporting, version, launch counter and the host-side relation shard planner.
#include <stdarg.h>
#include <algorithm>
#include <numeric>
#include <vector>
#include "common.cuh"

namespace rgcn {
static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace rgcn

extern "C" const char* rgcn_last_error(void) { return rgcn::g_err; }
extern "C" int rgcn_abi_version(void) { return RGCN_ABI_VERSION; }
extern "C" int64_t rgcn_launch_count(void) { return rgcn::g_launches.load(); }

// Longest-processing-time greedy: heaviest relation first onto the lightest rank.
// Ties broken by rank id then relation id so every rank computes the same plan.
extern "C" int rgcn_shard_plan(const int64_t* rel_nnz, int64_t num_rels, int32_t world, int32_t* rel_to_rank) {
    RGCN_REQUIRE(rel_nnz && rel_to_rank && num_rels >= 0 && world >= 1, RGCN_ERR_ARG, "rgcn_shard_plan: bad arguments");
    std::vector<int64_t> order(num_rels);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return rel_nnz[a] > rel_nnz[b]; });
    std::vector<int64_t> load(world, 0);
    for (int64_t r : order) {
        int best = 0;
        for (int k = 1; k < world; ++k)
            if (load[k] < load[best]) best = k;
        rel_to_rank[r] = best;
        load[best] += rel_nnz[r];
    }
    return RGCN_OK;
}
