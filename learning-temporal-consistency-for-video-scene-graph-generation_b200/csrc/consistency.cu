// Pairwise graph temporal-consistency reduction (lib/teatgt.py:325-334 of the reference): for every
// frame pair u < v of a clip, KLDivLoss(batchmean)(log_softmax(g_u), softmax(g_v)) / (v - u) over the
// per-frame graph embeddings g [frames, D].  One warp per pair: both rows are read once with 128-bit
// loads, the two log-sum-exps and the weighted difference are warp-shuffle reductions.
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/b200vsgg.h"
#include "common.cuh"

namespace vsgg {

__global__ void consistency_kl_kernel(const float* __restrict__ g, int D, const int32_t* __restrict__ pair_u,
                                      const int32_t* __restrict__ pair_v, int n_pairs, float* __restrict__ out) {
    const int warps_per_block = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (int p = blockIdx.x * warps_per_block + (threadIdx.x >> 5); p < n_pairs; p += gridDim.x * warps_per_block) {
        const int u = __ldg(pair_u + p), v = __ldg(pair_v + p);
        const float* gu = g + static_cast<size_t>(u) * D;
        const float* gv = g + static_cast<size_t>(v) * D;
        float mu = -INFINITY, mv = -INFINITY;
        for (int d = lane; d < D; d += 32) {
            mu = fmaxf(mu, gu[d]);
            mv = fmaxf(mv, gv[d]);
        }
        mu = warp_max(mu);
        mv = warp_max(mv);
        float su = 0.f, sv = 0.f;
        for (int d = lane; d < D; d += 32) {
            su += __expf(gu[d] - mu);
            sv += __expf(gv[d] - mv);
        }
        su = warp_sum(su);
        sv = warp_sum(sv);
        const float lse_u = mu + __logf(su), lse_v = mv + __logf(sv);
        // KL(q || p) with q = softmax(g_v), log p = log_softmax(g_u):  sum_d q_d (log q_d - log p_d)
        float kl = 0.f;
        for (int d = lane; d < D; d += 32) {
            const float lq = gv[d] - lse_v, lp = gu[d] - lse_u;
            kl += __expf(lq) * (lq - lp);
        }
        kl = warp_sum(kl);
        if (lane == 0) out[p] = kl / static_cast<float>(v - u);
    }
}

}  // namespace vsgg

using namespace vsgg;

extern "C" int b200vsgg_consistency_kl(const float* g, int32_t d, const int32_t* pair_u, const int32_t* pair_v,
                                       int32_t n_pairs, float* out, void* stream) {
    if (!g || !pair_u || !pair_v || !out || d <= 0) return set_error(B200VSGG_ERR_BAD_ARG, "consistency_kl: bad arg");
    if (n_pairs == 0) return 0;
    int grid = (n_pairs + 7) / 8;
    if (grid > 148 * 16) grid = 148 * 16;
    consistency_kl_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g, d, pair_u, pair_v, n_pairs, out);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}
