// Trainer loss block of the relation heads, forward AND gradient in one launch:
//   TEMPURA_train.py:181-206 / TEATGT_train.py:153-175 of the reference —
//   attention:  nn.CrossEntropyLoss(reduction='none') applied to the (already soft-maxed) distribution [N,3]
//   spatial / contacting:  nn.BCELoss(reduction='none') on the sigmoid mixtures [N,6] / [N,17] against multi-hot
//   labels built by a Python loop (:185-187), each followed by .mean().
// Labels are consumed as the ragged lists the dataloader produces (CSR offsets + class ids), so the multi-hot
// matrices never exist; `row_w` carries the reduction (1/N for one video, 1/(N_v * V) for a batch = the average of
// the per-video means).  HBM-bound row kernel: 26 floats in, 26 floats out per pair; one thread per pair, block
// reduction of the three partial sums, one atomicAdd per block and loss.
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/b200vsgg.h"
#include "common.cuh"

namespace vsgg {

constexpr int LOSS_MAX_C = 32;

// nn.BCELoss: log terms clamped at -100 (forward); backward (p - t) / max(p (1 - p), 1e-12) like torch
__device__ __forceinline__ float bce_row(const float* __restrict__ p, float* __restrict__ dp, int C, const int32_t* off,
                                         const int32_t* idx, const float* dense, int row, float w) {
    uint32_t hot = 0u;
    if (off) {
        for (int e = off[row]; e < off[row + 1]; ++e) hot |= 1u << (idx[e] & 31);   // class ids are < C <= 32
    } else {
        for (int c = 0; c < C; ++c) hot |= (dense[static_cast<size_t>(row) * C + c] > 0.5f ? 1u : 0u) << c;
    }
    float loss = 0.f;
    const float wc = w / C;
    for (int c = 0; c < C; ++c) {
        const float x = p[static_cast<size_t>(row) * C + c];
        const float t = (hot >> c) & 1u ? 1.f : 0.f;
        loss -= t * fmaxf(logf(x), -100.f) + (1.f - t) * fmaxf(log1pf(-x), -100.f);
        if (dp) dp[static_cast<size_t>(row) * C + c] = wc * (x - t) / fmaxf(x * (1.f - x), 1e-12f);
    }
    return loss * wc;
}

__global__ void __launch_bounds__(256) rel_loss_kernel(
    const float* __restrict__ att, const float* __restrict__ spa, const float* __restrict__ con, int n, int ca, int cs, int cc,
    const int64_t* __restrict__ att_label, const float* __restrict__ spa_dense, const float* __restrict__ con_dense,
    const int32_t* __restrict__ spa_off, const int32_t* __restrict__ spa_idx, const int32_t* __restrict__ con_off,
    const int32_t* __restrict__ con_idx, const float* __restrict__ row_w, float* __restrict__ losses,
    float* __restrict__ d_att, float* __restrict__ d_spa, float* __restrict__ d_con) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    float la = 0.f, ls = 0.f, lc = 0.f;
    if (row < n) {
        const float w = row_w[row];
        // cross entropy with the probabilities used as logits: -x_y + logsumexp(x)
        float x[LOSS_MAX_C], mx = -INFINITY;
        for (int c = 0; c < ca; ++c) { x[c] = att[static_cast<size_t>(row) * ca + c]; mx = fmaxf(mx, x[c]); }
        float den = 0.f;
        for (int c = 0; c < ca; ++c) den += expf(x[c] - mx);
        const int y = min(max(static_cast<int>(att_label[row]), 0), ca - 1);   // never index outside x[]
        la = w * (mx + logf(den) - x[y]);
        if (d_att)
            for (int c = 0; c < ca; ++c)
                d_att[static_cast<size_t>(row) * ca + c] = w * (expf(x[c] - mx) / den - (c == y ? 1.f : 0.f));
        ls = bce_row(spa, d_spa, cs, spa_off, spa_idx, spa_dense, row, w);
        lc = bce_row(con, d_con, cc, con_off, con_idx, con_dense, row, w);
    }
    la = warp_sum(la); ls = warp_sum(ls); lc = warp_sum(lc);
    __shared__ float red[3][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = la; red[1][warp] = ls; red[2][warp] = lc; }
    __syncthreads();
    if (threadIdx.x < 3) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += red[threadIdx.x][i];
        atomicAdd(losses + threadIdx.x, s);
    }
}

}  // namespace vsgg

using namespace vsgg;

extern "C" int b200vsgg_rel_loss(const float* att, const float* spa, const float* con, int32_t n, int32_t ca, int32_t cs,
                                 int32_t cc, const int64_t* att_label, const float* spa_dense, const float* con_dense,
                                 const int32_t* spa_off, const int32_t* spa_idx, const int32_t* con_off,
                                 const int32_t* con_idx, const float* row_w, float* losses, float* d_att, float* d_spa,
                                 float* d_con, void* stream) {
    if (!att || !spa || !con || !att_label || !row_w || !losses || ca < 1 || ca > LOSS_MAX_C || cs < 1 || cs > LOSS_MAX_C ||
        cc < 1 || cc > LOSS_MAX_C || (!spa_dense && !(spa_off && spa_idx)) || (!con_dense && !(con_off && con_idx)))
        return set_error(B200VSGG_ERR_BAD_ARG, "rel_loss: bad arg (class counts <= 32; labels dense or CSR)");
    if (n == 0) return 0;
    rel_loss_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(att, spa, con, n, ca, cs, cc, att_label, spa_dense,
                                                                        con_dense, spa_off, spa_idx, con_off, con_idx,
                                                                        row_w, losses, d_att, d_spa, d_con);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}
