// Relationship contrastive loss of the trainers (TEMPURA_train.py:103,209-212, TEATGT_train.py:81,176-179):
//   con_loss = pytorch_metric_learning.losses.ContrastiveLoss(pos_margin=0, neg_margin=1)
//   losses['spatial_con_loss'] = 0.2 * con_loss(spatial_distribution, argmax(spatial_label, 1))     (same for contacting)
// The package is absent from the reference tree and unversioned (PARITY UNPINNED); restated from its published
// algorithm: embeddings are L2-normalised, d_ij = |e_i - e_j| over ALL pairs i != j of the call (the
// reference's call = one video), positives (equal labels) cost max(d - pos_margin, 0), negatives max(neg_margin - d, 0),
// each group is averaged over its NON-ZERO entries (AvgNonZeroReducer) and the two means are added.
//
// One CTA per video segment: the normalised embeddings (n x C fp32, C = 6 / 17) live in shared memory; pass 1 reduces the
// four group statistics over the n(n-1)/2 unordered pairs, pass 2 (thread = row) accumulates dL/de_i over the row's
// partners and chains through the normalisation — loss AND gradient from one launch, nothing of size n^2 anywhere.
// HBM-bound by construction (n*C*4 bytes in, the same out); the n^2*C flops are register/shared-memory work.
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/b200vsgg.h"
#include "common.cuh"

namespace vsgg {

constexpr int CL_THREADS = 256;
constexpr int CL_CMAX = 32;

__global__ void __launch_bounds__(CL_THREADS) contrastive_loss_kernel(
    const float* __restrict__ x, const int32_t* __restrict__ label, const int32_t* __restrict__ seg_off, int C,
    float pos_margin, float neg_margin, float* __restrict__ loss, float* __restrict__ dx, const float* __restrict__ gscale) {
    extern __shared__ float cl_sm[];
    const int v = blockIdx.x;
    const int r0 = seg_off[v], n = seg_off[v + 1] - r0;
    float* e = cl_sm;                                   // [n][C] normalised embeddings
    float* inv_norm = e + static_cast<size_t>(n) * C;   // [n]
    int* lab = reinterpret_cast<int*>(inv_norm + n);    // [n]
    __shared__ float red[4][CL_THREADS / 32];
    __shared__ float stat[4];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int i = t; i < n; i += CL_THREADS) {
        const float* xi = x + static_cast<size_t>(r0 + i) * C;
        float s = 0.f;
        for (int c = 0; c < C; ++c) s = fmaf(xi[c], xi[c], s);
        const float inv = 1.f / fmaxf(sqrtf(s), 1e-12f);          // F.normalize(eps = 1e-12)
        inv_norm[i] = inv;
        for (int c = 0; c < C; ++c) e[i * C + c] = xi[c] * inv;
        lab[i] = label[r0 + i];
    }
    __syncthreads();
    // ---- pass 1: unordered pairs (i < j), linearised and strided over the CTA
    float ps = 0.f, pc = 0.f, ns = 0.f, nc = 0.f;
    const long long pairs = static_cast<long long>(n) * (n - 1) / 2;
    for (long long p = t; p < pairs; p += CL_THREADS) {
        // row i of the strict upper triangle that contains linear index p
        int i = static_cast<int>((2.0 * n - 1.0 - sqrt((2.0 * n - 1.0) * (2.0 * n - 1.0) - 8.0 * static_cast<double>(p))) * 0.5);
        long long base = static_cast<long long>(i) * (2 * n - i - 1) / 2;
        if (base > p) { --i; base = static_cast<long long>(i) * (2 * n - i - 1) / 2; }
        else if (base + (n - i - 1) <= p) { base += n - i - 1; ++i; }
        const int j = i + 1 + static_cast<int>(p - base);
        float d2 = 0.f;
        for (int c = 0; c < C; ++c) { const float df = e[i * C + c] - e[j * C + c]; d2 = fmaf(df, df, d2); }
        const float d = sqrtf(d2);
        if (lab[i] == lab[j]) { const float l = d - pos_margin; if (l > 0.f) { ps += l; pc += 1.f; } }
        else { const float l = neg_margin - d; if (l > 0.f) { ns += l; nc += 1.f; } }
    }
    ps = warp_sum(ps); pc = warp_sum(pc); ns = warp_sum(ns); nc = warp_sum(nc);
    if (lane == 0) { red[0][warp] = ps; red[1][warp] = pc; red[2][warp] = ns; red[3][warp] = nc; }
    __syncthreads();
    if (t < 4) {
        float s = 0.f;
        for (int w = 0; w < CL_THREADS / 32; ++w) s += red[t][w];
        stat[t] = s;
    }
    __syncthreads();
    const float pos_cnt = stat[1], neg_cnt = stat[3];
    if (t == 0) loss[v] = (pos_cnt > 0.f ? stat[0] / pos_cnt : 0.f) + (neg_cnt > 0.f ? stat[2] / neg_cnt : 0.f);
    if (dx == nullptr) return;
    // ---- pass 2: dL/de_i = sum_j w_ij (e_i - e_j) / d_ij, w = +1/pos_cnt (active positives), -1/neg_cnt (active
    //      negatives); an unordered pair contributes to both of its rows.  Then through e = x / |x|.
    const float wp = pos_cnt > 0.f ? 1.f / pos_cnt : 0.f, wn = neg_cnt > 0.f ? 1.f / neg_cnt : 0.f;
    const float gs = gscale ? gscale[0] : 1.f;
    for (int i = t; i < n; i += CL_THREADS) {
        float g[CL_CMAX], ei[CL_CMAX];
#pragma unroll
        for (int c = 0; c < CL_CMAX; ++c) { g[c] = 0.f; ei[c] = c < C ? e[i * C + c] : 0.f; }
        const int li = lab[i];
        for (int j = 0; j < n; ++j) {
            if (j == i) continue;
            float d2 = 0.f;
#pragma unroll
            for (int c = 0; c < CL_CMAX; ++c)
                if (c < C) { const float df = ei[c] - e[j * C + c]; d2 = fmaf(df, df, d2); }
            const float d = sqrtf(d2);
            float w = 0.f;
            if (li == lab[j]) { if (d - pos_margin > 0.f) w = wp; }
            else if (neg_margin - d > 0.f) w = -wn;
            if (w != 0.f && d > 0.f) {
                const float f = w / d;
#pragma unroll
                for (int c = 0; c < CL_CMAX; ++c)
                    if (c < C) g[c] = fmaf(f, ei[c] - e[j * C + c], g[c]);
            }
        }
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < CL_CMAX; ++c)
            if (c < C) dot = fmaf(g[c], ei[c], dot);
        const float inv = inv_norm[i] * gs;
        float* o = dx + static_cast<size_t>(r0 + i) * C;
#pragma unroll
        for (int c = 0; c < CL_CMAX; ++c)
            if (c < C) o[c] = (g[c] - dot * ei[c]) * inv;       // (I - e e^T) g / |x|
    }
}

}  // namespace vsgg

using namespace vsgg;

extern "C" int b200vsgg_contrastive_loss(const float* x, const int32_t* label, const int32_t* seg_off, int32_t n_seg,
                                         int32_t max_rows, int32_t C, float pos_margin, float neg_margin, float* loss,
                                         float* dx, const float* grad_scale, void* stream) {
    if (!x || !label || !seg_off || !loss || n_seg < 0 || C < 1 || C > CL_CMAX || max_rows < 0)
        return set_error(B200VSGG_ERR_BAD_ARG, "contrastive_loss: bad arg (C <= 32)");
    if (n_seg == 0) return 0;
    const size_t smem = static_cast<size_t>(max_rows) * (C + 2) * sizeof(float);
    if (smem > 200 * 1024)
        return set_error(B200VSGG_ERR_BAD_ARG, "contrastive_loss: a segment's embeddings exceed 200 KB of shared memory");
    static size_t cur = 48 * 1024;
    if (smem > cur) {
        cudaError_t e = cudaFuncSetAttribute(contrastive_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
        cur = smem;
    }
    contrastive_loss_kernel<<<n_seg, CL_THREADS, smem, (cudaStream_t)stream>>>(x, label, seg_off, C, pos_margin, neg_margin,
                                                                             loss, dx, grad_scale);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}
