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

// ------------------------------------------------------------------------------------------------
// Attention core of graph_transformer_pytorch.Attention on the per-frame graphs of the regulariser
// (lib/teatgt.py:316-317): heads x dim_head = 8 x 64, rotary position embedding on q / k by node index,
// per-edge key/value offsets e_ij = A_ij * we + be (edge_dim = 1):
//   sim_ij = (q_i . k_j + A_ij (q_i . we) + q_i . be) / 8,  out_i = sum_j a_ij v_j + (sum_j a_ij A_ij) we + be.
// One warp per (frame, head); lane l owns dims (2l, 2l+1) = one rotary pair; K / V of the frame live in
// shared memory; every dot product is a warp-shuffle reduction.  A is given as the upper-triangular
// predicate matrices of b200vsgg_teat_pair_flags (A = U + U^T).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) graph_attn_core_kernel(const float* __restrict__ qkv, int ld,
                                                              const int32_t* __restrict__ node_off,
                                                              const uint8_t* __restrict__ upper, int nmax,
                                                              const float* __restrict__ we, const float* __restrict__ be,
                                                              int n_frames, __nv_bfloat16* __restrict__ out, int ldo) {
    extern __shared__ float gsm[];
    constexpr int H = 8, DH = 64, INNER = H * DH;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.x, head = warp;
    if (f >= n_frames) return;
    const int r0 = node_off[f], n = node_off[f + 1] - r0;
    float* Ks = gsm + static_cast<size_t>(warp) * 2 * nmax * DH;
    float* Vs = Ks + nmax * DH;
    const uint8_t* U = upper + static_cast<size_t>(f) * nmax * nmax;
    const int d0 = lane * 2;
    const float inv_freq = __powf(10000.f, -static_cast<float>(d0) / DH);
    const float2 w2 = *reinterpret_cast<const float2*>(we + head * DH + d0);
    const float2 b2 = *reinterpret_cast<const float2*>(be + head * DH + d0);
    for (int j = 0; j < n; ++j) {
        const float* row = qkv + static_cast<size_t>(r0 + j) * ld + head * DH + d0;
        float2 kk = *reinterpret_cast<const float2*>(row + INNER);
        const float2 vv = *reinterpret_cast<const float2*>(row + 2 * INNER);
        float sn, cs;
        __sincosf(j * inv_freq, &sn, &cs);
        const float kx = kk.x * cs - kk.y * sn, ky = kk.y * cs + kk.x * sn;
        Ks[j * DH + d0] = kx; Ks[j * DH + d0 + 1] = ky;
        Vs[j * DH + d0] = vv.x; Vs[j * DH + d0 + 1] = vv.y;
    }
    __syncwarp();
    for (int i = 0; i < n; ++i) {
        const float2 qq = *reinterpret_cast<const float2*>(qkv + static_cast<size_t>(r0 + i) * ld + head * DH + d0);
        float sn, cs;
        __sincosf(i * inv_freq, &sn, &cs);
        const float qx = qq.x * cs - qq.y * sn, qy = qq.y * cs + qq.x * sn;
        const float qw = warp_sum(qx * w2.x + qy * w2.y);
        const float qb = warp_sum(qx * b2.x + qy * b2.y);
        float mx = -INFINITY;
        float my_score = -INFINITY;                     // lane j keeps score j
#pragma unroll 1
        for (int j = 0; j < n; ++j) {
            const float dot = warp_sum(qx * Ks[j * DH + d0] + qy * Ks[j * DH + d0 + 1]);
            const float a = (i < j ? U[i * nmax + j] : (i > j ? U[j * nmax + i] : 0)) ? 1.f : 0.f;
            const float s = (dot + a * qw + qb) * 0.125f;
            if (lane == j) my_score = s;
            mx = fmaxf(mx, s);
        }
        const float e = lane < n ? __expf(my_score - mx) : 0.f;
        const float den = warp_sum(e);
        const float p_l = e / den;                     // probability of key `lane`
        float ox = 0.f, oy = 0.f, pa = 0.f;
#pragma unroll 1
        for (int j = 0; j < n; ++j) {
            const float p = __shfl_sync(0xffffffffu, p_l, j);
            ox = fmaf(p, Vs[j * DH + d0], ox);
            oy = fmaf(p, Vs[j * DH + d0 + 1], oy);
            const float a = (i < j ? U[i * nmax + j] : (i > j ? U[j * nmax + i] : 0)) ? 1.f : 0.f;
            pa = fmaf(p, a, pa);
        }
        ox += pa * w2.x + b2.x;
        oy += pa * w2.y + b2.y;
        *reinterpret_cast<__nv_bfloat162*>(out + static_cast<size_t>(r0 + i) * ldo + head * DH + d0) =
            __floats2bfloat162_rn(ox, oy);
    }
}

// GatedResidual: g = sigmoid(out . wa + res . wb) with wa = w1 + w3, wb = w2 - w3;  x = out*g + res*(1-g).
// One warp per row, float4 accesses; writes x in place of `res`.
__global__ void gated_residual_kernel(const float* __restrict__ o, float* __restrict__ res, const float* __restrict__ w,
                                      int rows, int dim) {
    const int wpb = blockDim.x >> 5, lane = threadIdx.x & 31;
    const bool vec = (dim & 3) == 0;       // rows are 16-byte aligned: float4 accesses
    for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
        const float* op = o + static_cast<size_t>(r) * dim;
        float* rp = res + static_cast<size_t>(r) * dim;
        float acc = 0.f;
        if (vec) {
            const int nv = dim >> 2;
            const float4* o4 = reinterpret_cast<const float4*>(op);
            float4* r4 = reinterpret_cast<float4*>(rp);
            const float4* w1 = reinterpret_cast<const float4*>(w);
            const float4* w2 = reinterpret_cast<const float4*>(w + dim);
            const float4* w3 = reinterpret_cast<const float4*>(w + 2 * dim);
            for (int c = lane; c < nv; c += 32) {
                const float4 a = o4[c], b = r4[c], x1 = __ldg(w1 + c), x2 = __ldg(w2 + c), x3 = __ldg(w3 + c);
                acc += a.x * (x1.x + x3.x) + b.x * (x2.x - x3.x) + a.y * (x1.y + x3.y) + b.y * (x2.y - x3.y) +
                       a.z * (x1.z + x3.z) + b.z * (x2.z - x3.z) + a.w * (x1.w + x3.w) + b.w * (x2.w - x3.w);
            }
            acc = warp_sum(acc);
            const float g = 1.f / (1.f + __expf(-acc)), h = 1.f - g;
            for (int c = lane; c < nv; c += 32) {
                const float4 a = o4[c], b = r4[c];
                r4[c] = make_float4(a.x * g + b.x * h, a.y * g + b.y * h, a.z * g + b.z * h, a.w * g + b.w * h);
            }
        } else {
            for (int c = lane; c < dim; c += 32) {
                const float w1 = __ldg(w + c), w2 = __ldg(w + dim + c), w3 = __ldg(w + 2 * dim + c);
                acc += op[c] * (w1 + w3) + rp[c] * (w2 - w3);
            }
            acc = warp_sum(acc);
            const float g = 1.f / (1.f + __expf(-acc));
            for (int c = lane; c < dim; c += 32) rp[c] = op[c] * g + rp[c] * (1.f - g);
        }
    }
}

}  // namespace vsgg

using namespace vsgg;

extern "C" int b200vsgg_graph_attn_core(const float* qkv, int32_t ld, const int32_t* node_off, const uint8_t* upper,
                                        int32_t nmax, const float* we, const float* be, int32_t n_frames, void* out,
                                        int32_t ldo, void* stream) {
    if (!qkv || !node_off || !upper || !we || !be || !out || nmax <= 0 || nmax > 32 || (ld & 1) || (ldo & 1))
        return set_error(B200VSGG_ERR_BAD_ARG, "graph_attn_core: bad arg (<= 32 nodes per frame)");
    if (n_frames == 0) return 0;
    const size_t smem = 8ull * 2 * nmax * 64 * sizeof(float);
    static size_t cur = 48 * 1024;
    if (smem > cur) {
        cudaError_t e = cudaFuncSetAttribute(graph_attn_core_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
        cur = smem;
    }
    graph_attn_core_kernel<<<n_frames, 256, smem, (cudaStream_t)stream>>>(qkv, ld, node_off, upper, nmax, we, be, n_frames,
                                                                          (__nv_bfloat16*)out, ldo);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_gated_residual(const float* o, float* res, const float* w, int32_t rows, int32_t dim,
                                       void* stream) {
    if (!o || !res || !w || dim <= 0) return set_error(B200VSGG_ERR_BAD_ARG, "gated_residual: bad arg");
    if (rows == 0) return 0;
    int grid = (rows + 7) / 8;
    if (grid > 148 * 16) grid = 148 * 16;
    gated_residual_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(o, res, w, rows, dim);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

// GlobalAttentionPooling over compact node rows (dgl.nn.GlobalAttentionPooling as lib/teatgt.py:319-320 uses it): per frame
// gate_i = w . x_i + b, a = softmax_i(gate), out = sum_i a_i x_i.  One CTA per frame (<= 64 nodes): a warp per node for
// the gate dot products, then every thread owns output columns.
namespace vsgg {
constexpr int POOL_MAX_NODES = 64;
__global__ void __launch_bounds__(256)
attn_pool_kernel(const float* __restrict__ x, int D, const int32_t* __restrict__ node_off, const float* __restrict__ w,
                 const float* __restrict__ b, float* __restrict__ out) {
    __shared__ float gate[POOL_MAX_NODES];
    const int f = blockIdx.x, r0 = node_off[f], n = node_off[f + 1] - r0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = warp; i < n; i += 8) {
        const float* xi = x + static_cast<size_t>(r0 + i) * D;
        float acc = 0.f;
        for (int d = lane; d < D; d += 32) acc = fmaf(xi[d], __ldg(w + d), acc);
        acc = warp_sum(acc);
        if (lane == 0) gate[i] = acc + __ldg(b);
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int i = 0; i < n; ++i) mx = fmaxf(mx, gate[i]);
    float den = 0.f;
    for (int i = 0; i < n; ++i) den += __expf(gate[i] - mx);
    const float inv = n > 0 ? 1.f / den : 0.f;
    for (int d = threadIdx.x; d < D; d += 256) {
        float acc = 0.f;
        for (int i = 0; i < n; ++i) acc = fmaf(__expf(gate[i] - mx) * inv, x[static_cast<size_t>(r0 + i) * D + d], acc);
        out[static_cast<size_t>(f) * D + d] = acc;
    }
}
}  // namespace vsgg

extern "C" int b200vsgg_attn_pool(const float* x, int32_t d, const int32_t* node_off, int32_t n_frames, int32_t max_nodes,
                                  const float* w, const float* b, float* out, void* stream) {
    if (!x || !node_off || !w || !b || !out || d <= 0 || max_nodes > vsgg::POOL_MAX_NODES)
        return set_error(B200VSGG_ERR_BAD_ARG, "attn_pool: bad arg (<= 64 nodes per frame)");
    if (n_frames <= 0) return 0;
    vsgg::attn_pool_kernel<<<n_frames, 256, 0, (cudaStream_t)stream>>>(x, d, node_off, w, b, out);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

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
