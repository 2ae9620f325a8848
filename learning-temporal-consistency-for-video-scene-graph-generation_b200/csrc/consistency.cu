// Pairwise graph temporal-consistency reduction (lib/teatgt.py:325-334 of the reference): for every
// frame pair u < v of a clip, KLDivLoss(batchmean)(log_softmax(g_u), softmax(g_v)) / (v - u) over the
// per-frame graph embeddings g [frames, D].  One warp per pair: both rows are read once with 128-bit
// loads, the two log-sum-exps and the weighted difference are warp-shuffle reductions.
#include <cuda_runtime.h>
#include <algorithm>
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
template <bool OUT_F32>
__global__ void __launch_bounds__(256) graph_attn_core_kernel(const float* __restrict__ qkv, int ld,
                                                              const int32_t* __restrict__ node_off,
                                                              const uint8_t* __restrict__ upper, int nmax,
                                                              const float* __restrict__ we, const float* __restrict__ be,
                                                              int n_frames, void* __restrict__ out_, int ldo) {
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
        if (OUT_F32)
            *reinterpret_cast<float2*>(reinterpret_cast<float*>(out_) + static_cast<size_t>(r0 + i) * ldo + head * DH + d0) =
                make_float2(ox, oy);
        else
            *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(out_) + static_cast<size_t>(r0 + i) * ldo +
                                               head * DH + d0) = __floats2bfloat162_rn(ox, oy);
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

static int graph_attn_core_launch(const float* qkv, int32_t ld, const int32_t* node_off, const uint8_t* upper, int32_t nmax,
                                  const float* we, const float* be, int32_t n_frames, void* out, int32_t ldo, bool out_f32,
                                  void* stream) {
    if (!qkv || !node_off || !upper || !we || !be || !out || nmax <= 0 || nmax > 32 || (ld & 1) || (ldo & 1))
        return set_error(B200VSGG_ERR_BAD_ARG, "graph_attn_core: bad arg (<= 32 nodes per frame)");
    if (n_frames == 0) return 0;
    const size_t smem = 8ull * 2 * nmax * 64 * sizeof(float);
    static size_t cur[2] = {48 * 1024, 48 * 1024};
    if (smem > cur[out_f32]) {
        cudaError_t e = out_f32 ? cudaFuncSetAttribute(graph_attn_core_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                : cudaFuncSetAttribute(graph_attn_core_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
        cur[out_f32] = smem;
    }
    if (out_f32)
        graph_attn_core_kernel<true><<<n_frames, 256, smem, (cudaStream_t)stream>>>(qkv, ld, node_off, upper, nmax, we, be, n_frames, out, ldo);
    else
        graph_attn_core_kernel<false><<<n_frames, 256, smem, (cudaStream_t)stream>>>(qkv, ld, node_off, upper, nmax, we, be, n_frames, out, ldo);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_graph_attn_core(const float* qkv, int32_t ld, const int32_t* node_off, const uint8_t* upper,
                                        int32_t nmax, const float* we, const float* be, int32_t n_frames, void* out,
                                        int32_t ldo, void* stream) {
    return graph_attn_core_launch(qkv, ld, node_off, upper, nmax, we, be, n_frames, out, ldo, false, stream);
}

extern "C" int b200vsgg_graph_attn_core_f32(const float* qkv, int32_t ld, const int32_t* node_off, const uint8_t* upper,
                                            int32_t nmax, const float* we, const float* be, int32_t n_frames, float* out,
                                            int32_t ldo, void* stream) {
    return graph_attn_core_launch(qkv, ld, node_off, upper, nmax, we, be, n_frames, out, ldo, true, stream);
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

// ================================================================================================
// Backward kernels of the DIFFERENTIABLE consistency mode (the reference detaches both loss vectors,
// lib/teatgt.py:350-351; with `differentiable_consistency=True` the semantic-branch loss carries gradients into the
// regulariser's own parameters and into the encoder's hidden rows).  Each kernel recomputes its forward quantities from
// the saved inputs; parameter gradients that are sums over rows go through row-weight vectors + b200vsgg_weighted_colsum
// instead of one atomic per element.
// ================================================================================================
namespace vsgg {

// d/dg of out[p] = KL(softmax(g_v) || softmax(g_u)) / (v - u), scaled by gout[p] (0 for pairs the filter dropped):
//   dg_u += (p - q) * s,   dg_v += q * ((log q - log p) - KL) * s,   s = gout[p] / (v - u).   One warp per pair.
__global__ void consistency_kl_bwd_kernel(const float* __restrict__ g, int D, const int32_t* __restrict__ pair_u,
                                          const int32_t* __restrict__ pair_v, const float* __restrict__ gout, int n_pairs,
                                          float* __restrict__ dg) {
    const int wpb = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (int p = blockIdx.x * wpb + (threadIdx.x >> 5); p < n_pairs; p += gridDim.x * wpb) {
        const float go = gout[p];
        if (go == 0.f) continue;
        const int u = __ldg(pair_u + p), v = __ldg(pair_v + p);
        const float* gu = g + static_cast<size_t>(u) * D;
        const float* gv = g + static_cast<size_t>(v) * D;
        float mu = -INFINITY, mv = -INFINITY;
        for (int d = lane; d < D; d += 32) { mu = fmaxf(mu, gu[d]); mv = fmaxf(mv, gv[d]); }
        mu = warp_max(mu);
        mv = warp_max(mv);
        float su = 0.f, sv = 0.f;
        for (int d = lane; d < D; d += 32) { su += __expf(gu[d] - mu); sv += __expf(gv[d] - mv); }
        su = warp_sum(su);
        sv = warp_sum(sv);
        const float lse_u = mu + __logf(su), lse_v = mv + __logf(sv);
        float kl = 0.f;
        for (int d = lane; d < D; d += 32) {
            const float lq = gv[d] - lse_v, lp = gu[d] - lse_u;
            kl += __expf(lq) * (lq - lp);
        }
        kl = warp_sum(kl);
        const float sc = go / static_cast<float>(v - u);
        for (int d = lane; d < D; d += 32) {
            const float lq = gv[d] - lse_v, lp = gu[d] - lse_u;
            const float q = __expf(lq), pp = __expf(lp);
            atomicAdd(dg + static_cast<size_t>(u) * D + d, (pp - q) * sc);
            atomicAdd(dg + static_cast<size_t>(v) * D + d, q * ((lq - lp) - kl) * sc);
        }
    }
}

// Backward of attn_pool: dx_i = a_i dout_f + dgate_i w,  dgate_i = a_i (dout_f . x_i - dout_f . out_f);  dgate [rows] is
// written out so that dw = sum_i dgate_i x_i runs as one weighted column sum.  One CTA per frame.
__global__ void __launch_bounds__(256)
attn_pool_bwd_kernel(const float* __restrict__ x, int D, const int32_t* __restrict__ node_off, const float* __restrict__ w,
                     const float* __restrict__ b, const float* __restrict__ dout, float* __restrict__ dx,
                     float* __restrict__ dgate) {
    __shared__ float gate[POOL_MAX_NODES], dots[POOL_MAX_NODES];
    const int f = blockIdx.x, r0 = node_off[f], n = node_off[f + 1] - r0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* df = dout + static_cast<size_t>(f) * D;
    for (int i = warp; i < n; i += 8) {
        const float* xi = x + static_cast<size_t>(r0 + i) * D;
        float acc = 0.f, dd = 0.f;
        for (int d = lane; d < D; d += 32) {
            acc = fmaf(xi[d], __ldg(w + d), acc);
            dd = fmaf(xi[d], df[d], dd);
        }
        acc = warp_sum(acc);
        dd = warp_sum(dd);
        if (lane == 0) { gate[i] = acc + __ldg(b); dots[i] = dd; }
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int i = 0; i < n; ++i) mx = fmaxf(mx, gate[i]);
    float den = 0.f;
    for (int i = 0; i < n; ++i) den += __expf(gate[i] - mx);
    const float inv = n > 0 ? 1.f / den : 0.f;
    float dot_out = 0.f;                                   // dout . out_f = sum_i a_i (dout . x_i)
    for (int i = 0; i < n; ++i) dot_out = fmaf(__expf(gate[i] - mx) * inv, dots[i], dot_out);
    for (int i = 0; i < n; ++i) {
        const float a = __expf(gate[i] - mx) * inv;
        const float dgi = a * (dots[i] - dot_out);
        if (threadIdx.x == 0) dgate[r0 + i] = dgi;
        const size_t ro = static_cast<size_t>(r0 + i) * D;
        for (int d = threadIdx.x; d < D; d += 256) dx[ro + d] = fmaf(a, df[d], dgi * __ldg(w + d));
    }
}

// out[c] += sum_r wgt[r] * x[r, c]  (x fp32 or bf16 rows): parameter gradients that are row-weighted column sums.
template <typename T>
__global__ void __launch_bounds__(256) weighted_colsum_kernel(const T* __restrict__ x, int ld, int rows, int cols,
                                                              const float* __restrict__ wgt, float* __restrict__ out) {
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int col = blockIdx.x * 64 + tx;
    const int per = (rows + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
    float acc = 0.f;
    if (col < cols)
        for (int r = r0 + ty; r < r1; r += 4) {
            float v;
            if constexpr (sizeof(T) == 2) v = __bfloat162float(x[static_cast<size_t>(r) * ld + col]);
            else v = x[static_cast<size_t>(r) * ld + col];
            acc = fmaf(__ldg(wgt + r), v, acc);
        }
    __shared__ float sm[4][64];
    sm[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && col < cols) atomicAdd(out + col, sm[0][tx] + sm[1][tx] + sm[2][tx] + sm[3][tx]);
}

// Backward of GatedResidual (x = o g + r (1 - g), g = sigmoid(o.w1 + r.w2 + (o - r).w3)): per row
//   dg = dx . (o - r), da = dg g (1 - g), do = dx g + da (w1 + w3), dr = dx (1 - g) + da (w2 - w3);
// da [rows] is written out: dw1 = sum_r da_r o_r, dw2 = sum_r da_r r_r, dw3 = dw1 - dw2 (weighted column sums).
__global__ void gated_residual_bwd_kernel(const float* __restrict__ o, const float* __restrict__ res, const float* __restrict__ w,
                                          const float* __restrict__ dx, int rows, int dim, float* __restrict__ d_o,
                                          float* __restrict__ d_res, float* __restrict__ da_out) {
    const int wpb = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
        const float* op = o + static_cast<size_t>(r) * dim;
        const float* rp = res + static_cast<size_t>(r) * dim;
        const float* dp = dx + static_cast<size_t>(r) * dim;
        float acc = 0.f, dg = 0.f;
        for (int c = lane; c < dim; c += 32) {
            const float w1 = __ldg(w + c), w2 = __ldg(w + dim + c), w3 = __ldg(w + 2 * dim + c);
            acc += op[c] * (w1 + w3) + rp[c] * (w2 - w3);
            dg += dp[c] * (op[c] - rp[c]);
        }
        acc = warp_sum(acc);
        dg = warp_sum(dg);
        const float g = 1.f / (1.f + __expf(-acc));
        const float da = dg * g * (1.f - g);
        if (lane == 0) da_out[r] = da;
        for (int c = lane; c < dim; c += 32) {
            const float w1 = __ldg(w + c), w2 = __ldg(w + dim + c), w3 = __ldg(w + 2 * dim + c);
            d_o[static_cast<size_t>(r) * dim + c] = dp[c] * g + da * (w1 + w3);
            d_res[static_cast<size_t>(r) * dim + c] = dp[c] * (1.f - g) + da * (w2 - w3);
        }
    }
}

// Backward of graph_attn_core (same work split: one warp per (frame, head), lane l owns the rotary pair (2l, 2l+1)).
// With q~, k~ the rotated queries / keys, e_ij = A_ij we + be, z_ij = q~_i . (k~_j + e_ij), s = z / 8, p = softmax_j(s):
//   dv_j   = sum_i p_ij dout_i                      dwe += sum_i (sum_j p_ij A_ij) dout_i + sum_ij dz_ij A_ij q~_i
//   dp_ij  = dout_i . (v_j + e_ij)                   dbe += sum_i dout_i + sum_ij dz_ij q~_i
//   dz_ij  = p_ij (dp_ij - sum_j' p_ij' dp_ij') / 8
//   dq~_i  = sum_j dz_ij (k~_j + e_ij),   dk~_j = sum_i dz_ij q~_i,   dq / dk = inverse rotations.
// K~, V and the dK~ / dV accumulators of the frame live in shared memory; dqkv fp32 [rows, 3 * 512].
__global__ void __launch_bounds__(128)
graph_attn_core_bwd_kernel(const float* __restrict__ qkv, int ld, const int32_t* __restrict__ node_off,
                           const uint8_t* __restrict__ upper, int nmax, const float* __restrict__ we,
                           const float* __restrict__ be, const float* __restrict__ dout, int ldd, int n_frames,
                           float* __restrict__ dqkv, int ldg, float* __restrict__ dwe, float* __restrict__ dbe) {
    extern __shared__ float gsm[];
    constexpr int H = 8, DH = 64, INNER = H * DH;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.x, head = blockIdx.y * 4 + warp;
    if (f >= n_frames) return;
    const int r0 = node_off[f], n = node_off[f + 1] - r0;
    float* Ks = gsm + static_cast<size_t>(warp) * 4 * nmax * DH;
    float* Vs = Ks + nmax * DH;
    float* dKs = Vs + nmax * DH;
    float* dVs = dKs + nmax * DH;
    const uint8_t* U = upper + static_cast<size_t>(f) * nmax * nmax;
    const int d0 = lane * 2;
    const float inv_freq = __powf(10000.f, -static_cast<float>(d0) / DH);
    const float2 w2 = *reinterpret_cast<const float2*>(we + head * DH + d0);
    const float2 b2 = *reinterpret_cast<const float2*>(be + head * DH + d0);
    for (int j = 0; j < n; ++j) {
        const float* row = qkv + static_cast<size_t>(r0 + j) * ld + head * DH + d0;
        const float2 kk = *reinterpret_cast<const float2*>(row + INNER);
        const float2 vv = *reinterpret_cast<const float2*>(row + 2 * INNER);
        float sn, cs;
        __sincosf(j * inv_freq, &sn, &cs);
        Ks[j * DH + d0] = kk.x * cs - kk.y * sn; Ks[j * DH + d0 + 1] = kk.y * cs + kk.x * sn;
        Vs[j * DH + d0] = vv.x; Vs[j * DH + d0 + 1] = vv.y;
        dKs[j * DH + d0] = 0.f; dKs[j * DH + d0 + 1] = 0.f;
        dVs[j * DH + d0] = 0.f; dVs[j * DH + d0 + 1] = 0.f;
    }
    __syncwarp();
    float dwx = 0.f, dwy = 0.f, dbx = 0.f, dby = 0.f;
    for (int i = 0; i < n; ++i) {
        const float2 qq = *reinterpret_cast<const float2*>(qkv + static_cast<size_t>(r0 + i) * ld + head * DH + d0);
        const float2 go = *reinterpret_cast<const float2*>(dout + static_cast<size_t>(r0 + i) * ldd + head * DH + d0);
        float sn, cs;
        __sincosf(i * inv_freq, &sn, &cs);
        const float qx = qq.x * cs - qq.y * sn, qy = qq.y * cs + qq.x * sn;
        const float qw = warp_sum(qx * w2.x + qy * w2.y);
        const float qb = warp_sum(qx * b2.x + qy * b2.y);
        const float gw = warp_sum(go.x * w2.x + go.y * w2.y);      // dout_i . we
        const float gb = warp_sum(go.x * b2.x + go.y * b2.y);      // dout_i . be
        float mx = -INFINITY, my_score = -INFINITY, my_dp = 0.f, my_a = 0.f;     // lane j keeps the values of key j
#pragma unroll 1
        for (int j = 0; j < n; ++j) {
            const float dot = warp_sum(qx * Ks[j * DH + d0] + qy * Ks[j * DH + d0 + 1]);
            const float dv = warp_sum(go.x * Vs[j * DH + d0] + go.y * Vs[j * DH + d0 + 1]);
            const float a = (i < j ? U[i * nmax + j] : (i > j ? U[j * nmax + i] : 0)) ? 1.f : 0.f;
            const float sc = (dot + a * qw + qb) * 0.125f;
            if (lane == j) { my_score = sc; my_dp = dv + a * gw + gb; my_a = a; }
            mx = fmaxf(mx, sc);
        }
        const float e = lane < n ? __expf(my_score - mx) : 0.f;
        const float p_l = e / warp_sum(e);
        const float pdp = warp_sum(p_l * my_dp);
        const float dz_l = p_l * (my_dp - pdp) * 0.125f;           // d z_ij for key j = lane
        float dqx = 0.f, dqy = 0.f, pa = 0.f, za = 0.f, zs = 0.f;
#pragma unroll 1
        for (int j = 0; j < n; ++j) {
            const float p = __shfl_sync(0xffffffffu, p_l, j), dz = __shfl_sync(0xffffffffu, dz_l, j);
            const float a = __shfl_sync(0xffffffffu, my_a, j);
            dVs[j * DH + d0] = fmaf(p, go.x, dVs[j * DH + d0]);
            dVs[j * DH + d0 + 1] = fmaf(p, go.y, dVs[j * DH + d0 + 1]);
            dKs[j * DH + d0] = fmaf(dz, qx, dKs[j * DH + d0]);
            dKs[j * DH + d0 + 1] = fmaf(dz, qy, dKs[j * DH + d0 + 1]);
            dqx = fmaf(dz, Ks[j * DH + d0] + a * w2.x + b2.x, dqx);
            dqy = fmaf(dz, Ks[j * DH + d0 + 1] + a * w2.y + b2.y, dqy);
            pa = fmaf(p, a, pa);
            za = fmaf(dz, a, za);
            zs += dz;
        }
        dwx += pa * go.x + za * qx; dwy += pa * go.y + za * qy;
        dbx += go.x + zs * qx;      dby += go.y + zs * qy;
        // inverse rotation of dq~
        float* dq = dqkv + static_cast<size_t>(r0 + i) * ldg + head * DH + d0;
        dq[0] = dqx * cs + dqy * sn;
        dq[1] = dqy * cs - dqx * sn;
    }
    __syncwarp();
    for (int j = 0; j < n; ++j) {
        float sn, cs;
        __sincosf(j * inv_freq, &sn, &cs);
        const float kx = dKs[j * DH + d0], ky = dKs[j * DH + d0 + 1];
        float* row = dqkv + static_cast<size_t>(r0 + j) * ldg + head * DH + d0;
        row[INNER] = kx * cs + ky * sn;
        row[INNER + 1] = ky * cs - kx * sn;
        row[2 * INNER] = dVs[j * DH + d0];
        row[2 * INNER + 1] = dVs[j * DH + d0 + 1];
    }
    atomicAdd(dwe + head * DH + d0, dwx); atomicAdd(dwe + head * DH + d0 + 1, dwy);
    atomicAdd(dbe + head * DH + d0, dbx); atomicAdd(dbe + head * DH + d0 + 1, dby);
}

}  // namespace vsgg

// ------------------------------------------------------------------------------------------------
// SIMT fp32 helpers for the 10-wide STRUCTURE branch in differentiable mode: its linears have K = 10 or N = 10 (nothing a
// 128-wide tensor-core tile can use) and its LayerNorms have 10 columns.  The default (detached) mode runs the whole
// branch in ONE launch (graph_small.cu); these kernels exist so that the same network can be evaluated layer by layer
// with saved activations when its gradient is wanted.
// ------------------------------------------------------------------------------------------------
namespace vsgg {

// y[r, o] = act(sum_i x[r, i] * W(o, i) + b[o]),  W(o, i) = w[o * so + i * si]  (so / si choose W or W^T: forward / dgrad)
__global__ void __launch_bounds__(256)
simt_linear_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w, int so, int si,
                   const float* __restrict__ b, long long rows, int n_out, int n_in, int act, float* __restrict__ y, int ldy,
                   float* __restrict__ z_out) {
    const long long total = rows * n_out;
    for (long long idx = blockIdx.x * 256ll + threadIdx.x; idx < total; idx += gridDim.x * 256ll) {
        const long long r = idx / n_out;
        const int o = static_cast<int>(idx - r * n_out);
        const float* xr = x + r * ldx;
        float acc = b != nullptr ? __ldg(b + o) : 0.f;
        for (int i = 0; i < n_in; ++i) acc = fmaf(xr[i], __ldg(w + static_cast<size_t>(o) * so + static_cast<size_t>(i) * si), acc);
        if (z_out != nullptr) z_out[r * ldy + o] = acc;
        if (act == 2) acc = 0.5f * acc * (1.f + erff(acc * 0.70710678118654752f));
        y[r * ldy + o] = acc;
    }
}

// dW[o, i] += sum_r dy[r, o] * x[r, i]   (grid.y slabs of rows, one atomic per output and slab)
__global__ void __launch_bounds__(256)
simt_wgrad_kernel(const float* __restrict__ dy, int ldd, const float* __restrict__ x, int ldx, int rows, int n_out, int n_in,
                  float* __restrict__ dw) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= n_out * n_in) return;
    const int o = idx / n_in, i = idx - o * n_in;
    const int per = (rows + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
    float acc = 0.f;
    for (int r = r0; r < r1; ++r) acc = fmaf(dy[static_cast<size_t>(r) * ldd + o], x[static_cast<size_t>(r) * ldx + i], acc);
    atomicAdd(dw + idx, acc);
}

// dz = dy * gelu'(z)  (exact erf GELU)
__global__ void gelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ z, long long n, float* __restrict__ dz) {
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) {
        const float v = z[i];
        const float cdf = 0.5f * (1.f + erff(v * 0.70710678118654752f));
        dz[i] = dy[i] * (cdf + v * 0.3989422804014327f * __expf(-0.5f * v * v));
    }
}

// LayerNorm over <= 32 columns, thread = row
__global__ void ln_small_fwd_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
                                    int rows, int D, float* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float* xr = x + static_cast<size_t>(r) * D;
    float m = 0.f;
    for (int c = 0; c < D; ++c) m += xr[c];
    m /= D;
    float v = 0.f;
    for (int c = 0; c < D; ++c) { const float d = xr[c] - m; v += d * d; }
    const float rs = rsqrtf(v / D + 1e-5f);
    mean[r] = m;
    rstd[r] = rs;
    for (int c = 0; c < D; ++c) y[static_cast<size_t>(r) * D + c] = (xr[c] - m) * rs * __ldg(g + c) + __ldg(b + c);
}
// dx = base + rstd * (g dy - mean_c(g dy) - xhat mean_c(g dy xhat));  dgamma += sum_r dy xhat,  dbeta += sum_r dy
__global__ void __launch_bounds__(256)
ln_small_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ g,
                    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ base, int rows,
                    int D, float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta) {
    __shared__ float sg[32], sb[32];
    if (threadIdx.x < 32) { sg[threadIdx.x] = 0.f; sb[threadIdx.x] = 0.f; }
    __syncthreads();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) {
        const float m = mean[r], rs = rstd[r];
        float s1 = 0.f, s2 = 0.f;
        for (int c = 0; c < D; ++c) {
            const float xh = (x[static_cast<size_t>(r) * D + c] - m) * rs, gd = __ldg(g + c) * dy[static_cast<size_t>(r) * D + c];
            s1 += gd;
            s2 += gd * xh;
        }
        s1 /= D;
        s2 /= D;
        for (int c = 0; c < D; ++c) {
            const float d = dy[static_cast<size_t>(r) * D + c];
            const float xh = (x[static_cast<size_t>(r) * D + c] - m) * rs;
            const float v = rs * (__ldg(g + c) * d - s1 - xh * s2);
            dx[static_cast<size_t>(r) * D + c] = v + (base != nullptr ? base[static_cast<size_t>(r) * D + c] : 0.f);
            atomicAdd(&sg[c], d * xh);
            atomicAdd(&sb[c], d);
        }
    }
    __syncthreads();
    if (threadIdx.x < D) { atomicAdd(dgamma + threadIdx.x, sg[threadIdx.x]); atomicAdd(dbeta + threadIdx.x, sb[threadIdx.x]); }
}

}  // namespace vsgg

extern "C" int b200vsgg_simt_linear(const float* x, int32_t ldx, const float* w, int32_t so, int32_t si, const float* b,
                                    int64_t rows, int32_t n_out, int32_t n_in, int32_t act, float* y, int32_t ldy, float* z_out,
                                    void* stream) {
    if (!x || !w || !y || n_out <= 0 || n_in <= 0) return set_error(B200VSGG_ERR_BAD_ARG, "simt_linear: bad arg");
    if (rows == 0) return 0;
    const long long total = rows * n_out;
    const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 148LL * 32));
    vsgg::simt_linear_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, ldx, w, so, si, b, rows, n_out, n_in, act, y, ldy, z_out);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_simt_wgrad(const float* dy, int32_t ldd, const float* x, int32_t ldx, int32_t rows, int32_t n_out,
                                   int32_t n_in, float* dw, void* stream) {
    if (!dy || !x || !dw || n_out <= 0 || n_in <= 0) return set_error(B200VSGG_ERR_BAD_ARG, "simt_wgrad: bad arg");
    if (rows == 0) return 0;
    dim3 grid((n_out * n_in + 255) / 256, std::max(1, std::min(64, rows / 128)));
    vsgg::simt_wgrad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dy, ldd, x, ldx, rows, n_out, n_in, dw);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_gelu_bwd(const float* dy, const float* z, int64_t n, float* dz, void* stream) {
    if (!dy || !z || !dz) return set_error(B200VSGG_ERR_BAD_ARG, "gelu_bwd: bad arg");
    if (n == 0) return 0;
    vsgg::gelu_bwd_kernel<<<static_cast<int>(std::min<long long>((n + 255) / 256, 148LL * 16)), 256, 0, (cudaStream_t)stream>>>(dy, z, n, dz);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_ln_small_fwd(const float* x, const float* g, const float* b, int32_t rows, int32_t d, float* y,
                                     float* mean, float* rstd, void* stream) {
    if (!x || !g || !b || !y || !mean || !rstd || d < 1 || d > 32) return set_error(B200VSGG_ERR_BAD_ARG, "ln_small_fwd: bad arg (d <= 32)");
    if (rows == 0) return 0;
    vsgg::ln_small_fwd_kernel<<<(rows + 255) / 256, 256, 0, (cudaStream_t)stream>>>(x, g, b, rows, d, y, mean, rstd);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_ln_small_bwd(const float* dy, const float* x, const float* g, const float* mean, const float* rstd,
                                     const float* base, int32_t rows, int32_t d, float* dx, float* dgamma, float* dbeta,
                                     void* stream) {
    if (!dy || !x || !g || !mean || !rstd || !dx || !dgamma || !dbeta || d < 1 || d > 32)
        return set_error(B200VSGG_ERR_BAD_ARG, "ln_small_bwd: bad arg (d <= 32)");
    if (rows == 0) return 0;
    vsgg::ln_small_bwd_kernel<<<(rows + 255) / 256, 256, 0, (cudaStream_t)stream>>>(dy, x, g, mean, rstd, base, rows, d, dx, dgamma, dbeta);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_attn_pool_bwd(const float* x, int32_t d, const int32_t* node_off, int32_t n_frames, int32_t max_nodes,
                                      const float* w, const float* b, const float* dout, float* dx, float* dgate,
                                      void* stream) {
    if (!x || !node_off || !w || !b || !dout || !dx || !dgate || d <= 0 || max_nodes > vsgg::POOL_MAX_NODES)
        return set_error(B200VSGG_ERR_BAD_ARG, "attn_pool_bwd: bad arg (<= 64 nodes per frame)");
    if (n_frames <= 0) return 0;
    vsgg::attn_pool_bwd_kernel<<<n_frames, 256, 0, (cudaStream_t)stream>>>(x, d, node_off, w, b, dout, dx, dgate);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_consistency_kl_bwd(const float* g, int32_t d, const int32_t* pair_u, const int32_t* pair_v,
                                           const float* gout, int32_t n_pairs, float* dg, void* stream) {
    if (!g || !pair_u || !pair_v || !gout || !dg || d <= 0) return set_error(B200VSGG_ERR_BAD_ARG, "consistency_kl_bwd: bad arg");
    if (n_pairs == 0) return 0;
    int grid = (n_pairs + 7) / 8;
    if (grid > 148 * 16) grid = 148 * 16;
    vsgg::consistency_kl_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g, d, pair_u, pair_v, gout, n_pairs, dg);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_weighted_colsum(const void* x, int32_t x_is_bf16, int32_t ld, int32_t rows, int32_t cols,
                                        const float* wgt, float* out, void* stream) {
    if (!x || !wgt || !out || cols <= 0) return set_error(B200VSGG_ERR_BAD_ARG, "weighted_colsum: bad arg");
    if (rows == 0) return 0;
    int slabs = (rows + 255) / 256;
    if (slabs > 64) slabs = 64;
    dim3 grid((cols + 63) / 64, slabs);
    if (x_is_bf16)
        vsgg::weighted_colsum_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, ld, rows, cols, wgt, out);
    else
        vsgg::weighted_colsum_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, ld, rows, cols, wgt, out);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_gated_residual_bwd(const float* o, const float* res, const float* w, const float* dx, int32_t rows,
                                           int32_t dim, float* d_o, float* d_res, float* da, void* stream) {
    if (!o || !res || !w || !dx || !d_o || !d_res || !da || dim <= 0) return set_error(B200VSGG_ERR_BAD_ARG, "gated_residual_bwd: bad arg");
    if (rows == 0) return 0;
    int grid = (rows + 7) / 8;
    if (grid > 148 * 16) grid = 148 * 16;
    vsgg::gated_residual_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(o, res, w, dx, rows, dim, d_o, d_res, da);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_graph_attn_core_bwd(const float* qkv, int32_t ld, const int32_t* node_off, const uint8_t* upper,
                                            int32_t nmax, const float* we, const float* be, const float* dout, int32_t ldd,
                                            int32_t n_frames, float* dqkv, int32_t ldg, float* dwe, float* dbe, void* stream) {
    if (!qkv || !node_off || !upper || !we || !be || !dout || !dqkv || !dwe || !dbe || nmax <= 0 || nmax > 32 || (ld & 1) ||
        (ldd & 1) || (ldg & 1))
        return set_error(B200VSGG_ERR_BAD_ARG, "graph_attn_core_bwd: bad arg (<= 32 nodes per frame)");
    if (n_frames == 0) return 0;
    const size_t smem = 4ull * 4 * nmax * 64 * sizeof(float);
    static size_t cur = 48 * 1024;
    if (smem > cur) {
        cudaError_t e = cudaFuncSetAttribute(vsgg::graph_attn_core_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
        cur = smem;
    }
    vsgg::graph_attn_core_bwd_kernel<<<dim3(n_frames, 2), 128, smem, (cudaStream_t)stream>>>(
        qkv, ld, node_off, upper, nmax, we, be, dout, ldd, n_frames, dqkv, ldg, dwe, dbe);
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
