// Structure branch of the temporal-consistency regulariser (lib/teatgt.py:291-311,316,319 of
// the reference): per frame, the first 10 Laplacian eigenvector columns of the spatial graph go through
// graph_transformer_pytorch.GraphTransformer(dim 10, depth 4, heads 8 x 64, edge_dim 1, feed-forwards, gated
// residuals, rotary) and dgl GlobalAttentionPooling(Linear 10 -> 1)  ->  one structure embedding [10] per frame.
//
// The whole 4-layer network for one frame (<= 16 nodes x 10 features, 46 k parameters) runs in ONE CTA: the torch
// formulation of the same thing was ~150 launches over padded [frames, 11, ...] tensors and 5 ms of a 55 ms
// training step.  Warp h owns attention head h with q/k/v of every node in registers (lane = one rotary channel
// pair), scores by warp-shuffle reduction, softmax in registers; the 10-wide LayerNorms, projections back to 10,
// gated residuals and the feed-forward run on the CTA's threads over shared memory.  fp32 throughout.
// Roofline: 2048 frames x 4 layers x ~0.7 MFLOP = 6 GFLOP of SIMT work and ~0.7 GB of L2-resident weight reads:
// latency/issue-bound, target well under 0.5 ms; nothing here is GEMM-shaped enough for tcgen05 (K = 10).
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/b200vsgg.h"
#include "common.cuh"

namespace vsgg {

constexpr int GS_NMAX_LIMIT = 16; // nodes per frame (person + objects); larger frames use the generic path
constexpr int GS_DMAX = 16;       // feature width (10 in the reference)
constexpr int GS_DH = 64;         // dim_head of graph_transformer_pytorch (one rotary pair per lane)
constexpr int GS_THREADS = 256;

struct GsLayer {                  // offsets (in floats) into the packed parameter buffer of one layer
    int ln1_w, ln1_b, wq, bq, wkv, bkv, we, be, wo, bo, g1, ln2_w, ln2_b, w1, b1, w2, b2, g2, size;
};

__host__ __device__ inline GsLayer gs_layout(int D, int I) {
    GsLayer L;
    int o = 0;
    L.ln1_w = o; o += D;  L.ln1_b = o; o += D;
    L.wq = o; o += I * D; L.bq = o; o += I;
    L.wkv = o; o += 2 * I * D; L.bkv = o; o += 2 * I;
    L.we = o; o += I;     L.be = o; o += I;
    L.wo = o; o += D * I; L.bo = o; o += D;
    L.g1 = o; o += 3 * D;
    L.ln2_w = o; o += D;  L.ln2_b = o; o += D;
    L.w1 = o; o += 4 * D * D; L.b1 = o; o += 4 * D;
    L.w2 = o; o += 4 * D * D; L.b2 = o; o += D;
    L.g2 = o; o += 3 * D;
    L.size = o;
    return L;
}

__device__ __forceinline__ void gs_layernorm(const float* x, float* xn, int n, int D, const float* w, const float* b) {
    for (int i = threadIdx.x; i < n; i += GS_THREADS) {
        float mean = 0.f;
        for (int c = 0; c < D; ++c) mean += x[i * GS_DMAX + c];
        mean /= D;
        float var = 0.f;
        for (int c = 0; c < D; ++c) { const float d = x[i * GS_DMAX + c] - mean; var += d * d; }
        const float rstd = rsqrtf(var / D + 1e-5f);
        for (int c = 0; c < D; ++c) xn[i * GS_DMAX + c] = (x[i * GS_DMAX + c] - mean) * rstd * __ldg(w + c) + __ldg(b + c);
    }
}

// GatedResidual: g = sigmoid(W [o, res, o - res]);  x <- o * g + x * (1 - g)
__device__ __forceinline__ void gs_gate(const float* o, float* x, int n, int D, const float* gw) {
    for (int i = threadIdx.x; i < n; i += GS_THREADS) {
        float z = 0.f;
        for (int c = 0; c < D; ++c) {
            const float a = o[i * GS_DMAX + c], r = x[i * GS_DMAX + c];
            z += a * __ldg(gw + c) + r * __ldg(gw + D + c) + (a - r) * __ldg(gw + 2 * D + c);
        }
        const float g = 1.f / (1.f + __expf(-z));
        for (int c = 0; c < D; ++c) x[i * GS_DMAX + c] = o[i * GS_DMAX + c] * g + x[i * GS_DMAX + c] * (1.f - g);
    }
}

template <int GS_NMAX>
__global__ void __launch_bounds__(GS_THREADS, (GS_NMAX <= 12 ? 2 : 1)) graph_small_kernel(
    const float* __restrict__ nodes, const uint8_t* __restrict__ upper, const int32_t* __restrict__ counts, int nmax,
    int D, int heads, int depth, const float* __restrict__ params, const float* __restrict__ pool_w,
    const float* __restrict__ pool_b, float* __restrict__ out) {
    extern __shared__ float sm[];
    const int f = blockIdx.x;
    const int n = counts[f];
    const int I = heads * GS_DH;
    float* x = sm;                                // [GS_NMAX][GS_DMAX]
    float* xn = x + GS_NMAX * GS_DMAX;            // [GS_NMAX][GS_DMAX]
    float* o = xn + GS_NMAX * GS_DMAX;            // [GS_NMAX][GS_DMAX]
    float* adj = o + GS_NMAX * GS_DMAX;           // [GS_NMAX][GS_NMAX]
    float* hid = adj + GS_NMAX * GS_NMAX;         // [GS_NMAX][4*GS_DMAX]
    float* att = hid + GS_NMAX * 4 * GS_DMAX;     // [GS_NMAX][I]
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31, n_warps = GS_THREADS >> 5;
    for (int i = t; i < GS_NMAX * GS_DMAX; i += GS_THREADS) {
        const int r = i / GS_DMAX, c = i - r * GS_DMAX;
        x[i] = (r < n && c < D) ? nodes[(static_cast<size_t>(f) * nmax + r) * D + c] : 0.f;
    }
    for (int i = t; i < GS_NMAX * GS_NMAX; i += GS_THREADS) {
        const int r = i / GS_NMAX, c = i - r * GS_NMAX;
        float a = 0.f;
        if (r < n && c < n) {
            const uint8_t* u = upper + static_cast<size_t>(f) * nmax * nmax;
            a = static_cast<float>(u[r * nmax + c]) + static_cast<float>(u[c * nmax + r]);   // both directions are edges
        }
        adj[i] = a;
    }
    __syncthreads();
    const GsLayer L = gs_layout(D, I);
    const float scale = rsqrtf(static_cast<float>(GS_DH));
    const float inv_freq = __powf(10000.f, -static_cast<float>(2 * lane) / GS_DH);   // rotary pair `lane`

    for (int layer = 0; layer < depth; ++layer) {
        const float* P = params + static_cast<size_t>(layer) * L.size;
        // ---------------- attention block: x <- gate(to_out(attn(LN(x))), x)
        gs_layernorm(x, xn, n, D, P + L.ln1_w, P + L.ln1_b);
        __syncthreads();
        for (int h = warp; h < heads; h += n_warps) {
            const int e0 = h * GS_DH + 2 * lane;          // this lane's two channels of the head
            float q[GS_NMAX][2], k[GS_NMAX][2], v[GS_NMAX][2];
#pragma unroll
            for (int i = 0; i < GS_NMAX; ++i) {
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    q[i][s] = __ldg(P + L.bq + e0 + s);
                    k[i][s] = __ldg(P + L.bkv + e0 + s);
                    v[i][s] = __ldg(P + L.bkv + I + e0 + s);
                }
            }
            for (int c = 0; c < D; ++c) {             // weights of this lane's two channels: read once per layer
                const float wq0 = __ldg(P + L.wq + e0 * D + c), wq1 = __ldg(P + L.wq + (e0 + 1) * D + c);
                const float wk0 = __ldg(P + L.wkv + e0 * D + c), wk1 = __ldg(P + L.wkv + (e0 + 1) * D + c);
                const float wv0 = __ldg(P + L.wkv + (I + e0) * D + c), wv1 = __ldg(P + L.wkv + (I + e0 + 1) * D + c);
#pragma unroll
                for (int i = 0; i < GS_NMAX; ++i) {
                    if (i < n) {
                        const float xv = xn[i * GS_DMAX + c];
                        q[i][0] = fmaf(xv, wq0, q[i][0]); q[i][1] = fmaf(xv, wq1, q[i][1]);
                        k[i][0] = fmaf(xv, wk0, k[i][0]); k[i][1] = fmaf(xv, wk1, k[i][1]);
                        v[i][0] = fmaf(xv, wv0, v[i][0]); v[i][1] = fmaf(xv, wv1, v[i][1]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < GS_NMAX; ++i) {
                if (i < n) {
                    float sn, cs;
                    __sincosf(static_cast<float>(i) * inv_freq, &sn, &cs);   // rotary position = node index in the frame
                    const float q0 = q[i][0], q1 = q[i][1], k0 = k[i][0], k1 = k[i][1];
                    q[i][0] = q0 * cs - q1 * sn; q[i][1] = q1 * cs + q0 * sn;
                    k[i][0] = k0 * cs - k1 * sn; k[i][1] = k1 * cs + k0 * sn;
                }
            }
            const float we0 = __ldg(P + L.we + e0), we1 = __ldg(P + L.we + e0 + 1);
            const float be0 = __ldg(P + L.be + e0), be1 = __ldg(P + L.be + e0 + 1);
#pragma unroll
            for (int i = 0; i < GS_NMAX; ++i) {
                if (i < n) {
                    const float qw = warp_sum(q[i][0] * we0 + q[i][1] * we1);
                    const float qb = warp_sum(q[i][0] * be0 + q[i][1] * be1);
                    float s[GS_NMAX], mx = -INFINITY;
#pragma unroll
                    for (int j = 0; j < GS_NMAX; ++j) {
                        if (j < n) {
                            const float d = warp_sum(q[i][0] * k[j][0] + q[i][1] * k[j][1]);
                            s[j] = (d + qw * adj[i * GS_NMAX + j] + qb) * scale;
                            mx = fmaxf(mx, s[j]);
                        }
                    }
                    float den = 0.f;
#pragma unroll
                    for (int j = 0; j < GS_NMAX; ++j)
                        if (j < n) { s[j] = __expf(s[j] - mx); den += s[j]; }
                    const float inv = 1.f / den;
                    float o0 = 0.f, o1 = 0.f, pa = 0.f;
#pragma unroll
                    for (int j = 0; j < GS_NMAX; ++j) {
                        if (j < n) {
                            const float p = s[j] * inv;
                            o0 = fmaf(p, v[j][0], o0);
                            o1 = fmaf(p, v[j][1], o1);
                            pa = fmaf(p, adj[i * GS_NMAX + j], pa);
                        }
                    }
                    att[i * I + e0] = o0 + pa * we0 + be0;         // value offsets e_ij = A_ij * we + be
                    att[i * I + e0 + 1] = o1 + pa * we1 + be1;
                }
            }
        }
        __syncthreads();
        for (int idx = warp; idx < n * D; idx += n_warps) {       // to_out: [n, I] -> [n, D], one warp per output
            const int i = idx / D, c = idx - i * D;
            const float* w = P + L.wo + static_cast<size_t>(c) * I;
            float acc = 0.f;
            for (int e = lane; e < I; e += 32) acc = fmaf(att[i * I + e], __ldg(w + e), acc);
            acc = warp_sum(acc);
            if (lane == 0) o[i * GS_DMAX + c] = acc + __ldg(P + L.bo + c);
        }
        __syncthreads();
        gs_gate(o, x, n, D, P + L.g1);
        __syncthreads();
        // ---------------- feed-forward block: x <- gate(W2 gelu(W1 LN(x))), x)
        gs_layernorm(x, xn, n, D, P + L.ln2_w, P + L.ln2_b);
        __syncthreads();
        for (int idx = t; idx < n * 4 * D; idx += GS_THREADS) {
            const int i = idx / (4 * D), m = idx - i * 4 * D;
            float acc = __ldg(P + L.b1 + m);
            for (int c = 0; c < D; ++c) acc = fmaf(xn[i * GS_DMAX + c], __ldg(P + L.w1 + m * D + c), acc);
            hid[i * 4 * GS_DMAX + m] = 0.5f * acc * (1.f + erff(acc * 0.70710678118654752f));   // exact GELU
        }
        __syncthreads();
        for (int idx = t; idx < n * D; idx += GS_THREADS) {
            const int i = idx / D, c = idx - i * D;
            float acc = __ldg(P + L.b2 + c);
            for (int m = 0; m < 4 * D; ++m) acc = fmaf(hid[i * 4 * GS_DMAX + m], __ldg(P + L.w2 + c * 4 * D + m), acc);
            o[i * GS_DMAX + c] = acc;
        }
        __syncthreads();
        gs_gate(o, x, n, D, P + L.g2);
        __syncthreads();
    }
    // ---------------- GlobalAttentionPooling: softmax over the frame's nodes of gate_nn(x), weighted sum
    if (warp == 0) {
        float logit = -INFINITY;
        if (lane < n) {
            logit = __ldg(pool_b);
            for (int c = 0; c < D; ++c) logit = fmaf(x[lane * GS_DMAX + c], __ldg(pool_w + c), logit);
        }
        const float mx = warp_max(logit);
        const float e = lane < n ? __expf(logit - mx) : 0.f;
        const float w = e / warp_sum(e);
        for (int c = 0; c < D; ++c) {
            const float s = warp_sum(lane < n ? w * x[lane * GS_DMAX + c] : 0.f);
            if (lane == 0) out[static_cast<size_t>(f) * D + c] = s;
        }
    }
}

}  // namespace vsgg

using namespace vsgg;

extern "C" int b200vsgg_graph_small_params_per_layer(int32_t dim, int32_t heads) { return gs_layout(dim, heads * GS_DH).size; }

extern "C" int b200vsgg_graph_small_fwd(const float* nodes, const uint8_t* upper, const int32_t* counts, int32_t n_frames,
                                        int32_t nmax, int32_t dim, int32_t heads, int32_t depth, const float* params,
                                        const float* pool_w, const float* pool_b, float* out, void* stream) {
    if (!nodes || !upper || !counts || !params || !pool_w || !pool_b || !out || nmax < 1 || nmax > GS_NMAX_LIMIT || dim < 1 ||
        dim > GS_DMAX || heads < 1 || heads > 16 || depth < 1)
        return set_error(B200VSGG_ERR_BAD_ARG, "graph_small_fwd: bad arg (nmax <= 16, dim <= 16, heads <= 16)");
    if (n_frames == 0) return 0;
    auto launch = [&](auto kern, int NM) -> int {
        const size_t smem = sizeof(float) * (3 * NM * GS_DMAX + NM * NM + NM * 4 * GS_DMAX + static_cast<size_t>(NM) * heads * GS_DH);
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
        kern<<<n_frames, GS_THREADS, smem, (cudaStream_t)stream>>>(nodes, upper, counts, nmax, dim, heads, depth, params,
                                                                   pool_w, pool_b, out);
        return 0;
    };
    int rc = nmax <= 12 ? launch(graph_small_kernel<12>, 12) : launch(graph_small_kernel<16>, 16);
    if (rc) return rc;
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}
