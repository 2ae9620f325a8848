// GMM predicate-head epilogue (tools/utils/gmm_heads.py:37-76 and :25-35 of the reference).
// The 3*K*(2C+1) tiny Linear layers of the three heads are ONE tcgen05 GEMM [N,1936]x[336,1936]^T
// (b200vsgg_gemm_bf16, bias fused); this file turns its fp32 output z[N,ldz] into the mixture
// outputs.  Packed column layout of head h (C classes, K mixtures) starting at base_h:
//     [ mu_1..mu_K : K*C | var_1..var_K : K*C | pi_1..pi_K : K ]
// attention (C=3, softmax) | spatial (C=6, sigmoid) | contacting (C=17, sigmoid).
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>

#include "../../include/b200vsgg.h"
#include "common.cuh"

namespace vsgg {

constexpr int GMM_MAX_K = 8;
constexpr int GMM_MAX_C = 40;

struct GmmHeadDesc {
    int base;       // first packed column of this head
    int C;          // classes
    int softmax;    // 1: softmax over classes, 0: sigmoid
    const float* eps;   // [K,N,C] noise or nullptr
    float* out;     // [N,C]    distribution   (or aleatoric uncertainty in unc mode)
    float* out2;    // [N,C]    epistemic uncertainty (unc mode only)
    const float* dout;  // [N,C] upstream gradient (backward)
};
struct GmmParams {
    GmmHeadDesc h[3];
    int n_heads;
    int K;
    int N;
    int ldz;
    int mode;  // 0: test (mu only), 1: train (mu + sqrt(var)*eps), 2: uncertainty
    unsigned long long seed;  // used when eps == nullptr in train mode
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// standard normal from a counter (Box-Muller on two hashed uniforms)
__device__ __forceinline__ float hashed_normal(unsigned long long seed, unsigned long long idx) {
    const uint32_t a = hash_u32(seed, 2ull * idx), b = hash_u32(seed ^ 0xD1B54A32D192ED03ull, 2ull * idx + 1ull);
    const float u1 = (static_cast<float>(a) + 1.0f) * (1.0f / 4294967296.0f);
    const float u2 = static_cast<float>(b) * (1.0f / 4294967296.0f);
    return sqrtf(-2.f * __logf(u1)) * __cosf(6.28318530717958647692f * u2);
}

__device__ __forceinline__ float fetch_eps(const GmmParams& p, const GmmHeadDesc& hd, int head, int k, int n, int c) {
    if (hd.eps) return hd.eps[(static_cast<size_t>(k) * p.N + n) * hd.C + c];
    return hashed_normal(p.seed, ((static_cast<unsigned long long>(head) * GMM_MAX_K + k) * p.N + n) * GMM_MAX_C + c);
}

// activation of one mixture component into a[0..C)
__device__ __forceinline__ void act_component(const float* logit, int C, int softmax, float* a) {
    if (softmax) {
        float mx = -INFINITY;
        for (int c = 0; c < C; ++c) mx = fmaxf(mx, logit[c]);
        float den = 0.f;
        for (int c = 0; c < C; ++c) { a[c] = __expf(logit[c] - mx); den += a[c]; }
        const float inv = 1.f / den;
        for (int c = 0; c < C; ++c) a[c] *= inv;
    } else {
        for (int c = 0; c < C; ++c) a[c] = sigmoidf_(logit[c]);
    }
}

__global__ void gmm_head_fwd_kernel(const float* __restrict__ z, const GmmParams p) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = t / p.n_heads, head = t - n * p.n_heads;
    if (n >= p.N) return;
    const GmmHeadDesc hd = p.h[head];
    const int C = hd.C, K = p.K;
    const float* zr = z + static_cast<size_t>(n) * p.ldz + hd.base;
    float pi[GMM_MAX_K];
    float mx = -INFINITY;
    for (int k = 0; k < K; ++k) { pi[k] = zr[2 * K * C + k]; mx = fmaxf(mx, pi[k]); }
    float den = 0.f;
    for (int k = 0; k < K; ++k) { pi[k] = __expf(pi[k] - mx); den += pi[k]; }
    for (int k = 0; k < K; ++k) pi[k] /= den;

    float acc[GMM_MAX_C], acc2[GMM_MAX_C], logit[GMM_MAX_C], a[GMM_MAX_C];
    for (int c = 0; c < C; ++c) { acc[c] = 0.f; acc2[c] = 0.f; }
    if (p.mode == 2) {
        // aleatoric = sum_k pi_k var_k ; epistemic = sum_k pi_k (act(mu_k) - mean)^2
        float mean[GMM_MAX_C];
        for (int c = 0; c < C; ++c) mean[c] = 0.f;
        for (int k = 0; k < K; ++k) {
            for (int c = 0; c < C; ++c) logit[c] = zr[k * C + c];
            act_component(logit, C, hd.softmax, a);
            for (int c = 0; c < C; ++c) {
                mean[c] += pi[k] * a[c];
                acc[c] += pi[k] * sigmoidf_(zr[K * C + k * C + c]);
            }
        }
        for (int k = 0; k < K; ++k) {
            for (int c = 0; c < C; ++c) logit[c] = zr[k * C + c];
            act_component(logit, C, hd.softmax, a);
            for (int c = 0; c < C; ++c) { const float d = a[c] - mean[c]; acc2[c] += pi[k] * d * d; }
        }
        for (int c = 0; c < C; ++c) {
            hd.out[static_cast<size_t>(n) * C + c] = acc[c];
            hd.out2[static_cast<size_t>(n) * C + c] = acc2[c];
        }
        return;
    }
    // softmax == 2 (object head, test phase, gmm_heads.py:63-64 `mu[..., 1:]`): the background class is dropped BEFORE the
    // activation and the output has C - 1 columns
    const int c0 = (hd.softmax == 2 && p.mode == 0) ? 1 : 0, Cc = C - c0;
    for (int k = 0; k < K; ++k) {
        for (int c = 0; c < Cc; ++c) {
            float l = zr[k * C + c0 + c];
            if (p.mode == 1) l += sqrtf(sigmoidf_(zr[K * C + k * C + c])) * fetch_eps(p, hd, head, k, n, c);
            logit[c] = l;
        }
        act_component(logit, Cc, hd.softmax, a);
        for (int c = 0; c < Cc; ++c) acc[c] += pi[k] * a[c];
    }
    for (int c = 0; c < Cc; ++c) hd.out[static_cast<size_t>(n) * Cc + c] = acc[c];
}

// Backward of the train/test mixture output w.r.t. the packed logits z.  Writes dz (bf16, the A
// operand of the head dgrad/wgrad GEMMs) for this head's columns; padding columns are zeroed by the
// thread of head 0.
__global__ void gmm_head_bwd_kernel(const float* __restrict__ z, const GmmParams p, __nv_bfloat16* __restrict__ dz,
                                    int lddz, int total_cols) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = t / p.n_heads, head = t - n * p.n_heads;
    if (n >= p.N) return;
    const GmmHeadDesc hd = p.h[head];
    const int C = hd.C, K = p.K;
    const float* zr = z + static_cast<size_t>(n) * p.ldz + hd.base;
    __nv_bfloat16* dzr = dz + static_cast<size_t>(n) * lddz + hd.base;
    const float* go = hd.dout + static_cast<size_t>(n) * C;
    float pi[GMM_MAX_K], dpi[GMM_MAX_K];
    float mx = -INFINITY;
    for (int k = 0; k < K; ++k) { pi[k] = zr[2 * K * C + k]; mx = fmaxf(mx, pi[k]); }
    float den = 0.f;
    for (int k = 0; k < K; ++k) { pi[k] = __expf(pi[k] - mx); den += pi[k]; }
    for (int k = 0; k < K; ++k) pi[k] /= den;
    float logit[GMM_MAX_C], a[GMM_MAX_C];
    for (int k = 0; k < K; ++k) {
        float sd[GMM_MAX_C], ep[GMM_MAX_C], var[GMM_MAX_C];
        for (int c = 0; c < C; ++c) {
            float l = zr[k * C + c];
            if (p.mode == 1) {
                var[c] = sigmoidf_(zr[K * C + k * C + c]);
                sd[c] = sqrtf(var[c]);
                ep[c] = fetch_eps(p, hd, head, k, n, c);
                l += sd[c] * ep[c];
            }
            logit[c] = l;
        }
        act_component(logit, C, hd.softmax, a);
        float dp = 0.f, dot = 0.f;
        for (int c = 0; c < C; ++c) { dp += go[c] * a[c]; }
        dpi[k] = dp;
        // d a_k[c] = go[c] * pi_k
        if (hd.softmax) {
            for (int c = 0; c < C; ++c) dot += a[c] * go[c] * pi[k];
        }
        for (int c = 0; c < C; ++c) {
            const float da = go[c] * pi[k];
            const float dl = hd.softmax ? a[c] * (da - dot) : a[c] * (1.f - a[c]) * da;
            dzr[k * C + c] = __float2bfloat16(dl);
            float dvz = 0.f;
            if (p.mode == 1) dvz = dl * ep[c] * 0.5f / fmaxf(sd[c], 1e-20f) * var[c] * (1.f - var[c]);
            dzr[K * C + k * C + c] = __float2bfloat16(dvz);
        }
    }
    float s = 0.f;
    for (int k = 0; k < K; ++k) s += pi[k] * dpi[k];
    for (int k = 0; k < K; ++k) dzr[2 * K * C + k] = __float2bfloat16(pi[k] * (dpi[k] - s));
    if (head == 0) {
        int used = 0;
        for (int h = 0; h < p.n_heads; ++h) used += p.K * (2 * p.h[h].C + 1);
        for (int c = used; c < total_cols; ++c) dz[static_cast<size_t>(n) * lddz + c] = __float2bfloat16(0.f);
    }
}

}  // namespace vsgg

using namespace vsgg;

static int fill_params(GmmParams& p, const b200vsgg_gmm_head* heads, int n_heads, int K, int N, int ldz, int mode,
                       uint64_t seed) {
    if (n_heads < 1 || n_heads > 3 || K < 1 || K > GMM_MAX_K) return set_error(B200VSGG_ERR_BAD_ARG, "gmm_head: bad K / n_heads");
    p.n_heads = n_heads; p.K = K; p.N = N; p.ldz = ldz; p.mode = mode; p.seed = seed;
    for (int i = 0; i < n_heads; ++i) {
        if (heads[i].num_classes < 1 || heads[i].num_classes > GMM_MAX_C) return set_error(B200VSGG_ERR_BAD_ARG, "gmm_head: bad class count");
        p.h[i].base = heads[i].col_base;
        p.h[i].C = heads[i].num_classes;
        p.h[i].softmax = heads[i].softmax;
        p.h[i].eps = heads[i].eps;
        p.h[i].out = heads[i].out;
        p.h[i].out2 = heads[i].out2;
        p.h[i].dout = heads[i].dout;
    }
    return 0;
}

extern "C" int b200vsgg_gmm_head_fwd(const float* z, int32_t ldz, int32_t n_rows, int32_t K,
                                     const b200vsgg_gmm_head* heads, int32_t n_heads, int32_t mode, uint64_t seed,
                                     void* stream) {
    if (!z || !heads) return set_error(B200VSGG_ERR_BAD_ARG, "gmm_head_fwd: null pointer");
    if (n_rows == 0) return 0;
    GmmParams p;
    int rc = fill_params(p, heads, n_heads, K, n_rows, ldz, mode, seed);
    if (rc) return rc;
    for (int i = 0; i < n_heads; ++i)
        if (!p.h[i].out || (mode == 2 && !p.h[i].out2)) return set_error(B200VSGG_ERR_BAD_ARG, "gmm_head_fwd: missing output");
    const int total = n_rows * n_heads;
    gmm_head_fwd_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(z, p);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_gmm_head_bwd(const float* z, int32_t ldz, int32_t n_rows, int32_t K,
                                     const b200vsgg_gmm_head* heads, int32_t n_heads, int32_t mode, uint64_t seed,
                                     void* dz_bf16, int32_t lddz, int32_t total_cols, void* stream) {
    if (!z || !heads || !dz_bf16 || mode == 2) return set_error(B200VSGG_ERR_BAD_ARG, "gmm_head_bwd: bad arg");
    if (n_rows == 0) return 0;
    GmmParams p;
    int rc = fill_params(p, heads, n_heads, K, n_rows, ldz, mode, seed);
    if (rc) return rc;
    for (int i = 0; i < n_heads; ++i)
        if (!p.h[i].dout) return set_error(B200VSGG_ERR_BAD_ARG, "gmm_head_bwd: missing dout");
    const int total = n_rows * n_heads;
    gmm_head_bwd_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(z, p, (__nv_bfloat16*)dz_bf16, lddz,
                                                                              total_cols);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}
