// Variable-length multi-head attention for sequences of ANY length (< 4096) and head_dim <= 320: the class-sequence
// encoder of the SGCls object branch (lib/tempura.py:88-92,201 of the reference: nn.TransformerEncoderLayer(2376,
// 8 heads) => head_dim 297, zero-padded to 304 by the caller), whose sequences are object tracks — a person track is as
// long as the video.  The register/shared-memory budgets of attn_mma (<= 32 tokens, hd <= 248) and attn_small
// (whole segment resident) do not cover that, and the flash kernels are specialised for head_dim <= 64.
//
// Flash-style without tensor cores: the branch holds ~0.3 % of the step's flops, so the point is generality, not peak.
// One CTA per (sequence, head); each warp owns one row, the partner rows are staged 32 at a time in shared
// memory, lane l forms the dot products with partner l, and the weighted sums run with lanes over channel pairs.
//   fwd      own = query i, partners = keys:    online softmax, O_i, lse_i
//   bwd dQ   own = query i, partners = keys:    delta_i = dO_i.O_i,  dQ_i = scale * sum_j dS_ij K_j
//   bwd dKV  own = key j,   partners = queries: dV_j = sum_i P~_ij dO_i,  dK_j = scale * sum_i dS_ij Q_i
// with P~ = dropout(P), dS_ij = P_ij (keep_ij/(1-p) * dO_i.V_j - delta_i).  No atomics, nothing quadratic stored.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>

#include "../../include/b200vsgg.h"
#include "common.cuh"

namespace vsgg {

constexpr int AR_WARPS = 8;
constexpr int AR_THREADS = AR_WARPS * 32;
constexpr int AR_MAXP = 5;          // channel pairs per lane: head_dim <= 320
constexpr int AR_CHUNK = 32;

struct AttnRowsArgs {
    const __nv_bfloat16 *q, *k, *v, *ctx, *dctx;
    int ldq, ldk, ldv, ldc, lddc;
    const int32_t* seg_off;
    int n_heads, hd;
    float scale;
    float* lse;      // [rows, n_heads]
    float* delta;    // [rows, n_heads]
    __nv_bfloat16 *out0, *out1;   // fwd: ctx | dq: dq | dkv: dk (out0), dv (out1)
    int ldo0, ldo1;
    float drop_p;
    unsigned long long seed;
};

__device__ __forceinline__ float ar_keep(uint32_t thr, float inv_keep, unsigned long long seed, int row, int head, int j) {
    if (thr == 0u) return 1.f;   // same element index as attn_small / attn_mma: (query row, head, key)
    const uint32_t h = hash_u32(seed, (static_cast<unsigned long long>(row) * 64ull + head) * 4096ull + j);
    return h >= thr ? inv_keep : 0.f;
}

// rows [r0, r0+n) x head columns of a bf16 matrix -> fp32 smem [AR_CHUNK][pitch]; missing rows are zero-filled
__device__ __forceinline__ void ar_stage(float* dst, int pitch, const __nv_bfloat16* src, int ld, int r0, int n, int col0,
                                         int hd) {
    const int pairs = hd >> 1;
    for (int i = threadIdx.x; i < AR_CHUNK * pairs; i += AR_THREADS) {
        const int r = i / pairs, c = (i - r * pairs) * 2;
        float2 f = make_float2(0.f, 0.f);
        if (r < n)
            f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(src + static_cast<size_t>(r0 + r) * ld + col0 + c));
        dst[r * pitch + c] = f.x;
        dst[r * pitch + c + 1] = f.y;
    }
}

__device__ __forceinline__ void ar_load_own(float* dst, const __nv_bfloat16* src, int hd, int lane, float mul) {
    for (int c = lane * 2; c < hd; c += 64) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(src + c));
        dst[c] = f.x * mul;
        dst[c + 1] = f.y * mul;
    }
}

__device__ __forceinline__ float ar_dot(const float* a, const float* x, int hd) {
    float acc = 0.f;
    for (int d = 0; d < hd; ++d) acc = fmaf(a[d], x[d], acc);
    return acc;
}

// acc[t] += sum_l coef_l * M[l][channel pair of (lane, t)]
__device__ __forceinline__ void ar_axpy(float (&acc)[AR_MAXP][2], float coef, const float* M, int pitch, int hd, int lane) {
#pragma unroll 4
    for (int l = 0; l < AR_CHUNK; ++l) {
        const float c = __shfl_sync(0xffffffffu, coef, l);
        const float* row = M + l * pitch;
#pragma unroll
        for (int t = 0; t < AR_MAXP; ++t) {
            const int ch = lane * 2 + 64 * t;
            if (ch < hd) {
                acc[t][0] = fmaf(c, row[ch], acc[t][0]);
                acc[t][1] = fmaf(c, row[ch + 1], acc[t][1]);
            }
        }
    }
}

__device__ __forceinline__ void ar_store(__nv_bfloat16* dst, const float (&acc)[AR_MAXP][2], float mul, int hd, int lane) {
#pragma unroll
    for (int t = 0; t < AR_MAXP; ++t) {
        const int ch = lane * 2 + 64 * t;
        if (ch < hd) *reinterpret_cast<__nv_bfloat162*>(dst + ch) = __floats2bfloat162_rn(acc[t][0] * mul, acc[t][1] * mul);
    }
}

// MODE 0: forward, 1: backward dQ (+ delta), 2: backward dK/dV
template <int MODE>
__global__ void __launch_bounds__(AR_THREADS) attn_rows_kernel(AttnRowsArgs a) {
    extern __shared__ float sm[];
    const int seg = blockIdx.x, head = blockIdx.y;
    const int row0 = a.seg_off[seg];
    const int L = a.seg_off[seg + 1] - row0;
    if (L <= 0) return;
    const int hd = a.hd, pitch = hd + 1, col0 = head * hd;
    float* X = sm;                              // [AR_CHUNK][pitch]
    float* Y = X + AR_CHUNK * pitch;            // [AR_CHUNK][pitch]
    float* own_a = Y + AR_CHUNK * pitch;        // [AR_WARPS][pitch]
    float* own_b = own_a + AR_WARPS * pitch;    // [AR_WARPS][pitch] (backward only)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* va = own_a + warp * pitch;
    float* vb = own_b + warp * pitch;
    const float inv_keep = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;
    const uint32_t thr = a.drop_p > 0.f ? static_cast<uint32_t>(a.drop_p * 4294967296.0) : 0u;

    for (int base = 0; base < L; base += AR_WARPS) {
        const int i = base + warp;
        const bool valid = i < L;
        const int row = row0 + (valid ? i : 0);
        float acc0[AR_MAXP][2], acc1[AR_MAXP][2];
#pragma unroll
        for (int t = 0; t < AR_MAXP; ++t) acc0[t][0] = acc0[t][1] = acc1[t][0] = acc1[t][1] = 0.f;
        float m_run = -INFINITY, l_run = 0.f, lse_i = 0.f, delta_i = 0.f;
        if (valid) {
            if (MODE == 0) {
                ar_load_own(va, a.q + static_cast<size_t>(row) * a.ldq + col0, hd, lane, a.scale);
            } else if (MODE == 1) {
                ar_load_own(va, a.q + static_cast<size_t>(row) * a.ldq + col0, hd, lane, a.scale);
                ar_load_own(vb, a.dctx + static_cast<size_t>(row) * a.lddc + col0, hd, lane, 1.f);
                float d = 0.f;
                for (int c = lane * 2; c < hd; c += 64) {
                    const float2 o = __bfloat1622float2(
                        *reinterpret_cast<const __nv_bfloat162*>(a.ctx + static_cast<size_t>(row) * a.ldc + col0 + c));
                    d = fmaf(vb[c], o.x, d);
                    d = fmaf(vb[c + 1], o.y, d);
                }
                delta_i = warp_sum(d);
                if (lane == 0) a.delta[static_cast<size_t>(row) * a.n_heads + head] = delta_i;
                lse_i = a.lse[static_cast<size_t>(row) * a.n_heads + head];
            } else {
                ar_load_own(va, a.k + static_cast<size_t>(row) * a.ldk + col0, hd, lane, a.scale);
                ar_load_own(vb, a.v + static_cast<size_t>(row) * a.ldv + col0, hd, lane, 1.f);
            }
        }
        __syncwarp();
        for (int c0 = 0; c0 < L; c0 += AR_CHUNK) {
            const int n = min(AR_CHUNK, L - c0);
            __syncthreads();
            if (MODE == 2) {
                ar_stage(X, pitch, a.q, a.ldq, row0 + c0, n, col0, hd);
                ar_stage(Y, pitch, a.dctx, a.lddc, row0 + c0, n, col0, hd);
            } else {
                ar_stage(X, pitch, a.k, a.ldk, row0 + c0, n, col0, hd);
                ar_stage(Y, pitch, a.v, a.ldv, row0 + c0, n, col0, hd);
            }
            __syncthreads();
            if (!valid) continue;
            const int pl = c0 + lane;                 // partner row (within the sequence) of this lane
            const bool pv = lane < n;
            const float u = ar_dot(va, X + lane * pitch, hd);
            if (MODE == 0) {
                const float s = pv ? u : -INFINITY;
                const float m_new = fmaxf(m_run, warp_max(s));
                const float p = pv ? __expf(s - m_new) : 0.f;
                const float corr = __expf(m_run - m_new);     // first chunk: exp(-inf) = 0
                l_run = l_run * corr + warp_sum(p);
                m_run = m_new;
#pragma unroll
                for (int t = 0; t < AR_MAXP; ++t) { acc0[t][0] *= corr; acc0[t][1] *= corr; }
                ar_axpy(acc0, p * ar_keep(thr, inv_keep, a.seed, row, head, pl), Y, pitch, hd, lane);
            } else if (MODE == 1) {
                const float w = ar_dot(vb, Y + lane * pitch, hd);
                const float p = pv ? __expf(u - lse_i) : 0.f;
                const float ds = p * (w * ar_keep(thr, inv_keep, a.seed, row, head, pl) - delta_i);
                ar_axpy(acc0, ds, X, pitch, hd, lane);
            } else {
                const float w = ar_dot(vb, Y + lane * pitch, hd);
                const int prow = row0 + (pv ? pl : 0);
                const float lse_p = a.lse[static_cast<size_t>(prow) * a.n_heads + head];
                const float del_p = a.delta[static_cast<size_t>(prow) * a.n_heads + head];
                const float p = pv ? __expf(u - lse_p) : 0.f;
                const float kp = ar_keep(thr, inv_keep, a.seed, prow, head, i);
                ar_axpy(acc1, p * kp, Y, pitch, hd, lane);                   // dV_j += P~_ij dO_i
                ar_axpy(acc0, p * (w * kp - del_p), X, pitch, hd, lane);     // dK_j += dS_ij Q_i
            }
        }
        if (!valid) continue;
        if (MODE == 0) {
            ar_store(a.out0 + static_cast<size_t>(row) * a.ldo0 + col0, acc0, 1.f / l_run, hd, lane);
            if (a.lse && lane == 0) a.lse[static_cast<size_t>(row) * a.n_heads + head] = m_run + __logf(l_run);
        } else if (MODE == 1) {
            ar_store(a.out0 + static_cast<size_t>(row) * a.ldo0 + col0, acc0, a.scale, hd, lane);
        } else {
            ar_store(a.out0 + static_cast<size_t>(row) * a.ldo0 + col0, acc0, a.scale, hd, lane);
            ar_store(a.out1 + static_cast<size_t>(row) * a.ldo1 + col0, acc1, 1.f, hd, lane);
        }
    }
}

static size_t ar_smem(int hd) { return static_cast<size_t>(2 * AR_CHUNK + 2 * AR_WARPS) * (hd + 1) * sizeof(float); }

template <int MODE>
static int ar_launch(const AttnRowsArgs& a, int n_seg, cudaStream_t stream) {
    const size_t smem = ar_smem(a.hd);
    static size_t cur = 0;
    if (smem > cur) {
        cudaError_t e = cudaFuncSetAttribute(attn_rows_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
        cur = smem;
    }
    attn_rows_kernel<MODE><<<dim3(n_seg, a.n_heads), AR_THREADS, smem, stream>>>(a);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

}  // namespace vsgg

using namespace vsgg;

extern "C" int b200vsgg_attn_rows_fwd(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv,
                                      const int32_t* seg_off, int32_t n_seg, int32_t n_heads, int32_t head_dim, float scale,
                                      void* ctx, int32_t ldc, float* lse, float drop_p, uint64_t seed, void* stream) {
    if (!q || !k || !v || !seg_off || !ctx || n_heads <= 0 || n_heads > 64 || head_dim <= 0 || (head_dim & 1) ||
        head_dim > 64 * AR_MAXP || ((ldq | ldk | ldv | ldc) & 1))
        return set_error(B200VSGG_ERR_BAD_ARG, "attn_rows_fwd: bad arg (head_dim even and <= 320, even leading dimensions)");
    if (n_seg == 0) return 0;
    AttnRowsArgs a = {};
    a.q = (const __nv_bfloat16*)q; a.k = (const __nv_bfloat16*)k; a.v = (const __nv_bfloat16*)v;
    a.ldq = ldq; a.ldk = ldk; a.ldv = ldv;
    a.seg_off = seg_off; a.n_heads = n_heads; a.hd = head_dim; a.scale = scale; a.lse = lse;
    a.out0 = (__nv_bfloat16*)ctx; a.ldo0 = ldc; a.drop_p = drop_p; a.seed = seed;
    return ar_launch<0>(a, n_seg, (cudaStream_t)stream);
}

extern "C" int b200vsgg_attn_rows_bwd(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv,
                                      const void* ctx, int32_t ldc, const void* dctx, int32_t lddc, const float* lse,
                                      float* delta, const int32_t* seg_off, int32_t n_seg, int32_t n_heads,
                                      int32_t head_dim, float scale, void* dq, int32_t lddq, void* dk, int32_t lddk,
                                      void* dv, int32_t lddv, float drop_p, uint64_t seed, void* stream) {
    if (!q || !k || !v || !ctx || !dctx || !lse || !delta || !seg_off || !dq || !dk || !dv || n_heads <= 0 || n_heads > 64 ||
        head_dim <= 0 || (head_dim & 1) || head_dim > 64 * AR_MAXP || ((ldq | ldk | ldv | ldc | lddc | lddq | lddk | lddv) & 1))
        return set_error(B200VSGG_ERR_BAD_ARG, "attn_rows_bwd: bad arg");
    if (n_seg == 0) return 0;
    AttnRowsArgs a = {};
    a.q = (const __nv_bfloat16*)q; a.k = (const __nv_bfloat16*)k; a.v = (const __nv_bfloat16*)v;
    a.ctx = (const __nv_bfloat16*)ctx; a.dctx = (const __nv_bfloat16*)dctx;
    a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldc = ldc; a.lddc = lddc;
    a.seg_off = seg_off; a.n_heads = n_heads; a.hd = head_dim; a.scale = scale;
    a.lse = const_cast<float*>(lse); a.delta = delta; a.drop_p = drop_p; a.seed = seed;
    a.out0 = (__nv_bfloat16*)dq; a.ldo0 = lddq;
    if (int rc = ar_launch<1>(a, n_seg, (cudaStream_t)stream)) return rc;
    a.out0 = (__nv_bfloat16*)dk; a.ldo0 = lddk; a.out1 = (__nv_bfloat16*)dv; a.ldo1 = lddv;
    return ar_launch<2>(a, n_seg, (cudaStream_t)stream);
}
