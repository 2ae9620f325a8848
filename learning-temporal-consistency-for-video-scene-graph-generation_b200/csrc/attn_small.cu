// Variable-length multi-head attention for SHORT sequences (spatial: one frame's pairs, ~8 tokens;
// temporal: a 2-frame window, ~16 tokens; long-clip config: 64).  Replaces the padded
// nn.MultiheadAttention calls of tools/utils/transformer.py:23,50 on exactly the kept (unpadded)
// rows — padded query rows are dropped by the reference (transformer.py:196,240-241) and padded
// keys are masked to -inf, so a varlen kernel is equivalent on every row that survives.
//
// Roofline: per token the kernel reads q,k,v (3*D bf16) and writes ctx (D bf16) = 8*D bytes for
// 4*L*D flops => ~L/2 flop/byte (4..32), far below the B200 ridge (~210 flop/byte): HBM-bound, so it
// runs on CUDA cores with coalesced loads, shared-memory staging of K/V and warp-shuffle softmax
// rather than on tensor cores.  head_dim = 242 is not 16-byte aligned per head, so loads are 32-bit
// (bf16x2) granular and fully coalesced across the warp.
//
// Grid = (segments, heads); one CTA stages K_h, V_h ([L][hd] fp32, odd pitch) and its warps walk
// the query rows: lane j owns key j for the scores, lanes own channels for the P*V product.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>

#include "../../include/b200vsgg.h"
#include "common.cuh"

namespace vsgg {

constexpr int ATT_THREADS = 128;
constexpr int ATT_WARPS = ATT_THREADS / 32;

__device__ __forceinline__ void stage_rows(float* dst, int pitch, const __nv_bfloat16* src, int ld, int row0, int L,
                                           int col0, int hd) {
    // rows [row0,row0+L) x cols [col0,col0+hd) of a bf16 matrix -> fp32 smem; col0 and hd are even.
    const int pairs = hd >> 1;
    for (int i = threadIdx.x; i < L * pairs; i += blockDim.x) {
        const int r = i / pairs, c = (i - r * pairs) * 2;
        const __nv_bfloat162 v =
            *reinterpret_cast<const __nv_bfloat162*>(src + static_cast<size_t>(row0 + r) * ld + col0 + c);
        const float2 f = __bfloat1622float2(v);
        dst[r * pitch + c] = f.x;
        dst[r * pitch + c + 1] = f.y;
    }
}

// keep/scale factor of attention-prob dropout for element (row, head, key j)
__device__ __forceinline__ float drop_factor(uint32_t thr, float inv_keep, unsigned long long seed, int row, int head,
                                             int j) {
    if (thr == 0u) return 1.f;
    const uint32_t h = hash_u32(seed, (static_cast<unsigned long long>(row) * 64ull + head) * 4096ull + j);
    return h >= thr ? inv_keep : 0.f;
}

__global__ void __launch_bounds__(ATT_THREADS)
attn_small_fwd_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const __nv_bfloat16* __restrict__ k, int ldk,
                      const __nv_bfloat16* __restrict__ v, int ldv, const int32_t* __restrict__ seg_off, int hd,
                      float scale, int max_len, __nv_bfloat16* __restrict__ ctx, int ldc, float drop_p,
                      unsigned long long seed) {
    extern __shared__ float sm[];
    const int seg = blockIdx.x, head = blockIdx.y;
    const int row0 = seg_off[seg];
    const int L = seg_off[seg + 1] - row0;
    if (L <= 0) return;
    const int pitch = hd + 1;
    float* Ks = sm;
    float* Vs = Ks + max_len * pitch;
    float* qs = Vs + max_len * pitch;       // [ATT_WARPS][pitch]
    float* ps = qs + ATT_WARPS * pitch;     // [ATT_WARPS][max_len]
    const int col0 = head * hd;
    stage_rows(Ks, pitch, k, ldk, row0, L, col0, hd);
    stage_rows(Vs, pitch, v, ldv, row0, L, col0, hd);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    const uint32_t thr = drop_p > 0.f ? static_cast<uint32_t>(drop_p * 4294967296.0) : 0u;
    float* myq = qs + warp * pitch;
    float* myp = ps + warp * max_len;
    for (int i = warp; i < L; i += ATT_WARPS) {
        const int row = row0 + i;
        for (int c = lane * 2; c < hd; c += 64) {
            const float2 f = __bfloat1622float2(
                *reinterpret_cast<const __nv_bfloat162*>(q + static_cast<size_t>(row) * ldq + col0 + c));
            myq[c] = f.x * scale;
            myq[c + 1] = f.y * scale;
        }
        __syncwarp();
        // scores: lane owns keys lane, lane+32, ...
        float mx = -INFINITY;
        for (int j = lane; j < L; j += 32) {
            const float* kr = Ks + j * pitch;
            float acc = 0.f;
            for (int d = 0; d < hd; ++d) acc = fmaf(myq[d], kr[d], acc);
            myp[j] = acc;
            mx = fmaxf(mx, acc);
        }
        mx = warp_max(mx);
        float den = 0.f;
        for (int j = lane; j < L; j += 32) {
            const float e = __expf(myp[j] - mx);
            myp[j] = e;
            den += e;
        }
        den = warp_sum(den);
        const float inv = 1.f / den;
        for (int j = lane; j < L; j += 32) myp[j] = myp[j] * inv * drop_factor(thr, inv_keep, seed, row, head, j);
        __syncwarp();
        // ctx[d] = sum_j p_j V[j][d]; lanes own channel pairs
        for (int c = lane * 2; c < hd; c += 64) {
            float a0 = 0.f, a1 = 0.f;
            for (int j = 0; j < L; ++j) {
                const float p = myp[j];
                a0 = fmaf(p, Vs[j * pitch + c], a0);
                a1 = fmaf(p, Vs[j * pitch + c + 1], a1);
            }
            *reinterpret_cast<__nv_bfloat162*>(ctx + static_cast<size_t>(row) * ldc + col0 + c) =
                __floats2bfloat162_rn(a0, a1);
        }
        __syncwarp();
    }
}

// Backward: recomputes P from q,k; given dctx produces dq, dk, dv (bf16).
//   P~ = dropout(P);  dV_j = sum_i P~_ij dO_i;  dP~_ij = dO_i . V_j;  dP = mask/(1-p) * dP~
//   dS_ij = P_ij (dP_ij - sum_j P_ij dP_ij);  dQ_i = scale * sum_j dS_ij K_j;  dK_j = scale * sum_i dS_ij Q_i
__global__ void __launch_bounds__(ATT_THREADS)
attn_small_bwd_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const __nv_bfloat16* __restrict__ k, int ldk,
                      const __nv_bfloat16* __restrict__ v, int ldv, const __nv_bfloat16* __restrict__ dctx, int ldc,
                      const int32_t* __restrict__ seg_off, int hd, float scale, int max_len,
                      __nv_bfloat16* __restrict__ dq, int lddq, __nv_bfloat16* __restrict__ dk, int lddk,
                      __nv_bfloat16* __restrict__ dv, int lddv, float drop_p, unsigned long long seed) {
    extern __shared__ float sm[];
    const int seg = blockIdx.x, head = blockIdx.y;
    const int row0 = seg_off[seg];
    const int L = seg_off[seg + 1] - row0;
    if (L <= 0) return;
    const int pitch = hd + 1;
    float* Qs = sm;
    float* Ks = Qs + max_len * pitch;
    float* Vs = Ks + max_len * pitch;
    float* Os = Vs + max_len * pitch;              // dO
    float* Pt = Os + max_len * pitch;              // P~  [L][max_len]
    float* dS = Pt + max_len * max_len;            // dS  [L][max_len]
    const int col0 = head * hd;
    stage_rows(Qs, pitch, q, ldq, row0, L, col0, hd);
    stage_rows(Ks, pitch, k, ldk, row0, L, col0, hd);
    stage_rows(Vs, pitch, v, ldv, row0, L, col0, hd);
    stage_rows(Os, pitch, dctx, ldc, row0, L, col0, hd);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    const uint32_t thr = drop_p > 0.f ? static_cast<uint32_t>(drop_p * 4294967296.0) : 0u;
    // pass 1: per query row, P, P~, dS
    for (int i = warp; i < L; i += ATT_WARPS) {
        const int row = row0 + i;
        const float* qr = Qs + i * pitch;
        const float* orow = Os + i * pitch;
        float* prow = Pt + i * max_len;
        float* srow = dS + i * max_len;
        float mx = -INFINITY;
        for (int j = lane; j < L; j += 32) {
            const float* kr = Ks + j * pitch;
            const float* vr = Vs + j * pitch;
            float s = 0.f, dp = 0.f;
            for (int d = 0; d < hd; ++d) {
                s = fmaf(qr[d], kr[d], s);
                dp = fmaf(orow[d], vr[d], dp);
            }
            s *= scale;
            prow[j] = s;
            srow[j] = dp;  // dP~ for now
            mx = fmaxf(mx, s);
        }
        mx = warp_max(mx);
        float den = 0.f;
        for (int j = lane; j < L; j += 32) {
            const float e = __expf(prow[j] - mx);
            prow[j] = e;
            den += e;
        }
        den = warp_sum(den);
        const float inv = 1.f / den;
        float delta = 0.f;
        for (int j = lane; j < L; j += 32) {
            const float p = prow[j] * inv;
            const float f = drop_factor(thr, inv_keep, seed, row, head, j);
            const float dp = srow[j] * f;  // dP
            delta += p * dp;
            prow[j] = p;
            srow[j] = dp;
        }
        delta = warp_sum(delta);
        for (int j = lane; j < L; j += 32) {
            const float p = prow[j];
            const float f = drop_factor(thr, inv_keep, seed, row, head, j);
            srow[j] = p * (srow[j] - delta);
            prow[j] = p * f;  // P~
        }
        __syncwarp();
        // dQ_i = scale * sum_j dS_ij K_j
        for (int c = lane * 2; c < hd; c += 64) {
            float a0 = 0.f, a1 = 0.f;
            for (int j = 0; j < L; ++j) {
                const float s = srow[j];
                a0 = fmaf(s, Ks[j * pitch + c], a0);
                a1 = fmaf(s, Ks[j * pitch + c + 1], a1);
            }
            *reinterpret_cast<__nv_bfloat162*>(dq + static_cast<size_t>(row) * lddq + col0 + c) =
                __floats2bfloat162_rn(a0 * scale, a1 * scale);
        }
    }
    __syncthreads();
    // pass 2: per key row j, dK_j and dV_j
    for (int j = warp; j < L; j += ATT_WARPS) {
        const int row = row0 + j;
        for (int c = lane * 2; c < hd; c += 64) {
            float k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
            for (int i = 0; i < L; ++i) {
                const float s = dS[i * max_len + j];
                const float p = Pt[i * max_len + j];
                k0 = fmaf(s, Qs[i * pitch + c], k0);
                k1 = fmaf(s, Qs[i * pitch + c + 1], k1);
                v0 = fmaf(p, Os[i * pitch + c], v0);
                v1 = fmaf(p, Os[i * pitch + c + 1], v1);
            }
            *reinterpret_cast<__nv_bfloat162*>(dk + static_cast<size_t>(row) * lddk + col0 + c) =
                __floats2bfloat162_rn(k0 * scale, k1 * scale);
            *reinterpret_cast<__nv_bfloat162*>(dv + static_cast<size_t>(row) * lddv + col0 + c) =
                __floats2bfloat162_rn(v0, v1);
        }
    }
}

}  // namespace vsgg

namespace vsgg {
int attn_mma_fwd_try(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const int32_t* seg_off,
                     int n_seg, int max_len, int n_heads, int hd, float scale, void* ctx, int ldc, float drop_p,
                     unsigned long long seed, cudaStream_t stream);
int attn_mma_bwd_try(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const void* dctx, int ldc,
                     const int32_t* seg_off, int n_seg, int max_len, int n_heads, int hd, float scale, void* dq, int lddq,
                     void* dk, int lddk, void* dv, int lddv, float drop_p, unsigned long long seed, cudaStream_t stream);
}  // namespace vsgg

using namespace vsgg;

extern "C" int b200vsgg_attn_small_fwd(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v,
                                       int32_t ldv, const int32_t* seg_off, int32_t n_seg, int32_t max_len,
                                       int32_t n_heads, int32_t head_dim, float scale, void* ctx, int32_t ldc,
                                       float drop_p, uint64_t seed, void* stream) {
    if (!q || !k || !v || !seg_off || !ctx || n_heads <= 0 || head_dim <= 0 || (head_dim & 1) || max_len <= 0)
        return set_error(B200VSGG_ERR_BAD_ARG, "attn_small_fwd: bad arg (head_dim must be even)");
    if (n_seg == 0) return 0;
    {   // warp-level mma path for segments <= 32 tokens (attn_mma.cu); the SIMT kernel below is the long-segment path
        const int took = attn_mma_fwd_try(q, ldq, k, ldk, v, ldv, seg_off, n_seg, max_len, n_heads, head_dim, scale, ctx,
                                          ldc, drop_p, seed, (cudaStream_t)stream);
        if (took != 0) return took < 0 ? took : 0;
    }
    const int pitch = head_dim + 1;
    const size_t smem = sizeof(float) * (2ull * max_len * pitch + ATT_WARPS * pitch + ATT_WARPS * max_len);
    if (smem > 227 * 1024) return set_error(B200VSGG_ERR_BAD_ARG, "attn_small_fwd: segment too long for shared memory");
    static size_t attr = 0;
    if (smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(attn_small_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
        attr = smem;
    }
    dim3 grid(n_seg, n_heads);
    attn_small_fwd_kernel<<<grid, ATT_THREADS, smem, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, ldk, (const __nv_bfloat16*)v, ldv, seg_off, head_dim,
        scale, max_len, (__nv_bfloat16*)ctx, ldc, drop_p, seed);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_attn_small_bwd(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v,
                                       int32_t ldv, const void* dctx, int32_t ldc, const int32_t* seg_off,
                                       int32_t n_seg, int32_t max_len, int32_t n_heads, int32_t head_dim, float scale,
                                       void* dq, int32_t lddq, void* dk, int32_t lddk, void* dv, int32_t lddv,
                                       float drop_p, uint64_t seed, void* stream) {
    if (!q || !k || !v || !dctx || !seg_off || !dq || !dk || !dv || n_heads <= 0 || head_dim <= 0 || (head_dim & 1) ||
        max_len <= 0)
        return set_error(B200VSGG_ERR_BAD_ARG, "attn_small_bwd: bad arg");
    if (n_seg == 0) return 0;
    {
        const int took = attn_mma_bwd_try(q, ldq, k, ldk, v, ldv, dctx, ldc, seg_off, n_seg, max_len, n_heads, head_dim,
                                          scale, dq, lddq, dk, lddk, dv, lddv, drop_p, seed, (cudaStream_t)stream);
        if (took != 0) return took < 0 ? took : 0;
    }
    const int pitch = head_dim + 1;
    const size_t smem = sizeof(float) * (4ull * max_len * pitch + 2ull * max_len * max_len);
    if (smem > 227 * 1024) return set_error(B200VSGG_ERR_BAD_ARG, "attn_small_bwd: segment too long for shared memory");
    static size_t attr = 0;
    if (smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(attn_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
        attr = smem;
    }
    dim3 grid(n_seg, n_heads);
    attn_small_bwd_kernel<<<grid, ATT_THREADS, smem, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, ldk, (const __nv_bfloat16*)v, ldv,
        (const __nv_bfloat16*)dctx, ldc, seg_off, head_dim, scale, max_len, (__nv_bfloat16*)dq, lddq,
        (__nv_bfloat16*)dk, lddk, (__nv_bfloat16*)dv, lddv, drop_p, seed);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}
