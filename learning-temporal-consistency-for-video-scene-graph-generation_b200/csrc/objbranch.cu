// Object tokens of the SGCls object branch (lib/tempura.py:249-252 of the reference; TEAT-GT's copy
// tools/utils/object_classifier.py:230-233) fused with the gather into class-sequence order and the
// sinusoidal position add of PositionalEncoding.forward (lib/tempura.py:39-48):
//
//   row r of the output <- box s = src[r] (src == NULL: identity, the non-tracking path)
//   x[r] = drop_pe( [ features[s] | distribution[s] @ obj_embed.weight | drop_pos(relu(Wp bn(center_size(box_s)) + bp)) ]
//                   + pe[pos[r]] )                                   (pos == NULL: no position term, no drop_pe)
//
// HBM-bound row kernel: one CTA per row, float4 accesses, the 36 x 200 embedding matrix and the 128 x 4
// position MLP stay in L1/L2.  BatchNorm1d(4) enters as per-video (mean, rstd) so that a batch of videos
// keeps the reference's per-video statistics (its batch is one video).  Everything is fp32: this is 0.3 % of
// the branch's flops and feeds discrete decisions nowhere, but fp32 keeps `dist @ E` at reference accuracy.
//
// Backward: only parameters need gradients (features / distribution / boxes are frozen detector outputs):
//   d obj_embed.weight[k, c] = sum_r dist[s, k] * dx[r, F + c]
//   d Wp, d bp, d bn.gamma, d bn.beta of the position MLP.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>

#include "../../include/b200vsgg.h"
#include "common.cuh"

namespace vsgg {

constexpr int OBJ_MAX_CLS = 64;

struct ObjTokArgs {
    const float* features; int feat_dim;
    const float* dist; int n_cls;
    const float* embed; int e;
    const float* boxes;                       // [O,5] = (frame, x1, y1, x2, y2)
    const float *bn_mean, *bn_rstd;           // [V,4]
    const float *bn_gamma, *bn_beta;          // [4]
    const int32_t* video_of_box;
    const float *wp, *bp; int h;              // [h,4], [h]
    const float* pe;                          // [max_len, D]
    const int32_t *src, *pos;
    int rows;
    float p_pos; unsigned long long seed_pos;
    float p_pe; unsigned long long seed_pe;
};

// center_size (tools/utils/fpn/box_utils.py, neural-motifs): (cx, cy, w, h) with w = x2 - x1 + 1
__device__ __forceinline__ void bn_center_size(const ObjTokArgs& a, int s, float (&xhat)[4], float (&ybn)[4]) {
    const float* b = a.boxes + static_cast<size_t>(s) * 5;
    const float x1 = b[1], y1 = b[2], x2 = b[3], y2 = b[4];
    const float w = x2 - x1 + 1.0f, h = y2 - y1 + 1.0f;
    const float cs[4] = {x1 + 0.5f * w, y1 + 0.5f * h, w, h};
    const int v = __ldg(a.video_of_box + s);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        xhat[c] = (cs[c] - __ldg(a.bn_mean + v * 4 + c)) * __ldg(a.bn_rstd + v * 4 + c);
        ybn[c] = fmaf(xhat[c], __ldg(a.bn_gamma + c), __ldg(a.bn_beta + c));
    }
}

__device__ __forceinline__ float keep_scale(uint32_t thr, float inv_keep, unsigned long long seed, unsigned long long idx) {
    if (thr == 0u) return 1.f;
    return hash_u32(seed, idx) >= thr ? inv_keep : 0.f;
}

__global__ void __launch_bounds__(256) obj_tokens_fwd_kernel(ObjTokArgs a, float* __restrict__ x_f32,
                                                             __nv_bfloat16* __restrict__ x_bf16) {
    __shared__ float sd[OBJ_MAX_CLS];
    const int D = a.feat_dim + a.e + a.h;
    const int r = blockIdx.x;
    const int s = a.src ? __ldg(a.src + r) : r;
    if (threadIdx.x < a.n_cls) sd[threadIdx.x] = a.dist[static_cast<size_t>(s) * a.n_cls + threadIdx.x];
    __syncthreads();
    float xhat[4], ybn[4];
    bn_center_size(a, s, xhat, ybn);
    const uint32_t thr_pos = a.p_pos > 0.f ? static_cast<uint32_t>(a.p_pos * 4294967296.0) : 0u;
    const uint32_t thr_pe = (a.pos && a.p_pe > 0.f) ? static_cast<uint32_t>(a.p_pe * 4294967296.0) : 0u;
    const float ik_pos = a.p_pos > 0.f ? 1.f / (1.f - a.p_pos) : 1.f, ik_pe = a.p_pe > 0.f ? 1.f / (1.f - a.p_pe) : 1.f;
    const float* per = a.pos ? a.pe + static_cast<size_t>(__ldg(a.pos + r)) * D : nullptr;
    const int f4 = a.feat_dim >> 2, e4 = a.e >> 2, d4 = D >> 2;
    for (int c4 = threadIdx.x; c4 < d4; c4 += blockDim.x) {
        float o[4];
        if (c4 < f4) {
            const float4 v = *reinterpret_cast<const float4*>(a.features + static_cast<size_t>(s) * a.feat_dim + c4 * 4);
            o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
        } else if (c4 < f4 + e4) {
            const int c = (c4 - f4) * 4;
            o[0] = o[1] = o[2] = o[3] = 0.f;
            for (int k = 0; k < a.n_cls; ++k) {
                const float4 w = __ldg(reinterpret_cast<const float4*>(a.embed + static_cast<size_t>(k) * a.e + c));
                const float d = sd[k];
                o[0] = fmaf(d, w.x, o[0]); o[1] = fmaf(d, w.y, o[1]); o[2] = fmaf(d, w.z, o[2]); o[3] = fmaf(d, w.w, o[3]);
            }
        } else {
            const int j0 = (c4 - f4 - e4) * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 w = __ldg(reinterpret_cast<const float4*>(a.wp) + j0 + j);
                float hcur = __ldg(a.bp + j0 + j);
                hcur = fmaf(w.x, ybn[0], hcur); hcur = fmaf(w.y, ybn[1], hcur);
                hcur = fmaf(w.z, ybn[2], hcur); hcur = fmaf(w.w, ybn[3], hcur);
                o[j] = fmaxf(hcur, 0.f) * keep_scale(thr_pos, ik_pos, a.seed_pos, static_cast<unsigned long long>(s) * a.h + j0 + j);
            }
        }
        if (per) {
            const float4 p = __ldg(reinterpret_cast<const float4*>(per) + c4);
            o[0] += p.x; o[1] += p.y; o[2] += p.z; o[3] += p.w;
            if (thr_pe) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    o[j] *= keep_scale(thr_pe, ik_pe, a.seed_pe, static_cast<unsigned long long>(r) * D + c4 * 4 + j);
            }
        }
        if (x_f32) *reinterpret_cast<float4*>(x_f32 + static_cast<size_t>(r) * D + c4 * 4) = make_float4(o[0], o[1], o[2], o[3]);
        if (x_bf16) store_bf16x4(x_bf16 + static_cast<size_t>(r) * D + c4 * 4, o);
    }
}

// Threads [0, e) own one embedding column each (n_cls register accumulators); threads [e_pad, e_pad + h) own one
// hidden unit of the position MLP.  A CTA walks a strided set of rows and flushes with one atomicAdd per value.
__global__ void __launch_bounds__(384) obj_tokens_bwd_kernel(ObjTokArgs a, const float* __restrict__ dx, int e_pad,
                                                             float* __restrict__ dembed, float* __restrict__ dwp,
                                                             float* __restrict__ dbp, float* __restrict__ dgamma,
                                                             float* __restrict__ dbeta) {
    __shared__ float sd[OBJ_MAX_CLS];
    __shared__ float sdy[4];
    const int D = a.feat_dim + a.e + a.h;
    const int t = threadIdx.x;
    const bool emb_role = t < a.e;
    const int j = t - e_pad;
    const bool pos_role = j >= 0 && j < a.h;
    const uint32_t thr_pos = a.p_pos > 0.f ? static_cast<uint32_t>(a.p_pos * 4294967296.0) : 0u;
    const uint32_t thr_pe = (a.pos && a.p_pe > 0.f) ? static_cast<uint32_t>(a.p_pe * 4294967296.0) : 0u;
    const float ik_pos = a.p_pos > 0.f ? 1.f / (1.f - a.p_pos) : 1.f, ik_pe = a.p_pe > 0.f ? 1.f / (1.f - a.p_pe) : 1.f;
    float acc[OBJ_MAX_CLS];
#pragma unroll
    for (int k = 0; k < OBJ_MAX_CLS; ++k) acc[k] = 0.f;
    float aw[4] = {0.f, 0.f, 0.f, 0.f}, ab = 0.f, ag[4] = {0.f, 0.f, 0.f, 0.f}, abt[4] = {0.f, 0.f, 0.f, 0.f};
    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
    float bj = 0.f;
    if (pos_role) {
        w = __ldg(reinterpret_cast<const float4*>(a.wp) + j);
        bj = __ldg(a.bp + j);
    }
    if (t < 4) sdy[t] = 0.f;
    for (int r = blockIdx.x; r < a.rows; r += gridDim.x) {
        const int s = a.src ? __ldg(a.src + r) : r;
        __syncthreads();   // previous row's readers of sd / sdy are done
        if (t < a.n_cls) sd[t] = a.dist[static_cast<size_t>(s) * a.n_cls + t];
        if (t < 4) sdy[t] = 0.f;
        __syncthreads();
        const float* dxr = dx + static_cast<size_t>(r) * D;
        float xhat[4], ybn[4];
        bn_center_size(a, s, xhat, ybn);
        if (emb_role) {
            const int col = a.feat_dim + t;
            const float g = dxr[col] * keep_scale(thr_pe, ik_pe, a.seed_pe, static_cast<unsigned long long>(r) * D + col);
#pragma unroll
            for (int k = 0; k < OBJ_MAX_CLS; ++k)
                if (k < a.n_cls) acc[k] = fmaf(sd[k], g, acc[k]);
        }
        float dy[4] = {0.f, 0.f, 0.f, 0.f};
        if (pos_role) {
            const int col = a.feat_dim + a.e + j;
            float hcur = bj;
            hcur = fmaf(w.x, ybn[0], hcur); hcur = fmaf(w.y, ybn[1], hcur);
            hcur = fmaf(w.z, ybn[2], hcur); hcur = fmaf(w.w, ybn[3], hcur);
            float g = dxr[col] * keep_scale(thr_pe, ik_pe, a.seed_pe, static_cast<unsigned long long>(r) * D + col);
            g *= keep_scale(thr_pos, ik_pos, a.seed_pos, static_cast<unsigned long long>(s) * a.h + j);
            g = hcur > 0.f ? g : 0.f;
            ab += g;
#pragma unroll
            for (int c = 0; c < 4; ++c) aw[c] = fmaf(g, ybn[c], aw[c]);
            dy[0] = g * w.x; dy[1] = g * w.y; dy[2] = g * w.z; dy[3] = g * w.w;
        }
        if (t >= e_pad) {   // warps of the position role (e_pad is a multiple of 32): reduce d ybn over hidden units
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float v = warp_sum(dy[c]);
                if ((t & 31) == 0) atomicAdd(&sdy[c], v);
            }
        }
        __syncthreads();
        if (t < 4) {
            const float v = sdy[t];
            ag[0] = fmaf(v, xhat[t], ag[0]);   // thread c keeps its own channel in slot 0
            abt[0] += v;
        }
    }
    if (emb_role) {
#pragma unroll
        for (int k = 0; k < OBJ_MAX_CLS; ++k)
            if (k < a.n_cls) atomicAdd(dembed + static_cast<size_t>(k) * a.e + t, acc[k]);
    }
    if (pos_role) {
        atomicAdd(dbp + j, ab);
#pragma unroll
        for (int c = 0; c < 4; ++c) atomicAdd(dwp + j * 4 + c, aw[c]);
    }
    if (t < 4) {
        atomicAdd(dgamma + t, ag[0]);
        atomicAdd(dbeta + t, abt[0]);
    }
}

static int check_args(const b200vsgg_obj_tokens* p, const char* who) {
    if (!p || !p->features || !p->dist || !p->embed || !p->boxes || !p->bn_mean || !p->bn_rstd || !p->bn_gamma ||
        !p->bn_beta || !p->video_of_box || !p->wp || !p->bp || p->rows < 0 || (p->feat_dim & 3) || (p->e & 3) || (p->h & 3) ||
        p->feat_dim <= 0 || p->e <= 0 || p->h <= 0 || p->n_cls <= 0 || p->n_cls > OBJ_MAX_CLS || (p->pos && !p->pe))
        return set_error(B200VSGG_ERR_BAD_ARG, who);
    return 0;
}

static ObjTokArgs to_args(const b200vsgg_obj_tokens* p) {
    ObjTokArgs a;
    a.features = p->features; a.feat_dim = p->feat_dim; a.dist = p->dist; a.n_cls = p->n_cls; a.embed = p->embed; a.e = p->e;
    a.boxes = p->boxes; a.bn_mean = p->bn_mean; a.bn_rstd = p->bn_rstd; a.bn_gamma = p->bn_gamma; a.bn_beta = p->bn_beta;
    a.video_of_box = p->video_of_box; a.wp = p->wp; a.bp = p->bp; a.h = p->h; a.pe = p->pe; a.src = p->src; a.pos = p->pos;
    a.rows = p->rows; a.p_pos = p->p_pos; a.seed_pos = p->seed_pos; a.p_pe = p->p_pe; a.seed_pe = p->seed_pe;
    return a;
}

}  // namespace vsgg

using namespace vsgg;

extern "C" int b200vsgg_obj_tokens_fwd(const b200vsgg_obj_tokens* p, float* x_f32, void* x_bf16, void* stream) {
    if (int rc = check_args(p, "obj_tokens_fwd: bad arg")) return rc;
    if (!x_f32 && !x_bf16) return set_error(B200VSGG_ERR_BAD_ARG, "obj_tokens_fwd: no output");
    if (p->rows == 0) return 0;
    obj_tokens_fwd_kernel<<<p->rows, 256, 0, (cudaStream_t)stream>>>(to_args(p), x_f32, (__nv_bfloat16*)x_bf16);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_obj_tokens_bwd(const b200vsgg_obj_tokens* p, const float* dx, float* dembed, float* dwp,
                                       float* dbp, float* dgamma, float* dbeta, void* stream) {
    if (int rc = check_args(p, "obj_tokens_bwd: bad arg")) return rc;
    const int e_pad = (p->e + 31) / 32 * 32;
    if (!dx || !dembed || !dwp || !dbp || !dgamma || !dbeta || e_pad + p->h > 384)
        return set_error(B200VSGG_ERR_BAD_ARG, "obj_tokens_bwd: bad arg (e rounded up to 32 plus h must be <= 384)");
    if (p->rows == 0) return 0;
    int grid = p->rows < 148 * 2 ? p->rows : 148 * 2;
    obj_tokens_bwd_kernel<<<grid, 384, 0, (cudaStream_t)stream>>>(to_args(p), dx, e_pad, dembed, dwp, dbp, dgamma, dbeta);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}
