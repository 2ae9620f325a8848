// HBM-bound ragged kernels of the TEAT-GT / TokenGT path (lib/teatgt.py:118-240 and
// tools/TokenGT/tokengt/modules/tokenizer.py:217-295 of the reference): node-token gather/concat,
// pairwise edge predicates (centre distance within a frame, cosine similarity between consecutive
// frames of a clip), token assembly (node / edge / special tokens with Laplacian node identifiers and
// type identifiers) and their backward scatters, plus the GELU+dropout row op of the FFN.
// All row kernels use one warp per row and 128-bit accesses.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>

#include "../../include/b200vsgg.h"
#include "common.cuh"

namespace vsgg {

static inline int rows_grid(long long rows, int warps_per_block) {
    long long g = (rows + warps_per_block - 1) / warps_per_block;
    if (g < 1) g = 1;
    if (g > 148 * 32) g = 148 * 32;
    return static_cast<int>(g);
}

// ------------------------------------------------------------------------------------------------
// node tokens: tok[i] = so[feat_row[i], (is_person ? 0 : H1) + 0..H1) | embed[labels[feat_row[i]]]
// ------------------------------------------------------------------------------------------------
__global__ void node_tokens_fwd_kernel(const float* __restrict__ so, int ld_so, const int32_t* __restrict__ feat_row,
                                       const int32_t* __restrict__ is_person, const int64_t* __restrict__ labels,
                                       const float* __restrict__ embed, int n, int H1, int E, float* __restrict__ out_f32,
                                       __nv_bfloat16* __restrict__ out_bf16) {
    const int wpb = blockDim.x >> 5, lane = threadIdx.x & 31;
    const int D = H1 + E;
    for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
        const int r = __ldg(feat_row + i);
        const float* sp = so + static_cast<size_t>(r) * ld_so + (__ldg(is_person + i) ? 0 : H1);
        const float* ep = embed + static_cast<size_t>(__ldg(labels + r)) * E;
        for (int c = lane * 4; c < D; c += 128) {
            const float4 v = c < H1 ? *reinterpret_cast<const float4*>(sp + c) : *reinterpret_cast<const float4*>(ep + c - H1);
            if (out_f32) *reinterpret_cast<float4*>(out_f32 + static_cast<size_t>(i) * D + c) = v;
            if (out_bf16) {
                const float a[4] = {v.x, v.y, v.z, v.w};
                store_bf16x4(out_bf16 + static_cast<size_t>(i) * D + c, a);
            }
        }
    }
}

__global__ void node_tokens_bwd_kernel(const float* __restrict__ dtok, const int32_t* __restrict__ feat_row,
                                       const int32_t* __restrict__ is_person, const int64_t* __restrict__ labels, int n,
                                       int H1, int E, float* __restrict__ dso, int ld_dso, float* __restrict__ dembed) {
    const int wpb = blockDim.x >> 5, lane = threadIdx.x & 31;
    const int D = H1 + E;
    for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
        const int r = __ldg(feat_row + i);
        float* sp = dso + static_cast<size_t>(r) * ld_dso + (__ldg(is_person + i) ? 0 : H1);
        float* ep = dembed ? dembed + static_cast<size_t>(__ldg(labels + r)) * E : nullptr;
        for (int c = lane; c < D; c += 32) {
            const float g = dtok[static_cast<size_t>(i) * D + c];
            if (c < H1) atomicAdd(sp + c, g);
            else if (ep) atomicAdd(ep + c - H1, g);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// edge predicates.  One CTA per frame f (nodes [node_off[f], node_off[f+1]), n <= nmax):
//   spatial[f][a][b] (a < b) = || centre_a - centre_b || <= thr          (lib/teatgt.py:200-203)
//   temporal[f][p][c] = cos(tok[prev frame node p], tok[node c]) >= sim   (lib/teatgt.py:213-217), only
//                       when has_prev[f] (the previous frame belongs to the same clip)
// centres = box centres of the node's box row, fp32 (x1+x2)/2 like box[[0,2]].mean().
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) teat_pair_flags_kernel(const float* __restrict__ tok, int D,
                                                              const float* __restrict__ boxes,
                                                              const int32_t* __restrict__ feat_row,
                                                              const int32_t* __restrict__ node_off,
                                                              const int32_t* __restrict__ has_prev, float thr, float sim,
                                                              int nmax, uint8_t* __restrict__ spatial,
                                                              uint8_t* __restrict__ temporal) {
    const int f = blockIdx.x;
    const int n0 = node_off[f], n = node_off[f + 1] - n0;
    uint8_t* sp = spatial + static_cast<size_t>(f) * nmax * nmax;
    uint8_t* tp = temporal + static_cast<size_t>(f) * nmax * nmax;
    for (int i = threadIdx.x; i < nmax * nmax; i += blockDim.x) {
        const int a = i / nmax, b = i - a * nmax;
        uint8_t flag = 0;
        if (a < b && b < n) {
            const float* ba = boxes + static_cast<size_t>(feat_row[n0 + a]) * 5;
            const float* bb = boxes + static_cast<size_t>(feat_row[n0 + b]) * 5;
            const float ax = (ba[1] + ba[3]) / 2.f, ay = (ba[2] + ba[4]) / 2.f;
            const float bx = (bb[1] + bb[3]) / 2.f, by = (bb[2] + bb[4]) / 2.f;
            const float dx = ax - bx, dy = ay - by;
            const float dist = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
            flag = dist <= thr ? 1 : 0;
        }
        sp[i] = flag;
        tp[i] = 0;
    }
    if (!has_prev[f]) return;
    __syncthreads();
    const int p0 = node_off[f - 1], np = n0 - p0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int pr = warp; pr < np * n; pr += nwarps) {
        const int p = pr / n, c = pr - p * n;
        const float* u = tok + static_cast<size_t>(p0 + p) * D;
        const float* v = tok + static_cast<size_t>(n0 + c) * D;
        float dot = 0.f, uu = 0.f, vv = 0.f;
        for (int d = lane * 4; d < D; d += 128) {
            const float4 x = *reinterpret_cast<const float4*>(u + d);
            const float4 y = *reinterpret_cast<const float4*>(v + d);
            dot += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
            uu += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
            vv += y.x * y.x + y.y * y.y + y.z * y.z + y.w * y.w;
        }
        dot = warp_sum(dot);
        uu = warp_sum(uu);
        vv = warp_sum(vv);
        if (lane == 0) tp[p * nmax + c] = (dot / (sqrtf(uu) * sqrtf(vv)) >= sim) ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------------
// token assembly (tokenizer.py:217-295).  Token t is described by (kind, a, b, c):
//   kind 0: [graph] token, 1: [null] token
//   kind 2: node a      : x = NA[a] + temp[b] + PU[a] + PV[a] + order[1]     (b = frame - first frame of clip)
//   kind 3: edge (a, b) : x = eemb[c] + PU[a] + PV[b] + order[0]             (c = edge type)
// NA = atom_encoder(node tokens) [n, d];  PU / PV = eigvec @ lap_encoder.weight[:, :k]^T / [:, k:]^T  [n, d].
// ------------------------------------------------------------------------------------------------
__global__ void teat_assemble_fwd_kernel(const int32_t* __restrict__ desc, int T, int d, const float* __restrict__ NA,
                                         const float* __restrict__ PU, const float* __restrict__ PV,
                                         const float* __restrict__ temp, const float* __restrict__ eemb,
                                         const float* __restrict__ order, const float* __restrict__ graph_tok,
                                         const float* __restrict__ null_tok, float* __restrict__ x) {
    const int wpb = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (int t = blockIdx.x * wpb + (threadIdx.x >> 5); t < T; t += gridDim.x * wpb) {
        const int4 ds = *reinterpret_cast<const int4*>(desc + 4 * static_cast<size_t>(t));
        const int kind = ds.x, a = ds.y, b = ds.z, c = ds.w;
        float* xp = x + static_cast<size_t>(t) * d;
        for (int col = lane * 4; col < d; col += 128) {
            float4 v;
            if (kind == 0) v = *reinterpret_cast<const float4*>(graph_tok + col);
            else if (kind == 1) v = *reinterpret_cast<const float4*>(null_tok + col);
            else {
                const float4 pu = *reinterpret_cast<const float4*>(PU + static_cast<size_t>(a) * d + col);
                const float4 pv = *reinterpret_cast<const float4*>(PV + static_cast<size_t>(kind == 2 ? a : b) * d + col);
                const float4 od = *reinterpret_cast<const float4*>(order + (kind == 2 ? d : 0) + col);
                float4 base, extra;
                if (kind == 2) {
                    base = *reinterpret_cast<const float4*>(NA + static_cast<size_t>(a) * d + col);
                    extra = *reinterpret_cast<const float4*>(temp + static_cast<size_t>(b) * d + col);
                } else {
                    base = *reinterpret_cast<const float4*>(eemb + static_cast<size_t>(c) * d + col);
                    extra = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                v.x = base.x + extra.x + pu.x + pv.x + od.x;
                v.y = base.y + extra.y + pu.y + pv.y + od.y;
                v.z = base.z + extra.z + pu.z + pv.z + od.z;
                v.w = base.w + extra.w + pu.w + pv.w + od.w;
            }
            *reinterpret_cast<float4*>(xp + col) = v;
        }
    }
}

// backward: scatter dx into dNA (one writer per node: plain store), dPU / dPV (atomics), and the small
// tables dtemp / deemb / dorder / dgraph / dnull (atomics into [rows, d] fp32, pre-zeroed).
__global__ void teat_assemble_bwd_kernel(const int32_t* __restrict__ desc, int T, int d, const float* __restrict__ dx,
                                         float* __restrict__ dNA, float* __restrict__ dPU, float* __restrict__ dPV,
                                         float* __restrict__ dtemp, float* __restrict__ deemb, float* __restrict__ dorder,
                                         float* __restrict__ dgraph, float* __restrict__ dnull) {
    const int wpb = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (int t = blockIdx.x * wpb + (threadIdx.x >> 5); t < T; t += gridDim.x * wpb) {
        const int4 ds = *reinterpret_cast<const int4*>(desc + 4 * static_cast<size_t>(t));
        const int kind = ds.x, a = ds.y, b = ds.z, c = ds.w;
        const float* gp = dx + static_cast<size_t>(t) * d;
        for (int col = lane; col < d; col += 32) {
            const float g = gp[col];
            if (kind == 0) atomicAdd(dgraph + col, g);
            else if (kind == 1) atomicAdd(dnull + col, g);
            else if (kind == 2) {
                dNA[static_cast<size_t>(a) * d + col] = g;
                atomicAdd(dtemp + static_cast<size_t>(b) * d + col, g);
                atomicAdd(dPU + static_cast<size_t>(a) * d + col, g);
                atomicAdd(dPV + static_cast<size_t>(a) * d + col, g);
                atomicAdd(dorder + d + col, g);
            } else {
                atomicAdd(deemb + static_cast<size_t>(c) * d + col, g);
                atomicAdd(dPU + static_cast<size_t>(a) * d + col, g);
                atomicAdd(dPV + static_cast<size_t>(b) * d + col, g);
                atomicAdd(dorder + col, g);
            }
        }
    }
}

// out = bf16(dropout(act(x))), x bf16; act 0 none, 1 relu, 2 gelu(erf).  Dropout index = row*cols + col
// (the GEMM epilogue's convention, so the backward GEMM regenerates the same mask).
__global__ void act_dropout_kernel(const __nv_bfloat16* __restrict__ x, int ld_x, long long rows, int cols, int act,
                                   float p, unsigned long long seed, __nv_bfloat16* __restrict__ out, int ld_o) {
    const int vec = cols >> 3;
    const long long total = rows * vec;
    const float inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
    const uint32_t thr = p > 0.f ? static_cast<uint32_t>(p * 4294967296.0) : 0u;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / vec;
        const int c = static_cast<int>(i - r * vec) * 8;
        float v[8];
        load_bf16x8(x + static_cast<size_t>(r) * ld_x + c, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float y = v[j];
            if (act == 1) y = fmaxf(y, 0.f);
            else if (act == 2) y = 0.5f * y * (1.0f + erff(y * 0.70710678118654752f));
            if (thr != 0u) y = hash_u32(seed, static_cast<unsigned long long>(r) * cols + c + j) >= thr ? y * inv_keep : 0.f;
            v[j] = y;
        }
        store_bf16x8(out + static_cast<size_t>(r) * ld_o + c, v);
    }
}

}  // namespace vsgg

using namespace vsgg;

extern "C" int b200vsgg_node_tokens_fwd(const float* so, int32_t ld_so, const int32_t* feat_row, const int32_t* is_person,
                                        const int64_t* labels, const float* embed, int32_t n, int32_t h1, int32_t e,
                                        float* out_f32, void* out_bf16, void* stream) {
    if (!so || !feat_row || !is_person || !labels || !embed || (h1 & 3) || (e & 3) || (ld_so & 3))
        return set_error(B200VSGG_ERR_BAD_ARG, "node_tokens_fwd: bad arg");
    if (n == 0) return 0;
    node_tokens_fwd_kernel<<<rows_grid(n, 8), 256, 0, (cudaStream_t)stream>>>(so, ld_so, feat_row, is_person, labels, embed,
                                                                            n, h1, e, out_f32, (__nv_bfloat16*)out_bf16);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_node_tokens_bwd(const float* dtok, const int32_t* feat_row, const int32_t* is_person,
                                        const int64_t* labels, int32_t n, int32_t h1, int32_t e, float* dso,
                                        int32_t ld_dso, float* dembed, void* stream) {
    if (!dtok || !feat_row || !is_person || !labels || !dso) return set_error(B200VSGG_ERR_BAD_ARG, "node_tokens_bwd: bad arg");
    if (n == 0) return 0;
    node_tokens_bwd_kernel<<<rows_grid(n, 8), 256, 0, (cudaStream_t)stream>>>(dtok, feat_row, is_person, labels, n, h1, e,
                                                                            dso, ld_dso, dembed);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_teat_pair_flags(const float* tok, int32_t d, const float* boxes, const int32_t* feat_row,
                                        const int32_t* node_off, const int32_t* has_prev, int32_t n_frames, float thr,
                                        float sim, int32_t nmax, uint8_t* spatial, uint8_t* temporal, void* stream) {
    if (!tok || !boxes || !feat_row || !node_off || !has_prev || !spatial || !temporal || (d & 3) || nmax <= 0)
        return set_error(B200VSGG_ERR_BAD_ARG, "teat_pair_flags: bad arg");
    if (n_frames == 0) return 0;
    teat_pair_flags_kernel<<<n_frames, 256, 0, (cudaStream_t)stream>>>(tok, d, boxes, feat_row, node_off, has_prev, thr,
                                                                      sim, nmax, spatial, temporal);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_teat_assemble_fwd(const int32_t* desc, int32_t n_tokens, int32_t d, const float* na,
                                          const float* pu, const float* pv, const float* temp, const float* eemb,
                                          const float* order, const float* graph_tok, const float* null_tok, float* x,
                                          void* stream) {
    if (!desc || !na || !pu || !pv || !temp || !eemb || !order || !graph_tok || !null_tok || !x || (d & 3))
        return set_error(B200VSGG_ERR_BAD_ARG, "teat_assemble_fwd: bad arg");
    if (n_tokens == 0) return 0;
    teat_assemble_fwd_kernel<<<rows_grid(n_tokens, 8), 256, 0, (cudaStream_t)stream>>>(desc, n_tokens, d, na, pu, pv, temp,
                                                                                     eemb, order, graph_tok, null_tok, x);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_teat_assemble_bwd(const int32_t* desc, int32_t n_tokens, int32_t d, const float* dx, float* dna,
                                          float* dpu, float* dpv, float* dtemp, float* deemb, float* dorder, float* dgraph,
                                          float* dnull, void* stream) {
    if (!desc || !dx || !dna || !dpu || !dpv || !dtemp || !deemb || !dorder || !dgraph || !dnull)
        return set_error(B200VSGG_ERR_BAD_ARG, "teat_assemble_bwd: bad arg");
    if (n_tokens == 0) return 0;
    teat_assemble_bwd_kernel<<<rows_grid(n_tokens, 8), 256, 0, (cudaStream_t)stream>>>(desc, n_tokens, d, dx, dna, dpu, dpv,
                                                                                     dtemp, deemb, dorder, dgraph, dnull);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_act_dropout_bf16(const void* x, int32_t ld_x, int64_t rows, int32_t cols, int32_t act, float p,
                                         uint64_t seed, void* out, int32_t ld_o, void* stream) {
    if (!x || !out || cols <= 0 || (cols & 7) || (ld_x & 7) || (ld_o & 7)) return set_error(B200VSGG_ERR_BAD_ARG, "act_dropout: bad arg");
    if (rows == 0) return 0;
    long long items = rows * (cols >> 3);
    long long g = (items + 1023) / 1024;
    if (g > 148 * 32) g = 148 * 32;
    act_dropout_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, ld_x, rows, cols, act, p, seed,
                                                                 (__nv_bfloat16*)out, ld_o);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}
