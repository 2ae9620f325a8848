// HBM-bound row kernels of the relation path: segment offsets, row gathers (pair tokens, temporal
// windows, 'latter' scatter-back), LayerNorm forward/backward, casts with dropout, column sums.
// All use 128-bit loads/stores on rows whose length is a multiple of 4 (fp32) / 8 (bf16).
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>

#include "../../include/b200vsgg.h"
#include "common.cuh"

namespace vsgg {

static inline int grid_for(long long work_items, int per_block, int cap = 148 * 16) {
    long long g = (work_items + per_block - 1) / per_block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return static_cast<int>(g);
}

// ------------------------------------------------------------------------------------------------
// frame offsets: offsets[f] = first pair row whose frame id >= f  (im_idx sorted ascending)
// ------------------------------------------------------------------------------------------------
__global__ void frame_offsets_kernel(const float* __restrict__ im_idx, int n_pairs, int n_frames,
                                     int32_t* __restrict__ offsets) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f > n_frames) return;
    int lo = 0, hi = n_pairs;
    const float target = static_cast<float>(f);
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(im_idx + mid) < target) lo = mid + 1; else hi = mid;
    }
    offsets[f] = lo;
}

// ------------------------------------------------------------------------------------------------
// gather_rows: out[t, :] = src[idx[t], :] (+ add_table[add_idx[t], :]),  fp32 source.
//   writes any of: out_f32, out_bf16 (plain), out_bf16_added (with the additive row).
//   idx == nullptr means identity.  One warp per row, float4 accesses.
// ------------------------------------------------------------------------------------------------
__global__ void gather_rows_kernel(const float* __restrict__ src, int ld_src, const int32_t* __restrict__ idx,
                                   const float* __restrict__ add_table, const int32_t* __restrict__ add_idx,
                                   int rows, int cols, float* __restrict__ out_f32, int ld_f32,
                                   __nv_bfloat16* __restrict__ out_bf16, int ld_b, __nv_bfloat16* __restrict__ out_added,
                                   int ld_a) {
    const int warps_per_block = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < rows; r += gridDim.x * warps_per_block) {
        const int s = idx ? __ldg(idx + r) : r;
        const float* sp = src + static_cast<size_t>(s) * ld_src;
        const float* ap = (add_table && out_added) ? add_table + static_cast<size_t>(__ldg(add_idx + r)) * cols : nullptr;
        for (int c = lane * 4; c < cols; c += 128) {
            const float4 v = *reinterpret_cast<const float4*>(sp + c);
            if (out_f32) *reinterpret_cast<float4*>(out_f32 + static_cast<size_t>(r) * ld_f32 + c) = v;
            if (out_bf16) {
                const float a[4] = {v.x, v.y, v.z, v.w};
                store_bf16x4(out_bf16 + static_cast<size_t>(r) * ld_b + c, a);
            }
            if (out_added) {
                float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ap) p = __ldg(reinterpret_cast<const float4*>(ap + c));
                const float a[4] = {v.x + p.x, v.y + p.y, v.z + p.z, v.w + p.w};
                store_bf16x4(out_added + static_cast<size_t>(r) * ld_a + c, a);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// scatter_add_rows (gather form, deterministic): out[n,:] = sum_{k<2, idx[n*2+k]>=0} src[idx[n*2+k], :]
// (+ base[n,:]).  Backward of the temporal-window gather: each pair row is read by <= 2 windows.
// ------------------------------------------------------------------------------------------------
__global__ void gather2_sum_rows_kernel(const float* __restrict__ src, int ld_src, const int32_t* __restrict__ idx2,
                                        const float* __restrict__ base, int ld_base, int rows, int cols,
                                        float* __restrict__ out_f32, int ld_f32, __nv_bfloat16* __restrict__ out_bf16,
                                        int ld_b) {
    const int warps_per_block = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < rows; r += gridDim.x * warps_per_block) {
        const int i0 = __ldg(idx2 + 2 * r), i1 = __ldg(idx2 + 2 * r + 1);
        for (int c = lane * 4; c < cols; c += 128) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (base) v = *reinterpret_cast<const float4*>(base + static_cast<size_t>(r) * ld_base + c);
            if (i0 >= 0) {
                const float4 a = *reinterpret_cast<const float4*>(src + static_cast<size_t>(i0) * ld_src + c);
                v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
            }
            if (i1 >= 0) {
                const float4 a = *reinterpret_cast<const float4*>(src + static_cast<size_t>(i1) * ld_src + c);
                v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
            }
            if (out_f32) *reinterpret_cast<float4*>(out_f32 + static_cast<size_t>(r) * ld_f32 + c) = v;
            if (out_bf16) {
                const float a[4] = {v.x, v.y, v.z, v.w};
                store_bf16x4(out_bf16 + static_cast<size_t>(r) * ld_b + c, a);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// bf16 forms of the two gathers (16-byte chunks, one warp per row).  They move rows between the PAIR layout [N, .] and
// the WINDOW layout [M2, .] of the temporal decoder without a detour through fp32:
//   gather_rows_bf16       out[t,:] = idx[t] >= 0 ? src[idx[t],:] : 0        (negative index = zero row)
//   gather2_sum_rows_bf16  out[n,:] = bf16(sum_{k<2, idx2[2n+k] >= 0} float(src[idx2[2n+k],:]))
// ------------------------------------------------------------------------------------------------
__global__ void gather_rows_bf16_kernel(const __nv_bfloat16* __restrict__ src, int ld_src, const int32_t* __restrict__ idx,
                                        int rows, int cols, __nv_bfloat16* __restrict__ out, int ld_out) {
    const int warps_per_block = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < rows; r += gridDim.x * warps_per_block) {
        const int s = idx ? __ldg(idx + r) : r;
        const __nv_bfloat16* sp = src + static_cast<size_t>(s < 0 ? 0 : s) * ld_src;
        __nv_bfloat16* op = out + static_cast<size_t>(r) * ld_out;
        for (int c = lane * 8; c < cols; c += 256) {
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (s >= 0) v = *reinterpret_cast<const uint4*>(sp + c);
            *reinterpret_cast<uint4*>(op + c) = v;
        }
    }
}

__global__ void gather2_sum_rows_bf16_kernel(const __nv_bfloat16* __restrict__ src, int ld_src,
                                             const int32_t* __restrict__ idx2, int rows, int cols,
                                             __nv_bfloat16* __restrict__ out, int ld_out) {
    const int warps_per_block = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < rows; r += gridDim.x * warps_per_block) {
        const int i0 = __ldg(idx2 + 2 * r), i1 = __ldg(idx2 + 2 * r + 1);
        __nv_bfloat16* op = out + static_cast<size_t>(r) * ld_out;
        for (int c = lane * 8; c < cols; c += 256) {
            float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (i0 >= 0) load_bf16x8(src + static_cast<size_t>(i0) * ld_src + c, v);
            if (i1 >= 0) {
                float a[8];
                load_bf16x8(src + static_cast<size_t>(i1) * ld_src + c, a);
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] += a[k];
            }
            store_bf16x8(op + c, v);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// pair_concat: the [N,1936] pair token (lib/tempura.py:537-563).
//   cols [0,512)      = so[pair_idx[n,0], 0:512]        (subj_fc output of the person box)
//   cols [512,1024)   = so[pair_idx[n,1], 512:1024]     (obj_fc output of the object box)
//   cols [1024,1536)  = already written by the vr_fc GEMM into tok_f32 (read back for the bf16 copy)
//   cols [1536,1736)  = embed1[labels[pair_idx[n,0]]],  cols [1736,1936) = embed2[labels[pair_idx[n,1]]]
// ------------------------------------------------------------------------------------------------
__global__ void pair_concat_kernel(const float* __restrict__ so, const int64_t* __restrict__ pair_idx,
                                   const int64_t* __restrict__ labels, const float* __restrict__ embed1,
                                   const float* __restrict__ embed2, int n_pairs, float* __restrict__ tok_f32,
                                   __nv_bfloat16* __restrict__ tok_bf16) {
    constexpr int D = 1936;
    const int warps_per_block = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < n_pairs; r += gridDim.x * warps_per_block) {
        const int64_t s = __ldg(pair_idx + 2 * r), o = __ldg(pair_idx + 2 * r + 1);
        const int64_t ls = __ldg(labels + s), lo = __ldg(labels + o);
        float* tf = tok_f32 + static_cast<size_t>(r) * D;
        __nv_bfloat16* tb = tok_bf16 + static_cast<size_t>(r) * D;
        for (int c = lane * 4; c < D; c += 128) {
            float4 v;
            if (c < 512) v = *reinterpret_cast<const float4*>(so + s * 1024 + c);
            else if (c < 1024) v = *reinterpret_cast<const float4*>(so + o * 1024 + c);
            else if (c < 1536) v = *reinterpret_cast<const float4*>(tf + c);
            else if (c < 1736) v = __ldg(reinterpret_cast<const float4*>(embed1 + ls * 200 + (c - 1536)));
            else v = __ldg(reinterpret_cast<const float4*>(embed2 + lo * 200 + (c - 1736)));
            if (c < 1024 || c >= 1536) *reinterpret_cast<float4*>(tf + c) = v;
            const float a[4] = {v.x, v.y, v.z, v.w};
            store_bf16x4(tb + c, a);
        }
    }
}

// Backward of pair_concat for the gathered so-columns: dso[b, 0:512] = sum over pairs with subject b of
// dtok[n,0:512]; dso[b,512:1024] = sum over pairs with object b.  Pairs of a box are contiguous
// (subject) or unique (object) in Action-Genome order, but we do not rely on it: atomics on fp32.
__global__ void pair_concat_bwd_kernel(const float* __restrict__ dtok, const int64_t* __restrict__ pair_idx,
                                       const int64_t* __restrict__ labels, int n_pairs, float* __restrict__ dso,
                                       float* __restrict__ dembed1, float* __restrict__ dembed2) {
    constexpr int D = 1936;
    const int warps_per_block = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < n_pairs; r += gridDim.x * warps_per_block) {
        const int64_t s = __ldg(pair_idx + 2 * r), o = __ldg(pair_idx + 2 * r + 1);
        const int64_t ls = __ldg(labels + s), lo = __ldg(labels + o);
        const float* g = dtok + static_cast<size_t>(r) * D;
        for (int c = lane; c < D; c += 32) {
            const float v = g[c];
            if (c < 512) atomicAdd(dso + s * 1024 + c, v);
            else if (c < 1024) atomicAdd(dso + o * 1024 + c, v);
            else if (c < 1536) { /* vr_fc columns: consumed directly by the vr_fc backward GEMMs */ }
            else if (c < 1736) { if (dembed1) atomicAdd(dembed1 + ls * 200 + (c - 1536), v); }
            else { if (dembed2) atomicAdd(dembed2 + lo * 200 + (c - 1736), v); }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm forward: y = (x - mean) * rstd * gamma + beta over the last dim (eps inside sqrt, biased
// variance — torch.nn.LayerNorm).  One warp per row, row cached in registers (cols <= 4096).
// Optional outputs: y fp32, y bf16, bf16(y + add_table[add_idx[row]]).  Saves mean / rstd.
// ------------------------------------------------------------------------------------------------
// LN_MAX_VEC = float4 per lane (template): 8 -> cols <= 1024, 16 -> cols <= 2048, 20 -> cols <= 2560 (the
// 2376-wide object-sequence encoder, lib/tempura.py:88-92).
template <int LN_MAX_VEC>
__global__ void layernorm_fwd_kernel(const float* __restrict__ x, int ld_x, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, int rows, int cols, float eps,
                                     float* __restrict__ y_f32, int ld_y, __nv_bfloat16* __restrict__ y_bf16, int ld_b,
                                     const float* __restrict__ add_table, const int32_t* __restrict__ add_idx,
                                     __nv_bfloat16* __restrict__ y_added, int ld_a, float* __restrict__ mean_out,
                                     float* __restrict__ rstd_out) {
    const int warps_per_block = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nvec = cols >> 2;
    for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < rows; r += gridDim.x * warps_per_block) {
        const float4* xp = reinterpret_cast<const float4*>(x + static_cast<size_t>(r) * ld_x);
        float4 v[LN_MAX_VEC];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAX_VEC; ++i) {
            const int c = lane + i * 32;
            if (c < nvec) {
                v[i] = xp[c];
                sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
            }
        }
        const float mean = warp_sum(sum) / cols;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAX_VEC; ++i) {
            const int c = lane + i * 32;
            if (c < nvec) {
                const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
                sq += (a * a + b * b) + (cc * cc + d * d);
            }
        }
        const float rstd = rsqrtf(warp_sum(sq) / cols + eps);
        if (lane == 0) {
            if (mean_out) mean_out[r] = mean;
            if (rstd_out) rstd_out[r] = rstd;
        }
        const float* ap = (add_table && y_added) ? add_table + static_cast<size_t>(__ldg(add_idx + r)) * cols : nullptr;
#pragma unroll
        for (int i = 0; i < LN_MAX_VEC; ++i) {
            const int c = lane + i * 32;
            if (c < nvec) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
                const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c);
                float4 o;
                o.x = (v[i].x - mean) * rstd * g.x + b.x;
                o.y = (v[i].y - mean) * rstd * g.y + b.y;
                o.z = (v[i].z - mean) * rstd * g.z + b.z;
                o.w = (v[i].w - mean) * rstd * g.w + b.w;
                if (y_f32) reinterpret_cast<float4*>(y_f32 + static_cast<size_t>(r) * ld_y)[c] = o;
                if (y_bf16) {
                    const float a[4] = {o.x, o.y, o.z, o.w};
                    store_bf16x4(y_bf16 + static_cast<size_t>(r) * ld_b + c * 4, a);
                }
                if (y_added) {
                    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ap) p = __ldg(reinterpret_cast<const float4*>(ap) + c);
                    const float a[4] = {o.x + p.x, o.y + p.y, o.z + p.z, o.w + p.w};
                    store_bf16x4(y_added + static_cast<size_t>(r) * ld_a + c * 4, a);
                }
            }
        }
    }
}

// LayerNorm backward.  dx = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat)).
// Optional: dx_bf16 = bf16(dropout(dx)) for the next backward GEMM (dropout of the branch that fed
// the residual sum); dx_f32 is the undropped gradient of the skip path.
// ------------------------------------------------------------------------------------------------
// LayerNorm backward, v2: the row gradient (two light passes per row, second pass re-reads the row from
// L1/L2 instead of holding it in ~130 registers: 8x the resident warps of the register-blocked version)
// and the parameter gradients (column sums, seg_colstats-style) are separate kernels.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm_bwd_dx_kernel(
    const float* __restrict__ dy, int ld_dy, const float* __restrict__ x, int ld_x, const float* __restrict__ gamma,
    const float* __restrict__ mean, const float* __restrict__ rstd, int rows, int cols, float* __restrict__ dx_f32,
    int ld_dx, __nv_bfloat16* __restrict__ dx_bf16, int ld_b, float drop_p, unsigned long long drop_seed,
    const float* __restrict__ base, int ld_base) {
    const int warps_per_block = blockDim.x >> 5, lane = threadIdx.x & 31;
    const int nvec = cols >> 2;
    const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    const uint32_t thr = drop_p > 0.f ? static_cast<uint32_t>(drop_p * 4294967296.0) : 0u;
    const float4* gp = reinterpret_cast<const float4*>(gamma);
    for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < rows; r += gridDim.x * warps_per_block) {
        const float4* dyp = reinterpret_cast<const float4*>(dy + static_cast<size_t>(r) * ld_dy);
        const float4* xp = reinterpret_cast<const float4*>(x + static_cast<size_t>(r) * ld_x);
        const float mu = mean[r], rs = rstd[r];
        float s1 = 0.f, s2 = 0.f;
        for (int c = lane; c < nvec; c += 32) {
            const float4 d = dyp[c], xv = xp[c], g = __ldg(gp + c);
            const float g0 = d.x * g.x, g1 = d.y * g.y, g2 = d.z * g.z, g3 = d.w * g.w;
            s1 += (g0 + g1) + (g2 + g3);
            s2 += (g0 * (xv.x - mu) + g1 * (xv.y - mu)) + (g2 * (xv.z - mu) + g3 * (xv.w - mu));
        }
        const float m1 = warp_sum(s1) / cols, m2 = warp_sum(s2) * rs / cols;
        for (int c = lane; c < nvec; c += 32) {
            const float4 d = dyp[c], xv = xp[c], g = __ldg(gp + c);
            float4 o;
            o.x = rs * (d.x * g.x - m1 - (xv.x - mu) * rs * m2);
            o.y = rs * (d.y * g.y - m1 - (xv.y - mu) * rs * m2);
            o.z = rs * (d.z * g.z - m1 - (xv.z - mu) * rs * m2);
            o.w = rs * (d.w * g.w - m1 - (xv.w - mu) * rs * m2);
            if (base != nullptr) {
                const float4 bs = reinterpret_cast<const float4*>(base + static_cast<size_t>(r) * ld_base)[c];
                o.x += bs.x; o.y += bs.y; o.z += bs.z; o.w += bs.w;
            }
            if (dx_f32) reinterpret_cast<float4*>(dx_f32 + static_cast<size_t>(r) * ld_dx)[c] = o;
            if (dx_bf16) {
                float a[4] = {o.x, o.y, o.z, o.w};
                if (thr) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t h = hash_u32(drop_seed, static_cast<size_t>(r) * cols + c * 4 + j);
                        a[j] = h >= thr ? a[j] * inv_keep : 0.f;
                    }
                }
                store_bf16x4(dx_bf16 + static_cast<size_t>(r) * ld_b + c * 4, a);
            }
        }
    }
}

// dgamma[c] += sum_r dy[r,c] * (x[r,c] - mean[r]) * rstd[r];  dbeta[c] += sum_r dy[r,c].
// block = (32 float4 column vectors, 8 row phases); grid = (ceil(cols/128), row chunks of 256).
__global__ void __launch_bounds__(256) layernorm_bwd_param_kernel(const float* __restrict__ dy, int ld_dy,
                                                                  const float* __restrict__ x, int ld_x,
                                                                  const float* __restrict__ mean,
                                                                  const float* __restrict__ rstd, int rows, int cols,
                                                                  float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int col = (blockIdx.x * 32 + tx) * 4;
    const int r0 = blockIdx.y * 256, r1 = min(rows, r0 + 256);
    float4 sg = make_float4(0.f, 0.f, 0.f, 0.f), sb = sg;
    if (col < cols) {
#pragma unroll 4
        for (int r = r0 + ty; r < r1; r += 8) {
            const float4 d = *reinterpret_cast<const float4*>(dy + static_cast<size_t>(r) * ld_dy + col);
            const float4 xv = *reinterpret_cast<const float4*>(x + static_cast<size_t>(r) * ld_x + col);
            const float mu = __ldg(mean + r), rs = __ldg(rstd + r);
            sg.x += d.x * (xv.x - mu) * rs; sg.y += d.y * (xv.y - mu) * rs;
            sg.z += d.z * (xv.z - mu) * rs; sg.w += d.w * (xv.w - mu) * rs;
            sb.x += d.x; sb.y += d.y; sb.z += d.z; sb.w += d.w;
        }
    }
    __shared__ float4 red[2][8][32];
    red[0][ty][tx] = sg;
    red[1][ty][tx] = sb;
    __syncthreads();
    if (ty < 2 && col < cols) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int y = 0; y < 8; ++y) {
            const float4 v = red[ty][y][tx];
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        float* dst = (ty == 0 ? dgamma : dbeta) + col;
        atomicAdd(dst, s.x); atomicAdd(dst + 1, s.y); atomicAdd(dst + 2, s.z); atomicAdd(dst + 3, s.w);
    }
}

// ------------------------------------------------------------------------------------------------
// cast fp32 -> bf16 with optional dropout mask (same hash as the GEMM epilogue: index = row*cols+col).
// ------------------------------------------------------------------------------------------------
__global__ void cast_dropout_kernel(const float* __restrict__ x, int ld_x, int rows, int cols,
                                    __nv_bfloat16* __restrict__ out, int ld_o, float drop_p,
                                    unsigned long long seed) {
    const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    const uint32_t thr = drop_p > 0.f ? static_cast<uint32_t>(drop_p * 4294967296.0) : 0u;
    const int nvec = cols >> 2;
    const long long total = static_cast<long long>(rows) * nvec;
    const bool small = total <= 0xffffffffLL;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        long long rq;
        int c;
        divmod_idx(i, nvec, small, rq, c);
        const int r = static_cast<int>(rq);
        const float4 v = *reinterpret_cast<const float4*>(x + static_cast<size_t>(r) * ld_x + c * 4);
        float a[4] = {v.x, v.y, v.z, v.w};
        if (thr) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t h = hash_u32(seed, static_cast<size_t>(r) * cols + c * 4 + j);
                a[j] = h >= thr ? a[j] * inv_keep : 0.f;
            }
        }
        store_bf16x4(out + static_cast<size_t>(r) * ld_o + c * 4, a);
    }
}

// Split-precision operand copy: out[r] = [ hi | lo | hi ] (3 * cols bf16) with hi = bf16(x), lo = bf16(x - hi), so that
// a bf16 tensor-core GEMM against [ W_hi | W_hi | W_lo ] evaluates x.W to ~2^-17 relative (the hi*hi, lo*hi and
// hi*lo terms; lo*lo ~ 2^-18 is dropped).  Used for the predicate-head GEMM (tools/utils/gmm_heads.py:37-76):
// 0.1 % of the step's flops, but its logits feed softmax / sigmoid outputs compared at 1e-3.
__global__ void split3_bf16_kernel(const float* __restrict__ x, int ld_x, int rows, int cols,
                                   __nv_bfloat16* __restrict__ out, int ld_o) {
    const int nvec = cols >> 2;
    const long long total = static_cast<long long>(rows) * nvec;
    const bool small = total <= 0xffffffffLL;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        long long rq;
        int c;
        divmod_idx(i, nvec, small, rq, c);
        const int r = static_cast<int>(rq);
        const float4 v = *reinterpret_cast<const float4*>(x + static_cast<size_t>(r) * ld_x + c * 4);
        float hi[4] = {v.x, v.y, v.z, v.w}, lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float h = __bfloat162float(__float2bfloat16_rn(hi[j]));
            lo[j] = hi[j] - h;
            hi[j] = h;
        }
        __nv_bfloat16* o = out + static_cast<size_t>(r) * ld_o + c * 4;
        store_bf16x4(o, hi);
        store_bf16x4(o + cols, lo);
        store_bf16x4(o + 2 * cols, hi);
    }
}

// ------------------------------------------------------------------------------------------------
// column sums of a bf16 or fp32 [rows, cols] matrix into fp32 out[group, cols] (+=), where
// group = group_idx[row] (or 0).  Used for bias gradients and the position-embedding gradient.
// Each CTA owns a 64-column stripe x a slab of rows; partials go through one atomicAdd per column.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, int ld_x, int rows, int cols,
                              const int32_t* __restrict__ group_idx, int n_groups, float* __restrict__ out) {
    // blockDim = (64, 4): x -> column inside the stripe, y -> row phase
    const int col = blockIdx.x * 64 + threadIdx.x;
    const int rows_per_slab = (rows + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * rows_per_slab;
    const int r1 = min(rows, r0 + rows_per_slab);
    float acc[2] = {0.f, 0.f};  // up to 2 groups (position ids 0/1)
    if (col < cols) {
        for (int r = r0 + threadIdx.y; r < r1; r += blockDim.y) {
            float v;
            if constexpr (sizeof(T) == 2) v = __bfloat162float(x[static_cast<size_t>(r) * ld_x + col]);
            else v = x[static_cast<size_t>(r) * ld_x + col];
            const int g = group_idx ? __ldg(group_idx + r) : 0;
            if (g == 0) acc[0] += v; else acc[1] += v;
        }
    }
    __shared__ float sm[2][4][64];
    sm[0][threadIdx.y][threadIdx.x] = acc[0];
    sm[1][threadIdx.y][threadIdx.x] = acc[1];
    __syncthreads();
    if (threadIdx.y == 0 && col < cols) {
        for (int g = 0; g < n_groups; ++g) {
            const float s = sm[g][0][threadIdx.x] + sm[g][1][threadIdx.x] + sm[g][2][threadIdx.x] + sm[g][3][threadIdx.x];
            atomicAdd(out + static_cast<size_t>(g) * cols + col, s);
        }
    }
}

// Vectorised variant for bf16 rows (cols % 8 == 0, 16-byte aligned): thread = 8 columns (one 16-byte load per row),
// block = 32 column vectors x 8 row phases, grid = (ceil(cols/256), row slabs).  The scalar kernel above reads 2 bytes per
// thread and reached 1.4 TB/s on the [31.8 k, 3872] position-embedding sums of the decoder backward (0.26 ms per layer).
__global__ void __launch_bounds__(256) colsum_vec_kernel(const __nv_bfloat16* __restrict__ x, int ld_x, int rows, int cols,
                                                         const int32_t* __restrict__ group_idx, int n_groups,
                                                         float* __restrict__ out) {
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int col = (blockIdx.x * 32 + tx) * 8;
    const int rows_per_slab = (rows + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * rows_per_slab, r1 = min(rows, r0 + rows_per_slab);
    float acc[2][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[0][j] = 0.f; acc[1][j] = 0.f; }
    if (col < cols) {
#pragma unroll 4
        for (int r = r0 + ty; r < r1; r += 8) {
            float v[8];
            load_bf16x8(x + static_cast<size_t>(r) * ld_x + col, v);
            const bool g1 = group_idx != nullptr && __ldg(group_idx + r) != 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                acc[0][j] += g1 ? 0.f : v[j];
                acc[1][j] += g1 ? v[j] : 0.f;
            }
        }
    }
    __shared__ float red[2][8][32][9];
#pragma unroll
    for (int j = 0; j < 8; ++j) { red[0][ty][tx][j] = acc[0][j]; red[1][ty][tx][j] = acc[1][j]; }
    __syncthreads();
    for (int o = threadIdx.x; o < 256 * n_groups; o += 256) {
        const int g = o >> 8, c = o & 255;
        float sum = 0.f;
#pragma unroll
        for (int y = 0; y < 8; ++y) sum += red[g][y][c >> 3][c & 7];
        const int gc = blockIdx.x * 256 + c;
        if (gc < cols) atomicAdd(out + static_cast<size_t>(g) * cols + gc, sum);
    }
}

}  // namespace vsgg

using namespace vsgg;

// ================================================================================================
// C-ABI
// ================================================================================================
extern "C" int b200vsgg_frame_offsets(const float* im_idx, int32_t n_pairs, int32_t n_frames, int32_t* offsets,
                                      void* stream) {
    if (!im_idx || !offsets || n_pairs < 0 || n_frames < 0) return set_error(B200VSGG_ERR_BAD_ARG, "frame_offsets: bad arg");
    const int threads = 128;
    frame_offsets_kernel<<<(n_frames + 1 + threads - 1) / threads, threads, 0, (cudaStream_t)stream>>>(
        im_idx, n_pairs, n_frames, offsets);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_gather_rows(const float* src, int32_t ld_src, const int32_t* idx, const float* add_table,
                                    const int32_t* add_idx, int32_t rows, int32_t cols, float* out_f32, int32_t ld_f32,
                                    void* out_bf16, int32_t ld_bf16, void* out_bf16_added, int32_t ld_added,
                                    void* stream) {
    if (!src || rows < 0 || cols <= 0 || (cols & 7)) return set_error(B200VSGG_ERR_BAD_ARG, "gather_rows: cols % 8 != 0");
    if (rows == 0) return 0;
    gather_rows_kernel<<<grid_for(rows, 8), 256, 0, (cudaStream_t)stream>>>(
        src, ld_src, idx, add_table, add_idx, rows, cols, out_f32, ld_f32, (__nv_bfloat16*)out_bf16, ld_bf16,
        (__nv_bfloat16*)out_bf16_added, ld_added);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_gather2_sum_rows(const float* src, int32_t ld_src, const int32_t* idx2, const float* base,
                                         int32_t ld_base, int32_t rows, int32_t cols, float* out_f32, int32_t ld_f32,
                                         void* out_bf16, int32_t ld_bf16, void* stream) {
    if (!src || !idx2 || rows < 0 || cols <= 0 || (cols & 7)) return set_error(B200VSGG_ERR_BAD_ARG, "gather2_sum_rows: bad arg");
    if (rows == 0) return 0;
    gather2_sum_rows_kernel<<<grid_for(rows, 8), 256, 0, (cudaStream_t)stream>>>(
        src, ld_src, idx2, base, ld_base, rows, cols, out_f32, ld_f32, (__nv_bfloat16*)out_bf16, ld_bf16);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

static bool bf16_rows_ok(const void* src, int32_t ld_src, const void* out, int32_t ld_out, int32_t cols) {
    return src && out && cols > 0 && !(cols & 7) && !(ld_src & 7) && !(ld_out & 7) &&
           !(reinterpret_cast<uintptr_t>(src) & 15u) && !(reinterpret_cast<uintptr_t>(out) & 15u);
}

extern "C" int b200vsgg_gather_rows_bf16(const void* src, int32_t ld_src, const int32_t* idx, int32_t rows, int32_t cols,
                                         void* out, int32_t ld_out, void* stream) {
    if (rows < 0 || !bf16_rows_ok(src, ld_src, out, ld_out, cols))
        return set_error(B200VSGG_ERR_BAD_ARG, "gather_rows_bf16: cols / strides % 8 != 0 or pointers not 16-byte aligned");
    if (rows == 0) return 0;
    gather_rows_bf16_kernel<<<grid_for(rows, 8), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)src, ld_src, idx, rows, cols, (__nv_bfloat16*)out, ld_out);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_gather2_sum_rows_bf16(const void* src, int32_t ld_src, const int32_t* idx2, int32_t rows,
                                              int32_t cols, void* out, int32_t ld_out, void* stream) {
    if (!idx2 || rows < 0 || !bf16_rows_ok(src, ld_src, out, ld_out, cols))
        return set_error(B200VSGG_ERR_BAD_ARG, "gather2_sum_rows_bf16: cols / strides % 8 != 0 or pointers not 16-byte aligned");
    if (rows == 0) return 0;
    gather2_sum_rows_bf16_kernel<<<grid_for(rows, 8), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)src, ld_src, idx2, rows, cols, (__nv_bfloat16*)out, ld_out);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_pair_concat_fwd(const float* so, const int64_t* pair_idx, const int64_t* labels,
                                        const float* embed1, const float* embed2, int32_t n_pairs, float* tok_f32,
                                        void* tok_bf16, void* stream) {
    if (!so || !pair_idx || !labels || !embed1 || !embed2 || !tok_f32 || !tok_bf16)
        return set_error(B200VSGG_ERR_BAD_ARG, "pair_concat_fwd: null pointer");
    if (n_pairs == 0) return 0;
    pair_concat_kernel<<<grid_for(n_pairs, 8), 256, 0, (cudaStream_t)stream>>>(so, pair_idx, labels, embed1, embed2,
                                                                              n_pairs, tok_f32, (__nv_bfloat16*)tok_bf16);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_pair_concat_bwd(const float* dtok, const int64_t* pair_idx, const int64_t* labels,
                                        int32_t n_pairs, float* dso, float* dembed1, float* dembed2, void* stream) {
    if (!dtok || !pair_idx || !labels || !dso) return set_error(B200VSGG_ERR_BAD_ARG, "pair_concat_bwd: null pointer");
    if (n_pairs == 0) return 0;
    pair_concat_bwd_kernel<<<grid_for(n_pairs, 8), 256, 0, (cudaStream_t)stream>>>(dtok, pair_idx, labels, n_pairs, dso,
                                                                                  dembed1, dembed2);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_layernorm_fwd(const float* x, int32_t ld_x, const float* gamma, const float* beta, int32_t rows,
                                      int32_t cols, float eps, float* y_f32, int32_t ld_y, void* y_bf16, int32_t ld_b,
                                      const float* add_table, const int32_t* add_idx, void* y_bf16_added,
                                      int32_t ld_added, float* mean, float* rstd, void* stream) {
    if (!x || !gamma || !beta || cols <= 0 || (cols & 7) || cols > 2560)
        return set_error(B200VSGG_ERR_BAD_ARG, "layernorm_fwd: cols must be a multiple of 8 and <= 2560");
    if (rows == 0) return 0;
    auto kern = cols <= 1024 ? layernorm_fwd_kernel<8> : (cols <= 2048 ? layernorm_fwd_kernel<16> : layernorm_fwd_kernel<20>);
    kern<<<grid_for(rows, 4, 148 * 32), 128, 0, (cudaStream_t)stream>>>(
        x, ld_x, gamma, beta, rows, cols, eps, y_f32, ld_y, (__nv_bfloat16*)y_bf16, ld_b, add_table, add_idx,
        (__nv_bfloat16*)y_bf16_added, ld_added, mean, rstd);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_layernorm_bwd_add(const float* dy, int32_t ld_dy, const float* x, int32_t ld_x, const float* gamma,
                                      const float* mean, const float* rstd, int32_t rows, int32_t cols, float* dx_f32,
                                      int32_t ld_dx, void* dx_bf16, int32_t ld_b, float drop_p, uint64_t drop_seed,
                                      float* dgamma, float* dbeta, void* stream, const float* base, int32_t ld_base) {
    if (!dy || !x || !gamma || !mean || !rstd || cols <= 0 || (cols & 7) || (dgamma == nullptr) != (dbeta == nullptr) ||
        (rows + 255) / 256 > 65535)
        return set_error(B200VSGG_ERR_BAD_ARG, "layernorm_bwd: bad arg");
    if (rows == 0) return 0;
    {
        int grid = (rows + 7) / 8;
        if (grid > 148 * 8) grid = 148 * 8;
        layernorm_bwd_dx_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dy, ld_dy, x, ld_x, gamma, mean, rstd, rows, cols,
                                                                      dx_f32, ld_dx, (__nv_bfloat16*)dx_bf16, ld_b, drop_p,
                                                                      drop_seed, base, ld_base);
        if (dgamma != nullptr && dbeta != nullptr) {
            dim3 pgrid((cols + 127) / 128, (rows + 255) / 256);
            layernorm_bwd_param_kernel<<<pgrid, 256, 0, (cudaStream_t)stream>>>(dy, ld_dy, x, ld_x, mean, rstd, rows, cols,
                                                                              dgamma, dbeta);
        }
    }
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_layernorm_bwd(const float* dy, int32_t ld_dy, const float* x, int32_t ld_x, const float* gamma,
                                      const float* mean, const float* rstd, int32_t rows, int32_t cols, float* dx_f32,
                                      int32_t ld_dx, void* dx_bf16, int32_t ld_b, float drop_p, uint64_t drop_seed,
                                      float* dgamma, float* dbeta, void* stream) {
    return b200vsgg_layernorm_bwd_add(dy, ld_dy, x, ld_x, gamma, mean, rstd, rows, cols, dx_f32, ld_dx, dx_bf16, ld_b,
                                      drop_p, drop_seed, dgamma, dbeta, stream, nullptr, 0);
}

extern "C" int b200vsgg_cast_dropout_bf16(const float* x, int32_t ld_x, int32_t rows, int32_t cols, void* out,
                                          int32_t ld_o, float drop_p, uint64_t seed, void* stream) {
    if (!x || !out || cols <= 0 || (cols & 3)) return set_error(B200VSGG_ERR_BAD_ARG, "cast_dropout: cols % 4 != 0");
    if (rows == 0) return 0;
    cast_dropout_kernel<<<grid_for(static_cast<long long>(rows) * (cols / 4), 256), 256, 0, (cudaStream_t)stream>>>(
        x, ld_x, rows, cols, (__nv_bfloat16*)out, ld_o, drop_p, seed);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_split3_bf16(const float* x, int32_t ld_x, int32_t rows, int32_t cols, void* out, int32_t ld_o,
                                    void* stream) {
    if (!x || !out || cols <= 0 || (cols & 3) || ld_o < 3 * cols)
        return set_error(B200VSGG_ERR_BAD_ARG, "split3_bf16: cols % 4 != 0 or ld_o < 3 * cols");
    if (rows == 0) return 0;
    split3_bf16_kernel<<<grid_for(static_cast<long long>(rows) * (cols / 4), 256), 256, 0, (cudaStream_t)stream>>>(
        x, ld_x, rows, cols, (__nv_bfloat16*)out, ld_o);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_colsum(const void* x, int32_t x_is_bf16, int32_t ld_x, int32_t rows, int32_t cols,
                               const int32_t* group_idx, int32_t n_groups, float* out, void* stream) {
    if (!x || !out || n_groups < 1 || n_groups > 2) return set_error(B200VSGG_ERR_BAD_ARG, "colsum: bad arg");
    if (rows == 0) return 0;
    if (x_is_bf16 && (cols & 7) == 0 && (ld_x & 7) == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0) {
        int slabs = (rows + 127) / 128;
        const int col_blocks = (cols + 255) / 256;
        const int want = (148 * 4 + col_blocks - 1) / col_blocks;            // ~4 CTAs per SM in flight
        if (slabs > want) slabs = want;
        colsum_vec_kernel<<<dim3(col_blocks, slabs), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, ld_x, rows, cols,
                                                                                      group_idx, n_groups, out);
        VSGG_CUDA_CHECK_LAUNCH();
        return 0;
    }
    dim3 block(64, 4);
    int slabs = (rows + 255) / 256;
    if (slabs > 64) slabs = 64;
    dim3 grid((cols + 63) / 64, slabs);
    if (x_is_bf16)
        colsum_kernel<__nv_bfloat16><<<grid, block, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, ld_x, rows, cols,
                                                                             group_idx, n_groups, out);
    else
        colsum_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>((const float*)x, ld_x, rows, cols, group_idx,
                                                                       n_groups, out);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}
