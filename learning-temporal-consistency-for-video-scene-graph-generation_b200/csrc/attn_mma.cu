// Variable-length multi-head attention over SHORT segments (one frame's pairs / one 2-frame window,
// <= 32 tokens) with head_dim <= 248 — nn.MultiheadAttention of tools/utils/transformer.py:23,50 on
// exactly the rows the reference keeps.  Segments of <= 16 tokens: one WARP owns one (segment, head) unit; 17..32 tokens:
// one 4-warp CTA per unit (the QUAD kernels at the end of this file).  In both:
//   * Q/K/V (and dO) head slices are staged with 16-byte cp.async of the ALIGNED chunks that cover
//     the head's columns (head_dim 242 puts head h at a 4-byte-aligned offset h*484 B; the chunk
//     grid is kept, the few foreign columns at either end are zeroed in the contraction operands
//     and never stored),
//   * S = QK^T, O = PV (and the five backward products) are warp-level mma.sync m16n8k16 bf16 with
//     fp32 accumulation, fragments via ldmatrix(.trans); softmax, dropout and the dS algebra stay in
//     registers with quad shuffles.
// Why not tcgen05 here: a unit is a 16x16x242 problem (124 kFLOP) and the kernel moves 8*D bytes per
// token for 4*L*D flops (L/2 flop/byte, 4..16 << the ~210 flop/byte ridge of B200): it is HBM-bound,
// the tensor pipe is idle either way, and a 128-row UMMA tile would be >85 % padding.  Roofline for
// this kernel is therefore HBM bytes.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <algorithm>
#include <cstdint>

#include "../../include/b200vsgg.h"
#include "common.cuh"

namespace vsgg {

constexpr int AM_PITCH_B = 528;   // 33 chunks of 16 B: odd chunk count -> conflict-free ldmatrix
constexpr int AM_MAX_CH = 32;     // aligned 16-byte chunks per head row (head_dim <= 248)

__device__ __forceinline__ uint32_t s_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ float am_drop_factor(uint32_t thr, float inv_keep, unsigned long long seed, int row, int head,
                                                int j) {
    if (thr == 0u) return 1.f;
    const uint32_t h = hash_u32(seed, (static_cast<unsigned long long>(row) * 64ull + head) * 4096ull + j);
    return h >= thr ? inv_keep : 0.f;
}
__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

struct HeadGeom {
    int row0, L, col0, c_lo, phase, nch;
};
__device__ __forceinline__ bool head_geom(const int32_t* seg_off, int unit, int n_heads, int hd, HeadGeom& g, int& head) {
    const int seg = unit / n_heads;
    head = unit - seg * n_heads;
    g.row0 = __ldg(seg_off + seg);
    g.L = __ldg(seg_off + seg + 1) - g.row0;
    g.col0 = head * hd;
    g.c_lo = g.col0 >> 3;
    g.phase = g.col0 & 7;
    g.nch = ((g.col0 + hd - 1) >> 3) - g.c_lo + 1;
    return g.L > 0;
}

// Stage rows [row0,row0+L) x aligned chunks [c_lo, c_lo+nch) of `src` into a [LP][33-chunk] tile.
__device__ __forceinline__ void stage_tile(uint8_t* tile, const __nv_bfloat16* src, int ld, const HeadGeom& g, int lane,
                                           int nthreads = 32) {
    // lane = 16-byte chunk of the row (nch <= 32), warps stride over the rows: one cp.async per (thread, row) with
    // add-only address arithmetic (a flat index with a division per chunk cost more than the copy itself)
    const int ch = lane & 31, r0 = lane >> 5, rstep = nthreads >> 5;
    if (ch < g.nch) {
        const __nv_bfloat16* sp = src + static_cast<size_t>(g.row0 + r0) * ld + (g.c_lo + ch) * 8;
        uint32_t dp = s_u32(tile) + r0 * AM_PITCH_B + ch * 16;
        for (int r = r0; r < g.L; r += rstep) {
            cp_async16(dp, sp);
            sp += static_cast<size_t>(rstep) * ld;
            dp += rstep * AM_PITCH_B;
        }
    }
}
// Zero the foreign columns (before `phase`, after phase+hd) and the next chunk of rows < L, so that a
// contraction over whole 32-byte k-steps sees zeros outside the head.
__device__ __forceinline__ void zero_slop(uint8_t* tile, const HeadGeom& g, int hd, int lane) {
    for (int r = lane; r < g.L; r += 32) {
        uint32_t* row = reinterpret_cast<uint32_t*>(tile + r * AM_PITCH_B);   // bf16 pairs
        for (int e = 0; e < g.phase; e += 2) row[e >> 1] = 0u;
        for (int e = g.phase + hd; e < g.nch * 8; e += 2) row[e >> 1] = 0u;
        if (g.nch <= AM_MAX_CH) *reinterpret_cast<uint4*>(tile + r * AM_PITCH_B + g.nch * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
}

// acc[mt][nt] += A[rows of m-tile mt] . B[rows of n-tile nt]^T over the d-chunks (both tiles K-major).
template <int MT, int NT>
__device__ __forceinline__ void qk_product(float (&acc)[MT][NT][4], const uint8_t* A, const uint8_t* B, int ksteps, int lane,
                                           int ks0 = 0) {
    const uint32_t a_base = s_u32(A) + (lane & 15) * AM_PITCH_B + (lane >> 4) * 16;
    const uint32_t b_base = s_u32(B) + ((lane & 7) + ((lane >> 4) & 1) * 8) * AM_PITCH_B + ((lane >> 3) & 1) * 16;
    for (int ks = ks0; ks < ksteps; ++ks) {
        uint32_t a[MT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) ldsm_x4(a[mt], a_base + mt * 16 * AM_PITCH_B + ks * 32);
#pragma unroll
        for (int np = 0; np < NT / 2; ++np) {
            uint32_t b[4];
            ldsm_x4(b, b_base + np * 16 * AM_PITCH_B + ks * 32);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                mma_bf16(acc[mt][2 * np], a[mt], b[0], b[1]);
                mma_bf16(acc[mt][2 * np + 1], a[mt], b[2], b[3]);
            }
        }
    }
}

// out[rows of m-tile mt][d] = sum_k A_frag[mt][kk] * Bt[k][d] for the d-chunk group dg (4 chunks); Bt is a
// [k rows][d] tile read with ldmatrix.trans.  Results are scaled and stored as bf16 pairs for rows < L
// and columns inside the head.
template <int MT, int KK>
__device__ __forceinline__ void av_product_store(const uint32_t (&afrag)[MT][KK][4], const uint8_t* Bt, const HeadGeom& g,
                                                 int hd, float out_scale, __nv_bfloat16* out, int ldo, int lane,
                                                 int dg0 = 0, int dg1 = -1) {
    const uint32_t b_base = s_u32(Bt) + (lane & 15) * AM_PITCH_B + (lane >> 4) * 16;
    const int gq = lane >> 2, tq = lane & 3;
    const int ngroups = dg1 < 0 ? (g.nch + 3) >> 2 : dg1;
    for (int dg = dg0; dg < ngroups; ++dg) {
        float o[MT][4][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) o[mt][j][e] = 0.f;
#pragma unroll
        for (int kk = 0; kk < KK; ++kk) {
#pragma unroll
            for (int dp = 0; dp < 2; ++dp) {
                uint32_t b[4];
                ldsm_x4_t(b, b_base + kk * 16 * AM_PITCH_B + (dg * 4 + dp * 2) * 16);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    mma_bf16(o[mt][dp * 2], afrag[mt][kk], b[0], b[1]);
                    mma_bf16(o[mt][dp * 2 + 1], afrag[mt][kk], b[2], b[3]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gcol = (g.c_lo + dg * 4 + j) * 8 + tq * 2;
            if (gcol < g.col0 || gcol >= g.col0 + hd) continue;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const int r0 = mt * 16 + gq, r1 = r0 + 8;
                if (r0 < g.L)
                    *reinterpret_cast<uint32_t*>(out + static_cast<size_t>(g.row0 + r0) * ldo + gcol) =
                        pack_bf16(o[mt][j][0] * out_scale, o[mt][j][1] * out_scale);
                if (r1 < g.L)
                    *reinterpret_cast<uint32_t*>(out + static_cast<size_t>(g.row0 + r1) * ldo + gcol) =
                        pack_bf16(o[mt][j][2] * out_scale, o[mt][j][3] * out_scale);
            }
        }
    }
}

// In-register softmax over the key axis of s[mt][nt] (C-fragment layout), keys >= L masked.
template <int MT, int NT>
__device__ __forceinline__ void softmax_rows(float (&s)[MT][NT][4], float scale, int L, int lane) {
    const int tq = lane & 3;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float mx = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int col = nt * 8 + tq * 2 + e;
                    const float v = col < L ? s[mt][nt][half * 2 + e] * scale : -INFINITY;
                    s[mt][nt][half * 2 + e] = v;
                    mx = fmaxf(mx, v);
                }
            mx = quad_max(mx);
            float sum = 0.f;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float ex = __expf(s[mt][nt][half * 2 + e] - mx);
                    s[mt][nt][half * 2 + e] = ex;
                    sum += ex;
                }
            sum = quad_sum(sum);
            const float inv = 1.f / sum;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) s[mt][nt][half * 2 + e] *= inv;
        }
    }
}

template <int LP>
__global__ void __launch_bounds__(LP == 16 ? 256 : 128)
attn_mma_fwd_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const __nv_bfloat16* __restrict__ k, int ldk,
                    const __nv_bfloat16* __restrict__ v, int ldv, const int32_t* __restrict__ seg_off, int n_units,
                    int n_heads, int hd, float scale, __nv_bfloat16* __restrict__ ctx, int ldc, float drop_p,
                    unsigned long long seed) {
    constexpr int MT = LP / 16, NT = LP / 8, KK = LP / 16;
    constexpr int TILE = LP * AM_PITCH_B;
    extern __shared__ __align__(16) uint8_t am_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warps = blockDim.x >> 5;
    uint8_t* Qs = am_smem + warp * 3 * TILE;
    uint8_t* Ks = Qs + TILE;
    uint8_t* Vs = Ks + TILE;
    for (int i = lane; i < 3 * TILE / 16; i += 32) reinterpret_cast<uint4*>(Qs)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();
    const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    const uint32_t thr = drop_p > 0.f ? static_cast<uint32_t>(drop_p * 4294967296.0) : 0u;
    const int gq = lane >> 2, tq = lane & 3;
    for (int unit = blockIdx.x * warps + warp; unit < n_units; unit += gridDim.x * warps) {
        HeadGeom g;
        int head;
        if (!head_geom(seg_off, unit, n_heads, hd, g, head)) continue;
        stage_tile(Qs, q, ldq, g, lane);
        stage_tile(Ks, k, ldk, g, lane);
        stage_tile(Vs, v, ldv, g, lane);
        cp_async_wait_all();
        __syncwarp();
        zero_slop(Qs, g, hd, lane);
        __syncwarp();
        float s[MT][NT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) s[mt][nt][e] = 0.f;
        qk_product<MT, NT>(s, Qs, Ks, (g.nch + 1) >> 1, lane);
        softmax_rows<MT, NT>(s, scale, g.L, lane);
        uint32_t pa[MT][KK][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int kk = 0; kk < KK; ++kk) {
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
                    const int nt = 2 * kk + sub;
                    float p[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int row = mt * 16 + gq + (e >> 1) * 8, col = nt * 8 + tq * 2 + (e & 1);
                        p[e] = (row < g.L && col < g.L)
                                   ? s[mt][nt][e] * am_drop_factor(thr, inv_keep, seed, g.row0 + row, head, col) : 0.f;
                    }
                    pa[mt][kk][sub * 2] = pack_bf16(p[0], p[1]);
                    pa[mt][kk][sub * 2 + 1] = pack_bf16(p[2], p[3]);
                }
            }
        av_product_store<MT, KK>(pa, Vs, g, hd, 1.f, ctx, ldc, lane);
        __syncwarp();
    }
}

// Backward: recompute P; dP~ = dO V^T; dS = P o (f*dP~ - rowsum(P o f*dP~)); dQ = scale dS K;
// dK = scale dS^T Q; dV = (P o f)^T dO.
template <int LP>
__global__ void __launch_bounds__(LP == 16 ? 192 : 96)
attn_mma_bwd_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const __nv_bfloat16* __restrict__ k, int ldk,
                    const __nv_bfloat16* __restrict__ v, int ldv, const __nv_bfloat16* __restrict__ dctx, int ldc,
                    const int32_t* __restrict__ seg_off, int n_units, int n_heads, int hd, float scale,
                    __nv_bfloat16* __restrict__ dq, int lddq, __nv_bfloat16* __restrict__ dk, int lddk,
                    __nv_bfloat16* __restrict__ dv, int lddv, float drop_p, unsigned long long seed) {
    constexpr int MT = LP / 16, NT = LP / 8, KK = LP / 16;
    constexpr int TILE = LP * AM_PITCH_B;
    constexpr int TP = (LP + 8) * 2;            // pitch of the small [LP][LP] bf16 tiles (odd # of 16-B chunks)
    constexpr int SMALL = LP * TP;
    constexpr int PER_WARP = 4 * TILE + 2 * SMALL;
    extern __shared__ __align__(16) uint8_t am_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warps = blockDim.x >> 5;
    uint8_t* Qs = am_smem + warp * PER_WARP;
    uint8_t* Ks = Qs + TILE;
    uint8_t* Vs = Ks + TILE;
    uint8_t* Os = Vs + TILE;
    uint8_t* Tds = Os + TILE;
    uint8_t* Tp = Tds + SMALL;
    for (int i = lane; i < PER_WARP / 16; i += 32) reinterpret_cast<uint4*>(Qs)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();
    const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    const uint32_t thr = drop_p > 0.f ? static_cast<uint32_t>(drop_p * 4294967296.0) : 0u;
    const int gq = lane >> 2, tq = lane & 3;
    for (int unit = blockIdx.x * warps + warp; unit < n_units; unit += gridDim.x * warps) {
        HeadGeom g;
        int head;
        if (!head_geom(seg_off, unit, n_heads, hd, g, head)) continue;
        stage_tile(Qs, q, ldq, g, lane);
        stage_tile(Ks, k, ldk, g, lane);
        stage_tile(Vs, v, ldv, g, lane);
        stage_tile(Os, dctx, ldc, g, lane);
        cp_async_wait_all();
        __syncwarp();
        zero_slop(Qs, g, hd, lane);
        zero_slop(Os, g, hd, lane);
        __syncwarp();
        const int ksteps = (g.nch + 1) >> 1;
        float p[MT][NT][4], dp[MT][NT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) { p[mt][nt][e] = 0.f; dp[mt][nt][e] = 0.f; }
        qk_product<MT, NT>(p, Qs, Ks, ksteps, lane);
        softmax_rows<MT, NT>(p, scale, g.L, lane);
        qk_product<MT, NT>(dp, Os, Vs, ksteps, lane);
        // dS and P~ (C layout) -> A fragments for dQ, and bf16 tiles for the transposed products
        uint32_t dsa[MT][KK][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            float delta[2] = {0.f, 0.f};
            const int r0 = mt * 16 + gq;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                float pt[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int row = r0 + (e >> 1) * 8, col = nt * 8 + tq * 2 + (e & 1);
                    const bool ok = row < g.L && col < g.L;
                    const float f = ok ? am_drop_factor(thr, inv_keep, seed, g.row0 + row, head, col) : 0.f;
                    const float pe = ok ? p[mt][nt][e] : 0.f;
                    const float dpe = ok ? dp[mt][nt][e] * f : 0.f;
                    delta[e >> 1] += pe * dpe;
                    p[mt][nt][e] = pe;
                    dp[mt][nt][e] = dpe;
                    pt[e] = pe * f;
                }
                const int c = nt * 8 + tq * 2;
                *reinterpret_cast<uint32_t*>(Tp + r0 * TP + c * 2) = pack_bf16(pt[0], pt[1]);
                *reinterpret_cast<uint32_t*>(Tp + (r0 + 8) * TP + c * 2) = pack_bf16(pt[2], pt[3]);
            }
            delta[0] = quad_sum(delta[0]);
            delta[1] = quad_sum(delta[1]);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                float ds[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) ds[e] = p[mt][nt][e] * (dp[mt][nt][e] - delta[e >> 1]);
                const uint32_t d01 = pack_bf16(ds[0], ds[1]), d23 = pack_bf16(ds[2], ds[3]);
                dsa[mt][nt >> 1][(nt & 1) * 2] = d01;
                dsa[mt][nt >> 1][(nt & 1) * 2 + 1] = d23;
                const int c = nt * 8 + tq * 2;
                *reinterpret_cast<uint32_t*>(Tds + r0 * TP + c * 2) = d01;
                *reinterpret_cast<uint32_t*>(Tds + (r0 + 8) * TP + c * 2) = d23;
            }
        }
        __syncwarp();
        av_product_store<MT, KK>(dsa, Ks, g, hd, scale, dq, lddq, lane);          // dQ = scale * dS K
        // transposed A fragments: rows = keys, k = queries
        uint32_t ta[MT][KK][4];
        const uint32_t t_off = ((lane & 7) + ((lane >> 4) & 1) * 8) * TP + ((lane >> 3) & 1) * 16;
#pragma unroll
        for (int jt = 0; jt < MT; ++jt)
#pragma unroll
            for (int it = 0; it < KK; ++it) ldsm_x4_t(ta[jt][it], s_u32(Tds) + t_off + it * 16 * TP + jt * 32);
        av_product_store<MT, KK>(ta, Qs, g, hd, scale, dk, lddk, lane);           // dK = scale * dS^T Q
#pragma unroll
        for (int jt = 0; jt < MT; ++jt)
#pragma unroll
            for (int it = 0; it < KK; ++it) ldsm_x4_t(ta[jt][it], s_u32(Tp) + t_off + it * 16 * TP + jt * 32);
        av_product_store<MT, KK>(ta, Os, g, hd, 1.f, dv, lddv, lane);             // dV = P~^T dO
        __syncwarp();
    }
}


// ------------------------------------------------------------------------------------------------
// QUAD kernels for 17..32-token segments (the 2-frame temporal windows): one 128-thread CTA per (segment, head) unit,
// several CTAs per SM.  Their predecessors (one warp, then a warp pair per unit with 32-row tiles) ran 4..6 warps per SM
// and spent ~14 us per unit on a chain of dependent steps (ncu: 5 % active warps, 10 % of DRAM bandwidth; 0.29 / 0.73 ms
// per layer forward / backward at the headline shape).  Here (0.12 / 0.27 ms, profiles/r02_ncu_window_attention_*.txt)
//   * operand tiles have only as many rows as the longest segment of the launch (rounded to 8; ldmatrix row
//     addresses are clamped), so 3 (backward) / 5 (forward) CTAs = 12 / 20 warps fit per SM;
//   * S = Q K^T (and dP = dO V^T) are split by OUTPUT block: warp w owns the 16x16 block (w>>1, w&1) over the whole
//     head dimension — no partial sums to exchange, deterministic;
//   * softmax / dropout / dS run once per row (warp = row, lane = key) on the fp32 scores in shared memory, instead
//     of redundantly in every warp's fragment layout;
//   * the output products are split by head-dimension column group (warp w: groups w and w+4), their results are
//     written IN PLACE over the B operand they were computed from (dQ over K, dK over Q, dV over dO, O over V: a
//     warp only ever touches its own columns), and rows leave the CTA as 16-byte stores (lane = aligned chunk;
//     the two chunks a head shares with its neighbours use masked 4-byte stores).
// ------------------------------------------------------------------------------------------------
constexpr int AW_SP = 36;   // pitch (floats) of the fp32 score tiles
constexpr int AW_TP = 80;   // pitch (bytes) of the bf16 [32][32] probability tiles: 5 chunks (odd) -> conflict-free ldmatrix

__device__ __forceinline__ void st_shared_u32(uint32_t dst, uint32_t v) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst), "r"(v) : "memory");
}
// Stage one operand tile (lane = chunk, warps stride over the rows: at most 8 rows per warp).  One 64-bit address per
// operand, then a running pointer and predicated copies: no branches, no per-row multiplications.
__device__ __forceinline__ void aw_stage(uint8_t* tile, const __nv_bfloat16* src, int ld, const HeadGeom& g, int warp, int lane) {
    const char* sp = reinterpret_cast<const char*>(src + static_cast<size_t>(g.row0 + warp) * ld + (g.c_lo + lane) * 8);
    const size_t step = static_cast<size_t>(ld) * 8;             // four rows, in bytes
    const uint32_t dp = s_u32(tile) + warp * AM_PITCH_B + lane * 16;
    const int n = lane < g.nch ? g.L - warp : 0;                 // row warp + 4t is staged  <=>  4t < n
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        if (4 * t < n) cp_async16(dp + t * 4 * AM_PITCH_B, sp);
        sp += step;
    }
}
// Zero the foreign columns of the chunks THIS thread staged (first / last chunk of its rows) and the pad chunk: word masks,
// no loops over elements.
__device__ __forceinline__ void aw_fix_slop(uint8_t* tile, const HeadGeom& g, int hd, int warp, int lane) {
    const int last = g.nch - 1;
    const bool is_last = lane == last;
    if (!(is_last || (lane == 0 && g.phase > 0))) return;
    const int end = g.phase + hd - last * 8;          // valid elements of the last chunk (2..8, even)
    uint32_t m[4];
#pragma unroll
    for (int w = 0; w < 4; ++w)
        m[w] = ((lane != 0 || 2 * w >= g.phase) && (!is_last || 2 * w < end)) ? 0xffffffffu : 0u;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int r = warp + 4 * t;
        if (r < g.L) {
            uint4* p = reinterpret_cast<uint4*>(tile + r * AM_PITCH_B + lane * 16);
            uint4 v = *p;
            v.x &= m[0]; v.y &= m[1]; v.z &= m[2]; v.w &= m[3];
            *p = v;
            if (is_last) p[1] = make_uint4(0u, 0u, 0u, 0u);
        }
    }
}
// Dropout factors of 4 consecutive keys of one (query row, head) from ONE 64-bit hash (16 bits per key): the quad kernels'
// own stream — forward and backward of this kernel family regenerate the same mask, nothing else needs to agree with it.
__device__ __forceinline__ void aw_drop4(float (&f)[4], uint32_t thr16, float inv_keep, unsigned long long seed, int row,
                                         int head, int key0) {
    const unsigned long long z = hash_u64(seed, (static_cast<unsigned long long>(row) * 64ull + head) * 16ull + (key0 >> 2));
    const uint32_t lo = static_cast<uint32_t>(z), hi = static_cast<uint32_t>(z >> 32);
    f[0] = (lo & 0xffffu) >= thr16 ? inv_keep : 0.f;
    f[1] = (lo >> 16) >= thr16 ? inv_keep : 0.f;
    f[2] = (hi & 0xffffu) >= thr16 ? inv_keep : 0.f;
    f[3] = (hi >> 16) >= thr16 ? inv_keep : 0.f;
}
// Warp w computes the 16x16 block (mt = w>>1, key half nh = w&1) of A . B^T over all k-steps and writes it to the
// fp32 tile S (rows = A rows, columns = B rows).  Row addresses are clamped to the allocated tile rows.
__device__ __forceinline__ void aw_block_product(float* S, const uint8_t* A, const uint8_t* B, int ksteps, int L, int rows_alloc,
                                                 int warp, int lane) {
    const int mt = warp >> 1, nh = warp & 1;
    if (mt * 16 >= L || nh * 16 >= L) return;
    const int ar = min(mt * 16 + (lane & 15), rows_alloc - 1);
    const int br = min(nh * 16 + (lane & 7) + ((lane >> 4) & 1) * 8, rows_alloc - 1);
    const uint32_t a_base = s_u32(A) + ar * AM_PITCH_B + (lane >> 4) * 16;
    const uint32_t b_base = s_u32(B) + br * AM_PITCH_B + ((lane >> 3) & 1) * 16;
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll 4
    for (int ks = 0; ks < ksteps; ++ks) {
        uint32_t a[4], b[4];
        ldsm_x4(a, a_base + ks * 32);
        ldsm_x4(b, b_base + ks * 32);
        mma_bf16(acc[0], a, b[0], b[1]);
        mma_bf16(acc[1], a, b[2], b[3]);
    }
    const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        float* d = S + (mt * 16 + gq) * AW_SP + (nh * 2 + t) * 8 + tq * 2;
        *reinterpret_cast<float2*>(d) = make_float2(acc[t][0], acc[t][1]);
        *reinterpret_cast<float2*>(d + 8 * AW_SP) = make_float2(acc[t][2], acc[t][3]);
    }
}
// A fragments of a [32][32] bf16 tile T (pitch AW_TP): plain (rows of T are the output rows) or transposed.
__device__ __forceinline__ void aw_load_a(uint32_t (&a)[2][2][4], const uint8_t* T, bool transposed, int lane) {
    if (!transposed) {
        const uint32_t base = s_u32(T) + (lane & 15) * AW_TP + (lane >> 4) * 16;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) ldsm_x4(a[mt][kk], base + mt * 16 * AW_TP + kk * 32);
    } else {
        const uint32_t base = s_u32(T) + ((lane & 7) + ((lane >> 4) & 1) * 8) * AW_TP + ((lane >> 3) & 1) * 16;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) ldsm_x4_t(a[mt][kk], base + kk * 16 * AW_TP + mt * 32);
    }
}
// out[:, columns of group dg] = A . Bt for this warp's column groups, written over Bt's own columns.
__device__ __forceinline__ void aw_out_product_inplace(const uint32_t (&a)[2][2][4], uint8_t* Bt, int ngroups, int L, int rows_alloc,
                                                       int warp, int lane) {
    const int gq = lane >> 2, tq = lane & 3;
    const int k0 = min(lane & 15, rows_alloc - 1), k1 = min(16 + (lane & 15), rows_alloc - 1);
    const uint32_t b0 = s_u32(Bt) + k0 * AM_PITCH_B + (lane >> 4) * 16;
    const uint32_t b1 = s_u32(Bt) + k1 * AM_PITCH_B + (lane >> 4) * 16;
    const bool two_m = L > 16, two_k = L > 16;
    const uint32_t d0 = s_u32(Bt) + gq * AM_PITCH_B + tq * 4;
    const bool v0 = gq < L, v1 = gq + 8 < L, v2 = gq + 16 < L, v3 = gq + 24 < L;
    for (int dg = warp; dg < ngroups; dg += 4) {
        float o[2][4][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) o[mt][j][e] = 0.f;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
            if (kk == 1 && !two_k) break;
#pragma unroll
            for (int dp = 0; dp < 2; ++dp) {
                uint32_t b[4];
                ldsm_x4_t(b, (kk ? b1 : b0) + (dg * 4 + dp * 2) * 16);
                mma_bf16(o[0][dp * 2], a[0][kk], b[0], b[1]);
                mma_bf16(o[0][dp * 2 + 1], a[0][kk], b[2], b[3]);
                if (two_m) {
                    mma_bf16(o[1][dp * 2], a[1][kk], b[0], b[1]);
                    mma_bf16(o[1][dp * 2 + 1], a[1][kk], b[2], b[3]);
                }
            }
        }
        __syncwarp();      // every lane's ldmatrix of these columns has executed before they are overwritten
        const uint32_t d = d0 + dg * 64;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (v0) st_shared_u32(d + j * 16, pack_bf16(o[0][j][0], o[0][j][1]));
            if (v1) st_shared_u32(d + j * 16 + 8 * AM_PITCH_B, pack_bf16(o[0][j][2], o[0][j][3]));
            if (two_m) {
                if (v2) st_shared_u32(d + j * 16 + 16 * AM_PITCH_B, pack_bf16(o[1][j][0], o[1][j][1]));
                if (v3) st_shared_u32(d + j * 16 + 24 * AM_PITCH_B, pack_bf16(o[1][j][2], o[1][j][3]));
            }
        }
    }
}
// Rows [0,L) of a tile -> global rows, columns of the head only: 16-byte stores for chunks inside the head.
__device__ __forceinline__ void aw_store_tile(const uint8_t* tile, __nv_bfloat16* out, int ldo, const HeadGeom& g, int hd, bool vec,
                                              int warp, int lane) {
    if (lane >= g.nch) return;
    const int gc = (g.c_lo + lane) * 8;                            // first global column of this lane's chunk
    const int lo = max(g.col0, gc) - gc, hi = min(g.col0 + hd, gc + 8) - gc;
    const bool whole = vec && lo == 0 && hi == 8;
    for (int r = warp; r < g.L; r += 4) {
        const uint4 v = *reinterpret_cast<const uint4*>(tile + r * AM_PITCH_B + lane * 16);
        __nv_bfloat16* dst = out + static_cast<size_t>(g.row0 + r) * ldo + gc;
        if (whole) {
            *reinterpret_cast<uint4*>(dst) = v;
        } else {
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (2 * e >= lo && 2 * e < hi) *reinterpret_cast<uint32_t*>(dst + 2 * e) = w[e];
        }
    }
}

// Softmax of row i of the fp32 score tile for the 8 keys [c0, c0+8) of this thread (4 threads share a row): e[] holds the
// probabilities, exactly 0 for keys / rows outside the segment.  Branch-free: invalid rows reduce over -inf safely.
__device__ __forceinline__ void aw_softmax_row(float (&e)[8], const float* S, int i, int c0, int L, float scale) {
    const float4 s0 = *reinterpret_cast<const float4*>(S + i * AW_SP + c0);
    const float4 s1 = *reinterpret_cast<const float4*>(S + i * AW_SP + c0 + 4);
    const float sr[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    const bool row_ok = i < L;
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        e[t] = (row_ok && c0 + t < L) ? sr[t] * scale : -INFINITY;
        mx = fmaxf(mx, e[t]);
    }
    mx = quad_max(mx);
    if (!row_ok) mx = 0.f;
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        e[t] = __expf(e[t] - mx);          // exp(-inf) = 0 for the masked keys
        sum += e[t];
    }
    sum = quad_sum(sum);
    const float inv = row_ok ? 1.f / sum : 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) e[t] *= inv;
}

__global__ void __launch_bounds__(128, 5)
attn_win_fwd_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const __nv_bfloat16* __restrict__ k, int ldk,
                    const __nv_bfloat16* __restrict__ v, int ldv, const int32_t* __restrict__ seg_off, int n_units,
                    int n_heads, int hd, float scale, __nv_bfloat16* __restrict__ ctx, int ldc, float drop_p,
                    unsigned long long seed, int rows_alloc, int vec) {
    extern __shared__ __align__(16) uint8_t am_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int TILE = rows_alloc * AM_PITCH_B;
    uint8_t* Qs = am_smem;
    uint8_t* Ks = Qs + TILE;
    uint8_t* Vs = Ks + TILE;
    float* S = reinterpret_cast<float*>(Vs + TILE);
    uint8_t* Tp = reinterpret_cast<uint8_t*>(S + 32 * AW_SP);
    for (int i = threadIdx.x; i < 3 * TILE / 16; i += 128) reinterpret_cast<uint4*>(Qs)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    // 16-bit keep threshold; the scale is the reciprocal of the keep probability actually realised
    const uint32_t thr16 = drop_p > 0.f ? min(65535u, static_cast<uint32_t>(drop_p * 65536.f + 0.5f)) : 0u;
    const float inv_keep = 65536.f / static_cast<float>(65536u - thr16);
    int nseg = blockIdx.x / n_heads;
    int nr0 = blockIdx.x < n_units ? __ldg(seg_off + nseg) : 0, nr1 = blockIdx.x < n_units ? __ldg(seg_off + nseg + 1) : 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        // this unit's segment bounds were fetched one iteration ago; fetch the next unit's now (their latency was
        // fully exposed at the top of every unit)
        HeadGeom g;
        const int seg = nseg, head = unit - seg * n_heads;
        g.row0 = nr0;
        g.L = nr1 - nr0;
        {
            const int nu = unit + gridDim.x;
            nseg = nu / n_heads;
            if (nu < n_units) {
                nr0 = __ldg(seg_off + nseg);
                nr1 = __ldg(seg_off + nseg + 1);
            }
        }
        if (g.L <= 0) continue;                                             // uniform in the CTA
        g.col0 = head * hd;
        g.c_lo = g.col0 >> 3;
        g.phase = g.col0 & 7;
        g.nch = ((g.col0 + hd - 1) >> 3) - g.c_lo + 1;
        aw_stage(Qs, q, ldq, g, warp, lane);
        aw_stage(Ks, k, ldk, g, warp, lane);
        aw_stage(Vs, v, ldv, g, warp, lane);
        cp_async_wait_all();
        aw_fix_slop(Qs, g, hd, warp, lane);
        __syncthreads();
        aw_block_product(S, Qs, Ks, (g.nch + 1) >> 1, g.L, rows_alloc, warp, lane);
        __syncthreads();
        // softmax + dropout: 4 threads per row (8 keys each), all 32 rows in one pass, quad shuffles
        {
            const int i = threadIdx.x >> 2, c0 = (threadIdx.x & 3) * 8;
            float e[8];
            aw_softmax_row(e, S, i, c0, g.L, scale);
            if (thr16 && i < g.L) {
#pragma unroll
                for (int h4 = 0; h4 < 2; ++h4) {
                    float f[4];
                    aw_drop4(f, thr16, inv_keep, seed, g.row0 + i, head, c0 + 4 * h4);
#pragma unroll
                    for (int t = 0; t < 4; ++t) e[4 * h4 + t] *= f[t];
                }
            }
            uint32_t pk[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) pk[t] = pack_bf16(e[2 * t], e[2 * t + 1]);
            *reinterpret_cast<uint4*>(Tp + i * AW_TP + c0 * 2) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
        __syncthreads();
        uint32_t pa[2][2][4];
        aw_load_a(pa, Tp, false, lane);
        aw_out_product_inplace(pa, Vs, (g.nch + 3) >> 2, g.L, rows_alloc, warp, lane);
        __syncthreads();
        aw_store_tile(Vs, ctx, ldc, g, hd, vec != 0, warp, lane);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(128, 4)
attn_win_bwd_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const __nv_bfloat16* __restrict__ k, int ldk,
                    const __nv_bfloat16* __restrict__ v, int ldv, const __nv_bfloat16* __restrict__ dctx, int ldc,
                    const int32_t* __restrict__ seg_off, int n_units, int n_heads, int hd, float scale,
                    __nv_bfloat16* __restrict__ dq, int lddq, __nv_bfloat16* __restrict__ dk, int lddk,
                    __nv_bfloat16* __restrict__ dv, int lddv, float drop_p, unsigned long long seed, int rows_alloc, int vec) {
    extern __shared__ __align__(16) uint8_t am_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int TILE = rows_alloc * AM_PITCH_B;
    uint8_t* Qs = am_smem;
    uint8_t* Ks = Qs + TILE;
    uint8_t* Vs = Ks + TILE;
    uint8_t* Os = Vs + TILE;
    float* S = reinterpret_cast<float*>(Os + TILE);
    float* D = S + 32 * AW_SP;
    uint8_t* Tp = reinterpret_cast<uint8_t*>(D + 32 * AW_SP);
    uint8_t* Tds = Tp + 32 * AW_TP;
    for (int i = threadIdx.x; i < 4 * TILE / 16; i += 128) reinterpret_cast<uint4*>(Qs)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    // 16-bit keep threshold; the scale is the reciprocal of the keep probability actually realised
    const uint32_t thr16 = drop_p > 0.f ? min(65535u, static_cast<uint32_t>(drop_p * 65536.f + 0.5f)) : 0u;
    const float inv_keep = 65536.f / static_cast<float>(65536u - thr16);
    int nseg = blockIdx.x / n_heads;
    int nr0 = blockIdx.x < n_units ? __ldg(seg_off + nseg) : 0, nr1 = blockIdx.x < n_units ? __ldg(seg_off + nseg + 1) : 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        // this unit's segment bounds were fetched one iteration ago; fetch the next unit's now (their latency was
        // fully exposed at the top of every unit)
        HeadGeom g;
        const int seg = nseg, head = unit - seg * n_heads;
        g.row0 = nr0;
        g.L = nr1 - nr0;
        {
            const int nu = unit + gridDim.x;
            nseg = nu / n_heads;
            if (nu < n_units) {
                nr0 = __ldg(seg_off + nseg);
                nr1 = __ldg(seg_off + nseg + 1);
            }
        }
        if (g.L <= 0) continue;                                             // uniform in the CTA
        g.col0 = head * hd;
        g.c_lo = g.col0 >> 3;
        g.phase = g.col0 & 7;
        g.nch = ((g.col0 + hd - 1) >> 3) - g.c_lo + 1;
        aw_stage(Qs, q, ldq, g, warp, lane);
        aw_stage(Ks, k, ldk, g, warp, lane);
        aw_stage(Vs, v, ldv, g, warp, lane);
        aw_stage(Os, dctx, ldc, g, warp, lane);
        cp_async_wait_all();
        aw_fix_slop(Qs, g, hd, warp, lane);
        aw_fix_slop(Os, g, hd, warp, lane);
        __syncthreads();
        const int ksteps = (g.nch + 1) >> 1;
        aw_block_product(S, Qs, Ks, ksteps, g.L, rows_alloc, warp, lane);       // S  = Q K^T
        aw_block_product(D, Os, Vs, ksteps, g.L, rows_alloc, warp, lane);       // dP = dO V^T
        __syncthreads();
        // P, dropout, dS = P o (f dP - rowsum(P o f dP)): 4 threads per row (8 keys each), all rows in one pass
        {
            const int i = threadIdx.x >> 2, c0 = (threadIdx.x & 3) * 8;
            float e[8], dpe[8];
            aw_softmax_row(e, S, i, c0, g.L, scale);
            const float4 d0 = *reinterpret_cast<const float4*>(D + i * AW_SP + c0);
            const float4 d1 = *reinterpret_cast<const float4*>(D + i * AW_SP + c0 + 4);
            const float dr[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
            float delta = 0.f;
            uint32_t pk[4], dk_[4];
            float pt[8], fk[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) fk[t] = 1.f;
            if (thr16 && i < g.L) {
                float f[4];
                aw_drop4(f, thr16, inv_keep, seed, g.row0 + i, head, c0);
#pragma unroll
                for (int t = 0; t < 4; ++t) fk[t] = f[t];
                aw_drop4(f, thr16, inv_keep, seed, g.row0 + i, head, c0 + 4);
#pragma unroll
                for (int t = 0; t < 4; ++t) fk[4 + t] = f[t];
            }
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const bool ok = i < g.L && c0 + t < g.L;
                dpe[t] = ok ? dr[t] * fk[t] : 0.f;      // e[t] is already 0 outside the segment
                pt[t] = e[t] * fk[t];
                delta += e[t] * dpe[t];
            }
            delta = quad_sum(delta);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                pk[t] = pack_bf16(pt[2 * t], pt[2 * t + 1]);
                // the softmax scale of dQ / dK is folded into dS here
                dk_[t] = pack_bf16(e[2 * t] * scale * (dpe[2 * t] - delta), e[2 * t + 1] * scale * (dpe[2 * t + 1] - delta));
            }
            *reinterpret_cast<uint4*>(Tp + i * AW_TP + c0 * 2) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(Tds + i * AW_TP + c0 * 2) = make_uint4(dk_[0], dk_[1], dk_[2], dk_[3]);
        }
        __syncthreads();
        const int ngroups = (g.nch + 3) >> 2;
        uint32_t a[2][2][4];
        aw_load_a(a, Tds, false, lane);
        aw_out_product_inplace(a, Ks, ngroups, g.L, rows_alloc, warp, lane);      // dQ = (scale dS) K     (over K)
        aw_load_a(a, Tds, true, lane);
        aw_out_product_inplace(a, Qs, ngroups, g.L, rows_alloc, warp, lane);      // dK = (scale dS)^T Q   (over Q)
        aw_load_a(a, Tp, true, lane);
        aw_out_product_inplace(a, Os, ngroups, g.L, rows_alloc, warp, lane);      // dV = P~^T dO          (over dO)
        __syncthreads();
        aw_store_tile(Ks, dq, lddq, g, hd, vec != 0, warp, lane);
        aw_store_tile(Qs, dk, lddk, g, hd, vec != 0, warp, lane);
        aw_store_tile(Os, dv, lddv, g, hd, vec != 0, warp, lane);
        __syncthreads();
    }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename Kern>
static int ensure_smem(Kern kern, size_t smem, size_t& cur) {
    if (smem > cur) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
        cur = smem;
    }
    return 0;
}

static int am_num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// Returns 1 if the mma path took the call, 0 if the shapes are outside its envelope, < 0 on error.
int attn_mma_fwd_try(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const int32_t* seg_off,
                     int n_seg, int max_len, int n_heads, int hd, float scale, void* ctx, int ldc, float drop_p,
                     unsigned long long seed, cudaStream_t stream) {
    if (max_len > 32 || hd > 248 || (hd & 1) || n_heads > 64) return 0;
    {
        const int span = (n_heads * hd + 7) / 8 * 8;   // aligned chunks never leave the row
        if (span > ldq || span > ldk || span > ldv) return 0;
    }
    if (!aligned16(q) || !aligned16(k) || !aligned16(v) || (ldq & 7) || (ldk & 7) || (ldv & 7) || (ldc & 1) ||
        (reinterpret_cast<uintptr_t>(ctx) & 3u))
        return 0;
    const int n_units = n_seg * n_heads;
    int rc;
    if (max_len <= 16) {
        constexpr int W = 8;
        const size_t smem = (size_t)W * 3 * 16 * AM_PITCH_B;
        static size_t cur = 0;
        if ((rc = ensure_smem(attn_mma_fwd_kernel<16>, smem, cur))) return rc;
        int grid = (n_units + W - 1) / W;
        if (grid > am_num_sms()) grid = am_num_sms();
        attn_mma_fwd_kernel<16><<<grid, W * 32, smem, stream>>>(
            (const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, ldk, (const __nv_bfloat16*)v, ldv, seg_off, n_units,
            n_heads, hd, scale, (__nv_bfloat16*)ctx, ldc, drop_p, seed);
    } else {
        const int rows = (max_len + 7) & ~7;
        const size_t smem = (size_t)3 * rows * AM_PITCH_B + 32 * AW_SP * 4 + 32 * AW_TP;
        static size_t cur = 0;
        if ((rc = ensure_smem(attn_win_fwd_kernel, smem, cur))) return rc;
        const int per_sm = (int)std::min<size_t>(8, (227 * 1024) / (smem + 1024));
        int grid = std::min(n_units, am_num_sms() * per_sm);
        const int vec = (aligned16(ctx) && (ldc & 7) == 0) ? 1 : 0;
        attn_win_fwd_kernel<<<grid, 128, smem, stream>>>(
            (const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, ldk, (const __nv_bfloat16*)v, ldv, seg_off, n_units,
            n_heads, hd, scale, (__nv_bfloat16*)ctx, ldc, drop_p, seed, rows, vec);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
    return 1;
}

int attn_mma_bwd_try(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const void* dctx, int ldc,
                     const int32_t* seg_off, int n_seg, int max_len, int n_heads, int hd, float scale, void* dq, int lddq,
                     void* dk, int lddk, void* dv, int lddv, float drop_p, unsigned long long seed, cudaStream_t stream) {
    if (max_len > 32 || hd > 248 || (hd & 1) || n_heads > 64) return 0;
    {
        const int span = (n_heads * hd + 7) / 8 * 8;   // aligned chunks never leave the row
        if (span > ldq || span > ldk || span > ldv) return 0;
    }
    if (!aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(dctx) || (ldq & 7) || (ldk & 7) || (ldv & 7) ||
        (ldc & 7) || (lddq & 1) || (lddk & 1) || (lddv & 1) || (reinterpret_cast<uintptr_t>(dq) & 3u) ||
        (reinterpret_cast<uintptr_t>(dk) & 3u) || (reinterpret_cast<uintptr_t>(dv) & 3u))
        return 0;
    const int n_units = n_seg * n_heads;
    int rc;
    if (max_len <= 16) {
        constexpr int W = 6, LP = 16;
        const size_t smem = (size_t)W * (4 * LP * AM_PITCH_B + 2 * LP * (LP + 8) * 2);
        static size_t cur = 0;
        if ((rc = ensure_smem(attn_mma_bwd_kernel<16>, smem, cur))) return rc;
        int grid = (n_units + W - 1) / W;
        if (grid > am_num_sms()) grid = am_num_sms();
        attn_mma_bwd_kernel<16><<<grid, W * 32, smem, stream>>>(
            (const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, ldk, (const __nv_bfloat16*)v, ldv,
            (const __nv_bfloat16*)dctx, ldc, seg_off, n_units, n_heads, hd, scale, (__nv_bfloat16*)dq, lddq,
            (__nv_bfloat16*)dk, lddk, (__nv_bfloat16*)dv, lddv, drop_p, seed);
    } else {
        const int rows = (max_len + 7) & ~7;
        const size_t smem = (size_t)4 * rows * AM_PITCH_B + 2 * 32 * AW_SP * 4 + 2 * 32 * AW_TP;
        static size_t cur = 0;
        if ((rc = ensure_smem(attn_win_bwd_kernel, smem, cur))) return rc;
        const int per_sm = (int)std::min<size_t>(8, (227 * 1024) / (smem + 1024));
        int grid = std::min(n_units, am_num_sms() * per_sm);
        const int vec = (aligned16(dq) && aligned16(dk) && aligned16(dv) && ((lddq | lddk | lddv) & 7) == 0) ? 1 : 0;
        attn_win_bwd_kernel<<<grid, 128, smem, stream>>>(
            (const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, ldk, (const __nv_bfloat16*)v, ldv,
            (const __nv_bfloat16*)dctx, ldc, seg_off, n_units, n_heads, hd, scale, (__nv_bfloat16*)dq, lddq,
            (__nv_bfloat16*)dk, lddk, (__nv_bfloat16*)dv, lddv, drop_p, seed, rows, vec);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
    return 1;
}

}  // namespace vsgg
