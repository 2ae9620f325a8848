// Variable-length flash attention for the TokenGT encoder (tools/TokenGT/tokengt/modules/
// multihead_attention.py:135-183 of the reference): sequences = 5-frame clips (T = 2 + nodes + edges,
// ~10^2..10^3.5 tokens), 32 heads x 24 or 16 heads x 48, q scaled by head_dim^-0.5, fp32 softmax,
// attention dropout.  The reference materialises [heads, T, T] maps for all 12 layers; here nothing
// of size T^2 ever reaches HBM.
//
// CTA = 4 warps; warp w owns 16 query rows of a 64-row query block; key/value blocks of 64 rows are
// staged in shared memory with 16-byte cp.async (head slices are 48 B / 96 B: 16-byte aligned), S and
// PV run on mma.sync m16n8k16 bf16 (fp32 accumulate) with ldmatrix fragments, online softmax in
// registers.  head_dim is padded to a multiple of 16 in shared memory only.
// Backward = two kernels without atomics: dQ per query block (loop over key blocks) and dK/dV per key
// block (loop over query blocks, computing S^T = K Q^T directly so no transposes are needed).
// Roofline: per (clip, head) 4*T^2*hd flops over 8*T*hd bytes => T/2 flop/byte: HBM-bound for the
// C3 shapes (T ~ 450), mixed for C5 (T ~ 5k); tensor cores are used through mma.sync.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>

#include "../../include/b200vsgg.h"
#include "common.cuh"
#include "attn_dropout.cuh"

namespace vsgg {
namespace fa {

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
__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}
// Dropout keep-factor of probability (query row `row` (global), head, key `key_rel` relative to the sequence start):
// the shared counter-based mask of attn_dropout.cuh in its generic per-element form (these mma.sync kernels are the
// reference implementation / fallback for head_dim > 64; the tcgen05 kernels evaluate the same bits four at a time).
__device__ __forceinline__ uint32_t drop_row_key(uint32_t thr, unsigned long long seed, int row, int head) {
    return thr == 0u ? 0u : adrop::row_key(seed, row, head);
}
__device__ __forceinline__ float drop_factor(uint32_t thr, float inv_keep, uint32_t row_key, int key_rel) {
    if (thr == 0u) return 1.f;
    return adrop::keep(thr, row_key, key_rel) ? inv_keep : 0.f;
}

constexpr int BLK = 64;       // rows per query / key block
constexpr int WARPS = 4;

template <int HDP>
struct Geo {
    static constexpr int PITCH = (HDP + 8) * 2;          // bytes; (HDP+8)/8 chunks is odd for HDP = 32, 48, 64
    static constexpr int TILE = BLK * PITCH;
    static constexpr int KS = HDP / 16;                  // k-steps over head_dim
    static constexpr int NT_D = HDP / 8;                 // n-tiles over head_dim
};

// Stage rows [row0, row0+rows) x head slice of `src` into a [64][PITCH] tile; rows beyond `rows` and the
// padding columns [hd, HDP) are zero-filled.
template <int HDP>
__device__ __forceinline__ void stage(uint8_t* tile, const __nv_bfloat16* src, int ld, int row0, int rows, int col0, int hd) {
    constexpr int PITCH = Geo<HDP>::PITCH;
    const int chunks = hd >> 3, chunks_p = HDP >> 3;
    const uint32_t t = s_u32(tile);
    for (int i = threadIdx.x; i < BLK * chunks_p; i += WARPS * 32) {
        const int r = i / chunks_p, ch = i - r * chunks_p;
        if (r < rows && ch < chunks)
            cp_async16(t + r * PITCH + ch * 16, src + static_cast<size_t>(row0 + r) * ld + col0 + ch * 8);
        else
            *reinterpret_cast<uint4*>(tile + r * PITCH + ch * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
}

// A fragments of this warp's 16 rows over head_dim (row-major tile).
template <int HDP>
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[Geo<HDP>::KS][4], const uint8_t* tile, int warp, int lane) {
    const uint32_t base = s_u32(tile) + (warp * 16 + (lane & 15)) * Geo<HDP>::PITCH + (lane >> 4) * 16;
#pragma unroll
    for (int ks = 0; ks < Geo<HDP>::KS; ++ks) ldsm_x4(a[ks], base + ks * 32);
}

// acc[nt] (16 x 64) = A_frags (16 x HDP) . B_tile[64 rows][HDP]^T
template <int HDP>
__device__ __forceinline__ void mma_ab_t(float (&acc)[8][4], const uint32_t (&a)[Geo<HDP>::KS][4], const uint8_t* B, int lane) {
    const uint32_t b_base = s_u32(B) + ((lane & 7) + ((lane >> 4) & 1) * 8) * Geo<HDP>::PITCH + ((lane >> 3) & 1) * 16;
#pragma unroll
    for (int ks = 0; ks < Geo<HDP>::KS; ++ks) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            uint32_t b[4];
            ldsm_x4(b, b_base + np * 16 * Geo<HDP>::PITCH + ks * 32);
            mma_bf16(acc[2 * np], a[ks], b[0], b[1]);
            mma_bf16(acc[2 * np + 1], a[ks], b[2], b[3]);
        }
    }
}

// o[nt_d] (16 x HDP) += P_frags (16 x 64, A operand from registers) . Bt_tile[64 rows][HDP]
template <int HDP>
__device__ __forceinline__ void mma_p_b(float (&o)[Geo<HDP>::NT_D][4], const uint32_t (&pa)[4][4], const uint8_t* Bt, int lane) {
    const uint32_t b_base = s_u32(Bt) + (lane & 15) * Geo<HDP>::PITCH + (lane >> 4) * 16;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int dp = 0; dp < Geo<HDP>::NT_D / 2; ++dp) {
            uint32_t b[4];
            ldsm_x4_t(b, b_base + kk * 16 * Geo<HDP>::PITCH + dp * 32);
            mma_bf16(o[2 * dp], pa[kk], b[0], b[1]);
            mma_bf16(o[2 * dp + 1], pa[kk], b[2], b[3]);
        }
    }
}

__device__ __forceinline__ void pack_c_to_a(uint32_t (&pa)[4][4], const float (&s)[8][4]) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        pa[kk][0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
        pa[kk][1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
        pa[kk][2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pa[kk][3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
    }
}

// ------------------------------------------------------------------------------------------------
// forward: grid = (query blocks, heads).  blk_seq[b] = sequence of block b, blk_row0[b] = its first
// (global) query row.
// ------------------------------------------------------------------------------------------------
template <int HDP>
__global__ void __launch_bounds__(WARPS * 32)
flash_fwd_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const __nv_bfloat16* __restrict__ k, int ldk,
                 const __nv_bfloat16* __restrict__ v, int ldv, const int32_t* __restrict__ seq_off,
                 const int32_t* __restrict__ blk_seq, const int32_t* __restrict__ blk_row0, int hd, float scale,
                 __nv_bfloat16* __restrict__ ctx, int ldc, float* __restrict__ lse, int n_heads, float drop_p,
                 unsigned long long seed) {
    using G = Geo<HDP>;
    extern __shared__ __align__(16) uint8_t fa_smem[];
    uint8_t* Qs = fa_smem;
    uint8_t* Ks = Qs + G::TILE;
    uint8_t* Vs = Ks + G::TILE;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, tq = lane & 3;
    const int head = blockIdx.y, col0 = head * hd;
    const int seq = blk_seq[blockIdx.x];
    const int s0 = seq_off[seq], s1 = seq_off[seq + 1];
    const int qrow0 = blk_row0[blockIdx.x];
    const int qrows = min(BLK, s1 - qrow0);
    const uint32_t thr = adrop::thr8_of(drop_p);
    const float inv_keep = adrop::inv_keep_of(thr);

    stage<HDP>(Qs, q, ldq, qrow0, qrows, col0, hd);
    cp_async_wait_all();
    __syncthreads();
    uint32_t qa[G::KS][4];
    load_a_frags<HDP>(qa, Qs, warp, lane);
    float o[G::NT_D][4];
#pragma unroll
    for (int j = 0; j < G::NT_D; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[j][e] = 0.f;
    float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
    const int r_loc[2] = {warp * 16 + gq, warp * 16 + gq + 8};

    for (int kb = s0; kb < s1; kb += BLK) {
        const int krows = min(BLK, s1 - kb);
        __syncthreads();                      // previous block's tiles are no longer read
        stage<HDP>(Ks, k, ldk, kb, krows, col0, hd);
        stage<HDP>(Vs, v, ldv, kb, krows, col0, hd);
        cp_async_wait_all();
        __syncthreads();
        float s[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) s[nt][e] = 0.f;
        mma_ab_t<HDP>(s, qa, Ks, lane);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float mx = mrow[half];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int col = nt * 8 + tq * 2 + e;
                    const float x = col < krows ? s[nt][half * 2 + e] * scale : -INFINITY;
                    s[nt][half * 2 + e] = x;
                    mx = fmaxf(mx, x);
                }
            mx = quad_max(mx);
            const float corr = __expf(mrow[half] - mx);      // 0 on the first block (mrow = -inf)
            float sum = 0.f;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float p = __expf(s[nt][half * 2 + e] - mx);
                    sum += p;
                    const int col = nt * 8 + tq * 2 + e;
                    s[nt][half * 2 + e] = p * drop_factor(thr, inv_keep, drop_row_key(thr, seed, qrow0 + r_loc[half], head), kb - s0 + col);
                }
            sum = quad_sum(sum);
            lrow[half] = lrow[half] * corr + sum;
            mrow[half] = mx;
#pragma unroll
            for (int j = 0; j < G::NT_D; ++j) { o[j][half * 2] *= corr; o[j][half * 2 + 1] *= corr; }
        }
        uint32_t pa[4][4];
        pack_c_to_a(pa, s);
        mma_p_b<HDP>(o, pa, Vs, lane);
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int r = r_loc[half];
        if (r >= qrows) continue;
        const float inv = 1.f / lrow[half];
        const size_t grow = static_cast<size_t>(qrow0 + r);
#pragma unroll
        for (int j = 0; j < G::NT_D; ++j) {
            const int d = j * 8 + tq * 2;
            if (d < hd)
                *reinterpret_cast<uint32_t*>(ctx + grow * ldc + col0 + d) =
                    pack_bf16(o[j][half * 2] * inv, o[j][half * 2 + 1] * inv);
        }
        if (lse != nullptr && tq == 0) lse[grow * n_heads + head] = mrow[half] + __logf(lrow[half]);
    }
}

// delta[row, head] = sum_d dO[row, head, d] * O[row, head, d]
__global__ void flash_delta_kernel(const __nv_bfloat16* __restrict__ o, int ldo, const __nv_bfloat16* __restrict__ d_o,
                                   int lddo, int rows, int n_heads, int hd, float* __restrict__ delta) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= static_cast<long long>(rows) * n_heads) return;
    const int row = static_cast<int>(i / n_heads), head = static_cast<int>(i - static_cast<long long>(row) * n_heads);
    const __nv_bfloat16* op = o + static_cast<size_t>(row) * ldo + head * hd;
    const __nv_bfloat16* dp = d_o + static_cast<size_t>(row) * lddo + head * hd;
    float acc = 0.f;
    for (int c = 0; c < hd; c += 8) {
        float a[8], b[8];
        load_bf16x8(op + c, a);
        load_bf16x8(dp + c, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc = fmaf(a[j], b[j], acc);
    }
    delta[i] = acc;
}

// ------------------------------------------------------------------------------------------------
// backward dQ: same grid as forward.
// ------------------------------------------------------------------------------------------------
template <int HDP>
__global__ void __launch_bounds__(WARPS * 32)
flash_bwd_dq_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const __nv_bfloat16* __restrict__ k, int ldk,
                    const __nv_bfloat16* __restrict__ v, int ldv, const __nv_bfloat16* __restrict__ d_o, int lddo,
                    const float* __restrict__ lse, const float* __restrict__ delta, const int32_t* __restrict__ seq_off,
                    const int32_t* __restrict__ blk_seq, const int32_t* __restrict__ blk_row0, int hd, float scale,
                    __nv_bfloat16* __restrict__ dq, int lddq, int n_heads, float drop_p, unsigned long long seed) {
    using G = Geo<HDP>;
    extern __shared__ __align__(16) uint8_t fa_smem[];
    uint8_t* Qs = fa_smem;
    uint8_t* Os = Qs + G::TILE;
    uint8_t* Ks = Os + G::TILE;
    uint8_t* Vs = Ks + G::TILE;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, tq = lane & 3;
    const int head = blockIdx.y, col0 = head * hd;
    const int seq = blk_seq[blockIdx.x];
    const int s0 = seq_off[seq], s1 = seq_off[seq + 1];
    const int qrow0 = blk_row0[blockIdx.x];
    const int qrows = min(BLK, s1 - qrow0);
    const uint32_t thr = adrop::thr8_of(drop_p);
    const float inv_keep = adrop::inv_keep_of(thr);
    stage<HDP>(Qs, q, ldq, qrow0, qrows, col0, hd);
    stage<HDP>(Os, d_o, lddo, qrow0, qrows, col0, hd);
    cp_async_wait_all();
    __syncthreads();
    uint32_t qa[G::KS][4], oa[G::KS][4];
    load_a_frags<HDP>(qa, Qs, warp, lane);
    load_a_frags<HDP>(oa, Os, warp, lane);
    const int r_loc[2] = {warp * 16 + gq, warp * 16 + gq + 8};
    float lse_r[2], del_r[2];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const bool ok = r_loc[half] < qrows;
        const size_t idx = static_cast<size_t>(qrow0 + (ok ? r_loc[half] : 0)) * n_heads + head;
        lse_r[half] = ok ? lse[idx] : 0.f;
        del_r[half] = ok ? delta[idx] : 0.f;
    }
    float acc[G::NT_D][4];
#pragma unroll
    for (int j = 0; j < G::NT_D; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
    for (int kb = s0; kb < s1; kb += BLK) {
        const int krows = min(BLK, s1 - kb);
        __syncthreads();
        stage<HDP>(Ks, k, ldk, kb, krows, col0, hd);
        stage<HDP>(Vs, v, ldv, kb, krows, col0, hd);
        cp_async_wait_all();
        __syncthreads();
        float s[8][4], dp[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) { s[nt][e] = 0.f; dp[nt][e] = 0.f; }
        mma_ab_t<HDP>(s, qa, Ks, lane);
        mma_ab_t<HDP>(dp, oa, Vs, lane);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int half = e >> 1, col = nt * 8 + tq * 2 + (e & 1);
                const bool ok = col < krows && r_loc[half] < qrows;
                const float p = ok ? __expf(s[nt][e] * scale - lse_r[half]) : 0.f;
                const float f = drop_factor(thr, inv_keep, drop_row_key(thr, seed, qrow0 + r_loc[half], head), kb - s0 + col);
                s[nt][e] = p * (dp[nt][e] * f - del_r[half]);              // dS
            }
        uint32_t da[4][4];
        pack_c_to_a(da, s);
        mma_p_b<HDP>(acc, da, Ks, lane);
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int r = r_loc[half];
        if (r >= qrows) continue;
        const size_t grow = static_cast<size_t>(qrow0 + r);
#pragma unroll
        for (int j = 0; j < G::NT_D; ++j) {
            const int d = j * 8 + tq * 2;
            if (d < hd)
                *reinterpret_cast<uint32_t*>(dq + grow * lddq + col0 + d) =
                    pack_bf16(acc[j][half * 2] * scale, acc[j][half * 2 + 1] * scale);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward dK, dV: grid = (key blocks, heads); rows of the MMA tiles are KEYS, columns are queries.
// ------------------------------------------------------------------------------------------------
template <int HDP>
__global__ void __launch_bounds__(WARPS * 32)
flash_bwd_dkv_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const __nv_bfloat16* __restrict__ k, int ldk,
                     const __nv_bfloat16* __restrict__ v, int ldv, const __nv_bfloat16* __restrict__ d_o, int lddo,
                     const float* __restrict__ lse, const float* __restrict__ delta, const int32_t* __restrict__ seq_off,
                     const int32_t* __restrict__ blk_seq, const int32_t* __restrict__ blk_row0, int hd, float scale,
                     __nv_bfloat16* __restrict__ dk, int lddk, __nv_bfloat16* __restrict__ dv, int lddv, int n_heads,
                     float drop_p, unsigned long long seed) {
    using G = Geo<HDP>;
    extern __shared__ __align__(16) uint8_t fa_smem[];
    uint8_t* Ks = fa_smem;
    uint8_t* Vs = Ks + G::TILE;
    uint8_t* Qs = Vs + G::TILE;
    uint8_t* Os = Qs + G::TILE;
    float* lse_s = reinterpret_cast<float*>(Os + G::TILE);
    float* del_s = lse_s + BLK;
    uint32_t* rk_s = reinterpret_cast<uint32_t*>(del_s + BLK);     // dropout row keys of the staged queries
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, tq = lane & 3;
    const int head = blockIdx.y, col0 = head * hd;
    const int seq = blk_seq[blockIdx.x];
    const int s0 = seq_off[seq], s1 = seq_off[seq + 1];
    const int krow0 = blk_row0[blockIdx.x];
    const int krows = min(BLK, s1 - krow0);
    const uint32_t thr = adrop::thr8_of(drop_p);
    const float inv_keep = adrop::inv_keep_of(thr);
    stage<HDP>(Ks, k, ldk, krow0, krows, col0, hd);
    stage<HDP>(Vs, v, ldv, krow0, krows, col0, hd);
    cp_async_wait_all();
    __syncthreads();
    uint32_t ka[G::KS][4], va[G::KS][4];
    load_a_frags<HDP>(ka, Ks, warp, lane);
    load_a_frags<HDP>(va, Vs, warp, lane);
    const int r_loc[2] = {warp * 16 + gq, warp * 16 + gq + 8};       // key rows of this lane
    float dk_acc[G::NT_D][4], dv_acc[G::NT_D][4];
#pragma unroll
    for (int j = 0; j < G::NT_D; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) { dk_acc[j][e] = 0.f; dv_acc[j][e] = 0.f; }
    for (int qb = s0; qb < s1; qb += BLK) {
        const int qrows = min(BLK, s1 - qb);
        __syncthreads();
        stage<HDP>(Qs, q, ldq, qb, qrows, col0, hd);
        stage<HDP>(Os, d_o, lddo, qb, qrows, col0, hd);
        if (threadIdx.x < BLK) {
            const bool ok = threadIdx.x < qrows;
            const size_t idx = static_cast<size_t>(qb + (ok ? threadIdx.x : 0)) * n_heads + head;
            lse_s[threadIdx.x] = ok ? lse[idx] : 0.f;
            del_s[threadIdx.x] = ok ? delta[idx] : 0.f;
            rk_s[threadIdx.x] = drop_row_key(thr, seed, qb + threadIdx.x, head);
        }
        cp_async_wait_all();
        __syncthreads();
        float st[8][4], dpt[8][4];                                   // S^T, dP~^T : rows = keys, cols = queries
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) { st[nt][e] = 0.f; dpt[nt][e] = 0.f; }
        mma_ab_t<HDP>(st, ka, Qs, lane);
        mma_ab_t<HDP>(dpt, va, Os, lane);
        float pt[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int half = e >> 1, qc = nt * 8 + tq * 2 + (e & 1);  // query column
                const bool ok = qc < qrows && r_loc[half] < krows;
                const float p = ok ? __expf(st[nt][e] * scale - lse_s[qc]) : 0.f;
                const float f = drop_factor(thr, inv_keep, rk_s[qc], krow0 - s0 + r_loc[half]);
                pt[nt][e] = p * f;                                         // P~^T
                st[nt][e] = p * (dpt[nt][e] * f - del_s[qc]);              // dS^T
            }
        uint32_t pa[4][4], da[4][4];
        pack_c_to_a(pa, pt);
        pack_c_to_a(da, st);
        mma_p_b<HDP>(dv_acc, pa, Os, lane);                               // dV += P~^T dO
        mma_p_b<HDP>(dk_acc, da, Qs, lane);                               // dK += dS^T Q
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int r = r_loc[half];
        if (r >= krows) continue;
        const size_t grow = static_cast<size_t>(krow0 + r);
#pragma unroll
        for (int j = 0; j < G::NT_D; ++j) {
            const int d = j * 8 + tq * 2;
            if (d < hd) {
                *reinterpret_cast<uint32_t*>(dk + grow * lddk + col0 + d) =
                    pack_bf16(dk_acc[j][half * 2] * scale, dk_acc[j][half * 2 + 1] * scale);
                *reinterpret_cast<uint32_t*>(dv + grow * lddv + col0 + d) =
                    pack_bf16(dv_acc[j][half * 2], dv_acc[j][half * 2 + 1]);
            }
        }
    }
}

// ================================================================================================
// Resident variants for sequences whose K/V (forward) or Q/K/V/dO (backward) head slices fit in shared
// memory (T <= ~640 at head_dim 24): one CTA per (sequence, head) loads the operands ONCE, then its warps
// walk their 16-row blocks against the resident tiles without any block-level synchronisation.  This is
// the C3 regime (clips of ~450 tokens): the tiled kernels above re-stage K/V per 64-row query block and
// pay two barriers per key block, which dominates when the per-block math is only 64x64x32.
// ================================================================================================
template <int HDP>
__device__ __forceinline__ void stage_seq(uint8_t* tile, const __nv_bfloat16* src, int ld, int row0, int rows,
                                          int rows_pad, int col0, int hd) {
    constexpr int PITCH = Geo<HDP>::PITCH;
    const int chunks = hd >> 3, chunks_p = HDP >> 3;
    const uint32_t t = s_u32(tile);
    for (int i = threadIdx.x; i < rows_pad * chunks_p; i += blockDim.x) {
        const int r = i / chunks_p, ch = i - r * chunks_p;
        if (r < rows && ch < chunks)
            cp_async16(t + r * PITCH + ch * 16, src + static_cast<size_t>(row0 + r) * ld + col0 + ch * 8);
        else
            *reinterpret_cast<uint4*>(tile + r * PITCH + ch * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
}

// A fragments of 16 rows starting at `tile` (row-major, PITCH) — rows are addressed relative to `tile`.
template <int HDP>
__device__ __forceinline__ void load_a_frags_at(uint32_t (&a)[Geo<HDP>::KS][4], const uint8_t* tile, int lane) {
    const uint32_t base = s_u32(tile) + (lane & 15) * Geo<HDP>::PITCH + (lane >> 4) * 16;
#pragma unroll
    for (int ks = 0; ks < Geo<HDP>::KS; ++ks) ldsm_x4(a[ks], base + ks * 32);
}

// A fragments of 16 query rows straight from global memory (rows >= rows_valid and columns >= hd read as 0).
template <int HDP>
__device__ __forceinline__ void load_a_frags_global(uint32_t (&a)[Geo<HDP>::KS][4], const __nv_bfloat16* src, int ld,
                                                    int row0, int rows_valid, int col0, int hd, int lane) {
    const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int ks = 0; ks < Geo<HDP>::KS; ++ks) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int r = gq + (e & 1) * 8, c = ks * 16 + (e >> 1) * 8 + tq * 2;
            a[ks][e] = (r < rows_valid && c < hd)
                           ? *reinterpret_cast<const uint32_t*>(src + static_cast<size_t>(row0 + r) * ld + col0 + c) : 0u;
        }
    }
}

template <int HDP>
__global__ void __launch_bounds__(128)
flash_fwd_res_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const __nv_bfloat16* __restrict__ k, int ldk,
                     const __nv_bfloat16* __restrict__ v, int ldv, const int32_t* __restrict__ seq_off, int hd,
                     float scale, __nv_bfloat16* __restrict__ ctx, int ldc, float* __restrict__ lse, int n_heads,
                     int t_pad_max, float drop_p, unsigned long long seed) {
    using G = Geo<HDP>;
    extern __shared__ __align__(16) uint8_t fa_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, tq = lane & 3;
    const int seq = blockIdx.x, head = blockIdx.y, col0 = head * hd;
    const int s0 = seq_off[seq], T = seq_off[seq + 1] - s0;
    if (T <= 0) return;
    const int t_pad = (T + BLK - 1) / BLK * BLK;
    uint8_t* Ks = fa_smem;
    uint8_t* Vs = Ks + static_cast<size_t>(t_pad_max) * G::PITCH;
    const uint32_t thr = adrop::thr8_of(drop_p);
    const float inv_keep = adrop::inv_keep_of(thr);
    stage_seq<HDP>(Ks, k, ldk, s0, T, t_pad, col0, hd);
    stage_seq<HDP>(Vs, v, ldv, s0, T, t_pad, col0, hd);
    cp_async_wait_all();
    __syncthreads();
    for (int qb = warp * 16; qb < T; qb += 64) {
        const int qrows = min(16, T - qb);
        uint32_t qa[G::KS][4];
        load_a_frags_global<HDP>(qa, q, ldq, s0 + qb, qrows, col0, hd, lane);
        float o[G::NT_D][4];
#pragma unroll
        for (int j = 0; j < G::NT_D; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) o[j][e] = 0.f;
        float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
        const int r_loc[2] = {gq, gq + 8};
        for (int kb = 0; kb < T; kb += BLK) {
            const int krows = min(BLK, T - kb);
            float s[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) s[nt][e] = 0.f;
            mma_ab_t<HDP>(s, qa, Ks + static_cast<size_t>(kb) * G::PITCH, lane);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float mx = mrow[half];
#pragma unroll
                for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int col = nt * 8 + tq * 2 + e;
                        const float x = col < krows ? s[nt][half * 2 + e] * scale : -INFINITY;
                        s[nt][half * 2 + e] = x;
                        mx = fmaxf(mx, x);
                    }
                mx = quad_max(mx);
                const float corr = __expf(mrow[half] - mx);
                float sum = 0.f;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float p = __expf(s[nt][half * 2 + e] - mx);
                        sum += p;
                        const int col = nt * 8 + tq * 2 + e;
                        s[nt][half * 2 + e] = p * drop_factor(thr, inv_keep, drop_row_key(thr, seed, s0 + qb + r_loc[half], head), kb + col);
                    }
                sum = quad_sum(sum);
                lrow[half] = lrow[half] * corr + sum;
                mrow[half] = mx;
#pragma unroll
                for (int j = 0; j < G::NT_D; ++j) { o[j][half * 2] *= corr; o[j][half * 2 + 1] *= corr; }
            }
            uint32_t pa[4][4];
            pack_c_to_a(pa, s);
            mma_p_b<HDP>(o, pa, Vs + static_cast<size_t>(kb) * G::PITCH, lane);
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int r = r_loc[half];
            if (r >= qrows) continue;
            const float inv = 1.f / lrow[half];
            const size_t grow = static_cast<size_t>(s0 + qb + r);
#pragma unroll
            for (int j = 0; j < G::NT_D; ++j) {
                const int d = j * 8 + tq * 2;
                if (d < hd)
                    *reinterpret_cast<uint32_t*>(ctx + grow * ldc + col0 + d) =
                        pack_bf16(o[j][half * 2] * inv, o[j][half * 2 + 1] * inv);
            }
            if (lse != nullptr && tq == 0) lse[grow * n_heads + head] = mrow[half] + __logf(lrow[half]);
        }
    }
}

// Backward, resident: Q, K, V, dO of the (sequence, head) in shared memory, delta computed in the kernel;
// phase 1 (query-row blocks -> dQ), phase 2 (key-row blocks -> dK, dV).  8 warps, one barrier in total.
template <int HDP>
__global__ void __launch_bounds__(256)
flash_bwd_res_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const __nv_bfloat16* __restrict__ k, int ldk,
                     const __nv_bfloat16* __restrict__ v, int ldv, const __nv_bfloat16* __restrict__ o, int ldo,
                     const __nv_bfloat16* __restrict__ d_o, int lddo, const float* __restrict__ lse,
                     const int32_t* __restrict__ seq_off, int hd, float scale, __nv_bfloat16* __restrict__ dq, int lddq,
                     __nv_bfloat16* __restrict__ dk, int lddk, __nv_bfloat16* __restrict__ dv, int lddv, int n_heads,
                     int t_pad_max, float drop_p, unsigned long long seed) {
    using G = Geo<HDP>;
    extern __shared__ __align__(16) uint8_t fa_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, tq = lane & 3;
    const int seq = blockIdx.x, head = blockIdx.y, col0 = head * hd;
    const int s0 = seq_off[seq], T = seq_off[seq + 1] - s0;
    if (T <= 0) return;
    const int t_pad = (T + BLK - 1) / BLK * BLK;
    const size_t tile = static_cast<size_t>(t_pad_max) * G::PITCH;
    uint8_t* Qs = fa_smem;
    uint8_t* Ks = Qs + tile;
    uint8_t* Vs = Ks + tile;
    uint8_t* Os = Vs + tile;
    float* lse_s = reinterpret_cast<float*>(Os + tile);
    float* del_s = lse_s + t_pad_max;
    uint32_t* rk_s = reinterpret_cast<uint32_t*>(del_s + t_pad_max);
    const uint32_t thr = adrop::thr8_of(drop_p);
    const float inv_keep = adrop::inv_keep_of(thr);
    stage_seq<HDP>(Qs, q, ldq, s0, T, t_pad, col0, hd);
    stage_seq<HDP>(Ks, k, ldk, s0, T, t_pad, col0, hd);
    stage_seq<HDP>(Vs, v, ldv, s0, T, t_pad, col0, hd);
    stage_seq<HDP>(Os, d_o, lddo, s0, T, t_pad, col0, hd);
    for (int r = threadIdx.x; r < t_pad; r += blockDim.x) {
        float acc = 0.f, l = 0.f;
        if (r < T) {
            const __nv_bfloat16* op = o + static_cast<size_t>(s0 + r) * ldo + col0;
            const __nv_bfloat16* dp = d_o + static_cast<size_t>(s0 + r) * lddo + col0;
            for (int c = 0; c < hd; c += 8) {
                float a[8], b[8];
                load_bf16x8(op + c, a);
                load_bf16x8(dp + c, b);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc = fmaf(a[j], b[j], acc);
            }
            l = lse[static_cast<size_t>(s0 + r) * n_heads + head];
        }
        del_s[r] = acc;
        lse_s[r] = l;
        rk_s[r] = drop_row_key(thr, seed, s0 + r, head);
    }
    cp_async_wait_all();
    __syncthreads();
    const int nblk16 = (T + 15) / 16;
    // ---------------- phase 1: dQ
    for (int b = warp; b < nblk16; b += 8) {
        const int qb = b * 16;
        const int qrows = min(16, T - qb);
        uint32_t qa[G::KS][4], oa[G::KS][4];
        load_a_frags_at<HDP>(qa, Qs + static_cast<size_t>(qb) * G::PITCH, lane);
        load_a_frags_at<HDP>(oa, Os + static_cast<size_t>(qb) * G::PITCH, lane);
        const int r_loc[2] = {gq, gq + 8};
        const float lse_r[2] = {lse_s[qb + gq], lse_s[min(qb + gq + 8, t_pad - 1)]};
        const float del_r[2] = {del_s[qb + gq], del_s[min(qb + gq + 8, t_pad - 1)]};
        float acc[G::NT_D][4];
#pragma unroll
        for (int j = 0; j < G::NT_D; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
        for (int kb = 0; kb < T; kb += BLK) {
            const int krows = min(BLK, T - kb);
            float s[8][4], dp[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) { s[nt][e] = 0.f; dp[nt][e] = 0.f; }
            mma_ab_t<HDP>(s, qa, Ks + static_cast<size_t>(kb) * G::PITCH, lane);
            mma_ab_t<HDP>(dp, oa, Vs + static_cast<size_t>(kb) * G::PITCH, lane);
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int half = e >> 1, col = nt * 8 + tq * 2 + (e & 1);
                    const bool ok = col < krows && r_loc[half] < qrows;
                    const float p = ok ? __expf(s[nt][e] * scale - lse_r[half]) : 0.f;
                    const float f = drop_factor(thr, inv_keep, rk_s[min(qb + r_loc[half], t_pad - 1)], kb + col);
                    s[nt][e] = p * (dp[nt][e] * f - del_r[half]);
                }
            uint32_t da[4][4];
            pack_c_to_a(da, s);
            mma_p_b<HDP>(acc, da, Ks + static_cast<size_t>(kb) * G::PITCH, lane);
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int r = r_loc[half];
            if (r >= qrows) continue;
            const size_t grow = static_cast<size_t>(s0 + qb + r);
#pragma unroll
            for (int j = 0; j < G::NT_D; ++j) {
                const int d = j * 8 + tq * 2;
                if (d < hd)
                    *reinterpret_cast<uint32_t*>(dq + grow * lddq + col0 + d) =
                        pack_bf16(acc[j][half * 2] * scale, acc[j][half * 2 + 1] * scale);
            }
        }
    }
    // ---------------- phase 2: dK, dV (rows of the MMA tiles are keys, columns are queries)
    for (int b = warp; b < nblk16; b += 8) {
        const int kb = b * 16;
        const int krows = min(16, T - kb);
        uint32_t ka[G::KS][4], va[G::KS][4];
        load_a_frags_at<HDP>(ka, Ks + static_cast<size_t>(kb) * G::PITCH, lane);
        load_a_frags_at<HDP>(va, Vs + static_cast<size_t>(kb) * G::PITCH, lane);
        const int r_loc[2] = {gq, gq + 8};
        float dk_acc[G::NT_D][4], dv_acc[G::NT_D][4];
#pragma unroll
        for (int j = 0; j < G::NT_D; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) { dk_acc[j][e] = 0.f; dv_acc[j][e] = 0.f; }
        for (int qb = 0; qb < T; qb += BLK) {
            const int qrows = min(BLK, T - qb);
            float st[8][4], dpt[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) { st[nt][e] = 0.f; dpt[nt][e] = 0.f; }
            mma_ab_t<HDP>(st, ka, Qs + static_cast<size_t>(qb) * G::PITCH, lane);
            mma_ab_t<HDP>(dpt, va, Os + static_cast<size_t>(qb) * G::PITCH, lane);
            float pt[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int half = e >> 1, qc = nt * 8 + tq * 2 + (e & 1);
                    const bool ok = qc < qrows && r_loc[half] < krows;
                    const float p = ok ? __expf(st[nt][e] * scale - lse_s[qb + qc]) : 0.f;
                    const float f = drop_factor(thr, inv_keep, rk_s[qb + qc], kb + r_loc[half]);
                    pt[nt][e] = p * f;
                    st[nt][e] = p * (dpt[nt][e] * f - del_s[qb + qc]);
                }
            uint32_t pa[4][4], da[4][4];
            pack_c_to_a(pa, pt);
            pack_c_to_a(da, st);
            mma_p_b<HDP>(dv_acc, pa, Os + static_cast<size_t>(qb) * G::PITCH, lane);
            mma_p_b<HDP>(dk_acc, da, Qs + static_cast<size_t>(qb) * G::PITCH, lane);
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int r = r_loc[half];
            if (r >= krows) continue;
            const size_t grow = static_cast<size_t>(s0 + kb + r);
#pragma unroll
            for (int j = 0; j < G::NT_D; ++j) {
                const int d = j * 8 + tq * 2;
                if (d < hd) {
                    *reinterpret_cast<uint32_t*>(dk + grow * lddk + col0 + d) =
                        pack_bf16(dk_acc[j][half * 2] * scale, dk_acc[j][half * 2 + 1] * scale);
                    *reinterpret_cast<uint32_t*>(dv + grow * lddv + col0 + d) =
                        pack_bf16(dv_acc[j][half * 2], dv_acc[j][half * 2 + 1]);
                }
            }
        }
    }
}

template <typename Kern>
static int set_smem(Kern kern, size_t smem) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
    return 0;
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace fa
}  // namespace vsgg

using namespace vsgg;
using namespace vsgg::fa;

static int flash_args_ok(int hd, int n_heads, const void* q, int ldq, const void* k, int ldk, const void* v, int ldv) {
    if (hd <= 0 || (hd & 7) || hd > 64 || n_heads <= 0 || n_heads > 64)
        return set_error(B200VSGG_ERR_BAD_ARG, "attn_flash: head_dim must be a multiple of 8 and <= 64, heads <= 64");
    if (!al16(q) || !al16(k) || !al16(v) || (ldq & 7) || (ldk & 7) || (ldv & 7))
        return set_error(B200VSGG_ERR_BAD_ARG, "attn_flash: q/k/v must be 16-byte aligned with ld % 8 == 0");
    return 0;
}

// Largest padded sequence length whose resident tiles fit in shared memory (0 = use the tiled kernels).
static int resident_t_pad(int max_len, int pitch, int tiles, int extra_per_row) {
    if (max_len <= 0) return 0;
    const int t_pad = (max_len + BLK - 1) / BLK * BLK;
    const size_t need = static_cast<size_t>(t_pad) * (static_cast<size_t>(tiles) * pitch + extra_per_row);
    return need <= 220 * 1024 ? t_pad : 0;
}

extern "C" int b200vsgg_attn_flash_fwd(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv,
                                       const int32_t* seq_off, const int32_t* blk_seq, const int32_t* blk_row0,
                                       int32_t n_blocks, int32_t n_heads, int32_t head_dim, float scale, void* ctx,
                                       int32_t ldc, float* lse, float drop_p, uint64_t seed, void* stream, int32_t n_seq,
                                       int32_t max_len) {
    if (!q || !k || !v || !seq_off || !blk_seq || !blk_row0 || !ctx) return set_error(B200VSGG_ERR_BAD_ARG, "attn_flash_fwd: null pointer");
    int rc = flash_args_ok(head_dim, n_heads, q, ldq, k, ldk, v, ldv);
    if (rc) return rc;
    if (n_blocks == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    {   // resident path: K/V of a (sequence, head) fit in shared memory
        const int pitch = head_dim <= 32 ? Geo<32>::PITCH : (head_dim <= 48 ? Geo<48>::PITCH : Geo<64>::PITCH);
        const int t_pad = n_seq > 0 ? resident_t_pad(max_len, pitch, 2, 0) : 0;
        if (t_pad > 0 && t_pad <= 1024) {
            const size_t smem = 2ull * t_pad * pitch;
            dim3 grid(n_seq, n_heads);
#define FA_FWD_RES(HDP)                                                                                          \
    {                                                                                                            \
        static size_t cur = 0;                                                                                   \
        if (smem > cur) { if ((rc = set_smem(flash_fwd_res_kernel<HDP>, smem))) return rc; cur = smem; }         \
        flash_fwd_res_kernel<HDP><<<grid, 128, smem, st>>>(                                                      \
            (const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, ldk, (const __nv_bfloat16*)v, ldv, seq_off,   \
            head_dim, scale, (__nv_bfloat16*)ctx, ldc, lse, n_heads, t_pad, drop_p, seed);                       \
    }
            if (head_dim <= 32) FA_FWD_RES(32) else if (head_dim <= 48) FA_FWD_RES(48) else FA_FWD_RES(64)
#undef FA_FWD_RES
            VSGG_CUDA_CHECK_LAUNCH();
            return 0;
        }
    }
    dim3 grid(n_blocks, n_heads);
#define FA_FWD(HDP)                                                                                              \
    {                                                                                                            \
        const size_t smem = 3 * Geo<HDP>::TILE;                                                                  \
        static bool done = false;                                                                                \
        if (!done) { if ((rc = set_smem(flash_fwd_kernel<HDP>, smem))) return rc; done = true; }                 \
        flash_fwd_kernel<HDP><<<grid, WARPS * 32, smem, st>>>(                                                   \
            (const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, ldk, (const __nv_bfloat16*)v, ldv, seq_off,   \
            blk_seq, blk_row0, head_dim, scale, (__nv_bfloat16*)ctx, ldc, lse, n_heads, drop_p, seed);           \
    }
    if (head_dim <= 32) FA_FWD(32) else if (head_dim <= 48) FA_FWD(48) else FA_FWD(64)
#undef FA_FWD
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_attn_flash_bwd(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv,
                                       const void* ctx, int32_t ldc, const void* dctx, int32_t lddc, const float* lse,
                                       float* delta, const int32_t* seq_off, const int32_t* blk_seq,
                                       const int32_t* blk_row0, int32_t n_blocks, int32_t n_rows, int32_t n_heads,
                                       int32_t head_dim, float scale, void* dq, int32_t lddq, void* dk, int32_t lddk,
                                       void* dv, int32_t lddv, float drop_p, uint64_t seed, void* stream, int32_t n_seq,
                                       int32_t max_len) {
    if (!q || !k || !v || !ctx || !dctx || !lse || !delta || !seq_off || !blk_seq || !blk_row0 || !dq || !dk || !dv)
        return set_error(B200VSGG_ERR_BAD_ARG, "attn_flash_bwd: null pointer");
    int rc = flash_args_ok(head_dim, n_heads, q, ldq, k, ldk, v, ldv);
    if (rc) return rc;
    if (!al16(ctx) || !al16(dctx) || (ldc & 7) || (lddc & 7) || (lddq & 1) || (lddk & 1) || (lddv & 1))
        return set_error(B200VSGG_ERR_BAD_ARG, "attn_flash_bwd: ctx/dctx alignment");
    if (n_blocks == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    {   // resident path: Q, K, V, dO of a (sequence, head) fit in shared memory; one launch for dQ, dK, dV
        const int pitch = head_dim <= 32 ? Geo<32>::PITCH : (head_dim <= 48 ? Geo<48>::PITCH : Geo<64>::PITCH);
        const int t_pad = n_seq > 0 ? resident_t_pad(max_len, pitch, 4, 12) : 0;
        if (t_pad > 0) {
            const size_t smem = 4ull * t_pad * pitch + 12ull * t_pad;
            dim3 grid(n_seq, n_heads);
#define FA_BWD_RES(HDP)                                                                                          \
    {                                                                                                            \
        static size_t cur = 0;                                                                                   \
        if (smem > cur) { if ((rc = set_smem(flash_bwd_res_kernel<HDP>, smem))) return rc; cur = smem; }         \
        flash_bwd_res_kernel<HDP><<<grid, 256, smem, st>>>(                                                      \
            (const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, ldk, (const __nv_bfloat16*)v, ldv,            \
            (const __nv_bfloat16*)ctx, ldc, (const __nv_bfloat16*)dctx, lddc, lse, seq_off, head_dim, scale,     \
            (__nv_bfloat16*)dq, lddq, (__nv_bfloat16*)dk, lddk, (__nv_bfloat16*)dv, lddv, n_heads, t_pad, drop_p, \
            seed);                                                                                               \
    }
            if (head_dim <= 32) FA_BWD_RES(32) else if (head_dim <= 48) FA_BWD_RES(48) else FA_BWD_RES(64)
#undef FA_BWD_RES
            VSGG_CUDA_CHECK_LAUNCH();
            return 0;
        }
    }
    {
        const long long items = static_cast<long long>(n_rows) * n_heads;
        flash_delta_kernel<<<(unsigned)((items + 255) / 256), 256, 0, st>>>((const __nv_bfloat16*)ctx, ldc,
                                                                            (const __nv_bfloat16*)dctx, lddc, n_rows,
                                                                            n_heads, head_dim, delta);
    }
    dim3 grid(n_blocks, n_heads);
#define FA_BWD(HDP)                                                                                              \
    {                                                                                                            \
        const size_t smem_q = 4 * Geo<HDP>::TILE;                                                                \
        const size_t smem_kv = 4 * Geo<HDP>::TILE + 3 * BLK * sizeof(float);                                     \
        static bool done = false;                                                                                \
        if (!done) {                                                                                             \
            if ((rc = set_smem(flash_bwd_dq_kernel<HDP>, smem_q))) return rc;                                    \
            if ((rc = set_smem(flash_bwd_dkv_kernel<HDP>, smem_kv))) return rc;                                  \
            done = true;                                                                                         \
        }                                                                                                        \
        flash_bwd_dq_kernel<HDP><<<grid, WARPS * 32, smem_q, st>>>(                                              \
            (const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, ldk, (const __nv_bfloat16*)v, ldv,            \
            (const __nv_bfloat16*)dctx, lddc, lse, delta, seq_off, blk_seq, blk_row0, head_dim, scale,           \
            (__nv_bfloat16*)dq, lddq, n_heads, drop_p, seed);                                                    \
        flash_bwd_dkv_kernel<HDP><<<grid, WARPS * 32, smem_kv, st>>>(                                            \
            (const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, ldk, (const __nv_bfloat16*)v, ldv,            \
            (const __nv_bfloat16*)dctx, lddc, lse, delta, seq_off, blk_seq, blk_row0, head_dim, scale,           \
            (__nv_bfloat16*)dk, lddk, (__nv_bfloat16*)dv, lddv, n_heads, drop_p, seed);                          \
    }
    if (head_dim <= 32) FA_BWD(32) else if (head_dim <= 48) FA_BWD(48) else FA_BWD(64)
#undef FA_BWD
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}
