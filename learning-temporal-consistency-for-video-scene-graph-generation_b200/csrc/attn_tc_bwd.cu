// Blackwell-native backward of the TokenGT attention (multihead_attention.py:135-183 of the reference; the forward is
// attn_tc.cu).  Two kernels, no atomics, nothing of size T^2 in HBM; the probabilities are recomputed from the forward's
// log-sum-exp, the dropout mask from its counters (attn_dropout.cuh):
//
//   dQ kernel    CTA = 128-query tile of one (clip, head); loop over key tiles j:
//                  S = Q K_j^T, dP = dO V_j^T                      tcgen05.mma, operands in shared memory (TMA)
//                  dS = P o (M dP - delta)                         thread = query row, tcgen05.ld / tcgen05.st
//                  dQ += dS K_j                                    tcgen05.mma, A = dS from TENSOR MEMORY, B = K_j MN-major
//   dK/dV kernel CTA = 128-key tile; loop over query tiles i, everything transposed (thread = key row):
//                  S^T = K Q_i^T, dP^T = V dO_i^T
//                  P~^T = M o P^T, dS^T = P^T o (M dP^T - delta)   lse / delta / dropout seeds of the 128 queries are
//                                                                  staged in shared memory by the producer warp
//                  dV += P~^T dO_i, dK += dS^T Q_i                 A from tensor memory, B = dO_i / Q_i MN-major
// with P = exp2(S c - lse), M = keep mask / (1-p), delta = rowsum(dO o O).  Constant factors (softmax scale, 1/(1-p))
// are folded into the epilogues.  The SAME 16 KB swizzled tile serves as K-major operand of the S-type products and as
// MN-major operand of the accumulating ones, so every tile is fetched once per use site.
//
// 320 threads: warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 two softmax warpgroups that split the 128 columns of a
// tile (the backward needs no row reductions, so a row can be shared by two threads); both S-type accumulators are pulled
// into registers at once and released, so the next tile's products overlap this tile's exponentials.  One CTA per SM
// (512 TMEM columns: S, dP, the bf16 operands, the accumulators).
// Roofline: SFU/issue-bound like the forward (one exp per element and kernel: 2 x 1024 clk per tile pair of MUFU, ~9 and
// ~13 issue slots per element with dropout) against 5 x 128 clk of tcgen05.mma per kernel.
#include "attn_tc_common.cuh"
#include <cstdlib>

namespace vsgg {
namespace atc {

constexpr int BWD_THREADS = 320;
constexpr int BWD_TMEM_COLS = 512;
constexpr float LOG2E = 1.4426950408889634f;

// delta[row, head] = sum_d dO[row, head, d] * O[row, head, d]
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, int ldo, const __nv_bfloat16* __restrict__ d_o,
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

// ================================================================================================================
// dQ
// ================================================================================================================
constexpr int DQ_COL_S = 0, DQ_COL_DP = 128, DQ_COL_DS = 256, DQ_COL_DQ = 320;
constexpr int DQ_KST = 3;   // K_j feeds S_j (early) AND dQ_j (a whole softmax period later): 2 stages would stall the next S on the TMA
constexpr int DQ_SMEM_BYTES = TILE_BYTES * (4 + DQ_KST + 2) + 256 + 1024;   // Q[2], dO[2], K[DQ_KST], V[2]

// PERSISTENT like the forward: one CTA per SM walks units u = blockIdx.x, + gridDim.x, ... with running barrier phases.
template <int HDN>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_tc_dq_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                  const __grid_constant__ CUtensorMap tv, const __grid_constant__ CUtensorMap tdo,
                  const int32_t* __restrict__ seq_off, const int32_t* __restrict__ blk_seq,
                  const int32_t* __restrict__ blk_row0, int n_units, int n_heads, int hd, float scale,
                  const float* __restrict__ lse, const float* __restrict__ delta, __nv_bfloat16* __restrict__ dq, int lddq,
                  float drop_p, unsigned long long seed) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* Qs = smem;                       // [2]
    uint8_t* dOs = Qs + 2 * TILE_BYTES;       // [2]
    uint8_t* Ks = dOs + 2 * TILE_BYTES;       // [DQ_KST]
    uint8_t* Vs = Ks + DQ_KST * TILE_BYTES;   // [2]
    uint64_t* bars = reinterpret_cast<uint64_t*>(Vs + 2 * TILE_BYTES);
    uint64_t* qo_full = bars;                 // [2]
    uint64_t* qo_empty = qo_full + 2;         // [2]
    uint64_t* k_full = qo_empty + 2;
    uint64_t* k_empty = k_full + DQ_KST;
    uint64_t* v_full = k_empty + DQ_KST;
    uint64_t* v_empty = v_full + 2;
    uint64_t* sdp_full = v_empty + 2;
    uint64_t* s_free = sdp_full + 1;
    uint64_t* ds_full = s_free + 1;
    uint64_t* dq_done = ds_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dq_done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tq);
        ptx::prefetch_tmap(&tk);
        ptx::prefetch_tmap(&tv);
        ptx::prefetch_tmap(&tdo);
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&qo_full[i], 1);
            ptx::mbar_init(&qo_empty[i], 1);
            ptx::mbar_init(&v_full[i], 1);
            ptx::mbar_init(&v_empty[i], 1);
        }
        for (int i = 0; i < DQ_KST; ++i) {
            ptx::mbar_init(&k_full[i], 1);
            ptx::mbar_init(&k_empty[i], 1);
        }
        ptx::mbar_init(sdp_full, 1);
        ptx::mbar_init(s_free, 256);
        ptx::mbar_init(ds_full, 256);
        ptx::mbar_init(dq_done, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, BWD_TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t qc = 0, kc = 0;
            for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++qc) {
                const int blk = u / n_heads, head = u - blk * n_heads;
                const int seq = blk_seq[blk];
                const int s0 = seq_off[seq], s1 = seq_off[seq + 1];
                const int nkb = (s1 - s0 + BKV - 1) / BKV;
                const int qrow0 = blk_row0[blk];
                const uint32_t qs = qc & 1u;
                ptx::mbar_wait(&qo_empty[qs], ((qc >> 1) & 1u) ^ 1u);
                ptx::mbar_expect_tx(&qo_full[qs], 2 * TILE_BYTES);
                ptx::tma_load_3d(Qs + qs * TILE_BYTES, &tq, &qo_full[qs], 0, qrow0, head);
                ptx::tma_load_3d(dOs + qs * TILE_BYTES, &tdo, &qo_full[qs], 0, qrow0, head);
                for (int j = 0; j < nkb; ++j, ++kc) {
                    const uint32_t kst = kc % DQ_KST, vst = kc & 1u;
                    ptx::mbar_wait(&k_empty[kst], ((kc / DQ_KST) & 1u) ^ 1u);
                    ptx::mbar_expect_tx(&k_full[kst], TILE_BYTES);
                    ptx::tma_load_3d(Ks + kst * TILE_BYTES, &tk, &k_full[kst], 0, s0 + j * BKV, head);
                    ptx::mbar_wait(&v_empty[vst], ((kc >> 1) & 1u) ^ 1u);
                    ptx::mbar_expect_tx(&v_full[vst], TILE_BYTES);
                    ptx::tma_load_3d(Vs + vst * TILE_BYTES, &tv, &v_full[vst], 0, s0 + j * BKV, head);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = ptx::make_idesc_bf16(BQ, BKV, 0, 0);
            constexpr uint32_t idesc_a = ptx::make_idesc_bf16(BQ, HDN, 0, 1);     // A from TMEM, B MN-major
            const int ks_s = (hd + 15) >> 4;
            const uint32_t t_s = tmem_base + DQ_COL_S, t_dp = tmem_base + DQ_COL_DP, t_ds = tmem_base + DQ_COL_DS,
                           t_dq = tmem_base + DQ_COL_DQ;
            uint32_t qc = 0, g0 = 0;
            for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++qc) {
                const int blk = u / n_heads;
                const int seq = blk_seq[blk];
                const int s0 = seq_off[seq], s1 = seq_off[seq + 1];
                const int nkb = (s1 - s0 + BKV - 1) / BKV;
                const uint32_t qs = qc & 1u;
                const uint32_t qa = ptx::smem_u32(Qs + qs * TILE_BYTES), oa = ptx::smem_u32(dOs + qs * TILE_BYTES);
                auto issue_sdp = [&](int j) {
                    const uint32_t g = g0 + j, kst = g % DQ_KST, vst = g & 1u;
                    ptx::mbar_wait(&k_full[kst], (g / DQ_KST) & 1u);
                    ptx::mbar_wait(&v_full[vst], (g >> 1) & 1u);
                    if (g > 0) ptx::mbar_wait(s_free, (g - 1) & 1u);
                    ptx::tc_fence_after();
                    const uint32_t kb = ptx::smem_u32(Ks + kst * TILE_BYTES), vb = ptx::smem_u32(Vs + vst * TILE_BYTES);
                    for (int k = 0; k < ks_s; ++k)
                        ptx::umma_bf16(t_s, ptx::make_smem_desc_sw128(qa + k * 32, 16, 1024),
                                       ptx::make_smem_desc_sw128(kb + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
                    for (int k = 0; k < ks_s; ++k)
                        ptx::umma_bf16(t_dp, ptx::make_smem_desc_sw128(oa + k * 32, 16, 1024),
                                       ptx::make_smem_desc_sw128(vb + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
                    ptx::umma_commit(&v_empty[vst]);
                    if (j == nkb - 1) ptx::umma_commit(&qo_empty[qs]);       // last products that read Q / dO of this unit
                    ptx::umma_commit(sdp_full);
                };
                ptx::mbar_wait(&qo_full[qs], (qc >> 1) & 1u);
                issue_sdp(0);
                for (int j = 0; j < nkb; ++j) {
                    if (j + 1 < nkb) issue_sdp(j + 1);
                    const uint32_t g = g0 + j, kst = g % DQ_KST;
                    const int kvalid = min(BKV, s1 - (s0 + j * BKV));
                    const int ks_o = (kvalid + 15) >> 4;
                    ptx::mbar_wait(ds_full, g & 1u);
                    ptx::tc_fence_after();
                    const uint32_t kb = ptx::smem_u32(Ks + kst * TILE_BYTES);
                    for (int k = 0; k < ks_o; ++k)
                        ptx::umma_bf16_ts(t_dq, t_ds + k * 8, ptx::make_smem_desc_sw128(kb + k * 2048, 8192, 1024), idesc_a,
                                          (j != 0 || k != 0) ? 1u : 0u);
                    ptx::umma_commit(&k_empty[kst]);
                    ptx::umma_commit(dq_done);
                }
                g0 += nkb;
            }
        }
    } else {
        const int quad = warp & 3;
        const int ch = (warp - 2) >> 2;                 // column half of the tile this warpgroup owns
        const int row = quad * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
        const uint32_t t_s = tmem_base + lane_addr + DQ_COL_S + ch * 64, t_dp = tmem_base + lane_addr + DQ_COL_DP + ch * 64,
                       t_ds = tmem_base + lane_addr + DQ_COL_DS + ch * 32, t_dq = tmem_base + lane_addr + DQ_COL_DQ;
        const uint32_t thr = adrop::thr8_of(drop_p);
        const float inv_keep = thr ? adrop::inv_keep_of(thr) : 1.f;
        const uint32_t K8 = (256u - thr) * 0x00010001u;
        const float scale_log2 = scale * LOG2E;
        uint32_t g = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const int blk = u / n_heads, head = u - blk * n_heads;
            const int seq = blk_seq[blk];
            const int s0 = seq_off[seq], s1 = seq_off[seq + 1];
            const int qrow0 = blk_row0[blk];
            const int qrows = min(BQ, s1 - qrow0);
            const int nkb = (s1 - s0 + BKV - 1) / BKV;
            const bool row_ok = row < qrows;
            const size_t gidx = static_cast<size_t>(qrow0 + (row_ok ? row : 0)) * n_heads + head;
            const float lse2 = row_ok ? lse[gidx] * LOG2E : INFINITY;       // invalid rows: P = 0
            const float delta_s = row_ok ? delta[gidx] / inv_keep : 0.f;    // dS = ik * P (keep ? dP : 0  -  delta / ik)
            const uint32_t rk = thr ? adrop::row_key(seed, qrow0 + row, head) : 0u;
            const float nlse = -lse2, ndelta = -delta_s;
            for (int j = 0; j < nkb; ++j, ++g) {
                const int kvalid = min(BKV, s1 - (s0 + j * BKV)) - ch * 64;   // valid keys among this half's 64 columns
                uint32_t rs[64], rd[64];
                ptx::mbar_wait(sdp_full, g & 1u);
                ptx::tc_fence_after();
                ptx::tmem_ld_32x32b_x32(t_s, reinterpret_cast<uint32_t(&)[32]>(rs[0]));
                ptx::tmem_ld_32x32b_x32(t_s + 32, reinterpret_cast<uint32_t(&)[32]>(rs[32]));
                ptx::tmem_ld_32x32b_x32(t_dp, reinterpret_cast<uint32_t(&)[32]>(rd[0]));
                ptx::tmem_ld_32x32b_x32(t_dp + 32, reinterpret_cast<uint32_t(&)[32]>(rd[32]));
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                ptx::mbar_arrive(s_free);
                uint32_t pk[32];
                const uint32_t sd = thr ? adrop::stream_seed(rk, static_cast<uint32_t>(j) * 2u + ch) : 0u;
#pragma unroll
                for (int g4 = 0; g4 < 16; ++g4) {
                    uint32_t e = 0x01000100u, o = 0x01000100u;
                    if (thr) {
                        const uint32_t t = adrop::draw(sd, g4 >> 1, g4 & 1);
                        e = ((t & 0x00FF00FFu) + K8) & 0x01000100u;
                        o = (((t >> 8) & 0x00FF00FFu) + K8) & 0x01000100u;
                    }
                    // packed fp32x2 pipes (FFMA2 / FADD2 / FMUL2): one issue slot per pair of keys for the exponent
                    // argument, the (dP - delta) term and the product
                    float ds[4], a[4], dpe[4];
                    const int i0 = 4 * g4;
                    ptx::fma2(a[0], a[1], __uint_as_float(rs[i0]), __uint_as_float(rs[i0 + 1]), scale_log2, scale_log2, nlse, nlse);
                    ptx::fma2(a[2], a[3], __uint_as_float(rs[i0 + 2]), __uint_as_float(rs[i0 + 3]), scale_log2, scale_log2, nlse, nlse);
                    // keys (0,1) of the group sit in the even bytes (bits 8 / 24 of e), (2,3) in the odd ones
                    dpe[0] = (e & 0x100u) ? __uint_as_float(rd[i0]) : 0.f;
                    dpe[1] = (e & 0x1000000u) ? __uint_as_float(rd[i0 + 1]) : 0.f;
                    dpe[2] = (o & 0x100u) ? __uint_as_float(rd[i0 + 2]) : 0.f;
                    dpe[3] = (o & 0x1000000u) ? __uint_as_float(rd[i0 + 3]) : 0.f;
                    ptx::add2(dpe[0], dpe[1], dpe[0], dpe[1], ndelta, ndelta);
                    ptx::add2(dpe[2], dpe[3], dpe[2], dpe[3], ndelta, ndelta);
                    ptx::mul2(ds[0], ds[1], ptx::ex2_approx(a[0]), ptx::ex2_approx(a[1]), dpe[0], dpe[1]);
                    ptx::mul2(ds[2], ds[3], ptx::ex2_approx(a[2]), ptx::ex2_approx(a[3]), dpe[2], dpe[3]);
                    pk[2 * g4] = pack2(ds[0], ds[1]);
                    pk[2 * g4 + 1] = pack2(ds[2], ds[3]);
                }
                if (kvalid < 64) {                          // last key tile of the clip: columns beyond it contribute nothing
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        if (2 * i >= kvalid) pk[i] = 0u;
                        else if (2 * i + 1 >= kvalid) pk[i] &= 0x0000FFFFu;
                    }
                }
                if (j > 0) {
                    ptx::mbar_wait(dq_done, (g - 1) & 1u);   // the previous dS has been consumed
                    ptx::tc_fence_after();
                }
                ptx::tmem_st_32x32b_x32(t_ds, reinterpret_cast<const uint32_t(&)[32]>(pk[0]));
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
                ptx::mbar_arrive(ds_full);
            }
            // epilogue (warpgroup 0; warpgroup 1 moves on, but the next unit's first dQ MMA needs ALL 256 ds_full
            // arrivals, i.e. it comes after these loads of the accumulator)
            ptx::mbar_wait(dq_done, (g - 1) & 1u);
            ptx::tc_fence_after();
            if (ch == 0) {
                uint32_t o[HDN];
#pragma unroll
                for (int c = 0; c < HDN / 32; ++c)
                    ptx::tmem_ld_32x32b_x32(t_dq + c * 32, reinterpret_cast<uint32_t(&)[32]>(o[c * 32]));
                ptx::tmem_ld_wait();
                if (row_ok) {
                    const float f = scale * inv_keep;
                    __nv_bfloat16* dst = dq + static_cast<size_t>(qrow0 + row) * lddq + head * hd;
#pragma unroll
                    for (int c = 0; c < HDN / 8; ++c) {
                        if (c * 8 < hd) {
                            uint4 v4;
                            v4.x = pack2(__uint_as_float(o[c * 8]) * f, __uint_as_float(o[c * 8 + 1]) * f);
                            v4.y = pack2(__uint_as_float(o[c * 8 + 2]) * f, __uint_as_float(o[c * 8 + 3]) * f);
                            v4.z = pack2(__uint_as_float(o[c * 8 + 4]) * f, __uint_as_float(o[c * 8 + 5]) * f);
                            v4.w = pack2(__uint_as_float(o[c * 8 + 6]) * f, __uint_as_float(o[c * 8 + 7]) * f);
                            *reinterpret_cast<uint4*>(dst + c * 8) = v4;
                        }
                    }
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, BWD_TMEM_COLS);
    }
}

// ================================================================================================================
// dK / dV
// ================================================================================================================
constexpr int KV_COL_S = 0, KV_COL_DP = 128, KV_COL_P = 256, KV_COL_DS = 320, KV_COL_DV = 384, KV_COL_DK = 448;
constexpr int KV_QST = 3;   // Q_i / dO_i feed the S-type products (early) AND the accumulating ones (late): 3 ring stages
constexpr int KV_AUX_BYTES = KV_QST * 128 * 4 * 6;   // per ring stage: lse2, delta, 4 dropout seed rows (2 key blocks x 2 h)
constexpr int KV_SMEM_BYTES = TILE_BYTES * (4 + 2 * KV_QST) + KV_AUX_BYTES + 256 + 1024;   // K[2], V[2], Q[3], dO[3]

template <int HDN>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_tc_dkv_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                   const __grid_constant__ CUtensorMap tv, const __grid_constant__ CUtensorMap tdo,
                   const int32_t* __restrict__ seq_off, const int32_t* __restrict__ blk_seq,
                   const int32_t* __restrict__ blk_row0, int n_units, int n_heads, int hd, float scale,
                   const float* __restrict__ lse, const float* __restrict__ delta, __nv_bfloat16* __restrict__ dk, int lddk,
                   __nv_bfloat16* __restrict__ dv, int lddv, float drop_p, unsigned long long seed) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* Ks = smem;                           // [2]
    uint8_t* Vs = Ks + 2 * TILE_BYTES;            // [2]
    uint8_t* Qs = Vs + 2 * TILE_BYTES;            // [KV_QST]
    uint8_t* dOs = Qs + KV_QST * TILE_BYTES;      // [KV_QST]
    float* aux = reinterpret_cast<float*>(dOs + KV_QST * TILE_BYTES);   // [KV_QST stages][6][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(aux) + KV_AUX_BYTES);
    uint64_t* kv_full = bars;                // [2]
    uint64_t* kv_empty = kv_full + 2;        // [2]
    uint64_t* qo_full = kv_empty + 2;        // [KV_QST]
    uint64_t* qo_empty = qo_full + KV_QST;   // [KV_QST]
    uint64_t* sdp_full = qo_empty + KV_QST;
    uint64_t* s_free = sdp_full + 1;
    uint64_t* pds_full = s_free + 1;
    uint64_t* acc_done = pds_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t thr = adrop::thr8_of(drop_p);
    const float inv_keep = thr ? adrop::inv_keep_of(thr) : 1.f;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tq);
        ptx::prefetch_tmap(&tk);
        ptx::prefetch_tmap(&tv);
        ptx::prefetch_tmap(&tdo);
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&kv_full[i], 1);
            ptx::mbar_init(&kv_empty[i], 1);
        }
        for (int i = 0; i < KV_QST; ++i) {
            ptx::mbar_init(&qo_full[i], 33);      // expect_tx arrival of lane 0 + 32 lanes that staged lse / delta / seeds
            ptx::mbar_init(&qo_empty[i], 257);    // tcgen05.commit of the accumulating MMAs + 256 softmax threads
        }
        ptx::mbar_init(sdp_full, 1);
        ptx::mbar_init(s_free, 256);
        ptx::mbar_init(pds_full, 256);
        ptx::mbar_init(acc_done, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, BWD_TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================ producer: TMA tiles + per-query scalars ================================
        uint32_t uc = 0, qc = 0;                         // units / query tiles issued so far
        for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++uc) {
            const int blk = u / n_heads, head = u - blk * n_heads;
            const int seq = blk_seq[blk];
            const int s0 = seq_off[seq], s1 = seq_off[seq + 1];
            const int krow0 = blk_row0[blk];
            const int nqb = (s1 - s0 + BQ - 1) / BQ;
            const uint32_t ks = uc & 1u;
            ptx::mbar_wait(&kv_empty[ks], ((uc >> 1) & 1u) ^ 1u);
            if (lane == 0) {
                ptx::mbar_expect_tx(&kv_full[ks], 2 * TILE_BYTES);
                ptx::tma_load_3d(Ks + ks * TILE_BYTES, &tk, &kv_full[ks], 0, krow0, head);
                ptx::tma_load_3d(Vs + ks * TILE_BYTES, &tv, &kv_full[ks], 0, krow0, head);
            }
            const uint32_t kb64 = static_cast<uint32_t>(krow0 - s0) >> 6;        // first 64-key block of this tile
            for (int i = 0; i < nqb; ++i, ++qc) {
                const uint32_t st = qc % KV_QST, ph = (qc / KV_QST) & 1u;
                ptx::mbar_wait(&qo_empty[st], ph ^ 1u);
                const int qb = s0 + i * BQ;
                if (lane == 0) {
                    ptx::mbar_expect_tx(&qo_full[st], 2 * TILE_BYTES);
                    ptx::tma_load_3d(Qs + st * TILE_BYTES, &tq, &qo_full[st], 0, qb, head);
                    ptx::tma_load_3d(dOs + st * TILE_BYTES, &tdo, &qo_full[st], 0, qb, head);
                }
                float* a = aux + st * 6 * 128;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int c = r * 32 + lane, q = qb + c;
                    const bool ok = q < s1;
                    const size_t gi = static_cast<size_t>(ok ? q : s0) * n_heads + head;
                    a[c] = ok ? lse[gi] * LOG2E : INFINITY;                       // invalid query columns: P = 0
                    a[128 + c] = ok ? delta[gi] / inv_keep : 0.f;
                    if (thr) {
                        const uint32_t rk = adrop::row_key(seed, q, head);
                        const uint32_t sa = adrop::stream_seed(rk, kb64), sb = adrop::stream_seed(rk, kb64 + 1u);
                        uint32_t* w = reinterpret_cast<uint32_t*>(a);
                        w[256 + c] = sa;                    // key block 0, h = 0
                        w[384 + c] = sa + adrop::DELTA;     // key block 0, h = 1
                        w[512 + c] = sb;                    // key block 1, h = 0
                        w[640 + c] = sb + adrop::DELTA;     // key block 1, h = 1
                    }
                }
                ptx::mbar_arrive(&qo_full[st]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = ptx::make_idesc_bf16(BKV, BQ, 0, 0);
            constexpr uint32_t idesc_a = ptx::make_idesc_bf16(BKV, HDN, 0, 1);
            const int ks_s = (hd + 15) >> 4;
            const uint32_t t_s = tmem_base + KV_COL_S, t_dp = tmem_base + KV_COL_DP, t_p = tmem_base + KV_COL_P,
                           t_ds = tmem_base + KV_COL_DS, t_dv = tmem_base + KV_COL_DV, t_dk = tmem_base + KV_COL_DK;
            uint32_t uc = 0, g0 = 0;
            for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++uc) {
                const int blk = u / n_heads;
                const int seq = blk_seq[blk];
                const int s0 = seq_off[seq], s1 = seq_off[seq + 1];
                const int nqb = (s1 - s0 + BQ - 1) / BQ;
                const uint32_t ks = uc & 1u;
                const uint32_t ka = ptx::smem_u32(Ks + ks * TILE_BYTES), va = ptx::smem_u32(Vs + ks * TILE_BYTES);
                auto issue_sdp = [&](int i) {
                    const uint32_t g = g0 + i, st = g % KV_QST;
                    ptx::mbar_wait(&qo_full[st], (g / KV_QST) & 1u);
                    if (g > 0) ptx::mbar_wait(s_free, (g - 1) & 1u);
                    ptx::tc_fence_after();
                    const uint32_t qb = ptx::smem_u32(Qs + st * TILE_BYTES), ob = ptx::smem_u32(dOs + st * TILE_BYTES);
                    for (int k = 0; k < ks_s; ++k)
                        ptx::umma_bf16(t_s, ptx::make_smem_desc_sw128(ka + k * 32, 16, 1024),
                                       ptx::make_smem_desc_sw128(qb + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
                    for (int k = 0; k < ks_s; ++k)
                        ptx::umma_bf16(t_dp, ptx::make_smem_desc_sw128(va + k * 32, 16, 1024),
                                       ptx::make_smem_desc_sw128(ob + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
                    if (i == nqb - 1) ptx::umma_commit(&kv_empty[ks]);     // last products that read this unit's K / V
                    ptx::umma_commit(sdp_full);
                };
                ptx::mbar_wait(&kv_full[ks], (uc >> 1) & 1u);
                issue_sdp(0);
                for (int i = 0; i < nqb; ++i) {
                    if (i + 1 < nqb) issue_sdp(i + 1);
                    const uint32_t g = g0 + i, st = g % KV_QST;
                    const int qvalid = min(BQ, s1 - (s0 + i * BQ));
                    const int ks_o = (qvalid + 15) >> 4;
                    ptx::mbar_wait(pds_full, g & 1u);
                    ptx::tc_fence_after();
                    const uint32_t qb = ptx::smem_u32(Qs + st * TILE_BYTES), ob = ptx::smem_u32(dOs + st * TILE_BYTES);
                    for (int k = 0; k < ks_o; ++k)
                        ptx::umma_bf16_ts(t_dv, t_p + k * 8, ptx::make_smem_desc_sw128(ob + k * 2048, 8192, 1024), idesc_a,
                                          (i != 0 || k != 0) ? 1u : 0u);
                    for (int k = 0; k < ks_o; ++k)
                        ptx::umma_bf16_ts(t_dk, t_ds + k * 8, ptx::make_smem_desc_sw128(qb + k * 2048, 8192, 1024), idesc_a,
                                          (i != 0 || k != 0) ? 1u : 0u);
                    ptx::umma_commit(&qo_empty[st]);
                    ptx::umma_commit(acc_done);
                }
                g0 += nqb;
            }
        }
    } else {
        // ================================ thread = key row, warpgroup = query half ================================
        const int quad = warp & 3;
        const int ch = (warp - 2) >> 2;
        const int row = quad * 32 + lane;                 // key row of the tile
        const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
        const uint32_t t_s = tmem_base + lane_addr + KV_COL_S + ch * 64, t_dp = tmem_base + lane_addr + KV_COL_DP + ch * 64,
                       t_p = tmem_base + lane_addr + KV_COL_P + ch * 32, t_ds = tmem_base + lane_addr + KV_COL_DS + ch * 32,
                       t_dv = tmem_base + lane_addr + KV_COL_DV, t_dk = tmem_base + lane_addr + KV_COL_DK;
        const float scale_log2 = scale * LOG2E;
        // this key's position in its 64-key block: LCG jump (n), interleaved stream (h) and byte of the draw
        const int k64 = row & 63, n = k64 >> 3;
        uint32_t la = adrop::lcg_a(0), lc = adrop::lcg_c(0);
#pragma unroll
        for (int i = 1; i < 8; ++i)
            if (n == i) { la = adrop::lcg_a(i); lc = adrop::lcg_c(i); }
        const int sh = 8 * (((k64 & 1) << 1) | ((k64 >> 1) & 1));
        const uint32_t byte_mask = 0xFFu << sh, thr_sh = thr << sh;
        const int seed_row = 256 + ((row >> 6) * 2 + ((k64 >> 2) & 1)) * 128;      // which of the 4 staged seed rows
        uint32_t g = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const int blk = u / n_heads, head = u - blk * n_heads;
            const int seq = blk_seq[blk];
            const int s0 = seq_off[seq], s1 = seq_off[seq + 1];
            const int krow0 = blk_row0[blk];
            const int krows = min(BKV, s1 - krow0);
            const int nqb = (s1 - s0 + BQ - 1) / BQ;
            for (int i = 0; i < nqb; ++i, ++g) {
                const uint32_t st = g % KV_QST;
                uint32_t rs[64], rd[64];
                ptx::mbar_wait(sdp_full, g & 1u);
                ptx::tc_fence_after();
                ptx::tmem_ld_32x32b_x32(t_s, reinterpret_cast<uint32_t(&)[32]>(rs[0]));
                ptx::tmem_ld_32x32b_x32(t_s + 32, reinterpret_cast<uint32_t(&)[32]>(rs[32]));
                ptx::tmem_ld_32x32b_x32(t_dp, reinterpret_cast<uint32_t(&)[32]>(rd[0]));
                ptx::tmem_ld_32x32b_x32(t_dp + 32, reinterpret_cast<uint32_t(&)[32]>(rd[32]));
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                ptx::mbar_arrive(s_free);
                ptx::mbar_wait(&qo_full[st], (g / KV_QST) & 1u);                   // lse / delta / seeds of this tile
                const float* a = aux + st * 6 * 128 + ch * 64;
                const uint32_t* sdrow = reinterpret_cast<const uint32_t*>(aux + st * 6 * 128) + seed_row + ch * 64;
                uint32_t pp[32], pd[32];
#pragma unroll
                for (int g4 = 0; g4 < 16; ++g4) {
                    const float4 l4 = *reinterpret_cast<const float4*>(a + 4 * g4);
                    const float4 d4 = *reinterpret_cast<const float4*>(a + 128 + 4 * g4);
                    uint4 s4 = make_uint4(0u, 0u, 0u, 0u);
                    if (thr) s4 = *reinterpret_cast<const uint4*>(sdrow + 4 * g4);
                    const float lv[4] = {l4.x, l4.y, l4.z, l4.w}, dl[4] = {d4.x, d4.y, d4.z, d4.w};
                    const uint32_t sv[4] = {s4.x, s4.y, s4.z, s4.w};
                    float p[4], ds[4], a[4], pr[4], dpe[4];
                    const int c0 = 4 * g4;
                    ptx::fma2(a[0], a[1], __uint_as_float(rs[c0]), __uint_as_float(rs[c0 + 1]), scale_log2, scale_log2, -lv[0], -lv[1]);
                    ptx::fma2(a[2], a[3], __uint_as_float(rs[c0 + 2]), __uint_as_float(rs[c0 + 3]), scale_log2, scale_log2, -lv[2], -lv[3]);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        pr[c] = ptx::ex2_approx(a[c]);
                        bool kept = true;
                        if (thr) {
                            const uint32_t sx = sv[c] * la + lc;
                            kept = ((sx ^ (sx >> 16)) & byte_mask) >= thr_sh;
                        }
                        p[c] = kept ? pr[c] : 0.f;
                        dpe[c] = kept ? __uint_as_float(rd[c0 + c]) : 0.f;
                    }
                    ptx::add2(dpe[0], dpe[1], dpe[0], dpe[1], -dl[0], -dl[1]);
                    ptx::add2(dpe[2], dpe[3], dpe[2], dpe[3], -dl[2], -dl[3]);
                    ptx::mul2(ds[0], ds[1], pr[0], pr[1], dpe[0], dpe[1]);
                    ptx::mul2(ds[2], ds[3], pr[2], pr[3], dpe[2], dpe[3]);
                    pp[2 * g4] = pack2(p[0], p[1]);
                    pp[2 * g4 + 1] = pack2(p[2], p[3]);
                    pd[2 * g4] = pack2(ds[0], ds[1]);
                    pd[2 * g4 + 1] = pack2(ds[2], ds[3]);
                }
                ptx::mbar_arrive(&qo_empty[st]);                                   // the staged scalars have been read
                if (i > 0) {
                    ptx::mbar_wait(acc_done, (g - 1) & 1u);                        // previous P~ / dS consumed
                    ptx::tc_fence_after();
                }
                ptx::tmem_st_32x32b_x32(t_p, reinterpret_cast<const uint32_t(&)[32]>(pp[0]));
                ptx::tmem_st_32x32b_x32(t_ds, reinterpret_cast<const uint32_t(&)[32]>(pd[0]));
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
                ptx::mbar_arrive(pds_full);
            }
            ptx::mbar_wait(acc_done, (g - 1) & 1u);
            ptx::tc_fence_after();
            // warpgroup 0 stores dV, warpgroup 1 stores dK (the next unit's first accumulating MMA needs all 256
            // pds_full arrivals, so it comes after both warpgroups have read their accumulator)
            uint32_t o[HDN];
            const uint32_t t_acc = ch == 0 ? t_dv : t_dk;
#pragma unroll
            for (int c = 0; c < HDN / 32; ++c)
                ptx::tmem_ld_32x32b_x32(t_acc + c * 32, reinterpret_cast<uint32_t(&)[32]>(o[c * 32]));
            ptx::tmem_ld_wait();
            if (row < krows) {
                const float f = ch == 0 ? inv_keep : scale * inv_keep;
                __nv_bfloat16* dst = (ch == 0 ? dv + static_cast<size_t>(krow0 + row) * lddv
                                              : dk + static_cast<size_t>(krow0 + row) * lddk) + head * hd;
#pragma unroll
                for (int c = 0; c < HDN / 8; ++c) {
                    if (c * 8 < hd) {
                        uint4 v4;
                        v4.x = pack2(__uint_as_float(o[c * 8]) * f, __uint_as_float(o[c * 8 + 1]) * f);
                        v4.y = pack2(__uint_as_float(o[c * 8 + 2]) * f, __uint_as_float(o[c * 8 + 3]) * f);
                        v4.z = pack2(__uint_as_float(o[c * 8 + 4]) * f, __uint_as_float(o[c * 8 + 5]) * f);
                        v4.w = pack2(__uint_as_float(o[c * 8 + 6]) * f, __uint_as_float(o[c * 8 + 7]) * f);
                        *reinterpret_cast<uint4*>(dst + c * 8) = v4;
                    }
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, BWD_TMEM_COLS);
    }
}

}  // namespace atc
}  // namespace vsgg

using namespace vsgg;

extern "C" int b200vsgg_attn_tc_bwd(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv,
                                    const void* ctx, int32_t ldc, const void* dctx, int32_t lddc, const float* lse,
                                    float* delta, int32_t rows, const int32_t* seq_off, const int32_t* blk_seq,
                                    const int32_t* blk_row0, int32_t n_blocks, int32_t n_heads, int32_t head_dim, float scale,
                                    void* dq, int32_t lddq, void* dk, int32_t lddk, void* dv, int32_t lddv, float drop_p,
                                    uint64_t seed, void* stream) {
    if (!q || !k || !v || !ctx || !dctx || !lse || !delta || !seq_off || !blk_seq || !blk_row0 || !dq || !dk || !dv ||
        rows <= 0 || n_heads <= 0 || n_heads > 64 || head_dim < 8 || head_dim > 64 || (head_dim & 7) || (lddq & 7) ||
        (lddk & 7) || (lddv & 7) || (ldc & 7) || (lddc & 7))
        return set_error(B200VSGG_ERR_BAD_ARG, "attn_tc_bwd: bad arg (head_dim in 8..64, % 8 == 0; <= 64 heads)");
    for (const void* p : {static_cast<const void*>(dq), static_cast<const void*>(dk), static_cast<const void*>(dv), ctx, dctx})
        if (reinterpret_cast<uintptr_t>(p) & 15u) return set_error(B200VSGG_ERR_BAD_ARG, "attn_tc_bwd: 16-byte alignment");
    if (n_blocks == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    CUtensorMap tq, tk, tv, tdo;
    int rc;
    if ((rc = atc::make_tmap_heads(&tq, q, head_dim, rows, n_heads, ldq))) return rc;
    if ((rc = atc::make_tmap_heads(&tk, k, head_dim, rows, n_heads, ldk))) return rc;
    if ((rc = atc::make_tmap_heads(&tv, v, head_dim, rows, n_heads, ldv))) return rc;
    if ((rc = atc::make_tmap_heads(&tdo, dctx, head_dim, rows, n_heads, lddc))) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(atc::attn_tc_dq_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, atc::DQ_SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(atc::attn_tc_dq_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, atc::DQ_SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(atc::attn_tc_dkv_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, atc::KV_SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(atc::attn_tc_dkv_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, atc::KV_SMEM_BYTES);
        if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
        attr_set = true;
    }
    const long long total = static_cast<long long>(rows) * n_heads;
    atc::attn_delta_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(
        reinterpret_cast<const __nv_bfloat16*>(ctx), ldc, reinterpret_cast<const __nv_bfloat16*>(dctx), lddc, rows, n_heads,
        head_dim, delta);
    const long long units = static_cast<long long>(n_blocks) * n_heads;
    if (units > 0x7fffffffLL) return set_error(B200VSGG_ERR_BAD_ARG, "attn_tc_bwd: too many (tile, head) units");
    // persistent: one CTA per SM walks the unit list (B200VSGG_ATTN_PERSIST=0: one unit per CTA, for A/B timing)
    static const bool persist = []() { const char* e = getenv("B200VSGG_ATTN_PERSIST"); return !(e && e[0] == '0'); }();
    const unsigned grid = static_cast<unsigned>((!persist || units < num_sms()) ? units : num_sms());
    const int n_units = static_cast<int>(units);
    if (head_dim <= 32) {
        atc::attn_tc_dq_kernel<32><<<grid, atc::BWD_THREADS, atc::DQ_SMEM_BYTES, s>>>(
            tq, tk, tv, tdo, seq_off, blk_seq, blk_row0, n_units, n_heads, head_dim, scale, lse, delta,
            reinterpret_cast<__nv_bfloat16*>(dq), lddq, drop_p, seed);
        atc::attn_tc_dkv_kernel<32><<<grid, atc::BWD_THREADS, atc::KV_SMEM_BYTES, s>>>(
            tq, tk, tv, tdo, seq_off, blk_seq, blk_row0, n_units, n_heads, head_dim, scale, lse, delta,
            reinterpret_cast<__nv_bfloat16*>(dk), lddk, reinterpret_cast<__nv_bfloat16*>(dv), lddv, drop_p, seed);
    } else {
        atc::attn_tc_dq_kernel<64><<<grid, atc::BWD_THREADS, atc::DQ_SMEM_BYTES, s>>>(
            tq, tk, tv, tdo, seq_off, blk_seq, blk_row0, n_units, n_heads, head_dim, scale, lse, delta,
            reinterpret_cast<__nv_bfloat16*>(dq), lddq, drop_p, seed);
        atc::attn_tc_dkv_kernel<64><<<grid, atc::BWD_THREADS, atc::KV_SMEM_BYTES, s>>>(
            tq, tk, tv, tdo, seq_off, blk_seq, blk_row0, n_units, n_heads, head_dim, scale, lse, delta,
            reinterpret_cast<__nv_bfloat16*>(dk), lddk, reinterpret_cast<__nv_bfloat16*>(dv), lddv, drop_p, seed);
    }
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}
