// Persistent warp-specialised bf16 GEMM for sm_100a:
//   TMA (cp.async.bulk.tensor, SWIZZLE_128B) -> smem ring -> tcgen05.mma (cta_group::1, M=128,
//   N=BN, K=16) accumulating fp32 in TMEM (two accumulator stages) -> tcgen05.ld epilogue with
//   fused bias / activation / mask / dropout / residual / dual fp32+bf16 stores.
// Roles per CTA (384 threads): warp0 = TMA producer, warp1 = MMA issuer, warp2 = TMEM allocator,
//   warps4-11 = epilogue (warp w owns TMEM lanes 32*(w%4)..+31 and one column half of the tile; each
//   32x32 chunk is transposed through shared memory so every global access is row-contiguous).
// Operand majors: K-major ([rows,K], K contiguous) or MN-major ([K,rows], rows contiguous), so the
#include "gemm_common.cuh"
#include <cstdlib>

namespace vsgg {

template <int BN, int A_MN, int B_MN>
__global__ void __launch_bounds__(384, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_d,
                 const GemmEpi ep, const int M, const int N, const int K, const int splits, const int kb_per,
                 const int a_k_period) {
    // splits > 1: split-K.  Work unit u -> (tile = u % num_tiles, split = u / num_tiles); split s reduces
    // k-blocks [s*kb_per, min(num_kb, (s+1)*kb_per)) and its epilogue atomically adds into out_f32.
    using Cfg = GemmCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
    uint8_t* smem = smem_raw + pad;  // 1024-B aligned (SWIZZLE_128B atoms)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint64_t* tmem_empty_bar = tmem_full_bar + Cfg::ACC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + Cfg::ACC_STAGES);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int num_m = (M + BM - 1) / BM;
    const int num_n = (N + BN - 1) / BN;
    const int num_tiles = num_m * num_n;
    const int num_kb = (K + BK - 1) / BK;
    const int num_units = num_tiles * splits;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tma_a);
        ptx::prefetch_tmap(&tma_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            ptx::mbar_init(&full_bar[i], 1);
            ptx::mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < Cfg::ACC_STAGES; ++i) {
            ptx::mbar_init(&tmem_full_bar[i], 1);
            ptx::mbar_init(&tmem_empty_bar[i], 256);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Warpgroup register reallocation: warps 0-3 (TMA producer, MMA issuer, TMEM allocator, idle) need almost no
    // registers; the two epilogue warpgroups take them over (384 x 168 at launch -> 128 x 40 + 256 x 232), so the
    // epilogue's three 32-element register arrays and its state no longer spill to local memory.
    if (warp < 4) {
    ptx::setmaxnreg_dec<40>();
    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
                const int tile = unit % num_tiles, split = unit / num_tiles;
                const int m0 = (tile / num_n) * BM;
                const int n0 = (tile % num_n) * BN;
                const int kb0 = split * kb_per;
                const int kb1 = min(num_kb, kb0 + kb_per);
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
                    uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                    uint8_t* sb = sa + Cfg::A_BYTES;
                    ptx::mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                    if (A_MN == 0) {
                        // a_k_period > 0: A repeats along K with that period (split-precision weights
                        // [W_hi | W_lo] against one copy of the activations)
                        const int ka = a_k_period > 0 ? (kb * BK) % a_k_period : kb * BK;
                        ptx::tma_load_2d(sa, &tma_a, &full_bar[stage], ka, m0);
                    } else {
#pragma unroll
                        for (int c = 0; c < BM / 64; ++c)
                            ptx::tma_load_2d(sa + c * (64 * BK * 2), &tma_a, &full_bar[stage], m0 + c * 64, kb * BK);
                    }
                    if (B_MN == 0) {
                        ptx::tma_load_2d(sb, &tma_b, &full_bar[stage], kb * BK, n0);
                    } else {
#pragma unroll
                        for (int c = 0; c < BN / 64; ++c)
                            ptx::tma_load_2d(sb + c * (64 * BK * 2), &tma_b, &full_bar[stage], n0 + c * 64, kb * BK);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer (one thread) ================================
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, BN, A_MN, B_MN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
                const int split = unit / num_tiles;
                const int kb0 = split * kb_per;
                const int kb1 = min(num_kb, kb0 + kb_per);
                ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        // K-major: advance 16 bf16 = 32 B inside the 128-B swizzle row.
                        // MN-major: advance 16 K-rows of 128 B = 2048 B.
                        const uint64_t da = A_MN ? ptx::make_smem_desc_sw128(sa + k * 2048, 64 * BK * 2, 1024)
                                                 : ptx::make_smem_desc_sw128(sa + k * 32, 16, 1024);
                        const uint64_t db = B_MN ? ptx::make_smem_desc_sw128(sb + k * 2048, 64 * BK * 2, 1024)
                                                 : ptx::make_smem_desc_sw128(sb + k * 32, 16, 1024);
                        ptx::umma_bf16(tmem_d, da, db, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
                    }
                    ptx::umma_commit(&empty_bar[stage]);  // frees the smem slot when the MMAs retire
                    if (kb == kb1 - 1) ptx::umma_commit(&tmem_full_bar[acc]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                if (++acc == Cfg::ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
            }
        }
    }
    } else {
        ptx::setmaxnreg_inc<232>();
        // ================================ epilogue warps (8) ================================
        // Warp (4 + e): TMEM lane quadrant wq = warp % 4 (rows m0 + 32*wq ..), column half e / 4 of the tile.
        // Per 32-column chunk: tcgen05.ld gives each lane ONE ROW x 32 columns; the chunk is transposed
        // through a swizzled 4 KB shared-memory buffer so that global traffic is row-contiguous
        // (lane = column: 128 B per fp32 row access, one transaction) instead of 32 rows per instruction.
        // Residual / mask rows of the chunk are prefetched into registers before the accumulator is read.
        const int ew = warp - 4;
        const int wq = warp & 3;
        const int half = ew >> 2;
        EpiRegs E;
        E.stg = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::BAR_BYTES + ew * 4096);
        E.inv_keep = ep.dropout_p > 0.f ? 1.0f / (1.0f - ep.dropout_p) : 1.0f;
        E.drop_thr = ep.dropout_p > 0.f ? static_cast<uint32_t>(ep.dropout_p * 4294967296.0) : 0u;
        E.res_f32 = ep.residual_is_bf16 ? nullptr : reinterpret_cast<const float*>(ep.residual);
        E.res_b16 = ep.residual_is_bf16 ? reinterpret_cast<const __nv_bfloat16*>(ep.residual) : nullptr;
        E.mask_src = ep.mask_src;
        E.bias = ep.bias;
        E.act = ep.act; E.mask_mode = ep.mask_mode; E.accumulate = ep.accumulate;
        E.alpha = ep.alpha;
        E.drop_seed = ep.dropout_seed;
        E.out_f32 = ep.out_f32; E.out_bf16 = ep.out_bf16;
        E.ld_f32 = ep.ld_f32; E.ld_bf16 = ep.ld_bf16; E.ldr = ep.ldr; E.ldm = ep.ldm;
        E.N = N; E.lane = lane; E.atomic = splits > 1;
        int acc = 0;
        uint32_t acc_phase = 0;
        int issued = 0;                                   // TMA stores issued by this warp (epi_chunk_tma)
        for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
            const int tile = unit % num_tiles, split = unit / num_tiles;
            const int m0 = (tile / num_n) * BM;
            const int n0 = (tile % num_n) * BN;
            const int rbase = m0 + wq * 32;
            const int rows_here = min(32, M - rbase);      // <= 0: nothing to store for this warp
            E.use_bias = ep.bias != nullptr && (splits == 1 || split == 0);
            bool waited = false;
            if (rows_here > 0 && ep.tma_store) {   // split-K partial tiles: TMA reduce-add instead of per-element atomics
#pragma unroll 1
                for (int c = 0; c < BN / 64; ++c) {
                    const int nc = n0 + half * (BN / 2) + c * 32;
                    if (nc >= N) break;  // warp-uniform
                    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) +
                                           static_cast<uint32_t>(acc * BN + half * (BN / 2) + c * 32);
                    epi_chunk_tma(E, &tma_c, &tma_d, taddr, rbase, M, nc, reinterpret_cast<uint8_t*>(E.stg), issued,
                                  &tmem_full_bar[acc], acc_phase, waited);
                }
            } else if (rows_here > 0) {
#pragma unroll 1
                for (int c = 0; c < BN / 64; ++c) {
                    const int nc = n0 + half * (BN / 2) + c * 32;
                    if (nc >= N) break;  // warp-uniform
                    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) +
                                           static_cast<uint32_t>(acc * BN + half * (BN / 2) + c * 32);
                    if (rows_here == 32)
                        epi_chunk<true>(E, taddr, rbase, 32, nc, &tmem_full_bar[acc], acc_phase, waited);
                    else
                        epi_chunk<false>(E, taddr, rbase, rows_here, nc, &tmem_full_bar[acc], acc_phase, waited);
                }
            }
            if (!waited) {
                ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
                ptx::tc_fence_after();
            }
            ptx::tc_fence_before();
            ptx::mbar_arrive(&tmem_empty_bar[acc]);
            if (++acc == Cfg::ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
        }
        if (issued > 0 && lane == 0) ptx::bulk_wait_all();   // staging boxes stay valid until every store has drained
    }

    // ================================ teardown ================================
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
template <int BN, int A_MN, int B_MN>
static int launch_gemm(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const GemmEpi& ep,
                       int split_k, int a_k_period, cudaStream_t stream) {
    using Cfg = GemmCfg<BN>;
    CUtensorMap ta, tb;
    int rc;
    if (a_k_period < 0 || (a_k_period > 0 && (A_MN != 0 || a_k_period % BK != 0 || K % a_k_period != 0)))
        return set_error(B200VSGG_ERR_BAD_ARG, "gemm: a_k_period needs K-major A, period % 64 == 0, K % period == 0");
    if (A_MN == 0) rc = make_tmap_bf16(&ta, A, a_k_period > 0 ? a_k_period : K, M, lda, BK, BM);
    else rc = make_tmap_bf16(&ta, A, M, K, lda, 64, BK);
    if (rc) return rc;
    if (B_MN == 0) rc = make_tmap_bf16(&tb, B, K, N, ldb, BK, BN);
    else rc = make_tmap_bf16(&tb, B, N, K, ldb, 64, BK);
    if (rc) return rc;
    CUtensorMap tc = ta, td = ta;                          // placeholders when the TMA-store epilogue is off
    if (ep.tma_store && ep.out_bf16 != nullptr && (rc = make_tmap_out(&tc, ep.out_bf16, false, N, M, ep.ld_bf16))) return rc;
    if (ep.tma_store && ep.out_f32 != nullptr && (rc = make_tmap_out(&td, ep.out_f32, true, N, M, ep.ld_f32))) return rc;

    auto kern = gemm_bf16_kernel<BN, A_MN, B_MN>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
        attr_set = true;
    }
    const int num_tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    const int num_kb = (K + BK - 1) / BK;
    // Split-K when the tile count cannot fill the machine and K is long (weight gradients whose
    // reduction dimension is the token count): plain fp32 output only.
    int splits = 1;
    const bool can_split = ep.out_bf16 == nullptr && ep.act == 0 && ep.mask_src == nullptr && ep.residual == nullptr &&
                           ep.dropout_p == 0.f && ep.out_f32 != nullptr;
    if (split_k != 1 && can_split && num_tiles * 2 <= num_sms() && num_kb >= 32) {
        splits = split_k > 1 ? split_k : (2 * num_sms() + num_tiles - 1) / num_tiles;
        const int max_splits = num_kb / 8;
        if (splits > max_splits) splits = max_splits;
        if (splits < 1) splits = 1;
    }
    int kb_per = (num_kb + splits - 1) / splits;
    splits = (num_kb + kb_per - 1) / kb_per;  // no empty split
    if (splits > 1 && !ep.accumulate) {
        cudaError_t e = cudaMemset2DAsync(ep.out_f32, static_cast<size_t>(ep.ld_f32) * sizeof(float), 0,
                                          static_cast<size_t>(N) * sizeof(float), M, stream);
        if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
    }
    const int num_units = num_tiles * splits;
    int grid = num_units < num_sms() ? num_units : num_sms();
    kern<<<grid, 384, Cfg::SMEM_BYTES, stream>>>(ta, tb, tc, td, ep, M, N, K, splits, kb_per, a_k_period);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
    return 0;
}

}  // namespace vsgg

namespace vsgg {
int gemm2_launch(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, int M, int N, int K,
                 const GemmEpi& ep, cudaStream_t stream);   // gemm2_tcgen05.cu
static int use_2cta() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B200VSGG_GEMM_2CTA");
        v = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return v;
}
}  // namespace vsgg

extern "C" int b200vsgg_gemm_bf16(const void* A, int32_t lda, int32_t a_mn, const void* B, int32_t ldb, int32_t b_mn,
                                  int32_t M, int32_t N, int32_t K, const b200vsgg_gemm_epilogue* e, void* stream) {
    using namespace vsgg;
    if (A == nullptr || B == nullptr || e == nullptr || M <= 0 || N <= 0 || K <= 0)
        return set_error(B200VSGG_ERR_BAD_ARG, "gemm: null operand or non-positive size");
    if (e->out_f32 == nullptr && e->out_bf16 == nullptr)
        return set_error(B200VSGG_ERR_BAD_ARG, "gemm: no output pointer");
    if (a_mn == 1 && b_mn == 0) return set_error(B200VSGG_ERR_BAD_ARG, "gemm: (a_mn=1,b_mn=0) not instantiated");
    GemmEpi ep;
    ep.bias = e->bias;
    ep.residual = e->residual;
    ep.residual_is_bf16 = e->residual_is_bf16;
    ep.ldr = e->ldr;
    ep.mask_src = reinterpret_cast<const __nv_bfloat16*>(e->mask_src);
    ep.ldm = e->ldm;
    ep.mask_mode = e->mask_mode;
    ep.act = e->act;
    ep.out_f32 = e->out_f32;
    ep.ld_f32 = e->ld_f32;
    ep.out_bf16 = reinterpret_cast<__nv_bfloat16*>(e->out_bf16);
    ep.ld_bf16 = e->ld_bf16;
    ep.accumulate = e->accumulate;
    ep.alpha = e->alpha;
    ep.dropout_p = e->dropout_p;
    ep.dropout_seed = e->dropout_seed;
    {
        auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
        bool ok = true;
        if (ep.bias) ok = ok && al16(ep.bias);
        if (ep.residual) ok = ok && al16(ep.residual) && (ep.ldr % (ep.residual_is_bf16 ? 8 : 4) == 0);
        if (ep.mask_src) ok = ok && al16(ep.mask_src) && (ep.ldm % 8 == 0);
        if (ep.out_f32) ok = ok && al16(ep.out_f32) && (ep.ld_f32 % 4 == 0);
        if (ep.out_bf16) ok = ok && al16(ep.out_bf16) && (ep.ld_bf16 % 8 == 0);
        ep.vec_ok = ok ? 1 : 0;
    }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    // TMA-store epilogue unless the call may split K (atomics), accumulates in place, or has residual AND mask
    static const bool tma_epi = []() { const char* v = getenv("B200VSGG_GEMM_TMA_EPI"); return !(v && v[0] == '0'); }();
    ep.tma_store = (tma_epi && !ep.accumulate && ep.vec_ok && (N & 7) == 0 &&
                    !(ep.residual != nullptr && ep.mask_src != nullptr)) ? 1 : 0;
    // Large problems: 256x256 tiles on CTA pairs (cta_group::2), 2/3 of the operand traffic per flop.
    if (use_2cta() && e->a_k_period == 0 && e->split_k <= 1) {
        const long long tiles2 = static_cast<long long>((M + 255) / 256) * ((N + 255) / 256);
        if (tiles2 >= 4LL * (num_sms() / 2) && N >= 1024 && K >= 512)
            return gemm2_launch(A, lda, a_mn, B, ldb, b_mn, M, N, K, ep, s);
    }
    // Tile-N choice: 256-wide tiles when N is large enough to keep the padding waste small.
    const bool wide = (N >= 1024) || (N % 256 == 0);
    if (a_mn == 0 && b_mn == 0)
        return wide ? launch_gemm<256, 0, 0>(A, lda, B, ldb, M, N, K, ep, e->split_k, e->a_k_period, s)
                    : launch_gemm<128, 0, 0>(A, lda, B, ldb, M, N, K, ep, e->split_k, e->a_k_period, s);
    if (a_mn == 0 && b_mn == 1)
        return wide ? launch_gemm<256, 0, 1>(A, lda, B, ldb, M, N, K, ep, e->split_k, e->a_k_period, s)
                    : launch_gemm<128, 0, 1>(A, lda, B, ldb, M, N, K, ep, e->split_k, e->a_k_period, s);
    return wide ? launch_gemm<256, 1, 1>(A, lda, B, ldb, M, N, K, ep, e->split_k, e->a_k_period, s)
                : launch_gemm<128, 1, 1>(A, lda, B, ldb, M, N, K, ep, e->split_k, e->a_k_period, s);
}
