// 2-CTA (cta_group::2) variant of the persistent bf16 GEMM: a thread-block cluster of two CTAs on one
// TPC computes a 256 (M) x 256 (N) tile.  CTA r holds rows [m0 + 128 r, +128) of A and rows
// [n0 + 128 r, +128) of B in its own shared memory (each CTA loads only HALF of the B tile: 32 KB per
// pipeline stage per CTA instead of 48 KB), the leader CTA's single MMA thread issues
// tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16) which reads both CTAs' operands and writes each
// CTA's 128 x 256 fp32 accumulator into its own TMEM.  Per flop this moves 2/3 of the L2 -> SMEM bytes
// of the 128 x 256 single-CTA tile — the single-CTA kernel is L2-feed-bound on the large shapes.
// Barriers: TMA of both CTAs completes on the LEADER's full barrier; tcgen05.commit multicasts the
// "slot free" / "accumulator ready" arrivals to both CTAs; both CTAs' epilogue threads arrive on the
// leader's "accumulator free" barrier (remote mbarrier.arrive for the peer).
// Epilogue identical to the single-CTA kernel (gemm_common.cuh).
#include "gemm_common.cuh"

namespace vsgg {
namespace g2 {

constexpr int BM2 = 256;        // cluster tile rows (128 per CTA)
constexpr int BN2 = 256;
constexpr int BK2 = 64;
constexpr int A_BYTES = 128 * BK2 * 2;
constexpr int B_BYTES = 128 * BK2 * 2;          // half of the B tile
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;  // per CTA
constexpr int STAGES = 6;
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = 512;
constexpr int BAR_BYTES = 1024;   // keeps the epilogue staging boxes 1024-byte aligned (swizzled TMA stores)
constexpr int EPI_BYTES = 8 * 4096;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + EPI_BYTES + 1024;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t local, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}
// 2D TMA load into THIS CTA's shared memory, completion (bytes) on a barrier that may live in the peer CTA
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const void* tmap, uint32_t bar_cluster_addr, int c0,
                                                int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
        "[%2];\n" ::"r"(ptx::smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(ptx::smem_u32(smem_slot)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive (once all prior MMAs of this thread completed) on the barrier at the same offset in BOTH CTAs
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
    const uint16_t mask = 0x3;
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
            ptx::smem_u32(bar)),
        "h"(mask)
        : "memory");
}

template <int A_MN, int B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                  const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_d,
                 const GemmEpi ep, const int M, const int N, const int K) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
    uint8_t* smem = smem_raw + pad;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint64_t* tmem_empty_bar = tmem_full_bar + ACC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + ACC_STAGES);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1;
    const int n_clusters = gridDim.x >> 1;

    const int num_m = (M + BM2 - 1) / BM2;
    const int num_n = (N + BN2 - 1) / BN2;
    const int num_tiles = num_m * num_n;
    const int num_kb = (K + BK2 - 1) / BK2;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tma_a);
        ptx::prefetch_tmap(&tma_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            ptx::mbar_init(&full_bar[i], 1);      // leader: its own expect_tx arrival (+ 64 KB of transactions)
            ptx::mbar_init(&empty_bar[i], 1);     // one multicast commit per phase
        }
        for (int i = 0; i < ACC_STAGES; ++i) {
            ptx::mbar_init(&tmem_full_bar[i], 1);
            ptx::mbar_init(&tmem_empty_bar[i], 512);   // leader: 256 epilogue threads of each CTA
        }
        ptx::fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc2(tmem_slot, TMEM_COLS);
        tmem_relinquish2();
    }
    ptx::tc_fence_before();
    __syncthreads();
    cluster_sync();                                    // both CTAs' barriers are initialised before any remote use
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Warpgroup register reallocation: warps 0-3 (TMA producer, MMA issuer, TMEM allocator, idle) need almost no
    // registers; the two epilogue warpgroups take them over (384 x 168 at launch -> 128 x 40 + 256 x 232), so the
    // epilogue's three 32-element register arrays and its state no longer spill to local memory.
    if (warp < 4) {
    ptx::setmaxnreg_dec<40>();
    if (warp == 0) {
        // ================================ TMA producer (both CTAs) ================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
                const int m0 = (tile / num_n) * BM2 + static_cast<int>(rank) * 128;
                const int n0 = (tile % num_n) * BN2 + static_cast<int>(rank) * 128;
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    uint8_t* sb = sa + A_BYTES;
                    const uint32_t bar = mapa(ptx::smem_u32(&full_bar[stage]), 0);
                    if (leader) ptx::mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
                    if (A_MN == 0) {
                        tma_load_2d_2sm(sa, &tma_a, bar, kb * BK2, m0);
                    } else {
#pragma unroll
                        for (int c = 0; c < 2; ++c)
                            tma_load_2d_2sm(sa + c * (64 * BK2 * 2), &tma_a, bar, m0 + c * 64, kb * BK2);
                    }
                    if (B_MN == 0) {
                        tma_load_2d_2sm(sb, &tma_b, bar, kb * BK2, n0);
                    } else {
#pragma unroll
                        for (int c = 0; c < 2; ++c)
                            tma_load_2d_2sm(sb + c * (64 * BK2 * 2), &tma_b, bar, n0 + c * 64, kb * BK2);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer (leader CTA, one thread) ================================
        if (leader && lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(BM2, BN2, A_MN, B_MN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
                ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN2);
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + stage * STAGE_BYTES);
                    const uint32_t sb = sa + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK2 / 16; ++k) {
                        const uint64_t da = A_MN ? ptx::make_smem_desc_sw128(sa + k * 2048, 64 * BK2 * 2, 1024)
                                                 : ptx::make_smem_desc_sw128(sa + k * 32, 16, 1024);
                        const uint64_t db = B_MN ? ptx::make_smem_desc_sw128(sb + k * 2048, 64 * BK2 * 2, 1024)
                                                 : ptx::make_smem_desc_sw128(sb + k * 32, 16, 1024);
                        umma2_bf16(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma2_commit_mc(&empty_bar[stage]);                    // frees the slot in both CTAs
                    if (kb == num_kb - 1) umma2_commit_mc(&tmem_full_bar[acc]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
            }
        }
    }
    } else {
        ptx::setmaxnreg_inc<232>();
        // ================================ epilogue warps (8 per CTA) ================================
        const int ew = warp - 4;
        const int wq = warp & 3;
        const int half = ew >> 2;
        EpiRegs E;
        E.stg = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + BAR_BYTES + ew * 4096);
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
        E.N = N; E.lane = lane; E.atomic = false;
        E.use_bias = ep.bias != nullptr;
        const uint32_t empty_remote0 = mapa(ptx::smem_u32(&tmem_empty_bar[0]), 0);
        const uint32_t empty_remote1 = mapa(ptx::smem_u32(&tmem_empty_bar[1]), 0);
        int acc = 0;
        uint32_t acc_phase = 0;
        int issued = 0;                                   // TMA stores issued by this warp (epi_chunk_tma)
        for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
            const int m0 = (tile / num_n) * BM2 + static_cast<int>(rank) * 128;
            const int n0 = (tile % num_n) * BN2;
            const int rbase = m0 + wq * 32;
            const int rows_here = min(32, M - rbase);
            bool waited = false;
            if (rows_here > 0 && ep.tma_store) {
#pragma unroll 1
                for (int c = 0; c < BN2 / 64; ++c) {
                    const int nc = n0 + half * (BN2 / 2) + c * 32;
                    if (nc >= N) break;
                    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) +
                                           static_cast<uint32_t>(acc * BN2 + half * (BN2 / 2) + c * 32);
                    epi_chunk_tma(E, &tma_c, &tma_d, taddr, rbase, M, nc, reinterpret_cast<uint8_t*>(E.stg), issued,
                                  &tmem_full_bar[acc], acc_phase, waited);
                }
            } else if (rows_here > 0) {
#pragma unroll 1
                for (int c = 0; c < BN2 / 64; ++c) {
                    const int nc = n0 + half * (BN2 / 2) + c * 32;
                    if (nc >= N) break;
                    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) +
                                           static_cast<uint32_t>(acc * BN2 + half * (BN2 / 2) + c * 32);
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
            mbar_arrive_cluster(acc == 0 ? empty_remote0 : empty_remote1);   // leader's "accumulator free" barrier
            if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
        }
        if (issued > 0 && lane == 0) ptx::bulk_wait_all();   // staging boxes stay valid until every store has drained
    }

    // ================================ teardown ================================
    ptx::tc_fence_before();
    __syncthreads();
    cluster_sync();      // the peer's shared memory / TMEM stay alive until the leader's last MMA has been consumed
    if (warp == 2) {
        ptx::tc_fence_after();
        tmem_dealloc2(tmem_base, TMEM_COLS);
    }
}

template <int A_MN, int B_MN>
static int launch_gemm2(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const GemmEpi& ep,
                        cudaStream_t stream) {
    CUtensorMap ta, tb;
    int rc;
    if (A_MN == 0) rc = make_tmap_bf16(&ta, A, K, M, lda, BK2, 128);
    else rc = make_tmap_bf16(&ta, A, M, K, lda, 64, BK2);
    if (rc) return rc;
    if (B_MN == 0) rc = make_tmap_bf16(&tb, B, K, N, ldb, BK2, 128);
    else rc = make_tmap_bf16(&tb, B, N, K, ldb, 64, BK2);
    if (rc) return rc;
    CUtensorMap tc = ta, td = ta;                          // placeholders when the TMA-store epilogue is off
    if (ep.tma_store && ep.out_bf16 != nullptr && (rc = make_tmap_out(&tc, ep.out_bf16, false, N, M, ep.ld_bf16))) return rc;
    if (ep.tma_store && ep.out_f32 != nullptr && (rc = make_tmap_out(&td, ep.out_f32, true, N, M, ep.ld_f32))) return rc;
    auto kern = gemm2_bf16_kernel<A_MN, B_MN>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
        attr_set = true;
    }
    const int num_tiles = ((M + BM2 - 1) / BM2) * ((N + BN2 - 1) / BN2);
    int clusters = num_sms() / 2;
    if (clusters > num_tiles) clusters = num_tiles;
    kern<<<clusters * 2, 384, SMEM_BYTES, stream>>>(ta, tb, tc, td, ep, M, N, K);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
    return 0;
}

}  // namespace g2

// Entry used by b200vsgg_gemm_bf16's dispatcher (gemm_tcgen05.cu).
int gemm2_launch(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, int M, int N, int K,
                 const GemmEpi& ep, cudaStream_t stream) {
    if (a_mn == 0 && b_mn == 0) return g2::launch_gemm2<0, 0>(A, lda, B, ldb, M, N, K, ep, stream);
    if (a_mn == 0 && b_mn == 1) return g2::launch_gemm2<0, 1>(A, lda, B, ldb, M, N, K, ep, stream);
    return g2::launch_gemm2<1, 1>(A, lda, B, ldb, M, N, K, ep, stream);
}

}  // namespace vsgg
