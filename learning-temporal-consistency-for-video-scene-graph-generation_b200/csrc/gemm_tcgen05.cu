// Persistent warp-specialised bf16 GEMM for sm_100a:
//   TMA (cp.async.bulk.tensor, SWIZZLE_128B) -> smem ring -> tcgen05.mma (cta_group::1, M=128,
//   N=BN, K=16) accumulating fp32 in TMEM (two accumulator stages) -> tcgen05.ld epilogue with
//   fused bias / activation / mask / dropout / residual / dual fp32+bf16 stores.
// Roles per CTA (384 threads): warp0 = TMA producer, warp1 = MMA issuer, warp2 = TMEM allocator,
//   warps4-11 = epilogue (warp w owns TMEM lanes 32*(w%4)..+31 and one column half of the tile; each
//   32x32 chunk is transposed through shared memory so every global access is row-contiguous).
// Operand majors: K-major ([rows,K], K contiguous) or MN-major ([K,rows], rows contiguous), so the
// same kernel serves forward (K,K), dgrad (K,MN) and wgrad (MN,MN) without transposed copies.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/b200vsgg.h"
#include "common.cuh"
#include "ptx.cuh"

namespace vsgg {

struct GemmEpi {
    const float* bias;
    const void* residual;
    int residual_is_bf16;
    int ldr;
    const __nv_bfloat16* mask_src;
    int ldm;
    int mask_mode;
    int act;
    float* out_f32;
    int ld_f32;
    __nv_bfloat16* out_bf16;
    int ld_bf16;
    int accumulate;
    float alpha;
    float dropout_p;
    unsigned long long dropout_seed;
    int vec_ok;  // all epilogue pointers / leading dimensions allow 16-byte vector access
};

constexpr int BM = 128;
constexpr int BK = 64;

template <int BN>
struct GemmCfg {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BN == 256) ? 4 : ((BN == 128) ? 6 : 8);
    static constexpr int ACC_STAGES = 2;
    static constexpr int TMEM_COLS = (ACC_STAGES * BN <= 32) ? 32 : ((ACC_STAGES * BN <= 64) ? 64 : ((ACC_STAGES * BN <= 128) ? 128 : ((ACC_STAGES * BN <= 256) ? 256 : 512)));
    static constexpr int BAR_BYTES = 256;
    static constexpr int EPI_BYTES = 8 * 4096;  // one swizzled 32x32 fp32 transpose buffer per epilogue warp
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + EPI_BYTES + 1024 /*align slack*/;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
    const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
    return cdf + x * pdf;
}

// Epilogue state hoisted into registers once per kernel.
struct EpiRegs {
    float* stg;                       // this warp's swizzled 32x32 fp32 transpose buffer
    const float* res_f32;
    const __nv_bfloat16* res_b16;
    const __nv_bfloat16* mask_src;
    const float* bias;
    float* out_f32;
    __nv_bfloat16* out_bf16;
    unsigned long long drop_seed;
    float alpha, inv_keep;
    uint32_t drop_thr;
    int act, mask_mode, accumulate, ld_f32, ld_bf16, ldr, ldm, N, lane;
    bool atomic, use_bias;
};

// One 32-row x 32-column chunk of the accumulator: TMEM -> registers (lane = row) -> swizzled shared
// memory -> registers (lane = column, 32 rows) -> fused epilogue -> row-contiguous global stores.
// FULL = all 32 rows exist (every m-tile but the last): no row guards, so each row costs ~6 instructions.
// Loads are clamped in-bounds instead of predicated; stores sit under one lane predicate (col < N).
template <bool FULL>
__device__ __forceinline__ void epi_chunk(const EpiRegs& E, uint32_t taddr, int rbase, int rows_here, int nc,
                                          uint64_t* full_bar, uint32_t full_phase, bool& waited) {
    const int lane = E.lane;
    const int col = nc + lane;
    const bool col_ok = col < E.N;
    const int colc = col_ok ? col : E.N - 1;
    const int last = rows_here - 1;
    // aux = the chunk's residual values, or (when there is no residual) its mask values, prefetched
    // before the accumulator is touched.  With BOTH present the mask is read inline later (rare).
    float aux[32];
    const bool has_res = E.res_f32 != nullptr || E.res_b16 != nullptr;
    const bool has_mask = E.mask_src != nullptr;
    const __nv_bfloat16* mp = has_mask ? E.mask_src + static_cast<size_t>(rbase) * E.ldm + colc : nullptr;
    if (E.res_f32 != nullptr) {
        const float* rp = E.res_f32 + static_cast<size_t>(rbase) * E.ldr + colc;
#pragma unroll
        for (int i = 0; i < 32; ++i) aux[i] = __ldg(rp + (FULL ? i : min(i, last)) * E.ldr);
    } else if (E.res_b16 != nullptr) {
        const __nv_bfloat16* rp = E.res_b16 + static_cast<size_t>(rbase) * E.ldr + colc;
#pragma unroll
        for (int i = 0; i < 32; ++i) aux[i] = __bfloat162float(rp[(FULL ? i : min(i, last)) * E.ldr]);
    } else if (has_mask) {
#pragma unroll
        for (int i = 0; i < 32; ++i) aux[i] = __bfloat162float(mp[(FULL ? i : min(i, last)) * E.ldm]);
    }
    if (!waited) {
        ptx::mbar_wait(full_bar, full_phase);
        ptx::tc_fence_after();
        waited = true;
    }
    uint32_t r[32];
    ptx::tmem_ld_32x32b_x32(taddr, r);
    ptx::tmem_ld_wait();
    float* stg = E.stg;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float4 v4 = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                __uint_as_float(r[4 * j + 3]));
        *reinterpret_cast<float4*>(stg + lane * 32 + ((j ^ (lane & 7)) << 2)) = v4;
    }
    __syncwarp();
    const float bias_v = E.use_bias ? __ldg(E.bias + colc) : 0.f;
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaf(stg[i * 32 + (((lane >> 2) ^ (i & 7)) << 2) + (lane & 3)], E.alpha, bias_v);
    __syncwarp();   // the buffer may be overwritten by the next chunk from here on
    if (E.atomic) {
        if (col_ok) {
            float* op = E.out_f32 + static_cast<size_t>(rbase) * E.ld_f32 + col;
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (FULL || i < rows_here) atomicAdd(op + i * E.ld_f32, v[i]);
        }
        return;
    }
    if (E.act == 1) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
    } else if (E.act == 2) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
    }
    if (has_mask) {
        if (!has_res) {
            if (E.mask_mode == 1) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = aux[i] > 0.f ? v[i] : 0.f;
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] *= gelu_erf_grad(aux[i]);
            }
        } else {
#pragma unroll 4
            for (int i = 0; i < 32; ++i) {
                const float m = __bfloat162float(mp[(FULL ? i : min(i, last)) * E.ldm]);
                v[i] = E.mask_mode == 1 ? (m > 0.f ? v[i] : 0.f) : v[i] * gelu_erf_grad(m);
            }
        }
    }
    if (E.drop_thr != 0u) {
        const unsigned long long base_idx = static_cast<unsigned long long>(rbase) * E.N + col;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const uint32_t h = hash_u32(E.drop_seed, base_idx + static_cast<unsigned long long>(i) * E.N);
            v[i] = (h >= E.drop_thr) ? v[i] * E.inv_keep : 0.f;
        }
    }
    if (has_res) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += aux[i];
    }
    if (!col_ok) return;
    if (E.out_f32 != nullptr) {
        float* op = E.out_f32 + static_cast<size_t>(rbase) * E.ld_f32 + col;
        if (E.accumulate) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (FULL || i < rows_here) v[i] += op[i * E.ld_f32];
        }
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (FULL || i < rows_here) op[i * E.ld_f32] = v[i];
    }
    if (E.out_bf16 != nullptr) {
        __nv_bfloat16* op = E.out_bf16 + static_cast<size_t>(rbase) * E.ld_bf16 + col;
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (FULL || i < rows_here) op[i * E.ld_bf16] = __float2bfloat16(v[i]);
    }
}

template <int BN, int A_MN, int B_MN>
__global__ void __launch_bounds__(384, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
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
    } else if (warp >= 4) {
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
        for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
            const int tile = unit % num_tiles, split = unit / num_tiles;
            const int m0 = (tile / num_n) * BM;
            const int n0 = (tile % num_n) * BN;
            const int rbase = m0 + wq * 32;
            const int rows_here = min(32, M - rbase);      // <= 0: nothing to store for this warp
            E.use_bias = ep.bias != nullptr && (splits == 1 || split == 0);
            bool waited = false;
            if (rows_here > 0) {
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
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// 2D bf16 tensor map: `inner` contiguous elements, `outer` rows with pitch ld (elements); box = box_inner x box_outer.
static int make_tmap_bf16(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t ld,
                          uint32_t box_inner, uint32_t box_outer) {
    PFN_encodeTiled enc = get_encode_fn();
    if (enc == nullptr) return set_error(B200VSGG_ERR_NO_DRIVER, "cuTensorMapEncodeTiled unavailable (no CUDA driver)");
    if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || ((ld * 2) & 15u) != 0)
        return set_error(B200VSGG_ERR_BAD_ARG, "gemm operand must be 16-byte aligned with ld % 8 == 0");
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char msg[160];
        snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed (%d) inner=%llu outer=%llu ld=%llu", (int)r,
                 (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld);
        return set_error(B200VSGG_ERR_TMAP, msg);
    }
    return 0;
}

static int g_num_sms = 0;
static int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

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
    kern<<<grid, 384, Cfg::SMEM_BYTES, stream>>>(ta, tb, ep, M, N, K, splits, kb_per, a_k_period);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
    return 0;
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
