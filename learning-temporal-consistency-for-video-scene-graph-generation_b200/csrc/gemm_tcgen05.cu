// Persistent warp-specialised bf16 GEMM for sm_100a:
//   TMA (cp.async.bulk.tensor, SWIZZLE_128B) -> smem ring -> tcgen05.mma (cta_group::1, M=128,
//   N=BN, K=16) accumulating fp32 in TMEM (two accumulator stages) -> tcgen05.ld epilogue with
//   fused bias / activation / mask / dropout / residual / dual fp32+bf16 stores.
// Roles per CTA (256 threads): warp0 = TMA producer, warp1 = MMA issuer, warp2 = TMEM allocator,
//   warps4-7 = epilogue (warp w owns TMEM lanes 32*(w%4)..+31).
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
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
    const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
    return cdf + x * pdf;
}

template <int BN, int A_MN, int B_MN>
__global__ void __launch_bounds__(256, 1)
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
            ptx::mbar_init(&tmem_empty_bar[i], 128);
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
        // ================================ epilogue warps ================================
        const int wq = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        const float inv_keep = ep.dropout_p > 0.f ? 1.0f / (1.0f - ep.dropout_p) : 1.0f;
        const uint32_t drop_thr = ep.dropout_p > 0.f ? static_cast<uint32_t>(ep.dropout_p * 4294967296.0) : 0u;
        for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
            const int tile = unit % num_tiles, split = unit / num_tiles;
            const int m0 = (tile / num_n) * BM;
            const int n0 = (tile % num_n) * BN;
            ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
            ptx::tc_fence_after();
            const int row = m0 + wq * 32 + lane;
            const bool row_ok = row < M;
            const size_t rowz = static_cast<size_t>(row_ok ? row : 0);
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                const int nc = n0 + c * 32;
                if (nc >= N) break;  // warp-uniform
                uint32_t r[32];
                ptx::tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(wq * 32) << 16) +
                                            static_cast<uint32_t>(acc * BN + c * 32),
                                        r);
                ptx::tmem_ld_wait();
                if (row_ok) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int n = nc + g * 8;
                        if (n >= N) break;
                        float v[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]) * ep.alpha;
                        const bool full8 = (n + 8 <= N) && ep.vec_ok;
                        if (splits > 1) {
                            // split-K partial: bias once (split 0), then atomic accumulation into fp32
                            if (ep.bias != nullptr && split == 0)
                                for (int j = 0; j < 8 && n + j < N; ++j) v[j] += __ldg(ep.bias + n + j);
                            float* op = ep.out_f32 + rowz * ep.ld_f32 + n;
                            if (full8) {
                                atomicAdd(reinterpret_cast<float4*>(op), make_float4(v[0], v[1], v[2], v[3]));
                                atomicAdd(reinterpret_cast<float4*>(op + 4), make_float4(v[4], v[5], v[6], v[7]));
                            } else {
                                for (int j = 0; j < 8 && n + j < N; ++j) atomicAdd(op + j, v[j]);
                            }
                            continue;
                        }
                        if (ep.bias != nullptr) {
                            if (full8) {
                                const float4 b0 = __ldg(reinterpret_cast<const float4*>(ep.bias + n));
                                const float4 b1 = __ldg(reinterpret_cast<const float4*>(ep.bias + n + 4));
                                v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                                v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
                            } else {
                                for (int j = 0; j < 8 && n + j < N; ++j) v[j] += __ldg(ep.bias + n + j);
                            }
                        }
                        if (ep.act == 1) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
                        } else if (ep.act == 2) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] = gelu_erf(v[j]);
                        }
                        if (ep.mask_src != nullptr) {
                            const __nv_bfloat16* mp = ep.mask_src + rowz * ep.ldm + n;
                            float mv[8];
                            if (full8) {
                                load_bf16x8(mp, mv);
                            } else {
                                for (int j = 0; j < 8; ++j) mv[j] = (n + j < N) ? __bfloat162float(mp[j]) : 0.f;
                            }
                            if (ep.mask_mode == 1) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[j] = mv[j] > 0.f ? v[j] : 0.f;
                            } else {
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[j] *= gelu_erf_grad(mv[j]);
                            }
                        }
                        if (drop_thr != 0u) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const uint32_t h = hash_u32(ep.dropout_seed, rowz * static_cast<size_t>(N) + n + j);
                                v[j] = (h >= drop_thr) ? v[j] * inv_keep : 0.f;
                            }
                        }
                        if (ep.residual != nullptr) {
                            if (ep.residual_is_bf16) {
                                const __nv_bfloat16* rp =
                                    reinterpret_cast<const __nv_bfloat16*>(ep.residual) + rowz * ep.ldr + n;
                                float rv[8];
                                if (full8) {
                                    load_bf16x8(rp, rv);
                                } else {
                                    for (int j = 0; j < 8; ++j) rv[j] = (n + j < N) ? __bfloat162float(rp[j]) : 0.f;
                                }
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[j] += rv[j];
                            } else {
                                const float* rp = reinterpret_cast<const float*>(ep.residual) + rowz * ep.ldr + n;
                                if (full8) {
                                    const float4 a0 = *reinterpret_cast<const float4*>(rp);
                                    const float4 a1 = *reinterpret_cast<const float4*>(rp + 4);
                                    v[0] += a0.x; v[1] += a0.y; v[2] += a0.z; v[3] += a0.w;
                                    v[4] += a1.x; v[5] += a1.y; v[6] += a1.z; v[7] += a1.w;
                                } else {
                                    for (int j = 0; j < 8 && n + j < N; ++j) v[j] += rp[j];
                                }
                            }
                        }
                        if (ep.out_f32 != nullptr) {
                            float* op = ep.out_f32 + rowz * ep.ld_f32 + n;
                            if (full8) {
                                float4 o0 = make_float4(v[0], v[1], v[2], v[3]);
                                float4 o1 = make_float4(v[4], v[5], v[6], v[7]);
                                if (ep.accumulate) {
                                    const float4 p0 = *reinterpret_cast<const float4*>(op);
                                    const float4 p1 = *reinterpret_cast<const float4*>(op + 4);
                                    o0.x += p0.x; o0.y += p0.y; o0.z += p0.z; o0.w += p0.w;
                                    o1.x += p1.x; o1.y += p1.y; o1.z += p1.z; o1.w += p1.w;
                                    v[0] = o0.x; v[1] = o0.y; v[2] = o0.z; v[3] = o0.w;
                                    v[4] = o1.x; v[5] = o1.y; v[6] = o1.z; v[7] = o1.w;
                                }
                                *reinterpret_cast<float4*>(op) = o0;
                                *reinterpret_cast<float4*>(op + 4) = o1;
                            } else {
                                for (int j = 0; j < 8 && n + j < N; ++j) {
                                    if (ep.accumulate) v[j] += op[j];
                                    op[j] = v[j];
                                }
                            }
                        }
                        if (ep.out_bf16 != nullptr) {
                            __nv_bfloat16* op = ep.out_bf16 + rowz * ep.ld_bf16 + n;
                            if (full8) {
                                store_bf16x8(op, v);
                            } else {
                                for (int j = 0; j < 8 && n + j < N; ++j) op[j] = __float2bfloat16(v[j]);
                            }
                        }
                    }
                }
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
    kern<<<grid, 256, Cfg::SMEM_BYTES, stream>>>(ta, tb, ep, M, N, K, splits, kb_per, a_k_period);
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
