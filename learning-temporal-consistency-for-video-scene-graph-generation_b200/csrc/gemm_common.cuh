// Pieces shared by the 1-CTA and the 2-CTA (cta_group::2) tcgen05 GEMM kernels: epilogue parameter block,
// the fused chunk epilogue (TMEM -> swizzled smem transpose -> row-contiguous global access), and the TMA
// tensor-map helpers.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

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
    int tma_store;  // the epilogue writes through TMA stores (epi_chunk_tma)
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
    static constexpr int BAR_BYTES = 1024;   // keeps the epilogue staging boxes 1024-byte aligned (swizzled TMA stores)
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

// TMA-store epilogue (everything except split-K atomics, in-place accumulation and residual+mask together): one
// 32 x 32 chunk goes straight from the TMEM layout (lane = row, 32 consecutive columns in registers) through a
// hardware-swizzled shared-memory box to cp.async.bulk.tensor stores.  ~80-120 instructions per chunk instead of
// ~400: no transpose, no per-row address arithmetic, no 2- or 4-byte stores; rows >= M and columns >= N are clipped
// by the tensor map.  Residual / mask rows are read in the same layout (16-byte vector loads along the row).
//   fp32 output : box 32 rows x 128 B, SWIZZLE_128B (16-byte piece k of row r sits at piece k ^ (r & 7))
//   bf16 output : box 32 rows x  64 B, SWIZZLE_64B  (piece k of row r at k ^ ((r >> 1) & 3)); bf16-only outputs
//                 alternate between two boxes so the store of chunk c overlaps the math of chunk c + 1
// `box` = this warp's 4 KB staging area (1024-byte aligned); lane 0 owns the bulk groups.
__device__ __forceinline__ void epi_chunk_tma(const EpiRegs& E, const CUtensorMap* tm_bf16, const CUtensorMap* tm_f32,
                                              uint32_t taddr, int rbase, int M, int nc, uint8_t* box, int& issued,
                                              uint64_t* full_bar, uint32_t full_phase, bool& waited) {
    const int lane = E.lane;
    const int row = min(rbase + lane, M - 1);
    const bool has_mask = E.mask_src != nullptr;
    // ---- prefetch this lane's row segment of the residual / mask before the accumulator is waited for
    float4 ax[8];
    uint4 ab[4];
    if (E.res_f32 != nullptr) {
        const float4* rp = reinterpret_cast<const float4*>(E.res_f32 + static_cast<size_t>(row) * E.ldr + nc);
#pragma unroll
        for (int j = 0; j < 8; ++j) ax[j] = (nc + 4 * j + 4 <= E.N) ? __ldg(rp + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else if (E.res_b16 != nullptr || has_mask) {
        const __nv_bfloat16* bp = E.res_b16 != nullptr ? E.res_b16 + static_cast<size_t>(row) * E.ldr + nc
                                                      : E.mask_src + static_cast<size_t>(row) * E.ldm + nc;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            ab[j] = (nc + 8 * j + 8 <= E.N) ? __ldg(reinterpret_cast<const uint4*>(bp) + j) : make_uint4(0u, 0u, 0u, 0u);
    }
    const float bias_l = (E.use_bias && nc + lane < E.N) ? __ldg(E.bias + nc + lane) : 0.f;
    if (!waited) {
        ptx::mbar_wait(full_bar, full_phase);
        ptx::tc_fence_after();
        waited = true;
    }
    uint32_t r[32];
    ptx::tmem_ld_32x32b_x32(taddr, r);
    ptx::tmem_ld_wait();
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(r[j]), E.alpha, __shfl_sync(0xffffffffu, bias_l, j));
    if (E.act == 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    } else if (E.act == 2) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
    }
    if (has_mask) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&ab[j]);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float2 m = __bfloat1622float2(h[t]);
                const int c = 8 * j + 2 * t;
                if (E.mask_mode == 1) {
                    v[c] = m.x > 0.f ? v[c] : 0.f;
                    v[c + 1] = m.y > 0.f ? v[c + 1] : 0.f;
                } else {
                    v[c] *= gelu_erf_grad(m.x);
                    v[c + 1] *= gelu_erf_grad(m.y);
                }
            }
        }
    }
    if (E.drop_thr != 0u) {
        const unsigned long long base_idx = static_cast<unsigned long long>(rbase + lane) * E.N + nc;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const uint32_t h = hash_u32(E.drop_seed, base_idx + j);
            v[j] = (h >= E.drop_thr) ? v[j] * E.inv_keep : 0.f;
        }
    }
    if (E.res_f32 != nullptr) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            v[4 * j] += ax[j].x; v[4 * j + 1] += ax[j].y; v[4 * j + 2] += ax[j].z; v[4 * j + 3] += ax[j].w;
        }
    } else if (E.res_b16 != nullptr) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&ab[j]);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float2 m = __bfloat1622float2(h[t]);
                v[8 * j + 2 * t] += m.x;
                v[8 * j + 2 * t + 1] += m.y;
            }
        }
    }
    const bool f32_out = E.out_f32 != nullptr, b16_out = E.out_bf16 != nullptr;
    if (f32_out) {
        if (issued > 0) {                      // single 4 KB box: the previous store must have read it
            if (lane == 0) ptx::bulk_wait_read<0>();
            __syncwarp();
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
            *reinterpret_cast<float4*>(box + lane * 128 + ((k ^ (lane & 7)) << 4)) =
                make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            if (E.atomic) ptx::tma_reduce_add_2d(tm_f32, box, nc, rbase);     // split-K / stream-K partial tile
            else ptx::tma_store_2d(tm_f32, box, nc, rbase);
            ptx::bulk_commit();
        }
        ++issued;
    }
    if (b16_out) {
        uint8_t* dst = box;
        if (f32_out) {
            if (lane == 0) ptx::bulk_wait_read<0>();
            __syncwarp();
        } else {
            dst = box + (issued & 1) * 2048;
            if (issued >= 2) {                  // the store issued two chunks ago must have read this box
                if (lane == 0) ptx::bulk_wait_read<1>();
                __syncwarp();
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint4 u;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
            for (int t = 0; t < 4; ++t) h[t] = __floats2bfloat162_rn(v[8 * k + 2 * t], v[8 * k + 2 * t + 1]);
            *reinterpret_cast<uint4*>(dst + lane * 64 + ((k ^ ((lane >> 1) & 3)) << 4)) = u;
        }
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            ptx::tma_store_2d(tm_bf16, dst, nc, rbase);
            ptx::bulk_commit();
        }
        ++issued;
    }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline PFN_encodeTiled get_encode_fn() {
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

// Host-side cache of encoded tensor maps: a training step issues ~90 GEMMs with 2-4 maps each, and from the second step
// on every (pointer, shape, pitch, box) repeats (the caching allocator hands the same blocks back), so the driver's
// cuTensorMapEncodeTiled is paid once per distinct operand instead of ~350 times per step.  A map only depends on the
// key fields, so a stale pointer that happens to be reused with the same geometry yields the correct map anyway.
struct TmapKey {
    uint64_t base, d0, d1, ld, box, kind;
    bool operator==(const TmapKey& o) const {
        return base == o.base && d0 == o.d0 && d1 == o.d1 && ld == o.ld && box == o.box && kind == o.kind;
    }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey& k) const {
        uint64_t h = k.base * 0x9E3779B97F4A7C15ull;
        for (uint64_t v : {k.d0, k.d1, k.ld, k.box, k.kind}) h = (h ^ (h >> 29)) * 0xBF58476D1CE4E5B9ull + v;
        return static_cast<size_t>(h ^ (h >> 32));
    }
};
static inline bool tmap_cache_get(const TmapKey& k, CUtensorMap* out, const CUtensorMap* put) {
    static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
    static std::mutex mu;
    static const bool off = []() { const char* e = getenv("B200VSGG_NO_TMAP_CACHE"); return e && e[0] == '1'; }();
    if (off) return false;
    std::lock_guard<std::mutex> lock(mu);
    if (put != nullptr) {
        if (cache.size() >= 8192) cache.clear();
        cache.emplace(k, *put);
        return true;
    }
    auto it = cache.find(k);
    if (it == cache.end()) return false;
    *out = it->second;
    return true;
}

// Output tensor maps of the TMA-store epilogue: 32 x 32 boxes of a row-major [M, N] matrix (pitch ld elements).
static inline int make_tmap_out(CUtensorMap* tm, const void* base, bool is_f32, uint64_t N, uint64_t M, uint64_t ld) {
    PFN_encodeTiled enc = get_encode_fn();
    if (enc == nullptr) return set_error(B200VSGG_ERR_NO_DRIVER, "cuTensorMapEncodeTiled unavailable (no CUDA driver)");
    const uint64_t es = is_f32 ? 4 : 2;
    if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || ((ld * es) & 15u) != 0)
        return set_error(B200VSGG_ERR_BAD_ARG, "gemm output must be 16-byte aligned with a 16-byte multiple pitch");
    const TmapKey key{reinterpret_cast<uint64_t>(base), N, M, ld, 32, is_f32 ? 1ull : 2ull};
    if (tmap_cache_get(key, tm, nullptr)) return 0;
    cuuint64_t dims[2] = {N, M};
    cuuint64_t strides[1] = {ld * es};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                     const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     is_f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(B200VSGG_ERR_TMAP, "cuTensorMapEncodeTiled failed for a GEMM output");
    tmap_cache_get(key, nullptr, tm);
    return 0;
}

// 2D bf16 tensor map: `inner` contiguous elements, `outer` rows with pitch ld (elements); box = box_inner x box_outer.
static inline int make_tmap_bf16(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t ld,
                          uint32_t box_inner, uint32_t box_outer, bool swizzle128 = true) {
    PFN_encodeTiled enc = get_encode_fn();
    if (enc == nullptr) return set_error(B200VSGG_ERR_NO_DRIVER, "cuTensorMapEncodeTiled unavailable (no CUDA driver)");
    if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || ((ld * 2) & 15u) != 0)
        return set_error(B200VSGG_ERR_BAD_ARG, "gemm operand must be 16-byte aligned with ld % 8 == 0");
    const TmapKey key{reinterpret_cast<uint64_t>(base), inner, outer, ld,
                      (static_cast<uint64_t>(box_inner) << 32) | box_outer, swizzle128 ? 3ull : 4ull};
    if (tmap_cache_get(key, tm, nullptr)) return 0;
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char msg[160];
        snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed (%d) inner=%llu outer=%llu ld=%llu", (int)r,
                 (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld);
        return set_error(B200VSGG_ERR_TMAP, msg);
    }
    tmap_cache_get(key, nullptr, tm);
    return 0;
}

static inline int num_sms() {
    static int g_num_sms = 0;
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

}  // namespace vsgg
