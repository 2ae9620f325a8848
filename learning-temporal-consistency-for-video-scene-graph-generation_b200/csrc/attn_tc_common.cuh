// Shared by the tcgen05 attention kernels (attn_tc.cu forward, attn_tc_bwd.cu backward): tile geometry, the 3-D TMA
// tensor map that zero-pads a head to the 128-byte swizzled rows, small packing helpers.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>

#include "gemm_common.cuh"
#include "attn_dropout.cuh"

namespace vsgg {
namespace atc {

constexpr int BQ = 128;                 // queries per tile
constexpr int BKV = 128;                // keys per tile
constexpr int TILE_BYTES = 128 * 128;   // 128 rows x 128-byte swizzled rows (64 bf16, head_dim zero-padded by TMA)

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}

// {head_dim, rows, heads} view of a [rows, heads * head_dim] bf16 matrix (pitch ld), box {64, 128, 1}, SWIZZLE_128B.
static inline int make_tmap_heads(CUtensorMap* tm, const void* base, uint64_t hd, uint64_t rows, uint64_t heads, uint64_t ld) {
    PFN_encodeTiled enc = get_encode_fn();
    if (enc == nullptr) return set_error(B200VSGG_ERR_NO_DRIVER, "cuTensorMapEncodeTiled unavailable (no CUDA driver)");
    if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || ((ld * 2) & 15u) != 0 || ((hd * 2) & 15u) != 0)
        return set_error(B200VSGG_ERR_BAD_ARG, "attn_tc: q/k/v must be 16-byte aligned, ld % 8 == 0, head_dim % 8 == 0");
    cuuint64_t dims[3] = {hd, rows, heads};
    cuuint64_t strides[2] = {ld * 2, hd * 2};
    cuuint32_t box[3] = {64, 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char msg[160];
        snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed (%d) for an attention operand hd=%llu rows=%llu ld=%llu",
                 (int)r, (unsigned long long)hd, (unsigned long long)rows, (unsigned long long)ld);
        return set_error(B200VSGG_ERR_TMAP, msg);
    }
    return 0;
}

}  // namespace atc
}  // namespace vsgg
