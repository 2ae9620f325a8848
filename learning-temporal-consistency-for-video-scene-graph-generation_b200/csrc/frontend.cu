// Front-end layout kernels of the pair-token builder (lib/tempura.py:548 of the reference):
// the union ROI feature arrives as fp32 NCHW [N,1024,7,7]; the 1x1 conv `union_func1` is a GEMM over
// channels, so we re-lay it once as bf16 NHWC rows [N*49, 1024] (K-major for TMA/tcgen05) while
// converting.  This is the path's dominant HBM stream: 200,704 B read + 100,352 B written per pair.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>

#include "../../include/b200vsgg.h"
#include "common.cuh"

namespace vsgg {

// One CTA: one sample n, CT channels.  The CT*S source floats are contiguous -> float4 loads.
// Tile is transposed through shared memory; output rows (n, s, c0..c0+CT) are 2*CT bytes each.
template <int CT>
__global__ void __launch_bounds__(256)
nchw_to_nhwc_bf16_kernel(const float* __restrict__ in, int C, int S, __nv_bfloat16* __restrict__ out) {
    extern __shared__ float tile[];  // [CT][S+1]
    const int n = blockIdx.y;
    const int c0 = blockIdx.x * CT;
    const int pitch = S + 1;
    const float* src = in + (static_cast<size_t>(n) * C + c0) * S;
    const int total = CT * S;
    if (((reinterpret_cast<uintptr_t>(src) & 15u) == 0) && (total & 3) == 0) {
        for (int i = threadIdx.x * 4; i < total; i += blockDim.x * 4) {
            const float4 v = __ldcs(reinterpret_cast<const float4*>(src + i));  // streaming: read once
            const float a[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int e = i + j;
                const int c = e / S, s = e - c * S;
                tile[c * pitch + s] = a[j];
            }
        }
    } else {
        for (int e = threadIdx.x; e < total; e += blockDim.x) {
            const int c = e / S, s = e - c * S;
            tile[c * pitch + s] = src[e];
        }
    }
    __syncthreads();
    // each thread writes 8 consecutive channels (16 B) of one spatial position
    constexpr int VPR = CT / 8;  // vectors per output row
    for (int i = threadIdx.x; i < S * VPR; i += blockDim.x) {
        const int s = i / VPR, cv = (i - s * VPR) * 8;
        float a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = tile[(cv + j) * pitch + s];
        store_bf16x8(out + (static_cast<size_t>(n) * S + s) * C + c0 + cv, a);
    }
}

// Backward layout op: dX_nhwc bf16 [N*S, C] -> fp32 NCHW [N,C,S] is never needed (union_feat comes
// from the frozen detector, TEMPURA_train.py:160-161), so there is no inverse kernel.

// fp32 NCHW [N,C,S] -> fp32 NHWC rows [N*S, C] (used for the mask-branch output added in the
// union GEMM epilogue).  Same tiling, fp32 out.
template <int CT>
__global__ void __launch_bounds__(256)
nchw_to_nhwc_f32_kernel(const float* __restrict__ in, int C, int S, float* __restrict__ out) {
    extern __shared__ float tile[];
    const int n = blockIdx.y;
    const int c0 = blockIdx.x * CT;
    const int pitch = S + 1;
    const float* src = in + (static_cast<size_t>(n) * C + c0) * S;
    const int total = CT * S;
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        const int c = e / S, s = e - c * S;
        tile[c * pitch + s] = src[e];
    }
    __syncthreads();
    constexpr int VPR = CT / 4;
    for (int i = threadIdx.x; i < S * VPR; i += blockDim.x) {
        const int s = i / VPR, cv = (i - s * VPR) * 4;
        const float4 v = make_float4(tile[cv * pitch + s], tile[(cv + 1) * pitch + s], tile[(cv + 2) * pitch + s],
                                     tile[(cv + 3) * pitch + s]);
        *reinterpret_cast<float4*>(out + (static_cast<size_t>(n) * S + s) * C + c0 + cv) = v;
    }
}

// NHWC rows [N*S, C] (fp32) -> NCHW [N,C,S] fp32: gradient hand-off back to the mask branch.
template <int CT>
__global__ void __launch_bounds__(256)
nhwc_to_nchw_f32_kernel(const float* __restrict__ in, int C, int S, float* __restrict__ out) {
    extern __shared__ float tile[];
    const int n = blockIdx.y;
    const int c0 = blockIdx.x * CT;
    const int pitch = S + 1;
    for (int i = threadIdx.x; i < S * CT; i += blockDim.x) {
        const int s = i / CT, c = i - s * CT;
        tile[c * pitch + s] = in[(static_cast<size_t>(n) * S + s) * C + c0 + c];
    }
    __syncthreads();
    float* dst = out + (static_cast<size_t>(n) * C + c0) * S;
    for (int e = threadIdx.x; e < CT * S; e += blockDim.x) {
        const int c = e / S, s = e - c * S;
        dst[e] = tile[c * pitch + s];
    }
}

}  // namespace vsgg

using namespace vsgg;

extern "C" int b200vsgg_nchw_to_nhwc_bf16(const float* in, int32_t n, int32_t channels, int32_t spatial, void* out,
                                          void* stream) {
    constexpr int CT = 64;
    if (!in || !out || channels % CT != 0 || spatial <= 0 || n < 0)
        return set_error(B200VSGG_ERR_BAD_ARG, "nchw_to_nhwc_bf16: channels must be a multiple of 64");
    if (n == 0) return 0;
    if (n > 65535 * 32) return set_error(B200VSGG_ERR_BAD_ARG, "nchw_to_nhwc_bf16: too many samples");
    const size_t smem = sizeof(float) * CT * (spatial + 1);
    // gridDim.y is limited to 65535: fold samples into chunks
    for (int n0 = 0; n0 < n; n0 += 65535) {
        const int nn = (n - n0) < 65535 ? (n - n0) : 65535;
        dim3 grid(channels / CT, nn);
        nchw_to_nhwc_bf16_kernel<CT><<<grid, 256, smem, (cudaStream_t)stream>>>(
            in + static_cast<size_t>(n0) * channels * spatial, channels, spatial,
            (__nv_bfloat16*)out + static_cast<size_t>(n0) * channels * spatial);
    }
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_nchw_to_nhwc_f32(const float* in, int32_t n, int32_t channels, int32_t spatial, float* out,
                                         void* stream) {
    constexpr int CT = 64;
    if (!in || !out || channels % CT != 0 || spatial <= 0 || n < 0)
        return set_error(B200VSGG_ERR_BAD_ARG, "nchw_to_nhwc_f32: channels must be a multiple of 64");
    const size_t smem = sizeof(float) * CT * (spatial + 1);
    for (int n0 = 0; n0 < n; n0 += 65535) {
        const int nn = (n - n0) < 65535 ? (n - n0) : 65535;
        dim3 grid(channels / CT, nn);
        nchw_to_nhwc_f32_kernel<CT><<<grid, 256, smem, (cudaStream_t)stream>>>(
            in + static_cast<size_t>(n0) * channels * spatial, channels, spatial,
            out + static_cast<size_t>(n0) * channels * spatial);
    }
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_nhwc_to_nchw_f32(const float* in, int32_t n, int32_t channels, int32_t spatial, float* out,
                                         void* stream) {
    constexpr int CT = 64;
    if (!in || !out || channels % CT != 0 || spatial <= 0 || n < 0)
        return set_error(B200VSGG_ERR_BAD_ARG, "nhwc_to_nchw_f32: channels must be a multiple of 64");
    const size_t smem = sizeof(float) * CT * (spatial + 1);
    for (int n0 = 0; n0 < n; n0 += 65535) {
        const int nn = (n - n0) < 65535 ? (n - n0) : 65535;
        dim3 grid(channels / CT, nn);
        nhwc_to_nchw_f32_kernel<CT><<<grid, 256, smem, (cudaStream_t)stream>>>(
            in + static_cast<size_t>(n0) * channels * spatial, channels, spatial,
            out + static_cast<size_t>(n0) * channels * spatial);
    }
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}


// ------------------------------------------------------------------------------------------------
// Small host -> device uploads (index vectors of the segment plans) WITHOUT the copy engine: the source is
// pinned host memory, which is device-addressable under unified addressing, and a kernel on the compute
// stream reads it over PCIe.  DMA queues are FIFO: a cudaMemcpyAsync issued while a multi-GB input batch is
// being prefetched on another stream would wait for the whole batch; this does not.
// ------------------------------------------------------------------------------------------------
namespace vsgg {
__global__ void upload_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long n16) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n16;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        dst[i] = src[i];
}
}  // namespace vsgg

extern "C" int b200vsgg_upload(const void* h_pinned_src, void* dst, int64_t bytes, void* stream) {
    if (!h_pinned_src || !dst || bytes < 0 || (bytes & 15) || (reinterpret_cast<uintptr_t>(h_pinned_src) & 15) ||
        (reinterpret_cast<uintptr_t>(dst) & 15))
        return vsgg::set_error(B200VSGG_ERR_BAD_ARG, "upload: pointers and size must be 16-byte multiples");
    if (bytes == 0) return 0;
    const long long n16 = bytes / 16;
    long long g = (n16 + 255) / 256;
    if (g > 592) g = 592;
    vsgg::upload_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>((const uint4*)h_pinned_src, (uint4*)dst, n16);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}
