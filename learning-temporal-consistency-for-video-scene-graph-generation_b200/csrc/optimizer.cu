// Fused multi-tensor AdamW + global-norm gradient clipping: tools/utils/AdamW.py:53-113 of the reference
// (decay applied to the weights BEFORE the moment update; parameters without a gradient are skipped by
// the caller) and torch.nn.utils.clip_grad_norm_(max_norm=5) of TEMPURA_train.py:224, in two launches
// for the whole model instead of ~8 elementwise launches per parameter tensor.  HBM-bound: 16 B read +
// 12 B written per parameter (p, g, m, v -> p, m, v); the clip coefficient is computed on the device
// from the squared-norm scalar, so the step needs no host synchronisation.
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/b200vsgg.h"
#include "common.cuh"

namespace vsgg {

__global__ void __launch_bounds__(256) grad_sqnorm_kernel(const b200vsgg_opt_tensor* __restrict__ tensors,
                                                          const int32_t* __restrict__ chunk_tensor,
                                                          const int64_t* __restrict__ chunk_off, int chunk_elems,
                                                          float* __restrict__ sq_norm) {
    const b200vsgg_opt_tensor t = tensors[chunk_tensor[blockIdx.x]];
    const int64_t off = chunk_off[blockIdx.x];
    const int64_t end = min(t.n, off + static_cast<int64_t>(chunk_elems));
    const float* g = t.g;
    float acc = 0.f;
    if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
        const int64_t v_end = off + ((end - off) & ~3LL);
        for (int64_t i = off + threadIdx.x * 4LL; i < v_end; i += 1024) {
            const float4 x = *reinterpret_cast<const float4*>(g + i);
            acc += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
        }
        for (int64_t i = v_end + threadIdx.x; i < end; i += 256) acc += g[i] * g[i];
    } else {
        for (int64_t i = off + threadIdx.x; i < end; i += 256) acc += g[i] * g[i];
    }
    acc = warp_sum(acc);
    __shared__ float part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += part[w];
        atomicAdd(sq_norm, s);
    }
}

__global__ void __launch_bounds__(256) adamw_clip_kernel(const b200vsgg_opt_tensor* __restrict__ tensors,
                                                         const int32_t* __restrict__ chunk_tensor,
                                                         const int64_t* __restrict__ chunk_off, int chunk_elems,
                                                         const float* __restrict__ sq_norm, float max_norm, float lr,
                                                         float beta1, float beta2, float eps, float weight_decay) {
    const b200vsgg_opt_tensor t = tensors[chunk_tensor[blockIdx.x]];
    const int64_t off = chunk_off[blockIdx.x];
    const int64_t end = min(t.n, off + static_cast<int64_t>(chunk_elems));
    float coef = 1.f;
    if (sq_norm != nullptr) {   // clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to 1
        const float c = max_norm / (sqrtf(*sq_norm) + 1e-6f);
        coef = c < 1.f ? c : 1.f;
    }
    const float decay = 1.f - lr * weight_decay;
    const float step_size = lr * sqrtf(t.bias_correction2) / t.bias_correction1;
    for (int64_t i = off + threadIdx.x; i < end; i += 256) {
        const float g = t.g[i] * coef;
        float p = t.p[i] * decay;
        const float m = beta1 * t.m[i] + (1.f - beta1) * g;
        const float v = beta2 * t.v[i] + (1.f - beta2) * g * g;
        p -= step_size * (m / (sqrtf(v) + eps));
        t.p[i] = p;
        t.m[i] = m;
        t.v[i] = v;
    }
}

}  // namespace vsgg

using namespace vsgg;

extern "C" int b200vsgg_grad_sqnorm(const b200vsgg_opt_tensor* tensors, const int32_t* chunk_tensor,
                                    const int64_t* chunk_off, int32_t n_chunks, int32_t chunk_elems, float* sq_norm,
                                    void* stream) {
    if (!tensors || !chunk_tensor || !chunk_off || !sq_norm || chunk_elems <= 0)
        return set_error(B200VSGG_ERR_BAD_ARG, "grad_sqnorm: bad arg");
    if (n_chunks == 0) return 0;
    grad_sqnorm_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(tensors, chunk_tensor, chunk_off, chunk_elems, sq_norm);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_adamw_clip_step(const b200vsgg_opt_tensor* tensors, const int32_t* chunk_tensor,
                                        const int64_t* chunk_off, int32_t n_chunks, int32_t chunk_elems,
                                        const float* sq_norm, float max_norm, float lr, float beta1, float beta2, float eps,
                                        float weight_decay, void* stream) {
    if (!tensors || !chunk_tensor || !chunk_off || chunk_elems <= 0)
        return set_error(B200VSGG_ERR_BAD_ARG, "adamw_clip_step: bad arg");
    if (n_chunks == 0) return 0;
    adamw_clip_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(tensors, chunk_tensor, chunk_off, chunk_elems, sq_norm,
                                                                 max_norm, lr, beta1, beta2, eps, weight_decay);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}
