// Shared helpers for all b200vsgg translation units: error reporting, vectorised bf16 access,
// warp/block reductions, counter-based RNG.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>

namespace vsgg {

// Records a message for b200vsgg_last_error() and returns `code` (defined in capi.cu).
int set_error(int code, const char* msg);

#define VSGG_CUDA_CHECK_LAUNCH()                                            \
    do {                                                                    \
        cudaError_t e__ = cudaGetLastError();                               \
        if (e__ != cudaSuccess) return ::vsgg::set_error((int)e__, cudaGetErrorString(e__)); \
    } while (0)

__device__ __forceinline__ void load_bf16x8(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}
__device__ __forceinline__ void store_bf16x8(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void load_bf16x4(const __nv_bfloat16* p, float (&v)[4]) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
    const float2 a = __bfloat1622float2(h[0]);
    const float2 b = __bfloat1622float2(h[1]);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void store_bf16x4(__nv_bfloat16* p, const float (&v)[4]) {
    uint2 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
    h[0] = __floats2bfloat162_rn(v[0], v[1]);
    h[1] = __floats2bfloat162_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = u;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// q = i / d, r = i % d for grid-stride indices.  `small` (uniform: the whole index space fits 32 bits) selects 32-bit
// arithmetic: the 64-bit division sequence is ~60 instructions and made several 16-byte-per-thread kernels ALU-bound.
__device__ __forceinline__ void divmod_idx(long long i, int d, bool small, long long& q, int& r) {
    if (small) {
        const unsigned iu = static_cast<unsigned>(i), qq = iu / static_cast<unsigned>(d);
        q = qq;
        r = static_cast<int>(iu - qq * static_cast<unsigned>(d));
    } else {
        q = i / d;
        r = static_cast<int>(i - q * d);
    }
}

// Counter-based 32-bit hash (splitmix64 finaliser) used for dropout masks: the same (seed, index)
// gives the same bit in forward and backward, so masks are never stored.
__host__ __device__ __forceinline__ unsigned long long hash_u64(unsigned long long seed, unsigned long long idx) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (idx + 1ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint32_t hash_u32(unsigned long long seed, unsigned long long idx) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (idx + 1ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return static_cast<uint32_t>(z >> 32);
}

}  // namespace vsgg
