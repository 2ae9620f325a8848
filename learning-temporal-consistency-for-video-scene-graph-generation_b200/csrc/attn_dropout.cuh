// Attention-dropout mask shared by every TokenGT attention kernel (multihead_attention.py:175-177 of the reference:
// F.dropout on the fp32 softmax weights).  Masks are never stored: forward and backward regenerate the same bit from
// (seed, query row, head, key) — counter-based, so a 128-wide forward tile, a 64-wide mma.sync tile and the transposed
// dK/dV tiles all agree.
//
// Cost model: the softmax warps of the tcgen05 kernels are MUFU-bound with ~3 spare issue slots per element, and a
// murmur finaliser per element (the round-1 scheme, ~12 integer instructions) made them issue-bound 2.3x over.  Here
//   row_key            = hash(seed, row, head)                              once per (row, head)
//   stream s0          = mix(row_key, key block of 64)                      once per (row, 64 keys)
//   draw  t(n, h)      = xs16((s0 + h*DELTA) * A^(n+1) + C_(n+1))           one IMAD + shift + xor per FOUR keys
//   key k of the block : n = k >> 3, h = (k >> 2) & 1, byte {0,2,1,3}[k & 3] of t, kept iff byte >= thr8
//                        (keys 0,1 of a group sit in the even bytes, 2,3 in the odd ones: one SWAR compare per pair)
// i.e. 16 random-access LCG draws (two interleaved streams of 8 jumps) per 64 keys, 8 bits per key.  The dropout
// probability is therefore quantised to 1/256 (p = 0.1 -> 26/256 = 0.1016, and the survivors are scaled by
// 256/(256-26) so the estimator stays unbiased).  Measured on 4096 rows x 1024 keys: keep rate exact to 5e-5, serial /
// cross-row correlations < 7e-3 (tools/ history: see DESIGN.md).
#pragma once
#include <cstdint>

#include "common.cuh"

namespace vsgg {
namespace adrop {

constexpr uint32_t LCG_A = 1664525u, LCG_C = 1013904223u, DELTA = 0x7F4A7C15u;

__host__ __device__ constexpr uint32_t lcg_a(int n) {   // A^(n+1)
    uint32_t a = 1u;
    for (int i = 0; i <= n; ++i) a *= LCG_A;
    return a;
}
__host__ __device__ constexpr uint32_t lcg_c(int n) {   // C * (A^n + ... + A + 1)
    uint32_t c = 0u;
    for (int i = 0; i <= n; ++i) c = c * LCG_A + LCG_C;
    return c;
}

__host__ __device__ inline uint32_t thr8_of(float p) { return p > 0.f ? static_cast<uint32_t>(p * 256.f + 0.5f) : 0u; }
__host__ __device__ inline float inv_keep_of(uint32_t thr8) { return 256.f / static_cast<float>(256u - thr8); }

__device__ __forceinline__ uint32_t row_key(unsigned long long seed, int row, int head) {
    return hash_u32(seed, static_cast<unsigned long long>(row) * 64ull + static_cast<unsigned long long>(head));
}
// key_block64 = (key - first key of the sequence) >> 6
__device__ __forceinline__ uint32_t stream_seed(uint32_t rk, uint32_t key_block64) {
    uint32_t h = rk ^ (key_block64 * 0x9E3779B1u);
    h ^= h >> 15; h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    return h;
}
// 32 random bits for the keys 8n + 4h .. 8n + 4h + 3 of the 64-key block; n, h compile-time in the unrolled kernels
__device__ __forceinline__ uint32_t draw(uint32_t s0, int n, int h) {
    const uint32_t s = (s0 + (h ? DELTA : 0u)) * lcg_a(n) + lcg_c(n);
    return s ^ (s >> 16);
}
// Generic per-element form (any kernel, any access order): is key `key_rel` (relative to the sequence start) kept?
__device__ __forceinline__ bool keep(uint32_t thr8, uint32_t rk, int key_rel) {
    const uint32_t s0 = stream_seed(rk, static_cast<uint32_t>(key_rel) >> 6);
    const int k = key_rel & 63;
    const int n = k >> 3;
    // runtime n: the eight (A, C) pairs as a select chain (folds to immediates wherever n is known after unrolling)
    uint32_t a = lcg_a(0), c = lcg_c(0);
#pragma unroll
    for (int i = 1; i < 8; ++i)
        if (n == i) { a = lcg_a(i); c = lcg_c(i); }
    const uint32_t s = (s0 + (((k >> 2) & 1) ? DELTA : 0u)) * a + c;
    const uint32_t t = s ^ (s >> 16);
    const int byte = ((k & 1) << 1) | ((k >> 1) & 1);
    return ((t >> (8 * byte)) & 255u) >= thr8;
}
// SWAR: the four keep bits of one draw as two bf16x2 lane masks — lo = keys (0,1) of the group (bytes 0 and 2),
// hi = keys (2,3) (bytes 1 and 3): each 16-bit half is 0xFFFF where the key survives.  K8 = (256 - thr8) * 0x00010001.
__device__ __forceinline__ void keep_masks4(uint32_t t, uint32_t K8, uint32_t& lo, uint32_t& hi) {
    const uint32_t e = ((t & 0x00FF00FFu) + K8) & 0x01000100u;          // byte >= thr8  <=>  bit 8 of (byte + 256 - thr8)
    const uint32_t o = (((t >> 8) & 0x00FF00FFu) + K8) & 0x01000100u;
    lo = (e >> 8) * 0xFFFFu;                                            // bits 0 / 16 -> 0x0000FFFF / 0xFFFF0000
    hi = (o >> 8) * 0xFFFFu;
}

}  // namespace adrop
}  // namespace vsgg
