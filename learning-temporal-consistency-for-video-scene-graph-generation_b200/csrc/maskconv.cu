// Channels-last kernels of the spatial-mask branch of the pair token (lib/tempura.py:466-474 of the
// reference):  Conv7x7/s2 (2->128) + ReLU + BatchNorm2d -> MaxPool 3x3/s2/p1 -> Conv3x3 (128->256) +
// ReLU + BatchNorm2d.  Both convolutions run as tcgen05 GEMMs over im2col rows (gemm_tcgen05.cu);
// everything around them is HBM-bound and lives here: im2col / col2im, per-video BatchNorm
// statistics (the reference's batch IS one video, so statistics are segmented by video), the fused
// BN-apply + max-pool, and the BN/ReLU backward.  All tensors are NHWC rows, bf16, 128-bit accesses.
#include <cuda_runtime.h>
#include <algorithm>
#include <cuda_bf16.h>
#include <cstdint>

#include "../../include/b200vsgg.h"
#include "common.cuh"

namespace vsgg {

static inline int blocks_for(long long items, int per_block) {
    long long g = (items + per_block - 1) / per_block;
    if (g < 1) g = 1;
    if (g > 0x7fffffffLL) g = 0x7fffffffLL;
    return static_cast<int>(g);
}

// ------------------------------------------------------------------------------------------------
// masks fp32 [n,2,27,27] -> rows [n*196, ld] bf16, column c*49 + kh*7 + kw, stride 2, padding 3.
// One CTA per pair: the 5.8 kB of masks are staged in shared memory, then 196 x ld outputs are
// written with 16-byte stores.
// ------------------------------------------------------------------------------------------------
template <typename T>   // float: the reference's fp32 hand-off; __nv_bfloat16: the producer-side bf16 hand-off ((f).4)
__global__ void __launch_bounds__(256) mask_im2col_kernel(const T* __restrict__ masks, int n,
                                                          __nv_bfloat16* __restrict__ out, int ld) {
    __shared__ float sm[2 * 27 * 27];
    const int p = blockIdx.x;
    const T* src = masks + static_cast<size_t>(p) * (2 * 27 * 27);
    for (int i = threadIdx.x; i < 2 * 27 * 27; i += blockDim.x) sm[i] = static_cast<float>(src[i]);
    __syncthreads();
    const int vec_per_row = ld >> 3;
    __nv_bfloat16* dst = out + static_cast<size_t>(p) * 196 * ld;
    if (256 % vec_per_row == 0) {
        // a thread keeps ONE column vector for all its rows: the (channel, kh, kw) decomposition of its 8 columns is done
        // once instead of per element (the divisions were ~25 integer instructions per element: the kernel was ALU-bound
        // at 29 % of its store bandwidth)
        const int v = threadIdx.x % vec_per_row, rstep = 256 / vec_per_row;
        int off[8], kh[8], kw[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = v * 8 + j;
            const int c = col / 49, r = col - c * 49;
            kh[j] = col < 98 ? r / 7 : 100;                      // 100: never inside the image -> 0
            kw[j] = r - (r / 7) * 7;
            off[j] = c * 729 + kh[j] * 27 + kw[j];
        }
        for (int row = threadIdx.x / vec_per_row; row < 196; row += rstep) {
            const int oh = row / 14, ow = row - oh * 14;
            const int h0 = oh * 2 - 3, w0 = ow * 2 - 3, base = h0 * 27 + w0;
            float vals[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const unsigned ih = static_cast<unsigned>(h0 + kh[j]), iw = static_cast<unsigned>(w0 + kw[j]);
                vals[j] = (ih < 27u && iw < 27u) ? sm[base + off[j]] : 0.f;
            }
            store_bf16x8(dst + static_cast<size_t>(row) * ld + v * 8, vals);
        }
        return;
    }
    for (int i = threadIdx.x; i < 196 * vec_per_row; i += blockDim.x) {
        const int row = i / vec_per_row, v = i - row * vec_per_row;
        const int oh = row / 14, ow = row - oh * 14;
        float vals[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = v * 8 + j;
            float x = 0.f;
            if (col < 98) {
                const int c = col / 49, r = col - c * 49;
                const int kh = r / 7, kw = r - kh * 7;
                const int ih = oh * 2 - 3 + kh, iw = ow * 2 - 3 + kw;
                if (ih >= 0 && ih < 27 && iw >= 0 && iw < 27) x = sm[c * 729 + ih * 27 + iw];
            }
            vals[j] = x;
        }
        store_bf16x8(dst + static_cast<size_t>(row) * ld + v * 8, vals);
    }
}

// ------------------------------------------------------------------------------------------------
// Segmented column statistics.  chunk table: int32 [n_chunks][3] = (row_begin, row_end, group); all
// rows of a chunk belong to one group.  sum1[g,c] += sum_r a[r,c];  sum2[g,c] += sum_r a[r,c]*b[r,c].
// block = (32 column-vectors of 8, 8 row phases); grid = (ceil(cols/256), n_chunks).
// ------------------------------------------------------------------------------------------------
template <bool A_BF16, bool B_IS_A>
__global__ void __launch_bounds__(256) seg_colstats_kernel(const void* __restrict__ a_, int lda,
                                                           const __nv_bfloat16* __restrict__ b, int ldb, int cols,
                                                           const int32_t* __restrict__ chunks,
                                                           float* __restrict__ sum1, float* __restrict__ sum2, int cvecs) {
    // cvecs = column vectors (of 8) per block: 32 for wide matrices; 16 / 8 for the 128- / 64-column activations of the
    // mask branch, so that all 256 threads load (with a fixed 32 x 8 shape half of them idled on 128 columns: the two
    // largest launches of the step ran at 58 % of the HBM peak)
    const int tx = threadIdx.x % cvecs, ty = threadIdx.x / cvecs, phases = 256 / cvecs;
    const int col = (blockIdx.x * cvecs + tx) * 8;
    const int r0 = chunks[blockIdx.y * 3 + 0], r1 = chunks[blockIdx.y * 3 + 1], g = chunks[blockIdx.y * 3 + 2];
    float s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
    if (col < cols) {
#pragma unroll 4
        for (int r = r0 + ty; r < r1; r += phases) {
            float av[8], bv[8];
            if (A_BF16) {
                load_bf16x8(reinterpret_cast<const __nv_bfloat16*>(a_) + static_cast<size_t>(r) * lda + col, av);
            } else {
                const float* ap = reinterpret_cast<const float*>(a_) + static_cast<size_t>(r) * lda + col;
                const float4 x0 = *reinterpret_cast<const float4*>(ap);
                const float4 x1 = *reinterpret_cast<const float4*>(ap + 4);
                av[0] = x0.x; av[1] = x0.y; av[2] = x0.z; av[3] = x0.w;
                av[4] = x1.x; av[5] = x1.y; av[6] = x1.z; av[7] = x1.w;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) s1[j] += av[j];
            if (B_IS_A) {
#pragma unroll
                for (int j = 0; j < 8; ++j) s2[j] = fmaf(av[j], av[j], s2[j]);
            } else if (b != nullptr) {
                load_bf16x8(b + static_cast<size_t>(r) * ldb + col, bv);
#pragma unroll
                for (int j = 0; j < 8; ++j) s2[j] = fmaf(av[j], bv[j], s2[j]);
            }
        }
    }
    __shared__ float red[2][256][9];          // [which][phase * cvecs + column vector][element]
#pragma unroll
    for (int j = 0; j < 8; ++j) { red[0][threadIdx.x][j] = s1[j]; red[1][threadIdx.x][j] = s2[j]; }
    __syncthreads();
    // reduce 2 x (cvecs * 8) columns over the row phases
    const int bcols = cvecs * 8;
    for (int o = threadIdx.x; o < 2 * bcols; o += 256) {
        const int which = o / bcols, c = o - which * bcols;
        const int cx = c >> 3, cj = c & 7;
        float s = 0.f;
        for (int y = 0; y < phases; ++y) s += red[which][y * cvecs + cx][cj];
        const int gc = blockIdx.x * bcols + c;
        if (gc < cols) {
            if (which == 0) atomicAdd(sum1 + static_cast<size_t>(g) * cols + gc, s);
            else if (sum2 != nullptr && (B_IS_A || b != nullptr)) atomicAdd(sum2 + static_cast<size_t>(g) * cols + gc, s);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// out[r,c] = k1[g,c]*a[r,c] + k2[g,c]*b[r,c] + k3[g,c]   (terms with a null pointer are skipped),
// zeroed where relu_mask && b[r,c] <= 0;  g = group_of_unit[r / rows_per_unit].
// Serves BN apply (a = null, k2 = scale, k3 = shift) and the BN+ReLU backward.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) seg_affine_kernel(const __nv_bfloat16* __restrict__ a,
                                                         const __nv_bfloat16* __restrict__ b,
                                                         const float* __restrict__ k1, const float* __restrict__ k2,
                                                         const float* __restrict__ k3,
                                                         const int32_t* __restrict__ group_of_unit, long long rows,
                                                         int rows_per_unit, int cols, int relu_mask,
                                                         __nv_bfloat16* __restrict__ out) {
    const int vec_per_row = cols >> 3;
    const long long total = rows * vec_per_row;
    const bool small = total <= 0xffffffffLL;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        long long r, unit;
        int c, rem;
        divmod_idx(i, vec_per_row, small, r, c);
        c *= 8;
        divmod_idx(r, rows_per_unit, small, unit, rem);
        const int g = __ldg(group_of_unit + unit);
        const size_t off = static_cast<size_t>(r) * cols + c;
        const size_t koff = static_cast<size_t>(g) * cols + c;
        float av[8], bv[8], o[8];
        if (a) load_bf16x8(a + off, av);
        load_bf16x8(b + off, bv);
        const float4 q0 = __ldg(reinterpret_cast<const float4*>(k2 + koff));
        const float4 q1 = __ldg(reinterpret_cast<const float4*>(k2 + koff + 4));
        const float4 t0 = __ldg(reinterpret_cast<const float4*>(k3 + koff));
        const float4 t1 = __ldg(reinterpret_cast<const float4*>(k3 + koff + 4));
        const float kk2[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
        const float kk3[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(kk2[j], bv[j], kk3[j]);
        if (a) {
            const float4 p0 = __ldg(reinterpret_cast<const float4*>(k1 + koff));
            const float4 p1 = __ldg(reinterpret_cast<const float4*>(k1 + koff + 4));
            const float kk1[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = fmaf(kk1[j], av[j], o[j]);
        }
        if (relu_mask == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = bv[j] > 0.f ? o[j] : 0.f;
        } else if (relu_mask == 2) {   // ReLU of the result itself (Linear -> BatchNorm1d -> ReLU, lib/tempura.py:103-105)
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = fmaxf(o[j], 0.f);
        }
        store_bf16x8(out + off, o);
    }
}

// ------------------------------------------------------------------------------------------------
// z[n,oh,ow,c] = max over the 3x3/s2/p1 window of (scale[g,c]*y[n,ih,iw,c] + shift[g,c]); the first
// maximum in (kh,kw) scan order wins (torch's max_pool2d rule); argmax stores kh*3+kw.
// ------------------------------------------------------------------------------------------------
template <typename TY>
__global__ void __launch_bounds__(256) bn_pool_fwd_kernel(const TY* __restrict__ y,
                                                          const float* __restrict__ scale,
                                                          const float* __restrict__ shift,
                                                          const int32_t* __restrict__ group_of_unit, int n, int hin,
                                                          int C, __nv_bfloat16* __restrict__ z,
                                                          uint8_t* __restrict__ argmax) {
    const int hout = (hin + 1) / 2;
    const int vec = C >> 3;
    const long long total = static_cast<long long>(n) * hout * hout * vec;
    const bool small = total <= 0xffffffffLL;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        long long t, t2, t3;
        int c, ow, oh;
        divmod_idx(i, vec, small, t, c);
        c *= 8;
        divmod_idx(t, hout, small, t2, ow);
        divmod_idx(t2, hout, small, t3, oh);
        const int p = static_cast<int>(t3);
        const int g = __ldg(group_of_unit + p);
        float sc[8], sh[8], best[8];
        int arg[8];
        {
            const float4 a0 = __ldg(reinterpret_cast<const float4*>(scale + static_cast<size_t>(g) * C + c));
            const float4 a1 = __ldg(reinterpret_cast<const float4*>(scale + static_cast<size_t>(g) * C + c + 4));
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(shift + static_cast<size_t>(g) * C + c));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(shift + static_cast<size_t>(g) * C + c + 4));
            sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
            sh[0] = b0.x; sh[1] = b0.y; sh[2] = b0.z; sh[3] = b0.w; sh[4] = b1.x; sh[5] = b1.y; sh[6] = b1.z; sh[7] = b1.w;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; arg[j] = 0; }
        const TY* yp = y + static_cast<size_t>(p) * hin * hin * C;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int ih = oh * 2 - 1 + kh;
            if (ih < 0 || ih >= hin) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int iw = ow * 2 - 1 + kw;
                if (iw < 0 || iw >= hin) continue;
                float v[8];
                if constexpr (sizeof(TY) == 2) {
                    load_bf16x8(reinterpret_cast<const __nv_bfloat16*>(yp) + (static_cast<size_t>(ih) * hin + iw) * C + c, v);
                } else {
                    const float* fp = reinterpret_cast<const float*>(yp) + (static_cast<size_t>(ih) * hin + iw) * C + c;
                    const float4 x0 = *reinterpret_cast<const float4*>(fp);
                    const float4 x1 = *reinterpret_cast<const float4*>(fp + 4);
                    v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w; v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float x = fmaf(sc[j], v[j], sh[j]);
                    if (x > best[j]) { best[j] = x; arg[j] = kh * 3 + kw; }
                }
            }
        }
        const size_t o = ((static_cast<size_t>(p) * hout + oh) * hout + ow) * C + c;
        store_bf16x8(z + o, best);
        uint2 packed;
        packed.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
        packed.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
        *reinterpret_cast<uint2*>(argmax + o) = packed;
    }
}

// dy[n,ih,iw,c] = sum over the <= 4 windows that contain (ih,iw) of dz where argmax points here.
__global__ void __launch_bounds__(256) pool_bwd_kernel(const __nv_bfloat16* __restrict__ dz,
                                                       const uint8_t* __restrict__ argmax, int n, int hin, int C,
                                                       __nv_bfloat16* __restrict__ dy) {
    const int hout = (hin + 1) / 2;
    const int vec = C >> 3;
    const long long total = static_cast<long long>(n) * hin * hin * vec;
    const bool small = total <= 0xffffffffLL;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        long long t, t2, t3;
        int c, iw, ih;
        divmod_idx(i, vec, small, t, c);
        c *= 8;
        divmod_idx(t, hin, small, t2, iw);
        divmod_idx(t2, hin, small, t3, ih);
        const int p = static_cast<int>(t3);
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        // windows: oh with kh = ih - 2*oh + 1 in {0,1,2}
        const int oh_lo = ih >> 1;              // kh = 1 (even ih) or kh = 2 (odd ih)
        const int oh_hi = (ih + 1) >> 1;        // == oh_lo for even ih; kh = 0 for odd ih
        const int ow_lo = iw >> 1, ow_hi = (iw + 1) >> 1;
        for (int oh = oh_lo; oh <= oh_hi; ++oh) {
            if (oh >= hout) continue;
            const int kh = ih - 2 * oh + 1;
            for (int ow = ow_lo; ow <= ow_hi; ++ow) {
                if (ow >= hout) continue;
                const int kw = iw - 2 * ow + 1;
                const int code = kh * 3 + kw;
                const size_t o = ((static_cast<size_t>(p) * hout + oh) * hout + ow) * C + c;
                const uint2 packed = *reinterpret_cast<const uint2*>(argmax + o);
                float v[8];
                load_bf16x8(dz + o, v);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t word = j < 4 ? packed.x : packed.y;
                    const int a = (word >> ((j & 3) * 8)) & 0xff;
                    if (a == code) acc[j] += v[j];
                }
            }
        }
        store_bf16x8(dy + ((static_cast<size_t>(p) * hin + ih) * hin + iw) * C + c, acc);
    }
}

// Same for even `hin`: one thread owns a 2x2 block of input positions.  The <= 4 pooling windows that touch the block
// are loaded once and serve all four positions (9 (window, position) combinations instead of 9 separate loads),
// and the index arithmetic is shared: ~2x fewer instructions per output than the position-centric kernel above.
__global__ void __launch_bounds__(256) pool_bwd2x2_kernel(const __nv_bfloat16* __restrict__ dz,
                                                          const uint8_t* __restrict__ argmax, int n, int hin, int C,
                                                          __nv_bfloat16* __restrict__ dy) {
    const int hout = hin >> 1, vec = C >> 3;
    const long long total = static_cast<long long>(n) * hout * hout * vec;
    const bool small = total <= 0xffffffffLL;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        long long t, t2, t3;
        int c, b, a;
        divmod_idx(i, vec, small, t, c);
        c *= 8;
        divmod_idx(t, hout, small, t2, b);
        divmod_idx(t2, hout, small, t3, a);
        const int p = static_cast<int>(t3);
        float acc[2][2][8];
#pragma unroll
        for (int dh = 0; dh < 2; ++dh)
#pragma unroll
            for (int dw = 0; dw < 2; ++dw)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[dh][dw][j] = 0.f;
#pragma unroll
        for (int eh = 0; eh < 2; ++eh) {
#pragma unroll
            for (int ew = 0; ew < 2; ++ew) {
                const int oh = a + eh, ow = b + ew;
                if (oh >= hout || ow >= hout) continue;
                const size_t o = ((static_cast<size_t>(p) * hout + oh) * hout + ow) * C + c;
                const uint2 packed = *reinterpret_cast<const uint2*>(argmax + o);
                float v[8];
                load_bf16x8(dz + o, v);
#pragma unroll
                for (int dh = eh; dh < 2; ++dh) {          // window row oh reaches input row 2a+dh iff eh == 0 or dh == 1
#pragma unroll
                    for (int dw = ew; dw < 2; ++dw) {
                        const int code = (dh - 2 * eh + 1) * 3 + (dw - 2 * ew + 1);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const uint32_t word = j < 4 ? packed.x : packed.y;
                            const int am = (word >> ((j & 3) * 8)) & 0xff;
                            if (am == code) acc[dh][dw][j] += v[j];
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int dh = 0; dh < 2; ++dh)
#pragma unroll
            for (int dw = 0; dw < 2; ++dw)
                store_bf16x8(dy + ((static_cast<size_t>(p) * hin + 2 * a + dh) * hin + 2 * b + dw) * C + c, acc[dh][dw]);
    }
}

// z bf16 [n,hw,hw,C] -> rows [n*hw*hw, 9*C], column (kh*3+kw)*C + c, padding 1.
__global__ void __launch_bounds__(256) im2col3x3_kernel(const __nv_bfloat16* __restrict__ z, int n, int hw, int C,
                                                        __nv_bfloat16* __restrict__ out) {
    const int vec = C >> 3;
    const long long total = static_cast<long long>(n) * hw * hw * 9 * vec;
    const bool small = total <= 0xffffffffLL;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % vec) * 8;
        long long t = i / vec;
        const int k = static_cast<int>(t % 9); t /= 9;
        const int ow = static_cast<int>(t % hw); t /= hw;
        const int oh = static_cast<int>(t % hw);
        const int p = static_cast<int>(t / hw);
        const int ih = oh - 1 + k / 3, iw = ow - 1 + k % 3;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (ih >= 0 && ih < hw && iw >= 0 && iw < hw)
            v = *reinterpret_cast<const uint4*>(z + ((static_cast<size_t>(p) * hw + ih) * hw + iw) * C + c);
        *reinterpret_cast<uint4*>(out + ((static_cast<size_t>(p) * hw + oh) * hw + ow) * (9 * C) + k * C + c) = v;
    }
}

// Same, for C / 8 in {8, 16, 32} and fewer than 2^32 / 9 rows: VEC lanes copy the C contiguous channels of one (row, tap)
// — 32-bit index arithmetic, one division by 9 and two by hw per (row, tap) instead of five 64-bit div / mod per 16 bytes
// (the kernel above was ALU-bound at 43 % of its store bandwidth).
template <int VEC>
__global__ void __launch_bounds__(256) im2col3x3_fast_kernel(const __nv_bfloat16* __restrict__ z, unsigned rows, int hw,
                                                             __nv_bfloat16* __restrict__ out) {
    constexpr int C = VEC * 8, PER = 256 / VEC;              // (row, tap) items per block iteration
    const unsigned sub = threadIdx.x / VEC, lane = threadIdx.x % VEC;
    const unsigned total = rows * 9u;
    for (unsigned j = blockIdx.x * PER + sub; j < total; j += gridDim.x * PER) {
        const unsigned row = j / 9u, k = j - row * 9u;
        const unsigned prow = row / hw, ow = row - prow * hw;           // prow = p * hw + oh
        const unsigned oh = prow % hw;
        const int ih = static_cast<int>(oh) - 1 + static_cast<int>(k / 3u), iw = static_cast<int>(ow) - 1 + static_cast<int>(k % 3u);
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (ih >= 0 && ih < hw && iw >= 0 && iw < hw)
            v = *reinterpret_cast<const uint4*>(z + (static_cast<size_t>(row) + (ih - static_cast<int>(oh)) * hw + (iw - static_cast<int>(ow))) * C + lane * 8);
        *reinterpret_cast<uint4*>(out + static_cast<size_t>(row) * (9 * C) + k * C + lane * 8) = v;
    }
}

// dz[n,h,w,c] = sum_k dcol[n, h-kh+1, w-kw+1, k*C + c]  (transpose of im2col3x3, as a gather).
__global__ void __launch_bounds__(256) col2im3x3_kernel(const __nv_bfloat16* __restrict__ dcol, int n, int hw, int C,
                                                        __nv_bfloat16* __restrict__ dz) {
    const int vec = C >> 3;
    const long long total = static_cast<long long>(n) * hw * hw * vec;
    const bool small = total <= 0xffffffffLL;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        long long t, t2, t3;
        int c, w, h;
        divmod_idx(i, vec, small, t, c);
        c *= 8;
        divmod_idx(t, hw, small, t2, w);
        divmod_idx(t2, hw, small, t3, h);
        const int p = static_cast<int>(t3);
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const int oh = h + 1 - k / 3, ow = w + 1 - k % 3;
            if (oh < 0 || oh >= hw || ow < 0 || ow >= hw) continue;
            float v[8];
            load_bf16x8(dcol + ((static_cast<size_t>(p) * hw + oh) * hw + ow) * (9 * C) + k * C + c, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += v[j];
        }
        store_bf16x8(dz + ((static_cast<size_t>(p) * hw + h) * hw + w) * C + c, acc);
    }
}

}  // namespace vsgg

using namespace vsgg;

extern "C" int b200vsgg_mask_im2col(const float* masks, int32_t n, void* out, int32_t ld, void* stream) {
    if (!masks || !out || n < 0 || ld < 104 || (ld & 7)) return set_error(B200VSGG_ERR_BAD_ARG, "mask_im2col: bad arg");
    if (n == 0) return 0;
    mask_im2col_kernel<float><<<n, 256, 0, (cudaStream_t)stream>>>(masks, n, (__nv_bfloat16*)out, ld);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_mask_im2col_bf16(const void* masks, int32_t n, void* out, int32_t ld, void* stream) {
    if (!masks || !out || n < 0 || ld < 104 || (ld & 7)) return set_error(B200VSGG_ERR_BAD_ARG, "mask_im2col_bf16: bad arg");
    if (n == 0) return 0;
    mask_im2col_kernel<__nv_bfloat16><<<n, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)masks, n,
                                                                          (__nv_bfloat16*)out, ld);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_seg_colstats(const void* a, int32_t a_is_bf16, int32_t lda, const void* b, int32_t ldb,
                                     int32_t cols, const int32_t* chunks, int32_t n_chunks, float* sum1, float* sum2,
                                     void* stream) {
    if (!a || !chunks || !sum1 || cols <= 0 || (cols & 7) || n_chunks < 0)
        return set_error(B200VSGG_ERR_BAD_ARG, "seg_colstats: bad arg (cols % 8 == 0 required)");
    if (n_chunks == 0) return 0;
    if (n_chunks > 65535) return set_error(B200VSGG_ERR_BAD_ARG, "seg_colstats: more than 65535 chunks");
    const int cvecs = cols >= 256 ? 32 : (cols > 64 ? 16 : 8);
    dim3 grid((cols + cvecs * 8 - 1) / (cvecs * 8), n_chunks);
    const __nv_bfloat16* bb = (const __nv_bfloat16*)b;
    cudaStream_t st = (cudaStream_t)stream;
    const bool same = (b == a);
    if (a_is_bf16 && same) seg_colstats_kernel<true, true><<<grid, 256, 0, st>>>(a, lda, bb, ldb, cols, chunks, sum1, sum2, cvecs);
    else if (a_is_bf16) seg_colstats_kernel<true, false><<<grid, 256, 0, st>>>(a, lda, bb, ldb, cols, chunks, sum1, sum2, cvecs);
    else if (same) seg_colstats_kernel<false, true><<<grid, 256, 0, st>>>(a, lda, bb, ldb, cols, chunks, sum1, sum2, cvecs);
    else seg_colstats_kernel<false, false><<<grid, 256, 0, st>>>(a, lda, bb, ldb, cols, chunks, sum1, sum2, cvecs);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_seg_affine(const void* a, const void* b, const float* k1, const float* k2, const float* k3,
                                   const int32_t* group_of_unit, int64_t rows, int32_t rows_per_unit, int32_t cols,
                                   int32_t relu_mask, void* out, void* stream) {
    if (!b || !k2 || !k3 || !group_of_unit || !out || (a && !k1) || cols <= 0 || (cols & 7) || rows_per_unit <= 0)
        return set_error(B200VSGG_ERR_BAD_ARG, "seg_affine: bad arg");
    if (rows == 0) return 0;
    const long long items = rows * (cols >> 3);
    int grid = blocks_for(items, 256 * 4);
    seg_affine_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, k1, k2,
                                                             k3, group_of_unit, rows, rows_per_unit, cols, relu_mask,
                                                             (__nv_bfloat16*)out);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_bn_pool_fwd(const void* y, int32_t y_is_f32, const float* scale, const float* shift,
                                    const int32_t* group_of_unit, int32_t n, int32_t hw_in, int32_t channels, void* z,
                                    uint8_t* argmax, void* stream) {
    if (!y || !scale || !shift || !group_of_unit || !z || !argmax || channels <= 0 || (channels & 7) || hw_in <= 0)
        return set_error(B200VSGG_ERR_BAD_ARG, "bn_pool_fwd: bad arg");
    if (n == 0) return 0;
    const int hout = (hw_in + 1) / 2;
    const long long items = static_cast<long long>(n) * hout * hout * (channels >> 3);
    if (y_is_f32)
        bn_pool_fwd_kernel<float><<<blocks_for(items, 256 * 2), 256, 0, (cudaStream_t)stream>>>(
            (const float*)y, scale, shift, group_of_unit, n, hw_in, channels, (__nv_bfloat16*)z, argmax);
    else
        bn_pool_fwd_kernel<__nv_bfloat16><<<blocks_for(items, 256 * 2), 256, 0, (cudaStream_t)stream>>>(
            (const __nv_bfloat16*)y, scale, shift, group_of_unit, n, hw_in, channels, (__nv_bfloat16*)z, argmax);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_pool_bwd(const void* dz, const uint8_t* argmax, int32_t n, int32_t hw_in, int32_t channels,
                                 void* dy, void* stream) {
    if (!dz || !argmax || !dy || channels <= 0 || (channels & 7) || hw_in <= 0)
        return set_error(B200VSGG_ERR_BAD_ARG, "pool_bwd: bad arg");
    if (n == 0) return 0;
    const long long items = static_cast<long long>(n) * hw_in * hw_in * (channels >> 3);
    if ((hw_in & 1) == 0)
        pool_bwd2x2_kernel<<<blocks_for(items / 4, 256), 256, 0, (cudaStream_t)stream>>>(
            (const __nv_bfloat16*)dz, argmax, n, hw_in, channels, (__nv_bfloat16*)dy);
    else
        pool_bwd_kernel<<<blocks_for(items, 256 * 2), 256, 0, (cudaStream_t)stream>>>(
            (const __nv_bfloat16*)dz, argmax, n, hw_in, channels, (__nv_bfloat16*)dy);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_im2col3x3(const void* z, int32_t n, int32_t hw, int32_t channels, void* out, void* stream) {
    if (!z || !out || channels <= 0 || (channels & 7) || hw <= 0) return set_error(B200VSGG_ERR_BAD_ARG, "im2col3x3: bad arg");
    if (n == 0) return 0;
    const long long items = static_cast<long long>(n) * hw * hw * 9 * (channels >> 3);
    const long long rows = static_cast<long long>(n) * hw * hw;
    const int vec = channels >> 3;
    if (rows * 9 < 0xffffffffLL && (vec == 8 || vec == 16 || vec == 32)) {
        const int grid = static_cast<int>(std::min<long long>((rows * 9 * vec + 255) / 256, 148LL * 64));
        if (vec == 8) im2col3x3_fast_kernel<8><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)z, (unsigned)rows, hw, (__nv_bfloat16*)out);
        else if (vec == 16) im2col3x3_fast_kernel<16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)z, (unsigned)rows, hw, (__nv_bfloat16*)out);
        else im2col3x3_fast_kernel<32><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)z, (unsigned)rows, hw, (__nv_bfloat16*)out);
        VSGG_CUDA_CHECK_LAUNCH();
        return 0;
    }
    im2col3x3_kernel<<<blocks_for(items, 256 * 4), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)z, n, hw,
                                                                                   channels, (__nv_bfloat16*)out);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_col2im3x3(const void* dcol, int32_t n, int32_t hw, int32_t channels, void* dz, void* stream) {
    if (!dcol || !dz || channels <= 0 || (channels & 7) || hw <= 0) return set_error(B200VSGG_ERR_BAD_ARG, "col2im3x3: bad arg");
    if (n == 0) return 0;
    const long long items = static_cast<long long>(n) * hw * hw * (channels >> 3);
    col2im3x3_kernel<<<blocks_for(items, 256 * 2), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dcol, n, hw,
                                                                                   channels, (__nv_bfloat16*)dz);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}
