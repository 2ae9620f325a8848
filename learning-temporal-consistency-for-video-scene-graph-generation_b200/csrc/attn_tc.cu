// Blackwell-native variable-length attention for the TokenGT encoder (tools/TokenGT/tokengt/modules/
// multihead_attention.py:135-183 and tokengt_graph_encoder_layer.py:170-191 of the reference): sequences = 5-frame
// clips (T = 2 + nodes + edges: 250-600 tokens at the AG shapes, up to ~5.3 k in the long-clip config), 32 heads x 24
// or 16 heads x 48, fp32 softmax, attention dropout.  The reference keeps [heads, T, T] maps for all 12 layers;
// nothing of size T^2 leaves the SM here.
//
// Work unit = one 128-query tile of one (clip, head); PERSISTENT CTAs (two per SM, 192 threads) walk the unit list with
// running barrier phases, so the pipeline never drains between units:
//   warp 0    TMA producer: Q once, then a 2-stage ring of K tiles and one of V tiles.  The tensor maps are
//             3-D {head_dim, token row, head} views of the [rows, heads*head_dim] activations with a {64, 128, 1} box:
//             the columns head_dim..63 of a box lie outside dimension 0 and arrive as ZEROS, so a 24- or 48-wide
//             head lands directly in the 128-byte SWIZZLE_128B rows that tcgen05.mma consumes.
//   warp 1    MMA issuer (one thread): S = Q K^T  (M128 x N128, K = head_dim rounded to 16) into TMEM columns
//             [0,128); O += P V (M128 x N{32,64}, K = 128 keys) with P read FROM TENSOR MEMORY (tcgen05.mma, A in
//             TMEM) and V as an MN-major shared-memory operand — the same 16 KB tile format as K.
//   warps 2-5 softmax: thread = query row (tcgen05.ld 32x32b: lane = row), so the row maximum and the row sum are
//             thread-local — no shuffles.  The whole S row (128 fp32) is pulled into registers, the S buffer is
//             released at once (the next S MMA overlaps this block's exponentials), P goes back to TMEM as packed
//             bf16 (tcgen05.st) in columns [128,192).  The running maximum is only raised when it grew by more than
//             2^8 (lazy rescaling): O in TMEM columns [192,256) is rescaled by the row's own thread in that case.
// Pipeline barriers: q_full, k_full/k_empty[2], v_full/v_empty[2], s_full (MMA -> softmax), s_free, p_full
// (softmax -> MMA, 128 arrivals), o_done (tcgen05.commit after every P V).
//
// Roofline: 4 T^2 hd flops per (clip, head); per 128 x 128 tile the tensor pipe needs ~128 + 128 cycles, the 16 k
// exponentials need 1024 cycles of the SM's 16/clk MUFU: the kernel is SFU-bound at head_dim 24/48 (ceiling
// ~430 / ~860 TFLOP/s of algorithmic flops), not tensor- or HBM-bound.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>

#include "attn_tc_common.cuh"
#include <cstdlib>

namespace vsgg {
namespace atc {

constexpr int KV_STAGES = 2;
constexpr int threads_of(int nh) { return 64 + 128 * nh; }   // TMA warp, MMA warp, 4 * nh softmax warps
constexpr int TMEM_COLS = 256;
constexpr int COL_S = 0, COL_P = 128, COL_O = 192;
constexpr int XCH_BYTES = 2 * 2 * 128 * 4 + 128 * 4;   // row-maximum exchange (double-buffered, two halves) + row sums
constexpr int SMEM_BYTES = TILE_BYTES * (2 + 2 * KV_STAGES) + 256 + XCH_BYTES + 1024;   // tiles + barriers + exchange + slack
constexpr float RESCALE_THRESHOLD = 8.f;   // log2 domain: P stays below 2^8, exact in fp32 sums and fine in bf16

// Work-unit metadata (clip bounds, first query row): three DEPENDENT global loads whose latency (~1 k cycles) was exposed
// at the top of every unit in every role — a unit of a 250-600-token clip is only 2..5 key blocks long.  Each role now
// fetches the NEXT unit's metadata while it works on the current one.
struct UnitMeta {
    int s0, s1, qrow0, head;
};
__device__ __forceinline__ UnitMeta load_unit_meta(int u, int n_units, int n_heads, const int32_t* __restrict__ seq_off,
                                                   const int32_t* __restrict__ blk_seq, const int32_t* __restrict__ blk_row0) {
    UnitMeta m{0, 0, 0, 0};
    if (u < n_units) {
        const int blk = u / n_heads;
        m.head = u - blk * n_heads;
        const int seq = __ldg(blk_seq + blk);
        m.qrow0 = __ldg(blk_row0 + blk);
        m.s0 = __ldg(seq_off + seq);
        m.s1 = __ldg(seq_off + seq + 1);
    }
    return m;
}

// HDN: accumulator width of O, 32 (head_dim <= 32) or 64.  NH: softmax threads per query row (1: a thread owns the whole
// 128-key row of S; 2: two threads own 64 keys each — see the softmax section).
template <int HDN, int NH>
__global__ void __launch_bounds__(threads_of(NH), 2)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                   const __grid_constant__ CUtensorMap tv, const int32_t* __restrict__ seq_off,
                   const int32_t* __restrict__ blk_seq, const int32_t* __restrict__ blk_row0, int n_units, int n_heads,
                   int hd, float scale_log2, __nv_bfloat16* __restrict__ ctx, int ldc, float* __restrict__ lse,
                   float drop_p, unsigned long long seed) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* Qs = smem;                                  // [2]: the next unit's queries land while this one computes
    uint8_t* Ks = Qs + 2 * TILE_BYTES;
    uint8_t* Vs = Ks + KV_STAGES * TILE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(Vs + KV_STAGES * TILE_BYTES);
    uint64_t* q_full = bars;                             // [2]
    uint64_t* q_empty = q_full + 2;                      // [2]
    uint64_t* k_full = q_empty + 2;
    uint64_t* k_empty = k_full + KV_STAGES;
    uint64_t* v_full = k_empty + KV_STAGES;
    uint64_t* v_empty = v_full + KV_STAGES;
    uint64_t* s_full = v_empty + KV_STAGES;
    uint64_t* s_free = s_full + 1;
    uint64_t* p_full = s_free + 1;
    uint64_t* o_done = p_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 1);
    float* xmax = reinterpret_cast<float*>(bars + 32);     // [2 (buffer)][2 (half)][128 rows]
    float* xsum = xmax + 2 * 2 * 128;                       // [128 rows]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tq);
        ptx::prefetch_tmap(&tk);
        ptx::prefetch_tmap(&tv);
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&q_full[i], 1);
            ptx::mbar_init(&q_empty[i], 1);
        }
        for (int i = 0; i < KV_STAGES; ++i) {
            ptx::mbar_init(&k_full[i], 1);
            ptx::mbar_init(&k_empty[i], 1);
            ptx::mbar_init(&v_full[i], 1);
            ptx::mbar_init(&v_empty[i], 1);
        }
        ptx::mbar_init(s_full, 1);
        ptx::mbar_init(s_free, 128 * NH);
        ptx::mbar_init(p_full, 128 * NH);
        ptx::mbar_init(o_done, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // PERSISTENT: every role walks the same unit sequence u = blockIdx.x, + gridDim.x, ... (unit = 128-query tile x head;
    // consecutive units = the heads of one tile, so a clip's K/V rows stay hot in L2) with RUNNING barrier phases: the
    // producer is already fetching the next unit's Q / K / V and the MMA thread has issued its first S while the softmax
    // warps still write this unit's output — no pipeline drain, TMEM allocation or barrier setup per tile.
    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            uint32_t qc = 0, kc = 0;                          // units / key tiles issued so far
            UnitMeta nxt = load_unit_meta(blockIdx.x, n_units, n_heads, seq_off, blk_seq, blk_row0);
            for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++qc) {
                const UnitMeta cur = nxt;
                nxt = load_unit_meta(u + gridDim.x, n_units, n_heads, seq_off, blk_seq, blk_row0);
                const int head = cur.head, s0 = cur.s0, s1 = cur.s1;
                const int nkb = (s1 - s0 + BKV - 1) / BKV;
                const uint32_t qs = qc & 1u;
                ptx::mbar_wait(&q_empty[qs], ((qc >> 1) & 1u) ^ 1u);
                ptx::mbar_expect_tx(&q_full[qs], TILE_BYTES);
                ptx::tma_load_3d(Qs + qs * TILE_BYTES, &tq, &q_full[qs], 0, cur.qrow0, head);
                for (int j = 0; j < nkb; ++j, ++kc) {
                    const uint32_t st = kc % KV_STAGES, ph = (kc / KV_STAGES) & 1u;
                    ptx::mbar_wait(&k_empty[st], ph ^ 1u);
                    ptx::mbar_expect_tx(&k_full[st], TILE_BYTES);
                    ptx::tma_load_3d(Ks + st * TILE_BYTES, &tk, &k_full[st], 0, s0 + j * BKV, head);
                    ptx::mbar_wait(&v_empty[st], ph ^ 1u);
                    ptx::mbar_expect_tx(&v_full[st], TILE_BYTES);
                    ptx::tma_load_3d(Vs + st * TILE_BYTES, &tv, &v_full[st], 0, s0 + j * BKV, head);
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            constexpr uint32_t idesc_s = ptx::make_idesc_bf16(BQ, BKV, 0, 0);      // Q, K: K-major
            constexpr uint32_t idesc_o = ptx::make_idesc_bf16(BQ, HDN, 0, 1);      // P: TMEM (K-major), V: MN-major
            const int ks_s = (hd + 15) >> 4;                                       // k-steps over head_dim
            const uint32_t t_s = tmem_base + COL_S, t_p = tmem_base + COL_P, t_o = tmem_base + COL_O;
            uint32_t qc = 0, g0 = 0;                          // units done, key tiles done before this unit
            UnitMeta nxt = load_unit_meta(blockIdx.x, n_units, n_heads, seq_off, blk_seq, blk_row0);
            for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++qc) {
                const UnitMeta cur = nxt;
                nxt = load_unit_meta(u + gridDim.x, n_units, n_heads, seq_off, blk_seq, blk_row0);
                const int s0 = cur.s0, s1 = cur.s1;
                const int nkb = (s1 - s0 + BKV - 1) / BKV;
                const uint32_t qs = qc & 1u;
                const uint32_t qa = ptx::smem_u32(Qs + qs * TILE_BYTES);
                auto issue_s = [&](int j) {
                    const uint32_t g = g0 + j, st = g % KV_STAGES;
                    ptx::mbar_wait(&k_full[st], (g / KV_STAGES) & 1u);
                    if (g > 0) ptx::mbar_wait(s_free, (g - 1) & 1u);          // the previous S tile sits in registers
                    ptx::tc_fence_after();
                    const uint32_t kb = ptx::smem_u32(Ks + st * TILE_BYTES);
                    for (int k = 0; k < ks_s; ++k)
                        ptx::umma_bf16(t_s, ptx::make_smem_desc_sw128(qa + k * 32, 16, 1024),
                                       ptx::make_smem_desc_sw128(kb + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
                    ptx::umma_commit(&k_empty[st]);
                    if (j == nkb - 1) ptx::umma_commit(&q_empty[qs]);        // last product that reads this Q tile
                    ptx::umma_commit(s_full);
                };
                ptx::mbar_wait(&q_full[qs], (qc >> 1) & 1u);
                issue_s(0);
                for (int j = 0; j < nkb; ++j) {
                    if (j + 1 < nkb) issue_s(j + 1);
                    const uint32_t g = g0 + j, st = g % KV_STAGES;
                    const int kvalid = min(BKV, s1 - (s0 + j * BKV));
                    const int ks_o = (kvalid + 15) >> 4;                           // k-steps over the block's keys
                    ptx::mbar_wait(&v_full[st], (g / KV_STAGES) & 1u);
                    ptx::mbar_wait(p_full, g & 1u);
                    ptx::tc_fence_after();
                    const uint32_t vb = ptx::smem_u32(Vs + st * TILE_BYTES);
                    for (int k = 0; k < ks_o; ++k)
                        ptx::umma_bf16_ts(t_o, t_p + k * 8, ptx::make_smem_desc_sw128(vb + k * 2048, 8192, 1024), idesc_o,
                                          (j != 0 || k != 0) ? 1u : 0u);
                    ptx::umma_commit(&v_empty[st]);
                    ptx::umma_commit(o_done);
                }
                g0 += nkb;
            }
        }
    } else {
        // ================ softmax warps: NH threads per query row, 128 / NH of the tile's keys each ================
        // NH = 1: thread = query row (tcgen05.ld 32x32b: lane = row), the row maximum and the row sum are thread-local.
        // NH = 2: warps w and w + 4 may access the same TMEM lane quadrant (w % 4): the first four softmax warps own
        // columns [0,64) of S, the other four [64,128); the row maximum is exchanged through shared memory (double-
        // buffered, one 64-thread named barrier per block), the row sums are combined once per unit.  Measured on B200
        // (tools/attn_tc_profile.py): 32 heads x 24 — 2.28 -> 2.16 ms with dropout, 2.13 -> 1.89 ms without (the 168-
        // register single thread per row left the issue slots half empty); 16 heads x 48 — the O rescale / epilogue
        // state of the 64-wide accumulator spills at the 96-register budget of NH = 2 and the kernel gets SLOWER
        // (4.85 -> 6.1 ms per layer of the SGCls config), so that width keeps NH = 1.
        constexpr int CW = 128 / NH;                    // score columns per thread
        const int quad = warp & 3;                      // TMEM lane quadrant this warp may access
        const int ch = (warp - 2) >> 2;                 // column part (0 for NH = 1)
        const int row = quad * 32 + lane;               // row of the tile
        const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
        const uint32_t t_s = tmem_base + lane_addr + COL_S + ch * CW, t_p = tmem_base + lane_addr + COL_P + ch * (CW / 2),
                       t_o = tmem_base + lane_addr + COL_O;
        // attention dropout (attn_dropout.cuh): one LCG draw per four keys, SWAR compare, the keep masks are ANDed into
        // the packed bf16 pairs; the 1/(1-p) of the survivors is folded into the final normalisation of O
        const uint32_t thr = adrop::thr8_of(drop_p);
        const float inv_keep = thr ? adrop::inv_keep_of(thr) : 1.f;
        const uint32_t K8 = (256u - thr) * 0x00010001u;
        auto pair_sync = [&]() {
            if (NH == 2) asm volatile("bar.sync %0, 64;" ::"r"(quad + 1) : "memory");
        };
        uint32_t g = 0;                                 // key tiles processed so far (barrier phases)
        UnitMeta nxt = load_unit_meta(blockIdx.x, n_units, n_heads, seq_off, blk_seq, blk_row0);
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const UnitMeta cur = nxt;
            nxt = load_unit_meta(u + gridDim.x, n_units, n_heads, seq_off, blk_seq, blk_row0);
            const int head = cur.head, s0 = cur.s0, s1 = cur.s1, qrow0 = cur.qrow0;
            const int qrows = min(BQ, s1 - qrow0);
            const int nkb = (s1 - s0 + BKV - 1) / BKV;
            const uint32_t row_key = thr ? adrop::row_key(seed, qrow0 + row, head) : 0u;
            float m_ref = 0.f, l = 0.f;
            for (int j = 0; j < nkb; ++j, ++g) {
                const int kvalid = min(BKV, s1 - (s0 + j * BKV)) - ch * CW;      // valid keys among this thread's columns
                uint32_t r[CW];
                ptx::mbar_wait(s_full, g & 1u);
                ptx::tc_fence_after();
#pragma unroll
                for (int c = 0; c < CW / 32; ++c)
                    ptx::tmem_ld_32x32b_x32(t_s + c * 32, reinterpret_cast<uint32_t(&)[32]>(r[c * 32]));
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                ptx::mbar_arrive(s_free);                   // the next S MMA may overwrite the buffer
                float mx = -INFINITY;
                if (kvalid >= CW) {
#pragma unroll
                    for (int i = 0; i < CW / 2; ++i)    // three-input maximum (FMNMX3): one issue slot per two scores
                        mx = fmaxf(fmaxf(mx, __uint_as_float(r[2 * i])), __uint_as_float(r[2 * i + 1]));
                } else {
#pragma unroll
                    for (int i = 0; i < CW; ++i) {
                        if (i >= kvalid) r[i] = 0xff800000u;   // -inf: keys beyond the clip
                        mx = fmaxf(mx, __uint_as_float(r[i]));
                    }
                }
                if (NH == 2) {      // full-row maximum: exchange with the thread that owns the other 64 columns of this row
                    float* xb = xmax + (g & 1u) * 256;
                    xb[ch * 128 + row] = mx;
                    pair_sync();
                    mx = fmaxf(mx, xb[(ch ^ 1) * 128 + row]);
                }
                mx *= scale_log2;
                float alpha = 1.f;
                bool rescale = false;
                if (j == 0) {
                    m_ref = mx;
                } else if (mx > m_ref + RESCALE_THRESHOLD) {
                    alpha = ptx::ex2_approx(m_ref - mx);
                    m_ref = mx;
                    rescale = true;
                }
                // exponent arguments and the row sum run on the packed fp32x2 pipes (FFMA2 / FADD2): per PAIR of scores
                // one FFMA2, two MUFU.EX2, one FADD2, one pack
                float sum = 0.f, sum1 = 0.f;
                const float nm = -m_ref;
                uint32_t pk[CW / 2];
                if (thr == 0u) {
#pragma unroll
                    for (int i = 0; i < CW / 2; ++i) {
                        float a0, a1;
                        ptx::fma2(a0, a1, __uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]), scale_log2, scale_log2, nm, nm);
                        const float p0 = ptx::ex2_approx(a0), p1 = ptx::ex2_approx(a1);
                        ptx::add2(sum, sum1, sum, sum1, p0, p1);
                        pk[i] = pack2(p0, p1);
                    }
                } else {
                    // one stream per key block of 64 relative to the clip start, whichever thread draws from it
                    uint32_t sdv[CW / 64];
#pragma unroll
                    for (int c = 0; c < CW / 64; ++c)
                        sdv[c] = adrop::stream_seed(row_key, static_cast<uint32_t>(j) * 2u + static_cast<uint32_t>(ch * (CW / 64) + c));
#pragma unroll
                    for (int g4 = 0; g4 < CW / 4; ++g4) {                      // groups of four keys
                        const uint32_t sd = sdv[g4 >> 4];
                        float p[4], a[4];
                        ptx::fma2(a[0], a[1], __uint_as_float(r[4 * g4]), __uint_as_float(r[4 * g4 + 1]), scale_log2, scale_log2, nm, nm);
                        ptx::fma2(a[2], a[3], __uint_as_float(r[4 * g4 + 2]), __uint_as_float(r[4 * g4 + 3]), scale_log2, scale_log2, nm, nm);
#pragma unroll
                        for (int e = 0; e < 4; ++e) p[e] = ptx::ex2_approx(a[e]);
                        ptx::add2(sum, sum1, sum, sum1, p[0], p[1]);
                        ptx::add2(sum, sum1, sum, sum1, p[2], p[3]);
                        uint32_t lo, hi;
                        adrop::keep_masks4(adrop::draw(sd, (g4 & 15) >> 1, g4 & 1), K8, lo, hi);
                        pk[2 * g4] = pack2(p[0], p[1]) & lo;
                        pk[2 * g4 + 1] = pack2(p[2], p[3]) & hi;
                    }
                }
                l = l * alpha + (sum + sum1);               // this half's share of the row sum
                if (j > 0) {                                // P and O are free once the previous P V has retired
                    ptx::mbar_wait(o_done, (g - 1) & 1u);
                    ptx::tc_fence_after();
                    if (ch == 0 && __any_sync(0xffffffffu, rescale)) {          // one of the row's two threads rescales O
                        uint32_t o[HDN];
#pragma unroll
                        for (int c = 0; c < HDN / 32; ++c)
                            ptx::tmem_ld_32x32b_x32(t_o + c * 32, reinterpret_cast<uint32_t(&)[32]>(o[c * 32]));
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < HDN; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
#pragma unroll
                        for (int c = 0; c < HDN / 32; ++c)
                            ptx::tmem_st_32x32b_x32(t_o + c * 32, reinterpret_cast<const uint32_t(&)[32]>(o[c * 32]));
                    }
                }
                // (j == 0: the previous unit's epilogue below already waited for its last P V)
#pragma unroll
                for (int c = 0; c < CW / 64; ++c)
                    ptx::tmem_st_32x32b_x32(t_p + c * 32, reinterpret_cast<const uint32_t(&)[32]>(pk[c * 32]));
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
                ptx::mbar_arrive(p_full);
            }
            // ---- epilogue: O / l -> bf16 context rows, log-sum-exp for the backward (the half-0 thread of each row; the
            //      other half hands over its share of the row sum).  The next unit's first P V comes after BOTH threads'
            //      next p_full arrival, i.e. after these loads of O.
            if (NH == 2 && ch == 1) xsum[row] = l;
            pair_sync();
            ptx::mbar_wait(o_done, (g - 1) & 1u);
            ptx::tc_fence_after();
            if (ch == 0) {
                if (NH == 2) l += xsum[row];
                uint32_t o[HDN];
#pragma unroll
                for (int c = 0; c < HDN / 32; ++c)
                    ptx::tmem_ld_32x32b_x32(t_o + c * 32, reinterpret_cast<uint32_t(&)[32]>(o[c * 32]));
                ptx::tmem_ld_wait();
                if (row < qrows) {
                    const float inv = inv_keep / l;
                    const size_t grow = static_cast<size_t>(qrow0 + row);
                    __nv_bfloat16* dst = ctx + grow * ldc + head * hd;
#pragma unroll
                    for (int c = 0; c < HDN / 8; ++c) {
                        if (c * 8 < hd) {
                            uint4 v4;
                            v4.x = pack2(__uint_as_float(o[c * 8]) * inv, __uint_as_float(o[c * 8 + 1]) * inv);
                            v4.y = pack2(__uint_as_float(o[c * 8 + 2]) * inv, __uint_as_float(o[c * 8 + 3]) * inv);
                            v4.z = pack2(__uint_as_float(o[c * 8 + 4]) * inv, __uint_as_float(o[c * 8 + 5]) * inv);
                            v4.w = pack2(__uint_as_float(o[c * 8 + 6]) * inv, __uint_as_float(o[c * 8 + 7]) * inv);
                            *reinterpret_cast<uint4*>(dst + c * 8) = v4;
                        }
                    }
                    if (lse != nullptr) lse[grow * n_heads + head] = (m_ref + __log2f(l)) * 0.6931471805599453f;
                }
            }
            pair_sync();        // xsum[row] may be rewritten by the next unit only after it was read
        }
    }
    // ================================ teardown ================================
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace atc
}  // namespace vsgg

using namespace vsgg;

extern "C" int b200vsgg_attn_tc_fwd(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv,
                                    int32_t rows, const int32_t* seq_off, const int32_t* blk_seq, const int32_t* blk_row0,
                                    int32_t n_blocks, int32_t n_heads, int32_t head_dim, float scale, void* ctx, int32_t ldc,
                                    float* lse, float drop_p, uint64_t seed, void* stream) {
    if (!q || !k || !v || !seq_off || !blk_seq || !blk_row0 || !ctx || rows <= 0 || n_heads <= 0 || n_heads > 64 ||
        head_dim < 8 || head_dim > 64 || (head_dim & 7) || (ldc & 7) || (reinterpret_cast<uintptr_t>(ctx) & 15u))
        return set_error(B200VSGG_ERR_BAD_ARG, "attn_tc_fwd: bad arg (head_dim in 8..64, % 8 == 0; <= 64 heads)");
    if (n_blocks == 0) return 0;
    CUtensorMap tq, tk, tv;
    int rc;
    if ((rc = atc::make_tmap_heads(&tq, q, head_dim, rows, n_heads, ldq))) return rc;
    if ((rc = atc::make_tmap_heads(&tk, k, head_dim, rows, n_heads, ldk))) return rc;
    if ((rc = atc::make_tmap_heads(&tv, v, head_dim, rows, n_heads, ldv))) return rc;
    const float scale_log2 = scale * 1.4426950408889634f;
    const long long units = static_cast<long long>(n_blocks) * n_heads;
    if (units > 0x7fffffffLL) return set_error(B200VSGG_ERR_BAD_ARG, "attn_tc_fwd: too many (tile, head) units");
    static bool attr_set = false;
    if (!attr_set) {     // both instantiations share one function-pointer type: set the attribute on each explicitly
        cudaError_t e = cudaFuncSetAttribute(atc::attn_tc_fwd_kernel<32, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             atc::SMEM_BYTES);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(atc::attn_tc_fwd_kernel<64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     atc::SMEM_BYTES);
        if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
        attr_set = true;
    }
    auto launch = [&](auto kern, int threads) -> int {
        // Persistent (two CTAs per SM walk the unit list) when the softmax warps carry the dropout work; one unit per
        // CTA otherwise.  Measured on B200 (tools/flash_bench.py, 32 heads x 24): training p = 0.1  1.90 -> 1.78 ms at the
        // AG shapes; inference p = 0  1.50 ms per-unit vs 1.76 ms persistent (2.97 vs 3.38 ms on long clips) — with the
        // cheaper softmax the hardware's dynamic CTA dispatch balances the ragged units better than the static stride.
        // B200VSGG_ATTN_PERSIST=0/1 overrides (A/B timing).
        static const int force = []() { const char* e = getenv("B200VSGG_ATTN_PERSIST"); return e ? (e[0] == '0' ? 0 : 1) : -1; }();
        const bool persist = force >= 0 ? force == 1 : drop_p > 0.f;
        const long long resident = persist ? 2LL * num_sms() : units;
        kern<<<static_cast<unsigned>(units < resident ? units : resident), threads, atc::SMEM_BYTES, (cudaStream_t)stream>>>(
            tq, tk, tv, seq_off, blk_seq, blk_row0, static_cast<int>(units), n_heads, head_dim, scale_log2,
            reinterpret_cast<__nv_bfloat16*>(ctx), ldc, lse, drop_p, seed);
        return 0;
    };
    rc = head_dim <= 32 ? launch(atc::attn_tc_fwd_kernel<32, 2>, atc::threads_of(2)) : launch(atc::attn_tc_fwd_kernel<64, 1>, atc::threads_of(1));
    if (rc) return rc;
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}
