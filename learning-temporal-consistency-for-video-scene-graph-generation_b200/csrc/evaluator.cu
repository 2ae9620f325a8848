// Recall@K matching of one video on the device: the per-frame body of
// tools/utils/evaluation_recall.py:119-276 (evaluate_from_dict -> evaluate_recall -> _compute_pred_matches) for the three
// constraint modes, one CTA per frame.  The reference moves ~20 tensors to numpy per frame and loops over ground-truth
// triplets in Python; here a frame is: candidate triplets in the reference's order -> float64 scores with the reference's
// dtypes (float32 object-score product, float64 everything else) -> total order by (score desc, position desc = what a
// stable ascending argsort reversed gives) -> for every ground-truth relation "is there a matching candidate among the
// first K" for K = 10 / 20 / 50 / 100 (class triplet equal, both IoUs >= threshold, +1 pixel convention of bbox_overlaps).
// Integer / double work, a few kB per frame: bounded by latency, not by any roofline; the point is ONE launch and ONE
// small read-back per video instead of ~20 host synchronisations per frame.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>

#include "../../include/b200vsgg.h"
#include "common.cuh"

namespace vsgg {

constexpr int EV_THREADS = 256;
constexpr int EV_MAX_CAND = 3328;      // 3 * 42 pairs * 26 predicates rounded: candidates of a frame kept in shared memory
constexpr int EV_TOP = 100;            // evaluate_from_dict keeps the 100 best (relation, predicate) entries ("no" constraint)

struct EvalFrame {
    const long long* pair_idx;   // [N,2] box rows
    const float *att, *spa, *con;
    int na, ns, nc;
    int p0, n;                   // this frame's pairs: rows [p0, p0+n)
};

// rel_scores[r][c] of the reference's [3n, na+ns+nc] float64 matrix (float32 blocks next to float64 zero blocks)
__device__ __forceinline__ double ev_rel_score(const EvalFrame& f, int r, int c) {
    const int blk = r / f.n, i = f.p0 + (r - blk * f.n);
    if (blk == 0) return (c < f.na) ? static_cast<double>(f.att[static_cast<size_t>(i) * f.na + c]) : 0.0;
    if (blk == 1) return (c >= f.na && c < f.na + f.ns) ? static_cast<double>(f.spa[static_cast<size_t>(i) * f.ns + c - f.na]) : 0.0;
    return (c >= f.na + f.ns) ? static_cast<double>(f.con[static_cast<size_t>(i) * f.nc + c - f.na - f.ns]) : 0.0;
}
// (subject box row, object box row) of relation row r: attention / contacting rows are (person, object), spatial rows
// are reversed (evaluation_recall.py:171: pair_idx[:, ::-1])
__device__ __forceinline__ void ev_rel_inds(const EvalFrame& f, int r, int& sub, int& obj) {
    const int blk = r / f.n, i = f.p0 + (r - blk * f.n);
    const int a = static_cast<int>(f.pair_idx[2 * static_cast<size_t>(i)]), b = static_cast<int>(f.pair_idx[2 * static_cast<size_t>(i) + 1]);
    sub = (blk == 1) ? b : a;
    obj = (blk == 1) ? a : b;
}
__device__ __forceinline__ double ev_iou(const double* a, const double* b) {
    const double iw = fmin(a[2], b[2]) - fmax(a[0], b[0]) + 1.0;
    const double ih = fmin(a[3], b[3]) - fmax(a[1], b[1]) + 1.0;
    if (!(iw > 0.0 && ih > 0.0)) return 0.0;
    const double inter = iw * ih;
    const double ua = (a[2] - a[0] + 1.0) * (a[3] - a[1] + 1.0) + (b[2] - b[0] + 1.0) * (b[3] - b[1] + 1.0) - inter;
    return inter / ua;
}

// mode 0: "with" constraint (one predicate per relation row: argmax), 1: "no" (top-100 of score x object scores),
// 2: "semi" (attention rows: argmax; other rows: every predicate above the threshold)
__global__ void __launch_bounds__(EV_THREADS)
eval_recall_kernel(const long long* __restrict__ pair_idx, const int32_t* __restrict__ frame_off, const float* __restrict__ att,
                   int na, const float* __restrict__ spa, int ns, const float* __restrict__ con, int nc,
                   const float* __restrict__ pred_boxes, int box_ld, const long long* __restrict__ pred_classes,
                   const float* __restrict__ obj_scores, const double* __restrict__ gt_boxes,
                   const int32_t* __restrict__ gt_classes, const int32_t* __restrict__ gt_box_off,
                   const int32_t* __restrict__ gt_rels, const int32_t* __restrict__ gt_rel_off, int mode, double semi_thr,
                   double iou_thr, uint8_t* __restrict__ hits, int32_t* __restrict__ status) {
    __shared__ double key[EV_MAX_CAND];          // sort key of the current stage
    __shared__ uint16_t c_row[EV_MAX_CAND];      // candidate: relation row
    __shared__ uint8_t c_col[EV_MAX_CAND];       // candidate: predicate column
    __shared__ uint16_t s_row[EV_TOP + 28];      // sorted / selected candidates (<= 128)
    __shared__ uint8_t s_col[EV_TOP + 28];
    __shared__ int n_cand;
    __shared__ int row_cnt[128], row_off[129];
    const int fr = blockIdx.x, tid = threadIdx.x;
    EvalFrame f{pair_idx, att, spa, con, na, ns, nc, frame_off[fr], frame_off[fr + 1] - frame_off[fr]};
    const int R = 3 * f.n, C = na + ns + nc;
    const int g0 = gt_rel_off[fr], G = gt_rel_off[fr + 1] - g0;
    if (f.n <= 0 || R > 126 || R * C > EV_MAX_CAND) {
        if (tid == 0 && f.n > 0) atomicExch(status, 1);            // frame too large for the on-chip tables
        for (int g = tid; g < G * 4; g += EV_THREADS) hits[static_cast<size_t>(g0) * 4 + g] = 0;
        return;
    }
    // ---------------------------------------------------------------- candidates in the reference's order
    if (mode == 1) {
        // overall = float32(obj_s[sub] * obj_s[obj]) * rel_scores, all R*C entries; keep the 100 largest
        for (int e = tid; e < R * C; e += EV_THREADS) {
            const int r = e / C, c = e - r * C;
            int sub, obj;
            ev_rel_inds(f, r, sub, obj);
            const float per_rel = obj_scores[sub] * obj_scores[obj];
            key[e] = static_cast<double>(per_rel) * ev_rel_score(f, r, c);
        }
        __syncthreads();
        const int total = R * C, keep = min(EV_TOP, total);
        for (int e = tid; e < total; e += EV_THREADS) {
            const double k = key[e];
            int rank = 0;                           // position in argsort(-overall): larger first, ties by flat index
            for (int j = 0; j < total; ++j) {
                const double kj = key[j];
                rank += (kj > k) || (kj == k && j < e);
            }
            if (rank < keep) {
                s_row[rank] = static_cast<uint16_t>(e / C);
                s_col[rank] = static_cast<uint8_t>(e - (e / C) * C);
            }
        }
        __syncthreads();
        if (tid == 0) n_cand = keep;
        for (int e = tid; e < keep; e += EV_THREADS) { c_row[e] = s_row[e]; c_col[e] = s_col[e]; }
        __syncthreads();
    } else {
        // per relation row: how many candidates it contributes, then an exclusive scan keeps the row order
        for (int r = tid; r < R; r += EV_THREADS) {
            int cnt = 1;
            if (mode == 2) {
                const bool att_row = ev_rel_score(f, r, 0) + ev_rel_score(f, r, 1) > 0.0;
                if (!att_row) {
                    const bool other = (ev_rel_score(f, r, 3) + ev_rel_score(f, r, 4) > 0.0) ||
                                       (ev_rel_score(f, r, 9) + ev_rel_score(f, r, 10) > 0.0);
                    cnt = 0;
                    if (other)
                        for (int c = 0; c < C; ++c) cnt += ev_rel_score(f, r, c) > semi_thr;
                }
            }
            row_cnt[r] = cnt;
        }
        __syncthreads();
        if (tid == 0) {
            int acc = 0;
            for (int r = 0; r < R; ++r) { row_off[r] = acc; acc += row_cnt[r]; }
            row_off[R] = acc;
            n_cand = acc;
        }
        __syncthreads();
        for (int r = tid; r < R; r += EV_THREADS) {
            int o = row_off[r];
            if (row_cnt[r] == 0) continue;
            const bool argmax_row = (mode == 0) || (ev_rel_score(f, r, 0) + ev_rel_score(f, r, 1) > 0.0);
            if (argmax_row) {
                int best = 0;
                double bv = ev_rel_score(f, r, 0);
                for (int c = 1; c < C; ++c) {
                    const double v = ev_rel_score(f, r, c);
                    if (v > bv) { bv = v; best = c; }                 // np.argmax: first maximum
                }
                c_row[o] = static_cast<uint16_t>(r);
                c_col[o] = static_cast<uint8_t>(best);
            } else {
                for (int c = 0; c < C; ++c)
                    if (ev_rel_score(f, r, c) > semi_thr) {
                        c_row[o] = static_cast<uint16_t>(r);
                        c_col[o] = static_cast<uint8_t>(c);
                        ++o;
                    }
            }
        }
        __syncthreads();
    }
    const int nC = n_cand;
    // ---------------------------------------------------------------- triplet scores (float64 product of the three columns)
    for (int e = tid; e < nC; e += EV_THREADS) {
        int sub, obj;
        ev_rel_inds(f, c_row[e], sub, obj);
        key[e] = (static_cast<double>(obj_scores[sub]) * static_cast<double>(obj_scores[obj])) * ev_rel_score(f, c_row[e], c_col[e]);
    }
    __syncthreads();
    // order = argsort(scores)[::-1]: descending, ties in descending position.  Only the first 100 places matter.
    const int nS = min(nC, EV_TOP);
    for (int e = tid; e < nC; e += EV_THREADS) {
        const double k = key[e];
        int rank = 0;
        for (int j = 0; j < nC; ++j) {
            const double kj = key[j];
            rank += (kj > k) || (kj == k && j > e);
        }
        if (rank < nS) { s_row[rank] = c_row[e]; s_col[rank] = c_col[e]; }
    }
    __syncthreads();
    // ---------------------------------------------------------------- matching: one thread per ground-truth relation
    const int b0 = gt_box_off[fr];
    for (int g = tid; g < G; g += EV_THREADS) {
        const int gs = gt_rels[3 * (g0 + g)], go = gt_rels[3 * (g0 + g) + 1], gp = gt_rels[3 * (g0 + g) + 2];
        const int cls_s = gt_classes[b0 + gs], cls_o = gt_classes[b0 + go];
        const double* gbs = gt_boxes + 4 * static_cast<size_t>(b0 + gs);
        const double* gbo = gt_boxes + 4 * static_cast<size_t>(b0 + go);
        int first = EV_TOP + 1;                                    // first sorted position that matches
        for (int e = 0; e < nS; ++e) {
            if (s_col[e] != gp) continue;
            int sub, obj;
            ev_rel_inds(f, s_row[e], sub, obj);
            if (pred_classes[sub] != cls_s || pred_classes[obj] != cls_o) continue;
            double ps[4], po[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                ps[k] = static_cast<double>(pred_boxes[static_cast<size_t>(sub) * box_ld + k]);
                po[k] = static_cast<double>(pred_boxes[static_cast<size_t>(obj) * box_ld + k]);
            }
            if (ev_iou(gbs, ps) >= iou_thr && ev_iou(gbo, po) >= iou_thr) { first = e; break; }
        }
        uint8_t* h = hits + static_cast<size_t>(g0 + g) * 4;
        h[0] = first < 10;
        h[1] = first < 20;
        h[2] = first < 50;
        h[3] = first < 100;
    }
}

// Eval-time temporal-consistency score (tools/utils/temporal_consistency.py:45-66): for every interval [s, e) of pairs
//   KLDivLoss(batchmean)(input = log_softmax(one_hot(gt)), target = softmax(dist))
//     = 1/(e-s) * sum_rows sum_c q_c (log q_c - p_c),   q = softmax(dist row),
//       p_c = [c == gt] - log(e^1 + C - 1)                                   (log_softmax of a one-hot row).
// One warp per interval, lane = class (C <= 32), warp-shuffle reductions.
__global__ void __launch_bounds__(128)
interval_kl_kernel(const float* __restrict__ dist, int C, const int32_t* __restrict__ gt, const int32_t* __restrict__ itv,
                   int n_itv, float* __restrict__ out) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n_itv) return;
    const int s = itv[2 * w], e = itv[2 * w + 1];
    const float lse_onehot = logf(expf(1.f) + static_cast<float>(C - 1));
    float acc = 0.f;
    for (int r = s; r < e; ++r) {
        const float x = lane < C ? dist[static_cast<size_t>(r) * C + lane] : -INFINITY;
        const float mx = warp_max(x);
        const float ex = lane < C ? expf(x - mx) : 0.f;
        const float sum = warp_sum(ex);
        if (lane < C) {
            const float q = ex / sum;
            const float logq = (x - mx) - logf(sum);
            const float p = (lane == gt[r] ? 1.f : 0.f) - lse_onehot;
            acc += q > 0.f ? q * (logq - p) : 0.f;             // xlogy convention of F.kl_div for q == 0
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) out[w] = acc / static_cast<float>(e - s);
}

// Class-memory accumulation (tools/utils/Memory.py:53-117: rel_memory[rel] += batch_unc^T . rel_features per video, from
// .npy files written every step): A[cls, :] += w_e * F[row_e, :] over the (row, class, weight) entries of a step, with
// the features still resident on the device.  grid = (column chunks of 128, entry slices); a CTA keeps its [C][128]
// partial sums in shared memory (thread = column: no conflicts) and adds them to A once at the end.
constexpr int CM_MAX_CLASSES = 40;
__global__ void __launch_bounds__(128)
class_memory_accumulate_kernel(const float* __restrict__ feat, int ldf, int D, const int32_t* __restrict__ ent_row,
                               const int32_t* __restrict__ ent_cls, const float* __restrict__ ent_w, int n_ent, int C,
                               float* __restrict__ A) {
    __shared__ float acc[CM_MAX_CLASSES][128];
    const int tx = threadIdx.x, col = blockIdx.x * 128 + tx;
    for (int k = 0; k < C; ++k) acc[k][tx] = 0.f;
    const int per = (n_ent + gridDim.y - 1) / gridDim.y;
    const int e0 = blockIdx.y * per, e1 = min(n_ent, e0 + per);
    if (col < D)
        for (int e = e0; e < e1; ++e)
            acc[ent_cls[e]][tx] = fmaf(ent_w[e], feat[static_cast<size_t>(ent_row[e]) * ldf + col], acc[ent_cls[e]][tx]);
    if (col < D && e1 > e0)
        for (int k = 0; k < C; ++k)
            if (acc[k][tx] != 0.f) atomicAdd(A + static_cast<size_t>(k) * D + col, acc[k][tx]);
}

}  // namespace vsgg

extern "C" int b200vsgg_class_memory_accumulate(const float* feat, int32_t ldf, int32_t D, const int32_t* ent_row,
                                                const int32_t* ent_cls, const float* ent_w, int32_t n_ent, int32_t n_classes,
                                                float* A, void* stream) {
    using namespace vsgg;
    if (!feat || !ent_row || !ent_cls || !ent_w || !A || D <= 0 || n_classes < 1 || n_classes > CM_MAX_CLASSES)
        return set_error(B200VSGG_ERR_BAD_ARG, "class_memory_accumulate: bad arg (1 <= classes <= 40)");
    if (n_ent <= 0) return 0;
    dim3 grid((D + 127) / 128, (unsigned)std::min(64, (n_ent + 63) / 64));
    class_memory_accumulate_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(feat, ldf, D, ent_row, ent_cls, ent_w, n_ent,
                                                                         n_classes, A);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_interval_kl(const float* dist, int32_t n_classes, const int32_t* gt, const int32_t* intervals,
                                    int32_t n_intervals, float* out, void* stream) {
    using namespace vsgg;
    if (!dist || !gt || !intervals || !out || n_classes < 1 || n_classes > 32)
        return set_error(B200VSGG_ERR_BAD_ARG, "interval_kl: bad arg (1 <= classes <= 32)");
    if (n_intervals <= 0) return 0;
    interval_kl_kernel<<<(n_intervals + 3) / 4, 128, 0, (cudaStream_t)stream>>>(dist, n_classes, gt, intervals, n_intervals, out);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200vsgg_eval_recall(const int64_t* pair_idx, const int32_t* frame_off, int32_t n_frames, const float* att,
                                    int32_t na, const float* spa, int32_t ns, const float* con, int32_t nc,
                                    const float* pred_boxes, int32_t box_ld, const int64_t* pred_classes,
                                    const float* obj_scores, const double* gt_boxes, const int32_t* gt_classes,
                                    const int32_t* gt_box_off, const int32_t* gt_rels, const int32_t* gt_rel_off,
                                    int32_t mode, double semi_thr, double iou_thr, uint8_t* hits, int32_t* status,
                                    void* stream) {
    using namespace vsgg;
    if (!pair_idx || !frame_off || !att || !spa || !con || !pred_boxes || !pred_classes || !obj_scores || !gt_boxes ||
        !gt_classes || !gt_box_off || !gt_rels || !gt_rel_off || !hits || !status || mode < 0 || mode > 2 || na < 2 ||
        na + ns + nc < 11 || na + ns + nc > 255)
        return set_error(B200VSGG_ERR_BAD_ARG, "eval_recall: bad arg");
    if (n_frames <= 0) return 0;
    eval_recall_kernel<<<n_frames, EV_THREADS, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const long long*>(pair_idx), frame_off, att, na, spa, ns, con, nc, pred_boxes, box_ld,
        reinterpret_cast<const long long*>(pred_classes), obj_scores, gt_boxes, gt_classes, gt_box_off, gt_rels, gt_rel_off,
        mode, semi_thr, iou_thr, hits, status);
    VSGG_CUDA_CHECK_LAUNCH();
    return 0;
}
