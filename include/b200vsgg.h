/*
 * b200vsgg — C-ABI of the B200-native relation-classification hot path
 * (TEMPURA / TEAT-GT, J-PARK11/Learning-Temporal-Consistency-for-Video-Scene-Graph-Generation).
 *
 * The reference has no FFI of its own: its seam is the Python nn.Module API
 *   TEMPURA.forward(entry, phase, unc)   lib/tempura.py:512
 *   TEAT_GT.forward(entry, phase)        lib/teatgt.py:98
 * This header is what sits *under* that API: every entry point replaces the reference lines
 * cited in its comment.  Conventions for every function:
 *   - extern "C", plain pointers and sizes, no torch types;
 *   - every pointer is a DEVICE pointer borrowed for the duration of the call unless the
 *     parameter name starts with h_ (host);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - returns 0 on success, otherwise a cudaError_t / negative b200vsgg error; never throws,
 *     never allocates device memory (callers pass workspaces);
 *   - bf16 tensors are raw uint16 storage (__nv_bfloat16), row-major, `ld*` = leading
 *     dimension in ELEMENTS.
 */
#ifndef B200VSGG_H_
#define B200VSGG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200VSGG_ERR_BAD_ARG (-1)
#define B200VSGG_ERR_NO_DRIVER (-2)
#define B200VSGG_ERR_TMAP (-3)

/* Library / build info. Returns a static string "b200vsgg <version> sm_100a". */
const char* b200vsgg_version(void);
/* Last error message of the calling thread's most recent failing call (static storage). */
const char* b200vsgg_last_error(void);

/* ------------------------------------------------------------------------------------------
 * GEMM core: D = epilogue(alpha * op(A) * op(B)^T)  — bf16 operands, fp32 accumulation in TMEM.
 * Replaces every nn.Linear / 1x1-conv / packed in_proj on the path:
 *   lib/tempura.py:543-549 (subj_fc, obj_fc, union_func1, vr_fc), tools/utils/transformer.py:9-12,
 *   38-42 (MultiheadAttention in/out projections, linear1/linear2), tools/utils/gmm_heads.py:42-45,
 *   lib/teatgt.py:121-122, tools/TokenGT/tokengt/modules/multihead_attention.py:135-183,
 *   feedforward.py:31-36, models/tokengt.py:108-117 — and their backward passes.
 *   a_mn = 0: A is [M,K] row-major (K contiguous);  a_mn = 1: A is stored [K,M] row-major.
 *   b_mn = 0: B is [N,K] row-major (K contiguous);  b_mn = 1: B is stored [K,N] row-major.
 *   forward  Y = X W^T        : (a_mn,b_mn) = (0,0)   A=X[M,in]   B=W[out,in]
 *   dgrad    dX = dY W        : (0,1)                 A=dY[M,out] B=W[out,in]  (K=out)
 *   wgrad    dW = dY^T X      : (1,1)                 A=dY[tok,out] B=X[tok,in] (K=tok)
 * Epilogue order: v = alpha*acc; v += bias[n]; v = act(v); v *= mask(mask_src[m,n]);
 *                 v = dropout(v); v += residual[m,n]; (v += out_f32[m,n] if accumulate); store.
 */
typedef struct b200vsgg_gemm_epilogue {
    const float* bias;        /* [N] fp32 or NULL */
    const void* residual;     /* [M,N] fp32 or bf16, or NULL */
    int32_t residual_is_bf16; /* 0: fp32, 1: bf16 */
    int32_t ldr;
    const void* mask_src;     /* bf16 [M,N] or NULL */
    int32_t ldm;
    int32_t mask_mode;        /* 1: v *= (mask_src > 0)   (ReLU'), 2: v *= gelu'(mask_src) */
    int32_t act;              /* 0 none, 1 relu, 2 gelu (erf form) */
    float* out_f32;           /* [M,N] fp32 or NULL */
    int32_t ld_f32;
    void* out_bf16;           /* [M,N] bf16 or NULL */
    int32_t ld_bf16;
    int32_t accumulate;       /* out_f32 += v instead of = v */
    float alpha;
    float dropout_p;          /* 0 = off. keep = hash(seed, m*N+n) >= p, scaled by 1/(1-p) */
    uint64_t dropout_seed;
    int32_t split_k;          /* 0: automatic split-K for weight-gradient shapes (few tiles, long K; partial sums are
                                 added atomically into out_f32), 1: never split, >1: forced number of splits */
    int32_t a_k_period;       /* 0: off. >0 (multiple of 64, K-major A only): A has only a_k_period columns and is
                                 re-read periodically along K, i.e. D = A * (B[:, 0:p] + B[:, p:2p] + ...)^T — used with
                                 split-precision weights B = [W_hi | W_lo] so one bf16 activation copy gives ~fp32 products */
} b200vsgg_gemm_epilogue;

int b200vsgg_gemm_bf16(const void* A, int32_t lda, int32_t a_mn, const void* B, int32_t ldb, int32_t b_mn,
                       int32_t M, int32_t N, int32_t K, const b200vsgg_gemm_epilogue* ep, void* stream);


/* ------------------------------------------------------------------------------------------
 * Segment indexing (bit-exact int32).  Replaces the per-frame Python loops that pad pair rows
 * into [l, b, 1936] (tools/utils/transformer.py:184-192) and lib/teatgt.py:104-115.
 * offsets[f] = first pair row with frame id >= f, f in [0, n_frames]; im_idx is sorted fp32.
 */
int b200vsgg_frame_offsets(const float* im_idx, int32_t n_pairs, int32_t n_frames, int32_t* offsets, void* stream);

/* out[t,:] = src[idx[t],:] (idx NULL = identity); optional outputs fp32 / bf16 / bf16(row +
 * add_table[add_idx[t],:]).  Temporal window build + position embedding add
 * (tools/utils/transformer.py:203-215) and the 'latter' scatter-back (:236-242) are both this
 * gather with host-planned index vectors. cols % 8 == 0. */
int b200vsgg_gather_rows(const float* src, int32_t ld_src, const int32_t* idx, const float* add_table,
                         const int32_t* add_idx, int32_t rows, int32_t cols, float* out_f32, int32_t ld_f32,
                         void* out_bf16, int32_t ld_bf16, void* out_bf16_added, int32_t ld_added, void* stream);

/* Deterministic backward of a gather in which every destination row is read at most twice:
 * out[n,:] = base[n,:] + sum_{k<2, idx2[2n+k] >= 0} src[idx2[2n+k],:]. */
int b200vsgg_gather2_sum_rows(const float* src, int32_t ld_src, const int32_t* idx2, const float* base,
                              int32_t ld_base, int32_t rows, int32_t cols, float* out_f32, int32_t ld_f32,
                              void* out_bf16, int32_t ld_bf16, void* stream);

/* bf16 forms of the two gathers above (rows move between the pair layout [N,.] and the window layout [M2,.] of the
 * temporal decoder, tools/utils/transformer.py:203-215,236-242, without a detour through fp32):
 *   gather_rows_bf16       out[t,:] = idx[t] >= 0 ? src[idx[t],:] : 0   (idx NULL = identity; negative = zero row)
 *   gather2_sum_rows_bf16  out[n,:] = bf16(sum_{k<2, idx2[2n+k] >= 0} float(src[idx2[2n+k],:]))
 * cols, ld_src, ld_out % 8 == 0, pointers 16-byte aligned. */
int b200vsgg_gather_rows_bf16(const void* src, int32_t ld_src, const int32_t* idx, int32_t rows, int32_t cols, void* out,
                              int32_t ld_out, void* stream);
int b200vsgg_gather2_sum_rows_bf16(const void* src, int32_t ld_src, const int32_t* idx2, int32_t rows, int32_t cols,
                                   void* out, int32_t ld_out, void* stream);

/* Pair-token gather/concat (lib/tempura.py:537-563): tok[n] = so[pair_idx[n,0],0:512] |
 * so[pair_idx[n,1],512:1024] | (vr_fc output already in tok_f32[:,1024:1536]) |
 * embed1[labels[pair_idx[n,0]]] | embed2[labels[pair_idx[n,1]]];  writes fp32 and bf16 [N,1936].
 * `so` is the fused subj_fc|obj_fc GEMM output over ALL boxes, [O,1024] fp32. */
int b200vsgg_pair_concat_fwd(const float* so, const int64_t* pair_idx, const int64_t* labels, const float* embed1,
                             const float* embed2, int32_t n_pairs, float* tok_f32, void* tok_bf16, void* stream);
int b200vsgg_pair_concat_bwd(const float* dtok, const int64_t* pair_idx, const int64_t* labels, int32_t n_pairs,
                             float* dso /* [O,1024], pre-zeroed */, float* dembed1 /* nullable */,
                             float* dembed2 /* nullable */, void* stream);

/* LayerNorm over the last dim (nn.LayerNorm, eps inside the sqrt), tools/utils/transformer.py:14-15,
 * 45 and tokengt_graph_encoder_layer.py:170-191.  cols % 8 == 0, cols <= 2560. */
int b200vsgg_layernorm_fwd(const float* x, int32_t ld_x, const float* gamma, const float* beta, int32_t rows,
                           int32_t cols, float eps, float* y_f32, int32_t ld_y, void* y_bf16, int32_t ld_b,
                           const float* add_table, const int32_t* add_idx, void* y_bf16_added, int32_t ld_added,
                           float* mean, float* rstd, void* stream);
/* dgamma / dbeta are ACCUMULATED (+=) and may be NULL; dx_bf16 (nullable) = bf16(dropout(dx)). */
int b200vsgg_layernorm_bwd(const float* dy, int32_t ld_dy, const float* x, int32_t ld_x, const float* gamma,
                           const float* mean, const float* rstd, int32_t rows, int32_t cols, float* dx_f32,
                           int32_t ld_dx, void* dx_bf16, int32_t ld_b, float drop_p, uint64_t drop_seed,
                           float* dgamma, float* dbeta, void* stream);

/* Same, with dx_f32 = base + (gradient through the norm): the backward of a pre-LN residual branch
 * (tokengt_graph_encoder_layer.py:170-191); base nullable. */
int b200vsgg_layernorm_bwd_add(const float* dy, int32_t ld_dy, const float* x, int32_t ld_x, const float* gamma,
                               const float* mean, const float* rstd, int32_t rows, int32_t cols, float* dx_f32,
                               int32_t ld_dx, void* dx_bf16, int32_t ld_b, float drop_p, uint64_t drop_seed,
                               float* dgamma, float* dbeta, void* stream, const float* base, int32_t ld_base);

/* out = bf16(dropout(x)) with the GEMM epilogue's mask function (index = row*cols + col). */
int b200vsgg_cast_dropout_bf16(const float* x, int32_t ld_x, int32_t rows, int32_t cols, void* out, int32_t ld_o,
                               float drop_p, uint64_t seed, void* stream);

/* Split-precision operand copy for the predicate-head GEMM (tools/utils/gmm_heads.py:37-76, one packed
 * [N,1936]x[1936,330] product): out[r] = [hi | lo | hi] (3*cols bf16), hi = bf16(x), lo = bf16(x - hi); against
 * [W_hi | W_hi | W_lo] the tensor-core product equals the fp32 one to ~2^-17. */
int b200vsgg_split3_bf16(const float* x, int32_t ld_x, int32_t rows, int32_t cols, void* out, int32_t ld_o,
                         void* stream);

/* out[g, c] += sum over rows r with group_idx[r]==g of x[r,c]  (bias / position-embedding grads). */
int b200vsgg_colsum(const void* x, int32_t x_is_bf16, int32_t ld_x, int32_t rows, int32_t cols,
                    const int32_t* group_idx /* nullable */, int32_t n_groups /* 1 or 2 */, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Variable-length multi-head attention over short segments (one frame / one 2-frame window):
 * nn.MultiheadAttention in tools/utils/transformer.py:23 and :50, evaluated only on real rows.
 * q,k,v,ctx,dq,dk,dv: bf16 [rows, n_heads*head_dim] views with their own leading dimensions;
 * seg_off: int32 [n_seg+1] row offsets; max_len >= longest segment; scale = head_dim^-0.5.
 * drop_p: attention-probability dropout (train), regenerated from `seed` in backward. */
int b200vsgg_attn_small_fwd(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv,
                            const int32_t* seg_off, int32_t n_seg, int32_t max_len, int32_t n_heads,
                            int32_t head_dim, float scale, void* ctx, int32_t ldc, float drop_p, uint64_t seed,
                            void* stream);
int b200vsgg_attn_small_bwd(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv,
                            const void* dctx, int32_t ldc, const int32_t* seg_off, int32_t n_seg, int32_t max_len,
                            int32_t n_heads, int32_t head_dim, float scale, void* dq, int32_t lddq, void* dk,
                            int32_t lddk, void* dv, int32_t lddv, float drop_p, uint64_t seed, void* stream);

/* Same operation for sequences of any length (< 4096) and head_dim <= 320 (even): the class-sequence encoder of the
 * SGCls object branch (lib/tempura.py:88-92,201; head_dim 297 zero-padded to 304 by the caller, sequences = object
 * tracks as long as the video).  Flash-style: lse fp32 [rows, n_heads] is written by the forward (nullable in
 * inference) and read by the backward; delta fp32 [rows, n_heads] is backward workspace (rowsum(dO*O)). */
int b200vsgg_attn_rows_fwd(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv,
                           const int32_t* seg_off, int32_t n_seg, int32_t n_heads, int32_t head_dim, float scale,
                           void* ctx, int32_t ldc, float* lse, float drop_p, uint64_t seed, void* stream);
int b200vsgg_attn_rows_bwd(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv,
                           const void* ctx, int32_t ldc, const void* dctx, int32_t lddc, const float* lse, float* delta,
                           const int32_t* seg_off, int32_t n_seg, int32_t n_heads, int32_t head_dim, float scale,
                           void* dq, int32_t lddq, void* dk, int32_t lddk, void* dv, int32_t lddv, float drop_p,
                           uint64_t seed, void* stream);

/* ------------------------------------------------------------------------------------------
 * GMM predicate heads (tools/utils/gmm_heads.py:37-76, uncertainty :25-35).  z = fp32 output of the
 * packed head GEMM; head h occupies columns [col_base, col_base + K*(2C+1)) laid out as
 * mu[K][C] | var[K][C] | pi[K].  mode 0: test (mu), 1: train (mu + sqrt(sigmoid(var))*eps),
 * 2: uncertainty (out = aleatoric, out2 = epistemic).  eps NULL in train mode = counter-based
 * N(0,1) from `seed` (the reference draws it on the CPU, gmm_heads.py:57). */
typedef struct b200vsgg_gmm_head {
    int32_t col_base;
    int32_t num_classes;
    int32_t softmax;   /* 1: softmax over classes (attention / object head), 0: sigmoid, 2: like 1 but in mode 0 the first
                          (background) class is dropped before the softmax and `out` is [N, C-1] (object head, test phase) */
    const float* eps;  /* [K, N, C] or NULL */
    float* out;        /* [N, C] */
    float* out2;       /* [N, C], mode 2 only */
    const float* dout; /* [N, C], backward only */
} b200vsgg_gmm_head;
int b200vsgg_gmm_head_fwd(const float* z, int32_t ldz, int32_t n_rows, int32_t K, const b200vsgg_gmm_head* heads,
                          int32_t n_heads, int32_t mode, uint64_t seed, void* stream);
int b200vsgg_gmm_head_bwd(const float* z, int32_t ldz, int32_t n_rows, int32_t K, const b200vsgg_gmm_head* heads,
                          int32_t n_heads, int32_t mode, uint64_t seed, void* dz_bf16, int32_t lddz,
                          int32_t total_cols, void* stream);

/* ------------------------------------------------------------------------------------------
 * Layout conversion of the detector hand-off (lib/tempura.py:548): fp32 NCHW [n,channels,spatial]
 * -> bf16 (or fp32) NHWC rows [n*spatial, channels], and back (fp32) for gradients of the mask
 * branch.  channels % 64 == 0. */
int b200vsgg_nchw_to_nhwc_bf16(const float* in, int32_t n, int32_t channels, int32_t spatial, void* out, void* stream);
int b200vsgg_nchw_to_nhwc_f32(const float* in, int32_t n, int32_t channels, int32_t spatial, float* out, void* stream);
int b200vsgg_nhwc_to_nchw_f32(const float* in, int32_t n, int32_t channels, int32_t spatial, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Spatial-mask branch of the pair token (lib/tempura.py:466-474): Conv7x7/s2(2->128)+ReLU+BN2d ->
 * MaxPool3/s2/p1 -> Conv3x3(128->256)+ReLU+BN2d, channels-last: both convolutions are
 * b200vsgg_gemm_bf16 calls over im2col rows, BatchNorm statistics are taken per video (the
 * reference's batch is one video).  All activations bf16 NHWC rows, channels % 8 == 0.
 */
/* masks fp32 [n,2,27,27] -> bf16 rows [n*196, ld]; column c*49+kh*7+kw for the 98 taps, zero up to ld (ld>=104). */
int b200vsgg_mask_im2col(const float* masks, int32_t n, void* out, int32_t ld, void* stream);
/* Same for the producer-side hand-off ((f).4, tools/utils/object_detector.py:372-380 emitting what the path consumes):
 * masks bf16 [n,2,27,27].  The im2col rows are bf16 either way, so nothing is lost against the fp32 hand-off. */
int b200vsgg_mask_im2col_bf16(const void* masks, int32_t n, void* out, int32_t ld, void* stream);
/* Segmented column statistics: chunk table int32 [n_chunks,3] = (row_begin,row_end,group);
 * sum1[g,c] += sum_r a[r,c]; sum2[g,c] += sum_r a[r,c]*b[r,c] (b, sum2 nullable; b is bf16, or b == a for
 * sums of squares of either type). */
int b200vsgg_seg_colstats(const void* a, int32_t a_is_bf16, int32_t lda, const void* b, int32_t ldb, int32_t cols,
                          const int32_t* chunks, int32_t n_chunks, float* sum1, float* sum2, void* stream);
/* out[r,c] = bf16(k1[g,c]*a[r,c] + k2[g,c]*b[r,c] + k3[g,c]), zero where relu_mask == 1 && b[r,c] <= 0; relu_mask == 2:
 * out = max(out, 0) (Linear -> BatchNorm1d -> ReLU of lib/tempura.py:103-105);
 * g = group_of_unit[r / rows_per_unit]; a (and k1) may be NULL.  BN apply and BN+ReLU backward. */
int b200vsgg_seg_affine(const void* a, const void* b, const float* k1, const float* k2, const float* k3,
                        const int32_t* group_of_unit, int64_t rows, int32_t rows_per_unit, int32_t cols,
                        int32_t relu_mask, void* out, void* stream);
/* z = maxpool3x3/s2/p1(scale[g,c]*y + shift[g,c]); y bf16 or fp32 [n,hw_in,hw_in,C] -> z bf16 [n,ho,ho,C], ho=(hw_in+1)/2;
 * argmax u8 [n,ho,ho,C] = kh*3+kw of the first maximum (torch's tie rule). */
int b200vsgg_bn_pool_fwd(const void* y, int32_t y_is_f32, const float* scale, const float* shift, const int32_t* group_of_unit,
                         int32_t n, int32_t hw_in, int32_t channels, void* z, uint8_t* argmax, void* stream);
int b200vsgg_pool_bwd(const void* dz, const uint8_t* argmax, int32_t n, int32_t hw_in, int32_t channels, void* dy,
                      void* stream);
/* z bf16 [n,hw,hw,C] -> rows [n*hw*hw, 9*C], column (kh*3+kw)*C+c, padding 1; and its transpose (gather-sum). */
int b200vsgg_im2col3x3(const void* z, int32_t n, int32_t hw, int32_t channels, void* out, void* stream);
int b200vsgg_col2im3x3(const void* dcol, int32_t n, int32_t hw, int32_t channels, void* dz, void* stream);

/* Relationship contrastive loss of the trainers (TEMPURA_train.py:103,209-212; TEATGT_train.py:81,176-179):
 * pytorch_metric_learning ContrastiveLoss(pos_margin, neg_margin) on L2-normalised embeddings x fp32 [N,C] (C <= 32: the
 * spatial / contacting distributions) with integer labels, all pairs inside each segment [seg_off[v], seg_off[v+1]) (the
 * reference's call = one video), AvgNonZeroReducer per group.  loss fp32 [n_seg]; dx (nullable) fp32 [N,C] = gradient of
 * loss[v] w.r.t. the segment's rows times grad_scale[0] (nullable = 1).  max_rows = longest segment (shared memory). */
int b200vsgg_contrastive_loss(const float* x, const int32_t* label, const int32_t* seg_off, int32_t n_seg, int32_t max_rows,
                              int32_t C, float pos_margin, float neg_margin, float* loss, float* dx, const float* grad_scale,
                              void* stream);

/* ------------------------------------------------------------------------------------------
 * TEAT-GT / TokenGT path (lib/teatgt.py, tools/TokenGT/tokengt).
 */
/* Blackwell-native forward of the same attention (csrc/attn_tc.cu): 128-query x 128-key tiles, S = Q K^T and
 * O += P V on tcgen05.mma with S / P / O in tensor memory, Q / K / V tiles fetched by 3-D TMA maps
 * {head_dim, row, head} whose out-of-range columns zero-pad a 24- / 48-wide head to the 128-byte swizzled rows,
 * thread = query row softmax on tcgen05.ld with lazy rescaling.  Same arguments as b200vsgg_attn_flash_fwd except
 * that the host-planned blocks are 128 rows and `rows` (= total token rows of q/k/v) is passed for the tensor maps.
 * Writes the same lse and uses the same dropout mask function, so b200vsgg_attn_flash_bwd is its backward.
 * Replaces multihead_attention.py:135-183 (incl. the [heads,T,T] maps kept by tokengt_graph_encoder_layer.py:170-191). */
int b200vsgg_attn_tc_fwd(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv, int32_t rows,
                         const int32_t* seq_off, const int32_t* blk_seq, const int32_t* blk_row0, int32_t n_blocks,
                         int32_t n_heads, int32_t head_dim, float scale, void* ctx, int32_t ldc, float* lse, float drop_p,
                         uint64_t seed, void* stream);

/* Blackwell-native backward (csrc/attn_tc_bwd.cu): a dQ kernel (128-query tiles) and a dK/dV kernel (128-key tiles),
 * both with the S-type products and the accumulating products on tcgen05.mma (A operand of the latter read from tensor
 * memory), probabilities recomputed from `lse`, dropout regenerated from `seed`.  ctx / dctx: the forward output and its
 * gradient (bf16 [rows, n_heads*head_dim]); delta: fp32 workspace [rows, n_heads]; the block table has 128-row blocks. */
int b200vsgg_attn_tc_bwd(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv,
                         const void* ctx, int32_t ldc, const void* dctx, int32_t lddc, const float* lse, float* delta,
                         int32_t rows, const int32_t* seq_off, const int32_t* blk_seq, const int32_t* blk_row0,
                         int32_t n_blocks, int32_t n_heads, int32_t head_dim, float scale, void* dq, int32_t lddq, void* dk,
                         int32_t lddk, void* dv, int32_t lddv, float drop_p, uint64_t seed, void* stream);

/* Variable-length flash attention over clip sequences: multihead_attention.py:135-183 without the [heads,T,T]
 * maps.  q,k,v,ctx: bf16 [rows, n_heads*head_dim] views (ld in elements, 16-byte aligned, head_dim % 8 == 0,
 * head_dim <= 64); seq_off int32 [n_seq+1]; the query/key blocking is host-planned: block b covers rows
 * [blk_row0[b], min(blk_row0[b]+64, seq_off[blk_seq[b]+1])) of sequence blk_seq[b].  scale multiplies q.k
 * (the reference scales q by head_dim^-0.5 after the bias: same product).  lse fp32 [rows, n_heads]
 * (nullable in inference).  Dropout acts on the probabilities and is regenerated from `seed` in backward. */
int b200vsgg_attn_flash_fwd(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv,
                            const int32_t* seq_off, const int32_t* blk_seq, const int32_t* blk_row0, int32_t n_blocks,
                            int32_t n_heads, int32_t head_dim, float scale, void* ctx, int32_t ldc, float* lse,
                            float drop_p, uint64_t seed, void* stream, int32_t n_seq, int32_t max_len);
/* delta fp32 [rows, n_heads] is workspace (rowsum(dO*O), written here). */
int b200vsgg_attn_flash_bwd(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv,
                            const void* ctx, int32_t ldc, const void* dctx, int32_t lddc, const float* lse, float* delta,
                            const int32_t* seq_off, const int32_t* blk_seq, const int32_t* blk_row0, int32_t n_blocks,
                            int32_t n_rows, int32_t n_heads, int32_t head_dim, float scale, void* dq, int32_t lddq,
                            void* dk, int32_t lddk, void* dv, int32_t lddv, float drop_p, uint64_t seed, void* stream,
                            int32_t n_seq, int32_t max_len);
/* (n_seq, max_len = longest sequence: when its head slices fit in shared memory — T <= ~640 at head_dim 24 — the
 * resident kernels run: one CTA per (sequence, head), operands loaded once; pass n_seq = 0 to force the tiled path.) */
/* Node tokens (lib/teatgt.py:118-141): tok[i] = so[feat_row[i], half(is_person)] | embed[labels[feat_row[i]]];
 * so = features @ [subj_fc; obj_fc]^T + bias over all boxes, [O, 2*h1] fp32. */
int b200vsgg_node_tokens_fwd(const float* so, int32_t ld_so, const int32_t* feat_row, const int32_t* is_person,
                             const int64_t* labels, const float* embed, int32_t n, int32_t h1, int32_t e, float* out_f32,
                             void* out_bf16, void* stream);
int b200vsgg_node_tokens_bwd(const float* dtok, const int32_t* feat_row, const int32_t* is_person, const int64_t* labels,
                             int32_t n, int32_t h1, int32_t e, float* dso /* pre-zeroed */, int32_t ld_dso,
                             float* dembed /* nullable, pre-zeroed */, void* stream);
/* Edge predicates of the pseudo-graph (lib/teatgt.py:199-217): per frame f, uint8 [nmax,nmax] matrices
 * spatial[a][b] (a<b, centre distance <= thr) and temporal[p][c] (cosine(prev-frame node p, node c) >= sim, only
 * if has_prev[f]).  The edge LIST (reference order) is compacted on the host from these flags. */
int b200vsgg_teat_pair_flags(const float* tok, int32_t d, const float* boxes, const int32_t* feat_row,
                             const int32_t* node_off, const int32_t* has_prev, int32_t n_frames, float thr, float sim,
                             int32_t nmax, uint8_t* spatial, uint8_t* temporal, void* stream);
/* Token assembly (tokenizer.py:217-295): desc int32 [T,4] = (kind, a, b, c); kind 0 [graph], 1 [null],
 * 2 node a (b = frame - first frame of clip), 3 edge (a,b) of type c.  See teat_kernels.cu. */
int b200vsgg_teat_assemble_fwd(const int32_t* desc, int32_t n_tokens, int32_t d, const float* na, const float* pu,
                               const float* pv, const float* temp, const float* eemb, const float* order,
                               const float* graph_tok, const float* null_tok, float* x, void* stream);
int b200vsgg_teat_assemble_bwd(const int32_t* desc, int32_t n_tokens, int32_t d, const float* dx, float* dna, float* dpu,
                               float* dpv, float* dtemp, float* deemb, float* dorder, float* dgraph, float* dnull,
                               void* stream);
/* out = bf16(dropout(act(x))) for bf16 x (feedforward.py:31-36: GELU + activation dropout). */
int b200vsgg_act_dropout_bf16(const void* x, int32_t ld_x, int64_t rows, int32_t cols, int32_t act, float p, uint64_t seed,
                              void* out, int32_t ld_o, void* stream);

/* Pairwise graph temporal-consistency reduction (lib/teatgt.py:325-334): out[p] =
 * KLDiv_batchmean(log_softmax(g[pair_u[p]]), softmax(g[pair_v[p]])) / (pair_v[p] - pair_u[p]); g fp32 [frames, d]
 * (frame indices are absolute, so v - u is the frame distance inside the clip). */
int b200vsgg_consistency_kl(const float* g, int32_t d, const int32_t* pair_u, const int32_t* pair_v, int32_t n_pairs,
                            float* out, void* stream);

/* Backward kernels of the DIFFERENTIABLE consistency mode (the reference detaches the two loss vectors,
 * lib/teatgt.py:350-351 — `differentiable_consistency=True` is the "fixed" behaviour behind a flag).  All recompute their
 * forward quantities from the saved inputs.
 *   consistency_kl_bwd : dg [frames, d] += gradient of out[p] = KL / (v - u) scaled by gout[p] (0 = pair dropped)
 *   attn_pool_bwd      : dx [rows, d] and dgate [rows] (gate-logit gradients: dw = weighted_colsum(x, dgate))
 *   weighted_colsum    : out[c] += sum_r wgt[r] * x[r, c]   (x fp32 or bf16)
 *   gated_residual_bwd : d_o, d_res [rows, dim] and da [rows] (dw1 = weighted_colsum(o, da), dw2 = ...(res, da), dw3 = dw1 - dw2)
 *   graph_attn_core_bwd: dqkv fp32 [rows, 1536] from dout fp32 [rows, 512]; dwe / dbe [512] += (edges_to_kv gradients) */
int b200vsgg_consistency_kl_bwd(const float* g, int32_t d, const int32_t* pair_u, const int32_t* pair_v, const float* gout,
                                int32_t n_pairs, float* dg, void* stream);
int b200vsgg_attn_pool_bwd(const float* x, int32_t d, const int32_t* node_off, int32_t n_frames, int32_t max_nodes,
                           const float* w, const float* b, const float* dout, float* dx, float* dgate, void* stream);
int b200vsgg_weighted_colsum(const void* x, int32_t x_is_bf16, int32_t ld, int32_t rows, int32_t cols, const float* wgt,
                             float* out, void* stream);
int b200vsgg_gated_residual_bwd(const float* o, const float* res, const float* w, const float* dx, int32_t rows, int32_t dim,
                                float* d_o, float* d_res, float* da, void* stream);
int b200vsgg_graph_attn_core_bwd(const float* qkv, int32_t ld, const int32_t* node_off, const uint8_t* upper, int32_t nmax,
                                 const float* we, const float* be, const float* dout, int32_t ldd, int32_t n_frames,
                                 float* dqkv, int32_t ldg, float* dwe, float* dbe, void* stream);

/* SIMT fp32 helpers for the 10-wide structure branch of the regulariser in differentiable mode (its linears have K = 10 or
 * N = 10; the default detached mode runs the whole branch in one launch, b200vsgg_graph_small_fwd):
 *   simt_linear : y[r,o] = act(sum_i x[r,i] * w[o*so + i*si] + b[o])  (so/si select W or W^T; act 0 / 2 = GELU;
 *                 z_out optionally receives the pre-activation)
 *   simt_wgrad  : dw[o,i] += sum_r dy[r,o] x[r,i]
 *   gelu_bwd    : dz = dy * gelu'(z)
 *   ln_small_*  : LayerNorm over d <= 32 columns, forward (saves mean / rstd) and backward (dx = base + ..., dgamma/dbeta +=) */
int b200vsgg_simt_linear(const float* x, int32_t ldx, const float* w, int32_t so, int32_t si, const float* b, int64_t rows,
                         int32_t n_out, int32_t n_in, int32_t act, float* y, int32_t ldy, float* z_out, void* stream);
int b200vsgg_simt_wgrad(const float* dy, int32_t ldd, const float* x, int32_t ldx, int32_t rows, int32_t n_out, int32_t n_in,
                        float* dw, void* stream);
int b200vsgg_gelu_bwd(const float* dy, const float* z, int64_t n, float* dz, void* stream);
int b200vsgg_ln_small_fwd(const float* x, const float* g, const float* b, int32_t rows, int32_t d, float* y, float* mean,
                          float* rstd, void* stream);
int b200vsgg_ln_small_bwd(const float* dy, const float* x, const float* g, const float* mean, const float* rstd,
                          const float* base, int32_t rows, int32_t d, float* dx, float* dgamma, float* dbeta, void* stream);

/* GlobalAttentionPooling of the regulariser (lib/teatgt.py:319-320, dgl.nn.GlobalAttentionPooling with gate_nn =
 * Linear(d, 1)): per frame a = softmax_i(w . x_i + b), out[f] = sum_i a_i x_i.  x fp32 [rows, d] compact node rows,
 * node_off int32 [frames+1], max_nodes <= 64, out fp32 [frames, d].  One CTA per frame. */
int b200vsgg_attn_pool(const float* x, int32_t d, const int32_t* node_off, int32_t n_frames, int32_t max_nodes,
                       const float* w, const float* b, float* out, void* stream);

/* Per-frame graph attention core of the regulariser's GraphTransformer (8 heads x 64, rotary q/k, per-edge
 * key/value offsets from the adjacency; lib/teatgt.py:316-317 via graph_transformer_pytorch): qkv fp32
 * [rows, >= 1536] = q | k | v of the compact node rows, node_off int32 [frames+1], upper = uint8 predicate
 * matrices of b200vsgg_teat_pair_flags; out bf16 [rows, 512]. */
int b200vsgg_graph_attn_core(const float* qkv, int32_t ld, const int32_t* node_off, const uint8_t* upper, int32_t nmax,
                             const float* we, const float* be, int32_t n_frames, void* out, int32_t ldo, void* stream);
/* Same, fp32 output (the structure branch of the differentiable mode keeps fp32 throughout). */
int b200vsgg_graph_attn_core_f32(const float* qkv, int32_t ld, const int32_t* node_off, const uint8_t* upper, int32_t nmax,
                                 const float* we, const float* be, int32_t n_frames, float* out, int32_t ldo, void* stream);
/* GatedResidual: res <- o*g + res*(1-g), g = sigmoid(W [o, res, o-res]); w fp32 [3*dim]. */
int b200vsgg_gated_residual(const float* o, float* res, const float* w, int32_t rows, int32_t dim, void* stream);

/* ------------------------------------------------------------------------------------------
 * SGCls object branch (lib/tempura.py:185-255 / tools/utils/object_classifier.py:177-233).
 * Object tokens: row r <- box s = src[r] (src NULL = identity):
 *   x[r] = drop_pe([features[s] | dist[s] @ embed | drop_pos(relu(wp * bn(center_size(box_s)) + bp))] + pe[pos[r]])
 * = lib/tempura.py:249-252 (feature build) fused with the pad_sequence gather (:197) and PositionalEncoding.forward
 * (:39-48).  pos NULL: no position term / no drop_pe (the non-tracking path).  BatchNorm1d(4) enters as per-video
 * (mean, rstd) [V,4] (+ gamma, beta [4]); video_of_box int32 [O].  All tensors fp32; feat_dim, e, h multiples of 4.
 * The class-sequence encoder itself (3 x nn.TransformerEncoderLayer, :88-92) runs on b200vsgg_gemm_bf16 /
 * b200vsgg_attn_small_* / b200vsgg_layernorm_* with heads padded 297 -> 304 columns in the weight copies. */
typedef struct b200vsgg_obj_tokens {
    const float* features; int32_t feat_dim;   /* [O, feat_dim] */
    const float* dist; int32_t n_cls;          /* [O, n_cls] detector posterior */
    const float* embed; int32_t e;             /* obj_embed.weight [n_cls, e] */
    const float* boxes;                        /* [O,5] = (frame, x1, y1, x2, y2) */
    const float *bn_mean, *bn_rstd;            /* [V,4] */
    const float *bn_gamma, *bn_beta;           /* [4] */
    const int32_t* video_of_box;               /* [O] */
    const float *wp, *bp; int32_t h;           /* pos_embed Linear: [h,4], [h] */
    const float* pe;                           /* [max_len, feat_dim+e+h] sinusoid table (nullable with pos) */
    const int32_t *src, *pos;                  /* [rows] (nullable) */
    int32_t rows;
    float p_pos; uint64_t seed_pos;            /* nn.Dropout(0.1) inside pos_embed (indexed by box) */
    float p_pe; uint64_t seed_pe;              /* dropout of PositionalEncoding (indexed by row) */
} b200vsgg_obj_tokens;
int b200vsgg_obj_tokens_fwd(const b200vsgg_obj_tokens* p, float* x_f32, void* x_bf16, void* stream);
/* Parameter gradients only (features / dist / boxes are frozen detector outputs); outputs pre-zeroed, accumulated:
 * dembed [n_cls,e], dwp [h,4], dbp [h], dgamma [4], dbeta [4]; dx fp32 [rows, feat_dim+e+h]. */
int b200vsgg_obj_tokens_bwd(const b200vsgg_obj_tokens* p, const float* dx, float* dembed, float* dwp, float* dbp,
                            float* dgamma, float* dbeta, void* stream);

/* Trainer loss block of the relation heads, loss AND gradient in one launch (TEMPURA_train.py:181-206,
 * TEATGT_train.py:153-175): losses[0] += sum_i w_i * CE(att[i,:] used as logits, att_label[i]);
 * losses[1] += sum_i w_i * mean_c BCE(spa[i,c], t[i,c]); losses[2] likewise for con.  The multi-hot targets come
 * either dense (fp32 [n,C]) or as the ragged label lists the dataloader yields (CSR: off int32 [n+1], idx int32).
 * row_w carries the reduction (1/n for one video).  d_* (nullable) receive d(sum of the three losses)/d(input). */
int b200vsgg_rel_loss(const float* att, const float* spa, const float* con, int32_t n, int32_t ca, int32_t cs, int32_t cc,
                      const int64_t* att_label, const float* spa_dense, const float* con_dense, const int32_t* spa_off,
                      const int32_t* spa_idx, const int32_t* con_off, const int32_t* con_idx, const float* row_w,
                      float* losses /* [3], pre-zeroed */, float* d_att, float* d_spa, float* d_con, void* stream);

/* Structure branch of the regulariser, whole network in one launch (lib/teatgt.py:291-311,316,319 via
 * graph_transformer_pytorch.GraphTransformer(dim, depth, heads x 64, edge_dim 1, feed-forwards, gated residuals,
 * rotary) + dgl GlobalAttentionPooling): nodes fp32 [frames, nmax, dim] (first `dim` Laplacian eigenvector columns),
 * upper uint8 [frames, nmax, nmax] (b200vsgg_teat_pair_flags; adjacency = U + U^T), counts int32 [frames] ->
 * out fp32 [frames, dim].  params = `depth` packed layers of b200vsgg_graph_small_params_per_layer(dim, heads)
 * floats each, in the order ln1.w ln1.b to_q.w to_q.b to_kv.w to_kv.b edges_to_kv.w edges_to_kv.b to_out.w to_out.b
 * gate1.w ln2.w ln2.b ff1.w ff1.b ff2.w ff2.b gate2.w.  nmax <= 16, dim <= 16. */
int b200vsgg_graph_small_params_per_layer(int32_t dim, int32_t heads);
int b200vsgg_graph_small_fwd(const float* nodes, const uint8_t* upper, const int32_t* counts, int32_t n_frames,
                             int32_t nmax, int32_t dim, int32_t heads, int32_t depth, const float* params,
                             const float* pool_w, const float* pool_b, float* out, void* stream);

/* Upload `bytes` (multiple of 16) from PINNED host memory to the device with a kernel on `stream` instead of the
 * copy engine (see frontend.cu): used for the per-batch index vectors so they never queue behind a bulk
 * input prefetch. h_pinned_src must stay untouched until the kernel has run. */
int b200vsgg_upload(const void* h_pinned_src, void* dst, int64_t bytes, void* stream);

/* Class-memory accumulation of the trainers' uncertainty / memory bookkeeping (SURVEY.md 8(f).4;
 * tools/utils/Memory.py:53-117 `rel_memory[rel] += batch_unc.T @ rel_features`, fed by tools/utils/Uncertainty.py:105-178
 * which writes every video's features to .npy files): A[ent_cls[e], :] += ent_w[e] * feat[ent_row[e], :] for the
 * (row, class, weight) entries of one step; feat fp32 [N,D] with row pitch ldf, A fp32 [n_classes, D] (n_classes <= 40). */
int b200vsgg_class_memory_accumulate(const float* feat, int32_t ldf, int32_t D, const int32_t* ent_row, const int32_t* ent_cls,
                                     const float* ent_w, int32_t n_ent, int32_t n_classes, float* A, void* stream);

/* Eval-time temporal-consistency score (tools/utils/temporal_consistency.py:45-66): for each interval
 * [s, e) of pair rows, KLDivLoss(batchmean)(log_softmax(one_hot(gt[s:e])), softmax(dist[s:e])).  dist fp32 [N,n_classes]
 * (n_classes <= 32), gt int32 [N], intervals int32 [I,2], out fp32 [I].  One warp per interval. */
int b200vsgg_interval_kl(const float* dist, int32_t n_classes, const int32_t* gt, const int32_t* intervals,
                         int32_t n_intervals, float* out, void* stream);

/* Recall@K matching of one video (tools/utils/evaluation_recall.py:119-276: evaluate_from_dict,
 * evaluate_recall, _triplet, _compute_pred_matches), one CTA per frame.
 *   pair_idx int64 [N,2] box rows, frame_off int32 [F+1] pair offsets of the frames, att/spa/con fp32 distributions
 *   [N,na]/[N,ns]/[N,nc], pred_boxes fp32 rows of 4 coordinates with row pitch box_ld floats, pred_classes int64 [O],
 *   obj_scores fp32 [O]; ground truth concatenated over frames: gt_boxes fp64 [B,4], gt_classes int32 [B],
 *   gt_box_off int32 [F+1], gt_rels int32 [G,3] = (subject, object: frame-local box index; predicate id),
 *   gt_rel_off int32 [F+1].
 * mode 0 = "with" constraint (argmax predicate per relation row), 1 = "no" constraint (top-100 of score x object scores),
 * 2 = "semi" (attention rows: argmax; other rows: every predicate above semi_thr).  Scores are float64 built with the
 * reference's dtypes; ties: position-descending (a stable ascending argsort reversed).
 * hits uint8 [G,4]: ground-truth relation matched within the first 10 / 20 / 50 / 100 candidates (class triplet equal,
 * both IoUs >= iou_thr, +1 pixel convention).  *status is set to 1 if a frame exceeds the on-chip tables (> 42 pairs):
 * its hits are 0 and the caller evaluates that frame on the host. */
int b200vsgg_eval_recall(const int64_t* pair_idx, const int32_t* frame_off, int32_t n_frames, const float* att, int32_t na,
                         const float* spa, int32_t ns, const float* con, int32_t nc, const float* pred_boxes, int32_t box_ld,
                         const int64_t* pred_classes, const float* obj_scores, const double* gt_boxes,
                         const int32_t* gt_classes, const int32_t* gt_box_off, const int32_t* gt_rels,
                         const int32_t* gt_rel_off, int32_t mode, double semi_thr, double iou_thr, uint8_t* hits,
                         int32_t* status, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused multi-tensor optimiser step: tools/utils/AdamW.py:53-113 (weight decay before the moment update) with
 * torch.nn.utils.clip_grad_norm_ (TEMPURA_train.py:224) folded in.  `tensors` is a DEVICE array; the work is
 * split into chunks of chunk_elems elements: chunk c covers tensor chunk_tensor[c] from element chunk_off[c].
 * sq_norm: device scalar, pre-zeroed by the caller, filled by b200vsgg_grad_sqnorm and consumed by the step
 * (NULL = no clipping). */
typedef struct b200vsgg_opt_tensor {
    float* p;                 /* parameter, updated in place */
    const float* g;           /* gradient */
    float* m;                 /* exp_avg */
    float* v;                 /* exp_avg_sq */
    int64_t n;                /* elements */
    float bias_correction1;   /* 1 - beta1^step of this tensor */
    float bias_correction2;   /* 1 - beta2^step */
} b200vsgg_opt_tensor;
int b200vsgg_grad_sqnorm(const b200vsgg_opt_tensor* tensors, const int32_t* chunk_tensor, const int64_t* chunk_off,
                         int32_t n_chunks, int32_t chunk_elems, float* sq_norm, void* stream);
int b200vsgg_adamw_clip_step(const b200vsgg_opt_tensor* tensors, const int32_t* chunk_tensor, const int64_t* chunk_off,
                             int32_t n_chunks, int32_t chunk_elems, const float* sq_norm, float max_norm, float lr,
                             float beta1, float beta2, float eps, float weight_decay, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200VSGG_H_ */
