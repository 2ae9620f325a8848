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
} b200vsgg_gemm_epilogue;

int b200vsgg_gemm_bf16(const void* A, int32_t lda, int32_t a_mn, const void* B, int32_t ldb, int32_t b_mn,
                       int32_t M, int32_t N, int32_t K, const b200vsgg_gemm_epilogue* ep, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200VSGG_H_ */
