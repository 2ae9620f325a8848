"""Python wrappers (torch tensors in, raw device pointers out) around the C-ABI kernels.
These are plumbing only: every function launches hand-written sm_100a kernels from
libb200vsgg.so on torch's current CUDA stream and raises if the library is missing."""
import ctypes as C

import torch

from . import _lib
from ._lib import GemmEpilogue, check

ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
MASK_RELU, MASK_GELU = 1, 2

# Number of kernel launches issued through this module (bench.py reports it as gpu_launches).
launch_count = 0
# When set to a list, every GEMM launch appends (M, N, K, a_mn, b_mn, start_event, end_event):
# bench.py uses it to time the dominant kernel live with CUDA events on the launching stream.
gemm_profile = None
# Same for the TokenGT attention kernels: (kind, algorithmic flops, start_event, end_event); the caller passes the flops
# (4 * sum T^2 * heads * head_dim forward, 2.5x that backward).
attn_profile = None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


_DEBUG_SYNC = bool(int(__import__("os").environ.get("B200VSGG_DEBUG_SYNC", "0")))


def _count(n=1):
    global launch_count
    launch_count += n
    if _DEBUG_SYNC:  # debugging aid: surface asynchronous faults at the op that caused them
        import inspect
        try:
            torch.cuda.synchronize()
        except Exception as ex:
            raise RuntimeError("CUDA fault surfaced after ops.%s: %s" % (inspect.stack()[1].function, ex))


def gemm(a, b, *, a_mn=False, b_mn=False, bias=None, residual=None, mask_src=None, mask_mode=0,
         act=ACT_NONE, out_f32=None, out_bf16=None, accumulate=False, alpha=1.0,
         dropout_p=0.0, seed=0, split_k=0, a_k_period=0):
    """D = epilogue(alpha * op(a) @ op(b)^T); see b200vsgg_gemm_bf16 in include/b200vsgg.h.

    a: [M,K] (a_mn=False) or [K,M] (a_mn=True); b: [N,K] (b_mn=False) or [K,N] (b_mn=True); both
    bf16 with unit inner stride.  At least one of out_f32 / out_bf16 must be given ([M,N])."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.stride(-1) == 1 and b.stride(-1) == 1
    if a_mn:
        K, M = a.shape
    else:
        M, K = a.shape
    if b_mn:
        Kb, N = b.shape
    else:
        N, Kb = b.shape
    if a_k_period:
        assert not a_mn and K == a_k_period
        K = Kb
    assert K == Kb, (a.shape, b.shape, a_mn, b_mn)
    ep = GemmEpilogue()
    ep.bias = bias.data_ptr() if bias is not None else None
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
    if residual is not None:
        assert residual.shape == (M, N) and residual.stride(1) == 1
        ep.residual = residual.data_ptr()
        ep.residual_is_bf16 = 1 if residual.dtype == torch.bfloat16 else 0
        ep.ldr = residual.stride(0)
    if mask_src is not None:
        assert mask_src.dtype == torch.bfloat16 and mask_src.shape == (M, N) and mask_src.stride(1) == 1
        ep.mask_src = mask_src.data_ptr()
        ep.ldm = mask_src.stride(0)
        ep.mask_mode = mask_mode
    ep.act = act
    if out_f32 is not None:
        assert out_f32.dtype == torch.float32 and out_f32.shape == (M, N) and out_f32.stride(1) == 1
        ep.out_f32 = out_f32.data_ptr()
        ep.ld_f32 = out_f32.stride(0)
    if out_bf16 is not None:
        assert out_bf16.dtype == torch.bfloat16 and out_bf16.shape == (M, N) and out_bf16.stride(1) == 1
        ep.out_bf16 = out_bf16.data_ptr()
        ep.ld_bf16 = out_bf16.stride(0)
    ep.accumulate = 1 if accumulate else 0
    ep.alpha = alpha
    ep.dropout_p = dropout_p
    ep.dropout_seed = seed
    ep.split_k = split_k
    ep.a_k_period = a_k_period
    prof = gemm_profile
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = _lib.lib().b200vsgg_gemm_bf16(_ptr(a), a.stride(0), int(a_mn), _ptr(b), b.stride(0), int(b_mn),
                                       M, N, K, C.byref(ep), _stream())
    check(rc, "gemm_bf16")
    if prof is not None:
        e1.record()
        prof.append((M, N, K, int(a_mn), int(b_mn), e0, e1))
    _count()
    return out_f32 if out_f32 is not None else out_bf16


# ------------------------------------------------------------------------------------------------
# row kernels
# ------------------------------------------------------------------------------------------------
def _f32(t):
    assert t is None or (t.dtype == torch.float32 and t.stride(-1) == 1), "expected contiguous-row fp32"
    return t


def _bf(t):
    assert t is None or (t.dtype == torch.bfloat16 and t.stride(-1) == 1), "expected contiguous-row bf16"
    return t


def _ld(t):
    return t.stride(0) if t is not None and t.dim() == 2 else 0


def frame_offsets(im_idx, n_frames):
    """int32 [F+1] exclusive offsets of the sorted fp32 frame ids (device tensor)."""
    assert im_idx.dtype == torch.float32 and im_idx.is_contiguous()
    out = torch.empty(n_frames + 1, dtype=torch.int32, device=im_idx.device)
    check(_lib.lib().b200vsgg_frame_offsets(_ptr(im_idx), im_idx.numel(), n_frames, _ptr(out), _stream()),
          "frame_offsets")
    _count()
    return out


def gather_rows(src, idx=None, rows=None, add_table=None, add_idx=None, out_f32=None, out_bf16=None,
                out_bf16_added=None):
    _f32(src)
    cols = src.shape[1]
    rows = rows if rows is not None else (idx.numel() if idx is not None else src.shape[0])
    for t in (idx, add_idx):
        assert t is None or (t.dtype == torch.int32 and t.is_contiguous())
    check(_lib.lib().b200vsgg_gather_rows(
        _ptr(src), src.stride(0), _ptr(idx), _ptr(_f32(add_table)), _ptr(add_idx), rows, cols,
        _ptr(_f32(out_f32)), _ld(out_f32), _ptr(_bf(out_bf16)), _ld(out_bf16), _ptr(_bf(out_bf16_added)),
        _ld(out_bf16_added), _stream()), "gather_rows")
    _count()


def gather2_sum_rows(src, idx2, base=None, out_f32=None, out_bf16=None):
    _f32(src)
    assert idx2.dtype == torch.int32 and idx2.is_contiguous() and idx2.dim() == 2 and idx2.shape[1] == 2
    rows, cols = idx2.shape[0], src.shape[1]
    check(_lib.lib().b200vsgg_gather2_sum_rows(
        _ptr(src), src.stride(0), _ptr(idx2), _ptr(_f32(base)), _ld(base), rows, cols, _ptr(_f32(out_f32)),
        _ld(out_f32), _ptr(_bf(out_bf16)), _ld(out_bf16), _stream()), "gather2_sum_rows")
    _count()


def gather_rows_bf16(src, idx, out, rows=None):
    """out[t] = src[idx[t]] for bf16 rows; a negative index yields a zero row (idx None = identity copy)."""
    _bf(src), _bf(out)
    assert idx is None or (idx.dtype == torch.int32 and idx.is_contiguous())
    assert src.dim() == 2 and out.dim() == 2 and src.stride(1) == 1 and out.stride(1) == 1 and out.shape[1] == src.shape[1]
    rows = rows if rows is not None else (idx.numel() if idx is not None else src.shape[0])
    assert out.shape[0] == rows
    check(_lib.lib().b200vsgg_gather_rows_bf16(_ptr(src), src.stride(0), _ptr(idx), rows, src.shape[1], _ptr(out),
                                               out.stride(0), _stream()), "gather_rows_bf16")
    _count()


def gather2_sum_rows_bf16(src, idx2, out):
    """out[n] = bf16(sum of the <= 2 bf16 rows src[idx2[n, k]] with idx2[n, k] >= 0), accumulated in fp32."""
    _bf(src), _bf(out)
    assert idx2.dtype == torch.int32 and idx2.is_contiguous() and idx2.dim() == 2 and idx2.shape[1] == 2
    assert src.dim() == 2 and out.dim() == 2 and src.stride(1) == 1 and out.stride(1) == 1 and out.shape[1] == src.shape[1]
    assert out.shape[0] == idx2.shape[0]
    check(_lib.lib().b200vsgg_gather2_sum_rows_bf16(_ptr(src), src.stride(0), _ptr(idx2), idx2.shape[0], src.shape[1],
                                                    _ptr(out), out.stride(0), _stream()), "gather2_sum_rows_bf16")
    _count()


def pair_concat_fwd(so, pair_idx, labels, embed1, embed2, tok_f32, tok_bf16):
    assert so.is_contiguous() and so.shape[1] == 1024 and pair_idx.dtype == torch.int64 and pair_idx.is_contiguous()
    assert labels.dtype == torch.int64 and tok_f32.is_contiguous() and tok_bf16.is_contiguous()
    assert embed1.is_contiguous() and embed2.is_contiguous() and embed1.shape[1] == 200
    check(_lib.lib().b200vsgg_pair_concat_fwd(_ptr(_f32(so)), _ptr(pair_idx), _ptr(labels), _ptr(_f32(embed1)),
                                               _ptr(_f32(embed2)), pair_idx.shape[0], _ptr(_f32(tok_f32)),
                                               _ptr(_bf(tok_bf16)), _stream()), "pair_concat_fwd")
    _count()


def pair_concat_bwd(dtok, pair_idx, labels, dso, dembed1=None, dembed2=None):
    assert dtok.is_contiguous() and dso.is_contiguous()
    check(_lib.lib().b200vsgg_pair_concat_bwd(_ptr(_f32(dtok)), _ptr(pair_idx), _ptr(labels), pair_idx.shape[0],
                                               _ptr(_f32(dso)), _ptr(_f32(dembed1)), _ptr(_f32(dembed2)), _stream()),
          "pair_concat_bwd")
    _count()


def layernorm_fwd(x, gamma, beta, eps=1e-5, y_f32=None, y_bf16=None, add_table=None, add_idx=None,
                  y_bf16_added=None, mean=None, rstd=None):
    rows, cols = x.shape
    check(_lib.lib().b200vsgg_layernorm_fwd(
        _ptr(_f32(x)), x.stride(0), _ptr(_f32(gamma)), _ptr(_f32(beta)), rows, cols, eps, _ptr(_f32(y_f32)),
        _ld(y_f32), _ptr(_bf(y_bf16)), _ld(y_bf16), _ptr(_f32(add_table)), _ptr(add_idx), _ptr(_bf(y_bf16_added)),
        _ld(y_bf16_added), _ptr(mean), _ptr(rstd), _stream()), "layernorm_fwd")
    _count()


def layernorm_bwd(dy, x, gamma, mean, rstd, dx_f32=None, dx_bf16=None, drop_p=0.0, seed=0, dgamma=None, dbeta=None,
                  base=None):
    """dx = LN'(dy) (+ base: the skip-path gradient of a pre-LN residual branch)."""
    rows, cols = x.shape
    check(_lib.lib().b200vsgg_layernorm_bwd_add(
        _ptr(_f32(dy)), dy.stride(0), _ptr(_f32(x)), x.stride(0), _ptr(_f32(gamma)), _ptr(mean), _ptr(rstd), rows,
        cols, _ptr(_f32(dx_f32)), _ld(dx_f32), _ptr(_bf(dx_bf16)), _ld(dx_bf16), drop_p, seed, _ptr(_f32(dgamma)),
        _ptr(_f32(dbeta)), _stream(), _ptr(_f32(base)), _ld(base)), "layernorm_bwd")
    _count()


def cast_bf16(x, out=None, drop_p=0.0, seed=0):
    """bf16(dropout(x)) for a 2-D fp32 tensor (row stride allowed)."""
    rows, cols = x.shape
    if out is None:
        out = torch.empty(rows, cols, dtype=torch.bfloat16, device=x.device)
    check(_lib.lib().b200vsgg_cast_dropout_bf16(_ptr(_f32(x)), x.stride(0), rows, cols, _ptr(_bf(out)),
                                                 out.stride(0), drop_p, seed, _stream()), "cast_dropout_bf16")
    _count()
    return out


def split3_bf16(x, out=None):
    """[rows, cols] fp32 -> [rows, 3*cols] bf16 = (hi | lo | hi), hi + lo == x to ~2^-17 (b200vsgg_split3_bf16)."""
    rows, cols = x.shape
    if out is None:
        out = torch.empty(rows, 3 * cols, dtype=torch.bfloat16, device=x.device)
    check(_lib.lib().b200vsgg_split3_bf16(_ptr(_f32(x)), x.stride(0), rows, cols, _ptr(_bf(out)), out.stride(0),
                                           _stream()), "split3_bf16")
    _count()
    return out


# ------------------------------------------------------------------------------------------------
# bf16 operand copies of weights, refreshed only when a parameter changed
# ------------------------------------------------------------------------------------------------
_NO_WCACHE = bool(int(__import__("os").environ.get("B200VSGG_NO_WCACHE", "0")))


def cached_weight(tag, params, build):
    """`build()` -> tensor derived from the parameters `params` (bf16 copy, packed / permuted / split-precision layout),
    memoised until one of them changes.  The memo lives ON the first parameter object (so it dies with the model and
    can never be hit by another model whose tensors reuse the same addresses) and is validated by the identity of the
    other parameters plus (`_version`, data_ptr) of each: torch bumps `_version` on every in-place update
    (optimizer.step, load_state_dict, .copy_); b200vsgg.optim.FusedAdamW writes parameters from its own kernel and bumps
    the counters explicitly.  Inference and the repeated forwards of a training step therefore cast each weight once per
    optimiser step instead of once per call.  The returned tensor is read-only (saved activations may alias it)."""
    import weakref
    if _NO_WCACHE:
        with torch.no_grad():
            return build()
    p0 = params[0]
    memo = p0.__dict__.get("_b200vsgg_wcache")
    if memo is None:
        memo = p0.__dict__["_b200vsgg_wcache"] = {}
    stamp = tuple((p._version, p.data_ptr()) for p in params)
    hit = memo.get(tag)
    if hit is not None and hit[0] == stamp and len(hit[1]) == len(params) - 1 and \
            all(r() is p for r, p in zip(hit[1], params[1:])):
        return hit[2]
    with torch.no_grad():
        t = build()
    memo[tag] = (stamp, tuple(weakref.ref(p) for p in params[1:]), t)
    return t


_uniform_chunks = {}


def uniform_chunks(rows, device, chunk_rows=256):
    """int32 [n,3] chunk table (row_begin, row_end, group 0) covering `rows` rows (cached)."""
    key = (rows, chunk_rows, str(device))
    t = _uniform_chunks.get(key)
    if t is None:
        starts = torch.arange(0, rows, chunk_rows, dtype=torch.int32)
        t = torch.stack([starts, torch.clamp(starts + chunk_rows, max=rows), torch.zeros_like(starts)], 1).contiguous()
        t = t.to(device)
        _uniform_chunks[key] = t
    return t


def seg_colstats(a, chunks, sum1, b=None, sum2=None):
    """sum1[g,:] += column sums of a over the rows of each chunk's group; sum2[g,:] += sums of a*b."""
    rows, cols = a.shape
    assert a.stride(1) == 1 and chunks.dtype == torch.int32 and chunks.is_contiguous() and chunks.shape[1] == 3
    assert sum1.dtype == torch.float32 and sum1.is_contiguous() and sum1.shape[-1] == cols
    if b is not None:
        assert (b.dtype == torch.bfloat16 or b is a) and b.stride(1) == 1 and b.shape == a.shape and sum2 is not None
        assert sum2.dtype == torch.float32 and sum2.is_contiguous()
    check(_lib.lib().b200vsgg_seg_colstats(_ptr(a), 1 if a.dtype == torch.bfloat16 else 0, a.stride(0), _ptr(b),
                                            b.stride(0) if b is not None else 0, cols, _ptr(chunks), chunks.shape[0],
                                            _ptr(sum1), _ptr(sum2), _stream()), "seg_colstats")
    _count()


def colsum(x, out, group_idx=None, n_groups=1):
    """out[g,:] += column sums of x over rows of group g."""
    rows, cols = x.shape
    assert x.stride(1) == 1 and out.dtype == torch.float32 and out.is_contiguous()
    if group_idx is None and cols % 8 == 0 and x.stride(0) % 8 == 0 and x.data_ptr() % 16 == 0:
        return seg_colstats(x, uniform_chunks(rows, x.device), out)
    check(_lib.lib().b200vsgg_colsum(_ptr(x), 1 if x.dtype == torch.bfloat16 else 0, x.stride(0), rows, cols,
                                      _ptr(group_idx), n_groups, _ptr(out), _stream()), "colsum")
    _count()


def attn_small_fwd(q, k, v, seg_off, n_seg, max_len, n_heads, head_dim, ctx, drop_p=0.0, seed=0, scale=None):
    """scale defaults to head_dim^-0.5; pass it when heads are zero-padded (objbranch: 297 -> 304 columns)."""
    scale = float(head_dim) ** -0.5 if scale is None else float(scale)
    check(_lib.lib().b200vsgg_attn_small_fwd(
        _ptr(_bf(q)), q.stride(0), _ptr(_bf(k)), k.stride(0), _ptr(_bf(v)), v.stride(0), _ptr(seg_off), n_seg,
        max_len, n_heads, head_dim, scale, _ptr(_bf(ctx)), ctx.stride(0), drop_p, seed, _stream()), "attn_small_fwd")
    _count()


def attn_small_bwd(q, k, v, dctx, seg_off, n_seg, max_len, n_heads, head_dim, dq, dk, dv, drop_p=0.0, seed=0,
                   scale=None):
    scale = float(head_dim) ** -0.5 if scale is None else float(scale)
    check(_lib.lib().b200vsgg_attn_small_bwd(
        _ptr(_bf(q)), q.stride(0), _ptr(_bf(k)), k.stride(0), _ptr(_bf(v)), v.stride(0), _ptr(_bf(dctx)),
        dctx.stride(0), _ptr(seg_off), n_seg, max_len, n_heads, head_dim, scale, _ptr(_bf(dq)), dq.stride(0),
        _ptr(_bf(dk)), dk.stride(0), _ptr(_bf(dv)), dv.stride(0), drop_p, seed, _stream()), "attn_small_bwd")
    _count()


def attn_rows_fwd(q, k, v, seg_off, n_seg, n_heads, head_dim, ctx, lse=None, drop_p=0.0, seed=0, scale=None):
    """Varlen attention for sequences of any length / head_dim <= 320 (b200vsgg_attn_rows_fwd)."""
    scale = float(head_dim) ** -0.5 if scale is None else float(scale)
    check(_lib.lib().b200vsgg_attn_rows_fwd(
        _ptr(_bf(q)), q.stride(0), _ptr(_bf(k)), k.stride(0), _ptr(_bf(v)), v.stride(0), _ptr(seg_off), n_seg, n_heads,
        head_dim, scale, _ptr(_bf(ctx)), ctx.stride(0), _ptr(lse), drop_p, seed, _stream()), "attn_rows_fwd")
    _count()


def attn_rows_bwd(q, k, v, ctx, dctx, lse, seg_off, n_seg, n_heads, head_dim, dq, dk, dv, drop_p=0.0, seed=0, scale=None):
    scale = float(head_dim) ** -0.5 if scale is None else float(scale)
    delta = torch.empty_like(lse)
    check(_lib.lib().b200vsgg_attn_rows_bwd(
        _ptr(_bf(q)), q.stride(0), _ptr(_bf(k)), k.stride(0), _ptr(_bf(v)), v.stride(0), _ptr(_bf(ctx)), ctx.stride(0),
        _ptr(_bf(dctx)), dctx.stride(0), _ptr(lse), _ptr(delta), _ptr(seg_off), n_seg, n_heads, head_dim, scale,
        _ptr(_bf(dq)), dq.stride(0), _ptr(_bf(dk)), dk.stride(0), _ptr(_bf(dv)), dv.stride(0), drop_p, seed, _stream()),
        "attn_rows_bwd")
    _count(2)


def _gmm_heads_array(specs):
    from ._decls import GmmHead
    arr = (GmmHead * len(specs))()
    for i, s in enumerate(specs):
        arr[i].col_base = s["col_base"]
        arr[i].num_classes = s["num_classes"]
        arr[i].softmax = int(s["softmax"])        # 0 sigmoid, 1 softmax, 2 softmax without the background class (mode 0)
        for f in ("eps", "out", "out2", "dout"):
            t = s.get(f)
            if t is not None:
                assert t.dtype == torch.float32 and t.is_contiguous()
            setattr(arr[i], f, t.data_ptr() if t is not None else None)
    return arr


def gmm_head_fwd(z, K, specs, mode, seed=0):
    arr = _gmm_heads_array(specs)
    check(_lib.lib().b200vsgg_gmm_head_fwd(_ptr(_f32(z)), z.stride(0), z.shape[0], K, arr, len(specs), mode, seed,
                                            _stream()), "gmm_head_fwd")
    _count()


def gmm_head_bwd(z, K, specs, mode, dz_bf16, seed=0):
    arr = _gmm_heads_array(specs)
    check(_lib.lib().b200vsgg_gmm_head_bwd(_ptr(_f32(z)), z.stride(0), z.shape[0], K, arr, len(specs), mode, seed,
                                            _ptr(_bf(dz_bf16)), dz_bf16.stride(0), dz_bf16.shape[1], _stream()),
          "gmm_head_bwd")
    _count()


def nchw_to_nhwc_bf16(x, out=None):
    """fp32 [n,C,H,W] (contiguous) -> bf16 rows [n*H*W, C]."""
    assert x.dtype == torch.float32 and x.is_contiguous()
    n, c = x.shape[0], x.shape[1]
    s = x[0, 0].numel() if n > 0 else 1
    if out is None:
        out = torch.empty(n * s, c, dtype=torch.bfloat16, device=x.device)
    check(_lib.lib().b200vsgg_nchw_to_nhwc_bf16(_ptr(x), n, c, s, _ptr(out), _stream()), "nchw_to_nhwc_bf16")
    _count(max(1, (n + 65534) // 65535))
    return out


def nchw_to_nhwc_f32(x, out=None):
    assert x.dtype == torch.float32 and x.is_contiguous()
    n, c = x.shape[0], x.shape[1]
    s = x[0, 0].numel() if n > 0 else 1
    if out is None:
        out = torch.empty(n * s, c, dtype=torch.float32, device=x.device)
    check(_lib.lib().b200vsgg_nchw_to_nhwc_f32(_ptr(x), n, c, s, _ptr(out), _stream()), "nchw_to_nhwc_f32")
    _count(max(1, (n + 65534) // 65535))
    return out


def nhwc_to_nchw_f32(x, n, c, hw_shape):
    """fp32 rows [n*S, C] -> fp32 [n, C, *hw_shape]."""
    assert x.dtype == torch.float32 and x.is_contiguous()
    s = 1
    for d in hw_shape:
        s *= d
    out = torch.empty((n, c) + tuple(hw_shape), dtype=torch.float32, device=x.device)
    check(_lib.lib().b200vsgg_nhwc_to_nchw_f32(_ptr(x), n, c, s, _ptr(out), _stream()), "nhwc_to_nchw_f32")
    _count(max(1, (n + 65534) // 65535))
    return out


# ------------------------------------------------------------------------------------------------
# spatial-mask branch (lib/tempura.py:466-474), channels-last
# ------------------------------------------------------------------------------------------------
def mask_im2col(masks, out):
    """masks fp32 (the reference's hand-off) or bf16 (producer-side hand-off, (f).4) [n,2,27,27] -> im2col rows."""
    assert masks.dtype in (torch.float32, torch.bfloat16) and masks.is_contiguous() and masks.shape[1:] == (2, 27, 27)
    assert out.dtype == torch.bfloat16 and out.is_contiguous() and out.shape[0] == masks.shape[0] * 196
    fn = _lib.lib().b200vsgg_mask_im2col if masks.dtype == torch.float32 else _lib.lib().b200vsgg_mask_im2col_bf16
    check(fn(_ptr(masks), masks.shape[0], _ptr(out), out.shape[1], _stream()), "mask_im2col")
    _count()


def seg_affine(a, b, k1, k2, k3, group_of_unit, rows_per_unit, out, relu_mask=False):
    rows, cols = b.shape
    for t in (a, b, out):
        assert t is None or (t.dtype == torch.bfloat16 and t.is_contiguous())
    for t in (k1, k2, k3):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous() and t.shape[-1] == cols)
    assert group_of_unit.dtype == torch.int32
    check(_lib.lib().b200vsgg_seg_affine(_ptr(a), _ptr(b), _ptr(k1), _ptr(k2), _ptr(k3), _ptr(group_of_unit), rows,
                                          rows_per_unit, cols, int(relu_mask), _ptr(out), _stream()), "seg_affine")
    _count()


def bn_pool_fwd(y, scale, shift, group_of_unit, n, hw_in, channels, z, argmax):
    assert y.dtype in (torch.bfloat16, torch.float32) and y.is_contiguous()
    assert z.dtype == torch.bfloat16 and argmax.dtype == torch.uint8
    assert scale.is_contiguous() and shift.is_contiguous() and group_of_unit.dtype == torch.int32
    check(_lib.lib().b200vsgg_bn_pool_fwd(_ptr(y), 1 if y.dtype == torch.float32 else 0, _ptr(_f32(scale)), _ptr(_f32(shift)), _ptr(group_of_unit), n, hw_in,
                                           channels, _ptr(z), _ptr(argmax), _stream()), "bn_pool_fwd")
    _count()


def pool_bwd(dz, argmax, n, hw_in, channels, dy):
    assert dz.dtype == torch.bfloat16 and dz.is_contiguous() and dy.dtype == torch.bfloat16 and dy.is_contiguous()
    check(_lib.lib().b200vsgg_pool_bwd(_ptr(dz), _ptr(argmax), n, hw_in, channels, _ptr(dy), _stream()), "pool_bwd")
    _count()


def im2col3x3(z, n, hw, channels, out):
    assert z.dtype == torch.bfloat16 and z.is_contiguous() and out.dtype == torch.bfloat16 and out.is_contiguous()
    check(_lib.lib().b200vsgg_im2col3x3(_ptr(z), n, hw, channels, _ptr(out), _stream()), "im2col3x3")
    _count()


def col2im3x3(dcol, n, hw, channels, dz):
    assert dcol.dtype == torch.bfloat16 and dcol.is_contiguous() and dz.dtype == torch.bfloat16 and dz.is_contiguous()
    check(_lib.lib().b200vsgg_col2im3x3(_ptr(dcol), n, hw, channels, _ptr(dz), _stream()), "col2im3x3")
    _count()


# ------------------------------------------------------------------------------------------------
# TEAT-GT / TokenGT path
# ------------------------------------------------------------------------------------------------
def attn_flash_fwd(q, k, v, seq_off, blk_seq, blk_row0, n_heads, head_dim, ctx, lse=None, drop_p=0.0, seed=0, max_len=0):
    scale = float(head_dim) ** -0.5
    check(_lib.lib().b200vsgg_attn_flash_fwd(
        _ptr(_bf(q)), q.stride(0), _ptr(_bf(k)), k.stride(0), _ptr(_bf(v)), v.stride(0), _ptr(seq_off), _ptr(blk_seq),
        _ptr(blk_row0), blk_seq.numel(), n_heads, head_dim, scale, _ptr(_bf(ctx)), ctx.stride(0), _ptr(lse), drop_p, seed,
        _stream(), seq_off.numel() - 1 if max_len > 0 else 0, max_len), "attn_flash_fwd")
    _count()


def _attn_prof_begin():
    if attn_profile is None:
        return None
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record()
    return e0


def _attn_prof_end(kind, flops, e0):
    if e0 is not None:
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        attn_profile.append((kind, flops, e0, e1))


def attn_tc_fwd(q, k, v, seq_off, blk_seq128, blk_row0_128, n_heads, head_dim, ctx, lse=None, drop_p=0.0, seed=0,
                flops=0.0):
    """tcgen05 / TMEM / TMA forward (b200vsgg_attn_tc_fwd); the block table has 128-row blocks."""
    scale = float(head_dim) ** -0.5
    rows = q.shape[0]
    assert k.shape[0] == rows and v.shape[0] == rows and ctx.shape[0] == rows
    e0 = _attn_prof_begin()
    check(_lib.lib().b200vsgg_attn_tc_fwd(
        _ptr(_bf(q)), q.stride(0), _ptr(_bf(k)), k.stride(0), _ptr(_bf(v)), v.stride(0), rows, _ptr(seq_off),
        _ptr(blk_seq128), _ptr(blk_row0_128), blk_seq128.numel(), n_heads, head_dim, scale, _ptr(_bf(ctx)), ctx.stride(0),
        _ptr(lse), drop_p, seed, _stream()), "attn_tc_fwd")
    _attn_prof_end("fwd", flops, e0)
    _count()


def attn_tc_bwd(q, k, v, ctx, dctx, lse, seq_off, blk_seq128, blk_row0_128, n_heads, head_dim, dq, dk, dv, drop_p=0.0, seed=0,
                flops=0.0):
    """tcgen05 / TMEM backward (b200vsgg_attn_tc_bwd): delta, dQ and dK/dV kernels."""
    scale = float(head_dim) ** -0.5
    rows = q.shape[0]
    delta = torch.empty(rows, n_heads, device=q.device, dtype=torch.float32)
    e0 = _attn_prof_begin()
    check(_lib.lib().b200vsgg_attn_tc_bwd(
        _ptr(_bf(q)), q.stride(0), _ptr(_bf(k)), k.stride(0), _ptr(_bf(v)), v.stride(0), _ptr(_bf(ctx)), ctx.stride(0),
        _ptr(_bf(dctx)), dctx.stride(0), _ptr(lse), _ptr(delta), rows, _ptr(seq_off), _ptr(blk_seq128), _ptr(blk_row0_128),
        blk_seq128.numel(), n_heads, head_dim, scale, _ptr(_bf(dq)), dq.stride(0), _ptr(_bf(dk)), dk.stride(0),
        _ptr(_bf(dv)), dv.stride(0), drop_p, seed, _stream()), "attn_tc_bwd")
    _attn_prof_end("bwd", flops, e0)
    _count(3)


def attn_flash_bwd(q, k, v, ctx, dctx, lse, seq_off, blk_seq, blk_row0, n_heads, head_dim, dq, dk, dv, drop_p=0.0, seed=0,
                   max_len=0, flops=0.0):
    scale = float(head_dim) ** -0.5
    rows = q.shape[0]
    delta = torch.empty(rows, n_heads, device=q.device, dtype=torch.float32)
    e0 = _attn_prof_begin()
    check(_lib.lib().b200vsgg_attn_flash_bwd(
        _ptr(_bf(q)), q.stride(0), _ptr(_bf(k)), k.stride(0), _ptr(_bf(v)), v.stride(0), _ptr(_bf(ctx)), ctx.stride(0),
        _ptr(_bf(dctx)), dctx.stride(0), _ptr(lse), _ptr(delta), _ptr(seq_off), _ptr(blk_seq), _ptr(blk_row0),
        blk_seq.numel(), rows, n_heads, head_dim, scale, _ptr(_bf(dq)), dq.stride(0), _ptr(_bf(dk)), dk.stride(0),
        _ptr(_bf(dv)), dv.stride(0), drop_p, seed, _stream(), seq_off.numel() - 1 if max_len > 0 else 0, max_len),
          "attn_flash_bwd")
    _attn_prof_end("bwd", flops, e0)
    _count(3)


def node_tokens_fwd(so, feat_row, is_person, labels, embed, h1, out_f32, out_bf16):
    n = feat_row.numel()
    assert so.stride(1) == 1 and feat_row.dtype == torch.int32 and is_person.dtype == torch.int32
    assert labels.dtype == torch.int64 and embed.is_contiguous() and out_f32.is_contiguous()
    check(_lib.lib().b200vsgg_node_tokens_fwd(_ptr(_f32(so)), so.stride(0), _ptr(feat_row), _ptr(is_person), _ptr(labels),
                                               _ptr(_f32(embed)), n, h1, embed.shape[1], _ptr(_f32(out_f32)),
                                               _ptr(_bf(out_bf16)), _stream()), "node_tokens_fwd")
    _count()


def node_tokens_bwd(dtok, feat_row, is_person, labels, h1, e, dso, dembed):
    assert dtok.is_contiguous() and dso.stride(1) == 1
    check(_lib.lib().b200vsgg_node_tokens_bwd(_ptr(_f32(dtok)), _ptr(feat_row), _ptr(is_person), _ptr(labels),
                                               feat_row.numel(), h1, e, _ptr(_f32(dso)), dso.stride(0), _ptr(dembed),
                                               _stream()), "node_tokens_bwd")
    _count()


def teat_pair_flags(tok, boxes, feat_row, node_off, has_prev, thr, sim, nmax):
    n_frames = has_prev.numel()
    assert tok.is_contiguous() and boxes.is_contiguous() and boxes.shape[1] == 5
    spatial = torch.empty(n_frames, nmax, nmax, dtype=torch.uint8, device=tok.device)
    temporal = torch.empty(n_frames, nmax, nmax, dtype=torch.uint8, device=tok.device)
    check(_lib.lib().b200vsgg_teat_pair_flags(_ptr(_f32(tok)), tok.shape[1], _ptr(_f32(boxes)), _ptr(feat_row),
                                               _ptr(node_off), _ptr(has_prev), n_frames, thr, sim, nmax, _ptr(spatial),
                                               _ptr(temporal), _stream()), "teat_pair_flags")
    _count()
    return spatial, temporal


def teat_assemble_fwd(desc, na, pu, pv, temp, eemb, order, graph_tok, null_tok, x):
    for t in (na, pu, pv, temp, eemb, order, graph_tok, null_tok, x):
        assert t.dtype == torch.float32 and t.is_contiguous()
    assert desc.dtype == torch.int32 and desc.is_contiguous() and desc.shape[1] == 4
    check(_lib.lib().b200vsgg_teat_assemble_fwd(_ptr(desc), desc.shape[0], x.shape[1], _ptr(na), _ptr(pu), _ptr(pv),
                                                 _ptr(temp), _ptr(eemb), _ptr(order), _ptr(graph_tok), _ptr(null_tok),
                                                 _ptr(x), _stream()), "teat_assemble_fwd")
    _count()


def teat_assemble_bwd(desc, dx, dna, dpu, dpv, dtemp, deemb, dorder, dgraph, dnull):
    for t in (dx, dna, dpu, dpv, dtemp, deemb, dorder, dgraph, dnull):
        assert t.dtype == torch.float32 and t.is_contiguous()
    check(_lib.lib().b200vsgg_teat_assemble_bwd(_ptr(desc), desc.shape[0], dx.shape[1], _ptr(dx), _ptr(dna), _ptr(dpu),
                                                 _ptr(dpv), _ptr(dtemp), _ptr(deemb), _ptr(dorder), _ptr(dgraph),
                                                 _ptr(dnull), _stream()), "teat_assemble_bwd")
    _count()


def act_dropout(x, act, p=0.0, seed=0, out=None):
    rows, cols = x.shape
    if out is None:
        out = torch.empty(rows, cols, dtype=torch.bfloat16, device=x.device)
    check(_lib.lib().b200vsgg_act_dropout_bf16(_ptr(_bf(x)), x.stride(0), rows, cols, act, p, seed, _ptr(_bf(out)),
                                                out.stride(0), _stream()), "act_dropout_bf16")
    _count()
    return out


def consistency_kl(g, pair_u, pair_v):
    """out[p] = KL(softmax(g[v]) || softmax(g[u])) / (v - u) for the frame pairs (u < v) of each clip."""
    assert g.dtype == torch.float32 and g.is_contiguous() and pair_u.dtype == torch.int32 and pair_v.dtype == torch.int32
    out = torch.empty(pair_u.numel(), device=g.device, dtype=torch.float32)
    check(_lib.lib().b200vsgg_consistency_kl(_ptr(g), g.shape[1], _ptr(pair_u), _ptr(pair_v), pair_u.numel(), _ptr(out),
                                              _stream()), "consistency_kl")
    _count()
    return out


def graph_attn_core(qkv, node_off, upper, nmax, we, be, out):
    assert qkv.dtype == torch.float32 and qkv.stride(1) == 1 and upper.dtype == torch.uint8 and upper.is_contiguous()
    assert we.is_contiguous() and be.is_contiguous() and out.dtype in (torch.bfloat16, torch.float32)
    fn = _lib.lib().b200vsgg_graph_attn_core if out.dtype == torch.bfloat16 else _lib.lib().b200vsgg_graph_attn_core_f32
    check(fn(_ptr(qkv), qkv.stride(0), _ptr(node_off), _ptr(upper), nmax, _ptr(_f32(we)), _ptr(_f32(be)),
             node_off.numel() - 1, _ptr(out), out.stride(0), _stream()), "graph_attn_core")
    _count()


def gated_residual(o, res, w):
    assert o.is_contiguous() and res.is_contiguous() and w.is_contiguous() and w.numel() == 3 * o.shape[1]
    check(_lib.lib().b200vsgg_gated_residual(_ptr(_f32(o)), _ptr(_f32(res)), _ptr(_f32(w)), o.shape[0], o.shape[1],
                                              _stream()), "gated_residual")
    _count()


# ------------------------------------------------------------------------------------------------
# zero-copy uploads of small host arrays (segment plans) through a pinned ring buffer
# ------------------------------------------------------------------------------------------------
def _obj_tokens_struct(a):
    from ._decls import ObjTokens
    st = ObjTokens()
    for name in ("features", "dist", "embed", "boxes", "bn_mean", "bn_rstd", "bn_gamma", "bn_beta", "wp", "bp", "pe"):
        t = a.get(name)
        if t is not None:
            assert t.dtype == torch.float32 and t.is_contiguous(), name
        setattr(st, name, t.data_ptr() if t is not None else None)
    for name in ("video_of_box", "src", "pos"):
        t = a.get(name)
        if t is not None:
            assert t.dtype == torch.int32 and t.is_contiguous(), name
        setattr(st, name, t.data_ptr() if t is not None else None)
    st.feat_dim, st.n_cls, st.e, st.h = a["features"].shape[1], a["dist"].shape[1], a["embed"].shape[1], a["wp"].shape[0]
    assert a["embed"].shape[0] == st.n_cls and a["wp"].shape[1] == 4 and a["boxes"].shape[1] == 5
    st.rows = a["rows"]
    st.p_pos, st.seed_pos, st.p_pe, st.seed_pe = a.get("p_pos", 0.0), a.get("seed_pos", 0), a.get("p_pe", 0.0), a.get("seed_pe", 0)
    return st


def obj_tokens_fwd(args, x_f32=None, x_bf16=None):
    """b200vsgg_obj_tokens_fwd; `args` is a dict with the fields of struct b200vsgg_obj_tokens (tensors)."""
    st = _obj_tokens_struct(args)
    for t in (x_f32, x_bf16):
        assert t is None or (t.is_contiguous() and t.shape == (st.rows, st.feat_dim + st.e + st.h))
    check(_lib.lib().b200vsgg_obj_tokens_fwd(C.byref(st), _ptr(x_f32), _ptr(x_bf16), _stream()), "obj_tokens_fwd")
    _count()


def obj_tokens_bwd(args, dx, dembed, dwp, dbp, dgamma, dbeta):
    st = _obj_tokens_struct(args)
    assert dx.dtype == torch.float32 and dx.is_contiguous() and dx.shape == (st.rows, st.feat_dim + st.e + st.h)
    for t in (dembed, dwp, dbp, dgamma, dbeta):
        assert t.dtype == torch.float32 and t.is_contiguous()
    check(_lib.lib().b200vsgg_obj_tokens_bwd(C.byref(st), _ptr(dx), _ptr(dembed), _ptr(dwp), _ptr(dbp), _ptr(dgamma),
                                              _ptr(dbeta), _stream()), "obj_tokens_bwd")
    _count()


def rel_loss(att, spa, con, att_label, spa_t, con_t, row_w, want_grad=True):
    """b200vsgg_rel_loss.  spa_t / con_t: dense fp32 multi-hot [n,C] or a CSR pair (off int32 [n+1], idx int32).
    Returns (losses [3], d_att, d_spa, d_con) — gradients of the summed losses (None if not wanted)."""
    n = att.shape[0]
    for t in (att, spa, con, row_w):
        assert t.dtype == torch.float32 and t.is_contiguous()
    assert att_label.dtype == torch.int64 and att_label.is_contiguous()

    def lab(t):
        if isinstance(t, (tuple, list)):
            off, idx = t
            assert off.dtype == torch.int32 and idx.dtype == torch.int32 and off.numel() == n + 1
            return None, off, idx
        assert t.dtype == torch.float32 and t.is_contiguous()
        return t, None, None

    sd, so, si = lab(spa_t)
    cd, co, ci = lab(con_t)
    losses = torch.zeros(3, device=att.device)
    grads = [torch.empty_like(t) if want_grad else None for t in (att, spa, con)]
    check(_lib.lib().b200vsgg_rel_loss(_ptr(att), _ptr(spa), _ptr(con), n, att.shape[1], spa.shape[1], con.shape[1],
                                        _ptr(att_label), _ptr(sd), _ptr(cd), _ptr(so), _ptr(si), _ptr(co), _ptr(ci),
                                        _ptr(row_w), _ptr(losses), _ptr(grads[0]), _ptr(grads[1]), _ptr(grads[2]),
                                        _stream()), "rel_loss")
    _count()
    return (losses, *grads)


def contrastive_loss(x, label, seg_off, max_rows, pos_margin=0.0, neg_margin=1.0, want_grad=True):
    """b200vsgg_contrastive_loss: per-segment ContrastiveLoss values [n_seg] and d(loss[v])/dx (or None)."""
    assert x.dtype == torch.float32 and x.is_contiguous() and label.dtype == torch.int32 and seg_off.dtype == torch.int32
    n_seg = seg_off.numel() - 1
    loss = torch.empty(n_seg, device=x.device)
    dx = torch.empty_like(x) if want_grad else None
    check(_lib.lib().b200vsgg_contrastive_loss(_ptr(x), _ptr(label), _ptr(seg_off), n_seg, int(max_rows), x.shape[1],
                                                pos_margin, neg_margin, _ptr(loss), _ptr(dx), None, _stream()),
          "contrastive_loss")
    _count()
    return loss, dx


def graph_small_fwd(nodes, upper, counts, dim, heads, depth, params, pool_w, pool_b):
    """b200vsgg_graph_small_fwd: nodes fp32 [F, nmax, dim], upper uint8 [F, nmax, nmax], counts int32 [F] -> [F, dim]."""
    F_, nmax = nodes.shape[0], nodes.shape[1]
    assert nodes.dtype == torch.float32 and nodes.is_contiguous() and nodes.shape[2] == dim
    assert upper.dtype == torch.uint8 and upper.is_contiguous() and upper.shape == (F_, nmax, nmax)
    assert counts.dtype == torch.int32 and params.dtype == torch.float32 and params.is_contiguous()
    per_layer = _lib.lib().b200vsgg_graph_small_params_per_layer(dim, heads)
    assert params.numel() == depth * per_layer, (params.numel(), depth, per_layer)
    out = torch.empty(F_, dim, device=nodes.device)
    check(_lib.lib().b200vsgg_graph_small_fwd(_ptr(nodes), _ptr(upper), _ptr(counts), F_, nmax, dim, heads, depth,
                                               _ptr(params), _ptr(_f32(pool_w)), _ptr(_f32(pool_b)), _ptr(out), _stream()),
          "graph_small_fwd")
    _count()
    return out


class _UploadRing:
    """Pinned staging ring (default 64 MB) of ONE device.  upload(): memcpy into the ring on the host, then a kernel on
    the current stream reads it over PCIe (b200vsgg_upload).  A slot is reused only after the event recorded behind its
    last reader has completed: on wrap-around every outstanding event is synchronised, and the event list is only
    ever trimmed after synchronising what is dropped — readers may sit on any stream of the device."""

    def __init__(self, nbytes=64 << 20):
        self.buf = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        self.np = self.buf.numpy()
        self.pos = 0
        self.events = []          # events in issue order (one per upload)

    def _reserve(self, n):
        if self.pos + n > self.buf.numel():
            self.pos = 0
            for ev in self.events:             # wrap-around: everything issued so far must have been read
                ev.synchronize()
            self.events = []
        start = self.pos
        self.pos += n
        return start

    def _trim(self):
        if len(self.events) > 4096:            # completed long ago in practice; synchronise before forgetting them
            for ev in self.events[:-2048]:
                ev.synchronize()
            self.events = self.events[-2048:]


_rings = {}                       # device index -> _UploadRing
_ring_lock = __import__("threading").Lock()


def upload(arr, device, dtype=None):
    """numpy array -> device tensor of the same shape/dtype (or `dtype`), without using the DMA copy engine.
    Thread-safe; arrays larger than the ring take a one-off pinned staging buffer."""
    import numpy as np
    a = np.ascontiguousarray(arr if dtype is None else np.asarray(arr).astype(dtype))
    t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0].reshape(-1)).dtype, device=device)
    nbytes = a.nbytes
    if nbytes == 0:
        return t
    dev = torch.device(device)
    if dev.type != "cuda":
        t.copy_(torch.from_numpy(a))
        return t
    padded = (nbytes + 15) // 16 * 16
    out = t if nbytes == padded and t.data_ptr() % 16 == 0 else None
    dst = out if out is not None else torch.empty(padded, dtype=torch.uint8, device=device)
    with _ring_lock:
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        ring = _rings.get(idx)
        if ring is None:
            ring = _rings[idx] = _UploadRing()
        if padded > ring.buf.numel():                         # oversized: one-off pinned buffer, freed once it was read
            stage = torch.empty(padded, dtype=torch.uint8).pin_memory()
            stage.numpy()[:nbytes] = a.reshape(-1).view(np.uint8)
            check(_lib.lib().b200vsgg_upload(C.c_void_p(stage.data_ptr()), _ptr(dst), padded, _stream()), "upload")
            torch.cuda.current_stream().synchronize()
        else:
            start = ring._reserve(padded)
            ring.np[start:start + nbytes] = a.reshape(-1).view(np.uint8)
            check(_lib.lib().b200vsgg_upload(C.c_void_p(ring.buf.data_ptr() + start), _ptr(dst), padded, _stream()), "upload")
            ev = torch.cuda.Event()
            ev.record()
            ring.events.append(ev)
            ring._trim()
    _count()
    if out is None:
        t = dst[:nbytes].view(t.dtype).view(a.shape)
    return t


def grad_sqnorm(tensors, chunk_tensor, chunk_off, chunk_elems, sq_norm):
    check(_lib.lib().b200vsgg_grad_sqnorm(_ptr(tensors), _ptr(chunk_tensor), _ptr(chunk_off), chunk_tensor.numel(),
                                           chunk_elems, _ptr(sq_norm), _stream()), "grad_sqnorm")
    _count()


def adamw_clip_step(tensors, chunk_tensor, chunk_off, chunk_elems, sq_norm, max_norm, lr, beta1, beta2, eps, weight_decay):
    check(_lib.lib().b200vsgg_adamw_clip_step(_ptr(tensors), _ptr(chunk_tensor), _ptr(chunk_off), chunk_tensor.numel(),
                                               chunk_elems, _ptr(sq_norm), max_norm, lr, beta1, beta2, eps, weight_decay,
                                               _stream()), "adamw_clip_step")
    _count()


def eval_recall(pair_idx, frame_off, n_frames, att, spa, con, boxes5, classes, scores, gt_boxes, gt_classes, gt_box_off,
                gt_rels, gt_rel_off, mode, semi_thr, iou_thr):
    """b200vsgg_eval_recall: Recall@K hit flags [G,4] (K = 10/20/50/100) of every ground-truth relation of one video and a
    status word (1: a frame exceeded the kernel's tables).  boxes5 = pred['boxes'] [O,5] (column 0 is the frame index)."""
    dev = pair_idx.device
    G = gt_rels.numel() // 3
    hits = torch.empty(G, 4, dtype=torch.uint8, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    pi = pair_idx.contiguous().to(torch.int64)
    a, s_, c = (t.detach().contiguous().float() for t in (att, spa, con))
    b5 = boxes5.detach().contiguous().float()
    cl = classes.contiguous().to(torch.int64)
    sc = scores.detach().contiguous().float()
    check(_lib.lib().b200vsgg_eval_recall(
        _ptr(pi), _ptr(frame_off), n_frames, _ptr(a), a.shape[1], _ptr(s_), s_.shape[1], _ptr(c), c.shape[1],
        b5.data_ptr() + 4, b5.stride(0), _ptr(cl), _ptr(sc), _ptr(gt_boxes), _ptr(gt_classes), _ptr(gt_box_off),
        _ptr(gt_rels), _ptr(gt_rel_off), mode, semi_thr, iou_thr, _ptr(hits), _ptr(status), _stream()), "eval_recall")
    _count()
    return hits, status


def interval_kl(dist, gt, intervals):
    """b200vsgg_interval_kl: temporal-consistency score of each interval [s, e) of pair rows -> fp32 [I] (device)."""
    d = dist.detach().contiguous().float()
    out = torch.empty(intervals.shape[0], dtype=torch.float32, device=d.device)
    check(_lib.lib().b200vsgg_interval_kl(_ptr(d), d.shape[1], _ptr(gt), _ptr(intervals), intervals.shape[0], _ptr(out),
                                           _stream()), "interval_kl")
    _count()
    return out


def class_memory_accumulate(feat, ent_row, ent_cls, ent_w, A):
    """b200vsgg_class_memory_accumulate: A[ent_cls[e]] += ent_w[e] * feat[ent_row[e]] (fp32, in place)."""
    f = feat.detach()
    assert f.dtype == torch.float32 and f.stride(1) == 1 and A.dtype == torch.float32 and A.is_contiguous()
    check(_lib.lib().b200vsgg_class_memory_accumulate(_ptr(f), f.stride(0), f.shape[1], _ptr(ent_row), _ptr(ent_cls),
                                                       _ptr(ent_w), ent_row.numel(), A.shape[0], _ptr(A), _stream()),
          "class_memory_accumulate")
    _count()


def attn_pool(x, node_off, n_frames, max_nodes, w, b):
    """b200vsgg_attn_pool: per-frame attention pooling of compact node rows -> fp32 [frames, d]."""
    assert x.dtype == torch.float32 and x.is_contiguous() and node_off.dtype == torch.int32
    out = torch.empty(n_frames, x.shape[1], device=x.device, dtype=torch.float32)
    check(_lib.lib().b200vsgg_attn_pool(_ptr(x), x.shape[1], _ptr(node_off), n_frames, max_nodes,
                                         _ptr(w.detach().reshape(-1).contiguous().float()),
                                         _ptr(b.detach().reshape(-1).contiguous().float()), _ptr(out), _stream()), "attn_pool")
    _count()
    return out


# ---- backward kernels of the differentiable consistency mode (regulariser.py) --------------------------------------
def consistency_kl_bwd(g, pair_u, pair_v, gout, dg):
    check(_lib.lib().b200vsgg_consistency_kl_bwd(_ptr(g), g.shape[1], _ptr(pair_u), _ptr(pair_v), _ptr(gout), pair_u.numel(),
                                                  _ptr(dg), _stream()), "consistency_kl_bwd")
    _count()


def attn_pool_bwd(x, node_off, n_frames, max_nodes, w, b, dout, dx, dgate):
    check(_lib.lib().b200vsgg_attn_pool_bwd(_ptr(x), x.shape[1], _ptr(node_off), n_frames, max_nodes, _ptr(w), _ptr(b),
                                             _ptr(dout), _ptr(dx), _ptr(dgate), _stream()), "attn_pool_bwd")
    _count()


def weighted_colsum(x, wgt, out):
    assert x.stride(1) == 1 and wgt.dtype == torch.float32 and out.dtype == torch.float32
    check(_lib.lib().b200vsgg_weighted_colsum(_ptr(x), 1 if x.dtype == torch.bfloat16 else 0, x.stride(0), x.shape[0],
                                               x.shape[1], _ptr(wgt), _ptr(out), _stream()), "weighted_colsum")
    _count()


def gated_residual_bwd(o, res, w, dx, d_o, d_res, da):
    check(_lib.lib().b200vsgg_gated_residual_bwd(_ptr(o), _ptr(res), _ptr(w), _ptr(dx), o.shape[0], o.shape[1], _ptr(d_o),
                                                  _ptr(d_res), _ptr(da), _stream()), "gated_residual_bwd")
    _count()


def graph_attn_core_bwd(qkv, node_off, upper, nmax, we, be, dout, dqkv, dwe, dbe):
    check(_lib.lib().b200vsgg_graph_attn_core_bwd(_ptr(qkv), qkv.stride(0), _ptr(node_off), _ptr(upper), nmax, _ptr(we),
                                                   _ptr(be), _ptr(dout), dout.stride(0), node_off.numel() - 1, _ptr(dqkv),
                                                   dqkv.stride(0), _ptr(dwe), _ptr(dbe), _stream()), "graph_attn_core_bwd")
    _count()


def simt_linear(x, w, b=None, transposed=False, act=ACT_NONE, want_z=False):
    """fp32 y = act(x @ W^T + b) (transposed=False, w [out, in]) or x @ W (transposed=True: the dgrad of that linear)."""
    assert x.dtype == torch.float32 and x.stride(1) == 1 and w.dtype == torch.float32 and w.is_contiguous()
    n_out, n_in = (w.shape[1], w.shape[0]) if transposed else (w.shape[0], w.shape[1])
    assert x.shape[1] == n_in
    so, si = (1, w.shape[1]) if transposed else (w.shape[1], 1)
    y = torch.empty(x.shape[0], n_out, device=x.device)
    z = torch.empty_like(y) if want_z else None
    check(_lib.lib().b200vsgg_simt_linear(_ptr(x), x.stride(0), _ptr(w), so, si, _ptr(b), x.shape[0], n_out, n_in, act, _ptr(y),
                                           n_out, _ptr(z), _stream()), "simt_linear")
    _count()
    return (y, z) if want_z else y


def simt_wgrad(dy, x):
    dw = torch.zeros(dy.shape[1], x.shape[1], device=x.device)
    check(_lib.lib().b200vsgg_simt_wgrad(_ptr(dy), dy.stride(0), _ptr(x), x.stride(0), x.shape[0], dy.shape[1], x.shape[1],
                                          _ptr(dw), _stream()), "simt_wgrad")
    _count()
    return dw


def gelu_bwd(dy, z):
    dz = torch.empty_like(z)
    check(_lib.lib().b200vsgg_gelu_bwd(_ptr(dy), _ptr(z), z.numel(), _ptr(dz), _stream()), "gelu_bwd")
    _count()
    return dz


def ln_small_fwd(x, g, b):
    y, mean, rstd = torch.empty_like(x), torch.empty(x.shape[0], device=x.device), torch.empty(x.shape[0], device=x.device)
    check(_lib.lib().b200vsgg_ln_small_fwd(_ptr(x), _ptr(g), _ptr(b), x.shape[0], x.shape[1], _ptr(y), _ptr(mean), _ptr(rstd),
                                            _stream()), "ln_small_fwd")
    _count()
    return y, mean, rstd


def ln_small_bwd(dy, x, g, mean, rstd, base=None):
    dx = torch.empty_like(x)
    dg, db = torch.zeros(x.shape[1], device=x.device), torch.zeros(x.shape[1], device=x.device)
    check(_lib.lib().b200vsgg_ln_small_bwd(_ptr(dy), _ptr(x), _ptr(g), _ptr(mean), _ptr(rstd), _ptr(base), x.shape[0],
                                            x.shape[1], _ptr(dx), _ptr(dg), _ptr(db), _stream()), "ln_small_bwd")
    _count()
    return dx, dg, db
