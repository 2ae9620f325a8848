"""SGCls-train object branch (row S1 of SURVEY.md §8a) on the libb200vsgg kernels.

Reference: ObjectClassifier.forward / .classify, lib/tempura.py:185-255 (TEMPURA) and its copy
tools/utils/object_classifier.py:177-233 (TEAT-GT); class sequences from tools/utils/ds_track.py:18-39.

  x0   = [features 2048 | distribution @ obj_embed.weight 200 | pos_embed(center_size(box)) 128]      (:249-252)
  tracking: boxes grouped by arg-max class -> sequences; x = dropout(x0 + pe[frame rank]);
            3 x nn.TransformerEncoderLayer(2376, 8 heads, ffn 1024, post-norm, ReLU) with key padding; scatter back (:186-210)
  intermediate: Linear(2376 -> 1024) + BatchNorm1d + ReLU; head: GMM_head(1024 -> 37) or Linear            (:103-110)

B200 design: sequences are unpadded row segments (a permutation of the boxes, host-planned once per batch:
`ObjSeqPlan`), so padding, key masks and the per-sequence Python loops disappear; single-box sequences
(`indices[0]`) are ordinary length-1 segments.  Projections run on the tcgen05 GEMM; head_dim 297 is odd, so
the per-step bf16 weight copies carry each head zero-padded to 304 columns (q, k, v rows of in_proj, columns of
out_proj) — mathematically inert, keeps every head 16-byte aligned.  BatchNorm statistics are per video (the
reference's batch is one video).  No CPU / eager fallback.
"""
import numpy as np
import torch
import torch.nn as nn

from . import ops

F32, BF16 = torch.float32, torch.bfloat16
OBJ_DIM, EMBED_DIM, POS_DIM = 2048, 200, 128
D_OBJ = OBJ_DIM + EMBED_DIM + POS_DIM     # 2376
OBJ_HEADS, OBJ_FFN = 8, 1024


# ================================================================================================
# host logic
# ================================================================================================
def get_sequence(entry, gt_annotation=None, shape=None, task="sgcls"):
    """Mirror of tools/utils/ds_track.py:18-39 (same name and arguments): entry['indices'] = [rows of all
    single-box classes, rows of class a, rows of class b, ...] by the detector's arg-max class, ascending
    class id.  One device->host read instead of the reference's sync per class."""
    if task == "predcls":
        return
    dist = entry["distribution"]
    pred = torch.argmax(dist, 1).cpu().numpy()
    order = np.argsort(pred, kind="stable")
    cls, start, count = np.unique(pred[order], return_index=True, return_counts=True)
    singles, seqs = [], []
    for s, c in zip(start, count):
        rows = order[s:s + c]
        (singles if c == 1 else seqs).append(rows)
    dev = dist.device
    first = torch.as_tensor(np.concatenate(singles), dtype=torch.int64, device=dev) if singles else torch.tensor([])
    entry["indices"] = [first] + [torch.as_tensor(r, dtype=torch.int64, device=dev) for r in seqs]


class ObjSeqPlan:
    """Integer artefacts of the class-sequence encoder (all int32):
       seq_src[O]  box row feeding sequence row r (sequences of indices[1:] first, then the singles of indices[0])
       seq_off[S+1] segment offsets;  pos[O] frame rank inside the sequence (lib/tempura.py:191-195);
       inv[O]      sequence row that holds box b (the scatter of :202-203,207 as a gather)."""

    def __init__(self, indices, box_frames):
        bf = np.asarray(box_frames)
        # ONE device->host read for all class sequences (a batch of 64 videos holds ~850 of them: one `.cpu()` each was
        # ~850 stream synchronisations per forward); lengths come from the shapes, which live on the host
        seq_lens = [int(ix.shape[0]) for ix in indices[1:]]
        parts = [ix for ix in indices[1:]] + [indices[0]]
        parts = [t for t in parts if int(t.shape[0]) > 0]
        if parts and all(torch.is_tensor(t) for t in parts):
            same = all(t.dtype == parts[0].dtype and t.device == parts[0].device and t.dim() == 1 for t in parts)
            if not same:
                dev0 = next((t.device for t in parts if t.is_cuda), torch.device("cpu"))
                parts = [t.reshape(-1).long().to(dev0) for t in parts]
            src = torch.cat(parts).cpu().numpy().astype(np.int64)
        elif parts:
            src = np.concatenate([np.asarray(t.cpu() if torch.is_tensor(t) else t, dtype=np.int64).reshape(-1)
                                  for t in parts])
        else:
            src = np.zeros(0, np.int64)
        O = bf.shape[0]
        n_seq_rows = int(sum(seq_lens))
        n_single = src.shape[0] - n_seq_rows
        lens = seq_lens + [1] * n_single
        assert src.shape[0] == O and np.array_equal(np.sort(src), np.arange(O)), \
            "entry['indices'] must partition the boxes (tools/utils/ds_track.py:25-37)"
        # position of a sequence row (lib/tempura.py:191-195) — reference: unique(sorted) counts of the sequence's frame
        # ids, then rank k repeated count_k times, assigned in sequence order (rows are frame-sorted because boxes are).
        # All sequences at once: sort the frame ids inside every segment, dense-rank them, keep the sequence order.
        pos = np.zeros(O, dtype=np.int64)
        if n_seq_rows:
            seg = np.repeat(np.arange(len(seq_lens)), seq_lens)
            f = bf[src[:n_seq_rows]]
            order = np.lexsort((f, seg))
            fs = f[order]
            new = np.ones(n_seq_rows, dtype=np.int64)
            new[1:] = (fs[1:] != fs[:-1]) | (seg[1:] != seg[:-1])
            rank = np.cumsum(new) - 1
            seg_start = np.concatenate([[0], np.cumsum(seq_lens)[:-1]])
            pos[:n_seq_rows] = rank - np.repeat(rank[seg_start], seq_lens)
        off = np.zeros(len(lens) + 1, dtype=np.int64)
        off[1:] = np.cumsum(lens)
        inv = np.empty(O, dtype=np.int64)
        inv[src] = np.arange(O)
        self.O, self.S = O, len(lens)
        self.max_len = int(max(lens)) if lens else 0
        self.max_pos = int(pos.max()) if O else 0
        self.seq_src_h, self.seq_off_h, self.pos_h, self.inv_h = (a.astype(np.int32) for a in (src, off, pos, inv))

    def to(self, device):
        for n in ("seq_src", "seq_off", "pos", "inv"):
            setattr(self, n, ops.upload(getattr(self, n + "_h"), device))
        return self


class BoxGroups:
    """Per-video grouping of the box rows (BatchNorm statistics are per video)."""

    def __init__(self, box_frames, frames_per_video, device):
        bf = np.asarray(box_frames).astype(np.int64)
        fpv = np.asarray(frames_per_video, dtype=np.int64)
        video_of_frame = np.repeat(np.arange(fpv.shape[0]), fpv)
        vob = video_of_frame[bf]
        assert (np.diff(vob) >= 0).all(), "boxes must be ordered by frame"
        self.V = int(fpv.shape[0])
        self.count_h = np.bincount(vob, minlength=self.V).astype(np.int64)
        assert (self.count_h >= 2).all(), "BatchNorm1d in train mode needs at least two boxes per video"
        self.video_of_box = ops.upload(vob.astype(np.int32), device)
        self.video_of_box64 = ops.upload(vob, device)
        self.count = ops.upload(self.count_h.astype(np.float32), device)
        starts = np.concatenate([[0], np.cumsum(self.count_h)])
        rows = []
        for v in range(self.V):
            s = np.arange(starts[v], starts[v + 1], 1024, dtype=np.int64)
            rows.append(np.stack([s, np.minimum(s + 1024, starts[v + 1]), np.full_like(s, v)], 1))
        self.chunks = ops.upload(np.concatenate(rows).astype(np.int32), device)


def bn_stats(mean, var, cnt, bn, training):
    """(mean, rstd) [V,C] used by a BatchNorm layer.  Train mode: the per-video batch statistics passed in (biased
    variance), and the V sequential running-statistics updates the reference would have made video by video
    (momentum bn.momentum, unbiased variance).  Eval mode: the running statistics."""
    if not training:
        V = cnt.shape[0]
        return (bn.running_mean[None].expand(V, -1).contiguous(),
                torch.rsqrt(bn.running_var + bn.eps)[None].expand(V, -1).contiguous())
    rstd = torch.rsqrt(var + bn.eps)
    with torch.no_grad():
        V, m = mean.shape[0], bn.momentum
        w = m * (1 - m) ** torch.arange(V - 1, -1, -1, device=mean.device, dtype=mean.dtype)
        unb = var * (cnt / (cnt - 1))[:, None]
        bn.running_mean.mul_((1 - m) ** V).add_((w[:, None] * mean).sum(0))
        bn.running_var.mul_((1 - m) ** V).add_((w[:, None] * unb).sum(0))
        bn.num_batches_tracked += V
    return mean.contiguous(), rstd.contiguous()


def sinusoid_table(d_model, max_len):
    """PositionalEncoding.pe (lib/tempura.py:31-36), [1, max_len, d_model]."""
    import math
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(1, max_len, d_model)
    pe[0, :, 0::2] = torch.sin(position * div_term)
    pe[0, :, 1::2] = torch.cos(position * div_term)
    return pe


# ================================================================================================
# autograd blocks (forward AND backward are launch sequences over libb200vsgg)
# ================================================================================================
def _new(r, c, dt, dev):
    return torch.empty(r, c, device=dev, dtype=dt)


def _colsum(x):
    out = torch.zeros(1, x.shape[1], device=x.device)
    ops.colsum(x, out)
    return out[0]


def _bf(w):
    return ops.cast_bf16(w.detach().reshape(w.shape[0], -1).contiguous())


class ObjTokens(torch.autograd.Function):
    """b200vsgg_obj_tokens_{fwd,bwd}: x0 assembled directly in sequence order (+ position term, dropout)."""

    @staticmethod
    def forward(ctx, embed, bn_g, bn_b, wp, bp, args):
        a = dict(args, embed=embed.detach().contiguous(), bn_gamma=bn_g.detach().contiguous(),
                 bn_beta=bn_b.detach().contiguous(), wp=wp.detach().contiguous(), bp=bp.detach().contiguous())
        dev = a["features"].device
        x32, xb = _new(a["rows"], D_OBJ, F32, dev), _new(a["rows"], D_OBJ, BF16, dev)
        ops.obj_tokens_fwd(a, x32, xb)
        ctx.args = a
        ctx.mark_non_differentiable(xb)
        return x32, xb

    @staticmethod
    def backward(ctx, dx, _dxb):
        a = ctx.args
        dev = dx.device
        z = lambda *s: torch.zeros(*s, device=dev)
        dembed, dwp, dbp, dg, db = z(a["embed"].shape), z(a["wp"].shape), z(a["bp"].shape), z(4), z(4)
        ops.obj_tokens_bwd(a, dx.contiguous(), dembed, dwp, dbp, dg, db)
        return dembed, dg, db, dwp, dbp, None


def _pad_heads_rows(w, n_heads, hd, hdp, groups):
    """[groups*n_heads*hd, ...] -> [groups*n_heads*hdp, ...] with zero rows after every head."""
    if hd == hdp:
        return w
    tail = w.shape[1:]
    w = w.reshape(groups * n_heads, hd, *tail)
    out = w.new_zeros(groups * n_heads, hdp, *tail)
    out[:, :hd] = w
    return out.reshape(groups * n_heads * hdp, *tail)


def _unpad_heads_rows(w, n_heads, hd, hdp, groups):
    if hd == hdp:
        return w
    tail = w.shape[1:]
    return w.reshape(groups * n_heads, hdp, *tail)[:, :hd].reshape(groups * n_heads * hd, *tail)


class PostLNEncoderLayer(torch.autograd.Function):
    """nn.TransformerEncoderLayer(d, heads, ffn, dropout=p, batch_first, post-norm, ReLU) over unpadded row
    segments: y = LN2(t + drop(W2 drop(relu(W1 t)))), t = LN1(x + drop(Wo MHA(x)))  (lib/tempura.py:90-92,201)."""

    @staticmethod
    def forward(ctx, x32, xb, in_w, in_b, out_w, out_b, w1, b1, w2, b2, g1, be1, g2, be2, seg_off, n_seg, max_len,
                n_heads, p, seed):
        M, d = x32.shape
        dev = x32.device
        hd = d // n_heads
        hdp = (hd + 7) // 8 * 8
        dp = n_heads * hdp
        ffn = w1.shape[0]
        scale = float(hd) ** -0.5
        wqkv = _bf(_pad_heads_rows(in_w.detach(), n_heads, hd, hdp, 3))
        bqkv = _pad_heads_rows(in_b.detach(), n_heads, hd, hdp, 3).contiguous()
        wo = _bf(_pad_heads_rows(out_w.detach().t(), n_heads, hd, hdp, 1).t())
        w1b, w2b = _bf(w1), _bf(w2)
        qkv = _new(M, 3 * dp, BF16, dev)
        ops.gemm(xb, wqkv, bias=bqkv, out_bf16=qkv)
        ctxb = _new(M, dp, BF16, dev)
        lse = torch.empty(M, n_heads, device=dev)
        ops.attn_rows_fwd(qkv[:, :dp], qkv[:, dp:2 * dp], qkv[:, 2 * dp:], seg_off, n_seg, n_heads, hdp, ctxb, lse, p, seed,
                          scale=scale)
        u = _new(M, d, F32, dev)
        ops.gemm(ctxb, wo, bias=out_b.detach(), residual=x32, out_f32=u, dropout_p=p, seed=seed + 1)
        t32, tb = _new(M, d, F32, dev), _new(M, d, BF16, dev)
        m1, r1 = torch.empty(M, device=dev), torch.empty(M, device=dev)
        ops.layernorm_fwd(u, g1.detach(), be1.detach(), 1e-5, t32, tb, mean=m1, rstd=r1)
        h = _new(M, ffn, BF16, dev)
        ops.gemm(tb, w1b, bias=b1.detach(), act=ops.ACT_RELU, out_bf16=h, dropout_p=p, seed=seed + 2)
        v = _new(M, d, F32, dev)
        ops.gemm(h, w2b, bias=b2.detach(), residual=t32, out_f32=v, dropout_p=p, seed=seed + 3)
        y32, yb = _new(M, d, F32, dev), _new(M, d, BF16, dev)
        m2, r2 = torch.empty(M, device=dev), torch.empty(M, device=dev)
        ops.layernorm_fwd(v, g2.detach(), be2.detach(), 1e-5, y32, yb, mean=m2, rstd=r2)
        ctx.save_for_backward(xb, qkv, ctxb, lse, u, m1, r1, tb, h, v, m2, r2, wqkv, wo, w1b, w2b, g1, g2, seg_off)
        ctx.meta = (n_seg, max_len, n_heads, hd, hdp, p, seed, scale)
        ctx.mark_non_differentiable(yb)
        return y32, yb

    @staticmethod
    def backward(ctx, dy, _dyb):
        xb, qkv, ctxb, lse, u, m1, r1, tb, h, v, m2, r2, wqkv, wo, w1b, w2b, g1, g2, seg_off = ctx.saved_tensors
        n_seg, max_len, n_heads, hd, hdp, p, seed, scale = ctx.meta
        M, d = u.shape
        dev = u.device
        dp = n_heads * hdp
        ffn = h.shape[1]
        z = lambda *s: torch.zeros(*s, device=dev)
        # ---- norm2
        dv = _new(M, d, F32, dev)
        dg2, dbe2 = z(d), z(d)
        ops.layernorm_bwd(dy.contiguous(), v, g2.detach(), m2, r2, dv, None, 0.0, 0, dg2, dbe2)
        # ---- feed-forward: v = t + drop(W2 drop(relu(W1 t + b1)) + b2)
        dvb = ops.cast_bf16(dv, drop_p=p, seed=seed + 3)
        dw2 = _new(d, ffn, F32, dev)
        ops.gemm(dvb, h, a_mn=True, b_mn=True, out_f32=dw2)
        db2 = _colsum(dvb)
        dz = _new(M, ffn, BF16, dev)
        ops.gemm(dvb, w2b, b_mn=True, mask_src=h, mask_mode=ops.MASK_RELU, alpha=(1.0 / (1.0 - p)) if p > 0 else 1.0,
                 out_bf16=dz)
        dw1 = _new(ffn, d, F32, dev)
        ops.gemm(dz, tb, a_mn=True, b_mn=True, out_f32=dw1)
        db1 = _colsum(dz)
        dt = _new(M, d, F32, dev)
        ops.gemm(dz, w1b, b_mn=True, residual=dv, out_f32=dt)
        # ---- norm1 (its bf16 output carries the out_proj dropout mask for the next GEMMs)
        du, dub = _new(M, d, F32, dev), _new(M, d, BF16, dev)
        dg1, dbe1 = z(d), z(d)
        ops.layernorm_bwd(dt, u, g1.detach(), m1, r1, du, dub, p, seed + 1, dg1, dbe1)
        # ---- attention block: u = x + drop(Wo ctx + bo)
        dwo_p = _new(d, dp, F32, dev)
        ops.gemm(dub, ctxb, a_mn=True, b_mn=True, out_f32=dwo_p)
        dbo = _colsum(dub)
        dctx = _new(M, dp, BF16, dev)
        ops.gemm(dub, wo, b_mn=True, out_bf16=dctx)
        dqkv = _new(M, 3 * dp, BF16, dev)
        ops.attn_rows_bwd(qkv[:, :dp], qkv[:, dp:2 * dp], qkv[:, 2 * dp:], ctxb, dctx, lse, seg_off, n_seg, n_heads, hdp,
                          dqkv[:, :dp], dqkv[:, dp:2 * dp], dqkv[:, 2 * dp:], p, seed, scale=scale)
        dwqkv_p = _new(3 * dp, d, F32, dev)
        ops.gemm(dqkv, xb, a_mn=True, b_mn=True, out_f32=dwqkv_p)
        dbqkv_p = _colsum(dqkv)
        dx = _new(M, d, F32, dev)
        ops.gemm(dqkv, wqkv, b_mn=True, residual=du, out_f32=dx)
        din_w = _unpad_heads_rows(dwqkv_p, n_heads, hd, hdp, 3)
        din_b = _unpad_heads_rows(dbqkv_p, n_heads, hd, hdp, 3)
        dout_w = _unpad_heads_rows(dwo_p.t(), n_heads, hd, hdp, 1).t()
        return (dx, None, din_w, din_b, dout_w, dbo, dw1, db1, dw2, db2, dg1, dbe1, dg2, dbe2, None, None, None, None,
                None, None)


class GatherRows(torch.autograd.Function):
    """out[r] = x[idx[r]] for a permutation idx with inverse inv (the scatter-back of lib/tempura.py:202-207)."""

    @staticmethod
    def forward(ctx, x32, idx, inv):
        out32, outb = torch.empty_like(x32), torch.empty(x32.shape, device=x32.device, dtype=BF16)
        ops.gather_rows(x32, idx, out_f32=out32, out_bf16=outb)
        ctx.save_for_backward(inv)
        ctx.mark_non_differentiable(outb)
        return out32, outb

    @staticmethod
    def backward(ctx, d, _db):
        (inv,) = ctx.saved_tensors
        dx = torch.empty_like(d)
        ops.gather_rows(d.contiguous(), inv, out_f32=dx)
        return dx, None, None


class LinearBNReLU(torch.autograd.Function):
    """`intermediate` = Linear -> BatchNorm1d -> ReLU (lib/tempura.py:103-105) with per-video batch statistics."""

    @staticmethod
    def forward(ctx, x32, xb, w, b, gamma, beta, bn, groups, training):
        dev = xb.device
        O, n_out = xb.shape[0], w.shape[0]
        wb = _bf(w)
        zf, zb = _new(O, n_out, F32, dev), _new(O, n_out, BF16, dev)
        ops.gemm(xb, wb, bias=b.detach(), out_f32=zf, out_bf16=zb)
        mean = var = None
        if training:
            s1, s2 = torch.zeros(groups.V, n_out, device=dev), torch.zeros(groups.V, n_out, device=dev)
            ops.seg_colstats(zf, groups.chunks, s1, zf, s2)
            mean = s1 / groups.count[:, None]
            var = (s2 / groups.count[:, None] - mean * mean).clamp_min_(0.0)
        mean, rstd = bn_stats(mean, var, groups.count, bn, training)
        scale = (gamma.detach()[None] * rstd).contiguous()
        shift = (beta.detach()[None] - mean * scale).contiguous()
        y = _new(O, n_out, BF16, dev)
        ops.seg_affine(None, zb, None, scale, shift, groups.video_of_box, 1, y, relu_mask=2)
        ctx.save_for_backward(xb, wb, zb, y, mean, rstd, gamma)
        ctx.meta = (groups, training, x32 is not None and x32.requires_grad)
        return y

    @staticmethod
    def backward(ctx, dy):
        xb, wb, zb, y, mean, rstd, gamma = ctx.saved_tensors
        groups, training, need_dx = ctx.meta
        dev = xb.device
        O, n_out = zb.shape
        V = groups.V
        ones, zeros = torch.ones(V, n_out, device=dev), torch.zeros(V, n_out, device=dev)
        dyb = dy.contiguous() if dy.dtype == BF16 else ops.cast_bf16(dy.contiguous())
        dym = _new(O, n_out, BF16, dev)                      # dy where the ReLU let the value through
        ops.seg_affine(dyb, y, ones, zeros, zeros, groups.video_of_box, 1, dym, relu_mask=1)
        s1, s2 = torch.zeros(V, n_out, device=dev), torch.zeros(V, n_out, device=dev)
        ops.seg_colstats(dym, groups.chunks, s1, zb, s2)
        sx = (s2 - mean * s1) * rstd                          # sum dy * xhat per (video, channel)
        dgamma, dbeta = sx.sum(0), s1.sum(0)
        g = gamma.detach()[None]
        k1 = (g * rstd).expand(V, n_out).contiguous()
        if training:
            inv_n = (1.0 / groups.count)[:, None]
            k2 = (-g * rstd * rstd * sx * inv_n).contiguous()
            k3 = (-g * rstd * s1 * inv_n - k2 * mean).contiguous()
        else:
            k2, k3 = zeros, zeros
        dz = _new(O, n_out, BF16, dev)
        ops.seg_affine(dym, zb, k1, k2, k3, groups.video_of_box, 1, dz)
        dw = _new(n_out, xb.shape[1], F32, dev)
        ops.gemm(dz, xb, a_mn=True, b_mn=True, out_f32=dw)
        db = _colsum(dz)
        dx = None
        if need_dx:
            dx = _new(O, xb.shape[1], F32, dev)
            ops.gemm(dz, wb, b_mn=True, out_f32=dx)
        return dx, None, dw, db, dgamma, dbeta, None, None, None


class LinearHead(torch.autograd.Function):
    """decoder_lin = Linear(1024 -> 37) for obj_head='linear' (output width padded to a multiple of 8)."""

    @staticmethod
    def forward(ctx, xb, w, b):
        dev = xb.device
        n_out = w.shape[0]
        n_pad = (n_out + 7) // 8 * 8
        wb = torch.zeros(n_pad, w.shape[1], device=dev, dtype=BF16)
        ops.cast_bf16(w.detach().contiguous(), out=wb[:n_out])
        bias = torch.zeros(n_pad, device=dev)
        bias[:n_out] = b.detach()
        out = _new(xb.shape[0], n_pad, F32, dev)
        ops.gemm(xb, wb, bias=bias, out_f32=out)
        ctx.save_for_backward(xb, wb)
        ctx.n_out = n_out
        return out[:, :n_out]

    @staticmethod
    def backward(ctx, d):
        xb, wb = ctx.saved_tensors
        n_out, n_pad = ctx.n_out, wb.shape[0]
        dev = xb.device
        dp = torch.zeros(xb.shape[0], n_pad, device=dev)
        dp[:, :n_out] = d
        db_ = ops.cast_bf16(dp)
        dw = _new(n_pad, xb.shape[1], F32, dev)
        ops.gemm(db_, xb, a_mn=True, b_mn=True, out_f32=dw)
        dx = _new(xb.shape[0], xb.shape[1], BF16, dev)
        ops.gemm(db_, wb, b_mn=True, out_bf16=dx)
        return dx, dw[:n_out], _colsum(db_)[:n_out]


# ================================================================================================
# the branch
# ================================================================================================
def box_center_size(boxes):
    """center_size of tools/utils/fpn/box_utils.py (absent from the reference tree; neural-motifs definition)."""
    wh = boxes[:, 3:5] - boxes[:, 1:3] + 1.0
    return torch.cat((boxes[:, 1:3] + 0.5 * wh, wh), 1)


def run_object_branch(oc, entry, phase, frames_per_video, heads_fn, dropout_p, gmm_eps=None, unc=False):
    """ObjectClassifier.forward for mode='sgcls', phase='train' (lib/tempura.py:249-255 -> classify :185-241).
    `oc` is the parameter container (tempura.ObjectClassifier); returns entry with `distribution`,
    `object_features`, `object_mem_features`, `pred_labels`."""
    feats = entry["features"]
    if not feats.is_cuda:
        raise RuntimeError("b200vsgg object branch runs only on CUDA tensors (no CPU fallback)")
    dev = feats.device
    training = oc.training
    p = dropout_p if training else 0.0
    box_frames = entry.get("box_frames_host")
    if box_frames is None:
        box_frames = entry["boxes"][:, 0].cpu().numpy()
    groups = BoxGroups(box_frames, frames_per_video, dev)
    O = feats.shape[0]
    seed = int(torch.randint(0, 2 ** 40, (1,)).item()) if p > 0 else 0

    # BatchNorm1d(4) of pos_embed: per-video statistics of center_size(box) — [O,4] values, torch ops
    bn4 = oc.pos_embed[0]
    boxes = entry["boxes"].contiguous().float()
    mean4 = var4 = None
    if training:   # centred second moment: box coordinates are O(100) with O(10) spread
        cs = box_center_size(boxes)
        vob = groups.video_of_box64
        mean4 = torch.zeros(groups.V, 4, device=dev).index_add_(0, vob, cs) / groups.count[:, None]
        var4 = torch.zeros(groups.V, 4, device=dev).index_add_(0, vob, (cs - mean4[vob]) ** 2) / groups.count[:, None]
    mean4, rstd4 = bn_stats(mean4, var4, groups.count, bn4, training)

    args = dict(features=feats.contiguous(), dist=entry["distribution"].contiguous().float(), boxes=boxes,
                bn_mean=mean4, bn_rstd=rstd4, video_of_box=groups.video_of_box, rows=O, p_pos=p, seed_pos=seed + 11)
    lin_pos = oc.pos_embed[1]
    if oc.tracking:
        plan = ObjSeqPlan(entry["indices"], box_frames).to(dev)
        pe = oc.positional_encoder.pe[0]
        if plan.max_pos >= pe.shape[0]:
            raise RuntimeError("sequence spans %d frames, positional table holds %d" % (plan.max_pos + 1, pe.shape[0]))
        args.update(pe=pe.contiguous(), src=plan.seq_src, pos=plan.pos, p_pe=p, seed_pe=seed + 12)
        x32, xb = ObjTokens.apply(oc.obj_embed.weight, bn4.weight, bn4.bias, lin_pos.weight, lin_pos.bias, args)
        for i, L in enumerate(oc.encoder_tran.layers):
            x32, xb = PostLNEncoderLayer.apply(
                x32, xb, L.self_attn.in_proj_weight, L.self_attn.in_proj_bias, L.self_attn.out_proj.weight,
                L.self_attn.out_proj.bias, L.linear1.weight, L.linear1.bias, L.linear2.weight, L.linear2.bias,
                L.norm1.weight, L.norm1.bias, L.norm2.weight, L.norm2.bias, plan.seq_off, plan.S, plan.max_len,
                OBJ_HEADS, p, seed + 100 * (i + 1))
        f32, fb = GatherRows.apply(x32, plan.inv, plan.seq_src)
        entry["object_features"] = f32
        if oc.mem_compute and len(oc.obj_memory) != 0:
            f32 = oc.hallucinate(f32)
            fb = ops.cast_bf16(f32.contiguous()) if not f32.requires_grad else _CastBF16.apply(f32)
        entry["object_mem_features"] = f32
        inter = oc.intermediate
        y = LinearBNReLU.apply(f32, fb, inter[0].weight, inter[0].bias, inter[1].weight, inter[1].bias, inter[1], groups,
                               training)
    else:
        x32, xb = ObjTokens.apply(oc.obj_embed.weight, bn4.weight, bn4.bias, lin_pos.weight, lin_pos.bias, args)
        inter = oc.intermediate
        y = LinearBNReLU.apply(x32, xb, inter[0].weight, inter[0].bias, inter[1].weight, inter[1].bias, inter[1], groups,
                               training)
        entry["object_features"] = y.float()
        if oc.mem_compute and len(oc.obj_memory) != 0:
            # lib/tempura.py:217-221: without tracking the memory hallucinator works on the 1024-wide `intermediate`
            # output (single-head attention over the class memory through the C-ABI GEMM, blended by the selector)
            y = oc.hallucinate(y.float())
        entry["object_mem_features"] = y.float()

    if getattr(oc, "_debug", False):      # parity debugging: expose the branch's intermediate tensors
        oc._debug_last = dict(y=y, tokens=x32)
    if oc.obj_head == "gmm" and unc:
        # lib/tempura.py:226-228 (the trainer's uncertainty pass, Uncertainty.py:100): test-phase distribution (mixture of
        # the means, background class dropped) + aleatoric / epistemic uncertainties of the object head
        with torch.no_grad():
            (dist,) = heads_fn([oc.decoder_lin], y.float(), 0, [None], 0, skip_first=True)
            al, ep = heads_fn([oc.decoder_lin], y.float(), 2, [None], 0)
        entry["distribution"], entry["obj_al_uc"], entry["obj_ep_uc"] = dist, al, ep
    elif oc.obj_head == "gmm":
        eps = [gmm_eps.get("object")] if gmm_eps else [None]
        hseed = int(torch.randint(0, 2 ** 62, (1,)).item())
        (dist,) = heads_fn([oc.decoder_lin], y.float(), 1, eps, hseed)
        entry["distribution"] = dist
    else:
        lin = oc.decoder_lin[0]
        entry["distribution"] = LinearHead.apply(y, lin.weight, lin.bias)
    entry["pred_labels"] = entry["labels"]
    entry["box_groups"] = groups          # per-video box grouping, read by tempura_loss / object_loss
    return entry


class _CastBF16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return ops.cast_bf16(x.contiguous())

    @staticmethod
    def backward(ctx, d):
        return d.float()


def object_loss(pred, eos_coef=1.0, groups_count=None, video_of_box=None):
    """TEMPURA_train.py:97-100,191-195: class-weighted CE (weight[0] = eos_coef, reduction='none') on
    `distribution` as the model returns it, then the mean over boxes.  With a batch of videos: the mean over
    videos of the per-video mean (= the average of the reference's per-video losses)."""
    import torch.nn.functional as F
    dist = pred["distribution"]
    w = torch.ones(dist.shape[1], device=dist.device)
    w[0] = eos_coef
    ce = F.cross_entropy(dist, pred["labels"], weight=w, reduction="none")
    if groups_count is None:
        return ce.mean()
    V = groups_count.shape[0]
    return (ce / (groups_count[video_of_box] * V)).sum()
