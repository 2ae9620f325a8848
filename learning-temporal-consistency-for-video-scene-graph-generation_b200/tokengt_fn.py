"""Block-level autograd functions of the TEAT-GT / TokenGT path.  Each function is a hand-orchestrated
launch sequence over libb200vsgg kernels for forward AND backward (tcgen05 GEMMs with fused
epilogues, flash attention, LayerNorm, ragged gathers); torch only carries the graph between blocks.

Residual stream fp32 [T, d]; GEMM operands bf16; dropout masks are regenerated from (seed, index).
Reference lines: tokengt_graph_encoder_layer.py:170-191 (pre-LN layer), multihead_attention.py:135-183,
feedforward.py:31-36, tokenizer.py:217-295, models/tokengt.py:108-117, lib/teatgt.py:118-141.
"""
import torch

from . import ops

F32, BF16 = torch.float32, torch.bfloat16


def _bf(w):
    return ops.cast_bf16(w.detach().reshape(w.shape[0], -1).contiguous())


def _bfc(tag, *ws):
    """bf16 copy of one weight (or of several stacked along rows), memoised per parameter version."""
    return ops.cached_weight(tag, ws, lambda: _bf(ws[0] if len(ws) == 1 else torch.cat(ws, 0)))


def _new(rows, cols, dtype, dev):
    return torch.empty(rows, cols, device=dev, dtype=dtype)


def _colsum(x):
    out = torch.zeros(1, x.shape[1], device=x.device)
    ops.colsum(x, out)
    return out[0]


class AttnPlan:
    """Device copies of the varlen attention plan: sequence offsets, the 64-row block table of the mma.sync backward
    kernels and the 128-row block table of the tcgen05 forward (csrc/attn_tc.cu)."""

    def __init__(self, seq_off_h, device):
        import numpy as np
        from .plan import attention_blocks
        bs, br = attention_blocks(seq_off_h)
        self.seq_off = ops.upload(np.asarray(seq_off_h, dtype=np.int32), device)
        self.blk_seq = ops.upload(bs, device)
        self.blk_row0 = ops.upload(br, device)
        lens = np.diff(np.asarray(seq_off_h, dtype=np.float64))
        self.sum_t2 = float((lens * lens).sum())       # algorithmic attention flops = 4 * sum_t2 * heads * head_dim
        bs128, br128 = attention_blocks(seq_off_h, block=128)
        self.blk_seq128 = ops.upload(bs128, device)
        self.blk_row0_128 = ops.upload(br128, device)
        # 0 = tiled kernels.  The shared-memory-resident variants (pass the longest sequence length) are correct
        # but measured slower at the C3 shapes (occupancy: 1-2 CTAs/SM), see profiles/r01_flash_attention.txt
        self.max_len = 0


class PreLNAttention(torch.autograd.Function):
    """x -> x + dropout(out_proj(MHA(LN(x))))."""

    @staticmethod
    def forward(ctx, x, g, b, wq, bq, wk, bk, wv, bv, wo, bo, plan, n_heads, p_attn, p_out, seed):
        T, d = x.shape
        dev = x.device
        hd = d // n_heads
        h = _new(T, d, BF16, dev)
        mean, rstd = torch.empty(T, device=dev), torch.empty(T, device=dev)
        ops.layernorm_fwd(x, g.detach(), b.detach(), 1e-5, None, h, mean=mean, rstd=rstd)
        wqkv = _bfc("qkv", wq, wk, wv)
        bqkv = ops.cached_weight("bqkv", (bq, bk, bv), lambda: torch.cat([bq, bk, bv]).detach().clone())
        qkv = _new(T, 3 * d, BF16, dev)
        ops.gemm(h, wqkv, bias=bqkv, out_bf16=qkv)
        att = _new(T, d, BF16, dev)
        lse = torch.empty(T, n_heads, device=dev)
        # tcgen05 / TMEM / TMA forward; it writes the lse and uses the dropout mask function the backward kernels expect
        ops.attn_tc_fwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], plan.seq_off, plan.blk_seq128, plan.blk_row0_128,
                        n_heads, hd, att, lse, p_attn, seed, flops=4.0 * plan.sum_t2 * d)
        wob = _bfc("w", wo)
        y = _new(T, d, F32, dev)
        ops.gemm(att, wob, bias=bo.detach(), residual=x, out_f32=y, dropout_p=p_out, seed=seed + 1)
        ctx.save_for_backward(x, g, mean, rstd, h, qkv, att, lse, wqkv, wob)
        ctx.meta = (plan, n_heads, p_attn, p_out, seed)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, g, mean, rstd, h, qkv, att, lse, wqkv, wob = ctx.saved_tensors
        plan, n_heads, p_attn, p_out, seed = ctx.meta
        T, d = x.shape
        dev = x.device
        hd = d // n_heads
        dy = dy.contiguous()
        dyb = ops.cast_bf16(dy, drop_p=p_out, seed=seed + 1)
        dwo = _new(d, d, F32, dev)
        ops.gemm(dyb, att, a_mn=True, b_mn=True, out_f32=dwo)
        dbo = _colsum(dyb)
        datt = _new(T, d, BF16, dev)
        ops.gemm(dyb, wob, b_mn=True, out_bf16=datt)
        dqkv = _new(T, 3 * d, BF16, dev)
        ops.attn_tc_bwd(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], att, datt, lse, plan.seq_off, plan.blk_seq128,
                        plan.blk_row0_128, n_heads, hd, dqkv[:, :d], dqkv[:, d:2 * d], dqkv[:, 2 * d:], p_attn, seed,
                        flops=10.0 * plan.sum_t2 * d)
        dwqkv = _new(3 * d, d, F32, dev)
        ops.gemm(dqkv, h, a_mn=True, b_mn=True, out_f32=dwqkv)
        dbqkv = _colsum(dqkv)
        dh = _new(T, d, F32, dev)
        ops.gemm(dqkv, wqkv, b_mn=True, out_f32=dh)
        dx = _new(T, d, F32, dev)
        dg, db = torch.zeros(d, device=dev), torch.zeros(d, device=dev)
        ops.layernorm_bwd(dh, x, g.detach(), mean, rstd, dx, None, 0.0, 0, dg, db, base=dy)
        return (dx, dg, db, dwqkv[:d], dbqkv[:d], dwqkv[d:2 * d], dbqkv[d:2 * d], dwqkv[2 * d:], dbqkv[2 * d:], dwo, dbo,
                None, None, None, None, None)


class PreLNFeedForward(torch.autograd.Function):
    """x -> x + dropout(fc2(dropout(gelu(fc1(LN(x))))))."""

    @staticmethod
    def forward(ctx, x, g, b, w1, b1, w2, b2, p_act, p_out, seed):
        T, d = x.shape
        dev = x.device
        ffn = w1.shape[0]
        h = _new(T, d, BF16, dev)
        mean, rstd = torch.empty(T, device=dev), torch.empty(T, device=dev)
        ops.layernorm_fwd(x, g.detach(), b.detach(), 1e-5, None, h, mean=mean, rstd=rstd)
        w1b, w2b = _bfc("w", w1), _bfc("w", w2)
        z = _new(T, ffn, BF16, dev)
        ops.gemm(h, w1b, bias=b1.detach(), out_bf16=z)
        a = ops.act_dropout(z, ops.ACT_GELU, p_act, seed)
        y = _new(T, d, F32, dev)
        ops.gemm(a, w2b, bias=b2.detach(), residual=x, out_f32=y, dropout_p=p_out, seed=seed + 1)
        ctx.save_for_backward(x, g, mean, rstd, h, z, a, w1b, w2b)
        ctx.meta = (p_act, p_out, seed)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, g, mean, rstd, h, z, a, w1b, w2b = ctx.saved_tensors
        p_act, p_out, seed = ctx.meta
        T, d = x.shape
        dev = x.device
        ffn = z.shape[1]
        dy = dy.contiguous()
        dyb = ops.cast_bf16(dy, drop_p=p_out, seed=seed + 1)
        dw2 = _new(d, ffn, F32, dev)
        ops.gemm(dyb, a, a_mn=True, b_mn=True, out_f32=dw2)
        db2 = _colsum(dyb)
        dz = _new(T, ffn, BF16, dev)
        ops.gemm(dyb, w2b, b_mn=True, mask_src=z, mask_mode=ops.MASK_GELU, out_bf16=dz, dropout_p=p_act, seed=seed)
        dw1 = _new(ffn, d, F32, dev)
        ops.gemm(dz, h, a_mn=True, b_mn=True, out_f32=dw1)
        db1 = _colsum(dz)
        dh = _new(T, d, F32, dev)
        ops.gemm(dz, w1b, b_mn=True, out_f32=dh)
        dx = _new(T, d, F32, dev)
        dg, db = torch.zeros(d, device=dev), torch.zeros(d, device=dev)
        ops.layernorm_bwd(dh, x, g.detach(), mean, rstd, dx, None, 0.0, 0, dg, db, base=dy)
        return dx, dg, db, dw1, db1, dw2, db2, None, None, None


class NodeTokens(torch.autograd.Function):
    """lib/teatgt.py:118-141: [subj_fc | obj_fc](features) gathered per node + label embedding -> [n, 1168]."""

    @staticmethod
    def forward(ctx, feat_b, w_s, b_s, w_o, b_o, embed, labels, feat_row, is_person):
        dev = feat_b.device
        h1 = w_s.shape[0]
        wso = _bfc("so", w_s, w_o)
        so = _new(feat_b.shape[0], 2 * h1, F32, dev)
        ops.gemm(feat_b, wso, bias=torch.cat([b_s, b_o]).detach(), out_f32=so)
        n = feat_row.numel()
        D = h1 + embed.shape[1]
        tok = _new(n, D, F32, dev)
        tokb = _new(n, D, BF16, dev)
        ops.node_tokens_fwd(so, feat_row, is_person, labels, embed.detach().contiguous(), h1, tok, tokb)
        ctx.save_for_backward(feat_b, labels, feat_row, is_person)
        ctx.meta = (h1, embed.shape)
        ctx.mark_non_differentiable(tokb)
        return tok, tokb

    @staticmethod
    def backward(ctx, dtok, _dtokb):
        feat_b, labels, feat_row, is_person = ctx.saved_tensors
        h1, eshape = ctx.meta
        dev = dtok.device
        dso = torch.zeros(feat_b.shape[0], 2 * h1, device=dev)
        dembed = torch.zeros(eshape, device=dev)
        ops.node_tokens_bwd(dtok.contiguous(), feat_row, is_person, labels, h1, eshape[1], dso, dembed)
        dsob = ops.cast_bf16(dso)
        dw = _new(2 * h1, feat_b.shape[1], F32, dev)
        ops.gemm(dsob, feat_b, a_mn=True, b_mn=True, out_f32=dw)
        dbias = _colsum(dsob)
        return None, dw[:h1], dbias[:h1], dw[h1:], dbias[h1:], dembed, None, None, None


class AssembleTokens(torch.autograd.Function):
    """tokenizer.py:217-295 for all clips at once -> x0 [T, d] fp32."""

    @staticmethod
    def forward(ctx, tok, tokb, evb, wa, ba, wl, temp, eemb, order, graph_tok, null_tok, desc, lap_k):
        dev = tok.device
        n, d = tokb.shape[0], wa.shape[0]
        wab = _bfc("w", wa)
        kp = evb.shape[1]                                    # eigenvector columns padded to a multiple of 8
        wlu = torch.zeros(d, kp, device=dev, dtype=BF16)
        wlv = torch.zeros(d, kp, device=dev, dtype=BF16)
        wlu[:, :lap_k] = wl.detach()[:, :lap_k]           # [d, 50] weight slices: a few kB, cast by torch
        wlv[:, :lap_k] = wl.detach()[:, lap_k:]
        na, pu, pv = _new(n, d, F32, dev), _new(n, d, F32, dev), _new(n, d, F32, dev)
        ops.gemm(tokb, wab, bias=ba.detach(), out_f32=na)
        ops.gemm(evb, wlu, out_f32=pu)
        ops.gemm(evb, wlv, out_f32=pv)
        x = _new(desc.shape[0], d, F32, dev)
        ops.teat_assemble_fwd(desc, na, pu, pv, temp.detach().contiguous(), eemb.detach().contiguous(),
                              order.detach().contiguous(), graph_tok.detach().contiguous(),
                              null_tok.detach().contiguous(), x)
        ctx.save_for_backward(tokb, evb, wab, desc)
        ctx.meta = (lap_k, temp.shape, eemb.shape, order.shape, wl.shape)
        return x

    @staticmethod
    def backward(ctx, dx):
        tokb, evb, wab, desc = ctx.saved_tensors
        lap_k, tshape, eshape, oshape, wlshape = ctx.meta
        dev = dx.device
        n, d = tokb.shape[0], wab.shape[0]
        z = lambda *s: torch.zeros(*s, device=dev)
        dna, dpu, dpv = z(n, d), z(n, d), z(n, d)
        dtemp, deemb, dorder, dgraph, dnull = z(tshape), z(eshape), z(oshape), z(1, d), z(1, d)
        ops.teat_assemble_bwd(desc, dx.contiguous(), dna, dpu, dpv, dtemp, deemb, dorder, dgraph, dnull)
        dtemp[0].zero_()                                     # nn.Embedding(padding_idx=0): row 0 never gets a gradient
        deemb[0].zero_()
        dnab, dpub, dpvb = ops.cast_bf16(dna), ops.cast_bf16(dpu), ops.cast_bf16(dpv)
        dwa = _new(d, tokb.shape[1], F32, dev)
        ops.gemm(dnab, tokb, a_mn=True, b_mn=True, out_f32=dwa)
        dba = _colsum(dnab)
        dtok = _new(n, tokb.shape[1], F32, dev)
        ops.gemm(dnab, wab, b_mn=True, out_f32=dtok)
        kp = evb.shape[1]
        dwu, dwv = _new(d, kp, F32, dev), _new(d, kp, F32, dev)
        ops.gemm(dpub, evb, a_mn=True, b_mn=True, out_f32=dwu)
        ops.gemm(dpvb, evb, a_mn=True, b_mn=True, out_f32=dwv)
        dwl = torch.cat([dwu[:, :lap_k], dwv[:, :lap_k]], 1)
        return dtok, None, None, dwa, dba, dwl, dtemp, deemb, dorder, dgraph, dnull, None, None


class NodeHead(torch.autograd.Function):
    """models/tokengt.py:108-117 on the node rows only: LN(GELU(Linear)) -> hidden; Linear(d -> 26) + bias."""

    @staticmethod
    def forward(ctx, x, node_rows, wt, bt, g, b, we, bias):
        dev = x.device
        n, d = node_rows.numel(), x.shape[1]
        xn = _new(n, d, BF16, dev)
        ops.gather_rows(x, node_rows, out_bf16=xn)
        wtb = _bfc("w", wt)
        z = _new(n, d, BF16, dev)
        ops.gemm(xn, wtb, bias=bt.detach(), out_bf16=z)
        a32 = ops.act_dropout(z, ops.ACT_GELU).float()
        hid = _new(n, d, F32, dev)
        hidb = _new(n, d, BF16, dev)
        mean, rstd = torch.empty(n, device=dev), torch.empty(n, device=dev)
        ops.layernorm_fwd(a32, g.detach(), b.detach(), 1e-5, hid, hidb, mean=mean, rstd=rstd)
        n_out = we.shape[0]
        n_pad = (n_out + 7) // 8 * 8
        web = torch.zeros(n_pad, d, device=dev, dtype=BF16)
        ops.cast_bf16(we.detach().contiguous(), out=web[:n_out])
        bpad = torch.zeros(n_pad, device=dev)
        bpad[:n_out] = bias.detach()
        logits = _new(n, n_pad, F32, dev)
        ops.gemm(hidb, web, bias=bpad, out_f32=logits)
        ctx.save_for_backward(node_rows, xn, wtb, z, a32, g, mean, rstd, hidb, web)
        ctx.meta = (x.shape, n_out)
        return logits[:, :n_out], hid

    @staticmethod
    def backward(ctx, dlogits, dhid):
        node_rows, xn, wtb, z, a32, g, mean, rstd, hidb, web = ctx.saved_tensors
        xshape, n_out = ctx.meta
        dev = xn.device
        n, d = xn.shape
        n_pad = web.shape[0]
        dl = torch.zeros(n, n_pad, device=dev)
        dl[:, :n_out] = dlogits
        dlb = ops.cast_bf16(dl)
        dwe = _new(n_pad, d, F32, dev)
        ops.gemm(dlb, hidb, a_mn=True, b_mn=True, out_f32=dwe)
        dbias = _colsum(dlb)[:n_out]
        dh = _new(n, d, F32, dev)
        ops.gemm(dlb, web, b_mn=True, out_f32=dh)
        if dhid is not None:
            dh = dh + dhid
        da = _new(n, d, F32, dev)
        dg, db = torch.zeros(d, device=dev), torch.zeros(d, device=dev)
        ops.layernorm_bwd(dh, a32, g.detach(), mean, rstd, da, None, 0.0, 0, dg, db)
        zf = z.float()                                       # [nodes, d] only: gelu'(z) on the host-side graph
        cdf = 0.5 * (1.0 + torch.erf(zf * 0.7071067811865476))
        dz = (da * (cdf + zf * torch.exp(-0.5 * zf * zf) * 0.3989422804014327)).to(BF16)
        dwt = _new(d, d, F32, dev)
        ops.gemm(dz, xn, a_mn=True, b_mn=True, out_f32=dwt)
        dbt = _colsum(dz)
        dxn = _new(n, d, F32, dev)
        ops.gemm(dz, wtb, b_mn=True, out_f32=dxn)
        dx = torch.zeros(xshape, device=dev)
        dx.index_copy_(0, node_rows.long(), dxn)
        return dx, None, dwt, dbt, dg, db, dwe[:n_out], dbias
