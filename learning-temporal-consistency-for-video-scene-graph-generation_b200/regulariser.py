"""Adjacent-frame graph temporal-consistency regulariser (rows R1-R3 of SURVEY.md §8a;
lib/teatgt.py:285-334, 350-351 of the reference), batched over every frame of every clip.

As released, the reference detaches both loss vectors (`torch.tensor(list_of_scalars)`,
lib/teatgt.py:350-351), so they carry no gradient; that is the default here too.  `differentiable=True`
(`model.differentiable_consistency`) evaluates the same networks layer by layer with saved
activations and a hand-orchestrated backward (second half of this file).

Per frame: spatial-only graph -> normalised Laplacian -> eigenvectors (host LAPACK, the reference's call
:300) -> first 10 columns -> GraphTransformer(dim 10) -> attention pooling -> structure embedding [10];
clip hidden rows [0:n_f] (the reference's `savor` never advances, :312-314) -> GraphTransformer(dim 768)
-> attention pooling -> semantic embedding [768].  All frame pairs of a clip: KL / frame distance
(b200vsgg_consistency_kl, warp-shuffle reduction), kept where >= 0.

PARITY UNPINNED: GraphTransformer / GlobalAttentionPooling come from `graph_transformer_pytorch` and `dgl`,
absent from the reference tree and unversioned there; the arithmetic restates their published algorithms
 and is checked against the oracle's restatement only.  The 768/1936-wide semantic branch runs
row-compacted on the tcgen05 GEMM + `graph_attn_core` / `gated_residual` kernels; the 10-wide structure branch
(4 layers + pooling) is ONE launch of `b200vsgg_graph_small_fwd` (one CTA per frame).  Both kernels hold a frame's
graph on chip: frames with more than MAX_NODES_STRUCT / MAX_NODES_SEM nodes (person + objects) RAISE — there is no
eager-PyTorch or CPU route around the kernels (Action Genome frames have <= 10 annotated objects; the regulariser is
train-only, so the 33-node long-clip inference config never reaches it).
"""
import numpy as np
import torch
import torch.nn as nn

from . import ops


class _Attention(nn.Module):
    def __init__(self, dim, dim_head, heads, edge_dim):
        super().__init__()
        inner = dim_head * heads
        self.heads, self.dim_head = heads, dim_head
        self.to_q = nn.Linear(dim, inner)
        self.to_kv = nn.Linear(dim, inner * 2)
        self.edges_to_kv = nn.Linear(edge_dim, inner)
        self.to_out = nn.Linear(inner, dim)


class _PreNorm(nn.Module):
    def __init__(self, dim, fn):
        super().__init__()
        self.fn, self.norm = fn, nn.LayerNorm(dim)


class _GatedResidual(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.proj = nn.Sequential(nn.Linear(dim * 3, 1, bias=False), nn.Sigmoid())


class GraphTransformer(nn.Module):
    """Parameters of graph_transformer_pytorch.GraphTransformer(dim, depth, edge_dim=1, with_feedforwards=True,
    gated_residual=True, rel_pos_emb=True) (defaults dim_head 64, heads 8); evaluated by `run_compact` (wide graphs)
    or b200vsgg_graph_small_fwd (the 10-wide structure branch)."""

    def __init__(self, dim, depth, dim_head=64, heads=8, edge_dim=1):
        super().__init__()
        self.dim, self.dim_head, self.heads = dim, dim_head, heads
        self.layers = nn.ModuleList()
        for _ in range(depth):
            self.layers.append(nn.ModuleList([
                nn.ModuleList([_PreNorm(dim, _Attention(dim, dim_head, heads, edge_dim)), _GatedResidual(dim)]),
                nn.ModuleList([_PreNorm(dim, nn.Sequential(nn.Linear(dim, dim * 4), nn.GELU(), nn.Linear(dim * 4, dim))),
                               _GatedResidual(dim)])]))


MAX_NODES_STRUCT = 16   # b200vsgg_graph_small_fwd: q/k/v of every node of a frame live in one warp's registers
MAX_NODES_SEM = 32      # b200vsgg_graph_attn_core: lane j of a warp keeps the score of key j


def pack_small_params(gt):
    """GraphTransformer parameters in the packed order b200vsgg_graph_small_fwd reads (one torch.cat per call)."""
    parts = []
    for attn_block, ff_block in gt.layers:
        pre, gate = attn_block
        a = pre.fn
        pre2, gate2 = ff_block
        parts += [pre.norm.weight, pre.norm.bias, a.to_q.weight, a.to_q.bias, a.to_kv.weight, a.to_kv.bias,
                  a.edges_to_kv.weight, a.edges_to_kv.bias, a.to_out.weight, a.to_out.bias, gate.proj[0].weight,
                  pre2.norm.weight, pre2.norm.bias, pre2.fn[0].weight, pre2.fn[0].bias, pre2.fn[2].weight,
                  pre2.fn[2].bias, gate2.proj[0].weight]
    return torch.cat([p.detach().reshape(-1).float() for p in parts])


@torch.no_grad()
def run_compact(gt, x, node_off, upper, nmax):
    """Row-compacted GraphTransformer forward on the C-ABI kernels: x fp32 [R, dim] (dim % 8 == 0) = node rows of
    all frames back to back, node_off int32 [frames+1], upper uint8 [frames, nmax, nmax] (adjacency = U + U^T).
    LayerNorm, every projection (bf16 tcgen05 GEMMs, GELU fused), the per-frame attention core and the gated
    residuals are library kernels; nothing of the padded [frames, nmax, ...] shape is materialised."""
    R, dim = x.shape
    dev = x.device
    inner = gt.heads * gt.dim_head
    bf = lambda w: ops.cached_weight("gt", (w,), lambda: ops.cast_bf16(w.detach().contiguous()))
    x = x.contiguous().clone()
    xn = torch.empty(R, dim, device=dev, dtype=torch.bfloat16)
    qkv = torch.empty(R, 3 * inner, device=dev)
    att = torch.empty(R, inner, device=dev, dtype=torch.bfloat16)
    o = torch.empty(R, dim, device=dev)
    hid = torch.empty(R, 4 * dim, device=dev, dtype=torch.bfloat16)
    for attn_block, ff_block in gt.layers:
        pre, gate = attn_block
        a = pre.fn
        ops.layernorm_fwd(x, pre.norm.weight.detach(), pre.norm.bias.detach(), 1e-5, None, xn)
        wqkv = ops.cached_weight("gt_qkv", (a.to_q.weight, a.to_kv.weight),
                                 lambda: ops.cast_bf16(torch.cat([a.to_q.weight, a.to_kv.weight], 0).detach().contiguous()))
        bqkv = ops.cached_weight("gt_bqkv", (a.to_q.bias, a.to_kv.bias),
                                 lambda: torch.cat([a.to_q.bias, a.to_kv.bias]).detach().clone())
        ops.gemm(xn, wqkv, bias=bqkv, out_f32=qkv)
        ops.graph_attn_core(qkv, node_off, upper, nmax, a.edges_to_kv.weight.detach()[:, 0].contiguous(),
                            a.edges_to_kv.bias.detach().contiguous(), att)
        ops.gemm(att, bf(a.to_out.weight), bias=a.to_out.bias.detach(), out_f32=o)
        ops.gated_residual(o, x, gate.proj[0].weight.detach().reshape(-1).contiguous())
        pre2, gate2 = ff_block
        ops.layernorm_fwd(x, pre2.norm.weight.detach(), pre2.norm.bias.detach(), 1e-5, None, xn)
        ops.gemm(xn, bf(pre2.fn[0].weight), bias=pre2.fn[0].bias.detach(), act=ops.ACT_GELU, out_bf16=hid)
        ops.gemm(hid, bf(pre2.fn[2].weight), bias=pre2.fn[2].bias.detach(), out_f32=o)
        ops.gated_residual(o, x, gate2.proj[0].weight.detach().reshape(-1).contiguous())
    return x


# ================================================================================================
# differentiable mode: the same launch sequence with saved activations and a hand-orchestrated backward
# ================================================================================================
_PER_LAYER = 18      # parameters per GraphTransformer layer, in the order of `_layer_params`


def _layer_params(gt):
    out = []
    for attn_block, ff_block in gt.layers:
        pre, gate = attn_block
        a = pre.fn
        pre2, gate2 = ff_block
        out += [pre.norm.weight, pre.norm.bias, a.to_q.weight, a.to_q.bias, a.to_kv.weight, a.to_kv.bias,
                a.edges_to_kv.weight, a.edges_to_kv.bias, a.to_out.weight, a.to_out.bias, gate.proj[0].weight,
                pre2.norm.weight, pre2.norm.bias, pre2.fn[0].weight, pre2.fn[0].bias, pre2.fn[2].weight, pre2.fn[2].bias,
                gate2.proj[0].weight]
    return out


class _GraphTransformerFn(torch.autograd.Function):
    """`run_compact` with gradients: forward = the same kernels (GELU applied by a separate pass so that the pre-activation
    is kept), backward = GatedResidual / GraphTransformer-attention backward kernels of csrc/consistency.cu, the LayerNorm
    backward and the tcgen05 GEMM in its dgrad / wgrad roles."""

    @staticmethod
    def forward(ctx, x, node_off, upper, nmax, heads, dim_head, *params):
        R, dim = x.shape
        dev = x.device
        inner = heads * dim_head
        bf = lambda w: ops.cast_bf16(w.detach().contiguous())
        f32 = lambda *shape: torch.empty(*shape, device=dev)
        b16 = lambda *shape: torch.empty(*shape, device=dev, dtype=torch.bfloat16)
        x = x.detach().contiguous().clone()
        saved = []
        for li in range(len(params) // _PER_LAYER):
            (ln1w, ln1b, qw, qb, kvw, kvb, ew, eb, ow, ob, g1, ln2w, ln2b, f1w, f1b, f2w, f2b, g2) = \
                params[li * _PER_LAYER:(li + 1) * _PER_LAYER]
            x_in = x.clone()
            xn, mean1, rstd1 = b16(R, dim), f32(R), f32(R)
            ops.layernorm_fwd(x, ln1w.detach(), ln1b.detach(), 1e-5, None, xn, mean=mean1, rstd=rstd1)
            wqkv = bf(torch.cat([qw, kvw], 0))
            qkv = f32(R, 3 * inner)
            ops.gemm(xn, wqkv, bias=torch.cat([qb, kvb]).detach().clone(), out_f32=qkv)
            att = b16(R, inner)
            we, be = ew.detach()[:, 0].contiguous(), eb.detach().contiguous()
            ops.graph_attn_core(qkv, node_off, upper, nmax, we, be, att)
            wo = bf(ow)
            o1 = f32(R, dim)
            ops.gemm(att, wo, bias=ob.detach(), out_f32=o1)
            ops.gated_residual(o1, x, g1.detach().reshape(-1).contiguous())          # x <- gate(o1, x_in)
            x_mid = x.clone()
            xn2, mean2, rstd2 = b16(R, dim), f32(R), f32(R)
            ops.layernorm_fwd(x, ln2w.detach(), ln2b.detach(), 1e-5, None, xn2, mean=mean2, rstd=rstd2)
            w1, w2 = bf(f1w), bf(f2w)
            z = b16(R, 4 * dim)
            ops.gemm(xn2, w1, bias=f1b.detach(), out_bf16=z)
            hid = ops.act_dropout(z, ops.ACT_GELU)
            o2 = f32(R, dim)
            ops.gemm(hid, w2, bias=f2b.detach(), out_f32=o2)
            ops.gated_residual(o2, x, g2.detach().reshape(-1).contiguous())          # x <- gate(o2, x_mid)
            saved.append(dict(x_in=x_in, xn=xn, mean1=mean1, rstd1=rstd1, wqkv=wqkv, qkv=qkv, att=att, we=we, be=be, wo=wo,
                              o1=o1, x_mid=x_mid, xn2=xn2, mean2=mean2, rstd2=rstd2, w1=w1, w2=w2, z=z, hid=hid, o2=o2))
        ctx.saved, ctx.geom, ctx.params = saved, (node_off, upper, nmax, heads, dim_head), params
        return x

    @staticmethod
    def backward(ctx, dx):
        node_off, upper, nmax, heads, dim_head = ctx.geom
        params, saved = ctx.params, ctx.saved
        ctx.saved = None
        dev = dx.device
        inner = heads * dim_head
        R, dim = dx.shape
        f32 = lambda *shape: torch.empty(*shape, device=dev)
        zeros = lambda *shape: torch.zeros(*shape, device=dev)
        grads = [None] * len(params)
        dx = dx.contiguous().float()

        def gate_bwd(o, res, gw, dxx):
            d_o, d_res, da = f32(R, dim), f32(R, dim), f32(R)
            ops.gated_residual_bwd(o, res, gw.detach().reshape(-1).contiguous(), dxx, d_o, d_res, da)
            dw1, dw2 = zeros(dim), zeros(dim)
            ops.weighted_colsum(o, da, dw1)
            ops.weighted_colsum(res, da, dw2)
            return d_o, d_res, torch.cat([dw1, dw2, dw1 - dw2]).view(1, 3 * dim)

        def linear_bwd(dy32, x_b, w_b, n_out, k_in, mask=None):
            """dy fp32 [R, n_out] of y = x W^T + b -> (dW, db, dx fp32 or bf16-masked)."""
            dyb = ops.cast_bf16(dy32)
            dW = f32(n_out, k_in)
            ops.gemm(dyb, x_b, a_mn=True, b_mn=True, out_f32=dW)
            db = zeros(1, n_out)
            ops.colsum(dyb, db)
            return dyb, dW, db[0]

        for li in reversed(range(len(saved))):
            S = saved[li]
            base = li * _PER_LAYER
            (ln1w, ln1b, qw, qb, kvw, kvb, ew, eb, ow, ob, g1, ln2w, ln2b, f1w, f1b, f2w, f2b, g2) = params[base:base + _PER_LAYER]
            # ---- x_out = gate2(o2, x_mid)
            d_o2, d_mid_skip, grads[base + 17] = gate_bwd(S["o2"], S["x_mid"], g2, dx)
            # ---- o2 = gelu(z) W2^T + b2,  z = LN2(x_mid) W1^T + b1
            d_o2b, grads[base + 15], grads[base + 16] = linear_bwd(d_o2, S["hid"], S["w2"], dim, 4 * dim)
            dz = torch.empty(R, 4 * dim, device=dev, dtype=torch.bfloat16)
            ops.gemm(d_o2b, S["w2"], b_mn=True, mask_src=S["z"], mask_mode=ops.MASK_GELU, out_bf16=dz)
            dW1 = f32(4 * dim, dim)
            ops.gemm(dz, S["xn2"], a_mn=True, b_mn=True, out_f32=dW1)
            db1 = zeros(1, 4 * dim)
            ops.colsum(dz, db1)
            grads[base + 13], grads[base + 14] = dW1, db1[0]
            dxn2 = f32(R, dim)
            ops.gemm(dz, S["w1"], b_mn=True, out_f32=dxn2)
            d_mid, dg2, dbt2 = f32(R, dim), zeros(dim), zeros(dim)
            ops.layernorm_bwd(dxn2, S["x_mid"], ln2w.detach(), S["mean2"], S["rstd2"], d_mid, None, 0.0, 0, dg2, dbt2,
                              base=d_mid_skip)
            grads[base + 11], grads[base + 12] = dg2, dbt2
            # ---- x_mid = gate1(o1, x_in)
            d_o1, d_in_skip, grads[base + 10] = gate_bwd(S["o1"], S["x_in"], g1, d_mid)
            # ---- o1 = att Wo^T + bo
            d_o1b, grads[base + 8], grads[base + 9] = linear_bwd(d_o1, S["att"], S["wo"], dim, inner)
            datt = f32(R, inner)
            ops.gemm(d_o1b, S["wo"], b_mn=True, out_f32=datt)
            # ---- attention core
            dqkv, dwe, dbe = f32(R, 3 * inner), zeros(inner), zeros(inner)
            ops.graph_attn_core_bwd(S["qkv"], node_off, upper, nmax, S["we"], S["be"], datt, dqkv, dwe, dbe)
            grads[base + 6], grads[base + 7] = dwe.view(inner, 1), dbe
            # ---- qkv = LN1(x_in) Wqkv^T + bqkv
            dqkvb, dWqkv, dbqkv = linear_bwd(dqkv, S["xn"], S["wqkv"], 3 * inner, dim)
            grads[base + 2], grads[base + 3] = dWqkv[:inner], dbqkv[:inner]
            grads[base + 4], grads[base + 5] = dWqkv[inner:], dbqkv[inner:]
            dxn = f32(R, dim)
            ops.gemm(dqkvb, S["wqkv"], b_mn=True, out_f32=dxn)
            d_in, dg1, dbt1 = f32(R, dim), zeros(dim), zeros(dim)
            ops.layernorm_bwd(dxn, S["x_in"], ln1w.detach(), S["mean1"], S["rstd1"], d_in, None, 0.0, 0, dg1, dbt1,
                              base=d_in_skip)
            grads[base + 0], grads[base + 1] = dg1, dbt1
            dx = d_in
            saved[li] = None
        return (dx, None, None, None, None, None, *grads)


class _SmallGraphTransformerFn(torch.autograd.Function):
    """The 10-wide STRUCTURE branch, layer by layer in fp32 with saved activations (SIMT linears / LayerNorms of
    csrc/consistency.cu around the same attention-core and gated-residual kernels).  Only the differentiable mode uses it;
    the default mode evaluates the branch in one launch (b200vsgg_graph_small_fwd).  The node features are constants
    (Laplacian eigenvectors), so no gradient is returned for them."""

    @staticmethod
    def forward(ctx, x, node_off, upper, nmax, heads, dim_head, *params):
        R, dim = x.shape
        inner = heads * dim_head
        x = x.detach().contiguous().float().clone()
        saved = []
        for li in range(len(params) // _PER_LAYER):
            (ln1w, ln1b, qw, qb, kvw, kvb, ew, eb, ow, ob, g1, ln2w, ln2b, f1w, f1b, f2w, f2b, g2) = \
                [t.detach().contiguous().float() for t in params[li * _PER_LAYER:(li + 1) * _PER_LAYER]]
            x_in = x.clone()
            xn, mean1, rstd1 = ops.ln_small_fwd(x, ln1w, ln1b)
            wqkv = torch.cat([qw, kvw], 0).contiguous()
            qkv = ops.simt_linear(xn, wqkv, torch.cat([qb, kvb]).contiguous())
            att = torch.empty(R, inner, device=x.device)
            we, be = ew[:, 0].contiguous(), eb
            ops.graph_attn_core(qkv, node_off, upper, nmax, we, be, att)
            o1 = ops.simt_linear(att, ow, ob)
            ops.gated_residual(o1, x, g1.reshape(-1).contiguous())
            x_mid = x.clone()
            xn2, mean2, rstd2 = ops.ln_small_fwd(x, ln2w, ln2b)
            hid, z = ops.simt_linear(xn2, f1w, f1b, act=ops.ACT_GELU, want_z=True)
            o2 = ops.simt_linear(hid, f2w, f2b)
            ops.gated_residual(o2, x, g2.reshape(-1).contiguous())
            saved.append(dict(x_in=x_in, xn=xn, mean1=mean1, rstd1=rstd1, wqkv=wqkv, qkv=qkv, att=att, we=we, be=be, ow=ow,
                              o1=o1, g1=g1, x_mid=x_mid, xn2=xn2, mean2=mean2, rstd2=rstd2, f1w=f1w, f2w=f2w, z=z, hid=hid,
                              o2=o2, g2=g2, ln1w=ln1w, ln2w=ln2w))
        ctx.saved, ctx.geom, ctx.n_params = saved, (node_off, upper, nmax, heads, dim_head), len(params)
        return x

    @staticmethod
    def backward(ctx, dx):
        node_off, upper, nmax, heads, dim_head = ctx.geom
        saved = ctx.saved
        ctx.saved = None
        dev = dx.device
        inner = heads * dim_head
        R, dim = dx.shape
        zeros = lambda *shape: torch.zeros(*shape, device=dev)
        grads = [None] * ctx.n_params
        dx = dx.contiguous().float()

        def colsum(t):
            out = zeros(1, t.shape[1])
            ops.colsum(t.contiguous(), out)
            return out[0]

        def gate_bwd(o, res, gw, dxx):
            d_o, d_res, da = torch.empty_like(o), torch.empty_like(o), torch.empty(R, device=dev)
            ops.gated_residual_bwd(o, res, gw.reshape(-1).contiguous(), dxx, d_o, d_res, da)
            dw1, dw2 = zeros(dim), zeros(dim)
            ops.weighted_colsum(o, da, dw1)
            ops.weighted_colsum(res, da, dw2)
            return d_o, d_res, torch.cat([dw1, dw2, dw1 - dw2]).view(1, 3 * dim)

        for li in reversed(range(len(saved))):
            S = saved[li]
            base = li * _PER_LAYER
            d_o2, d_mid_skip, grads[base + 17] = gate_bwd(S["o2"], S["x_mid"], S["g2"], dx)
            grads[base + 15], grads[base + 16] = ops.simt_wgrad(d_o2, S["hid"]), colsum(d_o2)
            dz = ops.gelu_bwd(ops.simt_linear(d_o2, S["f2w"], transposed=True), S["z"])
            grads[base + 13], grads[base + 14] = ops.simt_wgrad(dz, S["xn2"]), colsum(dz)
            dxn2 = ops.simt_linear(dz, S["f1w"], transposed=True)
            d_mid, grads[base + 11], grads[base + 12] = ops.ln_small_bwd(dxn2, S["x_mid"], S["ln2w"], S["mean2"], S["rstd2"],
                                                                         base=d_mid_skip)
            d_o1, d_in_skip, grads[base + 10] = gate_bwd(S["o1"], S["x_in"], S["g1"], d_mid)
            grads[base + 8], grads[base + 9] = ops.simt_wgrad(d_o1, S["att"]), colsum(d_o1)
            datt = ops.simt_linear(d_o1, S["ow"], transposed=True)
            dqkv, dwe, dbe = torch.empty(R, 3 * inner, device=dev), zeros(inner), zeros(inner)
            ops.graph_attn_core_bwd(S["qkv"], node_off, upper, nmax, S["we"], S["be"], datt, dqkv, dwe, dbe)
            grads[base + 6], grads[base + 7] = dwe.view(inner, 1), dbe
            dWqkv, dbqkv = ops.simt_wgrad(dqkv, S["xn"]), colsum(dqkv)
            grads[base + 2], grads[base + 3] = dWqkv[:inner], dbqkv[:inner]
            grads[base + 4], grads[base + 5] = dWqkv[inner:], dbqkv[inner:]
            dxn = ops.simt_linear(dqkv, S["wqkv"], transposed=True)
            dx, grads[base + 0], grads[base + 1] = ops.ln_small_bwd(dxn, S["x_in"], S["ln1w"], S["mean1"], S["rstd1"],
                                                                    base=d_in_skip)
            saved[li] = None
        return (None, None, None, None, None, None, *grads)


class _AttnPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, node_off, n_frames, nmax, w, b):
        x = x.contiguous()
        ctx.save_for_backward(x, node_off, w, b)
        ctx.geom = (n_frames, nmax)
        return ops.attn_pool(x, node_off, n_frames, nmax, w, b)

    @staticmethod
    def backward(ctx, dout):
        x, node_off, w, b = ctx.saved_tensors
        n_frames, nmax = ctx.geom
        dx, dgate = torch.empty_like(x), torch.zeros(x.shape[0], device=x.device)
        wv, bv = w.detach().reshape(-1).contiguous().float(), b.detach().reshape(-1).contiguous().float()
        ops.attn_pool_bwd(x, node_off, n_frames, nmax, wv, bv, dout.contiguous().float(), dx, dgate)
        dw = torch.zeros(x.shape[1], device=x.device)
        ops.weighted_colsum(x, dgate, dw)
        return dx, None, None, None, dw.view_as(w), dgate.sum().view_as(b)


class _ConsistencyKLFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, g, pu, pv):
        g = g.contiguous()
        ctx.save_for_backward(g, pu, pv)
        return ops.consistency_kl(g, pu, pv)

    @staticmethod
    def backward(ctx, gout):
        g, pu, pv = ctx.saved_tensors
        dg = torch.zeros_like(g)
        ops.consistency_kl_bwd(g, pu, pv, gout.contiguous().float(), dg)
        return dg, None, None


def run_compact_differentiable(gt, x, node_off, upper, nmax):
    return _GraphTransformerFn.apply(x, node_off, upper, nmax, gt.heads, gt.dim_head, *_layer_params(gt))


def _pool_compact(x, node_off, n_frames, nmax, gate_nn):
    """GlobalAttentionPooling over compact rows: per-frame softmax of gate_nn(x), weighted sum -> [frames, dim]
    (b200vsgg_attn_pool, one CTA per frame)."""
    return ops.attn_pool(x.contiguous(), node_off, n_frames, nmax, gate_nn.weight, gate_nn.bias)


def consistency_losses(gat, gat_semantic, gate_nn, gate_sem_nn, plan, spatial_flags, hidden, clip_first_row=None,
                       clip_rows=None, flags_host=None, differentiable=False):
    """differentiable=False (default, the reference's behaviour: both vectors detached).  differentiable=True: both loss
    vectors carry gradients — the semantic one into gat_semantic, gate_sem_nn and `hidden` (pass it un-detached), the
    structure one into gat and gate_nn (its node features are constant Laplacian eigenvectors)."""
    if not differentiable:
        with torch.no_grad():
            return _consistency_losses(gat, gat_semantic, gate_nn, gate_sem_nn, plan, spatial_flags, hidden.detach(),
                                       clip_first_row, clip_rows, flags_host, False)
    return _consistency_losses(gat, gat_semantic, gate_nn, gate_sem_nn, plan, spatial_flags, hidden, clip_first_row,
                               clip_rows, flags_host, True)


def _consistency_losses(gat, gat_semantic, gate_nn, gate_sem_nn, plan, spatial_flags, hidden, clip_first_row, clip_rows,
                        flags_host, differentiable):
    """Returns (structure_temp_loss [P], semantic_temp_loss [P']) for the batch described by `plan`
    (teatgt.TeatPlan); spatial_flags uint8 [F, nmax, nmax] on the DEVICE (b200vsgg_teat_pair_flags); hidden
    [rows, d_sem] = per-clip feature rows (TEAT-GT: node order, the default; clip_first_row / clip_rows [F] override
    where a frame's clip starts in `hidden` and how many rows that clip owns)."""
    dev = hidden.device
    F_, nmax = plan.F, plan.nmax
    counts_h = np.diff(plan.node_off_h)
    n_nodes = int(plan.node_off_h[-1])
    up_ = lambda a: ops.upload(a, dev)
    # ---- R2 first (device only, asynchronous): semantic nodes = the clip's hidden rows [0:n_f] (`savor` quirk)
    if clip_first_row is None:
        clip_first_row = plan.clip_node_off[plan.clip_of_frame]
        clip_rows = np.diff(plan.clip_node_off)[plan.clip_of_frame]
    local = np.arange(n_nodes) - np.repeat(plan.node_off_h[:-1], counts_h)
    src = np.repeat(np.asarray(clip_first_row, dtype=np.int64), counts_h) + local
    valid = local < np.repeat(np.asarray(clip_rows, dtype=np.int64), counts_h)
    src_t = up_(np.where(valid, src, 0))
    x = hidden[src_t]
    if not valid.all():
        x = x * up_(valid.astype(np.float32))[:, None]
    if not hidden.is_cuda:
        raise RuntimeError("consistency_losses runs only on CUDA tensors (no CPU fallback for the hot path)")
    if nmax > MAX_NODES_STRUCT or nmax > MAX_NODES_SEM or gat.dim > 16 or gat.dim_head != 64 or hidden.shape[1] % 8:
        raise RuntimeError("consistency regulariser: a frame has %d nodes / widths (%d, %d); the on-chip graph kernels "
                           "take <= %d nodes per frame, structure width <= 16, semantic width %% 8 == 0 (no eager "
                           "fallback)" % (nmax, gat.dim, hidden.shape[1], min(MAX_NODES_STRUCT, MAX_NODES_SEM)))
    if differentiable:
        sem_rows = run_compact_differentiable(gat_semantic, x, plan.node_off, spatial_flags, nmax)
        sem = _AttnPoolFn.apply(sem_rows, plan.node_off, F_, nmax, gate_sem_nn.weight, gate_sem_nn.bias)
    else:
        sem_rows = run_compact(gat_semantic, x, plan.node_off, spatial_flags, nmax)
        sem = _pool_compact(sem_rows, plan.node_off, F_, nmax, gate_sem_nn)
    # ---- R1: per-frame Laplacian eigenvectors on the host (the reference's LAPACK call), grouped by node count;
    #      this runs while the device works on the semantic branch
    if flags_host is not None:        # (pinned host copy, event): the D2H was issued before the main path
        flags_host[1].synchronize()
        up = flags_host[0].numpy().astype(np.float64)
    else:
        up = spatial_flags.cpu().numpy().astype(np.float64)
    A = up + up.transpose(0, 2, 1)                                   # both directions were added as edges
    k = 10
    ev = np.zeros((F_, nmax, k), dtype=np.float32)

    def solve(idx):
        nf = int(counts_h[idx[0]])
        a = A[idx][:, :nf, :nf]
        deg = a.sum(1)                                              # in-degree (symmetric)
        nm = (torch.from_numpy(deg.astype(np.int64)).clip(1) ** -0.5).numpy().astype(np.float64)
        L = np.eye(nf)[None] - nm[:, :, None] * a * nm[:, None, :]
        _, vec = np.linalg.eigh(L)                                   # stacked LAPACK calls, GIL released
        vec = vec.astype(np.float32)
        if k > nf:
            vec = np.tile(vec, (1, 1, int(k / 2)))[:, :, :k]         # lib/teatgt.py:304-305
        else:
            vec = vec[:, :, :k]
        ev[idx, :nf] = vec

    jobs = []
    for nf in np.unique(counts_h):
        idx = np.nonzero(counts_h == nf)[0]
        jobs += [idx[i:i + 256] for i in range(0, idx.shape[0], 256)]
    if len(jobs) <= 2:
        for j in jobs:
            solve(j)
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(8) as pool:
            list(pool.map(solve, jobs))
    counts = up_(counts_h)
    nodes = up_(ev)
    # R1 in one launch: 4-layer GraphTransformer(dim 10) + attention pooling, one CTA per frame
    if differentiable:
        # same network, layer by layer with saved activations: compact node rows in frame order
        rows = ev[np.arange(nmax)[None, :] < counts_h[:, None]]
        st_rows = _SmallGraphTransformerFn.apply(up_(rows), plan.node_off, spatial_flags.contiguous(), nmax, gat.heads,
                                                 gat.dim_head, *_layer_params(gat))
        sym = _AttnPoolFn.apply(st_rows, plan.node_off, F_, nmax, gate_nn.weight, gate_nn.bias)
    else:
        sym = ops.graph_small_fwd(nodes, spatial_flags.contiguous(), counts.int(), gat.dim, gat.heads, len(gat.layers),
                                  pack_small_params(gat), gate_nn.weight.detach().reshape(-1).contiguous(),
                                  gate_nn.bias.detach().contiguous())
    # ---- R3: all frame pairs u < v inside each clip, reference order
    pu, pv = [], []
    frames_pc = np.bincount(plan.clip_of_frame, minlength=plan.n_clips)
    f0 = 0
    tri = {}
    for nf in frames_pc:
        if int(nf) not in tri:
            tri[int(nf)] = np.triu_indices(int(nf), 1)
        iu, iv = tri[int(nf)]
        pu.append(f0 + iu)
        pv.append(f0 + iv)
        f0 += int(nf)
    pu = up_(np.concatenate(pu).astype(np.int32))
    pv = up_(np.concatenate(pv).astype(np.int32))
    if differentiable:
        s, m = _ConsistencyKLFn.apply(sym, pu, pv), _ConsistencyKLFn.apply(sem, pu, pv)
    else:
        s, m = ops.consistency_kl(sym.contiguous(), pu, pv), ops.consistency_kl(sem.contiguous(), pu, pv)
    # the reference keeps a pair only if its score is >= 0 (lib/teatgt.py:327-333): a data-dependent length, hence a host
    # synchronisation — ONE flag for both branches; the usual case (no negative rounding residue) returns the kernels'
    # outputs as they are, without any gather
    if bool((torch.cat([s, m]) >= 0).all()):          # (NaN >= 0 is False: dropped like in the reference)
        return s, m
    return s[s >= 0], m[m >= 0]
