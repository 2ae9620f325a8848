"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

Inert / minimal stand-ins for third-party packages that the reference imports but that are absent
from this image, installed into ``sys.modules`` so that the UNMODIFIED reference modules can be
imported in the build container (oracle/make_golden*.py).  Nothing here is reference source.

  fairseq                  only wrappers are used (dropout / LayerNorm / softmax / registry decorators):
                           semantics follow fairseq exactly
  matplotlib.pyplot        imported for visualisation only
  dgl                      DGLGraph as a dense-adjacency holder (add_nodes / add_edges /
                           adjacency_matrix_scipy / in_degrees / out_degrees) and
                           dgl.nn.GlobalAttentionPooling (softmax-gated sum over nodes)
  graph_transformer_pytorch  lucidrains' GraphTransformer, restated from its published
                           architecture — UNPINNED: the package and its version are not
                           recorded by the reference (docker_cmd.txt:2-5 lists bare pip names)
  tools.fasterRCNN..., tools.utils.fpn, tools.utils.draw_rectangles   absent native ops, never called
                           on the PredCLS path
"""
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


# ------------------------------------------------------------------------------------------ fairseq
class FairseqDropout(nn.Module):
    def __init__(self, p, module_name=None):
        super().__init__()
        self.p = p
        self.module_name = module_name
        self.apply_during_inference = False

    def forward(self, x, inplace: bool = False):
        if self.p > 0 and (self.training or self.apply_during_inference):
            return F.dropout(x, p=self.p, training=True, inplace=inplace)
        return x


def LayerNorm(normalized_shape, eps=1e-5, elementwise_affine=True, export=False):
    return nn.LayerNorm(normalized_shape, eps, elementwise_affine)


class LayerDropModuleList(nn.ModuleList):
    def __init__(self, p, modules=None):
        super().__init__(modules)
        self.p = p


class FairseqEncoder(nn.Module):
    def __init__(self, dictionary):
        super().__init__()
        self.dictionary = dictionary


class FairseqEncoderModel(nn.Module):
    def __init__(self, encoder):
        super().__init__()
        self.encoder = encoder


def _register(*_a, **_k):
    def deco(x):
        return x
    return deco


def _get_activation_fn(name):
    return {"relu": F.relu, "gelu": F.gelu, "tanh": torch.tanh, "linear": lambda x: x}[name]


def _softmax(x, dim, onnx_trace=False):
    return F.softmax(x, dim=dim, dtype=torch.float32)


def _quant_noise(module, p, block_size):
    assert p <= 0
    return module


def install_fairseq():
    utils = _mod("fairseq.utils", get_activation_fn=_get_activation_fn, softmax=_softmax,
                 get_available_activation_fns=lambda: ["relu", "gelu", "tanh", "linear"],
                 safe_hasattr=lambda o, k: getattr(o, k, None) is not None)
    models = _mod("fairseq.models", FairseqEncoder=FairseqEncoder, FairseqEncoderModel=FairseqEncoderModel,
                  register_model=_register, register_model_architecture=_register)
    ln = _mod("fairseq.modules.layer_norm", LayerNorm=LayerNorm, LayerDropModuleList=LayerDropModuleList)
    fd = _mod("fairseq.modules.fairseq_dropout", FairseqDropout=FairseqDropout)
    qn = _mod("fairseq.modules.quant_noise", quant_noise=_quant_noise)
    modules = _mod("fairseq.modules", LayerNorm=LayerNorm, LayerDropModuleList=LayerDropModuleList,
                   FairseqDropout=FairseqDropout, layer_norm=ln, fairseq_dropout=fd, quant_noise=qn)
    _mod("fairseq", utils=utils, models=models, modules=modules)


# ---------------------------------------------------------------------------------------------- dgl
class DGLGraph:
    def __init__(self):
        self.n, self.src, self.dst = 0, [], []

    def to(self, device):
        return self

    def add_nodes(self, n):
        self.n += int(n)

    def add_edges(self, u, v):
        self.src += [int(x) for x in (u.tolist() if torch.is_tensor(u) else u)]
        self.dst += [int(x) for x in (v.tolist() if torch.is_tensor(v) else v)]

    def number_of_nodes(self):
        return self.n

    def adjacency_matrix_scipy(self, return_edge_ids=False):
        import scipy.sparse as sp
        data = np.ones(len(self.src))
        return sp.coo_matrix((data, (self.dst, self.src)), shape=(self.n, self.n)).tocsr()

    def in_degrees(self):
        return torch.bincount(torch.tensor(self.dst, dtype=torch.int64), minlength=self.n)

    def out_degrees(self):
        return torch.bincount(torch.tensor(self.src, dtype=torch.int64), minlength=self.n)


class GlobalAttentionPooling(nn.Module):
    def __init__(self, gate_nn, feat_nn=None):
        super().__init__()
        self.gate_nn, self.feat_nn = gate_nn, feat_nn

    def forward(self, graph, feat):
        gate = torch.softmax(self.gate_nn(feat), dim=0)
        feat = self.feat_nn(feat) if self.feat_nn is not None else feat
        return (gate * feat).sum(0, keepdim=True)


def install_dgl():
    fn = _mod("dgl.function")
    dnn = _mod("dgl.nn", GlobalAttentionPooling=GlobalAttentionPooling)
    _mod("dgl", DGLGraph=DGLGraph, function=fn, nn=dnn)


# ---------------------------------------------------------------------- graph_transformer_pytorch
class _Rotary(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.register_buffer("inv_freq", 1.0 / (10000 ** (torch.arange(0, dim, 2).float() / dim)), persistent=False)

    def forward(self, t):
        freqs = torch.einsum("i,j->ij", t.float(), self.inv_freq)
        return torch.repeat_interleave(freqs, 2, dim=-1)


def _rotate_half(x):
    x = x.reshape(*x.shape[:-1], -1, 2)
    x1, x2 = x.unbind(-1)
    return torch.stack((-x2, x1), -1).flatten(-2)


def _apply_rotary(freqs, t):
    return t * freqs.cos() + _rotate_half(t) * freqs.sin()


class _GTAttention(nn.Module):
    def __init__(self, dim, pos_emb, dim_head, heads, edge_dim):
        super().__init__()
        inner = dim_head * heads
        self.heads, self.scale, self.pos_emb = heads, dim_head ** -0.5, pos_emb
        self.to_q = nn.Linear(dim, inner)
        self.to_kv = nn.Linear(dim, inner * 2)
        self.edges_to_kv = nn.Linear(edge_dim, inner)
        self.to_out = nn.Linear(inner, dim)

    def forward(self, nodes, edges):
        b, n, _ = nodes.shape
        h = self.heads
        q = self.to_q(nodes)
        k, v = self.to_kv(nodes).chunk(2, dim=-1)
        e = self.edges_to_kv(edges)
        q, k, v = (t.view(b, n, h, -1).permute(0, 2, 1, 3).reshape(b * h, n, -1) for t in (q, k, v))
        e = e.view(b, n, n, h, -1).permute(0, 3, 1, 2, 4).reshape(b * h, n, n, -1)
        if self.pos_emb is not None:
            freqs = self.pos_emb(torch.arange(n, device=nodes.device))[None]
            q, k = _apply_rotary(freqs, q), _apply_rotary(freqs, k)
        k = k[:, None] + e
        v = v[:, None] + e
        sim = torch.einsum("bid,bijd->bij", q, k) * self.scale
        attn = sim.softmax(dim=-1)
        out = torch.einsum("bij,bijd->bid", attn, v)
        out = out.view(b, h, n, -1).permute(0, 2, 1, 3).reshape(b, n, -1)
        return self.to_out(out)


class _PreNorm(nn.Module):
    def __init__(self, dim, fn):
        super().__init__()
        self.fn, self.norm = fn, nn.LayerNorm(dim)

    def forward(self, x, *args):
        return self.fn(self.norm(x), *args)


class _GatedResidual(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.proj = nn.Sequential(nn.Linear(dim * 3, 1, bias=False), nn.Sigmoid())

    def forward(self, x, res):
        gate = self.proj(torch.cat((x, res, x - res), dim=-1))
        return x * gate + res * (1 - gate)


class GraphTransformer(nn.Module):
    def __init__(self, dim, depth, dim_head=64, edge_dim=None, heads=8, gated_residual=True, with_feedforwards=False,
                 norm_edges=False, rel_pos_emb=False, accept_adjacency_matrix=False):
        super().__init__()
        edge_dim = edge_dim if edge_dim is not None else dim
        self.norm_edges = nn.LayerNorm(edge_dim) if norm_edges else nn.Identity()
        pos_emb = _Rotary(dim_head) if rel_pos_emb else None
        self.layers = nn.ModuleList()
        for _ in range(depth):
            ff = nn.ModuleList([_PreNorm(dim, nn.Sequential(nn.Linear(dim, dim * 4), nn.GELU(), nn.Linear(dim * 4, dim))),
                                _GatedResidual(dim)]) if with_feedforwards else None
            self.layers.append(nn.ModuleList([
                nn.ModuleList([_PreNorm(dim, _GTAttention(dim, pos_emb, dim_head, heads, edge_dim)), _GatedResidual(dim)]),
                ff]))

    def forward(self, nodes, edges=None, adj_mat=None, mask=None):
        edges = self.norm_edges(edges)
        for attn_block, ff_block in self.layers:
            attn, res = attn_block
            nodes = res(attn(nodes, edges), nodes)
            if ff_block is not None:
                ff, res2 = ff_block
                nodes = res2(ff(nodes), nodes)
        return nodes, edges


def install_graph_transformer():
    _mod("graph_transformer_pytorch", GraphTransformer=GraphTransformer)


# ------------------------------------------------------------------------------ pytorch_metric_learning
def contrastive_loss(embeddings, labels, pos_margin=0.0, neg_margin=1.0):
    """Restatement of pytorch_metric_learning.losses.ContrastiveLoss(pos_margin, neg_margin) with its defaults
    (LpDistance(normalize_embeddings=True, p=2, power=1), all pairs i != j of the call, AvgNonZeroReducer per group,
    the two group means added) — the package is absent from the reference tree and unversioned: PARITY UNPINNED.
    Call sites: TEMPURA_train.py:103,209-212; TEATGT_train.py:81,176-179."""
    e = torch.nn.functional.normalize(embeddings, p=2, dim=1)
    d = torch.cdist(e, e, p=2, compute_mode="donot_use_mm_for_euclid_dist")   # LpDistance -> torch.cdist (gradient 0 at d = 0)
    n = e.shape[0]
    off = ~torch.eye(n, dtype=torch.bool)
    same = labels[:, None] == labels[None, :]
    pos = torch.relu(d - pos_margin)[same & off]
    neg = torch.relu(neg_margin - d)[(~same) & off]

    def avg_non_zero(v):
        nz = v > 0
        return v[nz].sum() / nz.sum() if nz.any() else v.sum() * 0.0

    return avg_non_zero(pos) + avg_non_zero(neg)


# ----------------------------------------------------------------------------- absent native ops
def install_reference_native_stubs():
    class _InertROIAlign(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    for n in ("tools.fasterRCNN", "tools.fasterRCNN.lib", "tools.fasterRCNN.lib.model", "tools.utils.fpn",
              "tools.utils.draw_rectangles"):
        _mod(n)
    _mod("tools.fasterRCNN.lib.model.roi_layers", ROIAlign=_InertROIAlign, nms=None)
    from oracle.tempura_oracle import center_size   # absent from the reference tree: injected (3 lines, unpinned)
    _mod("tools.utils.fpn.box_utils", center_size=center_size)
    _mod("tools.utils.draw_rectangles.draw_rectangles", draw_union_boxes=None)


def install_matplotlib():
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        plt = _mod("matplotlib.pyplot")
        _mod("matplotlib", pyplot=plt)


def install_all():
    install_fairseq()
    install_dgl()
    install_graph_transformer()
    install_reference_native_stubs()
    install_matplotlib()
