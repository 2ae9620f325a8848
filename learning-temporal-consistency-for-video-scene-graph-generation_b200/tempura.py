"""B200-native TEMPURA (PredCLS relation path) behind the reference's module API.

Drop-in for `lib/tempura.py::TEMPURA` of the reference (constructor lib/tempura.py:428-432,
`forward(entry, phase, unc)` :512-598): same keyword arguments, same `state_dict()` key names and
shapes (checkpoints load with strict=True), same entry-dict keys in and out, same mutable
attributes (`rel_memory`, `object_classifier.obj_memory`, `obj_classes`, `*_class_num`, `mode`).
Sub-modules are constructed in the reference's order with the same torch initialisers, so the
same `torch.manual_seed` yields the same initial weights (GloVe vectors are replaced by seeded
N(0,1) rows unless `embed_vecs` is given — there is no network in this environment).

Underneath, every tensor op of the path runs in hand-written sm_100a kernels from
libb200vsgg.so (ops.py): tcgen05/TMA GEMMs, varlen attention, LayerNorm, gathers, GMM epilogue.
There is NO CPU or eager-PyTorch fallback: CPU tensors raise.  Extensions over the reference API
(all optional): a *batch* of videos can be passed (see `collate_entries`); BatchNorm statistics of
the mask branch (lib/tempura.py:466-474) are then taken per video, because the reference's batch is
one video.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .plan import SegmentPlan, plan_from_im_idx

D_MODEL = 1936
N_HEADS = 8
HEAD_DIM = D_MODEL // N_HEADS
FFN_DIM = 2048
HEAD_COLS_PAD = 8  # packed head GEMM width is rounded up to a multiple of 8

# Two result-preserving savings in the temporal decoder ("legal algebraic savings"; window tokens are
# M2 ~ 1.94 N rows because every interior frame lives in two windows):
#   DEC_FIRST_ON_PAIRS  the window tokens of the FIRST decoder layer are copies of pair rows (transformer.py:203-215), so
#                       its q/k/v projections run once over the N pair rows — q, k = (x + pos) Wqk = x Wqk + pos Wqk — and
#                       are gathered into the windows; the backward sums the <= 2 window copies of dqkv before its GEMMs.
#   DEC_LATTER_ONLY     the 'latter' read-out (transformer.py:236-242) keeps N of the LAST layer's M2 output rows; the
#                       others never reach an output or a loss, so after the attention (whose keys / values need every
#                       row) out-proj, LayerNorm and the FFN run on the N kept rows only.  Bit-identical outputs.
# Environment B200VSGG_DEC_FIRST_ON_PAIRS=0 / B200VSGG_DEC_LATTER_ONLY=0 restore the dense M2-row schedule (A/B timing).
import os as _os
DEC_FIRST_ON_PAIRS = _os.environ.get("B200VSGG_DEC_FIRST_ON_PAIRS", "1") != "0"
DEC_LATTER_ONLY = _os.environ.get("B200VSGG_DEC_LATTER_ONLY", "1") != "0"


def _fgemm(*args, **kw):
    """Forward-path GEMM: never split-K.  Split-K partial tiles are summed in arrival order (TMA reduce-add), and a
    last-bit difference in a forward activation flips bf16 roundings downstream until, four layers later, two runs of
    the same input differ by a third of the total bf16 error (measured 3e-4 on the distributions).  Without it the eval
    forward is bit-reproducible; only tiny batches (< 148 output tiles) lose speed, the benchmark shapes never split."""
    return ops.gemm(*args, split_k=1, **kw)


# ================================================================================================
# parameter containers (names/shapes/initialisation order follow the reference)
# ================================================================================================
class GMMHead(nn.Module):
    """Parameters of tools/utils/gmm_heads.py::GMM_head; evaluation happens in the fused kernels."""

    def __init__(self, hid_dim, num_classes, rel_type=None, k=4):
        super().__init__()
        self.k, self.num_classes, self.rel_type = k, num_classes, rel_type
        self.heads = nn.ModuleDict()
        for i in range(k):
            self.heads.update({"mu_%d" % (i + 1): nn.Linear(hid_dim, num_classes),
                               "pi_%d" % (i + 1): nn.Linear(hid_dim, 1),
                               "var_%d" % (i + 1): nn.Linear(hid_dim, num_classes)})

    @property
    def softmax(self):
        return self.rel_type == "attention" or self.rel_type is None

    def packed_order(self):
        """The head's Linear layers in the kernel's column order mu_1..K | var_1..K | pi_1..K."""
        K = self.k
        order = ["mu_%d" % (i + 1) for i in range(K)] + ["var_%d" % (i + 1) for i in range(K)] + \
                ["pi_%d" % (i + 1) for i in range(K)]
        return [self.heads[n] for n in order]

    def packed(self):
        """([K*(2C+1), hid] weight, [K*(2C+1)] bias) in the kernel's column order mu|var|pi."""
        lins = self.packed_order()
        return (torch.cat([l.weight for l in lins], 0), torch.cat([l.bias for l in lins], 0))


class _SpatialLayer(nn.Module):
    def __init__(self, dim, heads, ffn, p):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(dim, heads, dropout=p)
        self.linear1 = nn.Linear(dim, ffn)
        self.linear2 = nn.Linear(ffn, dim)
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)


class _TemporalLayer(nn.Module):
    def __init__(self, dim, heads, ffn, p):
        super().__init__()
        self.multihead2 = nn.MultiheadAttention(dim, heads, dropout=p)
        self.linear1 = nn.Linear(dim, ffn)
        self.linear2 = nn.Linear(ffn, dim)
        self.norm3 = nn.LayerNorm(dim)


class _LayerStack(nn.Module):
    def __init__(self, first, n):
        super().__init__()
        import copy
        self.layers = nn.ModuleList([copy.deepcopy(first) for _ in range(n)])
        self.num_layers = n


class _STTran(nn.Module):
    """Parameters of tools/utils/transformer.py::transformer."""

    def __init__(self, enc_layer_num, dec_layer_num, embed_dim, nhead, dim_feedforward, dropout, mode, mem_compute,
                 mem_fusion, selection, selection_lambda):
        super().__init__()
        self.mode, self.mem_fusion, self.mem_compute, self.selection = mode, mem_fusion, mem_compute, selection
        self.dropout_p = dropout
        self.local_attention = _LayerStack(_SpatialLayer(embed_dim, nhead, dim_feedforward, dropout), enc_layer_num)
        if mem_compute:
            if mem_compute == "seperate":
                raise NotImplementedError("rel_mem_compute='seperate' is not on the accelerated path")
            self.mem_attention = nn.MultiheadAttention(embed_dim, 1, 0.0, bias=False)
            if selection == "manual":
                self.selector = selection_lambda
            else:
                self.selector = nn.Linear(embed_dim, 1)
        self.global_attention = _LayerStack(_TemporalLayer(embed_dim, nhead, dim_feedforward, dropout), dec_layer_num)
        self.position_embedding = nn.Embedding(2, embed_dim)
        nn.init.uniform_(self.position_embedding.weight)


class _PositionalEncoding(nn.Module):
    """Holds PositionalEncoding.pe (lib/tempura.py:26-37), a registered buffer of the reference's state_dict."""

    def __init__(self, d_model, max_len):
        super().__init__()
        from .objbranch import sinusoid_table
        self.register_buffer("pe", sinusoid_table(d_model, max_len))


class ObjectClassifier(nn.Module):
    """lib/tempura.py:51-423 (and TEAT-GT's copy tools/utils/object_classifier.py).  PredCLS: `pred_labels =
    labels` (:245-247).  SGCls, phase='train': the object branch of objbranch.py (feature build, class-sequence
    encoder when `tracking`, intermediate, GMM / linear head; :185-255).  Parameters are created in the
    reference's order with its names, so reference checkpoints load strictly."""

    def __init__(self, mode="sgdet", obj_head="gmm", K=4, obj_classes=None, mem_compute=None, selection=None,
                 selection_lambda=0.5, tracking=None, embed_vecs=None):
        super().__init__()
        self.classes, self.mode, self.GMM_K = obj_classes, mode, K
        self.obj_memory = []
        self.mem_compute, self.selection, self.tracking, self.obj_head = mem_compute, selection, tracking, obj_head
        self.obj_embed = nn.Embedding(len(obj_classes) - 1, 200)
        if embed_vecs is not None:
            self.obj_embed.weight.data = embed_vecs[1:].clone()
        self.pos_embed = nn.Sequential(nn.BatchNorm1d(4, momentum=0.01 / 10.0), nn.Linear(4, 128), nn.ReLU(inplace=True),
                                       nn.Dropout(0.1))
        self.obj_dim = 2048
        mem_embed = 1024
        if tracking:
            d_model = self.obj_dim + 200 + 128
            first = _SpatialLayer(d_model, 8, 1024, 0.1)   # = nn.TransformerEncoderLayer(d_model, 8, 1024, batch_first)
            self.positional_encoder = _PositionalEncoding(d_model, 600 if mode == "sgdet" else 400)
            self.encoder_tran = _LayerStack(first, 3)
            mem_embed = d_model
        if mem_compute:
            self.mem_attention = nn.MultiheadAttention(mem_embed, 1, 0.0, bias=False)
            if selection == "manual":
                self.selector = selection_lambda
            else:
                self.selector = nn.Linear(1024, 1)
        self.intermediate = nn.Sequential(nn.Linear(self.obj_dim + 200 + 128, 1024), nn.BatchNorm1d(1024), nn.ReLU())
        if obj_head == "gmm":
            self.decoder_lin = GMMHead(1024, len(obj_classes), None, K)
        else:
            self.decoder_lin = nn.Sequential(nn.Linear(1024, len(obj_classes)))
        self.dropout_p = 0.1
        self.gmm_eps = None             # {"object": [K,O,C]} noise to inject (parity tests); None = device RNG

    def hallucinate(self, feat):
        """memory_hallucinator (lib/tempura.py:165-182) against `obj_memory` [n_mem, d]."""
        bank = self.obj_memory.to(feat.device, torch.float32)
        D = feat.shape[1]
        w = self.mem_attention.in_proj_weight
        q = _GemmNT.apply(feat, w[:D])
        k = _GemmNT.apply(bank, w[D:2 * D])
        v = _GemmNT.apply(bank, w[2 * D:])
        p = torch.softmax(_GemmNT.apply(q, k) * (1.0 / math.sqrt(D)), -1)
        pad = (-p.shape[1]) % 8
        o = _GemmNT.apply(F.pad(p, (0, pad)), F.pad(v, (0, 0, 0, pad)).t().contiguous())
        mem = _GemmNT.apply(o, self.mem_attention.out_proj.weight)
        e = self.selector if self.selection == "manual" else self.selector(feat).sigmoid()
        return e * feat + (1 - e) * mem if e is not None else feat + mem

    def forward(self, entry, phase="train", unc=False):
        if self.mode == "predcls":
            entry["pred_labels"] = entry["labels"]
            return entry
        if self.mode != "sgcls" or phase != "train":
            raise NotImplementedError(
                "object branch: only mode='sgcls', phase='train' (with or without unc) is on the accelerated path — the "
                "test-time relabel / NMS / ROIAlign tail (lib/tempura.py:257-307, :310-421) needs the reference's absent "
                "CUDA ops")
        from .objbranch import run_object_branch
        fpv = entry.get("video_frames")
        if fpv is None:
            fpv = np.asarray([entry["human_idx"].shape[0] if "human_idx" in entry else int(entry["boxes"][-1, 0].item()) + 1])
        return run_object_branch(self, entry, phase, fpv, apply_heads, self.dropout_p, self.gmm_eps, unc=unc)


# ================================================================================================
# helpers
# ================================================================================================
def collate_entries(entries):
    """Concatenate per-video PredCLS entries into one batch entry (videos stay independent:
    windows, BatchNorm statistics and losses never cross `video_frames` boundaries)."""
    keys_cat = ["boxes", "labels", "scores", "im_idx", "pair_idx", "human_idx", "features", "union_feat", "union_box",
                "spatial_masks", "distribution"]
    out = {}
    box_base, frame_base = 0, 0
    parts = {k: [] for k in keys_cat}
    frames, gts = [], {"attention_gt": [], "spatial_gt": [], "contacting_gt": []}
    singles, seqs = [], []     # SGCls class sequences (tools/utils/ds_track.py): never cross a video
    for e in entries:
        nf = int(e["human_idx"].shape[0]) if "human_idx" in e else int(e["im_idx"][-1].item()) + 1
        for k in keys_cat:
            if k not in e:
                continue
            t = e[k]
            if k == "pair_idx" or k == "human_idx":
                t = t + box_base
            elif k == "im_idx":
                t = t + frame_base
            elif k in ("boxes", "union_box"):
                t = t.clone()
                t[:, 0] += frame_base
            parts[k].append(t)
        for k in gts:
            if k in e:
                gts[k].extend(e[k])
        if "indices" in e:
            if len(e["indices"][0]) > 0:
                singles.append(e["indices"][0].long() + box_base)
            seqs.extend(ix.long() + box_base for ix in e["indices"][1:])
        box_base += e["labels"].shape[0]
        frame_base += nf
        frames.append(nf)
    for k in keys_cat:
        if parts[k]:
            out[k] = torch.cat(parts[k], 0)
    for k, v in gts.items():
        if v:
            out[k] = v
    if singles or seqs:
        out["indices"] = [torch.cat(singles) if singles else torch.tensor([])] + seqs
    out["video_frames"] = np.asarray(frames, dtype=np.int64)
    out["video_size"] = entries[0].get("video_size")
    return out


def to_producer_contract(entry):
    """The entry dict as a B200-aware detector would hand it over (the reference produces it at
    tools/utils/object_detector.py:372-396 as fp32 NCHW): `union_feat` bf16 channels-last [N,7,7,1024] — the rows the
    union_func1 GEMM consumes, so the model runs no layout pass — and `spatial_masks` bf16 [N,2,27,27].  Half the bytes
    of the fp32 hand-off; every GEMM sees the same bf16 operands (the fp32 path rounds to bf16 at the same point)."""
    out = dict(entry)
    uf = entry["union_feat"]
    if uf.dtype != torch.bfloat16:
        if uf.is_cuda:
            out["union_feat"] = ops.nchw_to_nhwc_bf16(uf.contiguous()).view(uf.shape[0], 7, 7, 1024)
        else:
            out["union_feat"] = uf.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    out["spatial_masks"] = entry["spatial_masks"].to(torch.bfloat16).contiguous()
    return out


class _GemmNT(torch.autograd.Function):
    """y = a @ b^T through b200vsgg_gemm_bf16 (bf16 operands, fp32 out), differentiable in both."""

    @staticmethod
    def forward(ctx, a, b):
        ab, bb = ops.cast_bf16(a.contiguous()), ops.cast_bf16(b.contiguous())
        ctx.save_for_backward(ab, bb)
        out = torch.empty(a.shape[0], b.shape[0], device=a.device)
        _fgemm(ab, bb, out_f32=out)
        return out

    @staticmethod
    def backward(ctx, g):
        ab, bb = ctx.saved_tensors
        M, N = g.shape
        K = ab.shape[1]
        Np = (N + 7) // 8 * 8  # TMA row pitch must be a multiple of 16 bytes
        # (the cast kernel works on 4-column vectors: pad first when N is not a multiple of 8, e.g. the 26 memory slots)
        gb = ops.cast_bf16(F.pad(g, (0, Np - N)).contiguous() if Np != N else g.contiguous())
        da = db = None
        if ctx.needs_input_grad[0]:
            da = torch.empty(M, K, device=g.device)
            bpad = bb
            if Np != N:
                bpad = torch.zeros(Np, K, device=g.device, dtype=torch.bfloat16)
                bpad[:N] = bb
            ops.gemm(gb, bpad, b_mn=True, out_f32=da)
        if ctx.needs_input_grad[1]:
            dbp = torch.empty(Np, K, device=g.device)
            ops.gemm(gb, ab, a_mn=True, b_mn=True, out_f32=dbp)
            db = dbp[:N]
        return da, db


def _split_bf16(w):
    """[rows, k] fp32 -> [rows, 2k] bf16 = (hi | lo) with hi + lo == w to ~2^-17 relative."""
    hi = w.to(torch.bfloat16)
    lo = (w - hi.float()).to(torch.bfloat16)
    return torch.cat([hi, lo], 1).contiguous()


def _pos_rows_times_wt(pos, w_bf16):
    """[2, in] fp32 position rows times a bf16 [out, in] weight slice -> [2, out] fp32 on the tcgen05 GEMM; the rows
    enter as bf16 hi + lo parts (rows 0-1 / 2-3 of an 8-row operand), so only the weight is rounded."""
    hi = pos.to(torch.bfloat16)
    a = torch.zeros(8, pos.shape[1], device=pos.device, dtype=torch.bfloat16)
    a[0:2] = hi
    a[2:4] = (pos - hi.float()).to(torch.bfloat16)
    out = torch.empty(8, w_bf16.shape[0], device=pos.device, dtype=torch.float32)
    ops.gemm(a, w_bf16, out_f32=out, split_k=1)
    return (out[0:2] + out[2:4]).contiguous()


def _bn_train_stats(s1, s2, cnt, bn):
    """Per-video BatchNorm statistics from segmented column sums (s1 = sum x, s2 = sum x^2, [V,C]);
    returns (mean, rstd) and applies the V sequential running-statistics updates the reference would
    have made video by video (its batch IS one video; momentum 0.01)."""
    mean = s1 / cnt[:, None]
    var = (s2 / cnt[:, None] - mean * mean).clamp_min_(0.0)
    rstd = torch.rsqrt(var + bn.eps)
    with torch.no_grad():
        V = s1.shape[0]
        m = bn.momentum
        w = m * (1 - m) ** torch.arange(V - 1, -1, -1, device=s1.device, dtype=s1.dtype)
        unb = var * (cnt / (cnt - 1))[:, None]
        bn.running_mean.mul_((1 - m) ** V).add_((w[:, None] * mean).sum(0))
        bn.running_var.mul_((1 - m) ** V).add_((w[:, None] * unb).sum(0))
        bn.num_batches_tracked += V
    return mean, rstd


# ================================================================================================
# the model
# ================================================================================================
class TEMPURA(nn.Module):

    def __init__(self, mode="sgdet", attention_class_num=None, spatial_class_num=None, contact_class_num=None,
                 obj_classes=None, rel_classes=None, enc_layer_num=None, dec_layer_num=None, obj_mem_compute=None,
                 rel_mem_compute=None, mem_fusion=None, selection=None, selection_lambda=0.5, take_obj_mem_feat=False,
                 obj_head="gmm", rel_head="gmm", K=None, tracking=None, embed_vecs=None, consistency_regulariser=False):
        super().__init__()
        self.obj_classes, self.GMM_K, self.mem_fusion, self.rel_classes = obj_classes, K, mem_fusion, rel_classes
        self.attention_class_num, self.spatial_class_num, self.contact_class_num = (
            attention_class_num, spatial_class_num, contact_class_num)
        assert mode in ("sgdet", "sgcls", "predcls")
        if mode == "sgdet" or take_obj_mem_feat or rel_head != "gmm":
            raise NotImplementedError("b200vsgg.TEMPURA accelerates the PredCLS and SGCls-train / GMM-head paths "
                                      "(SURVEY.md §8); got mode=%s take_obj_mem_feat=%s rel_head=%s"
                                      % (mode, take_obj_mem_feat, rel_head))
        self.mode, self.tracking, self.take_obj_mem_feat = mode, tracking, take_obj_mem_feat
        self.obj_head, self.rel_head = obj_head, rel_head
        self.obj_mem_compute, self.rel_mem_compute = obj_mem_compute, rel_mem_compute
        self.selection_lambda = float(selection_lambda)
        self.rel_memory = []
        if embed_vecs is None:  # stand-in for GloVe-6B-200d (tools/utils/word_vectors.py), seeded
            embed_vecs = torch.randn(len(obj_classes), 200, generator=torch.Generator().manual_seed(len(obj_classes)))
        self.object_classifier = ObjectClassifier(mode=mode, obj_classes=obj_classes, obj_head=obj_head,
                                                  mem_compute=obj_mem_compute, K=K, selection=selection,
                                                  selection_lambda=self.selection_lambda, tracking=tracking,
                                                  embed_vecs=embed_vecs)
        self.union_func1 = nn.Conv2d(1024, 256, 1, 1)
        self.conv = nn.Sequential(
            nn.Conv2d(2, 256 // 2, kernel_size=7, stride=2, padding=3, bias=True), nn.ReLU(inplace=True),
            nn.BatchNorm2d(256 // 2, momentum=0.01), nn.MaxPool2d(kernel_size=3, stride=2, padding=1),
            nn.Conv2d(256 // 2, 256, kernel_size=3, stride=1, padding=1, bias=True), nn.ReLU(inplace=True),
            nn.BatchNorm2d(256, momentum=0.01))
        self.subj_fc = nn.Linear(2048, 512)
        self.obj_fc = nn.Linear(2048, 512)
        self.vr_fc = nn.Linear(256 * 7 * 7, 512)
        self.obj_embed = nn.Embedding(len(obj_classes), 200)
        self.obj_embed.weight.data = embed_vecs.clone()
        self.obj_embed2 = nn.Embedding(len(obj_classes), 200)
        self.obj_embed2.weight.data = embed_vecs.clone()
        self.glocal_transformer = _STTran(enc_layer_num, dec_layer_num, D_MODEL, N_HEADS, FFN_DIM, 0.1, "latter",
                                          rel_mem_compute, mem_fusion, selection, self.selection_lambda)
        self.a_rel_compress = GMMHead(D_MODEL, attention_class_num, "attention", K)
        self.s_rel_compress = GMMHead(D_MODEL, spatial_class_num, "spatial", K)
        self.c_rel_compress = GMMHead(D_MODEL, contact_class_num, "contact", K)
        # EXTENSION (not in the reference's lib/tempura.py, which never fills the *_temp_loss keys its trainer
        # reads, TEMPURA_train.py:215-218): the TEAT-GT regulariser R1-R3 on TEMPURA's graphs.  Adds parameters,
        # so it is opt-in and reference checkpoints still load strictly without it.
        self.consistency_regulariser = bool(consistency_regulariser)
        self._flags_pin = None
        if self.consistency_regulariser:
            from .regulariser import GraphTransformer
            self.gat = GraphTransformer(dim=10, depth=4)
            self.gat_semantic = GraphTransformer(dim=D_MODEL, depth=4)
            self.gate_nn = nn.Linear(10, 1)
            self.gate_sem_nn = nn.Linear(D_MODEL, 1)
        # knobs that are not part of the reference API
        self.dropout_p = 0.1            # nn.Dropout(0.1) / MHA dropout=0.1 everywhere in transformer.py
        self.gmm_eps = None             # dict head -> [K,N,C] noise to inject (parity tests); None = device RNG
        self.last_plan = None

    # ------------------------------------------------------------------------------------------
    def _path_params(self):
        g = self.glocal_transformer
        ps = [self.subj_fc.weight, self.subj_fc.bias, self.obj_fc.weight, self.obj_fc.bias, self.union_func1.weight,
              self.union_func1.bias, self.vr_fc.weight, self.vr_fc.bias, self.obj_embed.weight, self.obj_embed2.weight,
              g.position_embedding.weight, self.conv[0].weight, self.conv[0].bias, self.conv[2].weight,
              self.conv[2].bias, self.conv[4].weight, self.conv[4].bias, self.conv[6].weight, self.conv[6].bias]
        for l in g.local_attention.layers:
            ps += [l.self_attn.in_proj_weight, l.self_attn.in_proj_bias, l.self_attn.out_proj.weight,
                   l.self_attn.out_proj.bias, l.linear1.weight, l.linear1.bias, l.linear2.weight, l.linear2.bias,
                   l.norm1.weight, l.norm1.bias, l.norm2.weight, l.norm2.bias]
        for l in g.global_attention.layers:
            ps += [l.multihead2.in_proj_weight, l.multihead2.in_proj_bias, l.multihead2.out_proj.weight,
                   l.multihead2.out_proj.bias, l.linear1.weight, l.linear1.bias, l.linear2.weight, l.linear2.bias,
                   l.norm3.weight, l.norm3.bias]
        return ps

    def grad_layer_groups(self):
        """Parameter groups whose gradients the hand-written backward finishes together (one transformer layer's four
        weight matrices), last layer first = the order backward produces them: ddp.GradSync gives each group one flat
        buffer that the backward writes into and all-reduces it with ONE collective when the layer is done."""
        g = self.glocal_transformer
        groups = [[l.multihead2.in_proj_weight, l.multihead2.out_proj.weight, l.linear1.weight, l.linear2.weight]
                  for l in reversed(list(g.global_attention.layers))]
        groups += [[l.self_attn.in_proj_weight, l.self_attn.out_proj.weight, l.linear1.weight, l.linear2.weight]
                   for l in reversed(list(g.local_attention.layers))]
        return groups

    def _hallucinate(self, feat):
        """memory_hallucinator (tools/utils/transformer.py:143-175), joint memory, late fusion."""
        g = self.glocal_transformer
        if not (g.mem_compute and g.mem_fusion == "late") or len(self.rel_memory) == 0:
            return feat
        bank = torch.cat([v for _, v in self.rel_memory.items()], 0).to(feat.device, torch.float32)
        D = D_MODEL
        w = g.mem_attention.in_proj_weight
        q = _GemmNT.apply(feat, w[:D])
        k = _GemmNT.apply(bank, w[D:2 * D])
        v = _GemmNT.apply(bank, w[2 * D:])
        s = _GemmNT.apply(q, k) * (1.0 / math.sqrt(D))
        p = torch.softmax(s, -1)
        pad = (-p.shape[1]) % 8
        o = _GemmNT.apply(F.pad(p, (0, pad)), F.pad(v, (0, 0, 0, pad)).t().contiguous())
        mem = _GemmNT.apply(o, g.mem_attention.out_proj.weight)
        e = g.selector if g.selection == "manual" else g.selector(feat).sigmoid()
        return e * feat + (1 - e) * mem

    def _consistency_prepare(self, entry, plan):
        """Issue everything of the regulariser that depends on the boxes only (node layout, spatial edge
        predicates, their copy to pinned host memory) BEFORE the main path is launched, so that the host-side
        Laplacian eigen-decompositions later overlap the device's forward work."""
        from .teatgt import TeatPlan, edge_threshold
        dev = entry["features"].device
        pair_h = entry.get("pair_idx_host")
        if pair_h is None:
            pair_h = entry["pair_idx"].cpu().numpy()
        tp = TeatPlan(plan.counts_h, plan.frames_per_video, pair_h).to(dev)
        no_prev = torch.zeros_like(tp.has_prev)
        dummy = entry["boxes"].contiguous()
        sp, _ = ops.teat_pair_flags(dummy[:, :4].contiguous(), dummy, tp.feat_row, tp.node_off, no_prev,
                                    edge_threshold(entry["video_size"]), 2.0, tp.nmax)
        host = torch.empty(sp.shape, dtype=torch.uint8).pin_memory() if self._flags_pin is None or \
            self._flags_pin.shape != sp.shape else self._flags_pin
        self._flags_pin = host
        host.copy_(sp, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return tp, sp, (host, ev)

    def _consistency(self, entry, plan, rel_feats, prep, differentiable=False):
        """EXTENSION, see __init__: structure / semantic temporal-consistency losses (detached, like
        lib/teatgt.py:350-351) over 5-frame clips.  Graph nodes per frame = person + objects with spatial edges
        by box-centre distance (lib/teatgt.py:199-209); the semantic branch reads the clip's relation-feature
        rows [0:n_f] in place of TokenGT's hidden_x (same `savor` indexing, lib/teatgt.py:312-314)."""
        from .regulariser import consistency_losses
        tp, sp, flags_host = prep
        clips = tp.clip_of_frame
        pairs_pc = np.bincount(clips, weights=plan.counts_h.astype(np.float64), minlength=tp.n_clips).astype(np.int64)
        clip_pair_off = np.concatenate([[0], np.cumsum(pairs_pc)])
        entry["structure_temp_loss"], entry["semantic_temp_loss"] = consistency_losses(
            self.gat, self.gat_semantic, self.gate_nn, self.gate_sem_nn, tp, sp, rel_feats,
            clip_first_row=clip_pair_off[clips], clip_rows=pairs_pc[clips], flags_host=flags_host,
            differentiable=differentiable)

    # ------------------------------------------------------------------------------------------
    def forward(self, entry, phase="train", unc=False):
        self.object_classifier.gmm_eps = self.gmm_eps
        entry = self.object_classifier(entry, phase=phase, unc=unc)
        feats = entry["features"]
        if not feats.is_cuda:
            raise RuntimeError("b200vsgg.TEMPURA runs only on CUDA tensors (no CPU fallback for the hot path)")
        fpv = entry.get("video_frames")
        if fpv is None and "human_idx" in entry:
            fpv = np.asarray([entry["human_idx"].shape[0]])
        plan = entry.get("segment_plan")
        if plan is None:
            plan = plan_from_im_idx(entry["im_idx"], fpv, entry.get("frame_counts_host")).to(feats.device)
        self.last_plan = plan

        cons_prep = (self._consistency_prepare(entry, plan)
                     if self.consistency_regulariser and phase == "train" else None)
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self._path_params())
        runner = _PathRunner(self, entry, plan, train_dropout=self.training, save=need_grad)
        out, local = _PathFn.apply(runner, *self._path_params())
        mixed = self._hallucinate(out)
        g = self.glocal_transformer
        if g.mem_compute and g.mem_fusion == "late":
            rel_features, mem_features = out, mixed
        else:
            rel_features = mem_features = local
        entry["obj_class"] = entry["pred_labels"][entry["pair_idx"][:, 1]]
        entry["rel_features"] = rel_features
        entry["rel_mem_features"] = mem_features

        heads = [self.a_rel_compress, self.s_rel_compress, self.c_rel_compress]
        mode = 2 if unc else (1 if phase == "train" else 0)
        eps = self.gmm_eps or {}
        eps_list = [eps.get(n) for n in ("attention", "spatial", "contacting")]
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        res = apply_heads(heads, mixed, mode, eps_list, seed)
        if not unc:
            entry["attention_distribution"], entry["spatial_distribution"], entry["contacting_distribution"] = res[:3]
        else:
            (entry["attention_al_uc"], entry["spatial_al_uc"], entry["contacting_al_uc"],
             entry["attention_ep_uc"], entry["spatial_ep_uc"], entry["contacting_ep_uc"]) = res
        if cons_prep is not None:
            # LAST: the regulariser ends with a data-dependent filter (KL >= 0, lib/teatgt.py:327-333) that synchronises
            # the host; everything else of the forward is already queued behind it on the device by then
            diff = bool(getattr(self, "differentiable_consistency", False)) and torch.is_grad_enabled()
            self._consistency(entry, plan, mixed if diff else mixed.detach(), cons_prep, differentiable=diff)
        return entry


# ================================================================================================
# heads: packed GEMM + fused mixture epilogue
# ================================================================================================
DIRECT_HEAD_GRADS = True       # see _HeadsFn.backward


def apply_heads(heads, feat, mode, eps_list, seed, skip_first=False):
    """All mixture heads of `heads` (GMMHead containers) on `feat` [N, hid]: ONE packed GEMM + one epilogue kernel
    (tools/utils/gmm_heads.py:37-76 does 3K tiny Linears per head).  The 2 x 3K x len(heads) Linear parameters go to the
    autograd function individually — no differentiable torch.cat whose backward would split the packed gradient with
    one small kernel per Linear; the packed bf16 operand is memoised per parameter version."""
    lins = [l for h in heads for l in h.packed_order()]
    params = [l.weight for l in lins] + [l.bias for l in lins]
    direct = DIRECT_HEAD_GRADS and torch.is_grad_enabled() and all(p.is_leaf for p in params)
    anchor = next((i for i, p in enumerate(params) if p.requires_grad), 0)
    # skip_first (mode 0 only): the object head's test-phase output drops the background class (gmm_heads.py:63-64)
    softmaxes = [2 if (skip_first and mode == 0 and h.softmax) else h.softmax for h in heads]
    return _HeadsFn.apply(feat, mode, heads[0].k, [h.num_classes for h in heads], softmaxes, eps_list,
                          seed, len(lins), (tuple(params), anchor) if direct else None,
                          *(params[anchor:anchor + 1] if direct else params))


class _HeadsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, mode, K, Cs, softmaxes, eps_list, seed, n_lin, holder, *params):
        # holder: (the 2 * n_lin parameters as a plain tuple, index of the anchor) — no autograd edges; `params` is then only
        # the anchor (the first trainable parameter), which keeps the node connected to the graph — see backward
        direct = holder is not None
        ctx.anchor = 0
        if direct:
            params, ctx.anchor = holder
        ws, bs = params[:n_lin], params[n_lin:]
        N = feat.shape[0]
        K_in = ws[0].shape[1]
        cols = sum(w.shape[0] for w in ws)
        cols_pad = (cols + HEAD_COLS_PAD - 1) // HEAD_COLS_PAD * HEAD_COLS_PAD
        dev = feat.device

        # Split-precision product (SURVEY.md "fp32-accumulated logits ... max-abs <= 1e-3"): the head logits feed
        # softmax / sigmoid outputs that are compared at 1e-3, and a plain bf16 x bf16 product costs ~6e-4 of that
        # budget by itself (tools/parity_breakdown.py).  x = hi + lo, W = hi + lo, z = hi.hi + lo.hi + hi.lo as ONE
        # tcgen05 GEMM with K' = 3K over [hi|lo|hi] x [W_hi|W_hi|W_lo] — 0.3 % of the step's flops.
        def pack_w():
            Wf = torch.zeros(cols_pad, K_in, device=dev)
            Wf[:cols] = torch.cat([w.detach() for w in ws], 0)
            Wb = Wf.to(torch.bfloat16)
            return torch.cat([Wb, Wb, (Wf - Wb.float()).to(torch.bfloat16)], 1).contiguous()

        def pack_b():
            bias = torch.zeros(cols_pad, device=dev)
            bias[:cols] = torch.cat([b.detach() for b in bs])
            return bias

        Wb3 = ops.cached_weight("heads_w", ws, pack_w)
        bias = ops.cached_weight("heads_b", bs, pack_b)
        Wb = Wb3[:, :K_in]                                   # bf16(W): operand of the input-gradient GEMM
        fb3 = ops.split3_bf16(feat.contiguous())
        fb = fb3[:, :K_in]                                   # bf16(feat): operand of the weight-gradient GEMM
        z = torch.empty(N, cols_pad, device=dev)
        _fgemm(fb3, Wb3, bias=bias, out_f32=z)
        bases, b = [], 0
        for C in Cs:
            bases.append(b)
            b += K * (2 * C + 1)
        outs = [torch.empty(N, C - (1 if (mode == 0 and sm == 2) else 0), device=dev) for C, sm in zip(Cs, softmaxes)]
        outs2 = [torch.empty(N, C, device=dev) for C in Cs] if mode == 2 else [None] * len(Cs)
        eps_dev = [e.to(dev, torch.float32).contiguous() if e is not None else None for e in eps_list]
        specs = [dict(col_base=bases[i], num_classes=Cs[i], softmax=softmaxes[i], eps=eps_dev[i], out=outs[i],
                      out2=outs2[i]) for i in range(len(Cs))]
        ops.gmm_head_fwd(z, K, specs, mode, seed)
        ctx.save_for_backward(fb, Wb, z)
        ctx.params, ctx.direct = params, direct
        ctx.meta = (mode, K, Cs, softmaxes, eps_dev, seed, bases, cols, cols_pad, n_lin, [w.shape[0] for w in ws])
        if mode == 2:
            ctx.mark_non_differentiable(*outs, *outs2)
            return (*outs, *outs2)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        fb, Wb, z = ctx.saved_tensors
        mode, K, Cs, softmaxes, eps_dev, seed, bases, cols, cols_pad, n_lin, rows_of = ctx.meta
        N = fb.shape[0]
        dev = fb.device
        douts = [(g if g is not None else torch.zeros(N, C, device=dev)).contiguous().float()
                 for g, C in zip(grads, Cs)]
        specs = [dict(col_base=bases[i], num_classes=Cs[i], softmax=softmaxes[i], eps=eps_dev[i], dout=douts[i])
                 for i in range(len(Cs))]
        dz = torch.empty(N, cols_pad, device=dev, dtype=torch.bfloat16)
        ops.gmm_head_bwd(z, K, specs, mode, dz, seed)
        dfeat = None
        dws, dbs = [None] * n_lin, [None] * n_lin
        if ctx.needs_input_grad[0]:
            dfeat = torch.empty(N, Wb.shape[1], device=dev)
            ops.gemm(dz, Wb, b_mn=True, out_f32=dfeat)
        params, ctx.params = ctx.params, None
        need = [p.requires_grad for p in params] if ctx.direct else list(ctx.needs_input_grad[9:])
        if any(need[:n_lin]):
            dWp = torch.empty(cols_pad, Wb.shape[1], device=dev)
            ops.gemm(dz, fb, a_mn=True, b_mn=True, out_f32=dWp)
            dws = list(torch.split(dWp[:cols], rows_of, 0))              # views: no copies, no kernels
        if any(need[n_lin:]):
            dbp = torch.zeros(1, cols_pad, device=dev)
            ops.colsum(dz, dbp)
            dbs = list(torch.split(dbp[0, :cols], rows_of, 0))
        if ctx.direct:
            # 2 x 3K x 3 = 108 small Linear parameters: as autograd inputs they cost 108 AccumulateGrad node evaluations
            # (~0.4-0.5 ms of host time) exactly where the device has nothing queued yet (the forward ends with the
            # regulariser's host synchronisation, and the heads' backward kernels are tiny).  They are therefore NOT inputs
            # of this node (only the first weight is, to keep it in the graph); their gradients are accumulated here with
            # AccumulateGrad's semantics (adopt when .grad is None, add otherwise).  tempura.DIRECT_HEAD_GRADS = False
            # restores plain autograd inputs for wrappers that hook these parameters' graph nodes (torch DDP);
            # b200vsgg.ddp.GradSync reads .grad and is unaffected.
            with torch.no_grad():
                allg = list(dws) + list(dbs)
                for i, (p_, g_) in enumerate(zip(params, allg)):
                    if i == ctx.anchor or g_ is None or not need[i]:
                        continue
                    if p_.grad is None:
                        p_.grad = g_
                    else:
                        p_.grad.add_(g_)
            return (dfeat,) + (None,) * 8 + (allg[ctx.anchor] if need[ctx.anchor] else None,)
        return (dfeat, None, None, None, None, None, None, None, None, *dws, *dbs)


# ================================================================================================
# the pair-token + spatial/temporal transformer path (hand-orchestrated forward and backward)
# ================================================================================================
class _PathFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, runner, *params):
        out, local = runner.forward(params)
        ctx.runner = runner
        ctx.mark_non_differentiable(local)
        return out, local

    @staticmethod
    def backward(ctx, d_out, _d_local):
        runner = ctx.runner
        needs = ctx.needs_input_grad
        grads = runner.backward(d_out.contiguous(), need_params=needs[1:])
        ctx.runner = None
        return (None, *grads)


class _PathRunner:
    """Launch sequence of one forward (and its backward) over libb200vsgg kernels.  Holds the saved
    activations; one instance per model call."""

    def __init__(self, model, entry, plan, train_dropout, save):
        self.model, self.entry, self.plan = model, entry, plan
        self.p = model.dropout_p if train_dropout else 0.0
        self.save = save
        self.seed0 = int(torch.randint(0, 2 ** 40, (1,)).item()) if self.p > 0 else 0
        self.n_enc = len(model.glocal_transformer.local_attention.layers)
        self.n_dec = len(model.glocal_transformer.global_attention.layers)
        self.saved = {}

    def _seed(self, site):
        return self.seed0 + 7919 * site

    # -------------------------------------------------------------------------- parameter views
    def _unpack(self, params):
        it = iter(params)
        P = {}
        for n in ("subj_w", "subj_b", "obj_w", "obj_b", "union_w", "union_b", "vr_w", "vr_b", "emb1", "emb2", "pos",
                  "c1_w", "c1_b", "bn1_g", "bn1_b", "c2_w", "c2_b", "bn2_g", "bn2_b"):
            P[n] = next(it)
        P["enc"] = [dict(zip(("in_w", "in_b", "out_w", "out_b", "w1", "b1", "w2", "b2", "g1", "be1", "g2", "be2"),
                             [next(it) for _ in range(12)])) for _ in range(self.n_enc)]
        P["dec"] = [dict(zip(("in_w", "in_b", "out_w", "out_b", "w1", "b1", "w2", "b2", "g3", "be3"),
                             [next(it) for _ in range(10)])) for _ in range(self.n_dec)]
        return P

    @staticmethod
    def _bf(w):
        return ops.cast_bf16(w.detach().reshape(w.shape[0], -1).contiguous())

    # ------------------------------------------------------------------------------ forward
    # ---------------------------------------------------------------- spatial-mask branch (P4)
    def _bn_tables(self, y, rows_per_pair, bn, gamma, beta, tag):
        """(scale, shift) [V,C] of the BatchNorm that follows conv+ReLU output `y` (bf16 rows)."""
        plan, S = self.plan, self.saved
        C = y.shape[1]
        if self.model.training:
            s1 = torch.zeros(plan.V, C, device=y.device)
            s2 = torch.zeros(plan.V, C, device=y.device)
            ops.seg_colstats(y, plan.stat_chunks(rows_per_pair), s1, y, s2)
            cnt = plan.pairs_per_video_dev * float(rows_per_pair)
            mean, rstd = _bn_train_stats(s1, s2, cnt, bn)
        else:
            mean = bn.running_mean[None].expand(plan.V, C)
            rstd = torch.rsqrt(bn.running_var + bn.eps)[None].expand(plan.V, C)
            cnt = None
        scale = (gamma.detach()[None] * rstd).contiguous()
        shift = (beta.detach()[None] - mean * scale).contiguous()
        if self.save:
            S[tag] = (mean.contiguous(), rstd.contiguous(), cnt)
        return scale, shift

    def _mask_branch_fwd(self, P, W):
        """lib/tempura.py:466-474 channels-last; returns cm rows [N*49,256] bf16."""
        e, plan, S, model = self.entry, self.plan, self.saved, self.model
        dev = e["features"].device
        N = plan.N
        bf16 = torch.bfloat16
        # conv1 decides the max-pool argmax, a discrete choice: it is evaluated to ~fp32 accuracy (the
        # mask values are exact in bf16; the weights enter as bf16 hi + bf16 lo parts along K) and
        # pooled from an fp32 copy, so the routing equals the reference's except for fp32-level ties.
        A1 = torch.empty(N * 196, 128, device=dev, dtype=bf16)
        ops.mask_im2col(e["spatial_masks"].contiguous(), A1)
        y1 = torch.empty(N * 196, 128, device=dev, dtype=bf16)
        y1f = torch.empty(N * 196, 128, device=dev, dtype=torch.float32)
        _fgemm(A1, W["c1x"], bias=P["c1_b"].detach(), act=ops.ACT_RELU, out_bf16=y1, out_f32=y1f, a_k_period=128)
        sc1, sh1 = self._bn_tables(y1f, 196, model.conv[2], P["bn1_g"], P["bn1_b"], "bn1")
        z = torch.empty(N * 49, 128, device=dev, dtype=bf16)
        arg = torch.empty(N * 49, 128, device=dev, dtype=torch.uint8)
        ops.bn_pool_fwd(y1f, sc1, sh1, plan.video_of_pair32, N, 14, 128, z, arg)
        del y1f
        A2 = torch.empty(N * 49, 1152, device=dev, dtype=bf16)
        ops.im2col3x3(z, N, 7, 128, A2)
        y2 = torch.empty(N * 49, 256, device=dev, dtype=bf16)
        _fgemm(A2, W["c2"], bias=P["c2_b"].detach(), act=ops.ACT_RELU, out_bf16=y2)
        sc2, sh2 = self._bn_tables(y2, 49, model.conv[6], P["bn2_g"], P["bn2_b"], "bn2")
        cm = torch.empty(N * 49, 256, device=dev, dtype=bf16)
        ops.seg_affine(None, y2, None, sc2, sh2, plan.video_of_pair32, 49, cm)
        if self.save:
            S.update(A1=A1, y1=y1, arg=arg, A2=A2, y2=y2)
        return cm

    def _bn_relu_bwd(self, dout, y, rows_per_pair, gamma, tag, G, gkey, bkey):
        """Backward of BN(ReLU-output y) followed by the ReLU mask: returns d(conv output) bf16."""
        plan = self.plan
        mean, rstd, cnt = self.saved[tag]
        C = y.shape[1]
        s1 = torch.zeros(plan.V, C, device=y.device)
        s2 = torch.zeros(plan.V, C, device=y.device)
        ops.seg_colstats(dout, plan.stat_chunks(rows_per_pair), s1, y, s2)
        sx = (s2 - mean * s1) * rstd                     # sum dout * xhat per (video, channel)
        G[gkey], G[bkey] = sx.sum(0), s1.sum(0)
        g = gamma.detach()[None]
        k1 = (g * rstd).expand(plan.V, C).contiguous()
        if cnt is not None:                                  # batch statistics (train mode)
            inv_n = (1.0 / cnt)[:, None]
            k2 = (-g * rstd * rstd * sx * inv_n).contiguous()
            k3 = (-g * rstd * s1 * inv_n - k2 * mean).contiguous()
        else:                                                # running statistics: BN is a fixed affine map
            k2 = torch.zeros_like(k1)
            k3 = torch.zeros_like(k1)
        dconv = torch.empty_like(y)
        ops.seg_affine(dout, y, k1, k2, k3, plan.video_of_pair32, rows_per_pair, dconv, relu_mask=True)
        return dconv

    def _mask_branch_bwd(self, dcm, P, W, G):
        """dcm: bf16 [N*49,256] gradient of the branch output; fills conv / BatchNorm parameter grads."""
        S, plan = self.saved, self.plan
        N = plan.N
        dev = dcm.device
        bf16, f32 = torch.bfloat16, torch.float32
        d2 = self._bn_relu_bwd(dcm, S["y2"], 49, P["bn2_g"], "bn2", G, "bn2_g", "bn2_b")
        gw2 = torch.empty(256, 1152, device=dev, dtype=f32)
        ops.gemm(d2, S["A2"], a_mn=True, b_mn=True, out_f32=gw2)
        G["c2_w"] = gw2.view(256, 3, 3, 128).permute(0, 3, 1, 2)
        gb2 = torch.zeros(1, 256, device=dev)
        ops.colsum(d2, gb2)
        G["c2_b"] = gb2[0]
        dA2 = torch.empty(N * 49, 1152, device=dev, dtype=bf16)
        ops.gemm(d2, W["c2"], b_mn=True, out_bf16=dA2)
        dz = torch.empty(N * 49, 128, device=dev, dtype=bf16)
        ops.col2im3x3(dA2, N, 7, 128, dz)
        del dA2
        dpool = torch.empty(N * 196, 128, device=dev, dtype=bf16)
        ops.pool_bwd(dz, S["arg"], N, 14, 128, dpool)
        d1 = self._bn_relu_bwd(dpool, S["y1"], 196, P["bn1_g"], "bn1", G, "bn1_g", "bn1_b")
        gw1 = torch.empty(128, 128, device=dev, dtype=f32)
        ops.gemm(d1, S["A1"], a_mn=True, b_mn=True, out_f32=gw1)
        G["c1_w"] = gw1[:, :98].reshape(128, 2, 7, 7)
        gb1 = torch.zeros(1, 128, device=dev)
        ops.colsum(d1, gb1)
        G["c1_b"] = gb1[0]

    def forward(self, params):
        P = self._unpack(params)
        e, plan, S = self.entry, self.plan, self.saved
        dev = e["features"].device
        N, M2, p = plan.N, plan.M2, self.p
        bf16, f32 = torch.bfloat16, torch.float32
        new = lambda r, c, dt: torch.empty(r, c, device=dev, dtype=dt)

        # ---- weights in bf16 (K-major [out, in]); the same copies serve dgrad via MN-major B.  Memoised per parameter
        #      version (ops.cached_weight): cast once per optimiser step, not once per forward
        cw = ops.cached_weight
        W = {"so": cw("so", (P["subj_w"], P["obj_w"]), lambda: self._bf(torch.cat([P["subj_w"], P["obj_w"]], 0))),
             "union": cw("union", (P["union_w"],), lambda: self._bf(P["union_w"])),
             # vr_fc consumes vr in (h, w, c) order instead of the reference's (c, h, w): permute columns
             "vr": cw("vr", (P["vr_w"],), lambda: self._bf(P["vr_w"].detach().view(512, 256, 49).permute(0, 2, 1).reshape(512, 12544))),
             # conv1 taps in (c, kh, kw) order, weights as bf16 hi | lo halves; conv2 taps in (kh, kw, c) order
             "c1x": cw("c1x", (P["c1_w"],), lambda: _split_bf16(F.pad(P["c1_w"].detach().reshape(128, 98), (0, 30)))),
             "c2": cw("c2", (P["c2_w"],), lambda: self._bf(P["c2_w"].detach().permute(0, 2, 3, 1).reshape(256, 1152)))}
        b_so = cw("b_so", (P["subj_b"], P["obj_b"]), lambda: torch.cat([P["subj_b"], P["obj_b"]]).detach().clone())
        for i, L in enumerate(P["enc"] + P["dec"]):
            for n in ("in_w", "out_w", "w1", "w2"):
                W["%d%s" % (i, n)] = cw("layer", (L[n],), lambda w=L[n]: self._bf(w))
        S["W"] = W

        # ---- P1/P2: subj_fc|obj_fc over all boxes, then gather  (lib/tempura.py:537-544)
        featb = ops.cast_bf16(e["features"].contiguous())
        so = new(featb.shape[0], 1024, f32)
        _fgemm(featb, W["so"], bias=b_so, out_f32=so)
        # ---- P3: union_func1 as GEMM over NHWC rows, mask branch added in the epilogue (:548)
        uf = e["union_feat"]
        if uf.dtype == torch.bfloat16:
            # producer-side hand-off ((f).4): the ROIAlign already emitted bf16 channels-last rows [N,7,7,1024] /
            # [49N,1024] — exactly the A operand of the union_func1 GEMM: no layout pass, half the H2D bytes
            assert uf.is_contiguous() and uf.numel() == N * 49 * 1024 and uf.shape[-1] == 1024, \
                "bf16 union_feat must be channels-last [N,7,7,1024] or [49N,1024]"
            ub = uf.view(N * 49, 1024)
        else:
            ub = ops.nchw_to_nhwc_bf16(uf.contiguous())
        cm_rows = self._mask_branch_fwd(P, W)
        vrp = new(N * 49, 256, bf16)
        _fgemm(ub, W["union"], bias=P["union_b"].detach(), residual=cm_rows, out_bf16=vrp)
        del cm_rows
        # ---- P5: vr_fc straight into the token buffer (:549)
        tok = new(N, D_MODEL, f32)
        tokb = new(N, D_MODEL, bf16)
        _fgemm(vrp.view(N, 12544), W["vr"], bias=P["vr_b"].detach(), out_f32=tok[:, 1024:1536])
        # ---- P6: gather + label embeddings + concat (:554-563)
        ops.pair_concat_fwd(so, e["pair_idx"].contiguous(), e["pred_labels"].contiguous(), P["emb1"].detach().contiguous(),
                            P["emb2"].detach().contiguous(), tok, tokb)
        if self.save:
            S.update(featb=featb, ub=ub, vrp=vrp)

        # ---- T2: spatial encoder, sequences = frames (transformer.py:5-30,195)
        x32, xb = tok, tokb
        site = 0
        for i, L in enumerate(P["enc"]):
            qkv = new(N, 3 * D_MODEL, bf16)
            _fgemm(xb, W["%din_w" % i], bias=L["in_b"].detach(), out_bf16=qkv)
            ctxb = new(N, D_MODEL, bf16)
            ops.attn_small_fwd(qkv[:, :D_MODEL], qkv[:, D_MODEL:2 * D_MODEL], qkv[:, 2 * D_MODEL:], plan.frame_off,
                               plan.F, plan.max_frame_len, N_HEADS, HEAD_DIM, ctxb, p, self._seed(site))
            u = new(N, D_MODEL, f32)
            _fgemm(ctxb, W["%dout_w" % i], bias=L["out_b"].detach(), residual=x32, out_f32=u, dropout_p=p,
                     seed=self._seed(site + 1))
            t32, tb = new(N, D_MODEL, f32), new(N, D_MODEL, bf16)
            m1, r1 = torch.empty(N, device=dev), torch.empty(N, device=dev)
            ops.layernorm_fwd(u, L["g1"].detach(), L["be1"].detach(), 1e-5, t32, tb, mean=m1, rstd=r1)
            h = new(N, FFN_DIM, bf16)
            _fgemm(tb, W["%dw1" % i], bias=L["b1"].detach(), act=ops.ACT_RELU, out_bf16=h, dropout_p=p,
                     seed=self._seed(site + 2))
            v = new(N, D_MODEL, f32)
            _fgemm(h, W["%dw2" % i], bias=L["b2"].detach(), residual=t32, out_f32=v, dropout_p=p,
                     seed=self._seed(site + 3))
            y32, yb = new(N, D_MODEL, f32), new(N, D_MODEL, bf16)
            m2, r2 = torch.empty(N, device=dev), torch.empty(N, device=dev)
            ops.layernorm_fwd(v, L["g2"].detach(), L["be2"].detach(), 1e-5, y32, yb, mean=m2, rstd=r2)
            if self.save:
                S["enc%d" % i] = dict(xb=xb, qkv=qkv, ctxb=ctxb, u=u, m1=m1, r1=r1, tb=tb, h=h, v=v, m2=m2, r2=r2,
                                      site=site)
            x32, xb = y32, yb
            site += 4
        local = x32

        # ---- T3: 2-frame windows + position embedding (transformer.py:203-215)
        pos = P["pos"].detach().contiguous()
        D2 = 2 * D_MODEL
        on_pairs = DEC_FIRST_ON_PAIRS and self.n_dec > 0
        latter_only = DEC_LATTER_ONLY and self.n_dec > 0
        S["dec_flags"] = (on_pairs, latter_only)
        g32 = new(M2, D_MODEL, f32)
        gb = gpb = None
        if on_pairs:
            ops.gather_rows(local, plan.win_src, out_f32=g32)
        else:
            gb, gpb = new(M2, D_MODEL, bf16), new(M2, D_MODEL, bf16)
            ops.gather_rows(local, plan.win_src, add_table=pos, add_idx=plan.win_pos, out_f32=g32, out_bf16=gb,
                            out_bf16_added=gpb)
        # ---- T4: temporal decoder, sequences = windows (transformer.py:33-58,220)
        for j, L in enumerate(P["dec"]):
            i = self.n_enc + j
            first, last = on_pairs and j == 0, latter_only and j + 1 == self.n_dec
            w_in = W["%din_w" % i]
            qkv = new(M2, 3 * D_MODEL, bf16)
            in_b = L["in_b"].detach()
            if first:
                # window token = pair row (+ position row): project the N pair rows, gather into the windows; the
                # position term enters as pos @ Wqk^T (2 rows, memoised per optimiser step) and q, k are rounded once
                qk_n, v_n = new(N, D2, f32), new(N, D_MODEL, bf16)
                _fgemm(xb, w_in[:D2], bias=in_b[:D2], out_f32=qk_n)
                _fgemm(xb, w_in[D2:], bias=in_b[D2:], out_bf16=v_n)
                pos_qk = cw("pos_qk", (L["in_w"], P["pos"]), lambda: _pos_rows_times_wt(pos, w_in[:D2]))
                ops.gather_rows(qk_n, plan.win_src, add_table=pos_qk, add_idx=plan.win_pos, out_bf16_added=qkv[:, :D2])
                ops.gather_rows_bf16(v_n, plan.win_src, qkv[:, D2:])
                del qk_n, v_n
            else:
                _fgemm(gpb, w_in[:D2], bias=in_b[:D2], out_bf16=qkv[:, :D2])
                _fgemm(gb, w_in[D2:], bias=in_b[D2:], out_bf16=qkv[:, D2:])
            ctxb = new(M2, D_MODEL, bf16)
            ops.attn_small_fwd(qkv[:, :D_MODEL], qkv[:, D_MODEL:D2], qkv[:, D2:], plan.win_off,
                               plan.W, plan.max_win_len, N_HEADS, HEAD_DIM, ctxb, p, self._seed(site))
            rows, res = M2, g32
            if last:
                # only the rows the 'latter' read-out keeps go through out-proj / LayerNorm / FFN
                rows = N
                ctx_all, ctxb, res = ctxb, new(N, D_MODEL, bf16), new(N, D_MODEL, f32)
                ops.gather_rows_bf16(ctx_all, plan.latter_src, ctxb)
                ops.gather_rows(g32, plan.latter_src, out_f32=res)
                del ctx_all
            u = new(rows, D_MODEL, f32)
            _fgemm(ctxb, W["%dout_w" % i], bias=L["out_b"].detach(), residual=res, out_f32=u, dropout_p=p,
                     seed=self._seed(site + 1))
            t32, tb = new(rows, D_MODEL, f32), new(rows, D_MODEL, bf16)
            m3, r3 = torch.empty(rows, device=dev), torch.empty(rows, device=dev)
            ops.layernorm_fwd(u, L["g3"].detach(), L["be3"].detach(), 1e-5, t32, tb, mean=m3, rstd=r3)
            h = new(rows, FFN_DIM, bf16)
            _fgemm(tb, W["%dw1" % i], bias=L["b1"].detach(), act=ops.ACT_RELU, out_bf16=h, dropout_p=p,
                     seed=self._seed(site + 2))
            y32 = new(rows, D_MODEL, f32)
            _fgemm(h, W["%dw2" % i], bias=L["b2"].detach(), residual=t32, out_f32=y32, dropout_p=p,
                     seed=self._seed(site + 3))
            if self.save:
                S["dec%d" % j] = dict(gb=gb, gpb=gpb, xb=xb if first else None, qkv=qkv, ctxb=ctxb, u=u, m3=m3, r3=r3,
                                      tb=tb, h=h, site=site)
            g32 = y32
            if j + 1 < self.n_dec:
                gb, gpb = new(M2, D_MODEL, bf16), new(M2, D_MODEL, bf16)
                ops.gather_rows(g32, None, rows=M2, add_table=pos, add_idx=plan.win_pos, out_bf16=gb, out_bf16_added=gpb)
            site += 4
        # ---- T5: 'latter' scatter-back as a gather (transformer.py:236-242)
        if latter_only:
            out = g32                 # the last layer already ran on the kept rows, in pair order
        else:
            out = new(N, D_MODEL, f32)
            ops.gather_rows(g32, plan.latter_src, out_f32=out)
        return out, local

    # ------------------------------------------------------------------------------ backward
    def backward(self, d_out, need_params):
        S, plan, e, model = self.saved, self.plan, self.entry, self.model
        P = self._unpack(self.model._path_params())
        W = S["W"]
        dev = d_out.device
        N, M2, p = plan.N, plan.M2, self.p
        bf16, f32 = torch.bfloat16, torch.float32
        new = lambda r, c, dt: torch.empty(r, c, device=dev, dtype=dt)
        # every small zero-initialised accumulator of this backward (bias / LayerNorm / embedding gradients, ~60 of
        # them) is a 256-byte aligned slice of ONE buffer cleared by one fill
        pool = torch.zeros(1 << 18, device=dev, dtype=f32)
        pool_pos = [0]

        def zeros(*shape):
            n = 1
            for d in shape:
                n *= d
            a = pool_pos[0]
            if n > (1 << 16) or a + n > pool.numel():
                return torch.zeros(*shape, device=dev, dtype=f32)
            pool_pos[0] = a + (n + 63) // 64 * 64
            return pool[a:a + n].view(*shape)

        G = {}  # grads by the names of _unpack
        hook = getattr(model, "_grad_ready_hook", None)   # ddp.GradSync: start all-reducing finished gradients
        galloc = getattr(model, "_grad_alloc", None)      # ddp.GradSync: persistent per-layer gradient buffers

        def ready(*tensors):
            if hook is not None:
                hook(tensors)

        def gnew(param, r, c):
            """fp32 [r, c] gradient buffer of `param`: its view inside the layer's all-reduce bucket when data
            parallel (autograd adopts it as .grad — no copies around the collective), else a fresh tensor."""
            v = galloc(param) if galloc is not None else None
            return v if v is not None else new(r, c, f32)

        def ffn_and_norm_bwd(dy32, R, L, i, rows, norm_g, norm_mean, norm_rstd, pre_norm, gkey, bkey, site):
            """Backward of  y = t + drop(W2 drop(relu(W1 t + b1)) + b2),  t = LN(pre_norm).
            Returns (d_pre_norm fp32, bf16(dropout_mask * d_pre_norm)) and fills weight grads."""
            dyb = ops.cast_bf16(dy32, drop_p=p, seed=self._seed(site + 3))
            gL = G.setdefault(i, {})
            gL["w2"] = gnew(L["w2"], D_MODEL, FFN_DIM)
            ops.gemm(dyb, R["h"], a_mn=True, b_mn=True, out_f32=gL["w2"])
            gL["b2"] = zeros(1, D_MODEL)
            ops.colsum(dyb, gL["b2"])
            dz = new(rows, FFN_DIM, bf16)
            ops.gemm(dyb, W["%dw2" % i], b_mn=True, mask_src=R["h"], mask_mode=ops.MASK_RELU,
                     alpha=(1.0 / (1.0 - p)) if p > 0 else 1.0, out_bf16=dz)
            gL["w1"] = gnew(L["w1"], FFN_DIM, D_MODEL)
            ops.gemm(dz, R["tb"], a_mn=True, b_mn=True, out_f32=gL["w1"])
            gL["b1"] = zeros(1, FFN_DIM)
            ops.colsum(dz, gL["b1"])
            dt = new(rows, D_MODEL, f32)
            ops.gemm(dz, W["%dw1" % i], b_mn=True, residual=dy32, out_f32=dt)
            du = new(rows, D_MODEL, f32)
            dub = new(rows, D_MODEL, bf16)
            gL[gkey], gL[bkey] = zeros(D_MODEL), zeros(D_MODEL)
            ops.layernorm_bwd(dt, pre_norm, L[norm_g].detach(), norm_mean, norm_rstd, du, dub, p, self._seed(site + 1),
                              gL[gkey], gL[bkey])
            return du, dub

        def attn_block_bwd(du, dub, R, L, i, rows, seg_off, n_seg, max_len, site, pos_groups=False, expand=None):
            """Backward of u = x + drop(Wo attn(q,k,v) + bo); returns dqkv bf16 [rows, 3D].  `expand` (int32 [rows], -1 =
            none): du / dub / R["ctxb"] hold only the kept rows of a pruned last layer; the context gradient is spread
            back over all `rows` attention rows (zero where the output was dropped)."""
            gL = G[i]
            gL["out_w"] = gnew(L["out_w"], D_MODEL, D_MODEL)
            ops.gemm(dub, R["ctxb"], a_mn=True, b_mn=True, out_f32=gL["out_w"])
            gL["out_b"] = zeros(1, D_MODEL)
            ops.colsum(dub, gL["out_b"])
            dctx = new(dub.shape[0], D_MODEL, bf16)
            ops.gemm(dub, W["%dout_w" % i], b_mn=True, out_bf16=dctx)
            if expand is not None:
                dctx_kept, dctx = dctx, new(rows, D_MODEL, bf16)
                ops.gather_rows_bf16(dctx_kept, expand, dctx)
                del dctx_kept
            qkv = R["qkv"]
            dqkv = new(rows, 3 * D_MODEL, bf16)
            ops.attn_small_bwd(qkv[:, :D_MODEL], qkv[:, D_MODEL:2 * D_MODEL], qkv[:, 2 * D_MODEL:], dctx, seg_off, n_seg,
                               max_len, N_HEADS, HEAD_DIM, dqkv[:, :D_MODEL], dqkv[:, D_MODEL:2 * D_MODEL],
                               dqkv[:, 2 * D_MODEL:], p, self._seed(site))
            if not pos_groups:
                gL["in_b"] = zeros(1, 3 * D_MODEL)
                ops.colsum(dqkv, gL["in_b"])
            return dqkv

        # ---- T5 backward: scatter d_out into window-token rows (a pruned last layer consumes d_out as it is)
        on_pairs, latter_only = S["dec_flags"]
        D2 = 2 * D_MODEL
        if latter_only:
            dy = d_out.contiguous()
        else:
            dy = new(M2, D_MODEL, f32)
            ops.gather2_sum_rows(d_out, plan.inv_latter2, out_f32=dy)
        dlocal = None
        # ---- T4 backward
        G["pos"] = zeros(2, D_MODEL)
        for j in reversed(range(self.n_dec)):
            i = self.n_enc + j
            L, R = P["dec"][j], S["dec%d" % j]
            site = R["site"]
            first, last = on_pairs and j == 0, latter_only and j + 1 == self.n_dec
            du, dub = ffn_and_norm_bwd(dy, R, L, i, N if last else M2, "g3", R["m3"], R["r3"], R["u"], "g3", "be3", site)
            dqkv = attn_block_bwd(du, dub, R, L, i, M2, plan.win_off, plan.W, plan.max_win_len, site, pos_groups=True,
                                  expand=plan.inv_latter if last else None)
            if last:        # residual path of the kept rows, back in window-token order
                du_kept, du = du, new(M2, D_MODEL, f32)
                ops.gather2_sum_rows(du_kept, plan.inv_latter2, out_f32=du)
                del du_kept
            gL = G[i]
            gL["in_w"] = gnew(L["in_w"], 3 * D_MODEL, D_MODEL)
            w_in = W["%din_w" % i]
            # ONE pass over dqkv gives its column sums per position id: their total is the in_proj bias gradient, the
            # [dq|dk] part per position feeds the position embedding: d pos[k] = (sum over tokens with position k) @ W_qk
            by_pos = zeros(2, 3 * D_MODEL)
            ops.colsum(dqkv, by_pos, plan.win_pos, 2)
            gL["in_b"] = by_pos.sum(0, keepdim=True)
            dpos_qk = by_pos[:, :D2]
            dposb = torch.zeros(8, D2, device=dev, dtype=bf16)
            dposb[:2] = dpos_qk
            dpos_l = new(8, D_MODEL, f32)
            ops.gemm(dposb, w_in[:D2], b_mn=True, out_f32=dpos_l)
            G["pos"] += dpos_l[:2]
            if first:
                # the projections ran on the N pair rows: sum the <= 2 window copies of every pair row first
                dqkv_n, du_n = new(N, 3 * D_MODEL, bf16), new(N, D_MODEL, f32)
                ops.gather2_sum_rows_bf16(dqkv, plan.pair_win2, dqkv_n)
                ops.gather2_sum_rows(du, plan.pair_win2, out_f32=du_n)
                ops.gemm(dqkv_n, R["xb"], a_mn=True, b_mn=True, out_f32=gL["in_w"])
                dlocal = new(N, D_MODEL, f32)
                ops.gemm(dqkv_n, w_in, b_mn=True, residual=du_n, out_f32=dlocal)
            else:
                ops.gemm(dqkv[:, :D2], R["gpb"], a_mn=True, b_mn=True, out_f32=gL["in_w"][:D2])
                ops.gemm(dqkv[:, D2:], R["gb"], a_mn=True, b_mn=True, out_f32=gL["in_w"][D2:])
                dx = new(M2, D_MODEL, f32)
                ops.gemm(dqkv, w_in, b_mn=True, residual=du, out_f32=dx)
                dy = dx
            ready(gL["in_w"], gL["out_w"], gL["w1"], gL["w2"])
        # ---- T3 backward: each pair row was read by <= 2 windows
        if dlocal is None:
            dlocal = new(N, D_MODEL, f32)
            ops.gather2_sum_rows(dy, plan.pair_win2, out_f32=dlocal)
        # ---- T2 backward
        dy = dlocal
        for i in reversed(range(self.n_enc)):
            L, R = P["enc"][i], S["enc%d" % i]
            site = R["site"]
            # y = LN2(v): first undo norm2
            G.setdefault(i, {})
            dv = new(N, D_MODEL, f32)
            G[i]["g2"], G[i]["be2"] = zeros(D_MODEL), zeros(D_MODEL)
            ops.layernorm_bwd(dy, R["v"], L["g2"].detach(), R["m2"], R["r2"], dv, None, 0.0, 0, G[i]["g2"], G[i]["be2"])
            du, dub = ffn_and_norm_bwd(dv, R, L, i, N, "g1", R["m1"], R["r1"], R["u"], "g1", "be1", site)
            dqkv = attn_block_bwd(du, dub, R, L, i, N, plan.frame_off, plan.F, plan.max_frame_len, site)
            G[i]["in_w"] = gnew(L["in_w"], 3 * D_MODEL, D_MODEL)
            ops.gemm(dqkv, R["xb"], a_mn=True, b_mn=True, out_f32=G[i]["in_w"])
            dx = new(N, D_MODEL, f32)
            ops.gemm(dqkv, W["%din_w" % i], b_mn=True, residual=du, out_f32=dx)
            dy = dx
            ready(G[i]["in_w"], G[i]["out_w"], G[i]["w1"], G[i]["w2"])
        dtok = dy
        # ---- P6/P1 backward
        O = S["featb"].shape[0]
        dso = zeros(O, 1024)
        G["emb1"], G["emb2"] = zeros(*P["emb1"].shape), zeros(*P["emb2"].shape)
        ops.pair_concat_bwd(dtok, e["pair_idx"].contiguous(), e["pred_labels"].contiguous(), dso, G["emb1"], G["emb2"])
        dsob = ops.cast_bf16(dso)
        gso = new(1024, 2048, f32)
        ops.gemm(dsob, S["featb"], a_mn=True, b_mn=True, out_f32=gso)
        bso = zeros(1, 1024)
        ops.colsum(dsob, bso)
        G["subj_w"], G["obj_w"], G["subj_b"], G["obj_b"] = gso[:512], gso[512:], bso[0, :512], bso[0, 512:]
        ready(gso)
        # ---- P5 backward
        dvr = ops.cast_bf16(dtok[:, 1024:1536])
        gvr = new(512, 12544, f32)
        ops.gemm(dvr, S["vrp"].view(N, 12544), a_mn=True, b_mn=True, out_f32=gvr)
        G["vr_w"] = gvr.view(512, 49, 256).permute(0, 2, 1).reshape(512, 12544)
        G["vr_b"] = zeros(1, 512)
        ops.colsum(dvr, G["vr_b"])
        dvrpb = new(N, 12544, bf16)
        ops.gemm(dvr, W["vr"], b_mn=True, out_bf16=dvrpb)
        # ---- P3 backward (union_feat itself is a frozen detector output: no dgrad)
        G["union_w"] = new(256, 1024, f32)
        ops.gemm(dvrpb.view(N * 49, 256), S["ub"], a_mn=True, b_mn=True, out_f32=G["union_w"])
        G["union_b"] = zeros(1, 256)
        ops.colsum(dvrpb.view(N * 49, 256), G["union_b"])
        # ---- P4 backward (the masks are inputs: no dgrad below conv1)
        self._mask_branch_bwd(dvrpb.view(N * 49, 256), P, W, G)

        # ---- assemble in _path_params order
        out = [G["subj_w"], G["subj_b"], G["obj_w"], G["obj_b"], G["union_w"].view(256, 1024, 1, 1), G["union_b"][0],
               G["vr_w"], G["vr_b"][0], G["emb1"], G["emb2"], G["pos"], G["c1_w"], G["c1_b"], G["bn1_g"], G["bn1_b"],
               G["c2_w"], G["c2_b"], G["bn2_g"], G["bn2_b"]]
        for i in range(self.n_enc):
            g = G[i]
            out += [g["in_w"], g["in_b"][0], g["out_w"], g["out_b"][0], g["w1"], g["b1"][0], g["w2"], g["b2"][0],
                    g["g1"], g["be1"], g["g2"], g["be2"]]
        for j in range(self.n_dec):
            g = G[self.n_enc + j]
            out += [g["in_w"], g["in_b"][0], g["out_w"], g["out_b"][0], g["w1"], g["b1"][0], g["w2"], g["b2"][0],
                    g["g3"], g["be3"]]
        out = [g if need else None for g, need in zip(out, need_params)]
        self.saved = {}
        flush = getattr(model, "_grad_flush_hook", None)
        if flush is not None:
            flush()
        return out


class _RelLossFn(torch.autograd.Function):
    """The three relation losses and their gradients from ONE kernel launch (b200vsgg_rel_loss)."""

    @staticmethod
    def forward(ctx, dist_a, dist_s, dist_c, att, spa, con, w):
        want = any(t.requires_grad for t in (dist_a, dist_s, dist_c))
        losses, da, ds, dc = ops.rel_loss(dist_a.contiguous().float(), dist_s.contiguous().float(),
                                          dist_c.contiguous().float(), att.contiguous(), spa, con, w.contiguous(), want)
        ctx.grads = (da, ds, dc)
        return losses[0], losses[1], losses[2]

    @staticmethod
    def backward(ctx, ga, gs, gc):
        da, ds, dc = ctx.grads
        return da * ga, ds * gs, dc * gc, None, None, None, None


class _ContrastiveFn(torch.autograd.Function):
    """pytorch_metric_learning ContrastiveLoss(pos_margin=0, neg_margin=1) per video, loss and gradient in one launch."""

    @staticmethod
    def forward(ctx, x, label, seg_off, max_rows):
        loss, dx = ops.contrastive_loss(x.contiguous().float(), label, seg_off, max_rows, 0.0, 1.0, x.requires_grad)
        ctx.dx, ctx.seg_off = dx, seg_off
        return loss

    @staticmethod
    def backward(ctx, g):
        rows = torch.repeat_interleave(g, (ctx.seg_off[1:] - ctx.seg_off[:-1]).long())
        return ctx.dx * rows[:, None], None, None, None


def contrastive_relation_losses(pred, plan, spatial_label, contact_label, weight=0.2):
    """`--use_ctl_loss` of the trainers (TEMPURA_train.py:209-212, TEATGT_train.py:176-179):
    0.2 * ContrastiveLoss(spatial_distribution, argmax(spatial_label, 1)) and the same for contacting; spatial_label /
    contact_label are the trainer's multi-hot matrices [N,6] / [N,17] (or int class indices [N]).  With a batch of videos
    each video is one ContrastiveLoss call (pairs never cross a video) and the results are averaged."""
    dev = pred["spatial_distribution"].device
    if plan is not None and plan.V > 1:
        off = np.asarray(plan.pair_off_video_h, dtype=np.int32)
    else:
        off = np.asarray([0, pred["spatial_distribution"].shape[0]], dtype=np.int32)
    seg_off = ops.upload(off, dev)
    max_rows = int(np.diff(off).max())
    out = {}
    for name, dist, lab in (("spatial_con_loss", pred["spatial_distribution"], spatial_label),
                            ("contact_con_loss", pred["contacting_distribution"], contact_label)):
        idx = (lab.argmax(1) if lab.dim() == 2 else lab).to(torch.int32).contiguous()
        out[name] = weight * _ContrastiveFn.apply(dist, idx, seg_off, max_rows).mean()
    return out


def gt_label_csr(entry, device, num_classes=None):
    """Ragged predicate labels as the dataloader yields them (lists of class ids per pair) -> what the loss kernel
    consumes: attention class index int64 [N] and CSR (offsets int32 [N+1], ids int32) for spatial / contacting.
    Replaces the per-pair Python loop that builds multi-hot matrices (TEMPURA_train.py:181-187).
    num_classes = (attention, spatial, contacting) class counts: labels are validated on the host like the
    reference path does implicitly (nn.CrossEntropyLoss / the index assignment into the multi-hot matrix raise on
    out-of-range ids) — the kernel only clamps for memory safety."""
    import itertools
    for a in entry["attention_gt"]:
        if isinstance(a, (list, tuple)) and len(a) != 1:
            raise ValueError("attention_gt must hold exactly one class per pair (TEMPURA_train.py:183), got %r" % (a,))
    att = np.fromiter((a[0] if isinstance(a, (list, tuple)) else int(a) for a in entry["attention_gt"]), dtype=np.int64)

    def check(ids, C, what):
        if num_classes is not None and ids.size and (ids.min() < 0 or ids.max() >= C):
            raise IndexError("%s label out of range: ids span [%d, %d], the head has %d classes"
                             % (what, int(ids.min()), int(ids.max()), C))

    def csr(lists, C, what):
        lens = np.fromiter((len(l) for l in lists), dtype=np.int64, count=len(lists))
        off = np.zeros(len(lists) + 1, dtype=np.int32)
        off[1:] = np.cumsum(lens)
        idx = np.fromiter(itertools.chain.from_iterable(lists), dtype=np.int32, count=int(off[-1]))
        check(idx, C, what)
        return ops.upload(off, device), ops.upload(idx, device)

    ca, cs, cc = num_classes if num_classes is not None else (0, 0, 0)
    check(att, ca, "attention")
    return ops.upload(att, device), csr(entry["spatial_gt"], cs, "spatial"), csr(entry["contacting_gt"], cc, "contacting")


def tempura_loss(pred, plan=None, eos_coef=1.0):
    """The reference trainer's losses (TEMPURA_train.py:181-206) on the model output dict; with a batch of
    videos each loss is the mean over videos of the per-video mean, i.e. exactly the average of the losses
    the reference would compute video by video.  SGCls outputs (the object branch ran) add `object_loss`
    (:191-195, class-weighted CE with weight[0] = eos_coef)."""
    dist_a, dist_s, dist_c = pred["attention_distribution"], pred["spatial_distribution"], pred["contacting_distribution"]
    dev = dist_a.device
    N = dist_a.shape[0]
    if "gt_tensors" in pred:  # label tensors prepared by the data loader (same values as below)
        att, spa, con = pred["gt_tensors"]
    elif dist_a.is_cuda:      # ragged label lists go to the loss kernel as CSR, no multi-hot matrices
        att, spa, con = gt_label_csr(pred, dev, (dist_a.shape[1], dist_s.shape[1], dist_c.shape[1]))
    else:
        att = torch.tensor([a[0] if isinstance(a, (list, tuple)) else int(a) for a in pred["attention_gt"]], device=dev)
        spa = torch.zeros(N, dist_s.shape[1])
        con = torch.zeros(N, dist_c.shape[1])
        for i in range(N):
            spa[i, pred["spatial_gt"][i]] = 1
            con[i, pred["contacting_gt"][i]] = 1
        spa, con = spa.to(dev), con.to(dev)
    if plan is None or plan.V == 1:
        w = torch.full((N,), 1.0 / N, device=dev)
    else:
        ppv = plan.pairs_per_video_dev
        w = 1.0 / (ppv[plan.video_of_pair] * plan.V)
    if dist_a.is_cuda:
        la, ls, lc = _RelLossFn.apply(dist_a, dist_s, dist_c, att, spa, con, w)
    else:
        la = (F.cross_entropy(dist_a, att, reduction="none") * w).sum()
        ls = (F.binary_cross_entropy(dist_s, spa, reduction="none").mean(1) * w).sum()
        lc = (F.binary_cross_entropy(dist_c, con, reduction="none").mean(1) * w).sum()
    losses = {}
    if "box_groups" in pred:
        from .objbranch import object_loss
        grp = pred["box_groups"]
        losses["object_loss"] = object_loss(pred, eos_coef, grp.count if grp.V > 1 else None, grp.video_of_box64)
    losses.update({"attention_relation_loss": la, "spatial_relation_loss": ls, "contacting_relation_loss": lc})
    return losses
