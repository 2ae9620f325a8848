"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

CPU (plain torch, fp32) restatement of the reference's TEMPURA PredCLS relation path, used only as
the parity checker (tests/, __graft_entry__.smoke()) and as the timed CPU baseline
(bench.py cpu_baseline / --impl reference).  The product path (b200vsgg.*) never imports this.

Parity status: PINNED.  oracle/make_golden.py imports the *unmodified* reference classes
(lib/tempura.py::TEMPURA, tools/utils/transformer.py, tools/utils/gmm_heads.py) in the build
container, loads identical weights into this restatement, asserts agreement (max-abs <= 2e-5) and
writes the reference's outputs to tests/golden/tempura_*.pt; tests/test_oracle_golden.py re-checks
this file against those vectors on any machine.

What is restated, with the reference lines each piece follows (paths relative to the reference root):
  GMMHeadOracle      tools/utils/gmm_heads.py:3-76
  STTranOracle       tools/utils/transformer.py:5-58 (layers), :104-253 (windows, scatter-back, memory)
  TempuraOracle      lib/tempura.py:465-510 (layer definitions), :537-596 (forward)
  ObjectClassifierOracle  lib/tempura.py:51-255 (SGCls-train object branch: feature build, class-sequence
                     encoder, intermediate + head), get_sequence = tools/utils/ds_track.py:18-39; pinned by
                     oracle/make_golden_sgcls.py the same way (tests/golden/sgcls_*.pt)
  tempura_losses     TEMPURA_train.py:181-206
The state_dict key names equal the reference's so checkpoints interchange (strict=True).

Differences in *how* (not what): padding is built with index tensors instead of per-frame Python
loops; the temporal key mask is structural (padded slots) rather than `row-sum == 0`
(transformer.py:217) — identical unless a real token sums to exactly 0.0.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

D_MODEL = 1936
N_HEADS = 8


# --------------------------------------------------------------------------------------------
# attention building blocks
# --------------------------------------------------------------------------------------------
class PackedMHA(nn.Module):
    """Parameter container with nn.MultiheadAttention's names (in_proj_weight/bias, out_proj.*)."""

    def __init__(self, dim, heads, bias=True):
        super().__init__()
        self.dim, self.heads = dim, heads
        self.in_proj_weight = nn.Parameter(torch.empty(3 * dim, dim))
        nn.init.xavier_uniform_(self.in_proj_weight)
        if bias:
            self.in_proj_bias = nn.Parameter(torch.zeros(3 * dim))
        else:
            self.register_parameter("in_proj_bias", None)
        self.out_proj = nn.Linear(dim, dim, bias=bias)

    def attend(self, q_in, k_in, v_in, key_pad=None, drop_p=0.0, training=False):
        """Batch-first scaled-dot-product attention.  q_in [B,Lq,D]; k_in, v_in [B,Lk,D];
        key_pad [B,Lk] bool (True = ignore).  Returns [B,Lq,D]."""
        D, H = self.dim, self.heads
        hd = D // H
        w, b = self.in_proj_weight, self.in_proj_bias
        bq = bk = bv = None
        if b is not None:
            bq, bk, bv = b[:D], b[D:2 * D], b[2 * D:]
        q = F.linear(q_in, w[:D], bq)
        k = F.linear(k_in, w[D:2 * D], bk)
        v = F.linear(v_in, w[2 * D:], bv)
        B, Lq, _ = q.shape
        Lk = k.shape[1]
        q = q.view(B, Lq, H, hd).transpose(1, 2) * (1.0 / math.sqrt(hd))
        k = k.view(B, Lk, H, hd).transpose(1, 2)
        v = v.view(B, Lk, H, hd).transpose(1, 2)
        s = q @ k.transpose(-1, -2)
        if key_pad is not None:
            s = s.masked_fill(key_pad[:, None, None, :], float("-inf"))
        p = torch.softmax(s, dim=-1)
        if drop_p > 0.0 and training:
            p = F.dropout(p, drop_p, True)
        ctx = (p @ v).transpose(1, 2).reshape(B, Lq, D)
        return self.out_proj(ctx)


class SpatialLayer(nn.Module):
    """Post-LN encoder layer (transformer.py:5-30)."""

    def __init__(self, dim, heads, ffn, p):
        super().__init__()
        self.self_attn = PackedMHA(dim, heads)
        self.linear1 = nn.Linear(dim, ffn)
        self.linear2 = nn.Linear(ffn, dim)
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.p = p

    def forward(self, x, key_pad):
        a = self.self_attn.attend(x, x, x, key_pad, self.p, self.training)
        x = self.norm1(x + F.dropout(a, self.p, self.training))
        h = F.dropout(F.relu(self.linear1(x)), self.p, self.training)
        x = self.norm2(x + F.dropout(self.linear2(h), self.p, self.training))
        return x


class TemporalLayer(nn.Module):
    """Decoder layer: q = k = x + pos, v = x; LN after attention only (transformer.py:33-58)."""

    def __init__(self, dim, heads, ffn, p):
        super().__init__()
        self.multihead2 = PackedMHA(dim, heads)
        self.linear1 = nn.Linear(dim, ffn)
        self.linear2 = nn.Linear(ffn, dim)
        self.norm3 = nn.LayerNorm(dim)
        self.p = p

    def forward(self, x, key_pad, pos):
        qk = x + pos
        a = self.multihead2.attend(qk, qk, x, key_pad, self.p, self.training)
        t = self.norm3(x + F.dropout(a, self.p, self.training))
        h = F.dropout(F.relu(self.linear1(t)), self.p, self.training)
        return t + F.dropout(self.linear2(h), self.p, self.training)


class LayerStack(nn.Module):
    def __init__(self, make_layer, n):
        super().__init__()
        first = make_layer()
        layers = [first]
        for _ in range(n - 1):  # reference deep-copies one layer (transformer.py:255-256): identical init
            twin = make_layer()
            twin.load_state_dict(first.state_dict())
            layers.append(twin)
        self.layers = nn.ModuleList(layers[:n])


def frame_segments(im_idx):
    """im_idx (sorted frame id per pair, float or int) -> (counts[F], offsets[F+1]) as int64."""
    ids = im_idx.to(torch.int64)
    F_ = int(ids[-1].item()) + 1
    counts = torch.bincount(ids, minlength=F_)
    offsets = torch.zeros(F_ + 1, dtype=torch.int64, device=ids.device)
    offsets[1:] = torch.cumsum(counts, 0)
    return counts, offsets


class STTranOracle(nn.Module):
    """Spatial encoder -> 2-frame sliding-window temporal decoder -> 'latter' scatter-back ->
    optional memory hallucinator (transformer.py:104-253)."""

    def __init__(self, enc_layer_num=1, dec_layer_num=3, embed_dim=D_MODEL, nhead=N_HEADS, dim_feedforward=2048,
                 dropout=0.1, mode="latter", mem_compute=True, mem_fusion=None, selection=None, selection_lambda=0.5):
        super().__init__()
        assert mode == "latter"  # the only mode TEMPURA constructs (lib/tempura.py:498)
        self.mode, self.mem_compute, self.mem_fusion, self.selection = mode, mem_compute, mem_fusion, selection
        self.local_attention = LayerStack(lambda: SpatialLayer(embed_dim, nhead, dim_feedforward, dropout),
                                          enc_layer_num)
        if mem_compute:
            assert mem_compute != "seperate", "only the default joint memory is restated"
            self.mem_attention = PackedMHA(embed_dim, 1, bias=False)
            if selection == "manual":
                self.selector = float(selection_lambda)
            else:
                self.selector = nn.Linear(embed_dim, 1)
        self.global_attention = LayerStack(lambda: TemporalLayer(embed_dim, nhead, dim_feedforward, dropout),
                                           dec_layer_num)
        self.position_embedding = nn.Embedding(2, embed_dim)
        nn.init.uniform_(self.position_embedding.weight)

    def hallucinate(self, memory, feat):
        if len(memory) == 0:
            return feat
        e = self.selector if self.selection == "manual" else self.selector(feat).sigmoid()
        bank = torch.cat([v for _, v in memory.items()], 0)  # [3+6+17, D]
        q = feat[None]            # one "batch", N queries
        kv = bank[None]
        mem = self.mem_attention.attend(q, kv, kv)[0]
        return e * feat + (1 - e) * mem

    def forward(self, features, im_idx, memory=()):
        x = features
        N, D = x.shape
        counts, off = frame_segments(im_idx)
        Fn = counts.numel()
        l = int(counts.max())
        dev = x.device
        ar = torch.arange(l, device=dev)

        # ---- spatial: one sequence per frame, padded to l ----
        valid = ar[None, :] < counts[:, None]                       # [F,l]
        rows = (off[:-1, None] + ar[None, :]).clamp(max=N - 1)
        xp = x[rows] * valid[..., None]
        for layer in self.local_attention.layers:
            xp = layer(xp, ~valid)
        local = xp[valid]                                           # back to [N,D] in pair order

        # ---- temporal: window j = frames (j, j+1) = rows [off[j], off[j+2]) ----
        ar2 = torch.arange(2 * l, device=dev)
        wlen = counts[:-1] + counts[1:]                              # [F-1]
        wvalid = ar2[None, :] < wlen[:, None]
        wrows = (off[:-2, None] + ar2[None, :]).clamp(max=N - 1)
        g = local[wrows] * wvalid[..., None]
        second = (ar2[None, :] >= counts[:-1, None]) & wvalid        # slots holding frame j+1
        pos = torch.zeros(Fn - 1, 2 * l, D, device=dev, dtype=x.dtype)
        pos[wvalid & ~second] = self.position_embedding.weight[0]
        pos[second] = self.position_embedding.weight[1]
        for layer in self.global_attention.layers:
            g = layer(g, ~wvalid, pos)

        # ---- 'latter': frame 0 from window 0, frame j+1 from window j ----
        out = torch.zeros_like(x)
        first0 = ar2 < counts[0]
        out[: int(counts[0])] = g[0][first0]
        out[int(counts[0]):] = g[second]

        local_output, mem_feats = local, local
        if self.mem_compute and self.mem_fusion == "late":
            local_output = out
            out = self.hallucinate(memory, out)
            mem_feats = out
        return out, local_output, mem_feats


# --------------------------------------------------------------------------------------------
# GMM heads
# --------------------------------------------------------------------------------------------
class GMMHeadOracle(nn.Module):
    """K-component mixture head (tools/utils/gmm_heads.py:3-76)."""

    def __init__(self, hid_dim, num_classes, rel_type=None, k=4):
        super().__init__()
        self.k, self.num_classes, self.rel_type = k, num_classes, rel_type
        self.heads = nn.ModuleDict()
        for i in range(1, k + 1):
            self.heads["mu_%d" % i] = nn.Linear(hid_dim, num_classes)
            self.heads["pi_%d" % i] = nn.Linear(hid_dim, 1)
            self.heads["var_%d" % i] = nn.Linear(hid_dim, num_classes)
        self.softmax_act = rel_type == "attention" or rel_type is None

    def act(self, z):
        return torch.softmax(z, -1) if self.softmax_act else torch.sigmoid(z)

    def forward(self, x, phase="train", unc=False, eps=None):
        """eps: optional [K,N,C] noise to inject; default draws K CPU tensors like gmm_heads.py:57."""
        K = self.k
        mu = torch.stack([self.heads["mu_%d" % i](x) for i in range(1, K + 1)])            # [K,N,C]
        var = torch.stack([self.heads["var_%d" % i](x) for i in range(1, K + 1)]).sigmoid()
        pi = torch.softmax(torch.cat([self.heads["pi_%d" % i](x) for i in range(1, K + 1)], 1), 1)  # [N,K]
        pik = pi.t()[..., None]                                                                  # [K,N,1]
        if unc:
            prob = self.act(mu)
            mean = (prob * pik).sum(0)
            return (var * pik).sum(0), (((prob - mean) ** 2) * pik).sum(0)
        if eps is None:
            eps = torch.stack([torch.randn(var.shape[1:]) for _ in range(K)]).to(x.device)  # CPU RNG (quirk)
        if phase == "train":
            z = mu + var.sqrt() * eps
        elif self.rel_type is not None:
            z = mu
        else:
            z = mu[..., 1:]
        return (self.act(z) * pik).sum(0)


# --------------------------------------------------------------------------------------------
# the PredCLS model
# --------------------------------------------------------------------------------------------
def center_size(boxes):
    """(x1,y1,x2,y2) -> (cx,cy,w,h).  tools/utils/fpn/box_utils.py is ABSENT from the reference tree
    (imported at lib/tempura.py:18); this is the neural-motifs definition that file is copied from
    upstream.  UNPINNED for these three lines: the golden generator injects the same function."""
    wh = boxes[:, 2:] - boxes[:, :2] + 1.0
    return torch.cat((boxes[:, :2] + 0.5 * wh, wh), 1)


def sinusoid_table(d_model, max_len):
    """PositionalEncoding.pe (lib/tempura.py:31-36)."""
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(1, max_len, d_model)
    pe[0, :, 0::2] = torch.sin(position * div_term)
    pe[0, :, 1::2] = torch.cos(position * div_term)
    return pe


def get_sequence(entry, task="sgcls"):
    """tools/utils/ds_track.py:18-39: boxes grouped by the detector's arg-max class; entry['indices'][0] =
    all single-box groups concatenated, entry['indices'][1:] = the longer groups in class order."""
    if task == "predcls":
        return
    indices = [[]]
    pred = torch.argmax(entry["distribution"], 1)
    for c in pred.unique():
        index = torch.where(pred == c)[0]
        (indices[0] if len(index) == 1 else indices).append(index)
    indices[0] = torch.cat(indices[0]) if len(indices[0]) > 0 else torch.tensor([])
    entry["indices"] = indices


class _PE(nn.Module):
    def __init__(self, d_model, max_len):
        super().__init__()
        self.register_buffer("pe", sinusoid_table(d_model, max_len))


class ObjectClassifierOracle(nn.Module):
    """ObjectClassifier (lib/tempura.py:51-255; TEAT-GT's copy tools/utils/object_classifier.py:43-233 is
    line-for-line the same arithmetic).  Restated: PredCLS (:245-247), the SGCls feature build (:249-252),
    `classify` (:185-241) with and without the class-sequence encoder (`tracking`), for phase='train' and for
    the classification part of phase='test' (the relabel / NMS / ROIAlign tail :257-307 needs the reference's
    absent CUDA ops and is out of scope)."""

    def __init__(self, num_classes, mode="predcls", obj_head="gmm", K=4, mem_compute=None, selection=None,
                 selection_lambda=0.5, tracking=None, dropout=0.1):
        super().__init__()
        self.mode, self.obj_head, self.tracking, self.mem_compute, self.selection = (
            mode, obj_head, tracking, mem_compute, selection)
        self.obj_memory, self.p = [], dropout
        self.obj_embed = nn.Embedding(num_classes - 1, 200)
        self.pos_embed = nn.Sequential(nn.BatchNorm1d(4, momentum=0.001), nn.Linear(4, 128), nn.ReLU(inplace=True),
                                       nn.Dropout(dropout))
        d_model = 2048 + 200 + 128
        mem_embed = 1024
        if tracking:
            self.positional_encoder = _PE(d_model, 600 if mode == "sgdet" else 400)
            self.encoder_tran = LayerStack(lambda: SpatialLayer(d_model, 8, 1024, dropout), 3)
            mem_embed = d_model
        if mem_compute:
            self.mem_attention = PackedMHA(mem_embed, 1, bias=False)
            if selection == "manual":
                self.selector = selection_lambda
            else:
                self.selector = nn.Linear(1024, 1)
        self.intermediate = nn.Sequential(nn.Linear(d_model, 1024), nn.BatchNorm1d(1024), nn.ReLU())
        if obj_head == "gmm":
            self.decoder_lin = GMMHeadOracle(1024, num_classes, None, K)
        else:
            self.decoder_lin = nn.Sequential(nn.Linear(1024, num_classes))

    def hallucinate(self, feat):
        """lib/tempura.py:165-182."""
        if len(self.obj_memory) == 0:
            return feat
        e = self.selector if self.selection == "manual" else self.selector(feat).sigmoid()
        mem = self.mem_attention.attend(feat[None], self.obj_memory[None], self.obj_memory[None])[0]
        return e * feat + (1 - e) * mem if e is not None else feat + mem

    def encode_sequences(self, entry, x):
        """lib/tempura.py:186-210: every class sequence is one padded batch row; position = rank of the box's
        frame among the sequence's frames; single-box sequences run as length-1 batch rows at position 0."""
        indices = entry["indices"]
        pe = self.positional_encoder.pe[0]
        final = torch.zeros_like(x)
        seqs = [ix.long() for ix in indices[1:]]
        if seqs:
            L = max(len(ix) for ix in seqs)
            S = len(seqs)
            pad = torch.zeros(S, L, x.shape[1], dtype=x.dtype)
            pos = torch.zeros(S, L, dtype=torch.int64)
            mask = torch.ones(S, L, dtype=torch.bool)
            for s, ix in enumerate(seqs):
                _, counts = torch.unique(entry["boxes"][ix][:, 0].view(-1), return_counts=True, sorted=True)
                pos[s, :len(ix)] = torch.repeat_interleave(torch.arange(len(counts)), counts)
                pad[s, :len(ix)] = x[ix]
                mask[s, :len(ix)] = False
            h = F.dropout(pad + pe[pos], self.p, self.training)
            for layer in self.encoder_tran.layers:
                h = layer(h, mask)
            final[torch.cat(seqs)] = torch.cat([h[s, :len(ix)] for s, ix in enumerate(seqs)])
        if len(indices[0]) > 0:
            ix = indices[0].long()
            h = F.dropout(x[ix].unsqueeze(1) + pe[None, :1], self.p, self.training)
            for layer in self.encoder_tran.layers:
                h = layer(h, None)
            final[ix] = h[:, 0]
        return final

    def classify(self, entry, x, phase, unc, eps=None):
        if self.tracking:
            x = self.encode_sequences(entry, x)
            entry["object_features"] = x
            if self.mem_compute:
                x = self.hallucinate(x)
            entry["object_mem_features"] = x
            x = self.intermediate(x)
        else:
            x = self.intermediate(x)
            entry["object_features"] = x
            if self.mem_compute:
                x = self.hallucinate(x)
            entry["object_mem_features"] = x
        gmm = self.obj_head == "gmm"
        if phase == "train":
            if not gmm:
                entry["distribution"] = self.decoder_lin(x)
            elif not unc:
                entry["distribution"] = self.decoder_lin(x, phase, False, eps)
            else:
                entry["distribution"] = self.decoder_lin(x, "test", False)
                entry["obj_al_uc"], entry["obj_ep_uc"] = self.decoder_lin(x, unc=True)
            entry["pred_labels"] = entry["labels"]
        elif gmm:
            entry["distribution"] = self.decoder_lin(x, phase, unc, eps)
        else:
            entry["distribution"] = torch.softmax(self.decoder_lin(x)[:, 1:], dim=1)
        return entry

    def forward(self, entry, phase="train", unc=False, eps=None):
        if self.mode == "predcls":
            entry["pred_labels"] = entry["labels"]
            return entry
        assert self.mode == "sgcls"
        obj_embed = entry["distribution"] @ self.obj_embed.weight
        pos_embed = self.pos_embed(center_size(entry["boxes"][:, 1:]))
        x = torch.cat((entry["features"], obj_embed, pos_embed), 1)
        return self.classify(entry, x, phase, unc, eps)


class TempuraOracle(nn.Module):
    def __init__(self, mode="predcls", attention_class_num=3, spatial_class_num=6, contact_class_num=17,
                 obj_classes=None, rel_classes=None, enc_layer_num=1, dec_layer_num=3, obj_mem_compute=None,
                 rel_mem_compute=None, mem_fusion=None, selection=None, selection_lambda=0.5,
                 take_obj_mem_feat=False, obj_head="gmm", rel_head="gmm", K=6, tracking=None, dropout=0.1):
        super().__init__()
        assert mode in ("predcls", "sgcls") and not take_obj_mem_feat and rel_head == "gmm"
        self.mode, self.obj_classes, self.rel_memory = mode, obj_classes, []
        self.attention_class_num, self.spatial_class_num, self.contact_class_num = (
            attention_class_num, spatial_class_num, contact_class_num)
        ncls = len(obj_classes)
        self.object_classifier = ObjectClassifierOracle(ncls, mode, obj_head, K, obj_mem_compute, selection,
                                                        selection_lambda, tracking, dropout)
        self.union_func1 = nn.Conv2d(1024, 256, kernel_size=1)
        self.conv = nn.Sequential(
            nn.Conv2d(2, 128, kernel_size=7, stride=2, padding=3), nn.ReLU(inplace=True),
            nn.BatchNorm2d(128, momentum=0.01), nn.MaxPool2d(kernel_size=3, stride=2, padding=1),
            nn.Conv2d(128, 256, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=True),
            nn.BatchNorm2d(256, momentum=0.01))
        self.subj_fc = nn.Linear(2048, 512)
        self.obj_fc = nn.Linear(2048, 512)
        self.vr_fc = nn.Linear(256 * 7 * 7, 512)
        self.obj_embed = nn.Embedding(ncls, 200)
        self.obj_embed2 = nn.Embedding(ncls, 200)
        self.glocal_transformer = STTranOracle(enc_layer_num, dec_layer_num, D_MODEL, N_HEADS, 2048, dropout, "latter",
                                               rel_mem_compute, mem_fusion, selection, selection_lambda)
        self.a_rel_compress = GMMHeadOracle(D_MODEL, attention_class_num, "attention", K)
        self.s_rel_compress = GMMHeadOracle(D_MODEL, spatial_class_num, "spatial", K)
        self.c_rel_compress = GMMHeadOracle(D_MODEL, contact_class_num, "contact", K)

    def pair_tokens(self, entry):
        """lib/tempura.py:537-563 — [N,1936] = subj 512 | obj 512 | union+mask 512 | 2 x label-embedding 200."""
        pi = entry["pair_idx"]
        feats = entry["features"]
        s = self.subj_fc(feats[pi[:, 0]])
        o = self.obj_fc(feats[pi[:, 1]])
        vr = self.union_func1(entry["union_feat"]) + self.conv(entry["spatial_masks"])
        vr = self.vr_fc(vr.reshape(-1, 256 * 7 * 7))
        lab = entry["pred_labels"]
        return torch.cat([s, o, vr, self.obj_embed(lab[pi[:, 0]]), self.obj_embed2(lab[pi[:, 1]])], 1)

    def forward(self, entry, phase="train", unc=False, eps=None):
        e = eps or {}
        assert self.mode == "predcls" or phase == "train", "SGCls test-time relabel/NMS tail is not restated"
        entry = self.object_classifier(entry, phase, unc, e.get("object"))
        tok = self.pair_tokens(entry)
        out, rel_feats, mem_feats = self.glocal_transformer(tok, entry["im_idx"], self.rel_memory)
        entry["obj_class"] = entry["pred_labels"][entry["pair_idx"][:, 1]]
        entry["rel_features"] = rel_feats
        entry["rel_mem_features"] = mem_feats
        entry["global_output"] = out  # oracle-only key, used by parity tests
        if not unc:
            entry["attention_distribution"] = self.a_rel_compress(out, phase, False, e.get("attention"))
            entry["spatial_distribution"] = self.s_rel_compress(out, phase, False, e.get("spatial"))
            entry["contacting_distribution"] = self.c_rel_compress(out, phase, False, e.get("contacting"))
        else:
            entry["attention_al_uc"], entry["attention_ep_uc"] = self.a_rel_compress(out, phase, True)
            entry["spatial_al_uc"], entry["spatial_ep_uc"] = self.s_rel_compress(out, phase, True)
            entry["contacting_al_uc"], entry["contacting_ep_uc"] = self.c_rel_compress(out, phase, True)
        return entry


def object_loss(pred, eos_coef=1.0):
    """TEMPURA_train.py:97-100,191-195 / TEATGT_train.py:163-165: class-weighted CrossEntropyLoss
    (weight[0] = eos_coef, reduction='none') on `distribution` exactly as the model returns it (for the GMM
    head that is already a probability mixture), then a plain mean over the boxes."""
    dist = pred["distribution"]
    w = torch.ones(dist.shape[1], device=dist.device)
    w[0] = eos_coef
    return F.cross_entropy(dist, pred["labels"], weight=w, reduction="none").mean()


def tempura_losses(pred, attention_label, spatial_label, contact_label):
    """TEMPURA_train.py:196-205: CrossEntropyLoss applied to the (already soft-maxed) attention
    distribution, BCELoss on the two sigmoid mixtures, each reduced by mean."""
    return {
        "attention_relation_loss": F.cross_entropy(pred["attention_distribution"], attention_label),
        "spatial_relation_loss": F.binary_cross_entropy(pred["spatial_distribution"], spatial_label),
        "contacting_relation_loss": F.binary_cross_entropy(pred["contacting_distribution"], contact_label),
    }
