"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

CPU (plain torch fp32 + numpy fp64 eigh) restatement of the reference's TEAT-GT PredCLS classifier path
(rows G0-G10 of SURVEY.md §8a), used only as the parity checker (tests/, smoke) and CPU baseline.

Parity status: PINNED for the classifier path.  oracle/make_golden_teatgt.py runs the UNMODIFIED
reference `lib/teatgt.py::TEAT_GT.forward(phase='test')` (with the unmodified TokenGT modules of
tools/TokenGT/tokengt) in the build container behind the stand-ins of oracle/ref_shims.py, checks this
restatement against it (distributions <= 2e-5, edge_index / edge_data / eigenvectors bit-exact) and
writes the reference's outputs to tests/golden/teatgt_*.pt.
UNPINNED: the train-only consistency regulariser (rows R1-R3) — it runs through
`graph_transformer_pytorch` and `dgl`, which are absent from the reference tree and whose versions the
reference does not record; `regulariser()` below restates their published algorithms.

Reference lines followed (paths relative to the reference root):
  node tokens / ordering / clips      lib/teatgt.py:104-169
  pseudo-graph (spatial + temporal)   lib/teatgt.py:174-240
  Laplacian eigenvectors              lib/teatgt.py:243-254
  tokenizer                           tools/TokenGT/tokengt/modules/tokenizer.py:217-295
  encoder layer (pre-LN)              modules/tokengt_graph_encoder_layer.py:170-191, multihead_attention.py:135-183,
                                      feedforward.py:31-36
  head                                models/tokengt.py:99-134
  output split                        lib/teatgt.py:336-348
State-dict names equal the reference's (TokenGT_encoder.* and TokenGT_model.encoder.* alias one module).
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

CLIP_SIZE = 5
SPATIAL_THR = 0.5
SIM_THR = 0.75


# --------------------------------------------------------------------------------------------
# graph artefacts (integer parity contract)
# --------------------------------------------------------------------------------------------
def node_layout(entry):
    """Per-node gather recipe in the reference's token order (per frame: person, then objects in pair
    order).  Returns dict of int64 tensors: feat_row (row of entry['features'] / boxes / labels),
    is_person, frame, and obj_pair (pair row for object nodes, -1 for persons)."""
    im = entry["im_idx"].to(torch.int64)
    pair = entry["pair_idx"]
    N = im.numel()
    F_ = int(im.max()) + 1
    counts = torch.bincount(im, minlength=F_)
    off = torch.zeros(F_ + 1, dtype=torch.int64)
    off[1:] = torch.cumsum(counts, 0)
    feat_row, is_person, frame, obj_pair = [], [], [], []
    for f in range(F_):
        if counts[f] == 0:
            continue
        p0 = int(off[f])
        feat_row.append(int(pair[p0, 0])); is_person.append(1); frame.append(f); obj_pair.append(-1)
        for p in range(p0, int(off[f + 1])):
            feat_row.append(int(pair[p, 1])); is_person.append(0); frame.append(f); obj_pair.append(p)
    t = lambda x: torch.tensor(x, dtype=torch.int64)
    return {"feat_row": t(feat_row), "is_person": t(is_person), "frame": t(frame), "obj_pair": t(obj_pair), "N": N}


def edge_threshold(video_size):
    return float(np.round(np.sqrt(video_size[0] ** 2 + video_size[1] ** 2) * SPATIAL_THR, 4))


def build_clip_graph(frames, centers, tokens, edge_thr):
    """One clip: frames [n] (absolute frame id per node, sorted), centers [n,2] fp32, tokens [n,D].
    Returns edge_index [2,E] int64 (clip-local node ids), edge_data [E] int32 (0 spatial, 1 temporal),
    per-frame spatial (u, v) lists (frame-local ids) — all in the reference's order."""
    thr = torch.tensor(edge_thr, dtype=torch.float32)
    org_u, org_v, feat = [], [], []
    spatial_uv = []
    prev_idx = None
    f0, f1 = int(frames.min()), int(frames.max()) + 1
    for f in range(f0, f1):
        idx = (frames == f).nonzero().flatten()
        n = idx.numel()
        su, sv = [], []
        if n > 0:
            base = int(idx[0])
            c = centers[idx]
            iu, iv = torch.triu_indices(n, n, offset=1)
            if iu.numel() > 0:
                dist = torch.sqrt((c[iu, 0] - c[iv, 0]) ** 2 + (c[iu, 1] - c[iv, 1]) ** 2)
                keep = dist <= thr
                for a, b in zip(iu[keep].tolist(), iv[keep].tolist()):
                    org_u += [base + a, base + b]; org_v += [base + b, base + a]; feat += [0, 0]
                    su += [a, b]; sv += [b, a]
            if prev_idx is not None and prev_idx.numel() > 0:
                tp, tc = tokens[prev_idx], tokens[idx]
                cos = (tp @ tc.t()) / (tp.norm(dim=1)[:, None] * tc.norm(dim=1)[None, :])
                pu, cv = (cos >= SIM_THR).nonzero(as_tuple=True)
                for a, b in zip(pu.tolist(), cv.tolist()):
                    ga, gb = int(prev_idx[a]), int(idx[b])
                    org_u += [ga, gb]; org_v += [gb, ga]; feat += [1, 1]
        prev_idx = idx
        spatial_uv.append((su, sv))
    assert len(org_u) > 0, "edge-less clip: the reference's fallback reads stale loop variables (SURVEY A.3 #6)"
    edge_index = torch.tensor([org_u, org_v], dtype=torch.int64)
    return edge_index, torch.tensor(feat, dtype=torch.int32), spatial_uv


def laplacian_eigvec(n, edge_index):
    """lib/teatgt.py:243-254: dense A with multiplicity, in-degree clipped at 1, fp64 eigh -> fp32."""
    A = np.zeros((n, n), dtype=np.float64)
    np.add.at(A, (edge_index[1].numpy(), edge_index[0].numpy()), 1.0)
    deg = torch.bincount(edge_index[1], minlength=n)
    Nm = np.diag(deg.clip(1) ** -0.5)              # float32 values, like the reference
    L = np.eye(n) - Nm @ A @ Nm
    val, vec = np.linalg.eigh(L)
    return torch.tensor(val).type(torch.float32), torch.tensor(vec).type(torch.float32), A


# --------------------------------------------------------------------------------------------
# TokenGT (parameter names follow tools/TokenGT/tokengt)
# --------------------------------------------------------------------------------------------
class _Tokenizer(nn.Module):
    def __init__(self, num_atoms, hidden, lap_k, type_id=True):
        super().__init__()
        self.atom_encoder = nn.Linear(num_atoms, hidden)
        self.temp_encoder = nn.Embedding(100, hidden, padding_idx=0)
        self.edge_encoder = nn.Embedding(5, hidden, padding_idx=0)
        self.graph_token = nn.Embedding(1, hidden)
        self.null_token = nn.Embedding(1, hidden)
        self.lap_encoder = nn.Linear(2 * lap_k, hidden, bias=False)
        self.order_encoder = nn.Embedding(3, hidden)
        self.lap_k = lap_k


class _MHA(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.k_proj = nn.Linear(dim, dim)
        self.v_proj = nn.Linear(dim, dim)
        self.q_proj = nn.Linear(dim, dim)
        self.out_proj = nn.Linear(dim, dim)
        self.heads = heads


class _FFN(nn.Module):
    def __init__(self, dim, ffn):
        super().__init__()
        self.fc1 = nn.Linear(dim, ffn)
        self.fc2 = nn.Linear(ffn, dim)


class _Layer(nn.Module):
    def __init__(self, dim, ffn, heads):
        super().__init__()
        self.self_attn = _MHA(dim, heads)
        self.self_attn_layer_norm = nn.LayerNorm(dim)
        self.feedforward = _FFN(dim, ffn)
        self.final_layer_norm = nn.LayerNorm(dim)


class _GraphEncoder(nn.Module):
    def __init__(self, args):
        super().__init__()
        d = args.encoder_embed_dim
        self.graph_feature = _Tokenizer(args.num_atoms, d, args.lap_node_id_k)
        self.final_layer_norm = nn.LayerNorm(d)        # created by the reference, never applied
        self.layers = nn.ModuleList([_Layer(d, args.encoder_ffn_embed_dim, args.encoder_attention_heads)
                                     for _ in range(args.encoder_layers)])


class TokenGTEncoderOracle(nn.Module):
    def __init__(self, args):
        super().__init__()
        d = args.encoder_embed_dim
        self.args = args
        self.graph_encoder = _GraphEncoder(args)
        self.masked_lm_pooler = nn.Linear(d, d)        # unused by the reference forward
        self.lm_head_transform_weight = nn.Linear(d, d)
        self.layer_norm = nn.LayerNorm(d)
        self.lm_output_learned_bias = nn.Parameter(torch.zeros(args.num_output))
        self.embed_out = nn.Linear(d, args.num_output, bias=False)
        self.p = 0.1

    def tokens(self, node_data, frame_rel, edge_index, edge_data, eigvec, eig_keep=None):
        """tokenizer.py:217-295 for one graph -> [2+n+E, d]."""
        tk = self.graph_encoder.graph_feature
        n, E = node_data.shape[0], edge_index.shape[1]
        node_feat = tk.atom_encoder(node_data) + tk.temp_encoder(frame_rel)
        edge_feat = tk.edge_encoder(edge_data.long())
        u = torch.cat([torch.arange(n), edge_index[0]])
        v = torch.cat([torch.arange(n), edge_index[1]])
        k = tk.lap_k
        ev = F.pad(eigvec, (0, k - eigvec.shape[1])) if k > eigvec.shape[1] else eigvec[:, :k]
        if eig_keep is not None:                       # Dropout2d(0.2) on single elements, train only
            ev = ev * eig_keep
        lap = tk.lap_encoder(torch.cat([ev[u], ev[v]], 1))
        typ = tk.order_encoder((u == v).long())
        feat = torch.cat([node_feat, edge_feat], 0) + lap + typ
        return torch.cat([tk.graph_token.weight, tk.null_token.weight, feat], 0)

    def encode(self, x):
        """12 x pre-LN layer on one sequence [T, d] (no padding: B == 1 in the reference)."""
        training = self.training
        for layer in self.graph_encoder.layers:
            a = layer.self_attn
            h = layer.self_attn_layer_norm(x)
            T, D = h.shape
            H = a.heads
            hd = D // H
            q = a.q_proj(h) * (hd ** -0.5)
            k_, v_ = a.k_proj(h), a.v_proj(h)
            q, k_, v_ = (t.view(T, H, hd).transpose(0, 1) for t in (q, k_, v_))
            w = torch.softmax((q @ k_.transpose(1, 2)).float(), -1)
            w = F.dropout(w, self.p, training)
            o = (w @ v_).transpose(0, 1).reshape(T, D)
            x = x + F.dropout(a.out_proj(o), self.p, training)
            h = layer.final_layer_norm(x)
            f = layer.feedforward
            h = F.dropout(F.gelu(f.fc1(h)), self.p, training)
            x = x + F.dropout(f.fc2(h), self.p, training)
        return x

    def head(self, x_nodes):
        h = self.layer_norm(F.gelu(self.lm_head_transform_weight(x_nodes)))
        return self.embed_out(h) + self.lm_output_learned_bias, h


class TeatgtOracle(nn.Module):
    def __init__(self, mode="predcls", attention_class_num=3, spatial_class_num=6, contact_class_num=17,
                 obj_classes=None, tracking=None, args=None, with_regulariser=True):
        super().__init__()
        assert mode in ("predcls", "sgcls")
        from oracle import ref_shims
        from oracle.tempura_oracle import ObjectClassifierOracle
        self.mode, self.obj_classes, self.args = mode, obj_classes, args
        self.attention_class_num, self.spatial_class_num, self.contact_class_num = (
            attention_class_num, spatial_class_num, contact_class_num)
        # lib/teatgt.py:44-46: obj_head='linear', K=4, no object memory (restated in oracle/tempura_oracle.py; the
        # SGCls-train branch is the same arithmetic as TEMPURA's, tools/utils/object_classifier.py:177-233)
        self.object_classifier = ObjectClassifierOracle(len(obj_classes), mode, "linear", 4, None, None, None, tracking,
                                                        dropout=0.0)
        self.subj_fc = nn.Linear(2048, 968)
        self.obj_fc = nn.Linear(2048, 968)
        self.node_label_tokenizer = nn.Embedding(len(obj_classes), 200)
        self.TokenGT_encoder = TokenGTEncoderOracle(args)
        self.TokenGT_model = nn.Module()
        self.TokenGT_model.encoder = self.TokenGT_encoder          # alias, lib/teatgt.py:61-62
        if with_regulariser:
            self.gat = ref_shims.GraphTransformer(dim=10, depth=4, edge_dim=1, with_feedforwards=True,
                                                  gated_residual=True, rel_pos_emb=True)
            self.gat_semantic = ref_shims.GraphTransformer(dim=768, depth=4, edge_dim=1, with_feedforwards=True,
                                                           gated_residual=True, rel_pos_emb=True)
        self.gate_nn = nn.Linear(10, 1)
        self.gate_sem_nn = nn.Linear(768, 1)
        self.gate_gru_nn = nn.Linear(768, 1)
        for alias, lin in (("gap", self.gate_nn), ("gap_sem", self.gate_sem_nn), ("gap_gru", self.gate_gru_nn)):
            holder = nn.Module()                      # dgl GlobalAttentionPooling(gate_nn) re-registers the Linear
            holder.gate_nn = lin
            setattr(self, alias, holder)

    # -------------------------------------------------------------------------------------
    def node_tokens(self, entry, lay):
        feats = entry["features"][lay["feat_row"]]
        lab = entry["pred_labels"][lay["feat_row"]]
        person = lay["is_person"].bool()
        rep = torch.where(person[:, None], self.subj_fc(feats), self.obj_fc(feats))
        return torch.cat([rep, self.node_label_tokenizer(lab)], 1)                     # [O', 1168]

    def forward(self, entry, phase="train", eig_keep=None, return_artifacts=False):
        assert self.mode == "predcls" or phase == "train", "SGCls test-time relabel/NMS tail is not restated"
        entry = self.object_classifier(entry, phase)
        lay = node_layout(entry)
        tok = self.node_tokens(entry, lay)
        box = entry["boxes"][lay["feat_row"]][:, 1:]
        centers = torch.stack([(box[:, 0] + box[:, 2]) / 2, (box[:, 1] + box[:, 3]) / 2], 1)
        thr = edge_threshold(entry["video_size"])
        frames = lay["frame"]
        n_clips = math.ceil((int(frames.max()) + 1) / CLIP_SIZE)
        enc = self.TokenGT_encoder
        outs, arts = [], []
        str_kl, sem_kl = [], []
        for c in range(n_clips):
            sel = ((frames >= c * CLIP_SIZE) & (frames < (c + 1) * CLIP_SIZE)).nonzero().flatten()
            if sel.numel() == 0:
                continue
            cf, ct, cc = frames[sel], tok[sel], centers[sel]
            edge_index, edge_data, spatial_uv = build_clip_graph(cf, cc, ct, thr)
            n = sel.numel()
            _, eigvec, _ = laplacian_eigvec(n, edge_index)
            keep = eig_keep[c] if eig_keep is not None else None
            x = enc.tokens(ct, (cf - cf.min()), edge_index, edge_data, eigvec, keep)
            x = enc.encode(x)
            logits, hidden = enc.head(x[2:2 + n])
            is_obj = ~lay["is_person"][sel].bool()
            outs.append(logits[is_obj])
            arts.append({"edge_index": edge_index, "edge_data": edge_data, "eigvec": eigvec, "n": n})
            if phase == "train" and hasattr(self, "gat"):
                s, m = self.regulariser(cf, spatial_uv, hidden)
                str_kl += s
                sem_kl += m
        g = torch.cat(outs, 0)
        entry["attention_distribution"] = torch.softmax(g[:, :3], -1)
        entry["spatial_distribution"] = torch.sigmoid(g[:, 3:9])
        entry["contacting_distribution"] = torch.sigmoid(g[:, 9:])
        entry["logits"] = g                                          # oracle-only key
        # detached, like torch.tensor(list_of_scalars) at lib/teatgt.py:350-351
        entry["structure_temp_loss"] = torch.tensor([float(v) for v in str_kl], dtype=torch.float32)
        entry["semantic_temp_loss"] = torch.tensor([float(v) for v in sem_kl], dtype=torch.float32)
        if return_artifacts:
            entry["clip_artifacts"] = arts
        return entry

    # ------------------------------------------------------------------------------------- R1-R3
    def regulariser(self, clip_frames, spatial_uv, hidden):
        """lib/teatgt.py:285-334 (UNPINNED third-party arithmetic, see header)."""
        return regulariser_clip(self.gat, self.gat_semantic, self.gate_nn, self.gate_sem_nn, clip_frames, spatial_uv,
                                hidden)


def regulariser_clip(gat, gat_semantic, gate_nn, gate_sem_nn, clip_frames, spatial_uv, hidden):
    """lib/teatgt.py:285-334 for one clip: per-frame structure / semantic graph embeddings, then KL / frame distance
    over all frame pairs u < v, kept where >= 0.  `hidden` = the clip's feature rows; the semantic graph of every
    frame reads hidden[0:n_f] (`savor` never advances, :312-314).  If the clip owns fewer than n_f rows (possible
    only in the TEMPURA extension, where a clip's rows are its PAIRS while a frame has pairs + 1 nodes) the missing
    rows are zeros."""
    f0 = int(clip_frames.min())
    sym, sem = [], []
    for i, (su, sv) in enumerate(spatial_uv):
        nf = int((clip_frames == f0 + i).sum())
        A = np.zeros((nf, nf))
        if su:
            np.add.at(A, (np.asarray(sv), np.asarray(su)), 1.0)
        deg = torch.bincount(torch.tensor(sv, dtype=torch.int64), minlength=nf) if su else torch.zeros(nf, dtype=torch.int64)
        Nm = np.diag(deg.clip(1) ** -0.5)
        L = np.eye(nf) - Nm @ A @ Nm
        _, vec = np.linalg.eigh(L)
        vec = torch.tensor(vec).type(torch.float32)
        k = 10
        ev = vec.repeat(1, int(k / 2))[:, :k] if k > nf else vec[:, :k]
        nodes = ev[None]
        edges = torch.tensor(A.reshape(1, nf, nf, -1)).type(torch.float32)
        node_sem = hidden[0:nf]                                  # `savor` never advances (quirk kept)
        if node_sem.shape[0] < nf:
            node_sem = torch.cat([node_sem, node_sem.new_zeros(nf - node_sem.shape[0], node_sem.shape[1])], 0)
        no, _ = gat(nodes, edges)
        so, _ = gat_semantic(node_sem[None], edges)
        no, so = no.squeeze(0), so.squeeze(0)
        sym.append((torch.softmax(gate_nn(no), 0) * no).sum(0, keepdim=True))
        sem.append((torch.softmax(gate_sem_nn(so), 0) * so).sum(0, keepdim=True))
    s_out, m_out = [], []
    kl = nn.KLDivLoss(reduction="batchmean")
    for u in range(len(sym)):
        for v in range(u + 1, len(sym)):
            sc = kl(F.log_softmax(sym[u], 1), F.softmax(sym[v], 1)) / (v - u)
            ms = kl(F.log_softmax(sem[u], 1), F.softmax(sem[v], 1)) / (v - u)
            if sc >= 0:
                s_out.append(sc)
            if ms >= 0:
                m_out.append(ms)
    return s_out, m_out


def tempura_consistency(entry, rel_feats, gat, gat_semantic, gate_nn, gate_sem_nn):
    """Restatement of the build's TEMPURA EXTENSION (SURVEY.md A.3 #8 — the reference's lib/tempura.py never fills
    the *_temp_loss keys its trainer reads): the TEAT-GT regulariser R1-R3 unchanged on ONE video's graphs —
    5-frame clips, nodes per frame = person + objects in pair order, spatial edges by box-centre distance
    (lib/teatgt.py:199-209), structure branch on Laplacian eigenvectors, semantic branch on the clip's relation
    feature rows `rel_feats[pairs of the clip]` in place of TokenGT's hidden_x.  Returns two lists of scalars."""
    lay = node_layout(entry)
    box = entry["boxes"][lay["feat_row"]][:, 1:]
    centers = torch.stack([(box[:, 0] + box[:, 2]) / 2, (box[:, 1] + box[:, 3]) / 2], 1)
    thr = edge_threshold(entry["video_size"])
    frames = lay["frame"]
    pair_frame = entry["im_idx"].to(torch.int64)
    n_clips = math.ceil((int(frames.max()) + 1) / CLIP_SIZE)
    s_all, m_all = [], []
    for c in range(n_clips):
        sel = ((frames >= c * CLIP_SIZE) & (frames < (c + 1) * CLIP_SIZE)).nonzero().flatten()
        if sel.numel() == 0:
            continue
        cf, cc = frames[sel], centers[sel]
        # temporal edges play no role here: orthogonal dummy tokens keep build_clip_graph from adding any
        _, _, spatial_uv = _spatial_only(cf, cc, thr)
        rows = ((pair_frame >= c * CLIP_SIZE) & (pair_frame < (c + 1) * CLIP_SIZE)).nonzero().flatten()
        s, m = regulariser_clip(gat, gat_semantic, gate_nn, gate_sem_nn, cf, spatial_uv, rel_feats[rows])
        s_all += s
        m_all += m
    return s_all, m_all


def _spatial_only(frames, centers, edge_thr):
    """The spatial half of build_clip_graph (lib/teatgt.py:199-209): per frame, frame-local (u, v) edge lists."""
    thr = torch.tensor(edge_thr, dtype=torch.float32)
    out = []
    for f in range(int(frames.min()), int(frames.max()) + 1):
        idx = (frames == f).nonzero().flatten()
        n = idx.numel()
        su, sv = [], []
        if n > 1:
            c = centers[idx]
            iu, iv = torch.triu_indices(n, n, offset=1)
            dist = torch.sqrt((c[iu, 0] - c[iv, 0]) ** 2 + (c[iu, 1] - c[iv, 1]) ** 2)
            keep = dist <= thr
            for a, b in zip(iu[keep].tolist(), iv[keep].tolist()):
                su += [a, b]
                sv += [b, a]
        out.append((su, sv))
    return None, None, out


def teatgt_losses(pred, attention_label, spatial_label, contact_label):
    """TEATGT_train.py:167-175."""
    return {
        "attention_relation_loss": F.cross_entropy(pred["attention_distribution"], attention_label),
        "spatial_relation_loss": F.binary_cross_entropy(pred["spatial_distribution"], spatial_label),
        "contacting_relation_loss": F.binary_cross_entropy(pred["contacting_distribution"], contact_label),
    }
