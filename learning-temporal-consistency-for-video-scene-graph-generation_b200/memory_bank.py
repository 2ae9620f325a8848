"""Uncertainty / class-memory bookkeeping of the TEMPURA trainer on the device (second half).

The reference keeps, for every training video, the per-class uncertainties of its `model(entry, unc=True)` pass in nested
Python dicts (one `.item()` per label, tools/utils/Uncertainty.py:105-178), writes the video's relation features to
`rel_embeddings/<index>.npy`, and at the end of the epoch reloads every file to form
`rel_memory[rel][k] = sum_i w_ik * rel_features_i` with `w_ik = exp(u_ik) / (Z_k + 1e-12)` (tools/utils/Memory.py:5-135,
Uncertainty.py:194-246).  Because the weights factor as `exp(u_ik)` over an epoch-level constant `Z_k`, the sums can be
accumulated as the videos stream by: per step ONE kernel adds `exp(u_ik) * f_i` into a device-resident [classes, 1936]
accumulator per predicate group (b200vsgg_class_memory_accumulate) and a [classes] vector collects `Z_k`; `finalize()`
divides.  No files, no per-label host reads; the result equals the reference's up to fp32 summation order
(tests/test_memory_bank.py: golden values from the UNMODIFIED `memory_computation` / `normalize_batch_uncertainty` /
`uncertainty_values.stats2`).

Reference quirks reproduced: with weight type 'both' the weight exponent is `al + ep` but the class
normaliser of the relation memories is `sum(exp(al_list + al_list))` — a LIST concatenation (Uncertainty.py:64), i.e.
`2 * sum_j exp(al_jk)`; weight types 'al' / 'ep' normalise by `sum_j exp(u_jk)`; `None` / 'simple' use weight 1 and
'simple' alone divides by the label count (Memory.py:119-133).  Entries whose uncertainty is exactly 0.0 are skipped like
the reference's `np.where(batch_unc != 0)`.
Only the relation memories are built here: PredCLS trains with `obj_mem_compute=False`, and the reference's object-memory
branch with uncertainty weights cannot run as released (tools/utils/Memory.py:90-94 multiplies by `obj_features`, which
that branch never loads), so there is no reference behaviour to reproduce for it."""
import numpy as np
import torch

from . import ops

REL_KEYS = ("attention", "spatial", "contacting")


class ClassMemoryBank:
    """Drop-in for the trainer's `unc_vals = uncertainty_values(...)` + `uncertainty_computation(...)` per step +
    `memory_computation(...)` per epoch (TEMPURA_train.py:136-141,168-172,369-379)."""

    def __init__(self, rel_class_num, rel_feature_dim=1936, rel_weight_type="both", device="cuda"):
        if rel_weight_type not in ("both", "al", "ep", "simple", None):
            raise ValueError("rel_weight_type must be 'both', 'al', 'ep', 'simple' or None")
        self.rel_class_num = dict(rel_class_num)
        self.dim, self.weight_type, self.device = rel_feature_dim, rel_weight_type, torch.device(device)
        self.reset()

    def reset(self):
        self.acc = {k: torch.zeros(c, self.dim, device=self.device) for k, c in self.rel_class_num.items()}
        self.norm = {k: torch.zeros(c, device=self.device, dtype=torch.float64) for k, c in self.rel_class_num.items()}
        self.videos = 0

    @staticmethod
    def _entries(labels, device, dedup):
        """Ragged label lists (the dataloader's attention_gt / spatial_gt / contacting_gt) -> (row, class) index vectors.
        dedup=True: a class listed twice for one pair is one entry (the reference ASSIGNS `batch_unc[i, k] = ...`);
        dedup=False: one entry per occurrence (its per-class lists `cls_rel_uc[rel][k][u]` get one append per occurrence)."""
        rows, cls = [], []
        for i, l in enumerate(labels):
            ks = np.asarray(l).reshape(-1).astype(np.int64)
            if dedup:
                ks = np.unique(ks)
            rows.append(np.full(ks.shape, i, dtype=np.int64))
            cls.append(ks)
        rows = np.concatenate(rows) if rows else np.zeros(0, np.int64)
        cls = np.concatenate(cls) if cls else np.zeros(0, np.int64)
        return ops.upload(rows, device), ops.upload(cls, device)

    @torch.no_grad()
    def update(self, pred, gt=None):
        """One training video (or batch): pred = model(entry, unc=True) with `rel_features` and the `<rel>_{al,ep}_uc`
        tensors; gt = {'attention': lists, 'spatial': lists, 'contacting': lists} (default: pred['<rel>_gt'])."""
        feat = pred["rel_features"].detach().float().contiguous()
        if not feat.is_cuda:
            raise RuntimeError("ClassMemoryBank runs on CUDA tensors only (no CPU fallback)")
        for rel in REL_KEYS:
            labels = (gt or {}).get(rel, pred.get(rel + "_gt" if rel != "contacting" else "contacting_gt"))
            rows, cls = self._entries(labels, feat.device, dedup=True)
            if rows.numel() == 0:
                continue
            zrows, zcls = self._entries(labels, feat.device, dedup=False)
            wt = self.weight_type
            al = pred[rel + "_al_uc"].detach().float()
            ep = pred[rel + "_ep_uc"].detach().float()
            if wt in ("both", "al", "ep"):
                u = al + ep if wt == "both" else (al if wt == "al" else ep)
                ue = u[rows, cls]
                w = torch.where(ue != 0, torch.exp(ue), torch.zeros_like(ue))           # np.where(batch_unc != 0)
                # class normaliser: stats2() of the reference (list concatenation for 'both': 2 * sum exp(al))
                z = 2.0 * torch.exp(al[zrows, zcls].double()) if wt == "both" else torch.exp(u[zrows, zcls].double())
                self.norm[rel].index_add_(0, zcls, z)
            else:
                w = (al[rows, cls] != 0).float()                                        # weight 1 where the entry is set
                self.norm[rel].index_add_(0, cls, w.double())
            ops.class_memory_accumulate(feat, rows.int(), cls.int(), w.contiguous(), self.acc[rel])
        self.videos += 1

    @torch.no_grad()
    def finalize(self):
        """rel_memory dict as `memory_computation` returns it (float32 [classes, dim] per predicate group, on the device)."""
        out = {}
        for rel in REL_KEYS:
            a, z = self.acc[rel], self.norm[rel]
            if self.weight_type in ("both", "al", "ep"):
                out[rel] = (a.double() / (z + 1e-12)[:, None]).float()
            elif self.weight_type == "simple":
                out[rel] = torch.where((z != 0)[:, None], a.double() / z.clamp(min=1)[:, None], a.double()).float()
            else:
                out[rel] = a.clone()
        return out
