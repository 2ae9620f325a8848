"""Fused AdamW + global-norm gradient clipping behind the reference optimiser's interface.

Drop-in for `tools/utils/AdamW.py::AdamW` followed by `torch.nn.utils.clip_grad_norm_(params, max_norm)`
(TEMPURA_train.py:111, 224-225): weight decay multiplies the weights before the moment update, parameters
whose gradient is None are skipped (and keep their own step counter), bias corrections follow each
tensor's step count.  Two kernel launches per parameter group for the whole model (b200vsgg_grad_sqnorm,
b200vsgg_adamw_clip_step); the clip coefficient never visits the host.

It IS a `torch.optim.Optimizer`: `param_groups` (lr / betas / eps / weight_decay are read from the group at every
step, so `ExponentialLR(optimizer, gamma=0.8)` and `pytorch_warmup.ExponentialWarmup` of the reference recipe,
TEMPURA_train.py:113-114, drive it unchanged), `state_dict()` / `load_state_dict()` with the reference's state keys
(`step`, `exp_avg`, `exp_avg_sq`), `add_param_group`, `zero_grad`.
Differences from calling clip_grad_norm_ yourself: the clip coefficient is applied inside the update kernel, so `p.grad`
is left UNSCALED; the total norm is available afterwards as `optimizer.total_norm()` (a device tensor, no sync).
"""
import numpy as np
import torch

from . import ops


class FusedAdamW(torch.optim.Optimizer):
    CHUNK = 1 << 16

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=None):
        if lr < 0.0 or eps < 0.0 or not (0.0 <= betas[0] < 1.0) or not (0.0 <= betas[1] < 1.0):
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.max_grad_norm = max_grad_norm
        self._layouts = {}            # group index -> (numel key, chunk tables)
        self.last_sq_norm = None

    def total_norm(self):
        """L2 norm of all gradients of the last step() (device tensor; what clip_grad_norm_ would have returned)."""
        return None if self.last_sq_norm is None else self.last_sq_norm.sqrt()[0]

    _ROW = np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("n", "<i8"), ("bc1", "<f4"), ("bc2", "<f4")])

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._layouts = {}                       # the moment tensors were replaced: rebuild the pointer tables

    def _tables(self, gi, live, group, dev):
        """Per-tensor pointer / bias-correction table of one group.  The static columns (parameter and moment pointers,
        sizes, chunk layout) are cached and re-validated with two cheap passes; only the gradient pointers and the bias
        corrections are refreshed per step (this host loop used to leave the device idle for ~0.75 ms per step)."""
        b1, b2 = group["betas"]
        state = self.state
        lay = self._layouts.get(gi)
        pp = np.fromiter((p.data_ptr() for p in live), dtype=np.uint64, count=len(live))
        ok = lay is not None and lay["n"] == len(live) and np.array_equal(lay["table"]["p"], pp)
        if ok:
            for p, st in zip(live, lay["states"]):
                if state.get(p) is not st:
                    ok = False
                    break
        if not ok:
            states = []
            for p in live:
                st = state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                assert p.is_contiguous() and p.dtype == torch.float32
                states.append(st)
            table = np.zeros(len(live), dtype=self._ROW)
            table["p"] = pp
            table["m"] = [st["exp_avg"].data_ptr() for st in states]
            table["v"] = [st["exp_avg_sq"].data_ptr() for st in states]
            table["n"] = [p.numel() for p in live]
            ct, co = [], []
            for i, n in enumerate(table["n"].tolist()):
                offs = np.arange(0, n, self.CHUNK, dtype=np.int64)
                ct.append(np.full(offs.shape, i, dtype=np.int32))
                co.append(offs)
            lay = dict(n=len(live), table=table, states=states, ct=ops.upload(np.concatenate(ct), dev),
                       co=ops.upload(np.concatenate(co), dev))
            self._layouts[gi] = lay
        table, states = lay["table"], lay["states"]
        # (a loaded state_dict may hold a tensor / float step)
        steps = np.fromiter((int(st["step"]) + 1 for st in states), dtype=np.int64, count=len(states))
        for st, k in zip(states, steps.tolist()):
            st["step"] = k
        gp = []
        for p in live:
            g = p.grad
            if not g.is_contiguous() or g.dtype != torch.float32:
                g = p.grad = g.contiguous().float()
            gp.append(g.data_ptr())
        table["g"] = gp
        if steps.min() == steps.max():           # the usual case: one pow per step instead of one per tensor
            k = int(steps[0])
            table["bc1"], table["bc2"] = 1 - b1 ** k, 1 - b2 ** k
        else:
            table["bc1"] = [1 - b1 ** k for k in steps.tolist()]
            table["bc2"] = [1 - b2 ** k for k in steps.tolist()]
        return ops.upload(table.view(np.uint8), dev), lay["ct"], lay["co"]

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        work = []
        for gi, group in enumerate(self.param_groups):
            live = [p for p in group["params"] if p.grad is not None]     # tools/utils/AdamW.py:66-67
            if not live:
                continue
            dev = live[0].device
            if dev.type != "cuda":
                raise RuntimeError("FusedAdamW runs on CUDA parameters only")
            work.append((group, live, dev) + self._tables(gi, live, group, dev))
        if not work:
            return loss
        sq = None
        if self.max_grad_norm is not None:                                # ONE global norm over all groups
            sq = torch.zeros(1, device=work[0][2])
            for _, _, _, tens, ct, co in work:
                ops.grad_sqnorm(tens, ct, co, self.CHUNK, sq)
            self.last_sq_norm = sq
        for group, live, _, tens, ct, co in work:
            b1, b2 = group["betas"]
            ops.adamw_clip_step(tens, ct, co, self.CHUNK, sq, self.max_grad_norm or 0.0, float(group["lr"]), b1, b2,
                                group["eps"], group["weight_decay"])
            # the kernel wrote the parameters behind torch's back: bump their version counters so that anything keyed on
            # `param._version` (the bf16 weight cache of the models, autograd's saved-tensor checks) sees the update
            torch.autograd.graph.increment_version(live)
        return loss
