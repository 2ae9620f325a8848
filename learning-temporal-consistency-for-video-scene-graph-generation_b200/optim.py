"""Fused AdamW + global-norm gradient clipping behind the reference optimiser's interface.

Drop-in for `tools/utils/AdamW.py::AdamW` followed by `torch.nn.utils.clip_grad_norm_(params, max_norm)`
(TEMPURA_train.py:111, 224-225): weight decay multiplies the weights before the moment update, parameters
whose gradient is None are skipped (and keep their own step counter), bias corrections follow each
tensor's step count.  Two kernel launches per step for the whole model (b200vsgg_grad_sqnorm,
b200vsgg_adamw_clip_step); the clip coefficient never visits the host.
"""
import numpy as np
import torch

from . import ops


class FusedAdamW:
    CHUNK = 1 << 16

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=None):
        self.params = [p for p in params]
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.max_grad_norm = max_grad_norm
        self.state = {}
        self._layout_key, self._chunks = None, None
        self.last_sq_norm = None

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    @torch.no_grad()
    def step(self):
        live = [p for p in self.params if p.grad is not None]
        if not live:
            return
        dev = live[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdamW runs on CUDA parameters only")
        b1, b2 = self.betas
        table = np.zeros(len(live), dtype=np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("n", "<i8"),
                                                    ("bc1", "<f4"), ("bc2", "<f4")]))
        for i, p in enumerate(live):
            st = self.state.get(p)
            if st is None:
                st = self.state[p] = {"step": 0, "exp_avg": torch.zeros_like(p), "exp_avg_sq": torch.zeros_like(p)}
            st["step"] += 1
            g = p.grad
            if not g.is_contiguous() or g.dtype != torch.float32:
                g = p.grad = g.contiguous().float()
            assert p.is_contiguous() and p.dtype == torch.float32
            table[i] = (p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel(),
                        1 - b1 ** st["step"], 1 - b2 ** st["step"])
        key = tuple(p.numel() for p in live)
        if key != self._layout_key:
            ct, co = [], []
            for i, n in enumerate(key):
                offs = np.arange(0, n, self.CHUNK, dtype=np.int64)
                ct.append(np.full(offs.shape, i, dtype=np.int32))
                co.append(offs)
            self._chunks = (ops.upload(np.concatenate(ct), dev), ops.upload(np.concatenate(co), dev))
            self._layout_key = key
        tens = ops.upload(table.view(np.uint8), dev)
        chunk_tensor, chunk_off = self._chunks
        sq = None
        if self.max_grad_norm is not None:
            sq = torch.zeros(1, device=dev)
            ops.grad_sqnorm(tens, chunk_tensor, chunk_off, self.CHUNK, sq)
            self.last_sq_norm = sq
        ops.adamw_clip_step(tens, chunk_tensor, chunk_off, self.CHUNK, sq, self.max_grad_norm or 0.0, self.lr, b1, b2,
                            self.eps, self.weight_decay)
