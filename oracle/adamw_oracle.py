"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

CPU restatement of the reference optimiser step: tools/utils/AdamW.py:53-113 preceded by
torch.nn.utils.clip_grad_norm_(params, max_norm) (TEMPURA_train.py:224-225).
Parity status: PINNED — oracle/make_golden_adamw.py imports the unmodified reference AdamW (pure torch,
importable as is) and checks this restatement bit-exactly; tests/golden/adamw.pt holds its outputs.
"""
import math

import torch


class AdamWOracle:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=None):
        self.params = list(params)
        self.lr, self.betas, self.eps, self.wd, self.max_grad_norm = lr, betas, eps, weight_decay, max_grad_norm
        self.state = {}

    @torch.no_grad()
    def step(self):
        if self.max_grad_norm is not None:
            torch.nn.utils.clip_grad_norm_(self.params, max_norm=self.max_grad_norm, norm_type=2)
        b1, b2 = self.betas
        for p in self.params:
            if p.grad is None:                                   # AdamW.py:66-67
                continue
            p.mul_(1 - self.lr * self.wd)                        # decay first, AdamW.py:69
            st = self.state.setdefault(p, {"step": 0, "m": torch.zeros_like(p), "v": torch.zeros_like(p)})
            st["step"] += 1
            st["m"].mul_(b1).add_(p.grad, alpha=1 - b1)
            st["v"].mul_(b2).addcmul_(p.grad, p.grad, value=1 - b2)
            denom = st["v"].sqrt().add_(self.eps)
            step_size = self.lr * math.sqrt(1 - b2 ** st["step"]) / (1 - b1 ** st["step"])
            p.addcdiv_(st["m"], denom, value=-step_size)
