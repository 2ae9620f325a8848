"""Optimiser step: oracle vs the golden vector written by the UNMODIFIED reference AdamW (CPU, bit-exact),
and the fused CUDA step (b200vsgg.optim.FusedAdamW) vs the oracle (fp32, <= 2e-6 relative)."""
import os

import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "adamw.pt")


def _run(make_opt, device="cpu"):
    from oracle.make_golden_adamw import SHAPES, make_case
    g = torch.Generator().manual_seed(7)
    params = [torch.nn.Parameter(torch.randn(s, generator=g).to(device)) for s in SHAPES]
    step_fn = make_opt(params)
    for step in range(3):
        grads = make_case(step)
        for i, (p, gr) in enumerate(zip(params, grads)):
            p.grad = None if (i == 3 or (i == 1 and step == 1)) else gr.clone().to(device)
        step_fn()
    return [p.detach().cpu() for p in params]


def test_oracle_matches_reference_golden():
    from oracle.adamw_oracle import AdamWOracle
    gold = torch.load(GOLDEN, weights_only=False)

    def mk(params):
        o = AdamWOracle(params, max_grad_norm=5, **gold["kw"])
        return o.step
    for got, ref in zip(_run(mk), gold["params_after"]):
        assert torch.equal(got, ref)


@pytest.mark.gpu
def test_fused_adamw_matches_oracle(cuda_lib):
    from b200vsgg.optim import FusedAdamW
    gold = torch.load(GOLDEN, weights_only=False)

    def mk(params):
        o = FusedAdamW(params, max_grad_norm=5, **gold["kw"])
        return o.step
    for got, ref in zip(_run(mk, "cuda"), gold["params_after"]):
        assert (got - ref).abs().max().item() <= 2e-6 * ref.abs().max().item() + 1e-7
