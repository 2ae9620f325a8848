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


def test_fused_adamw_is_a_torch_optimizer():
    """the reference recipe wraps its optimiser in ExponentialLR(gamma=0.8) and a warm-up scheduler
    (TEMPURA_train.py:113-114), both of which need `param_groups`; checkpoints need state_dict / load_state_dict."""
    from b200vsgg.optim import FusedAdamW
    params = [torch.nn.Parameter(torch.zeros(4, 3)), torch.nn.Parameter(torch.zeros(5))]
    opt = FusedAdamW(params, lr=1e-5, weight_decay=0.1, max_grad_norm=5.0)
    assert isinstance(opt, torch.optim.Optimizer)
    assert opt.param_groups[0]["lr"] == 1e-5 and opt.param_groups[0]["weight_decay"] == 0.1
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.8)
    opt.param_groups[0]["lr"] *= 0.5                          # what a warm-up scheduler does
    sd = opt.state_dict()
    assert set(sd) == {"state", "param_groups"} and sd["param_groups"][0]["lr"] == 0.5e-5
    opt2 = FusedAdamW(params, lr=1.0)
    opt2.load_state_dict(sd)
    assert opt2.param_groups[0]["lr"] == 0.5e-5 and opt2.param_groups[0]["betas"] == (0.9, 0.999)
    opt.add_param_group({"params": [torch.nn.Parameter(torch.zeros(2))], "lr": 3e-4})
    assert len(opt.param_groups) == 2 and opt.param_groups[1]["weight_decay"] == 0.1
    assert sched.get_last_lr()[0] == 1e-5
    opt.zero_grad()


@pytest.mark.gpu
def test_fused_adamw_with_lr_schedule_and_checkpoint_matches_oracle(cuda_lib):
    """ExponentialLR + a warm-up factor drive the fused optimiser through `param_groups`; after a state_dict round trip
    into a fresh instance, 4 steps equal the oracle (reference AdamW + clip_grad_norm_(5)) run with the same lr values."""
    from b200vsgg.optim import FusedAdamW
    from oracle.adamw_oracle import AdamWOracle
    from oracle.make_golden_adamw import SHAPES, make_case
    gold = torch.load(GOLDEN, weights_only=False)
    g = torch.Generator().manual_seed(11)
    init = [torch.randn(s, generator=g) for s in SHAPES]
    pc = [torch.nn.Parameter(t.clone()) for t in init]
    pg = [torch.nn.Parameter(t.clone().cuda()) for t in init]
    ref = AdamWOracle(pc, max_grad_norm=5, **gold["kw"])
    opt = FusedAdamW(pg, max_grad_norm=5, **gold["kw"])
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.8)
    base_lr = gold["kw"]["lr"]
    for step in range(4):
        if step == 2:                                          # checkpoint / resume in the middle
            sd = opt.state_dict()
            opt = FusedAdamW(pg, max_grad_norm=5, **gold["kw"])
            opt.load_state_dict(sd)
            sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.8, last_epoch=step - 1)
        warm = min(1.0, (step + 1) / 3.0)
        lr_sched = opt.param_groups[0]["lr"]
        opt.param_groups[0]["lr"] = lr_sched * warm
        ref.lr = opt.param_groups[0]["lr"]
        grads = make_case(step)
        for i, gr in enumerate(grads):
            pc[i].grad = None if i == 3 else gr.clone()
            pg[i].grad = None if i == 3 else gr.clone().cuda()
        ref.step()
        opt.step()
        opt.param_groups[0]["lr"] = lr_sched
        sched.step()
        assert abs(opt.param_groups[0]["lr"] - base_lr * 0.8 ** (step + 1)) < 1e-12
    norm = opt.total_norm()
    want = torch.sqrt(sum((gr.float() ** 2).sum() for i, gr in enumerate(make_case(3)) if i != 3))
    assert abs(norm.item() - want.item()) <= 1e-4 * want.item()
    for got, want_p in zip(pg, pc):
        assert (got.detach().cpu() - want_p.detach()).abs().max().item() <= 2e-6 * want_p.abs().max().item() + 1e-7
