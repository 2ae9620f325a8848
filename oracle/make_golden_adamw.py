"""TEST INFRASTRUCTURE.  Runs the UNMODIFIED reference optimiser (tools/utils/AdamW.py) for 3 steps with
clip_grad_norm_(5) on seeded tensors (one parameter never receives a gradient, one skips a step), checks
oracle/adamw_oracle.py bit-exactly and writes tests/golden/adamw.pt."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.environ.get("VSGG_REFERENCE", "/root/reference"))
KW = dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.1)
SHAPES = [(37, 200), (1936,), (300, 1936), (5,), (2048, 3)]


def make_case(step):
    g = torch.Generator().manual_seed(100 + step)
    return [torch.randn(s, generator=g) * (3.0 if step == 1 else 0.05) for s in SHAPES]


def run(make_opt):
    g = torch.Generator().manual_seed(7)
    params = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in SHAPES]
    opt = make_opt(params)
    for step in range(3):
        grads = make_case(step)
        for i, (p, gr) in enumerate(zip(params, grads)):
            p.grad = None if (i == 3 or (i == 1 and step == 1)) else gr.clone()
        opt(params)
    return [p.detach().clone() for p in params]


def main():
    from tools.utils.AdamW import AdamW
    from oracle.adamw_oracle import AdamWOracle

    def ref(params):
        o = AdamW(params, **KW)

        def step(ps):
            torch.nn.utils.clip_grad_norm_(ps, max_norm=5, norm_type=2)
            o.step()
        return step

    def orc(params):
        o = AdamWOracle(params, max_grad_norm=5, **KW)
        return lambda ps: o.step()

    a, b = run(ref), run(orc)
    for x, y in zip(a, b):
        assert torch.equal(x, y), "oracle differs from the reference optimiser"
    torch.save({"kw": KW, "shapes": SHAPES, "params_after": a}, os.path.join(ROOT, "tests", "golden", "adamw.pt"))
    print("oracle == reference AdamW (bit-exact); wrote tests/golden/adamw.pt")


if __name__ == "__main__":
    main()
