"""Kernel-level checks of the differentiable consistency mode's building blocks (csrc/consistency.cu) against plain PyTorch
fp32 autograd of the same op on the same seeded inputs: attention pooling, pairwise KL, gated residual, the SIMT linears /
LayerNorm of the 10-wide structure branch and the row-weighted column sum.  (The GraphTransformer attention-core backward is
checked end to end against the oracle's autograd in tests/test_tempura_gpu.py.)  Tolerance 2e-5 relative: all fp32."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _close(a, b, tol=2e-5):
    a, b = a.float().cpu(), b.float().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    assert (a - b).abs().max().item() <= tol * max(1.0, b.abs().max().item()), (a - b).abs().max().item()


def _frames(g, counts, d):
    off = torch.tensor([0] + list(torch.tensor(counts).cumsum(0)), dtype=torch.int32, device=DEV)
    x = torch.randn(int(off[-1]), d, generator=g, device=DEV)
    return off, x


@pytest.mark.parametrize("d", [10, 768])
def test_attn_pool_fwd_bwd(cuda_lib, d):
    from b200vsgg.regulariser import _AttnPoolFn
    g = torch.Generator(device=DEV).manual_seed(1)
    counts = [3, 7, 1, 11, 5]
    off, x = _frames(g, counts, d)
    w = torch.randn(1, d, generator=g, device=DEV).requires_grad_(True)
    b = torch.randn(1, generator=g, device=DEV).requires_grad_(True)
    x.requires_grad_(True)
    out = _AttnPoolFn.apply(x, off, len(counts), max(counts), w, b)
    go = torch.randn(len(counts), d, generator=g, device=DEV)
    out.backward(go)
    xr, wr, br = (t.detach().clone().requires_grad_(True) for t in (x, w, b))
    refs = []
    for f in range(len(counts)):
        rows = xr[int(off[f]):int(off[f + 1])]
        a = torch.softmax(rows @ wr[0] + br, 0)
        refs.append((a[:, None] * rows).sum(0))
    ref = torch.stack(refs)
    ref.backward(go)
    _close(out, ref)
    _close(x.grad, xr.grad)
    _close(w.grad, wr.grad)
    assert b.grad.abs().max().item() <= 1e-4 * max(1.0, w.grad.abs().max().item())       # structurally zero


@pytest.mark.parametrize("d", [10, 1936])
def test_consistency_kl_fwd_bwd(cuda_lib, d):
    from b200vsgg.regulariser import _ConsistencyKLFn
    g = torch.Generator(device=DEV).manual_seed(2)
    Fr = 9
    emb = torch.randn(Fr, d, generator=g, device=DEV).requires_grad_(True)
    pairs = [(u, v) for u in range(Fr) for v in range(u + 1, min(Fr, u + 5))]
    pu = torch.tensor([p[0] for p in pairs], dtype=torch.int32, device=DEV)
    pv = torch.tensor([p[1] for p in pairs], dtype=torch.int32, device=DEV)
    out = _ConsistencyKLFn.apply(emb, pu, pv)
    go = torch.rand(len(pairs), generator=g, device=DEV)
    go[3] = 0.0                                              # a pair the `>= 0` filter dropped
    out.backward(go)
    er = emb.detach().clone().requires_grad_(True)
    ref = torch.stack([F.kl_div(F.log_softmax(er[u][None], 1), F.softmax(er[v][None], 1), reduction="batchmean") / (v - u)
                       for u, v in pairs])
    ref.backward(go)
    _close(out, ref)
    _close(emb.grad, er.grad)


@pytest.mark.parametrize("dim", [10, 768])
def test_gated_residual_bwd(cuda_lib, dim):
    from b200vsgg import ops
    g = torch.Generator(device=DEV).manual_seed(3)
    R = 37
    o, res, dx = (torch.randn(R, dim, generator=g, device=DEV) for _ in range(3))
    w = 0.3 * torch.randn(3 * dim, generator=g, device=DEV)
    d_o, d_res, da = torch.empty_like(o), torch.empty_like(o), torch.empty(R, device=DEV)
    ops.gated_residual_bwd(o, res, w, dx, d_o, d_res, da)
    dw1, dw2 = torch.zeros(dim, device=DEV), torch.zeros(dim, device=DEV)
    ops.weighted_colsum(o, da, dw1)
    ops.weighted_colsum(res, da, dw2)
    orr, rr, wr = (t.clone().requires_grad_(True) for t in (o, res, w))
    gate = torch.sigmoid(torch.cat([orr, rr, orr - rr], 1) @ wr)[:, None]
    x = orr * gate + rr * (1 - gate)
    # forward kernel agrees with this formula (in place on `res`)
    xk = res.clone()
    ops.gated_residual(o, xk, w)
    _close(xk, x)
    x.backward(dx)
    _close(d_o, orr.grad)
    _close(d_res, rr.grad)
    _close(torch.cat([dw1, dw2, dw1 - dw2]), wr.grad, 1e-4)


def test_simt_linear_wgrad_and_small_layernorm(cuda_lib):
    from b200vsgg import ops
    g = torch.Generator(device=DEV).manual_seed(4)
    R = 301
    for n_in, n_out in ((10, 1536), (512, 10), (10, 40), (40, 10)):
        x = torch.randn(R, n_in, generator=g, device=DEV)
        w = torch.randn(n_out, n_in, generator=g, device=DEV) / n_in ** 0.5
        b = torch.randn(n_out, generator=g, device=DEV)
        y, z = ops.simt_linear(x, w, b, act=ops.ACT_GELU, want_z=True)
        _close(z, x @ w.t() + b)
        _close(y, F.gelu(x @ w.t() + b))
        dy = torch.randn(R, n_out, generator=g, device=DEV)
        _close(ops.simt_linear(dy, w, transposed=True), dy @ w, 5e-5)
        _close(ops.simt_wgrad(dy, x), dy.t() @ x, 1e-4)
        zr = z.clone().requires_grad_(True)
        F.gelu(zr).backward(dy)
        _close(ops.gelu_bwd(dy, z), zr.grad)
    x = torch.randn(R, 10, generator=g, device=DEV)
    gam, bet = torch.rand(10, generator=g, device=DEV) + 0.5, torch.randn(10, generator=g, device=DEV)
    y, mean, rstd = ops.ln_small_fwd(x, gam, bet)
    xr, gr, br = (t.clone().requires_grad_(True) for t in (x, gam, bet))
    ref = F.layer_norm(xr, (10,), gr, br, 1e-5)
    _close(y, ref)
    dy, base = torch.randn(R, 10, generator=g, device=DEV), torch.randn(R, 10, generator=g, device=DEV)
    ref.backward(dy)
    dx, dg, db = ops.ln_small_bwd(dy, x, gam, mean, rstd, base=base)
    _close(dx, xr.grad + base)
    _close(dg, gr.grad, 1e-4)
    _close(db, br.grad, 1e-4)


def test_graph_attn_core_fwd_bwd(cuda_lib):
    """Attention core of graph_transformer_pytorch.Attention (8 heads x 64, rotary by node index, per-edge key / value
    offsets e_ij = A_ij we + be) per frame: forward (fp32 output variant) and the hand-written backward against autograd
    through a direct torch restatement."""
    from b200vsgg import ops
    g = torch.Generator(device=DEV).manual_seed(5)
    counts, nmax, H, DH = [4, 9, 1, 11, 6], 12, 8, 64
    inner = H * DH
    off = torch.tensor([0] + list(torch.tensor(counts).cumsum(0)), dtype=torch.int32, device=DEV)
    R = int(off[-1])
    qkv = torch.randn(R, 3 * inner, generator=g, device=DEV)
    upper = torch.zeros(len(counts), nmax, nmax, dtype=torch.uint8, device=DEV)
    for f, n in enumerate(counts):
        u = (torch.rand(n, n, generator=g, device=DEV) < 0.5).triu(1)
        upper[f, :n, :n] = u.to(torch.uint8)
    we, be = 0.3 * torch.randn(inner, generator=g, device=DEV), 0.3 * torch.randn(inner, generator=g, device=DEV)
    out = torch.empty(R, inner, device=DEV)
    ops.graph_attn_core(qkv, off, upper, nmax, we, be, out)
    dout = torch.randn(R, inner, generator=g, device=DEV)
    dqkv, dwe, dbe = torch.empty(R, 3 * inner, device=DEV), torch.zeros(inner, device=DEV), torch.zeros(inner, device=DEV)
    ops.graph_attn_core_bwd(qkv, off, upper, nmax, we, be, dout, dqkv, dwe, dbe)

    qr, wr, br = (t.clone().requires_grad_(True) for t in (qkv, we, be))
    inv_freq = 10000.0 ** (-torch.arange(0, DH, 2, device=DEV).float() / DH)

    def rot(x, n):                                     # x [n, H, DH]: rotate the pairs (2l, 2l+1) by position * inv_freq[l]
        ang = torch.arange(n, device=DEV).float()[:, None] * inv_freq[None]
        cs, sn = ang.cos()[:, None, :], ang.sin()[:, None, :]
        x0, x1 = x[..., 0::2], x[..., 1::2]
        return torch.stack([x0 * cs - x1 * sn, x1 * cs + x0 * sn], -1).flatten(-2)

    refs = []
    for f, n in enumerate(counts):
        a, b = int(off[f]), int(off[f + 1])
        q, k, v = (qr[a:b, i * inner:(i + 1) * inner].view(n, H, DH) for i in range(3))
        q, k = rot(q, n), rot(k, n)
        A = (upper[f, :n, :n] + upper[f, :n, :n].t()).float()
        e = A[:, :, None, None] * wr.view(H, DH) + br.view(H, DH)                    # [i, j, H, DH]
        sim = torch.einsum("ihd,ijhd->hij", q, k[None] + e) / 8.0
        p = sim.softmax(-1)
        refs.append(torch.einsum("hij,ijhd->ihd", p, v[None] + e).reshape(n, inner))
    ref = torch.cat(refs)
    ref.backward(dout)
    _close(out, ref, 2e-4)
    _close(dqkv, qr.grad, 2e-4)
    _close(dwe, wr.grad, 2e-4)
    _close(dbe, br.grad, 2e-4)
