"""Variable-length flash attention (csrc/attn_flash.cu) against a plain torch fp32 reference, through the
C-ABI.  Tolerance: bf16 inputs/outputs, fp32 accumulation -> 2^-7 relative to the tensor's max + 1e-3."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _plan(lens):
    from b200vsgg.plan import attention_blocks
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    bs, br = attention_blocks(off)
    t = lambda a: torch.from_numpy(a).to(DEV)
    return off, t(off), t(bs), t(br)


def _ref(q, k, v, off, H, hd, mask=None, p=0.0):
    outs = []
    for s in range(len(off) - 1):
        a, b = int(off[s]), int(off[s + 1])
        L = b - a
        Q, K, V = (t[a:b].view(L, H, hd).transpose(0, 1) for t in (q, k, v))
        P = torch.softmax(Q @ K.transpose(1, 2) / math.sqrt(hd), -1)
        if mask is not None:
            P = P * mask[s] / (1 - p)
        outs.append((P @ V).transpose(0, 1).reshape(L, H * hd))
    return torch.cat(outs)


@pytest.mark.parametrize("resident", [False, True])
@pytest.mark.parametrize("H,hd,lens", [(32, 24, [130, 64, 7, 200, 1]), (16, 48, [65, 300]), (4, 64, [64, 40, 129]),
                                       (32, 24, [447]), (32, 24, [640, 3])])
def test_flash_fwd_bwd(cuda_lib, H, hd, lens, resident):
    """resident=True passes the longest sequence length, which selects the shared-memory-resident kernels
    (one CTA per (sequence, head)); False forces the tiled kernels.  Same results either way."""
    ml = max(lens) if resident else 0
    from b200vsgg import ops
    D, M = H * hd, sum(lens)
    off_h, off, bs, br = _plan(lens)
    g = torch.Generator(device=DEV).manual_seed(M)
    qkv = torch.randn(M, 3 * D, generator=g, device=DEV).bfloat16()
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    ctx = torch.empty(M, D, device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(M, H, device=DEV)
    ops.attn_flash_fwd(q, k, v, off, bs, br, H, hd, ctx, lse, max_len=ml)
    qf, kf, vf = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    ref = _ref(qf, kf, vf, off_h, H, hd)
    assert (ctx.float() - ref).abs().max().item() < 2 ** -7 * ref.abs().max().item() + 1e-3
    dctx = torch.randn(M, D, generator=g, device=DEV).bfloat16()
    ref.backward(dctx.float())
    dqkv = torch.empty(M, 3 * D, device=DEV, dtype=torch.bfloat16)
    ops.attn_flash_bwd(q, k, v, ctx, dctx, lse, off, bs, br, H, hd, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:],
                       max_len=ml)
    for name, got, want in (("dq", dqkv[:, :D], qf.grad), ("dk", dqkv[:, D:2 * D], kf.grad), ("dv", dqkv[:, 2 * D:], vf.grad)):
        tol = 2 ** -6 * want.abs().max().item() + 1e-3
        err = (got.float() - want).abs().max().item()
        assert err < tol, (name, err, tol)


@pytest.mark.parametrize("resident", [False, True])
def test_flash_dropout_mask_consistent_between_fwd_and_bwd(cuda_lib, resident):
    """One-hot values recover the dropped probabilities (ctx = P~ V); dq/dk/dv must equal autograd through
    P * mask / (1-p) with that mask."""
    from b200vsgg import ops
    H, hd, lens, p, seed = 2, 64, [64, 40], 0.25, 99
    ml = max(lens) if resident else 0
    D, M = H * hd, sum(lens)
    off_h, off, bs, br = _plan(lens)
    g = torch.Generator(device=DEV).manual_seed(5)
    qkv = torch.randn(M, 3 * D, generator=g, device=DEV).bfloat16()
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    onehot = torch.zeros(M, D, device=DEV, dtype=torch.bfloat16)
    for s in range(len(lens)):
        for j in range(lens[s]):
            onehot[int(off_h[s]) + j].view(H, hd)[:, j] = 1
    probe = torch.empty(M, D, device=DEV, dtype=torch.bfloat16)
    ops.attn_flash_fwd(q, k, onehot, off, bs, br, H, hd, probe, None, p, seed, max_len=ml)
    masks = []
    for s in range(len(lens)):
        a, L = int(off_h[s]), lens[s]
        masks.append((probe[a:a + L].float().view(L, H, hd)[:, :, :L].transpose(0, 1) != 0).float())
    kept = sum(m.sum().item() for m in masks) / sum(H * n * n for n in lens)
    assert 0.65 < kept < 0.85
    ctx = torch.empty(M, D, device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(M, H, device=DEV)
    ops.attn_flash_fwd(q, k, v, off, bs, br, H, hd, ctx, lse, p, seed, max_len=ml)
    qf, kf, vf = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    ref = _ref(qf, kf, vf, off_h, H, hd, masks, p)
    assert (ctx.float() - ref).abs().max().item() < 2 ** -7 * ref.abs().max().item() + 1e-3
    dctx = torch.randn(M, D, generator=g, device=DEV).bfloat16()
    ref.backward(dctx.float())
    dqkv = torch.empty(M, 3 * D, device=DEV, dtype=torch.bfloat16)
    ops.attn_flash_bwd(q, k, v, ctx, dctx, lse, off, bs, br, H, hd, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], p, seed,
                       max_len=ml)
    for got, want in ((dqkv[:, :D], qf.grad), (dqkv[:, D:2 * D], kf.grad), (dqkv[:, 2 * D:], vf.grad)):
        assert (got.float() - want).abs().max().item() < 2 ** -6 * want.abs().max().item() + 1e-3


def _plan128(lens):
    from b200vsgg.plan import attention_blocks
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    bs, br = attention_blocks(off, block=128)
    t = lambda a: torch.from_numpy(a).to(DEV)
    return t(bs), t(br)


@pytest.mark.parametrize("H,hd,lens", [(32, 24, [130, 64, 7, 200, 1]), (16, 48, [65, 300]), (4, 64, [64, 40, 129]),
                                       (32, 24, [447]), (32, 24, [640, 3]), (2, 24, [1500]), (3, 40, [128, 256, 129])])
def test_attn_tc_fwd_matches_torch_and_feeds_the_backward(cuda_lib, H, hd, lens):
    """tcgen05 forward (csrc/attn_tc.cu) vs the torch fp32 reference; its lse must equal the mma.sync forward's
    (the backward kernels consume it)."""
    from b200vsgg import ops
    D, M = H * hd, sum(lens)
    off_h, off, bs, br = _plan(lens)
    bs128, br128 = _plan128(lens)
    g = torch.Generator(device=DEV).manual_seed(M)
    qkv = (torch.randn(M, 3 * D, generator=g, device=DEV) * 1.5).bfloat16()
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    ctx = torch.full((M, D), float("nan"), device=DEV, dtype=torch.bfloat16)
    lse = torch.full((M, H), float("nan"), device=DEV)
    ops.attn_tc_fwd(q, k, v, off, bs128, br128, H, hd, ctx, lse)
    ref = _ref(q.float(), k.float(), v.float(), off_h, H, hd)
    err = (ctx.float() - ref).abs().max().item()
    assert err < 2 ** -7 * ref.abs().max().item() + 1e-3, err
    ctx0 = torch.empty_like(ctx)
    lse0 = torch.empty_like(lse)
    ops.attn_flash_fwd(q, k, v, off, bs, br, H, hd, ctx0, lse0)
    assert (lse - lse0).abs().max().item() < 2e-3
    # backward from the tcgen05 forward's outputs
    qf, kf, vf = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    dctx = torch.randn(M, D, generator=g, device=DEV).bfloat16()
    _ref(qf, kf, vf, off_h, H, hd).backward(dctx.float())
    dqkv = torch.empty(M, 3 * D, device=DEV, dtype=torch.bfloat16)
    ops.attn_flash_bwd(q, k, v, ctx, dctx, lse, off, bs, br, H, hd, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:])
    for name, got, want in (("dq", dqkv[:, :D], qf.grad), ("dk", dqkv[:, D:2 * D], kf.grad), ("dv", dqkv[:, 2 * D:], vf.grad)):
        tol = 2 ** -6 * want.abs().max().item() + 1e-3
        assert (got.float() - want).abs().max().item() < tol, name


def test_attn_tc_lazy_rescale_and_dropout(cuda_lib):
    """(a) Keys sorted so that the row maximum keeps growing by > 2^8 per block: exercises the O-rescale path.
    (b) Dropout: same mask function as the mma.sync kernels -> identical dropped outputs."""
    from b200vsgg import ops
    H, hd, lens = 2, 24, [700]
    D, M = H * hd, sum(lens)
    off_h, off, bs, br = _plan(lens)
    bs128, br128 = _plan128(lens)
    g = torch.Generator(device=DEV).manual_seed(1)
    q = torch.randn(M, D, generator=g, device=DEV).abs().bfloat16()
    k = (torch.randn(M, D, generator=g, device=DEV).abs() * torch.linspace(0.1, 8.0, M, device=DEV)[:, None]).bfloat16()
    v = torch.randn(M, D, generator=g, device=DEV).bfloat16()
    ctx = torch.empty(M, D, device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(M, H, device=DEV)
    ops.attn_tc_fwd(q, k, v, off, bs128, br128, H, hd, ctx, lse)
    ref = _ref(q.float(), k.float(), v.float(), off_h, H, hd)
    assert torch.isfinite(ctx.float()).all()
    assert (ctx.float() - ref).abs().max().item() < 2 ** -7 * ref.abs().max().item() + 1e-3
    p, seed = 0.25, 77
    qkv = torch.randn(M, 3 * D, generator=g, device=DEV).bfloat16()
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    a, b = torch.empty_like(ctx), torch.empty_like(ctx)
    ops.attn_tc_fwd(q, k, v, off, bs128, br128, H, hd, a, lse, p, seed)
    ops.attn_flash_fwd(q, k, v, off, bs, br, H, hd, b, None, p, seed)
    assert (a.float() - b.float()).abs().max().item() < 2 ** -6 * b.float().abs().max().item() + 1e-3


@pytest.mark.parametrize("H,hd,lens", [(32, 24, [130, 64, 7, 200, 1]), (16, 48, [65, 300]), (4, 64, [64, 40, 129]),
                                       (32, 24, [447]), (2, 24, [1500]), (3, 40, [128, 256, 129])])
def test_attn_tc_bwd_matches_torch(cuda_lib, H, hd, lens):
    """tcgen05 backward (csrc/attn_tc_bwd.cu) vs autograd through the torch fp32 reference."""
    from b200vsgg import ops
    D, M = H * hd, sum(lens)
    off_h, off, bs, br = _plan(lens)
    bs128, br128 = _plan128(lens)
    g = torch.Generator(device=DEV).manual_seed(M + 1)
    qkv = torch.randn(M, 3 * D, generator=g, device=DEV).bfloat16()
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    ctx = torch.empty(M, D, device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(M, H, device=DEV)
    ops.attn_tc_fwd(q, k, v, off, bs128, br128, H, hd, ctx, lse)
    qf, kf, vf = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    dctx = torch.randn(M, D, generator=g, device=DEV).bfloat16()
    _ref(qf, kf, vf, off_h, H, hd).backward(dctx.float())
    dqkv = torch.full((M, 3 * D), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.attn_tc_bwd(q, k, v, ctx, dctx, lse, off, bs128, br128, H, hd, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:])
    for name, got, want in (("dq", dqkv[:, :D], qf.grad), ("dk", dqkv[:, D:2 * D], kf.grad), ("dv", dqkv[:, 2 * D:], vf.grad)):
        tol = 2 ** -6 * want.abs().max().item() + 1e-3
        err = (got.float() - want).abs().max().item()
        assert err < tol, (name, err, tol)


def test_attn_tc_bwd_dropout_matches_mma_sync_backward(cuda_lib):
    """Same counters -> same masks: with dropout the tcgen05 backward equals the mma.sync backward (which is itself
    checked against autograd with the recovered mask above)."""
    from b200vsgg import ops
    H, hd, lens, p, seed = 4, 24, [300, 77, 129], 0.25, 31
    D, M = H * hd, sum(lens)
    off_h, off, bs, br = _plan(lens)
    bs128, br128 = _plan128(lens)
    g = torch.Generator(device=DEV).manual_seed(2)
    qkv = torch.randn(M, 3 * D, generator=g, device=DEV).bfloat16()
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    ctx = torch.empty(M, D, device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(M, H, device=DEV)
    ops.attn_tc_fwd(q, k, v, off, bs128, br128, H, hd, ctx, lse, p, seed)
    dctx = torch.randn(M, D, generator=g, device=DEV).bfloat16()
    a = torch.empty(M, 3 * D, device=DEV, dtype=torch.bfloat16)
    b = torch.empty_like(a)
    ops.attn_tc_bwd(q, k, v, ctx, dctx, lse, off, bs128, br128, H, hd, a[:, :D], a[:, D:2 * D], a[:, 2 * D:], p, seed)
    ops.attn_flash_bwd(q, k, v, ctx, dctx, lse, off, bs, br, H, hd, b[:, :D], b[:, D:2 * D], b[:, 2 * D:], p, seed)
    for name, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
        x, y = a[:, sl].float(), b[:, sl].float()
        assert (x - y).abs().max().item() < 2 ** -6 * y.abs().max().item() + 1e-3, name
