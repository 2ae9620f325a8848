"""The two row savings of the temporal decoder (b200vsgg.tempura.DEC_FIRST_ON_PAIRS / DEC_LATTER_ONLY) are legal in the
reference's own algorithm: shown here on the oracle restatement of tools/utils/transformer.py:177-246 (pinned against the
unmodified reference by oracle/make_golden.py) in float64, where "equal" means equal to round-off.

  1. 'latter' read-out (transformer.py:236-242): of the last decoder layer's window rows only the second-frame rows (and
     the first-frame rows of window 0) are read.  Replacing everything the last layer computes AFTER its attention for
     the other rows by zeros changes neither the output nor any parameter gradient.
  2. Layer-1 window tokens are copies of pair rows plus one of two position rows (transformer.py:203-215):
     (x + pos) Wqk^T + b == (x Wqk^T + b) + pos Wqk^T, so the projections can run on the pair rows.
The CUDA path's implementation of both is compared with its own dense schedule and with the oracle in
tests/test_tempura_gpu.py."""
import torch
import torch.nn.functional as F

from oracle.tempura_oracle import STTranOracle, frame_segments


def _model_and_input(seed=0, counts=(3, 1, 4, 2, 5)):
    torch.manual_seed(seed)
    m = STTranOracle(enc_layer_num=1, dec_layer_num=3, embed_dim=32, nhead=4, dim_feedforward=48, dropout=0.0,
                     mem_compute=True, mem_fusion="late", selection="manual").double()
    with torch.no_grad():                      # de-twin the deep-copied decoder layers
        for p in m.parameters():
            p.add_(0.05 * torch.randn_like(p))
    im_idx = torch.repeat_interleave(torch.arange(len(counts)), torch.tensor(counts)).double()
    x = torch.randn(int(sum(counts)), 32, dtype=torch.float64)
    return m, x, im_idx


def _kept_mask(im_idx):
    """[F-1, 2l] bool: window slots the 'latter' read-out keeps (oracle forward, same construction)."""
    counts, _ = frame_segments(im_idx)
    l = int(counts.max())
    ar2 = torch.arange(2 * l)
    wlen = counts[:-1] + counts[1:]
    wvalid = ar2[None, :] < wlen[:, None]
    second = (ar2[None, :] >= counts[:-1, None]) & wvalid
    kept = second.clone()
    kept[0] |= ar2 < counts[0]
    return kept


def _run(m, x, im_idx):
    m.zero_grad()
    out, _, _ = m(x, im_idx)
    w = torch.linspace(-1.0, 1.0, out.numel(), dtype=out.dtype).view_as(out)
    (out * w).sum().backward()
    return out.detach().clone(), {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}


def test_last_layer_rows_outside_the_latter_readout_are_dead():
    m, x, im_idx = _model_and_input()
    ref_out, ref_g = _run(m, x, im_idx)
    kept = _kept_mask(im_idx)
    last = m.global_attention.layers[-1]

    def pruned_forward(xw, key_pad, pos):
        qk = xw + pos
        a = last.multihead2.attend(qk, qk, xw, key_pad, 0.0, False)      # keys / values: every row
        y = torch.zeros_like(xw)
        t = last.norm3(xw[kept] + a[kept])                               # kept rows only from here on
        y[kept] = t + last.linear2(F.relu(last.linear1(t)))
        return y

    last.forward = pruned_forward
    out, g = _run(m, x, im_idx)
    assert torch.equal(out, ref_out)                                     # row-wise arithmetic: bit-identical
    assert g.keys() == ref_g.keys()
    for n in ref_g:
        assert torch.allclose(g[n], ref_g[n], rtol=1e-11, atol=1e-12), n
    assert int(kept.sum()) == x.shape[0]                                 # exactly one kept window row per pair


def test_first_layer_projections_commute_with_the_window_gather():
    m, x, im_idx = _model_and_input(seed=1)
    ref_out, ref_g = _run(m, x, im_idx)
    first = m.global_attention.layers[0]
    mha = first.multihead2
    D = mha.dim

    def on_pairs_forward(xw, key_pad, pos):
        # xw = gathered pair rows: projecting xw row by row IS projecting the pair rows and gathering; the position
        # rows enter through their own product with Wqk
        w, b = mha.in_proj_weight, mha.in_proj_bias
        q = F.linear(xw, w[:D], b[:D]) + F.linear(pos, w[:D])
        k = F.linear(xw, w[D:2 * D], b[D:2 * D]) + F.linear(pos, w[D:2 * D])
        v = F.linear(xw, w[2 * D:], b[2 * D:])
        B, L, _ = q.shape
        H, hd = mha.heads, D // mha.heads
        qh = q.view(B, L, H, hd).transpose(1, 2) * hd ** -0.5
        kh, vh = k.view(B, L, H, hd).transpose(1, 2), v.view(B, L, H, hd).transpose(1, 2)
        s = (qh @ kh.transpose(-1, -2)).masked_fill(key_pad[:, None, None, :], float("-inf"))
        a = mha.out_proj((torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(B, L, D))
        t = first.norm3(xw + a)
        return t + first.linear2(F.relu(first.linear1(t)))

    first.forward = on_pairs_forward
    out, g = _run(m, x, im_idx)
    assert torch.allclose(out, ref_out, rtol=1e-11, atol=1e-12)
    for n in ref_g:
        assert torch.allclose(g[n], ref_g[n], rtol=1e-10, atol=1e-11), n
