"""SGCls-train object branch on the CUDA path vs
  (1) golden vectors of the UNMODIFIED reference (tests/golden/sgcls_*.pt, oracle/make_golden_sgcls.py),
  (2) the CPU oracle forward and backward, single video and batches.
Tolerances (bf16 tensor-core operands, fp32 accumulation): distributions max-abs <= 4e-3 (GMM head:
probabilities) / linear-head logits <= 3e-2 * max|ref|; object_features <= 3e-2 * max|ref|; parameter
gradients rel-L2 <= 6e-2 for the head and BatchNorm1d(1024) tensors, <= 0.15 for everything upstream of the
`intermediate` ReLU gate: Linear(2376->1024) has bf16 operands, so ~0.1 % of the 1024 x O gates (pre-activations
within bf16 rounding of zero) open differently from the fp32 oracle — measured 31 of 27 648 on sgcls_track_gmm
(tools/debug_sgcls.py) — and with only ~27 boxes every flipped gate is a full-magnitude element of the gated
gradient: sqrt(0.0011 / 0.5) = 4.7 % rel-L2 there, up to 10 % on column-sum (bias) gradients.  Same effect and
same bound as the mask branch's ReLU gates (tests/test_tempura_gpu.py)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = ["sgcls_track_gmm", "sgcls_track_linear", "sgcls_notrack_gmm"]
DIST_TOL, FEAT_REL_TOL, GRAD_REL_TOL, GATED_GRAD_REL_TOL = 4e-3, 3e-2, 6e-2, 0.15


def _clone(e, device=None):
    out = {}
    for k, v in e.items():
        if isinstance(v, torch.Tensor):
            out[k] = v.clone().to(device) if device is not None else v.clone()
        elif k == "indices":
            out[k] = [ix.clone().to(device) if device is not None else ix.clone() for ix in v]
        else:
            out[k] = v
    return out


def _setup(name):
    from b200vsgg import objbranch, synthetic, tempura
    from oracle.tempura_oracle import TempuraOracle
    gold = torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)
    classes = synthetic.ag_object_classes()
    m = tempura.TEMPURA(obj_classes=classes, **gold["model_kw"])
    synthetic.seeded_init_(m)
    o = TempuraOracle(obj_classes=classes, dropout=0.0, **gold["model_kw"])
    o.load_state_dict(m.state_dict(), strict=True)
    vid = gold["case"]["video_index"]
    entry = synthetic.add_sgcls_inputs(synthetic.make_video_entry(**gold["case"]), vid)
    objbranch.get_sequence(entry, None, None, "sgcls")
    return gold, entry, m.cuda().train(), o.train()


def test_get_sequence_is_bit_exact(cuda_lib):
    for name in CASES:
        gold, entry, _, _ = _setup(name)
        assert len(entry["indices"]) == len(gold["indices"])
        for a, b in zip(entry["indices"], gold["indices"]):
            assert torch.equal(torch.as_tensor(a).long().cpu(), b)


def test_obj_tokens_kernel_matches_torch(cuda_lib):
    """b200vsgg_obj_tokens_{fwd,bwd} vs plain torch fp32 on permuted rows with a position table."""
    from b200vsgg import ops
    g = torch.Generator().manual_seed(0)
    O, V = 53, 3
    feats = torch.randn(O, 2048, generator=g)
    dist = torch.softmax(torch.randn(O, 36, generator=g), 1)
    E = torch.randn(36, 200, generator=g)
    boxes = torch.cat([torch.arange(O)[:, None].float(), torch.rand(O, 2, generator=g) * 200,
                       200 + torch.rand(O, 2, generator=g) * 200], 1)
    vob = torch.sort(torch.randint(0, V, (O,), generator=g)).values
    mean, rstd = torch.randn(V, 4, generator=g) * 50 + 200, torch.rand(V, 4, generator=g) * 0.02 + 0.01
    gamma, beta = torch.randn(4, generator=g), torch.randn(4, generator=g)
    wp, bp = torch.randn(128, 4, generator=g), torch.randn(128, generator=g)
    pe = torch.randn(40, 2376, generator=g)
    src = torch.randperm(O, generator=g)
    pos = torch.randint(0, 40, (O,), generator=g)
    P = dict(embed=E, bn_gamma=gamma, bn_beta=beta, wp=wp, bp=bp)
    P = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    wh = boxes[:, 3:5] - boxes[:, 1:3] + 1.0
    cs = torch.cat([boxes[:, 1:3] + 0.5 * wh, wh], 1)
    ybn = (cs - mean[vob]) * rstd[vob] * P["bn_gamma"] + P["bn_beta"]
    x0 = torch.cat([feats, dist @ P["embed"], torch.relu(ybn @ P["wp"].t() + P["bp"])], 1)
    ref = x0[src] + pe[pos]
    dx = torch.randn(O, 2376, generator=g)
    ref.backward(dx)
    cu = lambda t: t.detach().cuda().contiguous()
    args = dict(features=cu(feats), dist=cu(dist), embed=cu(E), boxes=cu(boxes), bn_mean=cu(mean), bn_rstd=cu(rstd),
                bn_gamma=cu(gamma), bn_beta=cu(beta), video_of_box=cu(vob.int()), wp=cu(wp), bp=cu(bp), pe=cu(pe),
                src=cu(src.int()), pos=cu(pos.int()), rows=O)
    x32 = torch.empty(O, 2376, device="cuda")
    xb = torch.empty(O, 2376, device="cuda", dtype=torch.bfloat16)
    ops.obj_tokens_fwd(args, x32, xb)
    assert (x32.cpu() - ref.detach()).abs().max().item() <= 2e-4 * ref.abs().max().item()
    assert (xb.float().cpu() - ref.detach()).abs().max().item() <= 1e-2 * ref.abs().max().item()
    z = lambda *s: torch.zeros(*s, device="cuda")
    dE, dwp, dbp, dg, db = z(36, 200), z(128, 4), z(128), z(4), z(4)
    ops.obj_tokens_bwd(args, cu(dx), dE, dwp, dbp, dg, db)
    for got, want in ((dE, P["embed"].grad), (dwp, P["wp"].grad), (dbp, P["bp"].grad), (dg, P["bn_gamma"].grad),
                      (db, P["bn_beta"].grad)):
        rel = (got.cpu() - want).norm().item() / want.norm().item()
        assert rel <= 1e-4, rel
    # identity order, no position term (the non-tracking path)
    args2 = dict(args, src=None, pos=None, pe=None)
    ops.obj_tokens_fwd(args2, x32, None)
    assert (x32.cpu() - x0.detach()).abs().max().item() <= 2e-4 * x0.abs().max().item()


@pytest.mark.parametrize("name", CASES)
def test_forward_matches_reference_golden(cuda_lib, name):
    gold, entry, m, _ = _setup(name)
    m.dropout_p = 0.0
    m.object_classifier.dropout_p = 0.0
    m.gmm_eps = gold["eps"]
    state = {k: v.clone() for k, v in m.state_dict().items()}
    with torch.no_grad():
        out = m(_clone(entry, "cuda"), phase="train")
    n = 0
    for key, ref in gold.items():
        if not isinstance(key, str) or not key.startswith("train_eps/"):
            continue
        k = key.split("/")[1]
        got = out[k].float().cpu()
        tol = DIST_TOL if (k != "distribution" or gold["model_kw"]["obj_head"] == "gmm") else FEAT_REL_TOL * ref.abs().max().item()
        err = (got - ref).abs().max().item()
        assert err <= tol, (key, err, tol)
        n += 1
    assert n == 4
    if "train_seed99/object_features" in gold:      # independent of the GMM noise
        ref = gold["train_seed99/object_features"]
        err = (out["object_features"].float().cpu() - ref).abs().max().item()
        assert err <= FEAT_REL_TOL * ref.abs().max().item(), err
    # BatchNorm running statistics of the object branch after one train step
    after = m.state_dict()
    for key, ref in gold.items():
        if isinstance(key, str) and key.startswith("bn_after/"):
            k = key.split("/", 1)[1]
            got = after[k].float().cpu()
            assert (got - ref).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item()), key
            assert not torch.equal(after[k], state[k])


@pytest.mark.parametrize("name", CASES)
def test_backward_matches_oracle(cuda_lib, name):
    from b200vsgg import synthetic, tempura
    from oracle.tempura_oracle import object_loss, tempura_losses
    gold, entry, m, o = _setup(name)
    eps = gold["eps"]
    att, spa, con = synthetic.build_gt_tensors(entry)
    po = o(_clone(entry), phase="train", eps=eps)
    lo = sum(tempura_losses(po, att, spa, con).values()) + object_loss(po, 0.5)
    lo.backward()
    m.dropout_p = 0.0
    m.object_classifier.dropout_p = 0.0
    m.gmm_eps = eps
    pm = m(_clone(entry, "cuda"), phase="train")
    losses = tempura.tempura_loss(pm, m.last_plan, eos_coef=0.5)
    assert "object_loss" in losses
    lm = sum(losses.values())
    lm.backward()
    assert abs(lm.item() - lo.item()) < 2e-3 * abs(lo.item()), (lm.item(), lo.item())
    og = dict(o.named_parameters())
    errs = []
    norms = sorted(v.grad.norm().item() for k, v in og.items() if k.startswith("object_classifier.") and v.grad is not None)
    typical = norms[len(norms) // 2]
    for pname, p in m.named_parameters():
        if not pname.startswith("object_classifier."):
            continue
        ref = og[pname].grad
        if ref is None:
            assert p.grad is None or p.grad.abs().max().item() == 0, pname
            continue
        assert p.grad is not None, pname
        denom = ref.norm().item()
        got = p.grad.float().cpu()
        if denom < 1e-7:   # exactly-zero gradients (column sums behind a BatchNorm / softmax key bias): stay negligible
            assert got.norm().item() < 0.05 * typical, (pname, got.norm().item(), typical)
            continue
        errs.append(((got - ref).norm().item() / denom, pname))
    errs.sort(reverse=True)
    print("largest object-branch gradient rel-L2 errors:", errs[:8])
    assert len(errs) >= (40 if gold["model_kw"]["tracking"] else 8)
    for rel, pname in errs:
        below_gate = not (pname.startswith("object_classifier.decoder_lin") or pname.startswith("object_classifier.intermediate.1"))
        assert rel <= (GATED_GRAD_REL_TOL if below_gate else GRAD_REL_TOL), (pname, rel, errs[:6])


def test_batched_videos_equal_per_video_runs(cuda_lib):
    """A batch of videos = independent videos: class sequences and BatchNorm statistics never cross a video."""
    from b200vsgg import objbranch, synthetic, tempura
    gold, _, m, _ = _setup("sgcls_track_gmm")
    m.dropout_p = 0.0
    m.object_classifier.dropout_p = 0.0
    entries = []
    for i, (f, ppf) in enumerate([(4, (1, 3)), (6, (2, 5)), (3, 4)]):
        e = synthetic.add_sgcls_inputs(synthetic.make_video_entry(20 + i, f, ppf), 20 + i)
        objbranch.get_sequence(e, None, None, "sgcls")
        entries.append(e)
    K, C = gold["model_kw"]["K"], 37
    g = torch.Generator().manual_seed(1)
    eps_obj = [torch.randn(K, e["labels"].shape[0], C, generator=g) for e in entries]
    state = {k: v.clone() for k, v in m.state_dict().items()}
    singles = []
    with torch.no_grad():
        for e, ep in zip(entries, eps_obj):
            m.load_state_dict(state)
            m.gmm_eps = {"object": ep}
            singles.append(m(_clone(e, "cuda"), phase="train"))
        m.load_state_dict(state)
        m.gmm_eps = {"object": torch.cat(eps_obj, 1)}
        outb = m(tempura.collate_entries([_clone(e, "cuda") for e in entries]), phase="train")
    cat = torch.cat([s["distribution"] for s in singles])
    assert (outb["distribution"] - cat).abs().max().item() <= 1e-3
    catf = torch.cat([s["object_features"] for s in singles])
    assert (outb["object_features"] - catf).abs().max().item() <= 1e-2 * catf.abs().max().item()
    loss_b = tempura.tempura_loss(outb, m.last_plan)["object_loss"].item()
    loss_s = sum(tempura.tempura_loss(s, None)["object_loss"].item() for s in singles) / 3
    assert abs(loss_b - loss_s) <= 1e-3 * abs(loss_s)


def test_train_step_with_dropout_is_finite(cuda_lib):
    from b200vsgg import objbranch, synthetic, tempura
    _, _, m, _ = _setup("sgcls_track_gmm")
    entries = []
    for i in range(3):
        e = synthetic.add_sgcls_inputs(synthetic.make_video_entry(40 + i, 6, (2, 6), device="cuda"), 40 + i)
        objbranch.get_sequence(e, None, None, "sgcls")
        entries.append(e)
    torch.manual_seed(0)
    pred = m(tempura.collate_entries(entries), phase="train")
    loss = sum(tempura.tempura_loss(pred, m.last_plan).values())
    loss.backward()
    assert torch.isfinite(loss)
    n = 0
    for pname, p in m.named_parameters():
        if pname.startswith("object_classifier.") and "mem_attention" not in pname:
            assert p.grad is not None and torch.isfinite(p.grad).all(), pname
            n += 1
    assert n > 40


@pytest.mark.parametrize("hd,n_heads,lens", [(304, 8, [1, 70, 5, 33, 1, 130]), (48, 4, [17, 1, 64, 32]), (242, 2, [40, 9])])
def test_attn_rows_kernel_matches_torch(cuda_lib, hd, n_heads, lens):
    """b200vsgg_attn_rows_{fwd,bwd} (any sequence length, head_dim <= 320) vs fp32 torch attention per segment.
    Tolerances: bf16 in/out -> ctx max-abs <= 2e-2 * max|ref|, gradients rel-L2 <= 2e-2."""
    from b200vsgg import ops
    g = torch.Generator().manual_seed(hd + len(lens))
    M, D = sum(lens), n_heads * hd
    q, k, v, do = (torch.randn(M, D, generator=g).bfloat16() for _ in range(4))
    scale = 0.7 * hd ** -0.5
    off = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32)
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    outs = []
    for s in range(len(lens)):
        a, b = int(off[s]), int(off[s + 1])
        Q, K, V = (t[a:b].view(b - a, n_heads, hd).transpose(0, 1) for t in (qf, kf, vf))
        P = torch.softmax(Q @ K.transpose(1, 2) * scale, -1)
        outs.append((P @ V).transpose(0, 1).reshape(b - a, D))
    ref = torch.cat(outs)
    ref.backward(do.float())
    cu = lambda t: t.cuda().contiguous()
    qc, kc, vc, doc, offc = cu(q), cu(k), cu(v), cu(do), cu(off)
    ctx = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(M, n_heads, device="cuda")
    ops.attn_rows_fwd(qc, kc, vc, offc, len(lens), n_heads, hd, ctx, lse, scale=scale)
    assert (ctx.float().cpu() - ref.detach()).abs().max().item() <= 2e-2 * ref.abs().max().item()
    dq, dk, dv = (torch.empty(M, D, device="cuda", dtype=torch.bfloat16) for _ in range(3))
    ops.attn_rows_bwd(qc, kc, vc, ctx, doc, lse, offc, len(lens), n_heads, hd, dq, dk, dv, scale=scale)
    for got, want in ((dq, qf.grad), (dk, kf.grad), (dv, vf.grad)):
        rel = (got.float().cpu() - want).norm().item() / want.norm().item()
        assert rel <= 2e-2, rel
    # dropout: forward/backward regenerate the same mask -> finite differences are not needed, check the identity
    # sum(dV) = sum_i (sum_j P~_ij) dO_i is consistent with the forward's P~ on V = ones
    ones = torch.ones_like(vc)
    ctx1 = torch.empty_like(ctx)
    ops.attn_rows_fwd(qc, kc, ones, offc, len(lens), n_heads, hd, ctx1, lse, drop_p=0.3, seed=5, scale=scale)
    rowsum = ctx1.float().view(M, n_heads, hd)[:, :, 0]                     # sum_j P~_ij per (row, head)
    ops.attn_rows_bwd(qc, kc, ones, ctx1, doc, lse, offc, len(lens), n_heads, hd, dq, dk, dv, drop_p=0.3, seed=5, scale=scale)
    lhs = dv.float().view(M, n_heads, hd).sum(0)                            # sum_j dV_j  [heads, hd]
    rhs = (rowsum[:, :, None] * doc.float().view(M, n_heads, hd)).sum(0)
    assert (lhs - rhs).abs().max().item() <= 3e-2 * rhs.abs().max().item()
    assert 0.5 < rowsum.mean().item() < 1.5 and (rowsum == 0).float().mean().item() < 0.5


@pytest.mark.parametrize("selection", ["manual"])
def test_object_memory_with_tracking_matches_oracle(cuda_lib, selection):
    """Object-memory hallucinator of the tracking branch (lib/tempura.py:160-183,213-222: single-head attention of the
    class-sequence features over the per-class object memory, blended by the fixed selector — the reference's LEARNED
    selector is nn.Linear(1024, 1) (lib/tempura.py:101) and cannot take the 2376-wide tracking features at all) with a NON-EMPTY
    memory bank: object_mem_features and the object distribution vs the CPU oracle, and the gradients of the memory
    attention / selector / sequence encoder vs the oracle's autograd."""
    from b200vsgg import objbranch, synthetic, tempura
    from oracle.tempura_oracle import TempuraOracle, object_loss, tempura_losses
    gold = torch.load(os.path.join(GOLDEN, "sgcls_track_gmm.pt"), weights_only=False)
    kw = dict(gold["model_kw"], obj_mem_compute=True, selection=selection)
    classes = synthetic.ag_object_classes()
    m = tempura.TEMPURA(obj_classes=classes, **kw)
    synthetic.seeded_init_(m)
    o = TempuraOracle(obj_classes=classes, dropout=0.0, **kw)
    o.load_state_dict(m.state_dict(), strict=True)
    bank = 0.5 * torch.randn(len(classes) - 1, 2376, generator=torch.Generator().manual_seed(77))
    m.object_classifier.obj_memory = bank.cuda()
    o.object_classifier.obj_memory = bank.clone()
    vid = gold["case"]["video_index"]
    entry = synthetic.add_sgcls_inputs(synthetic.make_video_entry(**gold["case"]), vid)
    objbranch.get_sequence(entry, None, None, "sgcls")
    m, o = m.cuda().train(), o.train()
    m.dropout_p = m.object_classifier.dropout_p = 0.0
    m.gmm_eps = gold["eps"]
    att, spa, con = synthetic.build_gt_tensors(entry)
    po = o(_clone(entry), phase="train", eps=gold["eps"])
    (sum(tempura_losses(po, att, spa, con).values()) + object_loss(po, 0.5)).backward()
    pm = m(_clone(entry, "cuda"), phase="train")
    sum(tempura.tempura_loss(pm, m.last_plan, eos_coef=0.5).values()).backward()
    # the memory changed the features (the test would be vacuous otherwise)
    assert (po["object_mem_features"] - po["object_features"]).abs().max().item() > 1e-2
    ref = po["object_mem_features"].detach()
    err = (pm["object_mem_features"].float().cpu() - ref).abs().max().item()
    assert err <= FEAT_REL_TOL * ref.abs().max().item(), err
    assert (pm["distribution"].float().cpu() - po["distribution"].detach()).abs().max().item() <= DIST_TOL
    og = dict(o.named_parameters())
    checked = 0
    for pname, p in m.named_parameters():
        if not (pname.startswith("object_classifier.mem_attention") or pname.startswith("object_classifier.selector")):
            continue
        r = og[pname].grad
        assert r is not None and p.grad is not None, pname
        rel = (p.grad.float().cpu() - r).norm().item() / max(r.norm().item(), 1e-12)
        assert rel <= GATED_GRAD_REL_TOL, (pname, rel)
        checked += 1
    assert checked >= (2 if selection == "manual" else 4)


@pytest.mark.parametrize("selection", ["manual", "learned"])
def test_object_memory_without_tracking_matches_oracle(cuda_lib, selection):
    """Object memory of the NON-tracking branch (lib/tempura.py:217-221: hallucinator on the 1024-wide `intermediate`
    output; here the learned selector nn.Linear(1024, 1) fits) with a non-empty bank: features, distribution and the
    gradients of the memory attention / selector vs the CPU oracle."""
    from b200vsgg import objbranch, synthetic, tempura
    from oracle.tempura_oracle import TempuraOracle, object_loss, tempura_losses
    gold = torch.load(os.path.join(GOLDEN, "sgcls_notrack_gmm.pt"), weights_only=False)
    kw = dict(gold["model_kw"], obj_mem_compute=True, selection=selection)
    classes = synthetic.ag_object_classes()
    m = tempura.TEMPURA(obj_classes=classes, **kw)
    synthetic.seeded_init_(m)
    o = TempuraOracle(obj_classes=classes, dropout=0.0, **kw)
    o.load_state_dict(m.state_dict(), strict=True)
    bank = 0.5 * torch.randn(len(classes) - 1, 1024, generator=torch.Generator().manual_seed(78))
    m.object_classifier.obj_memory = bank.cuda()
    o.object_classifier.obj_memory = bank.clone()
    vid = gold["case"]["video_index"]
    entry = synthetic.add_sgcls_inputs(synthetic.make_video_entry(**gold["case"]), vid)
    objbranch.get_sequence(entry, None, None, "sgcls")
    m, o = m.cuda().train(), o.train()
    m.dropout_p = m.object_classifier.dropout_p = 0.0
    m.gmm_eps = gold["eps"]
    att, spa, con = synthetic.build_gt_tensors(entry)
    po = o(_clone(entry), phase="train", eps=gold["eps"])
    (sum(tempura_losses(po, att, spa, con).values()) + object_loss(po, 0.5)).backward()
    pm = m(_clone(entry, "cuda"), phase="train")
    sum(tempura.tempura_loss(pm, m.last_plan, eos_coef=0.5).values()).backward()
    assert (po["object_mem_features"] - po["object_features"]).abs().max().item() > 1e-2
    ref = po["object_mem_features"].detach()
    assert (pm["object_mem_features"].float().cpu() - ref).abs().max().item() <= FEAT_REL_TOL * ref.abs().max().item()
    assert (pm["distribution"].float().cpu() - po["distribution"].detach()).abs().max().item() <= DIST_TOL
    og = dict(o.named_parameters())
    checked = 0
    for pname, p in m.named_parameters():
        if not (pname.startswith("object_classifier.mem_attention") or pname.startswith("object_classifier.selector")):
            continue
        r = og[pname].grad
        assert r is not None and p.grad is not None, pname
        rel = (p.grad.float().cpu() - r).norm().item() / max(r.norm().item(), 1e-12)
        assert rel <= GATED_GRAD_REL_TOL, (pname, rel)
        checked += 1
    assert checked >= (2 if selection == "manual" else 4)


@pytest.mark.parametrize("name", ["sgcls_track_gmm", "sgcls_notrack_gmm"])
def test_uncertainty_pass_matches_oracle(cuda_lib, name):
    """`model(entry, unc=True)` in eval mode — the call of the trainer's uncertainty bookkeeping (tools/utils/Uncertainty.py:
    95-100; lib/tempura.py:226-228 for the object head): test-phase object distribution (background class dropped before the
    softmax: 36 columns), aleatoric / epistemic uncertainties of the object head and of the three relation heads vs the CPU
    oracle."""
    gold, entry, m, o = _setup(name)
    m.eval()
    o.eval()
    with torch.no_grad():
        out = m(_clone(entry, "cuda"), phase="train", unc=True)
        ref = o(_clone(entry), phase="train", unc=True)
    assert out["distribution"].shape == ref["distribution"].shape and out["distribution"].shape[1] == 36
    for k in ("distribution", "obj_al_uc", "obj_ep_uc", "attention_al_uc", "spatial_al_uc", "contacting_al_uc",
              "attention_ep_uc", "spatial_ep_uc", "contacting_ep_uc"):
        err = (out[k].float().cpu() - ref[k]).abs().max().item()
        # object uncertainties: sums of sigmoid(variance logit) read off the bf16 `intermediate` output (1024-wide, un-
        # normalised in the non-tracking branch: measured 1.7e-2 there, < 4e-3 with tracking); values live in [0, 1]
        tol = 2.5e-2 if k in ("obj_al_uc", "obj_ep_uc") else DIST_TOL
        assert err <= tol, (k, err)
    assert torch.equal(out["pred_labels"].cpu(), entry["labels"])
