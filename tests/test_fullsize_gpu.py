"""Size-independent properties at BASELINE.json's full single-GPU size (configs[1]: 64 synthetic AG videos x 32
frames x 6-10 pairs, N ~ 16.4 k pairs, 3.4 GB of inputs) — the oracle is too slow here, so the checks are the
invariants the path must satisfy at any size:
  * every distribution is finite and inside [0,1]; attention mixtures sum to 1 per pair
  * determinism: the same batch twice gives bit-identical outputs (eval mode)
  * videos are independent units: reversing the video order of the batch permutes the per-video outputs and changes
    nothing else (<= 1e-3: bf16 GEMM tiles see different row neighbours, no cross-video term exists)
  * segment plan: frame offsets from the device kernel equal the host bincount, window/latter indices are in range
    and the 'latter' gather is a bijection onto the pair rows
  * SGCls object branch at the same size: class sequences partition the boxes, GMM object posteriors sum to 1."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
KW = dict(mode="predcls", attention_class_num=3, spatial_class_num=6, contact_class_num=17, enc_layer_num=1,
          dec_layer_num=3, obj_mem_compute=False, rel_mem_compute="joint", mem_fusion="late", selection="manual",
          selection_lambda=0.5, take_obj_mem_feat=False, obj_head="gmm", rel_head="gmm", K=6, tracking=False)
V, FRAMES = 64, 32


def _entries(with_sgcls=False):
    from b200vsgg import objbranch, synthetic
    out = []
    for i in range(V):
        e = synthetic.make_video_entry(i, FRAMES, (6, 10), device="cuda", with_gt=False, big_on_device="cuda")
        if with_sgcls:
            synthetic.add_sgcls_inputs(e, i)
            objbranch.get_sequence(e, None, None, "sgcls")
        out.append(e)
    return out


def test_fullsize_predcls_properties(cuda_lib):
    from b200vsgg import ops, synthetic, tempura
    m = tempura.TEMPURA(obj_classes=synthetic.ag_object_classes(), **KW)
    synthetic.seeded_init_(m, 11)
    m = m.cuda().eval()
    entries = _entries()
    batch = tempura.collate_entries(entries)
    N = batch["pair_idx"].shape[0]
    assert N > 15000
    keys = ("attention_distribution", "spatial_distribution", "contacting_distribution")
    with torch.no_grad():
        a = m(dict(batch), phase="test")
        plan = m.last_plan
        b = m(dict(batch), phase="test")
        rev = m(tempura.collate_entries(entries[::-1]), phase="test")
    for k in keys:
        assert torch.isfinite(a[k]).all() and (a[k] >= 0).all() and (a[k] <= 1).all(), k
        assert torch.equal(a[k], b[k]), "non-deterministic " + k
    assert (a["attention_distribution"].sum(1) - 1).abs().max().item() <= 1e-3
    # reversed video order: per-video blocks come back in reverse
    ppv = [e["pair_idx"].shape[0] for e in entries]
    off = np.concatenate([[0], np.cumsum(ppv)])
    off_r = np.concatenate([[0], np.cumsum(ppv[::-1])])
    for k in keys:
        for v in (0, 17, V - 1):
            blk = a[k][off[v]:off[v + 1]]
            blk_r = rev[k][off_r[V - 1 - v]:off_r[V - v]]
            assert (blk - blk_r).abs().max().item() <= 1e-3, (k, v)
    # segment plan invariants at full size
    counts = torch.bincount(batch["im_idx"].long()).cpu().numpy()
    dev_off = ops.frame_offsets(batch["im_idx"].contiguous(), V * FRAMES).cpu().numpy()
    assert np.array_equal(dev_off, np.concatenate([[0], np.cumsum(counts)]))
    assert plan.N == N and plan.F == V * FRAMES and plan.W == V * (FRAMES - 1)
    assert plan.win_src_h.min() >= 0 and plan.win_src_h.max() < N and plan.M2 == plan.win_src_h.shape[0]
    assert np.array_equal(np.sort(plan.win_src_h[plan.latter_src_h]), np.arange(N))   # 'latter' is a bijection


def test_fullsize_sgcls_object_branch_properties(cuda_lib):
    from b200vsgg import synthetic, tempura
    kw = dict(KW, mode="sgcls", tracking=True)
    m = tempura.TEMPURA(obj_classes=synthetic.ag_object_classes(), **kw)
    synthetic.seeded_init_(m, 12)
    m = m.cuda().train()
    entries = _entries(with_sgcls=True)
    batch = tempura.collate_entries(entries)
    O = batch["labels"].shape[0]
    rows = torch.cat([ix.long() for ix in batch["indices"] if len(ix) > 0])
    assert rows.numel() == O and torch.equal(rows.sort().values, torch.arange(O, device=rows.device))
    with torch.no_grad():
        out = m(batch, phase="train")
    d = out["distribution"]
    assert d.shape == (O, 37) and torch.isfinite(d).all() and (d >= 0).all()
    assert (d.sum(1) - 1).abs().max().item() <= 1e-3
    assert torch.isfinite(out["object_features"]).all()
    assert (out["attention_distribution"].sum(1) - 1).abs().max().item() <= 1e-3
