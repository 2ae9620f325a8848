"""Build-container only (skipped wherever /root/reference is absent, e.g. on the GPU box): the CPU oracle against the
UNMODIFIED reference run live, on cases the committed fixtures do not hold — the reference's minimum video (3 frames,
one pair per frame), equal counts in every frame, a longer ragged video — forward AND parameter gradients.  The
fixtures under tests/golden pin the oracle's forward; this pins its autograd too, which is what the GPU tests use as the
gradient reference (tests/test_tempura_gpu.py::test_backward_matches_oracle).  Same import machinery as
oracle/make_golden.py (inert stubs for the modules that are absent from the reference tree; nothing copied)."""
import os

import pytest
import torch

REF = os.environ.get("VSGG_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "lib")), reason="reference tree not mounted")

CASES = [(31, 3, 1), (32, 5, 4), (33, 12, (1, 10))]        # (video index, frames, pairs per frame)


@pytest.fixture(scope="module")
def ref_and_oracle():
    from b200vsgg import synthetic
    from oracle import make_golden
    from oracle.tempura_oracle import TempuraOracle
    torch.backends.mha.set_fastpath_enabled(False)
    ref_mod = make_golden.import_reference_tempura()
    classes = synthetic.ag_object_classes()
    ref = ref_mod.TEMPURA(obj_classes=classes, **make_golden.MODEL_KW)
    synthetic.seeded_init_(ref)
    orc = TempuraOracle(obj_classes=classes, **make_golden.MODEL_KW)
    orc.load_state_dict(ref.state_dict(), strict=True)
    for m in list(ref.modules()) + list(orc.modules()):      # dropout off on both sides (make_golden.py does the same)
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if hasattr(m, "p") and isinstance(getattr(m, "p"), float):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0
    return ref, orc


@pytest.mark.parametrize("case", CASES, ids=["3f_1p", "5f_4p", "12f_1-10p"])
def test_oracle_forward_and_gradients_equal_live_reference(ref_and_oracle, case):
    from b200vsgg import synthetic
    from oracle.make_golden import clone_entry
    from oracle.tempura_oracle import tempura_losses
    ref, orc = ref_and_oracle
    vid, frames, ppf = case
    entry = synthetic.make_video_entry(vid, frames, ppf)
    att, spa, con = synthetic.build_gt_tensors(entry)
    state = {k: v.clone() for k, v in ref.state_dict().items()}
    keys = ("attention_distribution", "spatial_distribution", "contacting_distribution", "rel_features")
    # ---- eval forward
    ref.eval(), orc.eval()
    ref.rel_memory, orc.rel_memory = [], []
    with torch.no_grad():
        r, o = ref(clone_entry(entry), phase="test"), orc(clone_entry(entry), phase="test")
    for k in keys:
        assert (r[k] - o[k]).abs().max().item() <= 2e-5, k
    # ---- train forward + backward, the reference's own CPU noise stream on both sides (gmm_heads.py:57)
    ref.train(), orc.train()
    grads = []
    for model in (ref, orc):
        model.zero_grad()
        torch.manual_seed(99)
        pred = model(clone_entry(entry), phase="train")
        loss = sum(tempura_losses(pred, att, spa, con).values())
        loss.backward()
        grads.append((float(loss.detach()), {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}))
        model.load_state_dict(state)                          # undo the BatchNorm running-statistics update
    (lr, gr), (lo, go) = grads
    assert abs(lr - lo) <= 1e-5 * max(1.0, abs(lr))
    assert gr.keys() == go.keys() and len(gr) > 150
    for n, g in gr.items():
        scale = g.abs().max().item()
        assert (g - go[n]).abs().max().item() <= 2e-4 * scale + 1e-7, n
    ref.eval(), orc.eval()


# ------------------------------------------------------------------------------------------------
# TEAT-GT (PredCLS classifier path, lib/teatgt.py:98-283): phase='test' like the fixtures — the train-only regulariser
# runs through third-party packages that are absent here (unpinned) — but with autograd on, so that the oracle's
# gradients are pinned as well.  Cases: one full clip, a trailing one-frame clip, one pair per frame.
# ------------------------------------------------------------------------------------------------
TEAT_CASES = [(41, 5, (2, 4)), (42, 6, (1, 3)), (43, 11, 1)]


@pytest.fixture(scope="module")
def teat_ref_and_oracle():
    import types
    from b200vsgg import synthetic
    from oracle import make_golden_teatgt as mg
    from oracle.teatgt_oracle import TeatgtOracle
    ref_mod = mg.import_reference_teatgt()
    classes = synthetic.ag_object_classes()
    args = types.SimpleNamespace(**dict(mg.ARGS, encoder_layers=3))      # 3 of the 12 identical layers: same code, 4x faster
    ref = ref_mod.TEAT_GT(obj_classes=classes, args=args, **mg.MODEL_KW)
    synthetic.teatgt_seeded_init_(ref, synthetic.BASE_SEED)
    orc = TeatgtOracle(obj_classes=classes, args=args, **mg.MODEL_KW)
    orc.load_state_dict(ref.state_dict(), strict=True)
    return ref.eval(), orc.eval()


@pytest.mark.parametrize("case", TEAT_CASES, ids=["5f_one_clip", "6f_trailing_1-frame_clip", "11f_1p"])
def test_teatgt_oracle_forward_and_gradients_equal_live_reference(teat_ref_and_oracle, case):
    import torch.nn.functional as F
    from b200vsgg import synthetic
    ref, orc = teat_ref_and_oracle
    vid, frames, ppf = case
    entry = synthetic.make_video_entry(vid, frames, ppf)
    for k in ("union_feat", "spatial_masks"):
        entry.pop(k)
    att, spa, con = synthetic.build_gt_tensors(entry)
    outs = []
    for model in (ref, orc):
        model.zero_grad()
        pred = model({k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in entry.items()}, phase="test")
        loss = (F.cross_entropy(pred["attention_distribution"], att) + F.binary_cross_entropy(pred["spatial_distribution"], spa)
                + F.binary_cross_entropy(pred["contacting_distribution"], con))
        loss.backward()
        outs.append((pred, {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}))
    (r, gr), (o, go) = outs
    for k in ("attention_distribution", "spatial_distribution", "contacting_distribution"):
        assert (r[k] - o[k]).abs().max().item() <= 2e-5, k
    assert len(gr) > 40
    for n, g in gr.items():
        assert n in go, n
        scale = g.abs().max().item()
        assert (g - go[n]).abs().max().item() <= 5e-4 * scale + 1e-7, n


# ------------------------------------------------------------------------------------------------
# SGCls, phase='train' (object branch S1: lib/tempura.py:185-255 with the reference's own get_sequence,
# tools/utils/ds_track.py:18-39): forward, object + relation loss gradients.  lib/tempura.py:201 hard-codes
# `masks.cuda()`; Tensor.cuda is the identity for the duration of the test (as in oracle/make_golden_sgcls.py).
# ------------------------------------------------------------------------------------------------
SGCLS_CASES = [(51, 3, 1, dict(tracking=True, obj_head="gmm")), (52, 8, (1, 6), dict(tracking=True, obj_head="linear")),
               (53, 4, (2, 3), dict(tracking=False, obj_head="gmm"))]


@pytest.mark.parametrize("case", SGCLS_CASES, ids=["3f_1p_track_gmm", "8f_1-6p_track_linear", "4f_notrack_gmm"])
def test_sgcls_oracle_forward_and_gradients_equal_live_reference(monkeypatch, case):
    from b200vsgg import synthetic
    from oracle import make_golden
    from oracle.make_golden import clone_entry
    from oracle.make_golden_sgcls import zero_dropout
    from oracle.tempura_oracle import TempuraOracle, get_sequence, object_loss, tempura_losses
    torch.backends.mha.set_fastpath_enabled(False)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    ref_mod = make_golden.import_reference_tempura()
    from tools.utils.ds_track import get_sequence as ref_get_sequence          # the reference's own, unmodified
    vid, frames, ppf, over = case
    kw = dict(make_golden.MODEL_KW, mode="sgcls", **over)
    classes = synthetic.ag_object_classes()
    ref = ref_mod.TEMPURA(obj_classes=classes, **kw)
    synthetic.seeded_init_(ref)
    orc = TempuraOracle(obj_classes=classes, **kw)
    orc.load_state_dict(ref.state_dict(), strict=True)
    entry = synthetic.add_sgcls_inputs(synthetic.make_video_entry(vid, frames, ppf), vid)
    e_ref, e_orc = clone_entry(entry), clone_entry(entry)
    ref_get_sequence(e_ref, None, None, "sgcls")
    get_sequence(e_orc, "sgcls")
    assert len(e_ref["indices"]) == len(e_orc["indices"])
    for a, b in zip(e_ref["indices"], e_orc["indices"]):
        assert torch.equal(torch.as_tensor(a).long(), torch.as_tensor(b).long())
    att, spa, con = synthetic.build_gt_tensors(entry)
    ref.train(), orc.train()
    zero_dropout(ref, orc)
    outs = []
    for model, e in ((ref, e_ref), (orc, e_orc)):
        model.zero_grad()
        torch.manual_seed(99)
        pred = model(clone_entry(e), phase="train")
        loss = sum(tempura_losses(pred, att, spa, con).values()) + object_loss(pred)
        loss.backward()
        outs.append((pred, float(loss.detach()), {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}))
    (r, lr, gr), (o, lo, go) = outs
    for k in ("distribution", "attention_distribution", "spatial_distribution", "contacting_distribution"):
        assert (r[k] - o[k]).abs().max().item() <= 2e-5, k
    assert abs(lr - lo) <= 1e-5 * max(1.0, abs(lr))
    assert gr.keys() == go.keys() and len(gr) > 150
    for n, g in gr.items():
        scale = g.abs().max().item()
        assert (g - go[n]).abs().max().item() <= 5e-4 * scale + 1e-7, n


# ------------------------------------------------------------------------------------------------
# Evaluator (f).3: the product's host backend (it runs on the CPU) against the UNMODIFIED reference evaluator run live on
# seeds and shapes the fixture does not hold: one pair per frame (no-constraint candidate lists shorter than K), many pairs
# per frame, a long video.  Tables must be identical, recalls equal as floats (as in tests/test_evaluator.py).
# ------------------------------------------------------------------------------------------------
EVAL_CASES = [(61, 3, 1), (62, 10, (1, 2)), (63, 7, (6, 10)), (64, 24, (2, 5))]


@pytest.mark.parametrize("mode", ["predcls", "sgcls"])
@pytest.mark.parametrize("constraint,semi", [("with", None), ("semi", 0.9), ("no", None)])
def test_evaluator_host_backend_equals_live_reference(mode, constraint, semi):
    import sys
    import types
    from oracle.make_golden_eval import evaluator_kwargs, synthetic_prediction
    from b200vsgg import evaluator as mine
    for n in ("h5py", "dill", "tools.utils.fpn", "tools.utils.fpn.box_intersections_cpu"):
        sys.modules.setdefault(n, types.ModuleType(n))
    m = types.ModuleType("tools.utils.fpn.box_intersections_cpu.bbox")
    m.bbox_overlaps = mine.bbox_overlaps                       # absent Cython helper: Fast R-CNN definition (unpinned)
    sys.modules["tools.utils.fpn.box_intersections_cpu.bbox"] = m
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import tools.utils.evaluation_recall as ref
    ev_ref = ref.BasicSceneGraphEvaluator(mode=mode, constraint=constraint, semithreshold=semi, **evaluator_kwargs())
    ev = mine.BasicSceneGraphEvaluator(mode=mode, constraint=constraint, semithreshold=semi, **evaluator_kwargs())
    for vid, frames, ppf in EVAL_CASES:
        pred, gt = synthetic_prediction(vid, frames, ppf, mode)
        ev_ref.evaluate_scene_graph(gt, {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in pred.items()})
        ev.evaluate_scene_graph(gt, pred)
    for k in (10, 20, 50, 100):
        assert ev.result_dict[mode + "_recall"][k] == ev_ref.result_dict[mode + "_recall"][k], k
        assert list(ev.result_dict[mode + "_recall_count"][k]) == list(ev_ref.result_dict[mode + "_recall_count"][k]), k
        assert list(ev.result_dict[mode + "_recall_hit"][k]) == list(ev_ref.result_dict[mode + "_recall_hit"][k]), k
    assert ev.calc_mrecall() == ev_ref.calc_mrecall()
