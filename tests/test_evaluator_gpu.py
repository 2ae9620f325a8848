"""backend="cuda" of b200vsgg.evaluator.BasicSceneGraphEvaluator (b200vsgg_eval_recall, one launch per video) against
(a) the golden result tables written by the UNMODIFIED reference evaluator and (b) the host backend (itself bit-identical to
the reference, tests/test_evaluator.py) on larger seeded videos: per-frame recalls, per-predicate hit / count tables and
mean recalls must be IDENTICAL for PredCLS / SGCls and the three constraint modes."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "evaluator.pt")
pytestmark = pytest.mark.gpu


def _cuda(pred):
    return {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in pred.items()}


def _same_tables(got, ref, mode):
    for k in (10, 20, 50, 100):
        assert got[mode + "_recall"][k] == ref[mode + "_recall"][k], k
        assert list(got[mode + "_recall_count"][k]) == list(ref[mode + "_recall_count"][k]), k
        assert list(got.get(mode + "_recall_hit", {}).get(k, [])) == list(ref.get(mode + "_recall_hit", {}).get(k, [])), k


@pytest.mark.parametrize("mode", ["predcls", "sgcls"])
@pytest.mark.parametrize("constraint,semi", [("with", None), ("semi", 0.9), ("no", None)])
def test_cuda_evaluator_matches_reference_golden(cuda_lib, mode, constraint, semi):
    from make_golden_eval import CASES, evaluator_kwargs, synthetic_prediction
    from b200vsgg import ops
    from b200vsgg.evaluator import BasicSceneGraphEvaluator
    gold = torch.load(GOLDEN, weights_only=False)["%s/%s" % (mode, constraint)]
    ev = BasicSceneGraphEvaluator(mode=mode, constraint=constraint, semithreshold=semi, backend="cuda", **evaluator_kwargs())
    n0 = ops.launch_count
    for vid, frames, ppf in CASES:
        pred, gt = synthetic_prediction(vid, frames, ppf, mode)
        ev.evaluate_scene_graph(gt, _cuda(pred))
    assert ops.launch_count > n0                      # the kernel ran (no silent host path)
    _same_tables(ev.result_dict, gold["result_dict"], mode)
    mr = ev.calc_mrecall()
    for k, v in gold["mrecall"].items():
        assert mr[k] == v, (k, mr[k], v)


@pytest.mark.parametrize("mode", ["predcls", "sgcls"])
@pytest.mark.parametrize("constraint,semi", [("with", None), ("semi", 0.9), ("semi", 0.5), ("no", None)])
def test_cuda_evaluator_equals_host_backend_on_larger_videos(cuda_lib, mode, constraint, semi):
    """32-frame videos of the headline shape (6-10 pairs per frame), a crowded one (up to 30 pairs) and one with very few
    pairs per frame (the no-constraint frames below four pairs take the host path by design)."""
    from make_golden_eval import evaluator_kwargs, synthetic_prediction
    from b200vsgg.evaluator import BasicSceneGraphEvaluator
    host = BasicSceneGraphEvaluator(mode=mode, constraint=constraint, semithreshold=semi, **evaluator_kwargs())
    dev = BasicSceneGraphEvaluator(mode=mode, constraint=constraint, semithreshold=semi, backend="cuda", **evaluator_kwargs())
    for vid, frames, ppf in [(40, 32, (6, 10)), (41, 12, (20, 30)), (42, 9, (1, 4)), (43, 32, (6, 10))]:
        pred, gt = synthetic_prediction(vid, frames, ppf, mode)
        host.evaluate_scene_graph(gt, pred)
        dev.evaluate_scene_graph(gt, _cuda(pred))
    _same_tables(dev.result_dict, host.result_dict, mode)
    assert dev.calc_mrecall() == host.calc_mrecall()
    if constraint == "no" and mode != "predcls":
        assert len(dev.gt_obj_list) == len(host.gt_obj_list) and len(dev.pred_obj_list) == len(host.pred_obj_list)


def test_cuda_temporal_consistency_score_matches_reference_golden(cuda_lib):
    """backend="cuda" of evaluate_temp_cons (b200vsgg_interval_kl): the same intervals as the unmodified reference (their
    count is part of the golden record) and scores equal to its values to fp32 rounding (2e-6 absolute: the device sums a
    row's classes with warp shuffles, the reference with ATen's reduction order)."""
    from make_golden_eval import TC_CASES, tc_prediction
    from b200vsgg import temporal_consistency as tc
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "temporal_consistency.pt"), weights_only=False)
    s, c = torch.tensor([]), torch.tensor([])
    for vid, frames in TC_CASES:
        pred = _cuda(tc_prediction(vid, frames))
        s, c = tc.evaluate_temp_cons(pred, s, c, "predcls", backend="cuda")
        ref_s, ref_c = gold["after_%d" % vid]
        assert s.shape == ref_s.shape and c.shape == ref_c.shape, (vid, s.shape, ref_s.shape)
        assert (s.cpu() - ref_s).abs().max().item() <= 2e-6 and (c.cpu() - ref_c).abs().max().item() <= 2e-6
    assert s.numel() > 0 and c.numel() > 0
