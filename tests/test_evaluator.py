"""b200vsgg.evaluator.BasicSceneGraphEvaluator vs golden values written by the UNMODIFIED reference evaluator
(tools/utils/evaluation_recall.py; oracle/make_golden_eval.py): every per-frame recall, the per-predicate hit /
count tables and the mean recalls must be IDENTICAL (bit-exact integer tables, recalls equal as floats) for
PredCLS / SGCls and the three constraint modes (with / semi / no)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "evaluator.pt")


@pytest.mark.parametrize("mode", ["predcls", "sgcls"])
@pytest.mark.parametrize("constraint,semi", [("with", None), ("semi", 0.9), ("no", None)])
def test_evaluator_matches_reference_golden(mode, constraint, semi):
    from make_golden_eval import CASES, evaluator_kwargs, synthetic_prediction
    from b200vsgg.evaluator import BasicSceneGraphEvaluator
    gold = torch.load(GOLDEN, weights_only=False)["%s/%s" % (mode, constraint)]
    ev = BasicSceneGraphEvaluator(mode=mode, constraint=constraint, semithreshold=semi, **evaluator_kwargs())
    for vid, frames, ppf in CASES:
        pred, gt = synthetic_prediction(vid, frames, ppf, mode)
        ev.evaluate_scene_graph(gt, pred)
    ref = gold["result_dict"]
    got = ev.result_dict
    for k in (10, 20, 50, 100):
        assert got[mode + "_recall"][k] == ref[mode + "_recall"][k], k                      # per-frame recalls
        assert list(got[mode + "_recall_count"][k]) == list(ref[mode + "_recall_count"][k]), k
        assert list(got[mode + "_recall_hit"][k]) == list(ref[mode + "_recall_hit"][k]), k
    mr = ev.calc_mrecall()
    for k, v in gold["mrecall"].items():
        assert mr[k] == v, (k, mr[k], v)
    n_frames = sum(c[1] for c in CASES)
    assert all(len(v) == n_frames for v in got[mode + "_recall"].values())


def test_bbox_overlaps_definition():
    """Inclusive-pixel IoU of Fast R-CNN's bbox.pyx (the reference's absent Cython helper)."""
    import numpy as np
    from b200vsgg.evaluator import bbox_overlaps
    a = np.array([[0, 0, 9, 9], [0, 0, 4, 4], [20, 20, 30, 30]], dtype=float)
    b = np.array([[0, 0, 9, 9], [5, 5, 14, 14]], dtype=float)
    iou = bbox_overlaps(a, b)
    assert iou[0, 0] == 1.0
    assert abs(iou[0, 1] - 25.0 / (100 + 100 - 25)) < 1e-12
    assert iou[1, 1] == 0.0 and iou[2, 0] == 0.0 and iou[2, 1] == 0.0


def test_temporal_consistency_score_matches_reference_golden():
    """Eval-time temporal-consistency score (tools/utils/temporal_consistency.py): intervals and KL scores identical to
    the unmodified reference, including its end-of-list interval quirk."""
    import numpy as np
    from make_golden_eval import TC_CASES, tc_prediction
    from b200vsgg import temporal_consistency as tc
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "temporal_consistency.pt"), weights_only=False)
    s, c = torch.tensor([]), torch.tensor([])
    n_end = 0
    for vid, frames in TC_CASES:
        pred = tc_prediction(vid, frames)
        labels = pred["pred_labels"].numpy()
        obj_cls = labels[labels != 1]
        sgt = np.asarray([i[0] for i in pred["spatial_gt"]])
        for k, ref_itv in gold["itv_%d" % vid].items():
            got = tc.find_consecutive_duplicates(obj_cls == k, sgt)
            assert got == [list(map(int, iv)) for iv in ref_itv], (vid, k, got, ref_itv)
            n_end += sum(1 for iv in got if iv[1] == len(sgt) - 1)
        s, c = tc.evaluate_temp_cons(pred, s, c, "predcls")
        rs, rc = gold["after_%d" % vid]
        assert torch.equal(s, rs) and torch.equal(c, rc), vid
    assert len(s) > 0 and len(c) > 0 and n_end > 0        # the end-of-list case is exercised
    assert tc.evaluate_temp_cons(pred, s, c, "sgdet") == (None, None)
