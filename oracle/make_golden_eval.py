"""TEST INFRASTRUCTURE.  Golden values for b200vsgg.evaluator: runs the UNMODIFIED reference evaluator
(tools/utils/evaluation_recall.py) on seeded synthetic predictions / ground truth and stores its result_dict.

    python oracle/make_golden_eval.py

Absent third-party / native imports are stubbed: h5py and dill (imported by tools/utils/pytorch_misc.py, never
used here) and the Cython `bbox_overlaps` (tools/utils/fpn/box_intersections_cpu, absent from the reference
tree: the Fast-R-CNN definition from b200vsgg.evaluator is injected — unpinned for that one function)."""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("VSGG_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden", "evaluator.pt")
CASES = [(3, 6, (3, 5)), (11, 9, (1, 7)), (21, 5, (2, 3))]


def synthetic_prediction(vid, frames, ppf, mode):
    """A seeded 'prediction': distributions correlated with the ground truth (so recall is neither 0 nor 1)."""
    from b200vsgg import synthetic
    e = synthetic.make_video_entry(vid, frames, ppf)
    for k in ("union_feat", "spatial_masks", "features"):
        e.pop(k)
    g = torch.Generator().manual_seed(1000 + vid)
    N = e["pair_idx"].shape[0]
    att_gt, spa_gt, con_gt = synthetic.build_gt_tensors(e)
    e["attention_distribution"] = torch.softmax(torch.randn(N, 3, generator=g) + 1.5 * torch.nn.functional.one_hot(att_gt, 3), 1)
    e["spatial_distribution"] = torch.sigmoid(2 * torch.randn(N, 6, generator=g) + 2.5 * spa_gt - 1)
    e["contacting_distribution"] = torch.sigmoid(2 * torch.randn(N, 17, generator=g) + 2.5 * con_gt - 1.5)
    e["scores"] = 0.5 + 0.5 * torch.rand(e["labels"].shape[0], generator=g)
    if mode != "predcls":
        wrong = torch.rand(e["labels"].shape[0], generator=g) < 0.2
        e["pred_labels"] = torch.where(wrong & (e["labels"] != 1), torch.randint(2, 37, e["labels"].shape, generator=g), e["labels"])
        e["pred_scores"] = e["scores"].clone()
    return e, synthetic.make_gt_annotation(e)


def evaluator_kwargs():
    from b200vsgg import synthetic
    return dict(AG_object_classes=synthetic.ag_object_classes(),
                AG_all_predicates=synthetic.AG_ATTENTION + synthetic.AG_SPATIAL + synthetic.AG_CONTACTING,
                AG_attention_predicates=synthetic.AG_ATTENTION, AG_spatial_predicates=synthetic.AG_SPATIAL,
                AG_contacting_predicates=synthetic.AG_CONTACTING, iou_threshold=0.5, output_dir=None)


def main():
    from b200vsgg import evaluator as mine
    for n in ("h5py", "dill", "tools.utils.fpn", "tools.utils.fpn.box_intersections_cpu"):
        sys.modules.setdefault(n, types.ModuleType(n))
    m = types.ModuleType("tools.utils.fpn.box_intersections_cpu.bbox")
    m.bbox_overlaps = mine.bbox_overlaps
    sys.modules["tools.utils.fpn.box_intersections_cpu.bbox"] = m
    sys.path.insert(0, REF)
    import tools.utils.evaluation_recall as ref
    gold = {}
    for mode in ("predcls", "sgcls"):
        for constraint, semi in (("with", None), ("semi", 0.9), ("no", None)):
            ev = ref.BasicSceneGraphEvaluator(mode=mode, constraint=constraint, semithreshold=semi, **evaluator_kwargs())
            for vid, frames, ppf in CASES:
                pred, gt = synthetic_prediction(vid, frames, ppf, mode)
                ev.evaluate_scene_graph(gt, pred)
            mr = ev.calc_mrecall()
            gold["%s/%s" % (mode, constraint)] = {"result_dict": ev.result_dict, "mrecall": mr}
            print(mode, constraint, {k: round(sum(v) / len(v), 4) for k, v in ev.result_dict[mode + "_recall"].items()},
                  {k: round(v, 4) for k, v in mr.items()})
    torch.save(gold, GOLDEN)
    print("->", GOLDEN, "%.1f kB" % (os.path.getsize(GOLDEN) / 1e3))


if __name__ == "__main__" and "--tc" not in sys.argv:
    main()


# ------------------------------------------------------------------------------------------------
# eval-time temporal-consistency score (tools/utils/temporal_consistency.py): golden from the unmodified reference
# ------------------------------------------------------------------------------------------------
TC_GOLDEN = os.path.join(ROOT, "tests", "golden", "temporal_consistency.pt")
TC_CASES = [(3, 12), (11, 30), (21, 9)]


def tc_prediction(vid, frames):
    """Track-like video: every frame holds the same object slots, ground-truth labels change slowly, so the flattened
    pair list contains runs of >= 6 equal labels for some classes (and a run that reaches the end of the list)."""
    from b200vsgg import synthetic
    e = synthetic.make_video_entry(vid, frames, 1)
    for k in ("union_feat", "spatial_masks", "features"):
        e.pop(k)
    g = torch.Generator().manual_seed(500 + vid)
    N = e["pair_idx"].shape[0]
    seg = torch.arange(N) // 8
    e["spatial_gt"] = [[int((s + vid) % 6)] for s in seg]
    e["contacting_gt"] = [[int((s * 3 + vid) % 17), int((s + 1) % 17)] for s in seg]
    e["labels"][e["pair_idx"][:, 1]] = 2 + (torch.arange(N) // 20) % 3          # a few classes, long stretches each
    e["pred_labels"] = e["labels"].clone()
    e["spatial_distribution"] = torch.sigmoid(torch.randn(N, 6, generator=g))
    e["contacting_distribution"] = torch.sigmoid(torch.randn(N, 17, generator=g))
    return e


def main_tc():
    for n in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.path.insert(0, REF)
    import tools.utils.temporal_consistency as ref
    ref.device = torch.device("cpu")            # hard-coded cuda:0 at module level (:6)
    gold = {}
    s, c = torch.tensor([]), torch.tensor([])
    for vid, frames in TC_CASES:
        pred = tc_prediction(vid, frames)
        s, c = ref.evaluate_temp_cons(pred, s, c, "predcls")
        gold["after_%d" % vid] = (s.clone(), c.clone())
        print(vid, frames, "spatial intervals", len(s), "contact intervals", len(c))
        obj_cls = pred["pred_labels"][pred["pred_labels"] != 1]
        sgt = torch.tensor([i[0] for i in pred["spatial_gt"]])
        gold["itv_%d" % vid] = {int(k): ref.find_consecutive_duplicates(obj_cls == k, sgt, pred["spatial_distribution"])
                                for k in torch.unique(obj_cls)}
    torch.save(gold, TC_GOLDEN)
    print("->", TC_GOLDEN)


if __name__ == "__main__" and "--tc" in sys.argv:
    main_tc()
