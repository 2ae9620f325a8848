"""Recall@K evaluator: host backend vs backend="cuda" (b200vsgg_eval_recall) on 32-frame videos of the headline shape, the
three constraint modes a test script runs per video (TEMPURA_test.py:62-92).  Prints ms per video."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
from make_golden_eval import evaluator_kwargs, synthetic_prediction
from b200vsgg.evaluator import BasicSceneGraphEvaluator

vids = [synthetic_prediction(50 + i, 32, (6, 10), "predcls") for i in range(6)]
cuda = [({k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in p.items()}, g) for p, g in vids]
for backend, data in (("host", cuda), ("cuda", cuda)):
    evs = [BasicSceneGraphEvaluator(mode="predcls", constraint=c, semithreshold=s, backend=backend, **evaluator_kwargs())
           for c, s in (("with", None), ("semi", 0.9), ("no", None))]
    for ev in evs:                      # warm-up
        ev.evaluate_scene_graph(data[0][1], data[0][0])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for pred, gt in data:
        for ev in evs:
            ev.evaluate_scene_graph(gt, pred)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / len(data)
    print("backend=%-4s  %.1f ms per 32-frame video (three evaluators, predictions resident on the device)  R@20 with-constraint %.4f"
          % (backend, ms, sum(evs[0].result_dict["predcls_recall"][20]) / len(evs[0].result_dict["predcls_recall"][20])))
