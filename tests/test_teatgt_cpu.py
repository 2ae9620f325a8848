"""TEAT-GT on CPU: the oracle against the golden vectors written by the UNMODIFIED reference
(oracle/make_golden_teatgt.py), and the product's host plan (b200vsgg.teatgt.TeatPlan) against the
reference's integer artefacts — edge_index / edge_data per clip bit-exact, eigenvectors bit-exact."""
import os
import types

import numpy as np
import pytest
import torch

from b200vsgg import synthetic

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def _entry(gold):
    e = synthetic.make_video_entry(**gold["case"])
    e.pop("union_feat"), e.pop("spatial_masks")
    return e


@pytest.fixture(scope="module")
def oracle_model():
    from oracle.teatgt_oracle import TeatgtOracle
    gold = _load("teatgt_small")
    m = TeatgtOracle(obj_classes=synthetic.ag_object_classes(), args=types.SimpleNamespace(**gold["args"]),
                     **gold["model_kw"])
    synthetic.teatgt_seeded_init_(m, gold["seed"])
    return m.eval()


@pytest.mark.parametrize("name", ["teatgt_small", "teatgt_ragged"])
def test_oracle_matches_reference_golden(oracle_model, name):
    gold = _load(name)
    with torch.no_grad():
        out = oracle_model(_entry(gold), phase="test", return_artifacts=True)
    for k in ("attention_distribution", "spatial_distribution", "contacting_distribution"):
        assert (out[k] - gold["test/" + k]).abs().max().item() <= 2e-5, k
    assert len(out["clip_artifacts"]) == len(gold["clips"])
    for art, ref in zip(out["clip_artifacts"], gold["clips"]):
        assert torch.equal(art["edge_index"], ref["edge_index"])
        assert torch.equal(art["edge_data"], ref["edge_data"])
        assert torch.equal(art["eigvec"], ref["lap_eigvec"])


@pytest.mark.parametrize("name", ["teatgt_small", "teatgt_ragged"])
def test_host_plan_reproduces_reference_graph(oracle_model, name):
    """TeatPlan.build_graph from predicate matrices (computed here with torch on the CPU exactly as the
    CUDA kernel defines them) yields the reference's edge lists, in order, and its eigenvectors."""
    from b200vsgg.teatgt import SIM_THR, TeatPlan, edge_threshold
    from oracle.teatgt_oracle import node_layout
    gold = _load(name)
    e = _entry(gold)
    e["pred_labels"] = e["labels"]
    counts = torch.bincount(e["im_idx"].long()).numpy()
    plan = TeatPlan(counts, [len(counts)], e["pair_idx"].numpy())
    lay = node_layout(e)
    assert np.array_equal(plan.feat_row_h, lay["feat_row"].numpy())
    assert np.array_equal(plan.is_person_h, lay["is_person"].numpy())
    with torch.no_grad():
        tok = oracle_model.node_tokens(e, lay)
    box = e["boxes"][lay["feat_row"]][:, 1:]
    ctr = torch.stack([(box[:, 0] + box[:, 2]) / 2, (box[:, 1] + box[:, 3]) / 2], 1)
    thr = torch.tensor(edge_threshold(e["video_size"]), dtype=torch.float32)
    F, nmax = plan.F, plan.nmax
    sp = np.zeros((F, nmax, nmax), dtype=np.uint8)
    tp = np.zeros((F, nmax, nmax), dtype=np.uint8)
    off = plan.node_off_h
    for f in range(F):
        idx = torch.arange(off[f], off[f + 1])
        c = ctr[idx]
        d = torch.sqrt((c[:, None, 0] - c[None, :, 0]) ** 2 + (c[:, None, 1] - c[None, :, 1]) ** 2)
        n = idx.numel()
        sp[f, :n, :n] = (torch.triu(torch.ones(n, n), 1).bool() & (d <= thr)).numpy()
        if plan.has_prev_h[f]:
            pidx = torch.arange(off[f - 1], off[f])
            a, b = tok[pidx], tok[idx]
            cos = (a @ b.t()) / (a.norm(dim=1)[:, None] * b.norm(dim=1)[None, :])
            tp[f, :pidx.numel(), :n] = (cos >= SIM_THR).numpy()
    plan.build_graph(sp, tp, gold["args"]["lap_node_id_k"], eig_threads=1)
    assert plan.n_clips == len(gold["clips"])
    for cl, ref in enumerate(gold["clips"]):
        ei, ed = plan.clip_edge_index(cl)
        assert torch.equal(ei, ref["edge_index"]) and torch.equal(ed, ref["edge_data"]), cl
        n = ref["node_num"]
        k = min(n, gold["args"]["lap_node_id_k"])
        got = torch.from_numpy(plan.eigvec_h[plan.clip_node_off[cl]:plan.clip_node_off[cl + 1], :k])
        assert torch.equal(got, ref["lap_eigvec"][:, :k]), cl
    # token descriptors: [graph], [null], nodes (u == v), edges
    assert plan.T == sum(2 + c["node_num"] + c["edge_index"].shape[1] for c in gold["clips"])
    assert (plan.desc_h[plan.seq_off_h[:-1], 0] == 0).all() and (plan.desc_h[plan.seq_off_h[:-1] + 1, 0] == 1).all()


def test_sgcls_oracle_matches_reference_golden():
    """TEAT-GT SGCls, phase='train' (6 layers / 16 heads + the object branch): oracle vs the unmodified reference."""
    import os
    import types
    import torch
    from b200vsgg import synthetic
    from oracle.teatgt_oracle import TeatgtOracle
    from oracle.tempura_oracle import get_sequence
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "teatgt_sgcls.pt"), weights_only=False)
    o = TeatgtOracle(obj_classes=synthetic.ag_object_classes(), args=types.SimpleNamespace(**gold["args"]),
                     with_regulariser=False, **gold["model_kw"])
    synthetic.teatgt_seeded_init_(o, gold["seed"])
    o.train()
    o.TokenGT_encoder.p = 0.0            # dropout off on both sides when the golden vectors were made
    vid = gold["case"]["video_index"]
    e = synthetic.add_sgcls_inputs(synthetic.make_video_entry(**gold["case"]), vid)
    get_sequence(e, "sgcls")
    with torch.no_grad():
        out = o(e, phase="train")
    for k in ("distribution", "attention_distribution", "spatial_distribution", "contacting_distribution"):
        assert (out[k] - gold["train/" + k]).abs().max().item() <= 2e-5, k


def test_worker_thread_graph_build_equals_inline_build_and_propagates_errors():
    """SGCls-train hands TeatPlan.build_graph to a worker thread (TEAT_GT._prepare(background=True) -> _build_graph_job,
    joined in _finish).  The job on the pool yields the same plan arrays as the inline call, and what build_graph raises
    (an edge-less clip: the reference's stale-variable fallback, lib/teatgt.py:229-234, is not reproduced) reaches the
    caller through the future."""
    from b200vsgg import teatgt
    rng = np.random.default_rng(3)
    counts = rng.integers(2, 7, size=14)
    pair_idx, base = [], 0
    for c in counts:                                    # person row first, then its objects (object_detector.py:324-341)
        pair_idx += [(base, base + 1 + i) for i in range(c)]
        base += c + 1
    pair_idx = np.asarray(pair_idx, dtype=np.int64)

    def predicates(plan, density):
        nn_ = np.diff(plan.node_off_h)
        sp = np.zeros((plan.F, plan.nmax, plan.nmax), dtype=np.uint8)
        tp = np.zeros_like(sp)
        for f in range(plan.F):
            n = nn_[f]
            sp[f, :n, :n] = np.triu(rng.random((n, n)) < density, 1)
            sp[f, 0, 1] = density > 0                   # at least one edge per frame keeps every clip connected
            if plan.has_prev_h[f]:
                tp[f, :nn_[f - 1], :n] = rng.random((nn_[f - 1], n)) < 0.3 * density
        return sp, tp

    class _Ready:                                       # stands in for the CUDA event of the predicate copy
        def synchronize(self):
            pass

    stub = types.SimpleNamespace(lap_k=8, eig_threads=2, eig_backend="host")
    inline = teatgt.TeatPlan(counts, [8, 6], pair_idx)
    sp, tp = predicates(inline, 0.7)
    inline.build_graph(sp, tp, 8, 2, "host")
    threaded = teatgt.TeatPlan(counts, [8, 6], pair_idx)
    pr = dict(plan=threaded, sp_h=torch.from_numpy(sp), tp_h=torch.from_numpy(tp), ev=_Ready())
    ms = teatgt._graph_pool().submit(teatgt.TEAT_GT._build_graph_job, stub, pr).result(timeout=60)
    assert ms >= 0.0
    for name in ("seq_off_h", "desc_h", "node_tok_h", "eigvec_h"):
        assert np.array_equal(getattr(threaded, name), getattr(inline, name)), name
    assert threaded.T == inline.T and threaded.max_T == inline.max_T
    for a, b in zip(threaded.edges, inline.edges):
        assert np.array_equal(a, b)
    empty = teatgt.TeatPlan(counts, [8, 6], pair_idx)
    pr = dict(plan=empty, sp_h=torch.from_numpy(np.zeros_like(sp)), tp_h=torch.from_numpy(np.zeros_like(tp)), ev=_Ready())
    with pytest.raises(RuntimeError, match="edge-less clip"):
        teatgt._graph_pool().submit(teatgt.TEAT_GT._build_graph_job, stub, pr).result(timeout=60)
