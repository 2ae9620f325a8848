"""Host-side helpers that need no GPU: label-entry expansion of the class-memory bank, parameter ordering of the
regulariser's differentiable path, evaluator backend argument."""
import numpy as np
import pytest
import torch


def test_class_memory_entries_dedup_and_occurrences():
    """Weights are ASSIGNED per (pair, class) (a class listed twice counts once) while the reference's per-class normaliser
    lists get one append per occurrence (tools/utils/Uncertainty.py:170-173)."""
    from b200vsgg.memory_bank import ClassMemoryBank
    labels = [[2, 2, 5], [0], [], [4, 1]]
    rows, cls = ClassMemoryBank._entries(labels, "cpu", dedup=True)
    assert rows.tolist() == [0, 0, 1, 3, 3] and cls.tolist() == [2, 5, 0, 1, 4]
    rows, cls = ClassMemoryBank._entries(labels, "cpu", dedup=False)
    assert rows.tolist() == [0, 0, 0, 1, 3, 3] and cls.tolist() == [2, 2, 5, 0, 4, 1]
    with pytest.raises(ValueError):
        ClassMemoryBank({"attention": 3, "spatial": 6, "contacting": 17}, rel_weight_type="bogus", device="cpu")


def test_regulariser_parameter_order_is_consistent():
    """`_layer_params` (differentiable path) and `pack_small_params` (single-launch structure kernel) walk the same 18
    tensors per layer in the same order; the packed size equals the kernel's layout size."""
    from b200vsgg import regulariser
    gt = regulariser.GraphTransformer(dim=10, depth=3)
    ps = regulariser._layer_params(gt)
    assert len(ps) == 3 * regulariser._PER_LAYER
    packed = regulariser.pack_small_params(gt)
    assert packed.numel() == sum(p.numel() for p in ps)
    flat = torch.cat([p.detach().reshape(-1) for p in ps])
    assert torch.equal(flat, packed)
    names = {id(p): n for n, p in gt.named_parameters()}
    assert len({names[id(p)] for p in ps}) == len(ps) == len(list(gt.parameters()))


def test_evaluator_backend_argument():
    from b200vsgg.evaluator import BasicSceneGraphEvaluator
    kw = dict(mode="predcls", AG_object_classes=["a"], AG_all_predicates=["x", "y"], AG_attention_predicates=["x"],
              AG_spatial_predicates=["y"], AG_contacting_predicates=[])
    assert BasicSceneGraphEvaluator(**kw).backend == "host"
    assert BasicSceneGraphEvaluator(backend="cuda", **kw).backend == "cuda"
    with pytest.raises(ValueError):
        BasicSceneGraphEvaluator(backend="tpu", **kw)
    ev = BasicSceneGraphEvaluator(backend="cuda", **kw)
    with pytest.raises(RuntimeError):          # CPU tensors: no silent fallback to the host path
        ev.evaluate_scene_graph([[{"person_bbox": np.zeros((1, 4))}, {"bbox": np.zeros(4), "class": 1,
                                  "attention_relationship": torch.tensor([0]), "spatial_relationship": torch.tensor([0]),
                                  "contacting_relationship": torch.tensor([], dtype=torch.long)}]],
                                {"pair_idx": torch.zeros(1, 2, dtype=torch.long)})
