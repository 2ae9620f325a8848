"""b200vsgg.memory_bank.ClassMemoryBank (device-resident streaming class memories, b200vsgg_class_memory_accumulate) against
golden memories built by the UNMODIFIED reference (`memory_computation`, `normalize_batch_uncertainty`,
`uncertainty_values.stats2`; oracle/make_golden_memory.py) on the same seeded videos, for every weight type of the
trainer's `--rel_mem_weight_type`.  Tolerance 2e-5 of the largest entry: the reference sums `exp(u) / Z` products per
video in fp32 matmuls, the bank sums `exp(u) * f` in fp32 and divides by the float64 normaliser once."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "class_memory.pt")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("wt", ["both", "al", "ep", "simple", None])
def test_class_memory_matches_reference(cuda_lib, wt):
    from make_golden_memory import DIM, REL_CLASSES, VIDEOS, synthetic_video
    from b200vsgg import ops
    from b200vsgg.memory_bank import ClassMemoryBank
    gold = torch.load(GOLDEN, weights_only=False)[str(wt)]
    bank = ClassMemoryBank(REL_CLASSES, rel_feature_dim=DIM, rel_weight_type=wt)
    n0 = ops.launch_count
    for index, n in VIDEOS:
        pred = synthetic_video(index, n)
        bank.update({k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in pred.items()})
    assert ops.launch_count - n0 >= 3 * len(VIDEOS)            # the accumulation kernel ran for every predicate group
    mem = bank.finalize()
    for rel, ref in gold.items():
        got = mem[rel].cpu()
        assert got.shape == ref.shape
        err = (got - ref).abs().max().item()
        assert err <= 2e-5 * ref.abs().max().item() + 1e-7, (wt, rel, err)


def test_class_memory_feeds_the_model(cuda_lib):
    """The finalised dict is what TEMPURA.rel_memory expects (TEMPURA_train.py:379): the memory hallucinator runs with it."""
    from make_golden_memory import DIM, REL_CLASSES, VIDEOS, synthetic_video
    from b200vsgg.memory_bank import ClassMemoryBank
    bank = ClassMemoryBank(REL_CLASSES, rel_feature_dim=DIM)
    pred = synthetic_video(*VIDEOS[0])
    bank.update({k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in pred.items()})
    mem = bank.finalize()
    assert list(mem) == ["attention", "spatial", "contacting"]
    assert sum(v.shape[0] for v in mem.values()) == 26 and all(v.shape[1] == DIM and v.is_cuda for v in mem.values())
    assert all(torch.isfinite(v).all().item() for v in mem.values())
