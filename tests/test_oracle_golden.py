"""The CPU oracle (oracle/tempura_oracle.py) against the golden vectors produced by the UNMODIFIED
reference (oracle/make_golden.py, run in the build container).  Runs anywhere, no GPU, no
/root/reference.  Tolerance: fp32 CPU vs fp32 CPU, different op order only -> 2e-5 max-abs."""
import os

import pytest
import torch

from b200vsgg import synthetic

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL = 2e-5


def _load(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def _clone(e):
    return {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in e.items()}


@pytest.fixture(scope="module")
def oracle_model():
    from oracle.tempura_oracle import TempuraOracle
    gold = _load("tempura_small")
    m = TempuraOracle(obj_classes=synthetic.ag_object_classes(), dropout=0.0, **gold["model_kw"])
    synthetic.seeded_init_(m, gold["seed"])
    return m.eval()


@pytest.mark.parametrize("name", ["tempura_small", "tempura_ragged"])
def test_oracle_matches_reference_golden(oracle_model, name):
    gold = _load(name)
    entry = synthetic.make_video_entry(**gold["case"])
    chk = float(entry["features"].double().sum() + entry["union_feat"].double().sum()
                + entry["spatial_masks"].double().sum())
    assert abs(chk - gold["input_checksum"]) < 1e-6 * abs(gold["input_checksum"]), "synthetic generator drifted"
    m = oracle_model
    state = {k: v.clone() for k, v in m.state_dict().items()}
    for tag in ("nomem", "mem"):
        m.rel_memory = gold["rel_memory"] if tag == "mem" else []
        with torch.no_grad():
            out = m(_clone(entry), phase="test")
            unc = m(_clone(entry), phase="test", unc=True)
            m.train()
            torch.manual_seed(99)
            tr = m(_clone(entry), phase="train")
            m.load_state_dict(state)
            te = m(_clone(entry), phase="train", eps=gold["eps"])
            m.load_state_dict(state)
            m.eval()
        n = 0
        for key, ref in gold.items():
            if not key.startswith(tag + "/"):
                continue
            _, phase, k = key.split("/")
            got = {"test": out, "unc": unc, "train_seed99": tr, "train_eps": te}[phase][k]
            err = (got - ref).abs().max().item()
            assert err <= TOL, (key, err)
            n += 1
        assert n >= 15


def test_losses_match_trainer_formulas(oracle_model):
    from oracle.tempura_oracle import tempura_losses
    entry = synthetic.make_video_entry(3, 6, (3, 5))
    with torch.no_grad():
        out = oracle_model(_clone(entry), phase="test")
    att, spa, con = synthetic.build_gt_tensors(entry)
    losses = tempura_losses(out, att, spa, con)
    # TEMPURA_train.py:200-205 spelled out with the nn modules the trainer instantiates
    ce = torch.nn.CrossEntropyLoss(reduction="none")(out["attention_distribution"], att).mean()
    b1 = torch.nn.BCELoss(reduction="none")(out["spatial_distribution"], spa).mean()
    b2 = torch.nn.BCELoss(reduction="none")(out["contacting_distribution"], con).mean()
    assert torch.allclose(losses["attention_relation_loss"], ce)
    assert torch.allclose(losses["spatial_relation_loss"], b1)
    assert torch.allclose(losses["contacting_relation_loss"], b2)


# ------------------------------------------------------------------------------------------------
# SGCls-train object branch (S1): oracle vs golden vectors of the unmodified reference
# ------------------------------------------------------------------------------------------------
SGCLS_CASES = ["sgcls_track_gmm", "sgcls_track_linear", "sgcls_notrack_gmm"]


def sgcls_setup(name):
    """(gold, entry incl. `distribution` + `indices`, seeded oracle in train mode with dropout 0)."""
    from oracle.tempura_oracle import TempuraOracle, get_sequence
    gold = _load(name)
    vid = gold["case"]["video_index"]
    entry = synthetic.add_sgcls_inputs(synthetic.make_video_entry(**gold["case"]), vid)
    chk = float(entry["features"].double().sum() + entry["distribution"].double().sum())
    assert abs(chk - gold["input_checksum"]) < 1e-6 * abs(gold["input_checksum"]), "synthetic generator drifted"
    get_sequence(entry, "sgcls")
    m = TempuraOracle(obj_classes=synthetic.ag_object_classes(), dropout=0.0, **gold["model_kw"])
    synthetic.seeded_init_(m)
    return gold, entry, m.train()


@pytest.mark.parametrize("name", SGCLS_CASES)
def test_sgcls_oracle_matches_reference_golden(name):
    gold, entry, m = sgcls_setup(name)
    # class sequences: bit-exact vs the reference's own tools/utils/ds_track.py::get_sequence
    assert len(entry["indices"]) == len(gold["indices"])
    for a, b in zip(entry["indices"], gold["indices"]):
        assert torch.equal(torch.as_tensor(a).long(), b)
    state = {k: v.clone() for k, v in m.state_dict().items()}
    with torch.no_grad():
        torch.manual_seed(99)
        tr = m(_clone(entry), phase="train")
        after = {k: v.clone() for k, v in m.state_dict().items()}
        m.load_state_dict(state)
        te = m(_clone(entry), phase="train", eps=gold["eps"])
    n = 0
    for key, ref in gold.items():
        if key.startswith("train_seed99/") or key.startswith("train_eps/"):
            got = (tr if key.startswith("train_seed99/") else te)[key.split("/")[1]]
            assert (got - ref).abs().max().item() <= TOL, key
            n += 1
        elif key.startswith("bn_after/"):
            assert (after[key.split("/", 1)[1]] - ref).abs().max().item() <= TOL, key
            n += 1
    assert n >= 12


def test_object_loss_matches_trainer_formula():
    from oracle.tempura_oracle import object_loss
    gold, entry, m = sgcls_setup("sgcls_track_linear")
    with torch.no_grad():
        out = m(_clone(entry), phase="train")
    w = torch.ones(37)
    w[0] = 0.3
    ref = torch.nn.CrossEntropyLoss(weight=w, reduction="none")(out["distribution"], out["labels"]).mean()
    assert torch.allclose(object_loss(out, eos_coef=0.3), ref)
