"""TEST INFRASTRUCTURE.  Generates tests/golden/sgcls_*.pt by running the UNMODIFIED reference TEMPURA in
SGCls mode, phase='train' (object branch S1 of SURVEY.md §8a: lib/tempura.py:185-255 with the class-sequence
encoder, tools/utils/ds_track.py:18-39) and pins oracle/tempura_oracle.py::ObjectClassifierOracle against it.

    python oracle/make_golden_sgcls.py

Same import mechanism as oracle/make_golden.py.  `center_size` (tools/utils/fpn/box_utils.py) is absent from
the reference tree and injected from the oracle (3 lines, unpinned).  The SGCls *test* tail (relabel, NMS,
ROIAlign of new union boxes, lib/tempura.py:257-307) needs the reference's absent CUDA ops and is not run.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from make_golden import GOLDEN_DIR, MODEL_KW, clone_entry, import_reference_tempura  # noqa: E402

# (name, video index, frames, pairs/frame, overrides)
CASES = [
    ("sgcls_track_gmm", 5, 7, (2, 5), dict(tracking=True, obj_head="gmm")),
    ("sgcls_track_linear", 8, 6, (3, 4), dict(tracking=True, obj_head="linear")),
    ("sgcls_notrack_gmm", 9, 5, (2, 4), dict(tracking=False, obj_head="gmm")),
]


def zero_dropout(*models):
    for model in models:
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
            if hasattr(m, "p") and isinstance(getattr(m, "p"), float):
                m.p = 0.0
            if isinstance(m, torch.nn.MultiheadAttention):
                m.dropout = 0.0


def main():
    from b200vsgg import synthetic
    from oracle.tempura_oracle import TempuraOracle, get_sequence

    torch.backends.mha.set_fastpath_enabled(False)
    # lib/tempura.py:201 hard-codes `masks.cuda()`; there is no GPU in the build container, so Tensor.cuda is
    # made the identity for this process (the reference source itself stays untouched)
    torch.Tensor.cuda = lambda self, *a, **k: self
    ref_mod = import_reference_tempura()
    from tools.utils.ds_track import get_sequence as ref_get_sequence  # the reference's own, unmodified
    classes = synthetic.ag_object_classes()
    worst = 0.0
    for name, vid, frames, ppf, over in CASES:
        kw = dict(MODEL_KW, mode="sgcls", **over)
        ref = ref_mod.TEMPURA(obj_classes=classes, **kw)
        synthetic.seeded_init_(ref)
        orc = TempuraOracle(obj_classes=classes, **kw)
        print(name, "state_dict interchange (strict):", orc.load_state_dict(ref.state_dict(), strict=True))
        entry = synthetic.add_sgcls_inputs(synthetic.make_video_entry(vid, frames, ppf), vid)
        e_ref, e_orc = clone_entry(entry), clone_entry(entry)
        ref_get_sequence(e_ref, None, None, "sgcls")
        get_sequence(e_orc, "sgcls")
        assert len(e_ref["indices"]) == len(e_orc["indices"])
        for a, b in zip(e_ref["indices"], e_orc["indices"]):
            assert torch.equal(torch.as_tensor(a).long(), torch.as_tensor(b).long())
        O, N = entry["labels"].shape[0], entry["pair_idx"].shape[0]
        gold = {"case": dict(video_index=vid, num_frames=frames, pairs_per_frame=ppf), "model_kw": kw,
                "indices": [torch.as_tensor(ix).long() for ix in e_ref["indices"]],
                "input_checksum": float(entry["features"].double().sum() + entry["distribution"].double().sum())}
        ref.train(); orc.train()
        zero_dropout(ref, orc)
        bn_state = {k: v.clone() for k, v in ref.state_dict().items()}
        keys = ["distribution", "object_features", "attention_distribution", "spatial_distribution",
                "contacting_distribution"]
        with torch.no_grad():
            torch.manual_seed(99)
            r = ref(clone_entry(e_ref), phase="train")
            after_ref = {k: v.clone() for k, v in ref.state_dict().items()}
            ref.load_state_dict(bn_state)
            torch.manual_seed(99)
            o = orc(clone_entry(e_orc), phase="train")
            after_orc = {k: v.clone() for k, v in orc.state_dict().items()}
            orc.load_state_dict(bn_state)
        for k in keys:
            d = (r[k] - o[k]).abs().max().item()
            worst = max(worst, d)
            print("   %-26s %s  |oracle-ref| %.2e" % (k, tuple(r[k].shape), d))
            if k != "object_features" or over["tracking"]:
                gold["train_seed99/" + k] = r[k].clone()
        for k in after_ref:      # BatchNorm running statistics after one train step
            if "running" in k and "object_classifier" in k:
                d = (after_ref[k] - after_orc[k]).abs().max().item()
                worst = max(worst, d)
                gold["bn_after/" + k] = after_ref[k].clone()
        # eps-injected train outputs + gradients come from the oracle AFTER it has been pinned above
        ge = torch.Generator().manual_seed(77 + vid)
        K = kw["K"]
        eps = {"object": torch.randn(K, O, len(classes), generator=ge), "attention": torch.randn(K, N, 3, generator=ge),
               "spatial": torch.randn(K, N, 6, generator=ge), "contacting": torch.randn(K, N, 17, generator=ge)}
        gold["eps"] = eps
        with torch.no_grad():
            oe = orc(clone_entry(e_orc), phase="train", eps=eps)
            orc.load_state_dict(bn_state)
        for k in keys:
            if k != "object_features":
                gold["train_eps/" + k] = oe[k].clone()
        path = os.path.join(GOLDEN_DIR, name + ".pt")
        torch.save(gold, path)
        print("  ->", path, "%.1f kB" % (os.path.getsize(path) / 1e3), "O=%d N=%d sequences=%d singles=%d" % (
            O, N, len(e_ref["indices"]) - 1, len(e_ref["indices"][0])))
    print("max |oracle - reference| over all cases/outputs: %.3e" % worst)
    assert worst <= 2e-5, worst


if __name__ == "__main__":
    main()
