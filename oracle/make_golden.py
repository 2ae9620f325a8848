"""TEST INFRASTRUCTURE.  Generates tests/golden/tempura_*.pt by running the UNMODIFIED reference
(/root/reference, build container only — it does not travel to the GPU box) and pins
oracle/tempura_oracle.py against it.

    python oracle/make_golden.py            # writes fixtures, prints the oracle-vs-reference gaps

The reference's lib/tempura.py imports modules that are absent from its own tree (FasterRCNN CUDA
ops, Cython draw_union_boxes, fpn.box_utils) and reads a GloVe file at construction; none of them is
touched by the PredCLS forward, so they are replaced by inert stubs in sys.modules *before* import
and the GloVe loader returns seeded N(0,1) vectors.  No reference source is copied or modified.
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("VSGG_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
MODEL_KW = dict(mode="predcls", attention_class_num=3, spatial_class_num=6, contact_class_num=17,
                enc_layer_num=1, dec_layer_num=3, obj_mem_compute=False, rel_mem_compute="joint",
                mem_fusion="late", selection="manual", selection_lambda=0.5, take_obj_mem_feat=False,
                obj_head="gmm", rel_head="gmm", K=6, tracking=False)

# (name, video_index, frames, pairs/frame) — small enough that outputs are a few hundred kB
CASES = [
    ("tempura_small", 3, 6, (3, 5)),
    ("tempura_ragged", 11, 9, (1, 7)),
]


def import_reference_tempura():
    """Return the reference's lib.tempura module, imported unmodified behind inert stubs."""
    if REF not in sys.path:
        sys.path.insert(0, REF)

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m

    class _InertROIAlign(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    for n in ("tools.fasterRCNN", "tools.fasterRCNN.lib", "tools.fasterRCNN.lib.model", "tools.utils.fpn",
              "tools.utils.draw_rectangles"):
        stub(n)
    stub("tools.fasterRCNN.lib.model.roi_layers", ROIAlign=_InertROIAlign, nms=None)
    from oracle.tempura_oracle import center_size   # absent from the reference tree: injected (3 lines, unpinned)
    stub("tools.utils.fpn.box_utils", center_size=center_size)
    stub("tools.utils.draw_rectangles.draw_rectangles", draw_union_boxes=None)

    import tools.utils.word_vectors as wv

    def seeded_vectors(names, wv_type=None, wv_dir=None, wv_dim=200):
        return torch.randn(len(names), wv_dim, generator=torch.Generator().manual_seed(len(names)))

    wv.obj_edge_vectors = seeded_vectors
    import lib.tempura as ref_tempura
    ref_tempura.obj_edge_vectors = seeded_vectors
    return ref_tempura


def clone_entry(e):
    return {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in e.items()}


def main():
    from b200vsgg import synthetic
    from oracle.tempura_oracle import TempuraOracle

    torch.backends.mha.set_fastpath_enabled(False)
    ref_mod = import_reference_tempura()
    classes = synthetic.ag_object_classes()
    ref = ref_mod.TEMPURA(obj_classes=classes, **MODEL_KW)
    synthetic.seeded_init_(ref)
    ref.eval()
    orc = TempuraOracle(obj_classes=classes, **MODEL_KW)
    missing = orc.load_state_dict(ref.state_dict(), strict=True)
    print("state_dict interchange (strict):", missing)
    orc.eval()

    os.makedirs(GOLDEN_DIR, exist_ok=True)
    worst = 0.0
    for name, vid, frames, ppf in CASES:
        entry = synthetic.make_video_entry(vid, frames, ppf)
        gold = {"case": dict(video_index=vid, num_frames=frames, pairs_per_frame=ppf), "model_kw": MODEL_KW,
                "seed": synthetic.BASE_SEED,
                "input_checksum": float(entry["features"].double().sum() + entry["union_feat"].double().sum()
                                        + entry["spatial_masks"].double().sum())}
        # eps injected so that train-phase outputs are reproducible without sharing RNG state
        N = entry["pair_idx"].shape[0]
        ge = torch.Generator().manual_seed(77 + vid)
        eps = {"attention": torch.randn(6, N, 3, generator=ge), "spatial": torch.randn(6, N, 6, generator=ge),
               "contacting": torch.randn(6, N, 17, generator=ge)}
        gold["eps"] = eps
        for mem in (False, True):
            if mem:
                gm = torch.Generator().manual_seed(5)
                memory = {"attention": torch.randn(3, 1936, generator=gm), "spatial": torch.randn(6, 1936, generator=gm),
                          "contacting": torch.randn(17, 1936, generator=gm)}
                gold["rel_memory"] = memory
            else:
                memory = []
            ref.rel_memory = memory
            orc.rel_memory = memory
            tag = "mem" if mem else "nomem"
            with torch.no_grad():
                # --- test phase
                r = ref(clone_entry(entry), phase="test")
                o = orc(clone_entry(entry), phase="test")
                # --- uncertainty outputs
                ru = ref(clone_entry(entry), phase="test", unc=True)
                ou = orc(clone_entry(entry), phase="test", unc=True)
                # --- train phase with the reference's own CPU RNG noise: same seed on both sides
                ref.train(); orc.train()
                for m in list(ref.modules()) + list(orc.modules()):
                    if isinstance(m, torch.nn.Dropout):
                        m.p = 0.0
                    if hasattr(m, "p") and isinstance(getattr(m, "p"), float):
                        m.p = 0.0
                    if isinstance(m, torch.nn.MultiheadAttention):
                        m.dropout = 0.0
                bn_state = {k: v.clone() for k, v in ref.state_dict().items()}
                torch.manual_seed(99)
                rt = ref(clone_entry(entry), phase="train")
                ref.load_state_dict(bn_state)   # undo BatchNorm running-stat updates
                torch.manual_seed(99)
                ot = orc(clone_entry(entry), phase="train")
                orc.load_state_dict(bn_state)
                ref.eval(); orc.eval()
            keys = ["attention_distribution", "spatial_distribution", "contacting_distribution", "rel_features",
                    "rel_mem_features"]
            ukeys = [a + b for a in ("attention", "spatial", "contacting") for b in ("_al_uc", "_ep_uc")]
            for k in keys:
                worst = max(worst, (r[k] - o[k]).abs().max().item(), (rt[k] - ot[k]).abs().max().item())
                big = k.startswith("rel_")
                # keep fixtures small: the [N,1936] feature tensors are stored for the test phase only,
                # and only once when the two keys coincide (no memory => rel_mem_features == global output)
                if not big or (k == "rel_mem_features") or mem:
                    gold["%s/test/%s" % (tag, k)] = r[k].clone()
                if not big:
                    gold["%s/train_seed99/%s" % (tag, k)] = rt[k].clone()
            for k in ukeys:
                worst = max(worst, (ru[k] - ou[k]).abs().max().item())
                gold["%s/unc/%s" % (tag, k)] = ru[k].clone()
            # eps-injected train outputs come from the oracle AFTER it has been pinned above
            with torch.no_grad():
                orc.train()
                oe = orc(clone_entry(entry), phase="train", eps=eps)
                orc.load_state_dict(bn_state)
                orc.eval()
            for k in keys[:3]:
                gold["%s/train_eps/%s" % (tag, k)] = oe[k].clone()
        path = os.path.join(GOLDEN_DIR, name + ".pt")
        torch.save(gold, path)
        print(name, "N=%d" % N, "->", path, "%.1f kB" % (os.path.getsize(path) / 1e3))
    print("max |oracle - reference| over all cases/outputs: %.3e" % worst)
    assert worst <= 2e-5, worst


if __name__ == "__main__":
    main()
