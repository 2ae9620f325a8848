"""TEST INFRASTRUCTURE.  Generates tests/golden/teatgt_*.pt by running the UNMODIFIED reference
`lib/teatgt.py::TEAT_GT.forward(phase='test')` (/root/reference, build container only) and pins
oracle/teatgt_oracle.py against it.

    python oracle/make_golden_teatgt.py

The reference's third-party imports that are absent here (fairseq, dgl, graph_transformer_pytorch,
matplotlib, the FasterRCNN CUDA ops, GloVe vectors) are replaced by the stand-ins of
oracle/ref_shims.py before import; `lib.teatgt.device` (hard-coded cuda:0) is pointed at the CPU.
No reference source is copied or modified.  phase='test' is used because the train-only regulariser
runs entirely through the absent third-party packages (unpinned, see the oracle's header).
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("VSGG_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

ARGS = dict(num_atoms=1168, num_edges=1, num_output=26, lap_node_id=True, lap_node_id_k=50,
            lap_node_id_sign_flip=False, lap_node_id_eig_dropout=0.2, rand_node_id=False, rand_node_id_dim=50,
            orf_node_id=False, orf_node_id_dim=50, type_id=True, encoder_embed_dim=768, encoder_layers=12,
            encoder_attention_heads=32, encoder_ffn_embed_dim=768, return_attention=True)
MODEL_KW = dict(mode="predcls", attention_class_num=3, spatial_class_num=6, contact_class_num=17, tracking=False)
CASES = [("teatgt_small", 3, 7, (2, 5)), ("teatgt_ragged", 11, 12, (1, 7))]


def import_reference_teatgt():
    from oracle import ref_shims
    ref_shims.install_all()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import tools.utils.word_vectors as wv

    def seeded_vectors(names, wv_type=None, wv_dir=None, wv_dim=200):
        return torch.randn(len(names), wv_dim, generator=torch.Generator().manual_seed(len(names)))

    wv.obj_edge_vectors = seeded_vectors
    import tools.utils.object_classifier as oc
    oc.obj_edge_vectors = seeded_vectors
    import lib.teatgt as ref
    ref.obj_edge_vectors = seeded_vectors
    ref.device = torch.device("cpu")
    return ref


SGCLS_ARGS = dict(ARGS, encoder_layers=6, encoder_attention_heads=16)      # tools/utils/teatgt_config.py:11-14
SGCLS_KW = dict(MODEL_KW, mode="sgcls", tracking=True)
SGCLS_CASES = [("teatgt_sgcls", 6, 7, (2, 4))]


def main_sgcls():
    """SGCls, phase='train' (the SGCls test tail needs the reference's absent CUDA ops): dropout probabilities are
    set to 0 on both sides; the regulariser outputs (third-party arithmetic, unpinned) are not stored."""
    from b200vsgg import synthetic
    from oracle.teatgt_oracle import TeatgtOracle
    from oracle.tempura_oracle import get_sequence
    torch.backends.mha.set_fastpath_enabled(False)
    torch.Tensor.cuda = lambda self, *a, **k: self      # tools/utils/object_classifier.py hard-codes masks.cuda()
    ref_mod = import_reference_teatgt()
    from tools.utils.ds_track import get_sequence as ref_get_sequence
    classes = synthetic.ag_object_classes()
    args = types.SimpleNamespace(**SGCLS_ARGS)
    ref = ref_mod.TEAT_GT(obj_classes=classes, args=args, **SGCLS_KW)
    synthetic.teatgt_seeded_init_(ref, synthetic.BASE_SEED)
    orc = TeatgtOracle(obj_classes=classes, args=args, **SGCLS_KW)
    print("sgcls state_dict interchange (strict):", orc.load_state_dict(ref.state_dict(), strict=True))
    ref.train(); orc.train()
    for model in (ref, orc):
        for m in model.modules():
            if hasattr(m, "p") and isinstance(getattr(m, "p"), float):
                m.p = 0.0
            if isinstance(m, torch.nn.MultiheadAttention):
                m.dropout = 0.0
            if hasattr(m, "dropout") and isinstance(getattr(m, "dropout"), float):
                m.dropout = 0.0
    worst = 0.0
    for name, vid, frames, ppf in SGCLS_CASES:
        entry = synthetic.add_sgcls_inputs(synthetic.make_video_entry(vid, frames, ppf), vid)
        for k in ("union_feat", "spatial_masks"):
            entry.pop(k)
        e_ref = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in entry.items()}
        e_orc = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in entry.items()}
        ref_get_sequence(e_ref, None, None, "sgcls")
        get_sequence(e_orc, "sgcls")
        with torch.no_grad():
            r = ref(e_ref, phase="train")
            o = orc(e_orc, phase="train")
        gold = {"case": dict(video_index=vid, num_frames=frames, pairs_per_frame=ppf), "args": SGCLS_ARGS,
                "model_kw": SGCLS_KW, "seed": synthetic.BASE_SEED}
        for k in ("distribution", "attention_distribution", "spatial_distribution", "contacting_distribution"):
            d = (r[k] - o[k]).abs().max().item()
            worst = max(worst, d)
            print("   %-26s %s |oracle-ref| %.2e" % (k, tuple(r[k].shape), d))
            gold["train/" + k] = r[k].clone()
        path = os.path.join(GOLDEN_DIR, name + ".pt")
        torch.save(gold, path)
        print(name, "->", path, "%.1f kB" % (os.path.getsize(path) / 1e3))
    print("sgcls max |oracle - reference|: %.3e" % worst)
    assert worst <= 2e-5, worst


def main():
    from b200vsgg import synthetic
    from oracle.teatgt_oracle import TeatgtOracle
    ref_mod = import_reference_teatgt()
    classes = synthetic.ag_object_classes()
    args = types.SimpleNamespace(**ARGS)
    ref = ref_mod.TEAT_GT(obj_classes=classes, args=args, **MODEL_KW)
    synthetic.teatgt_seeded_init_(ref, synthetic.BASE_SEED)
    ref.eval()
    orc = TeatgtOracle(obj_classes=classes, args=args, **MODEL_KW)
    print("state_dict interchange (strict):", orc.load_state_dict(ref.state_dict(), strict=True))
    orc.eval()
    captured = []
    ref.TokenGT_model.register_forward_pre_hook(lambda mod, inp: captured.append(inp[0]))
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    worst = 0.0
    for name, vid, frames, ppf in CASES:
        entry = synthetic.make_video_entry(vid, frames, ppf)
        for k in ("union_feat", "spatial_masks"):        # not on the TEAT-GT path; keep fixtures/inputs light
            entry.pop(k)
        captured.clear()
        with torch.no_grad():
            r = ref({k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in entry.items()}, phase="test")
            o = orc({k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in entry.items()}, phase="test",
                    return_artifacts=True)
        gold = {"case": dict(video_index=vid, num_frames=frames, pairs_per_frame=ppf), "args": ARGS,
                "model_kw": MODEL_KW, "seed": synthetic.BASE_SEED, "clips": []}
        for k in ("attention_distribution", "spatial_distribution", "contacting_distribution"):
            worst = max(worst, (r[k] - o[k]).abs().max().item())
            gold["test/" + k] = r[k].clone()
        assert len(captured) == len(o["clip_artifacts"])
        for bd, art in zip(captured, o["clip_artifacts"]):
            assert torch.equal(bd["edge_index"], art["edge_index"]), "edge_index differs from the reference"
            assert torch.equal(bd["edge_data"].flatten(), art["edge_data"]), "edge_data differs"
            assert torch.equal(bd["lap_eigvec"], art["eigvec"]), "Laplacian eigenvectors differ"
            gold["clips"].append({"edge_index": bd["edge_index"].clone(), "edge_data": bd["edge_data"].flatten().clone(),
                                  "node_num": int(bd["node_num"][0]), "lap_eigvec": bd["lap_eigvec"].clone()})
        path = os.path.join(GOLDEN_DIR, name + ".pt")
        torch.save(gold, path)
        print(name, "pairs=%d clips=%d edges=%s ->" % (r["attention_distribution"].shape[0], len(captured),
              [c["edge_index"].shape[1] for c in gold["clips"]]), path, "%.1f kB" % (os.path.getsize(path) / 1e3))
    print("max |oracle - reference| over all cases/outputs: %.3e" % worst)
    assert worst <= 2e-5, worst


if __name__ == "__main__":
    if "--sgcls" in sys.argv:
        main_sgcls()
    else:
        main()
