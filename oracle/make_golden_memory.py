"""TEST INFRASTRUCTURE.  Golden class memories for b200vsgg.memory_bank: runs the UNMODIFIED reference functions
`uncertainty_values` (+ `.stats2`), `get_cls_rel_uncertainty`, `normalize_batch_uncertainty` (tools/utils/Uncertainty.py)
and `memory_computation` (tools/utils/Memory.py) on seeded synthetic uncertainties / relation features.

    python oracle/make_golden_memory.py

`uncertainty_computation` itself needs the detector and the dataset object, so its per-video bookkeeping of the RELATION
part (Uncertainty.py:148-178: np.save of rel_features, unc_list_rel[index][rel][u] = get_cls_rel_uncertainty(...),
cls_rel_uc[rel][k][u].append(unc[j, k].item())) is replayed here line by line on the synthetic tensors; everything after
that — stats2, the normalisation and the epoch-end memory build from the .npy files — is the reference's own code.
`tools/utils/ds_track.py` imports cv2 (absent): stubbed, never called."""
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("VSGG_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden", "class_memory.pt")
REL_CLASSES = {"attention": 3, "spatial": 6, "contacting": 17}
VIDEOS = [(0, 57), (1, 120), (2, 33)]            # (index, pairs)
DIM = 1936
WEIGHT_TYPES = ["both", "al", "ep", "simple", None]


def synthetic_video(index, n):
    """Seeded stand-in for model(entry, unc=True): relation features, aleatoric / epistemic uncertainties, label lists."""
    g = torch.Generator().manual_seed(900 + index)
    pred = {"rel_features": torch.randn(n, DIM, generator=g)}
    for rel, c in REL_CLASSES.items():
        pred[rel + "_al_uc"] = 0.4 * torch.randn(n, c, generator=g)
        pred[rel + "_ep_uc"] = 0.3 * torch.rand(n, c, generator=g)
    pred["attention_gt"] = [[int(torch.randint(0, 3, (1,), generator=g))] for _ in range(n)]
    for rel, c, key in (("spatial", 6, "spatial_gt"), ("contacting", 17, "contacting_gt")):
        pred[key] = [sorted(set(torch.randint(0, c, (int(torch.randint(1, 3, (1,), generator=g)),), generator=g).tolist()))
                     for _ in range(n)]
    return pred


def main():
    for n in ("cv2", "h5py", "dill"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.path.insert(0, REF)
    from tools.utils import Uncertainty as U
    from tools.utils.Memory import memory_computation
    gold = {}
    for wt in WEIGHT_TYPES:
        with tempfile.TemporaryDirectory() as tmp:
            out_dir = tmp + "/"
            os.makedirs(out_dir + "rel_embeddings/")
            os.makedirs(out_dir + "obj_embeddings/")
            unc_vals = U.uncertainty_values(37, 3, 6, 17)
            for index, n in VIDEOS:
                pred = synthetic_video(index, n)
                # ---- replay of Uncertainty.py:148-178 (relation part, rel_unc=True)
                np.save(out_dir + "rel_embeddings/" + str(index) + ".npy", pred["rel_features"].numpy(), allow_pickle=True)
                rel_labels = {"attention": pred["attention_gt"], "spatial": pred["spatial_gt"],
                              "contacting": pred["contacting_gt"]}
                tmp_dict = {}
                for rel in rel_labels:
                    tmp_dict[rel] = {}
                    labels = rel_labels[rel]
                    for u in ["al", "ep"]:
                        pred_rel_unc = pred[rel + "_" + u + "_uc"].cpu()
                        batch_unc = U.get_cls_rel_uncertainty(pred_rel_unc, labels, rel)
                        tmp_dict[rel][u] = batch_unc.numpy()
                        for j, l in enumerate(labels):
                            for k in l:
                                unc_vals.cls_rel_uc[rel][k][u].append(pred_rel_unc[j, k].item())
                unc_vals.unc_list_rel[index] = tmp_dict
                unc_vals.unc_list_obj[index] = {"al": np.zeros((1, 36), dtype=np.float32)}
            rel_memory, _ = memory_computation(unc_vals, out_dir, dict(REL_CLASSES), 37, obj_feature_dim=1024,
                                               rel_feature_dim=DIM, obj_weight_type=wt, rel_weight_type=wt,
                                               obj_mem=False, obj_unc=False, include_bg_mem=False)
            gold[str(wt)] = {k: v.clone() for k, v in rel_memory.items()}
            print(wt, {k: float(v.abs().max()) for k, v in rel_memory.items()})
    torch.save(gold, GOLDEN)
    print("->", GOLDEN)


if __name__ == "__main__":
    main()
