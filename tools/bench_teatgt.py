"""Secondary benchmark (not the driver's bench.py contract): TEAT-GT PredCLS fwd+loss+bwd on a batch of
synthetic AG videos (BASELINE configs[2]-shaped graphs, PredCLS heads).  Prints one JSON line."""
import argparse, json, os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from b200vsgg import ops, synthetic, teatgt, tempura

ARGS = dict(num_atoms=1168, num_edges=1, num_output=26, lap_node_id=True, lap_node_id_k=50, lap_node_id_sign_flip=False,
            lap_node_id_eig_dropout=0.2, rand_node_id=False, rand_node_id_dim=50, orf_node_id=False, orf_node_id_dim=50,
            type_id=True, encoder_embed_dim=768, encoder_layers=12, encoder_attention_heads=32, encoder_ffn_embed_dim=768,
            return_attention=True)
ap = argparse.ArgumentParser()
ap.add_argument("--videos", type=int, default=64)
ap.add_argument("--frames", type=int, default=32)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--infer", action="store_true")
ap.add_argument("--eig", default="host", choices=["host", "device"])
ap.add_argument("--mode", default="predcls", choices=["predcls", "sgcls"],
                help="sgcls = BASELINE configs[2]: object branch + 6-layer/16-head encoder (teatgt_config.py:11-14)")
ap.add_argument("--pairs", default="6,10", help="pairs per frame: 'lo,hi' or a single number (configs[4]: 32)")
ap.add_argument("--chunks", type=int, default=None, help="TEAT_GT.pipeline_chunks (video chunks of the host/device pipeline)")
a = ap.parse_args()
ppf = tuple(int(x) for x in a.pairs.split(",")) if "," in a.pairs else int(a.pairs)
sgcls = a.mode == "sgcls"
if sgcls:
    ARGS.update(encoder_layers=6, encoder_attention_heads=16)
dev = torch.device("cuda", 0)
m = teatgt.TEAT_GT(mode=a.mode, attention_class_num=3, spatial_class_num=6, contact_class_num=17,
                   obj_classes=synthetic.ag_object_classes(), tracking=sgcls, args=types.SimpleNamespace(**ARGS))
synthetic.teatgt_seeded_init_(m, 1123)
m = m.to(dev)
m.eig_backend = a.eig
if a.chunks is not None:
    m.pipeline_chunks = a.chunks
    m.pipeline_min_pairs = 0
if not sgcls:
    for p in m.object_classifier.parameters():
        p.requires_grad_(False)
entries = []
for i in range(a.videos):
    e = synthetic.make_video_entry(i, a.frames, ppf)
    e.pop("union_feat"), e.pop("spatial_masks")
    e = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in e.items()}
    if sgcls:
        from b200vsgg import objbranch
        synthetic.add_sgcls_inputs(e, i)
        objbranch.get_sequence(e, None, None, "sgcls")
    entries.append(e)
gts = [synthetic.build_gt_tensors(e, dev) for e in entries]
batch = tempura.collate_entries(entries)
for k in ("attention_gt", "spatial_gt", "contacting_gt"):
    batch.pop(k, None)
att, spa, con = (torch.cat([g[i] for g in gts]) for i in range(3))
batch["frame_counts_host"] = torch.bincount(batch["im_idx"].long()).cpu().numpy()
batch["pair_idx_host"] = batch["pair_idx"].cpu().numpy()
batch["box_frames_host"] = batch["boxes"][:, 0].cpu().numpy()
dist_in = batch.get("distribution")
N = batch["pair_idx"].shape[0]

def step():
    if a.infer:
        with torch.no_grad():
            return m(dict(batch), phase="test")["attention_distribution"].sum()
    m.zero_grad(set_to_none=True)
    out = m(dict(batch), phase="train")
    loss = 0.0
    if sgcls:
        from b200vsgg.objbranch import object_loss
        grp = out["box_groups"]
        loss = object_loss(out, 1.0, grp.count, grp.video_of_box64)
    loss = loss + (torch.nn.functional.cross_entropy(out["attention_distribution"], att)
            + torch.nn.functional.binary_cross_entropy(out["spatial_distribution"], spa)
            + torch.nn.functional.binary_cross_entropy(out["contacting_distribution"], con))
    loss.backward()
    return loss

m.train(not a.infer)
for _ in range(a.warmup):
    step()
torch.cuda.synchronize()
l0 = ops.launch_count
t0 = time.perf_counter()
for _ in range(a.steps):
    loss = step()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / a.steps
plan = m.last_plan
print(json.dumps({"workload": "TEAT-GT %s %s, %d videos x %d frames x %s pairs/frame" % (a.mode, "inference" if a.infer else "fwd+bwd", a.videos, a.frames, a.pairs),
                  "boxes": int(batch["labels"].shape[0]),
                  "eig_backend": a.eig, "pipeline_chunks": m.pipeline_chunks, "host_graph_ms": getattr(m, "last_host_graph_ms", None),
                  "pairs_per_s": N / dt, "ms_per_step": dt * 1e3, "pairs": N, "clips": plan.n_clips, "tokens": plan.T,
                  "max_tokens_per_clip": plan.max_T, "launches_per_step": (ops.launch_count - l0) // a.steps,
                  "loss": float(loss)}))
