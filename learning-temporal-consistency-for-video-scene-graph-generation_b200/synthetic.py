"""Synthetic Action-Genome-shaped inputs for the relation-classification path.

Produces the ``entry`` dict that the reference's frozen detector hands to the model in PredCLS mode
(tools/utils/object_detector.py:382-396 of the reference): per video F frames, per frame one
person box followed by its object boxes, one (person, object) pair per object.  Shapes and value
ranges follow SURVEY.md §8(d).  Everything is seeded (1123 + video index mirrors
tools/utils/env.py:11) so the oracle, the golden-vector generator and the CUDA path see identical
tensors without any fixture files for the inputs.
"""
import math
import zlib

import torch

AG_NUM_OBJ_CLASSES = 37  # '__background__', 'person', 35 objects (dataloader/AG/action_genome.py:24-32)
ATTENTION_CLASSES, SPATIAL_CLASSES, CONTACT_CLASSES = 3, 6, 17
VIDEO_SIZE = (480, 270)
BASE_SEED = 1123


def ag_object_classes():
    return ["__background__", "person"] + ["obj%02d" % i for i in range(35)]


def make_video_entry(video_index=0, num_frames=32, pairs_per_frame=(6, 10), device="cpu",
                     feat_dim=2048, with_gt=True, big_on_device=None):
    """One video's PredCLS entry dict (all tensors on ``device``).

    pairs_per_frame: int (fixed) or (lo, hi) inclusive range.
    big_on_device: if given, union_feat / spatial_masks are drawn directly on that device with a
    device generator (used by bench.py at sizes where CPU generation is too slow); parity tests
    leave it None so inputs are bit-identical everywhere."""
    g = torch.Generator().manual_seed(BASE_SEED + int(video_index))
    F = int(num_frames)
    assert F >= 3  # dataloader/AG/action_genome.py:152 drops shorter videos
    if isinstance(pairs_per_frame, int):
        counts = torch.full((F,), pairs_per_frame, dtype=torch.int64)
    else:
        lo, hi = pairs_per_frame
        counts = torch.randint(lo, hi + 1, (F,), generator=g)
    max_obj = int(counts.max())
    N = int(counts.sum())
    O = N + F
    W, H = VIDEO_SIZE

    # tracks: slot 0 = person, slots 1..max_obj = objects with a fixed class and a base feature
    obj_cls = torch.randint(2, AG_NUM_OBJ_CLASSES, (max_obj,), generator=g)
    base_feat = torch.randn(max_obj + 1, feat_dim, generator=g)
    base_ctr = torch.stack([torch.rand(max_obj + 1, generator=g) * W, torch.rand(max_obj + 1, generator=g) * H], 1)

    boxes = torch.zeros(O, 5)
    labels = torch.zeros(O, dtype=torch.int64)
    features = torch.empty(O, feat_dim)
    pair_idx = torch.zeros(N, 2, dtype=torch.int64)
    im_idx = torch.zeros(N)
    human_idx = torch.zeros(F, 1, dtype=torch.int64)
    row, p = 0, 0
    for f in range(F):
        n = int(counts[f])
        slots = torch.cat([torch.zeros(1, dtype=torch.int64), 1 + torch.randperm(max_obj, generator=g)[:n].sort().values])
        k = n + 1
        features[row:row + k] = base_feat[slots] + 0.3 * torch.randn(k, feat_dim, generator=g)
        ctr = base_ctr[slots] + 6.0 * torch.randn(k, 2, generator=g)
        wh = 20.0 + 60.0 * torch.rand(k, 2, generator=g)
        boxes[row:row + k, 0] = f
        boxes[row:row + k, 1:3] = ctr - wh / 2
        boxes[row:row + k, 3:5] = ctr + wh / 2
        labels[row] = 1
        labels[row + 1:row + k] = obj_cls[slots[1:] - 1]
        human_idx[f, 0] = row
        pair_idx[p:p + n, 0] = row
        pair_idx[p:p + n, 1] = torch.arange(row + 1, row + k)
        im_idx[p:p + n] = f
        row += k
        p += n

    if big_on_device is not None:
        gd = torch.Generator(device=big_on_device).manual_seed(BASE_SEED + int(video_index))
        union_feat = torch.randn(N, 1024, 7, 7, generator=gd, device=big_on_device).relu_()
        spatial_masks = torch.rand(N, 2, 27, 27, generator=gd, device=big_on_device).round_().sub_(0.5)
    else:
        union_feat = torch.randn(N, 1024, 7, 7, generator=g).relu_()
        spatial_masks = torch.rand(N, 2, 27, 27, generator=g).round_().sub_(0.5)

    sub = boxes[pair_idx[:, 0]]
    obj = boxes[pair_idx[:, 1]]
    union_box = torch.cat([im_idx[:, None], torch.minimum(sub[:, 1:3], obj[:, 1:3]),
                           torch.maximum(sub[:, 3:5], obj[:, 3:5])], 1)
    entry = {
        "boxes": boxes, "labels": labels, "scores": torch.ones(O), "im_idx": im_idx, "pair_idx": pair_idx,
        "human_idx": human_idx, "features": features, "union_feat": union_feat, "union_box": union_box,
        "spatial_masks": spatial_masks,
    }
    dev = torch.device(device)
    entry = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in entry.items()}
    entry["video_size"] = VIDEO_SIZE
    entry["video_id"] = "synthetic_%05d" % video_index
    if with_gt:
        att = torch.randint(0, ATTENTION_CLASSES, (N,), generator=g)
        entry["attention_gt"] = [[int(a)] for a in att]
        entry["spatial_gt"] = [sorted(set(torch.randint(0, SPATIAL_CLASSES, (int(torch.randint(1, 3, (1,), generator=g)),),
                                                          generator=g).tolist())) for _ in range(N)]
        entry["contacting_gt"] = [sorted(set(torch.randint(0, CONTACT_CLASSES, (int(torch.randint(1, 3, (1,), generator=g)),),
                                                             generator=g).tolist())) for _ in range(N)]
    return entry


def add_sgcls_inputs(entry, video_index=0, sharpness=3.0):
    """SGCls detector hand-off on top of a PredCLS entry (tools/utils/object_detector.py:415-430 of the
    reference): `distribution` [O,36] = the detector's posterior over the 36 non-background classes, peaked at
    the ground-truth class (logit + `sharpness`) so that the arg-max class sequences of
    tools/utils/ds_track.py:18-39 look like object tracks with a realistic share of misdetections."""
    g = torch.Generator().manual_seed(BASE_SEED * 7 + int(video_index))
    labels = entry["labels"].cpu()
    logits = torch.randn(labels.shape[0], AG_NUM_OBJ_CLASSES - 1, generator=g)
    logits[torch.arange(labels.shape[0]), labels - 1] += sharpness
    entry["distribution"] = torch.softmax(logits, 1).to(entry["labels"].device)
    return entry


AG_ATTENTION = ["looking_at", "not_looking_at", "unsure"]
AG_SPATIAL = ["above", "beneath", "in_front_of", "behind", "on_the_side_of", "in"]
AG_CONTACTING = ["carrying", "covered_by", "drinking_from", "eating", "have_it_on_the_back", "holding", "leaning_on",
                 "lying_on", "not_contacting", "other_relationship", "sitting_on", "standing_on", "touching", "twisting",
                 "wearing", "wiping", "writing_on"]


def make_gt_annotation(entry):
    """The per-frame ground-truth structure the AG dataloader hands to the evaluator (dataloader/AG/action_genome.py:
    one list per frame = [{'person_bbox', 'frame'}, {'class', 'bbox', 'attention_relationship',
    'spatial_relationship', 'contacting_relationship'}, ...]) rebuilt from a synthetic entry's labels."""
    boxes, labels = entry["boxes"].cpu(), entry["labels"].cpu()
    pair, im = entry["pair_idx"].cpu(), entry["im_idx"].cpu().long()
    frames = []
    for f in range(int(im.max()) + 1):
        rows = (im == f).nonzero().flatten().tolist()
        human = int(pair[rows[0], 0])
        fr = [{"person_bbox": boxes[human, 1:].numpy()[None].copy(), "frame": "%s/%06d.png" % (entry.get("video_id", "v"), f)}]
        for r in rows:
            o = int(pair[r, 1])
            fr.append({"class": int(labels[o]), "bbox": boxes[o, 1:].numpy().copy(),
                       "attention_relationship": torch.tensor(entry["attention_gt"][r], dtype=torch.long),
                       "spatial_relationship": torch.tensor(entry["spatial_gt"][r], dtype=torch.long),
                       "contacting_relationship": torch.tensor(entry["contacting_gt"][r], dtype=torch.long)})
        frames.append(fr)
    return frames


def seeded_init_(module, seed=BASE_SEED):
    """Deterministic, construction-order-independent parameter fill: every tensor of the state_dict
    is drawn from a generator keyed by (seed, crc32(name)).  Applied to the reference modules when
    the golden vectors are generated and to the oracle / CUDA modules in tests, so identical weights
    never have to be stored.  Scales are chosen so all branches of the path are exercised
    (non-trivial LayerNorm/BatchNorm affine terms and running statistics, biases != 0)."""
    sd = module.state_dict()
    with torch.no_grad():
        for name in sorted(sd.keys()):
            t = sd[name]
            if not torch.is_floating_point(t):
                continue
            g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63))
            leaf = name.rsplit(".", 1)[-1]
            if leaf == "running_var":
                v = 0.5 + torch.rand(t.shape, generator=g)
            elif leaf == "running_mean":
                v = 0.1 * torch.randn(t.shape, generator=g)
            elif leaf == "pe":
                continue
            elif "position_embedding" in name:
                v = torch.rand(t.shape, generator=g)
            elif "embed" in name and t.dim() == 2 and leaf == "weight" and "pos_embed" not in name:
                v = torch.randn(t.shape, generator=g)
            elif t.dim() >= 2:
                fan_in = t[0].numel()
                a = 1.0 / math.sqrt(fan_in)
                v = (torch.rand(t.shape, generator=g) * 2 - 1) * a
            elif leaf == "weight":  # 1-D weight = norm scale
                v = 1.0 + 0.1 * torch.randn(t.shape, generator=g)
            else:  # biases
                v = 0.05 * torch.randn(t.shape, generator=g)
            t.copy_(v.to(t.dtype))
    return module


def build_gt_tensors(entry, device=None):
    """Label tensors exactly as the reference trainer builds them (TEMPURA_train.py:181-187):
    attention -> class index [N]; spatial / contacting -> multi-hot float [N,6] / [N,17]."""
    N = len(entry["attention_gt"])
    dev = device if device is not None else entry["im_idx"].device
    att = torch.tensor([a[0] if isinstance(a, (list, tuple)) else int(a) for a in entry["attention_gt"]],
                       dtype=torch.int64)
    spa = torch.zeros(N, SPATIAL_CLASSES)
    con = torch.zeros(N, CONTACT_CLASSES)
    for i in range(N):
        spa[i, entry["spatial_gt"][i]] = 1
        con[i, entry["contacting_gt"][i]] = 1
    return att.to(dev), spa.to(dev), con.to(dev)


def teatgt_seeded_init_(model, seed=BASE_SEED):
    """seeded_init_ plus the zero rows that nn.Embedding(padding_idx=0) keeps in a real TokenGT model."""
    seeded_init_(model, seed)
    with torch.no_grad():
        for name, p in model.state_dict().items():
            if name.endswith("temp_encoder.weight") or name.endswith("edge_encoder.weight"):
                p[0].zero_()
            if name.endswith("embed_out.weight"):        # the generic rule treats "*embed*" as an N(0,1) table;
                p.mul_(1.0 / math.sqrt(p.shape[1]))      # this one is the output projection: keep logits O(1)
    return model
