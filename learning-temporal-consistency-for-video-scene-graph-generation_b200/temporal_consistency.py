"""Eval-time temporal-consistency score of the paper (tools/utils/temporal_consistency.py:8-83 of the reference;
called per video from TEMPURA_test.py:94 / TEATGT_test.py:85) — same function names and arguments, host side.

For every object class of a video the reference scans the flattened pair list for runs of >= 6 consecutive pairs of
that class whose first spatial (contacting) ground-truth label repeats, and scores each run with
KLDiv_batchmean(log_softmax(one_hot(gt)), softmax(predicted distribution)).  The scan is a Python loop over every pair
for every class (twice) with a device transfer per run; here the runs come from two vectorised comparisons per class
and all runs of a video are scored on the host in one pass.  The reference's end-of-list quirk is kept: a run that
reaches the last pair is reported one position early (`[id - cnt, id]` with `id` the LAST index, :22-23), i.e. it
includes the pair before the run and drops the final pair.  Results are bit-identical (tests/test_evaluator.py).
"""
import numpy as np
import torch
import torch.nn.functional as F


def find_consecutive_duplicates(target_bool, gt_tensor, pred_tensor=None, window=6):
    """Intervals [start, end) as the reference returns them (tools/utils/temporal_consistency.py:8-25)."""
    b = np.asarray(target_bool, dtype=bool)
    g = np.asarray(gt_tensor)
    n = b.shape[0]
    if n == 0:
        return []
    prev = np.concatenate([[-1], g[:-1]])              # prev_state before pair i is always gt[i-1] (or -1)
    cont = b & (g == prev)
    # maximal runs of `cont`
    d = np.diff(np.concatenate([[0], cont.astype(np.int8), [0]]))
    starts, ends = np.nonzero(d == 1)[0], np.nonzero(d == -1)[0]
    out = []
    for s, e in zip(starts.tolist(), ends.tolist()):
        cnt = e - s
        if cnt < window:
            continue
        if e == n:                                      # run reaches the end of the list: reported as [id - cnt, id]
            out.append([n - 1 - cnt, n - 1])
        else:
            out.append([s, e])
    return out


def evaluate_temp_cons(pred, temp_cons_eval_spatial, temp_cons_eval_contact, mode, backend="host"):
    """backend="host": everything on the host, bit-identical to the reference.  backend="cuda": the runs are still found on
    the host (they depend on the ground-truth label lists only) but all intervals of the video are scored by ONE launch
    per predicate group (b200vsgg_interval_kl) on the device-resident distributions — no transfer of the distributions, one
    small read-back; equal to the host scores to fp32 rounding (tests/test_evaluator_gpu.py)."""
    if mode == "sgdet":
        return None, None
    spatial_gt = np.asarray([i[0] for i in pred["spatial_gt"]], dtype=np.int64)
    contact_gt = np.asarray([i[0] for i in pred["contacting_gt"]], dtype=np.int64)
    if backend == "cuda":
        from . import ops
        dev = pred["spatial_distribution"].device
        if dev.type != "cuda":
            raise RuntimeError("backend='cuda' needs the distributions on a CUDA device")
        labels = pred["pred_labels"].detach().cpu().numpy()
        obj_cls = labels[labels != 1]
        itv_s, itv_c = [], []
        for cls in np.unique(obj_cls):
            itv_s += find_consecutive_duplicates(obj_cls == cls, spatial_gt)
            itv_c += find_consecutive_duplicates(obj_cls == cls, contact_gt)
        outs = []
        for itv, gt, key in ((itv_s, spatial_gt, "spatial_distribution"), (itv_c, contact_gt, "contacting_distribution")):
            if itv:
                outs.append(ops.interval_kl(pred[key], ops.upload(gt.astype(np.int32), dev),
                                            ops.upload(np.asarray(itv, dtype=np.int32), dev)))
            else:
                outs.append(torch.zeros(0, device=dev))
        both = torch.cat(outs).to(temp_cons_eval_spatial.device)                 # one read-back
        return (torch.cat([temp_cons_eval_spatial, both[:len(itv_s)]]),
                torch.cat([temp_cons_eval_contact.to(both.device), both[len(itv_s):]]))
    spatial_pred = pred["spatial_distribution"].detach().float().cpu()
    contact_pred = pred["contacting_distribution"].detach().float().cpu()
    labels = pred["pred_labels"].detach().cpu().numpy()
    obj_cls = labels[labels != 1]

    def scores(gt, dist, n_cls, cls):
        out = []
        for s, e in find_consecutive_duplicates(obj_cls == cls, gt):
            p = F.log_softmax(F.one_hot(torch.from_numpy(gt[s:e]), n_cls).type(torch.float32), dim=1)
            q = F.softmax(dist[s:e], dim=1)
            out.append(F.kl_div(p, q, reduction="batchmean").reshape(1))
        return out

    video_spatial, video_contact = [], []
    for cls in np.unique(obj_cls):
        video_spatial += scores(spatial_gt, spatial_pred, 6, cls)
        video_contact += scores(contact_gt, contact_pred, 17, cls)
    cat = lambda prev, new: torch.cat([prev] + new) if new else torch.cat([prev, torch.tensor([])])
    return cat(temp_cons_eval_spatial, video_spatial), cat(temp_cons_eval_contact, video_contact)


def print_temp_cons_score(temp_cons_eval_spatial, temp_cons_eval_contact, mode):
    if mode != "sgdet":
        s_score, c_score = temp_cons_eval_spatial.mean() * 100, temp_cons_eval_contact.mean() * 100
        print("Spatial Temporal Consistency Score: %.6f, %d Intervals" % (s_score, len(temp_cons_eval_spatial)))
        print("Contacting Temporal Consistency Score: %.6f, %d Intervals" % (c_score, len(temp_cons_eval_contact)))
        print("Temporal Consistency Score: %.6f" % ((s_score + c_score) / 2))
