"""Recall@K evaluator of the reference (tools/utils/evaluation_recall.py:9-276: BasicSceneGraphEvaluator,
evaluate_from_dict, evaluate_recall, _triplet, _compute_pred_matches), same constructor, same `result_dict`
contents, same `evaluate_scene_graph(gt, pred)` call — row (f).3 of SURVEY.md §8, HOST side.

What changes is the cost per video.  The reference converts ~20 device tensors to numpy PER FRAME (each a
device->host sync), loops over GT triplets in Python calling a Cython IoU per triplet, and updates the per-
predicate counters one `+= 1` at a time; three evaluators (with / semi / no constraint) do this for every frame of
every validation video.  Here one video costs ONE device->host transfer of the prediction tensors; triplet
equality and both IoU tests are a single [gt x pred] array expression per frame and the counters are bincounts.
The ORDER of the candidate triplets is produced by the same numpy calls on arrays of the same dtypes as the
reference builds them (float32 scores concatenated with float64 zero blocks), so ties resolve identically and the
results are bit-identical (tests/test_evaluator.py, golden values from the unmodified reference).

Two backends, identical results (tests/test_evaluator.py on the host, tests/test_evaluator_gpu.py on a B200):
`backend="host"` (default) — after the single transfer the frames are numpy array expressions; `backend="cuda"` — the
candidate ranking and the ground-truth matching of ALL frames of the video run in one launch of
`b200vsgg_eval_recall` (csrc/evaluator.cu, one CTA per frame, float64 scores with the reference's dtypes) and only the
[ground-truth relation x 4] hit flags come back.  Frames whose outcome depends on numpy's ordering of exactly tied
scores (the no-constraint mode with fewer than four pairs: zero-score entries enter its top-100 list) and frames beyond
the kernel's on-chip tables take the host path.  `bbox_overlaps`
(tools/utils/fpn/box_intersections_cpu, a Cython file absent from the reference tree) is restated with the
Fast-R-CNN definition (+1 pixel convention).
"""
import numpy as np
import torch


def bbox_overlaps(boxes, query_boxes):
    """IoU matrix [n, k], inclusive pixel coordinates (area = (x2-x1+1)(y2-y1+1)), Fast R-CNN bbox.pyx."""
    b = np.asarray(boxes, dtype=np.float64)
    q = np.asarray(query_boxes, dtype=np.float64)
    iw = np.minimum(b[:, None, 2], q[None, :, 2]) - np.maximum(b[:, None, 0], q[None, :, 0]) + 1
    ih = np.minimum(b[:, None, 3], q[None, :, 3]) - np.maximum(b[:, None, 1], q[None, :, 1]) + 1
    area_b = (b[:, 2] - b[:, 0] + 1) * (b[:, 3] - b[:, 1] + 1)
    area_q = (q[:, 2] - q[:, 0] + 1) * (q[:, 3] - q[:, 1] + 1)
    inter = iw * ih
    ua = area_b[:, None] + area_q[None, :] - inter
    return np.where((iw > 0) & (ih > 0), inter / ua, 0.0)


class BasicSceneGraphEvaluator:
    def __init__(self, mode, AG_object_classes, AG_all_predicates, AG_attention_predicates, AG_spatial_predicates,
                 AG_contacting_predicates, iou_threshold=0.5, constraint=False, semithreshold=None, output_dir="output/",
                 backend="host"):
        if backend not in ("host", "cuda"):
            raise ValueError("backend must be 'host' or 'cuda'")
        self.backend = backend
        self.AG_object_classes = AG_object_classes
        self.AG_all_predicates = AG_all_predicates
        self.AG_attention_predicates = AG_attention_predicates
        self.AG_spatial_predicates = AG_spatial_predicates
        self.AG_contacting_predicates = AG_contacting_predicates
        self.result_dict, self.per_class_recall = {}, {}
        self.mode, self.constraint = mode, constraint
        self.result_dict[self.mode + "_recall"] = {10: [], 20: [], 50: [], 100: []}
        self.iou_threshold, self.semithreshold = iou_threshold, semithreshold
        self.output_dir, self.tot_all_predicates = output_dir, len(AG_all_predicates)
        self.gt_obj_list, self.pred_obj_list = [], []
        # global predicate ids of the three groups (evaluation_recall.py:106-110 does a list.index per relation)
        self._att_id = np.asarray([AG_all_predicates.index(p) for p in AG_attention_predicates], dtype=np.int64)
        self._spa_id = np.asarray([AG_all_predicates.index(p) for p in AG_spatial_predicates], dtype=np.int64)
        self._con_id = np.asarray([AG_all_predicates.index(p) for p in AG_contacting_predicates], dtype=np.int64)

    def reset_result(self):
        self.result_dict[self.mode + "_recall"] = {10: [], 20: [], 50: [], 100: []}

    def calc_mrecall(self):
        for k in self.result_dict[self.mode + "_recall"]:
            hit = np.asarray(self.result_dict[self.mode + "_recall_hit"][k], dtype=np.float64)
            cnt = np.asarray(self.result_dict[self.mode + "_recall_count"][k], dtype=np.float64)
            per = hit / (cnt + 1e-10)
            self.per_class_recall[k] = {self.AG_all_predicates[i]: float(per[i]) for i in range(self.tot_all_predicates)}
            avg = 0                                     # plain left-to-right accumulation like the reference (:40-43);
            for v in per:                               # Python >= 3.12 sum() is compensated and differs in the last ulp
                avg += float(v)
            self.result_dict.setdefault(self.mode + "_Mrecall", {})[k] = avg / self.tot_all_predicates
        return self.result_dict[self.mode + "_Mrecall"]

    def print_stats(self, log_file=None, log_writer=None, log_epoch=None, metric=None):
        print("--------- %s_%s ---------" % (metric, self.mode))
        mrec = self.calc_mrecall()
        for k, v in self.result_dict[self.mode + "_recall"].items():
            print("R@%i: %f" % (k, np.mean(v)), flush=True)
            print("mR@%i: %f" % (k, mrec[k]), flush=True)
            if log_file:
                log_file.write("R@%i: %f \n" % (k, np.mean(v)))
                log_file.write("mR@%i: %f \n" % (k, mrec[k]))
            if log_writer:
                log_writer.add_scalar("%s_R@K/%s_R@%d" % (metric, metric, k), np.mean(v), log_epoch)
                log_writer.add_scalar("%s_MR@K/%s_MR@%d" % (metric, metric, k), mrec[k], log_epoch)

    # --------------------------------------------------------------------------------------------
    def _gt_frame(self, frame_gt):
        """evaluation_recall.py:91-116: boxes, classes and the (sub, obj, predicate) relations of one frame."""
        n = len(frame_gt)
        gt_boxes = np.zeros([n, 4])
        gt_classes = np.zeros(n)
        gt_classes[0] = 1
        gt_boxes[0] = np.asarray(frame_gt[0]["person_bbox"]).reshape(-1)[:4]
        rels = []
        for m, obj in enumerate(frame_gt[1:]):
            gt_boxes[m + 1, :] = obj["bbox"]
            gt_classes[m + 1] = obj["class"]
            a = obj["attention_relationship"]
            a = int(a.reshape(-1)[0]) if hasattr(a, "reshape") else int(a)
            rels.append([0, m + 1, self._att_id[a]])
            for s in np.asarray(obj["spatial_relationship"]).reshape(-1).tolist():
                rels.append([m + 1, 0, self._spa_id[int(s)]])
            for c in np.asarray(obj["contacting_relationship"]).reshape(-1).tolist():
                rels.append([0, m + 1, self._con_id[int(c)]])
        return gt_boxes, gt_classes, np.array(rels)

    def evaluate_scene_graph(self, gt, pred):
        if self.backend == "cuda":
            return self._evaluate_scene_graph_cuda(gt, pred)
        mode = self.mode
        host = lambda t: t.detach().cpu().numpy()
        pair_idx, im_idx = host(pred["pair_idx"]), host(pred["im_idx"])
        att, spa, con = (host(pred[k]) for k in ("attention_distribution", "spatial_distribution", "contacting_distribution"))
        pred_boxes = host(pred["boxes"][:, 1:]).astype(float)
        if mode == "predcls":
            pred_classes, obj_scores = host(pred["labels"]), host(pred["scores"])
        else:
            pred_classes, obj_scores = host(pred["pred_labels"]), host(pred["pred_scores"])
        na, ns, nc = att.shape[1], spa.shape[1], con.shape[1]
        counter = 0
        for idx, frame_gt in enumerate(gt):
            gt_boxes, gt_classes, gt_rels = self._gt_frame(frame_gt)
            if self.constraint == "no" and mode != "predcls":
                self.gt_obj_list.append({"boxes": torch.tensor(gt_boxes), "labels": torch.tensor(gt_classes)})
                self.pred_obj_list.append({
                    "boxes": pred["boxes"][counter:counter + len(frame_gt), 1:].cpu().clone(),
                    "scores": pred["pred_scores"][counter:counter + len(frame_gt)].cpu().clone(),
                    "labels": pred["pred_labels"][counter:counter + len(frame_gt)].cpu().clone()})
                counter += len(frame_gt)
            sel = im_idx == idx
            p = pair_idx[sel]
            n = p.shape[0]
            rels_i = np.concatenate((p, p[:, ::-1], p), axis=0)
            # same dtypes as the reference: float32 blocks next to float64 zeros -> float64
            rel_scores = np.concatenate((
                np.concatenate((att[sel], np.zeros([n, ns]), np.zeros([n, nc])), axis=1),
                np.concatenate((np.zeros([n, na]), spa[sel], np.zeros([n, nc])), axis=1),
                np.concatenate((np.zeros([n, na]), np.zeros([n, ns]), con[sel]), axis=1)), axis=0)
            self._evaluate_frame(gt_rels, gt_boxes.astype(float), gt_classes, rels_i, rel_scores, pred_boxes,
                                 pred_classes, obj_scores)

    def _evaluate_frame(self, gt_rels, gt_boxes, gt_classes, pred_rel_inds, rel_scores, pred_boxes, pred_classes, obj_scores):
        mode, method = self.mode, self.constraint
        # ---- candidate (subject, object, predicate) triplets per constraint (evaluate_from_dict :196-233)
        if method == "semi":
            thr = self.semithreshold
            att_row = rel_scores[:, 0] + rel_scores[:, 1] > 0
            other = ~att_row & ((rel_scores[:, 3] + rel_scores[:, 4] > 0) | (rel_scores[:, 9] + rel_scores[:, 10] > 0))
            rows, cols = [], []
            amax = rel_scores.argmax(1)
            above = rel_scores > thr
            for i in range(rel_scores.shape[0]):           # row order is part of the contract (ties in the later sort)
                if att_row[i]:
                    rows.append(i); cols.append(amax[i])
                elif other[i]:
                    ks = np.where(above[i])[0]
                    rows.extend([i] * len(ks)); cols.extend(ks.tolist())
            rows, cols = np.asarray(rows, dtype=np.int64), np.asarray(cols, dtype=np.int64)
            if rows.size:
                pred_rels = np.column_stack((pred_rel_inds[rows], cols))
                predicate_scores = rel_scores[rows, cols]
            else:
                pred_rels, predicate_scores = np.array([]), np.array([])
        elif method == "no":
            obj_scores_per_rel = obj_scores[pred_rel_inds].prod(1)
            overall = obj_scores_per_rel[:, None] * rel_scores
            order = np.argsort(-overall.ravel())[:100]
            r, c = np.unravel_index(order, overall.shape)
            pred_rels = np.column_stack((pred_rel_inds[r], c))
            predicate_scores = rel_scores[r, c]
        else:
            pred_rels = np.column_stack((pred_rel_inds, rel_scores.argmax(1)))
            predicate_scores = rel_scores.max(1)

        ks = list(self.result_dict[mode + "_recall"].keys())
        num_gt = gt_rels.shape[0]
        assert num_gt != 0
        if pred_rels.size == 0:
            hit_at = {k: np.zeros(num_gt, dtype=bool) for k in ks}
        else:
            # ---- evaluate_recall :280-347 and _compute_pred_matches :386-425 as array expressions
            gt_trip = np.column_stack((gt_classes[gt_rels[:, 0]], gt_rels[:, 2], gt_classes[gt_rels[:, 1]]))
            gt_tb = np.column_stack((gt_boxes[gt_rels[:, 0]], gt_boxes[gt_rels[:, 1]]))
            so = pred_classes[pred_rels[:, :2]]
            p_trip = np.column_stack((so[:, 0], pred_rels[:, 2], so[:, 1]))
            p_tb = np.column_stack((pred_boxes[pred_rels[:, 0]], pred_boxes[pred_rels[:, 1]]))
            trip_scores = np.column_stack((obj_scores[pred_rels[:, 0]], obj_scores[pred_rels[:, 1]], predicate_scores))
            order = trip_scores.prod(1).argsort()[::-1]
            p_trip, p_tb = p_trip[order], p_tb[order]
            same = (gt_trip[:, None, :] == p_trip[None, :, :]).all(2)                   # intersect_2d
            ok = same & (bbox_overlaps(gt_tb[:, :4], p_tb[:, :4]) >= self.iou_threshold) \
                      & (bbox_overlaps(gt_tb[:, 4:], p_tb[:, 4:]) >= self.iou_threshold)
            hit_at = {k: ok[:, :k].any(1) for k in ks}
        self._record(gt_rels, hit_at)

    def _record(self, gt_rels, hit_at):
        """Per-frame bookkeeping (evaluation_recall.py:236-262) from the hit flags of the frame's ground-truth relations."""
        mode = self.mode
        ks = list(self.result_dict[mode + "_recall"].keys())
        num_gt = gt_rels.shape[0]
        labels = gt_rels[:, 2].astype(np.int64)
        count = np.bincount(labels, minlength=self.tot_all_predicates)
        rd = self.result_dict
        for k in ks:
            hits = hit_at[k]
            cd = rd.setdefault(mode + "_recall_count", {})
            if hits.any():                      # the reference creates the hit tables at the first match (:247-253)
                hd = rd.setdefault(mode + "_recall_hit", {})
                add = np.bincount(labels[hits], minlength=self.tot_all_predicates)
                hd[k] = [int(a + b) for a, b in zip(hd.get(k, [0] * self.tot_all_predicates), add)]
            if k not in cd:
                cd[k] = [0] * self.tot_all_predicates
            cd[k] = [int(a + b) for a, b in zip(cd[k], count)]
            rd[mode + "_recall"][k].append(float(hits.sum()) / float(num_gt))

    # --------------------------------------------------------------------------------------------
    def _evaluate_scene_graph_cuda(self, gt, pred):
        """One launch + one read-back per video (see the module docstring)."""
        from . import ops
        mode, method = self.mode, self.constraint
        dev = pred["pair_idx"].device
        if dev.type != "cuda":
            raise RuntimeError("backend='cuda' needs the prediction tensors on a CUDA device")
        F_ = len(gt)
        frames = [self._gt_frame(fg) for fg in gt]
        for _, _, rels in frames:
            assert rels.shape[0] != 0
        box_off = np.concatenate([[0], np.cumsum([b.shape[0] for b, _, _ in frames])]).astype(np.int32)
        rel_off = np.concatenate([[0], np.cumsum([r.shape[0] for _, _, r in frames])]).astype(np.int32)
        gt_boxes = np.concatenate([b for b, _, _ in frames]).astype(np.float64)
        gt_classes = np.concatenate([c for _, c, _ in frames]).astype(np.int32)
        gt_rels = np.concatenate([r for _, _, r in frames]).astype(np.int32)
        if mode == "predcls":
            classes, scores = pred["labels"], pred["scores"]
        else:
            classes, scores = pred["pred_labels"], pred["pred_scores"]
        frame_off = ops.frame_offsets(pred["im_idx"].contiguous(), F_)
        kmode = {"no": 1, "semi": 2}.get(method, 0)
        hits, status = ops.eval_recall(
            pred["pair_idx"], frame_off, F_, pred["attention_distribution"], pred["spatial_distribution"],
            pred["contacting_distribution"], pred["boxes"], classes, scores, ops.upload(gt_boxes.reshape(-1), dev),
            ops.upload(gt_classes, dev), ops.upload(box_off, dev), ops.upload(gt_rels.reshape(-1), dev),
            ops.upload(rel_off, dev), kmode, float(self.semithreshold if self.semithreshold is not None else 0.0),
            float(self.iou_threshold))
        packed = torch.cat([hits.reshape(-1), status.view(torch.uint8), frame_off.view(torch.uint8)]).cpu().numpy()   # ONE read-back
        G = int(rel_off[-1])
        hits_h = packed[:4 * G].reshape(G, 4).astype(bool)
        too_big = bool(packed[4 * G:4 * G + 4].view(np.int32)[0])
        counts = np.diff(packed[4 * G + 4:].view(np.int32))
        n_pred = pred["attention_distribution"].shape[1] + pred["spatial_distribution"].shape[1] + \
            pred["contacting_distribution"].shape[1]
        host_frames = [i for i in range(F_) if (method == "no" and counts[i] * n_pred < 100) or (too_big and counts[i] > 42)]
        host_pred = None
        if host_frames or (method == "no" and mode != "predcls"):
            host = lambda t: t.detach().cpu().numpy()
            host_pred = dict(pair_idx=host(pred["pair_idx"]), im_idx=host(pred["im_idx"]),
                             att=host(pred["attention_distribution"]), spa=host(pred["spatial_distribution"]),
                             con=host(pred["contacting_distribution"]), boxes=host(pred["boxes"][:, 1:]).astype(float),
                             classes=host(classes), scores=host(scores))
        counter = 0
        ks = list(self.result_dict[mode + "_recall"].keys())
        for idx, (gt_b, gt_c, gt_r) in enumerate(frames):
            if method == "no" and mode != "predcls":
                n_box = len(gt[idx])
                self.gt_obj_list.append({"boxes": torch.tensor(gt_b), "labels": torch.tensor(gt_c)})
                self.pred_obj_list.append({
                    "boxes": pred["boxes"][counter:counter + n_box, 1:].cpu().clone(),
                    "scores": pred["pred_scores"][counter:counter + n_box].cpu().clone(),
                    "labels": pred["pred_labels"][counter:counter + n_box].cpu().clone()})
                counter += n_box
            if idx in host_frames:
                hp = host_pred
                sel = hp["im_idx"] == idx
                p = hp["pair_idx"][sel]
                n = p.shape[0]
                na, ns, nc = hp["att"].shape[1], hp["spa"].shape[1], hp["con"].shape[1]
                rels_i = np.concatenate((p, p[:, ::-1], p), axis=0)
                rel_scores = np.concatenate((
                    np.concatenate((hp["att"][sel], np.zeros([n, ns]), np.zeros([n, nc])), axis=1),
                    np.concatenate((np.zeros([n, na]), hp["spa"][sel], np.zeros([n, nc])), axis=1),
                    np.concatenate((np.zeros([n, na]), np.zeros([n, ns]), hp["con"][sel]), axis=1)), axis=0)
                self._evaluate_frame(gt_r, gt_b.astype(float), gt_c, rels_i, rel_scores, hp["boxes"], hp["classes"], hp["scores"])
                continue
            h = hits_h[rel_off[idx]:rel_off[idx + 1]]
            self._record(gt_r, {k: h[:, j] for j, k in enumerate(ks)})
