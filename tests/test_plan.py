"""Bit-exact check of the segment plan (b200vsgg.plan) against the index artefacts the reference
builds with Python loops (tools/utils/transformer.py:184-192 padding, :203-215 windows/position ids,
:236-242 'latter' scatter-back) — restated here loop-for-loop on CPU.  SURVEY.md A.1."""
import numpy as np
import pytest
import torch

from b200vsgg.plan import SegmentPlan, plan_from_im_idx


def reference_artifacts(im_idx):
    """Loop restatement for ONE video; returns pair-row lists per window token etc."""
    N = im_idx.shape[0]
    rel_idx = torch.arange(N)
    b = int(im_idx[-1] + 1)
    l = int(torch.sum(im_idx == torch.mode(im_idx)[0]))
    counts = [int(torch.sum(im_idx == i)) for i in range(b)]
    masks = torch.zeros(b, l, dtype=torch.bool)
    for i in range(b):
        masks[i, counts[i]:] = 1
    idx = -torch.ones(l * 2, b - 1)
    idx_plus = -torch.ones(l * 2, b - 1, dtype=torch.long)
    posid = -torch.ones(l * 2, b - 1, dtype=torch.long)
    for j in range(b - 1):
        sel = (im_idx == j) + (im_idx == j + 1)
        n = int(torch.sum(sel))
        idx[:n, j] = im_idx[sel]
        idx_plus[:n, j] = rel_idx[sel]
        posid[:counts[j], j] = 0
        posid[counts[j]:counts[j] + counts[j + 1], j] = 1
    # flattened window tokens in (window, slot) order, real slots only
    win_src, win_pos, win_len = [], [], []
    for j in range(b - 1):
        real = idx_plus[:, j] >= 0
        win_src += idx_plus[real, j].tolist()
        win_pos += posid[real, j].tolist()
        win_len.append(int(real.sum()))
    # 'latter': output row n comes from (window, slot)
    src_of_pair = {}
    tok_base = np.concatenate([[0], np.cumsum(win_len)])
    for j in range(b - 1):
        if j == 0:
            slots = torch.nonzero(idx[:, j] == j).flatten().tolist()
            rows = torch.nonzero(im_idx == j).flatten().tolist()
            for r, s in zip(rows, slots):
                src_of_pair[r] = tok_base[j] + s
        slots = torch.nonzero(idx[:, j] == j + 1).flatten().tolist()
        rows = torch.nonzero(im_idx == j + 1).flatten().tolist()
        for r, s in zip(rows, slots):
            src_of_pair[r] = tok_base[j] + s
    latter = [src_of_pair[n] for n in range(N)]
    return dict(counts=counts, l=l, b=b, masks=masks, win_src=win_src, win_pos=win_pos, win_len=win_len, latter=latter)


@pytest.mark.parametrize("counts", [[3, 1, 4, 1, 5], [2, 2], [1, 1, 1], [8, 6, 7, 10, 9, 6, 6, 8], [32] * 6])
def test_plan_single_video_bit_exact(counts):
    im_idx = torch.repeat_interleave(torch.arange(len(counts)), torch.tensor(counts)).float()
    ref = reference_artifacts(im_idx)
    plan = plan_from_im_idx(im_idx)
    assert plan.counts_h.tolist() == ref["counts"]
    assert plan.max_frame_len == ref["l"] and plan.F == ref["b"]
    assert plan.frame_off_h.tolist() == np.concatenate([[0], np.cumsum(ref["counts"])]).tolist()
    assert plan.win_src_h.tolist() == ref["win_src"]
    assert plan.win_pos_h.tolist() == ref["win_pos"]
    assert np.diff(plan.win_off_h).tolist() == ref["win_len"]
    assert plan.latter_src_h.tolist() == ref["latter"]
    # spatial key-padding mask [b,l]: col >= count
    mask = np.arange(plan.max_frame_len)[None, :] >= plan.counts_h[:, None]
    assert np.array_equal(mask, ref["masks"].numpy())
    # backward maps are consistent inverses
    inv = plan.inv_latter2_h
    for n, t in enumerate(plan.latter_src_h):
        assert inv[t, 0] == n
    assert (inv[:, 1] == -1).all() and (inv[:, 0] >= 0).sum() == plan.N
    assert np.array_equal(plan.inv_latter_h, inv[:, 0]) and plan.inv_latter_h.dtype == np.int32
    for n in range(plan.N):
        for t in plan.pair_win2_h[n]:
            if t >= 0:
                assert plan.win_src_h[t] == n
    assert (plan.pair_win2_h >= 0).sum() == plan.M2


def test_plan_multi_video_equals_concatenation():
    vids = [[3, 1, 4], [2, 2, 5, 1], [7, 7]]
    flat = [c for v in vids for c in v]
    plan = SegmentPlan(flat, [len(v) for v in vids])
    tok_base, pair_base = 0, 0
    src, pos, lat = [], [], []
    for v in vids:
        p1 = SegmentPlan(v, [len(v)])
        src += (p1.win_src_h + pair_base).tolist()
        pos += p1.win_pos_h.tolist()
        lat += (p1.latter_src_h + tok_base).tolist()
        tok_base += p1.M2
        pair_base += p1.N
    assert plan.win_src_h.tolist() == src and plan.win_pos_h.tolist() == pos and plan.latter_src_h.tolist() == lat
    assert plan.W == sum(len(v) - 1 for v in vids)
    assert plan.pairs_per_video.tolist() == [sum(v) for v in vids]
    assert plan.video_of_pair_h.tolist() == sum([[i] * sum(v) for i, v in enumerate(vids)], [])


def test_plan_rejects_empty_frames_and_single_frame_videos():
    with pytest.raises(AssertionError):
        SegmentPlan([3, 0, 2], [3])
    with pytest.raises(AssertionError):
        SegmentPlan([3, 2, 2], [1, 2])


# ------------------------------------------------------------------------------------------------
# SGCls object branch: class sequences (tools/utils/ds_track.py:18-39) and their positions (lib/tempura.py:186-199)
# ------------------------------------------------------------------------------------------------
def _reference_sequence_layout(entry):
    """The reference's own construction (lib/tempura.py:189-199): per sequence, unique(sorted) frame counts ->
    position = rank repeated count times; pad_sequence order = indices[1:] then the singles of indices[0]."""
    import torch
    from torch.nn.utils.rnn import pad_sequence
    indices = entry["indices"]
    pos_index = []
    for index in indices[1:]:
        _, counts = torch.unique(entry["boxes"][index][:, 0].view(-1), return_counts=True, sorted=True)
        counts = counts.tolist()
        pos_index.append(torch.cat([torch.LongTensor([im] * count) for im, count in zip(range(len(counts)), counts)]))
    padded = pad_sequence(pos_index, batch_first=True) if pos_index else torch.zeros(0, 0, dtype=torch.long)
    return indices, padded


@pytest.mark.parametrize("vid,frames,ppf", [(5, 7, (2, 5)), (8, 6, (3, 4)), (31, 20, (1, 9))])
def test_object_sequence_plan_bit_exact(vid, frames, ppf, monkeypatch):
    import numpy as np
    import torch
    from b200vsgg import objbranch, ops, synthetic
    from oracle.tempura_oracle import get_sequence as oracle_get_sequence
    monkeypatch.setattr(ops, "upload", lambda arr, dev, dtype=None: torch.as_tensor(np.asarray(arr)))
    e = synthetic.add_sgcls_inputs(synthetic.make_video_entry(vid, frames, ppf), vid)
    e2 = dict(e)
    objbranch.get_sequence(e, None, None, "sgcls")
    oracle_get_sequence(e2, "sgcls")
    assert len(e["indices"]) == len(e2["indices"])
    for a, b in zip(e["indices"], e2["indices"]):
        assert torch.equal(torch.as_tensor(a).long(), torch.as_tensor(b).long())
    plan = objbranch.ObjSeqPlan(e["indices"], e["boxes"][:, 0].numpy()).to("cpu")
    indices, padded = _reference_sequence_layout(e2)
    O = e["labels"].shape[0]
    # a permutation of the boxes: sequences first (class order), then the single-box classes
    want_src = torch.cat([ix.long() for ix in indices[1:]] + ([indices[0].long()] if len(indices[0]) else []))
    assert np.array_equal(plan.seq_src_h, want_src.numpy().astype(np.int32))
    assert np.array_equal(np.sort(plan.seq_src_h), np.arange(O))
    assert np.array_equal(plan.inv_h[plan.seq_src_h], np.arange(O))
    lens = [len(ix) for ix in indices[1:]] + [1] * len(indices[0])
    assert np.array_equal(plan.seq_off_h, np.concatenate([[0], np.cumsum(lens)]).astype(np.int32))
    r = 0
    for s, ix in enumerate(indices[1:]):
        assert np.array_equal(plan.pos_h[r:r + len(ix)], padded[s, :len(ix)].numpy().astype(np.int32)), s
        r += len(ix)
    assert (plan.pos_h[r:] == 0).all()                       # single-box sequences sit at position 0 (:205-207)
    assert plan.max_len == max(lens) and plan.S == len(lens)


def test_collate_keeps_sequences_inside_videos():
    import torch
    from b200vsgg import objbranch, synthetic, tempura
    entries = []
    for i, (f, ppf) in enumerate([(4, (1, 3)), (6, (2, 5)), (3, 4)]):
        e = synthetic.add_sgcls_inputs(synthetic.make_video_entry(60 + i, f, ppf), 60 + i)
        objbranch.get_sequence(e, None, None, "sgcls")
        entries.append(e)
    batch = tempura.collate_entries(entries)
    O = batch["labels"].shape[0]
    rows = torch.cat([ix.long() for ix in batch["indices"] if len(ix) > 0])
    assert torch.equal(rows.sort().values, torch.arange(O))
    bounds = torch.tensor([0] + [e["labels"].shape[0] for e in entries]).cumsum(0)
    video_of_box = torch.bucketize(torch.arange(O), bounds[1:], right=True)
    for ix in batch["indices"][1:]:
        assert video_of_box[ix.long()].unique().numel() == 1      # a class sequence never crosses a video
    assert batch["distribution"].shape == (O, 36)


def test_gt_label_csr_equals_trainer_multi_hot(monkeypatch):
    """Ragged label lists -> CSR (what b200vsgg_rel_loss consumes) describe exactly the multi-hot matrices the
    reference trainer builds with its per-pair loop (TEMPURA_train.py:181-187)."""
    import numpy as np
    import torch
    from b200vsgg import ops, synthetic, tempura
    monkeypatch.setattr(ops, "upload", lambda arr, dev, dtype=None: torch.as_tensor(np.asarray(arr)))
    e = synthetic.make_video_entry(13, 6, (2, 6))
    att, (s_off, s_idx), (c_off, c_idx) = tempura.gt_label_csr(e, "cpu")
    ref_att, ref_spa, ref_con = synthetic.build_gt_tensors(e)
    assert torch.equal(att, ref_att)
    for (off, idx), ref in (((s_off, s_idx), ref_spa), ((c_off, c_idx), ref_con)):
        dense = torch.zeros_like(ref)
        rows = torch.repeat_interleave(torch.arange(ref.shape[0]), torch.diff(off.long()))
        dense[rows, idx.long()] = 1
        assert torch.equal(dense, ref)


def test_object_sequence_plan_vectorised_equals_per_sequence_loop():
    """ObjSeqPlan computes the positions of all class sequences at once (segment-wise sort + dense rank, one device->host
    read); here against the reference's per-sequence construction (lib/tempura.py:191-195: unique(sorted) counts of the
    sequence's frame ids, rank k repeated count_k times) on random partitions, frame-sorted and not, tensor / numpy /
    mixed-dtype index lists."""
    import torch
    from b200vsgg import objbranch
    rng = np.random.default_rng(0)
    for trial in range(120):
        O = int(rng.integers(2, 120))
        bf = np.sort(rng.integers(0, 12, O)) if trial % 3 else rng.integers(0, 12, O)
        perm = rng.permutation(O)
        k = min(O - 1, int(rng.integers(0, 10)))
        cuts = np.sort(rng.choice(np.arange(1, O), size=k, replace=False)) if k else np.array([], dtype=np.int64)
        groups = np.split(perm, cuts)
        singles = [g for g in groups if len(g) == 1]
        seqs = [np.sort(g) if trial % 2 else g for g in groups if len(g) > 1]
        first = torch.as_tensor(np.concatenate(singles)) if singles else torch.tensor([])
        pos = np.zeros(O, dtype=np.int64)
        r = 0
        for s in seqs:
            _, counts = np.unique(bf[s], return_counts=True)
            pos[r:r + len(s)] = np.repeat(np.arange(len(counts)), counts)
            r += len(s)
        src = np.concatenate(seqs + [np.asarray(first, dtype=np.int64)])
        lens = [len(s) for s in seqs] + [1] * len(first)
        for ind in ([first] + [torch.as_tensor(g) for g in seqs], [np.asarray(first)] + seqs,
                    [first.int() if len(first) else first] + [torch.as_tensor(g) for g in seqs]):
            plan = objbranch.ObjSeqPlan(ind, bf)
            assert np.array_equal(plan.seq_src_h, src.astype(np.int32)), trial
            assert np.array_equal(plan.pos_h, pos.astype(np.int32)), trial
            assert np.array_equal(plan.seq_off_h, np.concatenate([[0], np.cumsum(lens)]).astype(np.int32))
            assert np.array_equal(plan.inv_h[plan.seq_src_h], np.arange(O))
            assert plan.max_len == max(lens) and plan.S == len(lens) and plan.max_pos == int(pos.max())
