"""Segment plan: every integer index the relation path needs, computed once per batch from the
per-frame pair counts.  These replace the Python loops of tools/utils/transformer.py:184-192
(padding), :203-215 (2-frame windows, position ids) and :236-242 ('latter' scatter-back) of the
reference and are compared bit-exactly against them in tests/test_plan.py.

Frames of several videos are concatenated; windows never cross a video boundary.
"""
import numpy as np
import torch


class SegmentPlan:
    """All arrays are int32 numpy on the host (`*_h`) and mirrored on the device after `.to(device)`.

    counts[F]        pairs per frame                     frame_off[F+1]  exclusive scan
    win_first[W]     first frame of window w             win_off[W+1]    window-token offsets
    win_src[M2]      pair row feeding window token t     win_pos[M2]     0 = former frame, 1 = latter
    latter_src[N]    window token that yields pair n's output ('latter' mode)
    inv_latter2[M2,2] (pair n or -1, -1): backward of the latter gather      inv_latter[M2] its first column
    pair_win2[N,2]   window tokens that read pair n (former, latter), -1 if none
    video_of_pair[N] video id per pair (int64, for per-video BatchNorm statistics / losses)
    """

    def __init__(self, counts, frames_per_video):
        c = np.asarray(counts, dtype=np.int64)
        fpv = np.asarray(frames_per_video, dtype=np.int64)
        assert c.ndim == 1 and fpv.sum() == c.shape[0], "frames_per_video must sum to the number of frames"
        assert (c > 0).all(), "every frame needs at least one pair (dataloader/AG/action_genome.py:96-105)"
        assert (fpv >= 2).all(), "every video needs at least two frames"
        F = c.shape[0]
        V = fpv.shape[0]
        off = np.zeros(F + 1, dtype=np.int64)
        off[1:] = np.cumsum(c)
        N = int(off[-1])
        vf_off = np.zeros(V + 1, dtype=np.int64)
        vf_off[1:] = np.cumsum(fpv)
        video_of_frame = np.repeat(np.arange(V), fpv)
        is_first = np.zeros(F, dtype=bool)
        is_first[vf_off[:-1]] = True
        is_last = np.zeros(F, dtype=bool)
        is_last[vf_off[1:] - 1] = True

        win_first = np.nonzero(~is_last)[0]                 # window w = frames (win_first[w], +1)
        W = win_first.shape[0]
        wlen = c[win_first] + c[win_first + 1]
        win_off = np.zeros(W + 1, dtype=np.int64)
        win_off[1:] = np.cumsum(wlen)
        M2 = int(win_off[-1])
        local = np.arange(M2) - np.repeat(win_off[:-1], wlen)
        win_src = local + np.repeat(off[win_first], wlen)
        win_pos = (local >= np.repeat(c[win_first], wlen)).astype(np.int64)

        widx_first = np.full(F, -1, dtype=np.int64)         # window whose FIRST frame is f
        widx_first[win_first] = np.arange(W)
        frame_of_pair = np.repeat(np.arange(F), c)
        k_in_frame = np.arange(N) - off[frame_of_pair]
        fpair = frame_of_pair
        former_tok = np.where(~is_last[fpair], win_off[np.maximum(widx_first[fpair], 0)] + k_in_frame, -1)
        prev = np.maximum(fpair - 1, 0)
        latter_tok = np.where(~is_first[fpair], win_off[np.maximum(widx_first[prev], 0)] + c[prev] + k_in_frame, -1)
        latter_src = np.where(is_first[fpair], former_tok, latter_tok)
        inv = np.full((M2, 2), -1, dtype=np.int64)
        inv[latter_src, 0] = np.arange(N)

        self.N, self.F, self.V, self.W, self.M2 = N, F, V, W, M2
        self.max_frame_len = int(c.max())
        self.max_win_len = int(wlen.max()) if W else 0
        self.frames_per_video = fpv
        self.pairs_per_video = np.add.reduceat(c, vf_off[:-1])
        i32 = np.int32
        self.counts_h = c.astype(i32)
        self.frame_off_h = off.astype(i32)
        self.win_first_h = win_first.astype(i32)
        self.win_off_h = win_off.astype(i32)
        self.win_src_h = win_src.astype(i32)
        self.win_pos_h = win_pos.astype(i32)
        self.latter_src_h = latter_src.astype(i32)
        self.inv_latter2_h = inv.astype(i32)
        self.inv_latter_h = inv[:, 0].astype(i32)
        self.pair_win2_h = np.stack([former_tok, latter_tok], 1).astype(i32)
        self.video_of_pair_h = video_of_frame[frame_of_pair].astype(np.int64)
        self.video_of_pair32_h = self.video_of_pair_h.astype(i32)
        self.pair_off_video_h = np.concatenate([[0], np.cumsum(self.pairs_per_video)]).astype(np.int64)
        self.device = None
        self._chunks = {}

    _DEVICE_FIELDS = ("frame_off", "win_off", "win_src", "win_pos", "latter_src", "inv_latter2", "inv_latter", "pair_win2",
                      "video_of_pair32", "video_of_pair")

    def stat_chunks(self, rows_per_pair, chunk_rows=2048):
        """int32 device table [n_chunks,3] = (row_begin, row_end, video) over the rows of a
        [N*rows_per_pair, C] activation: every chunk lies inside one video (per-video BatchNorm
        statistics)."""
        key = (rows_per_pair, chunk_rows)
        t = self._chunks.get(key)
        if t is None:
            rows = []
            for v in range(self.V):
                r0 = int(self.pair_off_video_h[v]) * rows_per_pair
                r1 = int(self.pair_off_video_h[v + 1]) * rows_per_pair
                starts = np.arange(r0, r1, chunk_rows, dtype=np.int64)
                rows.append(np.stack([starts, np.minimum(starts + chunk_rows, r1), np.full_like(starts, v)], 1))
            from . import ops
            t = ops.upload(np.concatenate(rows).astype(np.int32), self.device)
            self._chunks[key] = t
        return t

    def to(self, device):
        """One H2D copy for all int32 arrays through a persistent pinned staging buffer (allocating
        pinned memory per call would synchronise the whole device, including copy streams)."""
        names = [n for n in self._DEVICE_FIELDS if n != "video_of_pair"]
        arrays = [np.ascontiguousarray(getattr(self, n + "_h")).reshape(-1) for n in names]
        sizes = [a.shape[0] for a in arrays]
        dflat = _staged_h2d(np.concatenate(arrays), device)
        pos = 0
        for n, sz in zip(names, sizes):
            t = dflat[pos:pos + sz]
            h = getattr(self, n + "_h")
            setattr(self, n, t.view(h.shape))
            pos += sz
        from . import ops
        self.video_of_pair = ops.upload(self.video_of_pair_h, device)
        self.pairs_per_video_dev = ops.upload(self.pairs_per_video.astype(np.float32), device)
        self.device = device
        return self


_STAGE = {"buf": None, "event": None}


def _staged_h2d(flat_np, device):
    """int32 numpy vector -> device tensor via a reused pinned buffer (async copy on the current stream)."""
    from . import ops
    return ops.upload(flat_np, device)


def plan_from_im_idx(im_idx, frames_per_video=None, counts_host=None):
    """Build the plan from an entry's `im_idx` (sorted fp32 frame id per pair).

    If the per-frame counts are already known on the host (`counts_host`), no device->host sync is
    needed; otherwise the frame offsets are computed on the device (b200vsgg_frame_offsets, or
    torch.bincount for CPU tensors) and read back once — the reference syncs ~8x per frame here."""
    if counts_host is None:
        if im_idx.is_cuda:
            from . import ops
            n_frames = int(frames_per_video.sum()) if frames_per_video is not None else int(im_idx[-1].item()) + 1
            off = ops.frame_offsets(im_idx.contiguous(), n_frames).cpu().numpy().astype(np.int64)
            counts_host = np.diff(off)
        else:
            counts_host = torch.bincount(im_idx.to(torch.int64)).numpy()
    if frames_per_video is None:
        frames_per_video = np.asarray([len(counts_host)])
    return SegmentPlan(counts_host, frames_per_video)


def attention_blocks(seq_off_h, block=64):
    """Host plan of the flash-attention grid: (blk_seq, blk_row0) int32 arrays — block b covers rows
    [blk_row0[b], min(blk_row0[b] + block, seq_off[blk_seq[b] + 1])) of sequence blk_seq[b]."""
    off = np.asarray(seq_off_h, dtype=np.int64)
    lens = np.diff(off)
    nblk = (lens + block - 1) // block
    blk_seq = np.repeat(np.arange(lens.shape[0]), nblk)
    first = np.concatenate([[0], np.cumsum(nblk)])[:-1]
    local = np.arange(int(nblk.sum())) - np.repeat(first, nblk)
    blk_row0 = off[blk_seq] + local * block
    return blk_seq.astype(np.int32), blk_row0.astype(np.int32)
