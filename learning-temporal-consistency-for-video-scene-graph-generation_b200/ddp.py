"""Video-sharded data parallelism for the relation path.

One process per GPU; every rank owns a disjoint shard of the step's videos (videos are independent
units: windows, BatchNorm statistics and losses never cross a video boundary), so the data path has
NO collective.  The only exchange is the training gradient all-reduce (average over the ranks) over
NCCL / NVLink, issued on NCCL's stream so that it overlaps whatever is still running on the compute stream.

The reference has no distributed code at all (single process, `cuda:0`, TEMPURA_train.py:38); its
optimiser skips parameters whose gradient is None (tools/utils/AdamW.py:66-67), so parameters that
never receive a gradient (frozen object classifier, unused memory attention) are left out of the
buckets instead of being all-reduced as zeros.

What keeps the exchange off the critical path:
  * LAYER BUCKETS: the four weight gradients of a transformer layer (in_proj, out_proj, linear1, linear2: 85 % of the
    gradient bytes) are written by the hand-written backward DIRECTLY into one persistent flat buffer per layer
    (`model._grad_alloc`): autograd adopts those views as `.grad` (no copy in, no copy out), and the moment a layer's
    backward has finished, ONE all-reduce (AVG, in place) of its buffer is issued — 4 large collectives overlap the rest
    of backward instead of 17 medium ones;
  * everything else (biases, norms, heads, front-end) rides in flat buckets after backward: one `_foreach_copy_` in,
    AVG all-reduce, one `_foreach_copy_` out per bucket;
  * a fingerprint (count, numel) of the participating gradients is compared across ranks first, so a rank with a
    different set of gradients raises instead of hanging NCCL.
State consistency (torch DDP's broadcast_buffers): `broadcast_state(model)` copies rank 0's parameters and buffers to
every rank at start; `sync_buffers(model)` averages the floating-point buffers (BatchNorm running statistics, which
each rank updates from its own videos) and takes the maximum of integer counters before saving / evaluating.
"""
import torch
import torch.distributed as dist


def shard_videos(num_videos, rank, world_size):
    """Contiguous, balanced shard of video indices for `rank` (sizes differ by at most one)."""
    base, rem = divmod(num_videos, world_size)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


def _avg_all_reduce(t, group, world, async_op=True):
    """In-place average over the ranks.  NCCL has ReduceOp.AVG; gloo (CPU tests) sums and divides."""
    if t.is_cuda:
        return dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group, async_op=async_op), False
    return dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=async_op), True


def broadcast_state(module, src=0, process_group=None):
    """Copy rank `src`'s parameters and buffers to every rank (start of training / after loading a checkpoint)."""
    if not (dist.is_initialized() and dist.get_world_size(process_group) > 1):
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src, group=process_group)


def sync_buffers(module, process_group=None):
    """Make module buffers identical on all ranks: floating-point buffers (BatchNorm running_mean / running_var, which
    every rank updated from its own shard of videos) are averaged, integer counters (num_batches_tracked) take the
    maximum.  Call before saving a checkpoint or evaluating; cheap enough to call every step."""
    if not (dist.is_initialized() and dist.get_world_size(process_group) > 1):
        return
    world = dist.get_world_size(process_group)
    with torch.no_grad():
        fl = [b for b in module.buffers() if torch.is_floating_point(b)]
        it = [b for b in module.buffers() if not torch.is_floating_point(b)]
        if fl:
            flat = torch.cat([b.reshape(-1).float() for b in fl])
            work, div = _avg_all_reduce(flat, process_group, world, async_op=False)
            if div:
                flat /= world
            pos = 0
            for b in fl:
                b.copy_(flat[pos:pos + b.numel()].view_as(b))
                pos += b.numel()
        for b in it:
            dist.all_reduce(b, op=dist.ReduceOp.MAX, group=process_group)


class GradSync:
    """Gradient averaging over the ranks (see the module docstring).

    `params`: parameters in the order their gradients become available in backward (heads first,
    front-end last); the flat buckets of `sync()` are filled in that order."""

    def __init__(self, params, bucket_bytes=64 << 20, process_group=None, check_fingerprint=True):
        self.params = [p for p in params if p.requires_grad]
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.bucket_bytes = bucket_bytes
        self.check_fingerprint = check_fingerprint
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.cuda = dev.type == "cuda"
        self.stream = torch.cuda.Stream(device=dev) if self.cuda else None
        self._flat = {}
        self._early_works, self._early_ptrs = [], set()
        self._layer_flat, self._layer_views, self._layer_of_ptr = [], {}, {}
        self._fingerprint_ok = None
        self.comm_events = []         # (start, end) CUDA events around the post-backward exchange (bench: comm_exposed_ms)

    # ------------------------------------------------------------------ overlap with backward
    def attach(self, model, layer_groups=None):
        """Let `model` hand over gradient tensors the moment its hand-written backward has finished them.
        layer_groups: list of parameter lists (default: `model.grad_layer_groups()` if it exists); each group gets one
        persistent flat buffer whose views the backward writes into (`model._grad_alloc(param)`), reduced by ONE
        all-reduce when the group is handed over.  Other tensors >= 256 kB handed to the hook are reduced individually;
        `sync()` later waits for these and reduces whatever is left (heads, small tensors)."""
        self._early_works, self._early_ptrs = [], set()
        if self.world > 1:
            if layer_groups is None and hasattr(model, "grad_layer_groups"):
                layer_groups = model.grad_layer_groups()
            for grp in layer_groups or []:
                grp = [p for p in grp if p.requires_grad]
                if not grp:
                    continue
                flat = torch.empty(sum(p.numel() for p in grp), dtype=torch.float32, device=grp[0].device)
                pos = 0
                for p in grp:
                    self._layer_views[id(p)] = (len(self._layer_flat), pos, p.numel(), tuple(p.shape))
                    self._layer_of_ptr[flat.data_ptr() + 4 * pos] = len(self._layer_flat)
                    pos += p.numel()
                self._layer_flat.append(flat)
            model._grad_alloc = self._alloc
            model._grad_ready_hook = self._on_ready
            model._grad_flush_hook = self._flush
        return self

    def _alloc(self, param):
        """Gradient view of `param` inside its persistent layer bucket, or None (no bucket, or `param.grad` is still
        set: gradient ACCUMULATION over several backwards takes the ordinary copy path).  A fresh tensor object per
        call, so that autograd's AccumulateGrad finds it unshared and adopts it as `.grad` instead of cloning it."""
        slot = self._layer_views.get(id(param))
        if slot is None or param.grad is not None:
            return None
        li, pos, n, shape = slot
        return self._layer_flat[li][pos:pos + n].view(shape)

    def _flush(self):
        """Order the compute stream after every early all-reduce (called at the end of the model's backward,
        before autograd copies or accumulates the returned tensors)."""
        for work, t, div in self._early_works:
            work.wait()
            if div:
                t /= self.world
        self._early_works = []                            # waited once; _early_ptrs still marks them as reduced

    def _on_ready(self, tensors):
        done_layers = set()
        for t in tensors:
            if t is None:
                continue
            li = self._layer_of_ptr.get(t.data_ptr())
            if li is not None:                            # a view of a layer bucket: reduce the whole bucket once
                if li not in done_layers and self._layer_flat[li].data_ptr() not in self._early_ptrs:
                    done_layers.add(li)
                    flat = self._layer_flat[li]
                    self._early_ptrs.add(flat.data_ptr())
                    work, div = _avg_all_reduce(flat, self.group, self.world)
                    self._early_works.append((work, flat, div))
                continue
            if t.numel() * t.element_size() < (256 << 10):
                continue                                  # small ones ride in the flat buckets of sync()
            # only whole, contiguous tensors are reduced in place: a view may share its base with gradients that are
            # not finished yet (those fall through to the buckets of sync(), where they are copied)
            if t._base is not None or not t.is_contiguous() or t.data_ptr() in self._early_ptrs:
                continue
            self._early_ptrs.add(t.data_ptr())
            work, div = _avg_all_reduce(t, self.group, self.world)
            self._early_works.append((work, t, div))

    def _already_reduced(self, g):
        if g.data_ptr() in self._early_ptrs:
            return True
        li = self._layer_of_ptr.get(g.data_ptr())
        return li is not None and self._layer_flat[li].data_ptr() in self._early_ptrs

    def _buckets(self, with_grad):
        cur, size, out = [], 0, []
        for p in with_grad:
            n = p.grad.numel() * 4
            if cur and size + n > self.bucket_bytes:
                out.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += n
        if cur:
            out.append(cur)
        return out

    def _check(self, with_grad):
        """All ranks must bring the same set of gradients, or the collectives below would mismatch and hang."""
        key = (len(with_grad), sum(p.grad.numel() for p in with_grad))
        if self._fingerprint_ok == key or not self.check_fingerprint:
            return
        dev = with_grad[0].device if with_grad else (self.params[0].device if self.params else "cpu")
        t = torch.tensor([key[0], key[1], -key[0], -key[1]], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        mx_n, mx_e, mn_n, mn_e = t.tolist()
        if mx_n != -mn_n or mx_e != -mn_e:
            raise RuntimeError("GradSync: ranks disagree on the gradients to reduce (this rank: %d tensors / %d elements; "
                               "across ranks %d..%d tensors, %d..%d elements) — a parameter is frozen or unused on some "
                               "ranks only" % (key[0], key[1], -mn_n, mx_n, -mn_e, mx_e))
        self._fingerprint_ok = key            # same layout as last step: checked once

    def sync(self):
        """Average `.grad` of every parameter that has one over all ranks (in place)."""
        if self.world == 1:
            return
        ev0 = ev1 = None
        if self.cuda:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        self._flush()                                     # stream-level waits: the host does not block
        with_grad = []
        for p in self.params:
            if p.grad is None:
                continue
            if self._already_reduced(p.grad):
                continue                                  # already averaged during backward
            with_grad.append(p)
        self._early_works, self._early_ptrs = [], set()
        self._check(with_grad)
        buckets = self._buckets(with_grad)
        if self.cuda:
            self.stream.wait_stream(torch.cuda.current_stream())
            ctx = torch.cuda.stream(self.stream)
        else:
            import contextlib
            ctx = contextlib.nullcontext()
        with ctx:
            works = []
            for bi, bucket in enumerate(buckets):
                total = sum(p.grad.numel() for p in bucket)
                flat = self._flat.get(bi)
                if flat is None or flat.numel() != total:
                    flat = torch.empty(total, dtype=torch.float32, device=bucket[0].device)
                    self._flat[bi] = flat
                views, pos = [], 0
                for p in bucket:
                    n = p.grad.numel()
                    views.append(flat[pos:pos + n].view_as(p.grad))
                    pos += n
                torch._foreach_copy_(views, [p.grad for p in bucket])
                work, div = _avg_all_reduce(flat, self.group, self.world)
                works.append((work, bucket, views, flat, div))
            for work, bucket, views, flat, div in works:
                work.wait()
                if div:
                    flat /= self.world
                torch._foreach_copy_([p.grad for p in bucket], views)
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.stream)
            ev1.record()
            self.comm_events.append((ev0, ev1))
