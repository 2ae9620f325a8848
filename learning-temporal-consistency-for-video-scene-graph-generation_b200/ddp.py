"""Video-sharded data parallelism for the relation path.

One process per GPU; every rank owns a disjoint shard of the step's videos (videos are independent
units: windows, BatchNorm statistics and losses never cross a video boundary), so the data path has
NO collective.  The only exchange is the training gradient all-reduce (sum, then divide by the world
size) over NCCL / NVLink, issued bucket by bucket on a side stream so that it overlaps whatever is
still running on the compute stream.

The reference has no distributed code at all (single process, `cuda:0`, TEMPURA_train.py:38); its
optimiser skips parameters whose gradient is None (tools/utils/AdamW.py:66-67), so parameters that
never receive a gradient (frozen object classifier, unused memory attention) are left out of the
buckets instead of being all-reduced as zeros.
"""
import torch
import torch.distributed as dist


def shard_videos(num_videos, rank, world_size):
    """Contiguous, balanced shard of video indices for `rank` (sizes differ by at most one)."""
    base, rem = divmod(num_videos, world_size)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


class GradSync:
    """Bucketed gradient all-reduce.

    `params`: parameters in the order their gradients become available in backward (heads first,
    front-end last).  Buckets are filled in that order; each bucket is one flat fp32 buffer that is
    all-reduced asynchronously on `self.stream` and copied back into the `.grad` tensors."""

    def __init__(self, params, bucket_bytes=64 << 20, process_group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.bucket_bytes = bucket_bytes
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.cuda = dev.type == "cuda"
        self.stream = torch.cuda.Stream(device=dev) if self.cuda else None
        self._flat = {}

    # ------------------------------------------------------------------ overlap with backward
    def attach(self, model):
        """Let `model` (b200vsgg TEMPURA) hand over gradient tensors the moment its hand-written backward has
        finished them: they are all-reduced (AVG) asynchronously on NCCL's stream while the remaining backward
        kernels run.  `sync()` later waits for these and reduces whatever is left (heads, small tensors)."""
        self._early_works, self._early_ptrs = [], set()
        if self.world > 1:
            model._grad_ready_hook = self._on_ready
            model._grad_flush_hook = self._flush
        return self

    def _flush(self):
        """Order the compute stream after every early all-reduce (called at the end of the model's backward,
        before autograd copies or accumulates the returned tensors)."""
        for work, _ in self._early_works:
            work.wait()

    def _on_ready(self, tensors):
        for t in tensors:
            if t is None or t.numel() * t.element_size() < (256 << 10):
                continue                                  # small ones ride in the flat buckets of sync()
            base = t._base if t._base is not None else t
            ptr = base.data_ptr()
            if ptr in self._early_ptrs or not base.is_contiguous():
                continue
            self._early_ptrs.add(ptr)
            self._early_works.append((dist.all_reduce(base, op=dist.ReduceOp.AVG, group=self.group, async_op=True), base))

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

    def sync(self):
        """Average `.grad` of every parameter that has one over all ranks (in place)."""
        if self.world == 1:
            return
        early = getattr(self, "_early_works", [])
        early_ptrs = getattr(self, "_early_ptrs", set())
        for work, _ in early:
            work.wait()                                   # stream-level wait: the host does not block
        with_grad = []
        for p in self.params:
            if p.grad is None:
                continue
            g = p.grad
            base = g._base if g._base is not None else g
            if base.data_ptr() in early_ptrs:
                continue                                  # already averaged during backward
            with_grad.append(p)
        self._early_works, self._early_ptrs = [], set()
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
                works.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True), bucket,
                              views))
            inv = 1.0 / self.world
            for work, bucket, views in works:
                work.wait()
                torch._foreach_mul_(views, inv)
                torch._foreach_copy_([p.grad for p in bucket], views)
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.stream)
