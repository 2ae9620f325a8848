"""Host-side logic of the video-sharded data parallelism (b200vsgg/ddp.py) on CPU: shard assignment
and the bucketed gradient all-reduce over gloo with world_size 2 (parameters without a gradient are
skipped, like the reference optimiser does, tools/utils/AdamW.py:66-67)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_videos_balanced_and_disjoint():
    from b200vsgg.ddp import shard_videos
    for n, w in ((512, 8), (10, 4), (3, 4), (64, 1)):
        shards = [shard_videos(n, r, w) for r in range(w)]
        flat = [v for s in shards for v in s]
        assert flat == list(range(n))
        assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200vsgg.ddp import GradSync
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(s)) for s in ((300, 7), (5,), (1000,), (3, 3))]
    for i, p in enumerate(params):
        if i == 1:
            continue                                  # never receives a gradient: must be skipped
        p.grad = torch.full(p.shape, float(rank + 1) * (i + 1))
    sync = GradSync(params, bucket_bytes=4096)        # several buckets
    sync.sync()
    ok = params[1].grad is None
    for i, p in enumerate(params):
        if i != 1:
            ok = ok and torch.allclose(p.grad, torch.full(p.shape, 1.5 * (i + 1)))
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_grad_sync_gloo_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))


def _worker_early(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200vsgg.ddp import GradSync

    class Dummy:
        pass

    model = Dummy()
    big = torch.nn.Parameter(torch.zeros(400, 400))          # >= 256 kB: reduced early through the hook
    small = torch.nn.Parameter(torch.zeros(10))
    sync = GradSync([big, small], bucket_bytes=4096).attach(model)
    gbig = torch.full((400, 400), float(rank + 1))
    model._grad_ready_hook((gbig, None, torch.ones(3)))       # as the model's backward does
    model._grad_flush_hook()
    big.grad = gbig                                           # autograd keeps the (already averaged) buffer
    small.grad = torch.full((10,), float(rank + 1))
    sync.sync()
    out[rank] = bool(torch.allclose(big.grad, torch.full((400, 400), 1.5)) and
                     torch.allclose(small.grad, torch.full((10,), 1.5)))
    dist.destroy_process_group()


def test_grad_sync_early_hook_gloo_world2():
    """Gradients handed over during backward are averaged once (not again in sync()), the rest in sync()."""
    world = 2
    port = _free_port()
    out = mp.Manager().dict()
    mp.spawn(_worker_early, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))
