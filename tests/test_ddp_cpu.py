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


def _worker_layers(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200vsgg import ddp
    torch.manual_seed(rank)                                  # ranks start DIFFERENT on purpose
    net = torch.nn.Sequential(torch.nn.Linear(300, 300), torch.nn.BatchNorm1d(300), torch.nn.Linear(300, 4))
    ddp.broadcast_state(net)                                 # ... and are made identical here
    a, b = net[0].weight, net[2].weight
    net.grad_layer_groups = lambda: [[a], [b]]               # two "layers", one flat bucket each
    sync = ddp.GradSync(list(net.parameters())[::-1], bucket_bytes=4096).attach(net)
    ok = True
    for step in range(2):
        net.zero_grad(set_to_none=True)
        ga, gb = net._grad_alloc(a), net._grad_alloc(b)      # what the hand-written backward does
        ok = ok and ga is not None and ga.shape == a.shape and net._grad_alloc(a) is not ga
        ga.fill_(float(rank + 1))
        gb.fill_(float(10 * (rank + 1)))
        net._grad_ready_hook((ga,))
        net._grad_ready_hook((gb,))
        net._grad_flush_hook()
        a.grad, b.grad = ga, gb                              # autograd adopts the views
        for p in (net[0].bias, net[1].weight, net[1].bias, net[2].bias):
            p.grad = torch.full(p.shape, float(rank + 1))
        sync.sync()
        ok = ok and torch.allclose(a.grad, torch.full(a.shape, 1.5)) and torch.allclose(b.grad, torch.full(b.shape, 15.0))
        ok = ok and all(torch.allclose(p.grad, torch.full(p.shape, 1.5)) for p in (net[0].bias, net[1].weight, net[2].bias))
        ok = ok and net._grad_alloc(a) is None               # .grad still set: accumulation takes the copy path
    # each rank updates BatchNorm running statistics from its own shard ...
    net.train()
    net(torch.randn(16, 300) + rank)
    differ = not torch.allclose(net[1].running_mean, torch.zeros(300))
    ddp.sync_buffers(net)                                    # ... and they are averaged before saving / evaluating
    flat = torch.cat([t.detach().reshape(-1).float() for t in net.state_dict().values()])
    both = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(both, flat)
    ok = ok and differ and torch.equal(both[0], both[1])     # identical state_dict on both ranks
    # a rank with a different set of gradients must raise instead of hanging the collective
    net.zero_grad(set_to_none=True)
    for p in net.parameters():
        p.grad = torch.ones_like(p)
    if rank == 1:
        net[2].bias.grad = None
    sync2 = ddp.GradSync(list(net.parameters()), bucket_bytes=1 << 20)
    try:
        sync2.sync()
        raised = False
    except RuntimeError as ex:
        raised = "disagree" in str(ex)
    out[rank] = bool(ok and raised)
    dist.destroy_process_group()


def test_layer_buckets_state_broadcast_and_fingerprint_gloo_world2():
    """Layer buckets (gradients written into persistent flat buffers, one collective per layer, adopted as .grad
    without copies), broadcast_state / sync_buffers (identical state_dict on all ranks after a step that updated
    BatchNorm statistics per rank), and the fingerprint check (mismatching gradient sets raise)."""
    world = 2
    port = _free_port()
    out = mp.Manager().dict()
    mp.spawn(_worker_layers, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))
