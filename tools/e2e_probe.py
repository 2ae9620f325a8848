"""Where does the end-to-end step time go?  Times the H2D copies (copy-stream events) with and without
concurrent compute."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from b200vsgg import synthetic, tempura
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
model = tempura.TEMPURA(obj_classes=synthetic.ag_object_classes(), **bench.MODEL_KW)
synthetic.seeded_init_(model, 1123)
model = model.to(dev).train()
batch = bench.build_batch(list(range(64)), 32, dev)
host = {k: batch[k].cpu().pin_memory() for k in bench.TENSOR_KEYS if k in batch}
for k, v in host.items():
    print(k, tuple(v.shape), v.dtype, v.is_pinned(), "%.1f MB" % (v.numel() * v.element_size() / 1e6))
dst = {k: torch.empty_like(batch[k]) for k in host}
cs = torch.cuda.Stream()
def copy_all():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(cs):
        e0.record(cs)
        for k, t in host.items():
            n, chunk = t.numel(), (16 << 20) // t.element_size()
            if n <= chunk:
                dst[k].copy_(t, non_blocking=True)
            else:
                df, sf = dst[k].view(-1), t.view(-1)
                for i in range(0, n, chunk):
                    df[i:i + chunk].copy_(sf[i:i + chunk], non_blocking=True)
        e1.record(cs)
    return e0, e1
def step():
    model.zero_grad(set_to_none=True)
    pred = model(dict(batch), phase="train")
    l = tempura.tempura_loss(pred, model.last_plan)
    (l["attention_relation_loss"] + l["spatial_relation_loss"] + l["contacting_relation_loss"]).backward()
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = copy_all(); torch.cuda.synchronize()
print("copy alone: %.1f ms" % e0.elapsed_time(e1))
t0 = time.perf_counter(); e0, e1 = copy_all(); t1 = time.perf_counter(); step(); t2 = time.perf_counter(); torch.cuda.synchronize(); t3 = time.perf_counter()
print("copy || compute: copy %.1f ms; host issue copy %.1f ms, host issue step %.1f ms, total %.1f ms" % (e0.elapsed_time(e1), (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t0) * 1e3))
t0 = time.perf_counter(); step(); t2 = time.perf_counter(); torch.cuda.synchronize(); t3 = time.perf_counter()
print("compute alone: host issue %.1f ms total %.1f ms" % ((t2 - t0) * 1e3, (t3 - t0) * 1e3))
