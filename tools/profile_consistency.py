import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from b200vsgg import synthetic, tempura
dev = torch.device("cuda", 0)
m = tempura.TEMPURA(obj_classes=synthetic.ag_object_classes(), consistency_regulariser=True, **bench.MODEL_KW)
synthetic.seeded_init_(m, 1123); m = m.to(dev).train()
batch = bench.build_batch(list(range(64)), 32, dev)
def step():
    m.zero_grad(set_to_none=True)
    pred = m(dict(batch), phase="train")
    l = tempura.tempura_loss(pred, m.last_plan)
    (l["attention_relation_loss"] + l["spatial_relation_loss"] + l["contacting_relation_loss"]).backward()
for _ in range(2): step()
torch.cuda.synchronize()
orig = m._consistency
acc = []
def timed(*a, **k):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = orig(*a, **k); torch.cuda.synchronize(); acc.append(time.perf_counter() - t0); return r
m._consistency = timed
for _ in range(3): step()
print("consistency wall ms:", [round(x * 1e3, 1) for x in acc])
m._consistency = orig
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    with torch.no_grad():
        orig(dict(batch), m.last_plan, torch.randn(batch["pair_idx"].shape[0], 1936, device=dev))
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=50))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=10, max_name_column_width=50))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
with torch.no_grad():
    orig(dict(batch), m.last_plan, torch.randn(batch["pair_idx"].shape[0], 1936, device=dev))
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
