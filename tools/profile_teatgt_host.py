import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200vsgg import teatgt
orig = teatgt.TeatPlan.build_graph
acc = {"build": 0.0, "n": 0}
def timed(self, *a, **k):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = orig(self, *a, **k)
    acc["build"] += time.perf_counter() - t0; acc["n"] += 1
    return r
teatgt.TeatPlan.build_graph = timed
sys.argv = ["bench_teatgt.py", "--steps", "3", "--warmup", "1"] + sys.argv[1:]
exec(open(os.path.join(os.path.dirname(__file__), "bench_teatgt.py")).read())
print("build_graph (edges + eigh) per call: %.1f ms over %d calls" % (1e3 * acc["build"] / acc["n"], acc["n"]))
# GPU time of one step via events around kernels only (no host): run once more under profiler
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60))
