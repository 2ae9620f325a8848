"""Varlen window attention micro-benchmark (TEMPURA temporal windows: 2 frames = 12..20 tokens, 8 heads x 242):
CUDA-event time of b200vsgg_attn_small_{fwd,bwd} and achieved HBM bandwidth vs the algorithmic bytes
(fwd 8*D B/token, bwd 14*D B/token)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from b200vsgg import ops
rng = np.random.default_rng(0)
H, hd = 8, 242
D = H * hd
lens = rng.integers(6, 11, size=1984) + rng.integers(6, 11, size=1984)
off = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device="cuda")
M = int(lens.sum())
qkv = torch.randn(M, 3 * D, device="cuda").bfloat16()
q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
ctx = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
do = torch.randn(M, D, device="cuda").bfloat16()
dqkv = torch.empty(M, 3 * D, device="cuda", dtype=torch.bfloat16)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for p in (0.0, 0.1):
    f = timeit(lambda: ops.attn_small_fwd(q, k, v, off, len(lens), int(lens.max()), H, hd, ctx, p, 7))
    b = timeit(lambda: ops.attn_small_bwd(q, k, v, do, off, len(lens), int(lens.max()), H, hd, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], p, 7))
    print("tokens %d max_len %d dropout %.1f: fwd %.3f ms (%.2f TB/s of algorithmic bytes)  bwd %.3f ms (%.2f TB/s)" % (
        M, lens.max(), p, f, 8 * D * M / f / 1e9, b, 14 * D * M / b / 1e9))
