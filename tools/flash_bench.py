import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from b200vsgg import ops
from b200vsgg.plan import attention_blocks
H, hd = 32, 24
rng = np.random.default_rng(0)
lens = rng.integers(250, 460, 448)
off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
bs, br = attention_blocks(off)
t = lambda a: torch.from_numpy(a).cuda()
offd, bsd, brd = t(off), t(bs), t(br)
M, D = int(off[-1]), H * hd
qkv = torch.randn(M, 3 * D, device="cuda").bfloat16()
q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
ctx = torch.empty(M, D, device="cuda", dtype=torch.bfloat16); lse = torch.empty(M, H, device="cuda")
dctx = torch.randn(M, D, device="cuda").bfloat16(); dqkv = torch.empty_like(qkv)
flops = 4.0 * float((lens.astype(np.float64) ** 2).sum()) * H * hd
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for p in (0.0, 0.1):
    for ml, name in ((0, "tiled"), (int(lens.max()), "resident")):
        f = timeit(lambda: ops.attn_flash_fwd(q, k, v, offd, bsd, brd, H, hd, ctx, lse, p, 7, max_len=ml))
        b = timeit(lambda: ops.attn_flash_bwd(q, k, v, ctx, dctx, lse, offd, bsd, brd, H, hd, dqkv[:, :D], dqkv[:, D:2*D], dqkv[:, 2*D:], p, 7, max_len=ml))
        print("p=%.1f %-8s fwd %.3f ms (%.1f TF/s)  bwd %.3f ms (%.1f TF/s)  tokens %d" % (p, name, f, flops / f / 1e9, b, 2.5 * flops / b / 1e9, M))
