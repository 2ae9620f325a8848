"""One launch of each tcgen05 attention kernel at a long-clip shape (for ncu): python tools/attn_tc_profile.py [p]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from b200vsgg import ops
from b200vsgg.plan import attention_blocks
p = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
H, hd = 32, 24
if len(sys.argv) > 2 and sys.argv[2] == "short":      # the 64-video training batch: 448 clips of 250..589 tokens
    lens = np.random.default_rng(0).integers(250, 590, size=448)
else:                                                   # long clips (BASELINE configs[4])
    lens = np.asarray([3000, 2600, 3300, 2900, 3100, 2800, 3200, 2700])
off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
bs128, br128 = attention_blocks(off, block=128)
t = lambda a: torch.from_numpy(a).cuda()
offd, bsd, brd = t(off), t(bs128), t(br128)
M, D = int(off[-1]), H * hd
qkv = torch.randn(M, 3 * D, device="cuda").bfloat16()
q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
ctx = torch.empty(M, D, device="cuda", dtype=torch.bfloat16); lse = torch.empty(M, H, device="cuda")
dctx = torch.randn(M, D, device="cuda").bfloat16(); dqkv = torch.empty_like(qkv)
for _ in range(2):
    ops.attn_tc_fwd(q, k, v, offd, bsd, brd, H, hd, ctx, lse, p, 7)
    ops.attn_tc_bwd(q, k, v, ctx, dctx, lse, offd, bsd, brd, H, hd, dqkv[:, :D], dqkv[:, D:2*D], dqkv[:, 2*D:], p, 7)
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
for _ in range(5):
    ops.attn_tc_fwd(q, k, v, offd, bsd, brd, H, hd, ctx, lse, p, 7)
e1.record()
for _ in range(5):
    ops.attn_tc_bwd(q, k, v, ctx, dctx, lse, offd, bsd, brd, H, hd, dqkv[:, :D], dqkv[:, D:2*D], dqkv[:, 2*D:], p, 7)
e2.record()
torch.cuda.synchronize()
fl = 4.0 * float((lens.astype(np.float64) ** 2).sum()) * hd * H
print("ok tokens %d clips %d dropout %.2f: fwd %.3f ms (%.0f TFLOP/s)  bwd %.3f ms (%.0f TFLOP/s)" % (
    M, len(lens), p, e0.elapsed_time(e1) / 5, fl / (e0.elapsed_time(e1) / 5 * 1e-3) / 1e12,
    e1.elapsed_time(e2) / 5, 2.5 * fl / (e1.elapsed_time(e2) / 5 * 1e-3) / 1e12))
