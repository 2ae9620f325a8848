"""One launch of each tcgen05 attention kernel at a long-clip shape (for ncu): python tools/attn_tc_profile.py [p]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from b200vsgg import ops
from b200vsgg.plan import attention_blocks
p = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
H, hd = 32, 24
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
print("ok", M)
