"""TokenGT varlen attention: mma.sync kernels (attn_flash.cu) vs the tcgen05 / TMEM forward (attn_tc.cu), CUDA events.
Shapes: C3 = 448 clips x 250-460 tokens (32 heads x 24, and 16 x 48); LONG = 8 clips x 2000-5300 tokens."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from b200vsgg import ops
from b200vsgg.plan import attention_blocks
t = lambda a: torch.from_numpy(a).cuda()
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
rng = np.random.default_rng(0)
for tag, lens, H, hd in (("C3 32x24", rng.integers(250, 460, 448), 32, 24), ("C3 16x48", rng.integers(250, 460, 448), 16, 48),
                         ("LONG 32x24", rng.integers(2000, 5300, 16), 32, 24)):
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    bs, br = attention_blocks(off)
    bs128, br128 = attention_blocks(off, block=128)
    offd, bsd, brd, bs128d, br128d = t(off), t(bs), t(br), t(bs128), t(br128)
    M, D = int(off[-1]), H * hd
    qkv = torch.randn(M, 3 * D, device="cuda").bfloat16()
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    ctx = torch.empty(M, D, device="cuda", dtype=torch.bfloat16); lse = torch.empty(M, H, device="cuda")
    dctx = torch.randn(M, D, device="cuda").bfloat16(); dqkv = torch.empty_like(qkv)
    flops = 4.0 * float((lens.astype(np.float64) ** 2).sum()) * H * hd
    for p in (0.0, 0.1):
        f = timeit(lambda: ops.attn_flash_fwd(q, k, v, offd, bsd, brd, H, hd, ctx, lse, p, 7))
        c = timeit(lambda: ops.attn_tc_fwd(q, k, v, offd, bs128d, br128d, H, hd, ctx, lse, p, 7))
        b = timeit(lambda: ops.attn_flash_bwd(q, k, v, ctx, dctx, lse, offd, bsd, brd, H, hd, dqkv[:, :D], dqkv[:, D:2*D], dqkv[:, 2*D:], p, 7))
        d = timeit(lambda: ops.attn_tc_bwd(q, k, v, ctx, dctx, lse, offd, bs128d, br128d, H, hd, dqkv[:, :D], dqkv[:, D:2*D], dqkv[:, 2*D:], p, 7))
        print("%-10s p=%.1f  fwd mma.sync %.3f ms (%.1f TF/s) | fwd tcgen05 %.3f ms (%.1f TF/s) | bwd mma.sync %.3f ms (%.1f TF/s) | bwd tcgen05 %.3f ms (%.1f TF/s)  tokens %d" % (
            tag, p, f, flops / f / 1e9, c, flops / c / 1e9, b, 2.5 * flops / b / 1e9, d, 2.5 * flops / d / 1e9, M), flush=True)
