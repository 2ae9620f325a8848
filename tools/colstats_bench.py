import sys, torch
sys.path.insert(0, "/root/repo")
from b200vsgg import ops
rows, V = 3218124, 64
for cols in (128, 256):
    r = rows if cols == 128 else rows // 4
    a = torch.randn(r, cols, device="cuda").bfloat16(); b = torch.randn(r, cols, device="cuda").bfloat16()
    chunks = ops.uniform_chunks(r, a.device)
    s1 = torch.zeros(1, cols, device="cuda"); s2 = torch.zeros(1, cols, device="cuda")
    for _ in range(3): ops.seg_colstats(a, chunks, s1, b, s2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.seg_colstats(a, chunks, s1, b, s2)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("seg_colstats [%d x %d] bf16 x2: %.3f ms = %.2f TB/s" % (r, cols, ms, 2 * r * cols * 2 / ms / 1e9))
    ref = a.float().sum(0)
    s1.zero_(); s2.zero_(); ops.seg_colstats(a, chunks, s1, b, s2)
    print("  max rel err", ((s1[0] - ref).abs().max() / ref.abs().max()).item(), ((s2[0] - (a.float() * b.float()).sum(0)).abs().max() / (a.float() * b.float()).sum(0).abs().max()).item())
