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

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

N = 16419
masks = (torch.rand(N, 2, 27, 27, device="cuda").round() - 0.5)
A1 = torch.empty(N * 196, 128, device="cuda", dtype=torch.bfloat16)
ms = timeit(lambda: ops.mask_im2col(masks, A1))
print("mask_im2col [%d pairs] -> [%d x 128]: %.3f ms = %.2f TB/s written" % (N, N * 196, ms, A1.numel() * 2 / ms / 1e9))
z = torch.randn(N * 49, 128, device="cuda").bfloat16()
A2 = torch.empty(N * 49, 1152, device="cuda", dtype=torch.bfloat16)
ms = timeit(lambda: ops.im2col3x3(z, N, 7, 128, A2))
print("im2col3x3 [%d x 128] -> [%d x 1152]: %.3f ms = %.2f TB/s written" % (N * 49, N * 49, ms, A2.numel() * 2 / ms / 1e9))
# reference check of im2col3x3 against torch unfold on a small case
n = 5
zz = torch.randn(n * 49, 128, device="cuda").bfloat16()
out = torch.empty(n * 49, 1152, device="cuda", dtype=torch.bfloat16)
ops.im2col3x3(zz, n, 7, 128, out)
x = zz.view(n, 7, 7, 128).permute(0, 3, 1, 2).float()
ref = torch.nn.functional.unfold(x, 3, padding=1).view(n, 128, 9, 49).permute(0, 3, 2, 1).reshape(n * 49, 1152)
print("im2col3x3 equals unfold:", torch.equal(out.float(), ref))
