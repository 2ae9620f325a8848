"""Run one GEMM shape a few times (for ncu captures):  python tools/gemm_one.py M N K a_mn b_mn [res|mask|f32]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200vsgg import ops
M, N, K, a_mn, b_mn = (int(x) for x in sys.argv[1:6])
mode = sys.argv[6] if len(sys.argv) > 6 else ""
a = torch.randn((K, M) if a_mn else (M, K), device="cuda").bfloat16()
b = torch.randn((K, N) if b_mn else (N, K), device="cuda").bfloat16()
kw = {}
if "f32" in mode:
    kw["out_f32"] = torch.empty(M, N, device="cuda")
else:
    kw["out_bf16"] = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
if "res" in mode:
    kw["residual"] = torch.randn(M, N, device="cuda").bfloat16()
if "mask" in mode:
    kw["mask_src"] = torch.randn(M, N, device="cuda").bfloat16(); kw["mask_mode"] = 1
for _ in range(3):
    ops.gemm(a, b, a_mn=bool(a_mn), b_mn=bool(b_mn), **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.gemm(a, b, a_mn=bool(a_mn), b_mn=bool(b_mn), **kw)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("%d %d %d a%d b%d %s: %.4f ms %.1f TF" % (M, N, K, a_mn, b_mn, mode, ms, 2.0 * M * N * K / ms / 1e9))
