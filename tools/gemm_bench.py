"""Micro-benchmark of b200vsgg_gemm_bf16 on the path's GEMM shapes (CUDA-event timed, L2 flushed)."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200vsgg import ops

SHAPES = [  # name, M, N, K, a_mn, b_mn
    ("qkv_fwd", 16384, 5808, 1936, 0, 0),
    ("out_fwd", 16384, 1936, 1936, 0, 0),
    ("ffn1_fwd", 16384, 2048, 1936, 0, 0),
    ("ffn2_fwd", 16384, 1936, 2048, 0, 0),
    ("qkv_dgrad", 16384, 1936, 5808, 0, 1),
    ("qkv_wgrad", 5808, 1936, 16384, 1, 1),
    ("ffn1_wgrad", 2048, 1936, 16384, 1, 1),
    ("vr_fc", 16384, 512, 12544, 0, 0),
    ("union", 16384 * 49, 256, 1024, 0, 0),
    ("square8k", 8192, 8192, 8192, 0, 0),
]


def main():
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    res = []
    for name, M, N, K, a_mn, b_mn in SHAPES:
        a = torch.randn((K, M) if a_mn else (M, K), device="cuda").bfloat16()
        b = torch.randn((K, N) if b_mn else (N, K), device="cuda").bfloat16()
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        for _ in range(3):
            ops.gemm(a, b, a_mn=bool(a_mn), b_mn=bool(b_mn), out_bf16=out)
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.gemm(a, b, a_mn=bool(a_mn), b_mn=bool(b_mn), out_bf16=out)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
        # cuBLAS for comparison
        A2 = a.t() if a_mn else a
        B2 = b if b_mn else b.t()
        for _ in range(3):
            torch.matmul(A2, B2)
        ts2 = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(A2, B2)
            e1.record()
            torch.cuda.synchronize()
            ts2.append(e0.elapsed_time(e1))
        ts2.sort()
        ms2 = ts2[len(ts2) // 2]
        tf2 = 2.0 * M * N * K / (ms2 * 1e-3) / 1e12
        r = dict(name=name, M=M, N=N, K=K, a_mn=a_mn, b_mn=b_mn, ms=round(ms, 4), tflops=round(tf, 1),
                 cublas_ms=round(ms2, 4), cublas_tflops=round(tf2, 1))
        print(json.dumps(r), flush=True)
        res.append(r)
        del a, b, out
    return res


if __name__ == "__main__":
    main()
