"""Python wrappers (torch tensors in, raw device pointers out) around the C-ABI kernels.
These are plumbing only: every function launches hand-written sm_100a kernels from
libb200vsgg.so on torch's current CUDA stream and raises if the library is missing."""
import ctypes as C

import torch

from . import _lib
from ._lib import GemmEpilogue, check

ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
MASK_RELU, MASK_GELU = 1, 2

# Number of kernel launches issued through this module (bench.py reports it as gpu_launches).
launch_count = 0


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _count(n=1):
    global launch_count
    launch_count += n


def gemm(a, b, *, a_mn=False, b_mn=False, bias=None, residual=None, mask_src=None, mask_mode=0,
         act=ACT_NONE, out_f32=None, out_bf16=None, accumulate=False, alpha=1.0,
         dropout_p=0.0, seed=0):
    """D = epilogue(alpha * op(a) @ op(b)^T); see b200vsgg_gemm_bf16 in include/b200vsgg.h.

    a: [M,K] (a_mn=False) or [K,M] (a_mn=True); b: [N,K] (b_mn=False) or [K,N] (b_mn=True); both
    bf16 with unit inner stride.  At least one of out_f32 / out_bf16 must be given ([M,N])."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.stride(-1) == 1 and b.stride(-1) == 1
    if a_mn:
        K, M = a.shape
    else:
        M, K = a.shape
    if b_mn:
        Kb, N = b.shape
    else:
        N, Kb = b.shape
    assert K == Kb, (a.shape, b.shape, a_mn, b_mn)
    ep = GemmEpilogue()
    ep.bias = bias.data_ptr() if bias is not None else None
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
    if residual is not None:
        assert residual.shape == (M, N) and residual.stride(1) == 1
        ep.residual = residual.data_ptr()
        ep.residual_is_bf16 = 1 if residual.dtype == torch.bfloat16 else 0
        ep.ldr = residual.stride(0)
    if mask_src is not None:
        assert mask_src.dtype == torch.bfloat16 and mask_src.shape == (M, N) and mask_src.stride(1) == 1
        ep.mask_src = mask_src.data_ptr()
        ep.ldm = mask_src.stride(0)
        ep.mask_mode = mask_mode
    ep.act = act
    if out_f32 is not None:
        assert out_f32.dtype == torch.float32 and out_f32.shape == (M, N) and out_f32.stride(1) == 1
        ep.out_f32 = out_f32.data_ptr()
        ep.ld_f32 = out_f32.stride(0)
    if out_bf16 is not None:
        assert out_bf16.dtype == torch.bfloat16 and out_bf16.shape == (M, N) and out_bf16.stride(1) == 1
        ep.out_bf16 = out_bf16.data_ptr()
        ep.ld_bf16 = out_bf16.stride(0)
    ep.accumulate = 1 if accumulate else 0
    ep.alpha = alpha
    ep.dropout_p = dropout_p
    ep.dropout_seed = seed
    rc = _lib.lib().b200vsgg_gemm_bf16(_ptr(a), a.stride(0), int(a_mn), _ptr(b), b.stride(0), int(b_mn),
                                       M, N, K, C.byref(ep), _stream())
    check(rc, "gemm_bf16")
    _count()
    return out_f32 if out_f32 is not None else out_bf16
