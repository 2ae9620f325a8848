"""tcgen05 GEMM (b200vsgg_gemm_bf16) against a plain torch fp32 matmul of the same bf16 operands.
Tolerance: fp32 accumulation of exact bf16 products -> only summation order differs; we allow
max-abs 2e-3 * sqrt(K/64) relative to unit-variance operands (stated per test)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, b, a_mn, b_mn):
    A = a.float().t() if a_mn else a.float()
    B = b.float() if b_mn else b.float().t()
    return A.double() @ B.double()


SHAPES = [
    # M, N, K
    (128, 256, 64),
    (256, 512, 128),
    (300, 1936, 1936),
    (1000, 5808, 1936),
    (77, 336, 1936),
    (513, 2048, 1936),
    (260, 512, 12544),
    (49 * 40, 256, 1024),
    (2, 1936, 3872),
]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True)])
def test_gemm_plain(cuda_lib, M, N, K, a_mn, b_mn):
    from b200vsgg import ops
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = torch.randn((K, M) if a_mn else (M, K), generator=g, device="cuda").bfloat16()
    b = torch.randn((K, N) if b_mn else (N, K), generator=g, device="cuda").bfloat16()
    if a_mn and M % 8:
        pytest.skip("MN-major A needs M % 8 == 0 (TMA pitch)")
    if b_mn and N % 8:
        pytest.skip("MN-major B needs N % 8 == 0 (TMA pitch)")
    out = torch.full((M, N), float("nan"), device="cuda")
    ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, out_f32=out)
    torch.cuda.synchronize()
    ref = _ref(a, b, a_mn, b_mn)
    err = (out.double() - ref).abs().max().item()
    tol = 2e-3 * math.sqrt(K / 64.0) + 1e-4
    assert torch.isfinite(out).all()
    assert err <= tol, (err, tol)


def test_gemm_epilogue(cuda_lib):
    from b200vsgg import ops
    M, N, K = 391, 2048, 1936
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randn(M, K, generator=g, device="cuda").bfloat16()
    w = (torch.randn(N, K, generator=g, device="cuda") / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, generator=g, device="cuda")
    res = torch.randn(M, N, generator=g, device="cuda")
    o32 = torch.empty(M, N, device="cuda")
    o16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, bias=bias, act=ops.ACT_RELU, residual=res, out_f32=o32, out_bf16=o16)
    ref = torch.relu(a.float() @ w.float().t() + bias) + res
    assert (o32 - ref).abs().max().item() < 2e-3
    assert (o16.float() - ref).abs().max().item() < 3e-2
    # gelu + bf16 residual + accumulate
    resb = res.bfloat16()
    o32b = torch.ones(M, N, device="cuda")
    ops.gemm(a, w, bias=bias, act=ops.ACT_GELU, residual=resb, out_f32=o32b, accumulate=True, alpha=0.5)
    ref = torch.nn.functional.gelu(0.5 * (a.float() @ w.float().t()) + bias) + resb.float() + 1.0
    assert (o32b - ref).abs().max().item() < 2e-3
    # relu-mask (backward of relu) into strided output slice
    h = torch.randn(M, N, generator=g, device="cuda").bfloat16()
    big = torch.zeros(M, N + 64, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, mask_src=h, mask_mode=ops.MASK_RELU, out_bf16=big[:, 32:32 + N])
    ref = (a.float() @ w.float().t()) * (h.float() > 0)
    assert (big[:, 32:32 + N].float() - ref).abs().max().item() < 3e-2
    assert big[:, :32].abs().max().item() == 0 and big[:, 32 + N:].abs().max().item() == 0


def test_gemm_dropout_determinism(cuda_lib):
    from b200vsgg import ops
    M, N, K = 256, 512, 256
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = torch.randn(N, K, device="cuda").bfloat16()
    o1 = torch.empty(M, N, device="cuda")
    o2 = torch.empty(M, N, device="cuda")
    o0 = torch.empty(M, N, device="cuda")
    ops.gemm(a, w, out_f32=o0)
    ops.gemm(a, w, out_f32=o1, dropout_p=0.1, seed=1234)
    ops.gemm(a, w, out_f32=o2, dropout_p=0.1, seed=1234)
    assert torch.equal(o1, o2)
    dropped = (o1 == 0) & (o0 != 0)
    frac = dropped.float().mean().item()
    assert 0.08 < frac < 0.12, frac
    kept = ~dropped
    assert torch.allclose(o1[kept], o0[kept] / 0.9, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True)])
@pytest.mark.parametrize("M,N,K", [(31799, 1936, 1936), (16419, 5808, 1936), (20000, 2048, 600)])
def test_gemm_two_cta_path(cuda_lib, M, N, K, a_mn, b_mn):
    """Shapes large enough for the cta_group::2 kernel (256x256 tiles on CTA pairs), ragged in M and N, with
    the fused epilogue (bias + fp32 residual + dual fp32/bf16 stores) — against torch fp32 matmul of the same
    bf16 operands."""
    from b200vsgg import ops
    if a_mn:
        M = (M + 7) // 8 * 8          # an MN-major operand's leading dimension must be a multiple of 8 (TMA)
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn((K, M) if a_mn else (M, K), generator=g, device="cuda").bfloat16()
    b = torch.randn((K, N) if b_mn else (N, K), generator=g, device="cuda").bfloat16()
    bias = torch.randn(N, generator=g, device="cuda")
    res = torch.randn(M, N, generator=g, device="cuda")
    o32 = torch.empty(M, N, device="cuda")
    o16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, bias=bias, residual=res, out_f32=o32, out_bf16=o16)
    af = a.float().t() if a_mn else a.float()
    bf = b.float() if b_mn else b.float().t()
    ref = af @ bf + bias[None, :] + res
    tol = 2e-3 * ref.abs().max().item()
    assert (o32 - ref).abs().max().item() <= tol
    assert (o16.float() - ref).abs().max().item() <= tol + 2 ** -7 * ref.abs().max().item()
