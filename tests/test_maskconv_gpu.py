"""Channels-last kernels of the spatial-mask branch (csrc/maskconv.cu) and the split-K GEMM against
plain torch fp32 references.  Everything goes through the C-ABI.  Tolerances: kernels that only move
or select bf16 data are bit-exact; kernels that do arithmetic in fp32 and round once to bf16 are
compared after rounding the reference the same way (<= 1 bf16 ulp); fp32 sums 1e-5 relative."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _gen(seed):
    return torch.Generator(device=DEV).manual_seed(seed)


def _nhwc_rows(x):  # [n,C,H,W] -> [n*H*W, C]
    n, c, h, w = x.shape
    return x.permute(0, 2, 3, 1).reshape(n * h * w, c).contiguous()


def test_mask_im2col_bit_exact(cuda_lib):
    from b200vsgg import ops
    g = _gen(1)
    n = 5
    masks = (torch.rand(n, 2, 27, 27, generator=g, device=DEV).round() - 0.5)
    out = torch.empty(n * 196, 128, device=DEV, dtype=torch.bfloat16)
    ops.mask_im2col(masks, out)
    ref = F.unfold(masks, kernel_size=7, padding=3, stride=2)         # [n, 98, 196]
    ref = ref.transpose(1, 2).reshape(n * 196, 98)
    assert torch.equal(out[:, :98].float(), ref)
    assert out[:, 98:].abs().max().item() == 0


def test_seg_colstats_and_colsum(cuda_lib):
    from b200vsgg import ops
    from b200vsgg.plan import SegmentPlan
    g = _gen(2)
    plan = SegmentPlan([2, 3, 1, 4, 2, 2, 5], [3, 2, 2]).to(DEV)      # 3 videos: 6, 5, 7 pairs
    rows_per_pair = 49
    rows = plan.N * rows_per_pair
    for cols in (128, 256):
        a = torch.randn(rows, cols, generator=g, device=DEV).bfloat16()
        b = torch.randn(rows, cols, generator=g, device=DEV).bfloat16()
        s1 = torch.zeros(plan.V, cols, device=DEV)
        s2 = torch.zeros(plan.V, cols, device=DEV)
        ops.seg_colstats(a, plan.stat_chunks(rows_per_pair, chunk_rows=100), s1, b, s2)
        vid = plan.video_of_pair.repeat_interleave(rows_per_pair)
        r1 = torch.zeros(plan.V, cols, device=DEV).index_add_(0, vid, a.float())
        r2 = torch.zeros(plan.V, cols, device=DEV).index_add_(0, vid, a.float() * b.float())
        assert torch.allclose(s1, r1, rtol=1e-5, atol=1e-3)
        assert torch.allclose(s2, r2, rtol=1e-5, atol=1e-3)
    # plain column sums (bias gradients) route through the same kernel, fp32 and bf16 inputs
    x = torch.randn(3001, 1936, generator=g, device=DEV)
    out = torch.zeros(1, 1936, device=DEV)
    ops.colsum(x, out)
    assert torch.allclose(out[0], x.sum(0), rtol=1e-5, atol=1e-3)
    xb = x.bfloat16()
    out.zero_()
    ops.colsum(xb, out)
    assert torch.allclose(out[0], xb.float().sum(0), rtol=1e-5, atol=1e-3)


def test_bn_pool_fwd_and_bwd(cuda_lib):
    from b200vsgg import ops
    g = _gen(3)
    n, C, V = 7, 128, 3
    vid = torch.tensor([0, 0, 1, 1, 1, 2, 2], device=DEV, dtype=torch.int32)
    y = torch.randn(n, C, 14, 14, generator=g, device=DEV).relu().bfloat16()
    scale = torch.randn(V, C, generator=g, device=DEV)                # negative scales included
    shift = torch.randn(V, C, generator=g, device=DEV)
    y_rows = _nhwc_rows(y.float()).bfloat16()
    z = torch.empty(n * 49, C, device=DEV, dtype=torch.bfloat16)
    arg = torch.empty(n * 49, C, device=DEV, dtype=torch.uint8)
    ops.bn_pool_fwd(y_rows, scale, shift, vid, n, 14, C, z, arg)
    x = torch.addcmul(shift[vid.long()][:, :, None, None], scale[vid.long()][:, :, None, None], y.float())
    x.requires_grad_(True)
    ref, ref_idx = F.max_pool2d(x, 3, 2, 1, return_indices=True)
    assert torch.equal(z.float(), _nhwc_rows(ref).bfloat16().float())
    # argmax code kh*3+kw -> flat input index, must equal torch's choice (first maximum)
    code = arg.view(n, 7, 7, C).permute(0, 3, 1, 2).long()
    oh = torch.arange(7, device=DEV)[None, None, :, None]
    ow = torch.arange(7, device=DEV)[None, None, None, :]
    flat = (oh * 2 - 1 + code // 3) * 14 + (ow * 2 - 1 + code % 3)
    assert torch.equal(flat, ref_idx)
    # backward: route dz to the argmax positions
    dz = torch.randn(n, C, 7, 7, generator=g, device=DEV).bfloat16()
    ref.backward(dz.float())
    dy = torch.empty(n * 196, C, device=DEV, dtype=torch.bfloat16)
    ops.pool_bwd(_nhwc_rows(dz.float()).bfloat16(), arg, n, 14, C, dy)
    assert torch.equal(dy.float(), _nhwc_rows(x.grad).bfloat16().float())


def test_im2col3x3_and_col2im(cuda_lib):
    from b200vsgg import ops
    g = _gen(4)
    n, C = 6, 128
    z = torch.randn(n, C, 7, 7, generator=g, device=DEV).bfloat16()
    out = torch.empty(n * 49, 9 * C, device=DEV, dtype=torch.bfloat16)
    ops.im2col3x3(_nhwc_rows(z.float()).bfloat16(), n, 7, C, out)
    ref = F.unfold(z.float(), 3, padding=1)                           # [n, C*9, 49], column c*9+k
    ref = ref.view(n, C, 9, 49).permute(0, 3, 2, 1).reshape(n * 49, 9 * C)
    assert torch.equal(out.float(), ref)
    dcol = torch.randn(n * 49, 9 * C, generator=g, device=DEV).bfloat16()
    dz = torch.empty(n * 49, C, device=DEV, dtype=torch.bfloat16)
    ops.col2im3x3(dcol, n, 7, C, dz)
    cols = dcol.float().view(n, 49, 9, C).permute(0, 3, 2, 1).reshape(n, C * 9, 49)
    refz = F.fold(cols, (7, 7), 3, padding=1)
    assert torch.allclose(dz.float(), _nhwc_rows(refz).bfloat16().float(), rtol=2 ** -7, atol=1e-2)


def test_seg_affine(cuda_lib):
    from b200vsgg import ops
    g = _gen(5)
    V, C, rpu = 3, 256, 49
    units = torch.tensor([0, 1, 1, 2, 2, 2], device=DEV, dtype=torch.int32)
    rows = units.numel() * rpu
    a = torch.randn(rows, C, generator=g, device=DEV).bfloat16()
    b = torch.randn(rows, C, generator=g, device=DEV).relu().bfloat16()
    k1, k2, k3 = (torch.randn(V, C, generator=g, device=DEV) for _ in range(3))
    gi = units.long().repeat_interleave(rpu)
    out = torch.empty_like(a)
    ops.seg_affine(None, b, None, k2, k3, units, rpu, out)
    ref = torch.addcmul(k3[gi], k2[gi], b.float())
    assert torch.allclose(out.float(), ref.bfloat16().float(), rtol=2 ** -7, atol=1e-6)
    ops.seg_affine(a, b, k1, k2, k3, units, rpu, out, relu_mask=True)
    ref = (k1[gi] * a.float() + k2[gi] * b.float() + k3[gi]) * (b.float() > 0)
    assert torch.allclose(out.float(), ref.bfloat16().float(), rtol=2 ** -6, atol=1e-3)


def test_gemm_k_periodic_split_precision(cuda_lib):
    """a_k_period: one bf16 copy of A against B = [W_hi | W_lo] reproduces the fp32 product A @ W^T
    when A is exact in bf16 (mask values), to ~1e-5 relative instead of bf16's 4e-3."""
    from b200vsgg import ops, tempura
    g = _gen(7)
    a = (torch.rand(5000, 128, generator=g, device=DEV).round() - 0.5).bfloat16()
    w = torch.randn(128, 128, generator=g, device=DEV) * 0.1
    out = torch.empty(5000, 128, device=DEV)
    ops.gemm(a, tempura._split_bf16(w), out_f32=out, a_k_period=128)
    ref = a.double() @ w.double().t()
    assert (out.double() - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()


@pytest.mark.parametrize("M,N,K", [(256, 1024, 40000), (128, 104, 70000), (336, 1936, 9000)])
def test_gemm_split_k_wgrad(cuda_lib, M, N, K):
    """Weight-gradient shapes (few tiles, long K) take the split-K path: fp32 atomics into out_f32."""
    from b200vsgg import ops
    g = _gen(6)
    a = (torch.randn(K, M, generator=g, device=DEV) * 0.1).bfloat16()
    b = torch.randn(K, N, generator=g, device=DEV).bfloat16()
    out = torch.full((M, N), 7.0, device=DEV)                          # must be overwritten, not accumulated
    ops.gemm(a, b, a_mn=True, b_mn=True, out_f32=out)
    ref = a.float().t() @ b.float()
    err = (out - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item() + 1e-3, err
    out1 = torch.empty(M, N, device=DEV)
    ops.gemm(a, b, a_mn=True, b_mn=True, out_f32=out1, split_k=1)      # unsplit path agrees
    assert (out1 - out).abs().max().item() <= 2e-3 * ref.abs().max().item() + 1e-3


def test_mask_branch_matches_torch_conv_stack(cuda_lib):
    """The whole native branch (im2col -> GEMM -> per-video BN -> pool -> im2col -> GEMM -> BN) vs the
    reference layer stack nn.Sequential(Conv,ReLU,BN,MaxPool,Conv,ReLU,BN) run video by video in
    train mode, forward and parameter gradients."""
    from b200vsgg import synthetic, tempura
    kw = dict(mode="predcls", attention_class_num=3, spatial_class_num=6, contact_class_num=17, enc_layer_num=1,
              dec_layer_num=1, obj_mem_compute=False, rel_mem_compute=None, mem_fusion=None, selection="manual",
              K=2, tracking=False)
    m = tempura.TEMPURA(obj_classes=synthetic.ag_object_classes(), **kw)
    synthetic.seeded_init_(m, 5)
    m = m.cuda().train()
    entries = [synthetic.make_video_entry(60 + i, f, ppf, device=DEV) for i, (f, ppf) in enumerate([(3, (2, 4)), (4, (1, 3))])]
    batch = tempura.collate_entries(entries)
    import copy
    ref_conv = copy.deepcopy(m.conv).train()
    from b200vsgg.plan import plan_from_im_idx
    plan = plan_from_im_idx(batch["im_idx"], batch["video_frames"]).to(DEV)
    runner = tempura._PathRunner(m, batch, plan, train_dropout=False, save=True)
    P = runner._unpack(m._path_params())
    W = {"c1x": tempura._split_bf16(F.pad(P["c1_w"].detach().reshape(128, 98), (0, 30))),
         "c2": runner._bf(P["c2_w"].detach().permute(0, 2, 3, 1).reshape(256, 1152))}
    cm = runner._mask_branch_fwd(P, W)
    refs = [ref_conv(e["spatial_masks"]) for e in entries]                  # per-video batch statistics
    ref = torch.cat(refs)
    got = cm.float().view(plan.N, 49, 256).permute(0, 2, 1).reshape(plan.N, 256, 7, 7)
    err = (got - ref).abs().max().item()
    assert err <= 3e-2 * ref.abs().max().item(), err
    assert torch.allclose(m.conv[2].running_mean, ref_conv[2].running_mean, rtol=1e-2, atol=1e-4)
    assert torch.allclose(m.conv[6].running_var, ref_conv[6].running_var, rtol=1e-2, atol=1e-4)
    # backward
    dref = torch.randn(ref.shape, generator=_gen(9), device=DEV)
    ref.backward(dref)
    G = {}
    dcm = _nhwc_rows(dref).bfloat16()
    runner._mask_branch_bwd(dcm, P, W, G)
    names = {"c1_w": "0.weight", "c1_b": "0.bias", "bn1_g": "2.weight", "bn1_b": "2.bias", "c2_w": "4.weight",
             "c2_b": "4.bias", "bn2_g": "6.weight", "bn2_b": "6.bias"}
    refp = dict(ref_conv.named_parameters())
    errs = {}
    for k, n in names.items():
        r = refp[n].grad
        errs[k] = (G[k].reshape(r.shape) - r).norm().item() / r.norm().item()
    print("mask-branch gradient rel-L2 errors:", errs)
    # The only discontinuous steps are the ReLU gates: a conv2 pre-activation within bf16 rounding of 0
    # can gate differently from the fp32 reference (measured: 0.07 % of gates), and the random-init
    # BatchNorm here has channels with rstd ~ 30 that magnify each flipped element, so rel-L2 over a
    # RANDOM upstream gradient is dominated by those few elements (tools/debug_maskconv.py prints the
    # per-stage breakdown: every continuous stage agrees to < 1 %).  Hence 0.15 here; the end-to-end
    # gradient check with a real loss is tests/test_tempura_gpu.py::test_backward_matches_oracle.
    assert max(errs.values()) <= 0.15, errs
    y2_ref = torch.cat([F.relu(ref_conv[4](ref_conv[3](ref_conv[2](F.relu(ref_conv[0](e["spatial_masks"])))))) for e in entries])
    gates_ref = _nhwc_rows(y2_ref) > 0
    # (running statistics moved by this extra forward; the model copy is discarded)
    flips = (gates_ref != (runner.saved["y2"].float() > 0)).float().mean().item()
    assert flips < 5e-3, flips
