import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from b200vsgg import synthetic, tempura, ops
from b200vsgg.plan import plan_from_im_idx
DEV = "cuda"
def rows(x):
    n, c, h, w = x.shape
    return x.permute(0, 2, 3, 1).reshape(n * h * w, c).contiguous()
def rel(a, b): return ((a - b).norm() / b.norm()).item()
kw = dict(mode="predcls", attention_class_num=3, spatial_class_num=6, contact_class_num=17, enc_layer_num=1,
          dec_layer_num=1, obj_mem_compute=False, rel_mem_compute=None, mem_fusion=None, selection="manual", K=2, tracking=False)
m = tempura.TEMPURA(obj_classes=synthetic.ag_object_classes(), **kw)
synthetic.seeded_init_(m, 5)
m = m.cuda().train()
entries = [synthetic.make_video_entry(60 + i, f, ppf, device=DEV) for i, (f, ppf) in enumerate([(3, (2, 4)), (4, (1, 3))])]
batch = tempura.collate_entries(entries)
rc = copy.deepcopy(m.conv).train()
plan = plan_from_im_idx(batch["im_idx"], batch["video_frames"]).to(DEV)
runner = tempura._PathRunner(m, batch, plan, train_dropout=False, save=True)
P = runner._unpack(m._path_params())
W = {"c1x": tempura._split_bf16(F.pad(P["c1_w"].detach().reshape(128, 98), (0, 30))),
     "c2": runner._bf(P["c2_w"].detach().permute(0, 2, 3, 1).reshape(256, 1152))}
cm = runner._mask_branch_fwd(P, W)
S = runner.saved
# reference stage by stage, per video
inter = {k: [] for k in ("y1", "bn1", "z", "y2", "out")}
for e in entries:
    x = e["spatial_masks"]
    y1 = F.relu(rc[0](x)); y1.retain_grad()
    b1 = rc[2](y1); b1.retain_grad()
    z = rc[3](b1); z.retain_grad()
    y2 = F.relu(rc[4](z)); y2.retain_grad()
    o = rc[6](y2)
    for k, v in zip(("y1", "bn1", "z", "y2", "out"), (y1, b1, z, y2, o)): inter[k].append(v)
print("y1", rel(S["y1"].float(), rows(torch.cat(inter["y1"]))))
print("y2", rel(S["y2"].float(), rows(torch.cat(inter["y2"]))))
print("cm", rel(cm.float(), rows(torch.cat(inter["out"]))))
out = torch.cat(inter["out"])
dref = torch.randn(out.shape, generator=torch.Generator(device=DEV).manual_seed(9), device=DEV)
out.backward(dref)
G = {}
dcm = rows(dref).bfloat16()
# replicate _mask_branch_bwd with checks
N = plan.N
d2 = runner._bn_relu_bwd(dcm, S["y2"], 49, P["bn2_g"], "bn2", G, "bn2_g", "bn2_b")
# ref d(conv2 out) = y2.grad * (y2>0)
ref_d2 = rows(torch.cat([v.grad * (v > 0) for v in inter["y2"]]))
print("d2", rel(d2.float(), ref_d2))
print("d2 pre-mask ref vs ours on y>0:", rel(d2.float()[ref_d2 != 0], ref_d2[ref_d2 != 0]))
dA2 = torch.empty(N * 49, 1152, device=DEV, dtype=torch.bfloat16)
ops.gemm(d2, W["c2"], b_mn=True, out_bf16=dA2)
dz = torch.empty(N * 49, 128, device=DEV, dtype=torch.bfloat16)
ops.col2im3x3(dA2, N, 7, 128, dz)
print("dz", rel(dz.float(), rows(torch.cat([v.grad for v in inter["z"]]))))
dpool = torch.empty(N * 196, 128, device=DEV, dtype=torch.bfloat16)
ops.pool_bwd(dz, S["arg"], N, 14, 128, dpool)
print("dpool", rel(dpool.float(), rows(torch.cat([v.grad for v in inter["bn1"]]))))
d1 = runner._bn_relu_bwd(dpool, S["y1"], 196, P["bn1_g"], "bn1", G, "bn1_g", "bn1_b")
ref_d1 = rows(torch.cat([v.grad * (v > 0) for v in inter["y1"]]))
print("d1", rel(d1.float(), ref_d1))
# ---- formula check in fp32 torch
mean, rstd, cnt = S["bn2"]
vid = plan.video_of_pair.repeat_interleave(49)
y2r = rows(torch.cat(inter["y2"])).detach()
dout = rows(dref)
V = plan.V
s1 = torch.zeros(V, 256, device=DEV).index_add_(0, vid, dout)
s2 = torch.zeros(V, 256, device=DEV).index_add_(0, vid, dout * y2r)
g = P["bn2_g"].detach()[None]
print("mean err", rel(mean, torch.stack([v.mean((0, 2, 3)) for v in inter["y2"]])))
print("rstd err", rel(rstd, torch.stack([torch.rsqrt(v.var((0, 2, 3), unbiased=False) + 1e-5) for v in inter["y2"]])))
sx = (s2 - mean * s1) * rstd
inv_n = (1.0 / cnt)[:, None]
k1 = (g * rstd); k2 = -g * rstd * rstd * sx * inv_n; k3 = -g * rstd * s1 * inv_n - k2 * mean
dy = k1[vid] * dout + k2[vid] * y2r + k3[vid]
ref_dy = rows(torch.cat([v.grad for v in inter["y2"]]))
print("fp32 formula vs ref dy (unmasked)", rel(dy, ref_dy))
print("cnt", cnt, "pairs", plan.pairs_per_video)
dyk = (k1[vid] * dcm.float() + k2[vid] * S["y2"].float() + k3[vid]) * (S["y2"].float() > 0)
print("kernel vs torch-with-rounded-inputs", rel(d2.float(), dyk))
for v in range(V):
    msk = vid == v
    print("video", v, "kernel vs ref", rel(d2.float()[msk], ref_d2[msk]), "formula vs ref", rel((dy * (y2r > 0))[msk], ref_d2[msk]))
err = (d2.float() - ref_d2).abs()
i = err.argmax(); r, c = divmod(i.item(), 256)
print("worst", r, c, d2[r, c].item(), ref_d2[r, c].item(), "y2 ours", S["y2"][r, c].item(), "ref", y2r[r, c].item(), "k1", k1[vid[r], c].item(), "dout", dout[r, c].item())
mism = ((S["y2"].float() > 0) != (y2r > 0)).float().mean().item()
print("relu-mask mismatch fraction", mism)
