"""Where the distribution error of the CUDA path comes from, at the headline shape (one 32-frame video per run):
(a) end-to-end error vs the oracle, (b) error of the ORACLE heads applied to the CUDA path's features (upstream
error only), (c) error of the CUDA heads applied to the oracle's features (head GEMM + epilogue only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200vsgg import synthetic, tempura
from oracle.tempura_oracle import TempuraOracle

KW = dict(mode="predcls", attention_class_num=3, spatial_class_num=6, contact_class_num=17, enc_layer_num=1,
          dec_layer_num=3, obj_mem_compute=False, rel_mem_compute="joint", mem_fusion="late", selection="manual",
          selection_lambda=0.5, take_obj_mem_feat=False, obj_head="gmm", rel_head="gmm", K=6, tracking=False)
classes = synthetic.ag_object_classes()
m = tempura.TEMPURA(obj_classes=classes, **KW)
synthetic.seeded_init_(m, 11)
o = TempuraOracle(obj_classes=classes, dropout=0.0, **KW)
o.load_state_dict(m.state_dict(), strict=True)
m = m.cuda().eval()
o.eval()
keys = ("attention_distribution", "spatial_distribution", "contacting_distribution")
worst = {}
for v in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
    e = synthetic.make_video_entry(v, 32, (6, 10))
    with torch.no_grad():
        ref = o(dict(e), phase="test")
        got = m({k: (t.cuda() if isinstance(t, torch.Tensor) else t) for k, t in e.items()}, phase="test")
        feat_c = got["rel_mem_features"].float().cpu()
        feat_o = ref["global_output"]
        heads = (o.a_rel_compress, o.s_rel_compress, o.c_rel_compress)
        up = [h(feat_c, "test") for h in heads]
        res = tempura.apply_heads([m.a_rel_compress, m.s_rel_compress, m.c_rel_compress], feat_o.cuda(), 0, [None] * 3, 0)
    fe = (feat_c - feat_o).abs()
    print("video %d: feature err max %.3e (rel to max|ref| %.2e, rel-L2 %.2e)" % (
        v, fe.max().item(), fe.max().item() / feat_o.abs().max().item(), fe.norm().item() / feat_o.norm().item()))
    for i, k in enumerate(keys):
        worst[k] = max(worst.get(k, 0.0), (got[k].float().cpu() - ref[k]).abs().max().item())
        print("   %-24s end-to-end %.2e | upstream only %.2e | head only %.2e" % (
            k, (got[k].float().cpu() - ref[k]).abs().max().item(), (up[i] - ref[k]).abs().max().item(),
            (res[i].float().cpu() - ref[k]).abs().max().item()))
print("worst end-to-end over the videos:", worst)
