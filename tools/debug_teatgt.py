import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200vsgg import synthetic, teatgt
from oracle import teatgt_oracle as TO
gold = torch.load("tests/golden/teatgt_small.pt", weights_only=False)
args = types.SimpleNamespace(**gold["args"])
classes = synthetic.ag_object_classes()
m = teatgt.TEAT_GT(obj_classes=classes, args=args, **gold["model_kw"])
synthetic.teatgt_seeded_init_(m, gold["seed"])
o = TO.TeatgtOracle(obj_classes=classes, args=args, with_regulariser=True, **gold["model_kw"])
o.load_state_dict(m.state_dict(), strict=True)
m = m.cuda().eval(); o.eval()
e = synthetic.make_video_entry(**gold["case"]); e.pop("union_feat"); e.pop("spatial_masks")
# oracle intermediates
cap = {"x0": [], "xl": [], "logits": []}
enc = o.TokenGT_encoder
orig_tokens, orig_encode, orig_head = enc.tokens, enc.encode, enc.head
def tokens(*a, **k):
    x = orig_tokens(*a, **k); cap["x0"].append(x); return x
def encode(x):
    x = orig_encode(x); cap["xl"].append(x); return x
def head(x):
    l, h = orig_head(x); cap["logits"].append(l); return l, h
enc.tokens, enc.encode, enc.head = tokens, encode, head
with torch.no_grad():
    ro = o(dict(e), phase="test")
    lay = TO.node_layout({**e, "pred_labels": e["labels"]})
    tok_o = o.node_tokens({**e, "pred_labels": e["labels"]}, lay)
m._debug = {}
with torch.no_grad():
    rm = m({k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in e.items()}, phase="test")
d = m._debug
def rel(a, b): return ((a.float().cpu() - b).norm() / b.norm()).item()
print("tok", rel(d["tok"], tok_o))
print("x0", rel(d["x0"], torch.cat(cap["x0"])))
x0o = torch.cat(cap["x0"]); x0m = d["x0"].float().cpu()
kinds = m.last_plan.desc_h[:, 0]
for kd in range(4):
    sel = torch.from_numpy(kinds == kd)
    print(" kind", kd, int(sel.sum()), rel(x0m[sel], x0o[sel]))
print("x_final", rel(d["layers"][-1], torch.cat(cap["xl"])))
print("logits", rel(d["logits"][:, :26], torch.cat(cap["logits"])))
print("dist", (rm["attention_distribution"].cpu() - ro["attention_distribution"]).abs().max().item())
lo = torch.cat(cap["logits"]); lm = d["logits"][:, :26].float().cpu()
err = (lm - lo).abs()
print("logits std", lo.std().item(), "max", lo.abs().max().item(), "max abs err", err.max().item(), "row errs", err.max(1).values)
xo = torch.cat(cap["xl"]); xm = d["layers"][-1].float().cpu()
print("x_final max abs", xo.abs().max().item(), "err max", (xm - xo).abs().max().item())
for i, xl in enumerate(d["layers"]):
    print("layer", i, "norm", xl.float().norm().item())
