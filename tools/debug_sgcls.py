"""Where does the object-branch backward lose accuracy?  Compares intermediate gradients with the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_sgcls_gpu import _setup, _clone
from b200vsgg import synthetic, tempura
from oracle.tempura_oracle import object_loss

name = sys.argv[1] if len(sys.argv) > 1 else "sgcls_track_gmm"
gold, entry, m, o = _setup(name)
eps = gold["eps"]
oc = o.object_classifier
keep = {}
def hook(tag):
    def f(mod, inp, out):
        out.retain_grad(); keep[tag] = out
    return f
oc.intermediate.register_forward_hook(hook("y"))
oc.intermediate[0].register_forward_hook(hook("z"))
po = o(_clone(entry), phase="train", eps=eps)
po["object_features"].retain_grad()
lo = object_loss(po, 0.5)
lo.backward()
m.dropout_p = 0.0; m.object_classifier.dropout_p = 0.0; m.gmm_eps = eps
m.object_classifier._debug = True
pm = m(_clone(entry, "cuda"), phase="train")
dbg = m.object_classifier._debug_last
dbg["y"].retain_grad(); pm["object_features"].retain_grad()
lm = tempura.tempura_loss(pm, m.last_plan, eos_coef=0.5)["object_loss"]
lm.backward()
def rel(a, b):
    return ((a.float().cpu() - b).norm() / b.norm()).item()
print("loss", lm.item(), lo.item())
print("dist fwd rel", rel(pm["distribution"], po["distribution"].detach()))
print("y fwd rel", rel(dbg["y"], keep["y"].detach()), "gate flips", ((dbg["y"].float().cpu() > 0) != (keep["y"] > 0)).float().mean().item())
print("objfeat fwd rel", rel(pm["object_features"], po["object_features"].detach()))
print("dy rel", rel(dbg["y"].grad, keep["y"].grad))
print("d objfeat rel", rel(pm["object_features"].grad, po["object_features"].grad))
og = dict(o.named_parameters())
for pname, p in m.named_parameters():
    if pname.startswith("object_classifier.") and og[pname].grad is not None and p.grad is not None and og[pname].grad.norm() > 1e-7:
        print("  %-70s %.4f" % (pname, rel(p.grad, og[pname].grad)))
