import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200vsgg import synthetic, tempura, ops

kw = dict(mode="predcls", attention_class_num=3, spatial_class_num=6, contact_class_num=17,
          enc_layer_num=1, dec_layer_num=3, obj_mem_compute=False, rel_mem_compute="joint", mem_fusion="late",
          selection="manual", selection_lambda=0.5, obj_head="gmm", rel_head="gmm", K=6, tracking=False)
m = tempura.TEMPURA(obj_classes=synthetic.ag_object_classes(), **kw)
synthetic.seeded_init_(m)
m = m.cuda().eval()
e = synthetic.make_video_entry(3, 6, (3, 5), device="cuda")
with torch.no_grad():
    out = m(dict(e), phase="test")
    torch.cuda.synchronize()
    print("forward 1 ok")
    out = m(dict(e), phase="test", unc=True)
    torch.cuda.synchronize()
    print("forward unc ok")
