"""Single-layer driver for ncu: L0-sized SubM 3^3 layer, forward (k_conv_tc) and weight gradient
(k_conv_wgrad_tc).  Usage: python scripts/profile_conv.py [C] [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sparse_rcnn_b200 import scn
from sparse_rcnn_b200.scn import functions as Fn
from sparse_rcnn_b200.synthetic import make_batch

C = int(sys.argv[1]) if len(sys.argv) > 1 else 32
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
scn.set_precision("tf32")
coords, feats, size, bs, _ = make_batch(1, 0)
md = scn.Metadata(3)
scn.ioLayers.InputLayerFunction.apply(3, md, size, coords, feats.to(dev), bs, 4)
n = md.level(size).n
conv = scn.SubmanifoldConvolution(3, C, C, 3, True).to(dev)
x = Fn.tf32_exact(torch.randn(n, C, device=dev)).requires_grad_(True)
x._scn_tf32 = True
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(iters):
    y = conv(scn.SparseConvNetTensor(x, md, size)).features
    g = torch.randn_like(y)
    torch.cuda.synchronize()
    e0.record()
    y.backward(g)
    e1.record()
    torch.cuda.synchronize()
print("N", n, "C", C, "bwd (dgrad+wgrad+bias) ms", e0.elapsed_time(e1))
