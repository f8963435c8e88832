"""Time one SubM 3^3 layer forward (k_conv_tc) with CUDA events; prints ms.  usage: time_conv.py C [N_scale]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sparse_rcnn_b200 import scn
from sparse_rcnn_b200.scn import functions as Fn
from sparse_rcnn_b200.synthetic import make_batch
C = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
scn.set_precision("tf32")
coords, feats, size, bs, _ = make_batch(1, 0)
md = scn.Metadata(3)
scn.ioLayers.InputLayerFunction.apply(3, md, size, coords, feats.to(dev), bs, 4)
n = md.level(size).n
conv = scn.SubmanifoldConvolution(3, C, C, 3, True).to(dev)
x = Fn.tf32_exact(torch.randn(n, C, device=dev))
t = scn.SparseConvNetTensor(x, md, size)
with torch.no_grad():
    for _ in range(5): conv(t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): conv(t)
    e1.record(); torch.cuda.synchronize()
print("C=%d debug=%s: %.1f us/layer (warm L2)" % (C, os.environ.get("SCN_CONV_DEBUG", "0"), e0.elapsed_time(e1) / 20 * 1e3))
