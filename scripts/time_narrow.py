import os, sys
sys.path.insert(0, "/root/repo")
import torch
from sparse_rcnn_b200 import scn
from sparse_rcnn_b200.scn import functions as Fn
from sparse_rcnn_b200.synthetic import make_batch
dev = torch.device("cuda:0"); scn.set_precision("tf32")
coords, feats, sz, bs, _ = make_batch(1, 0)
md = scn.Metadata(3)
scn.ioLayers.InputLayerFunction.apply(3, md, sz, coords, feats.to(dev), bs, 4)
n = md.level(sz).n
for C in (16, 22, 32, 48):
    conv = scn.SubmanifoldConvolution(3, C, C, 3, True).to(dev)
    t = scn.SparseConvNetTensor(Fn.tf32_exact(torch.randn(n, C, device=dev)), md, sz)
    with torch.no_grad():
        for _ in range(5): conv(t)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): conv(t)
        e1.record(); torch.cuda.synchronize()
    print("N=%d C=%d: %.1f us (skip=%s)" % (n, C, e0.elapsed_time(e1) / 50 * 1e3, os.environ.get("SCN_CONV_SKIP", "default")))
