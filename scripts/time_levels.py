"""Time one SubM 3^3 conv (forward, tf32) at every level of the backbone pyramid, warm, CUDA events."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sparse_rcnn_b200 import scn
from sparse_rcnn_b200.scn import functions as Fn
from sparse_rcnn_b200.synthetic import make_batch
dev = torch.device("cuda:0"); scn.set_precision("tf32")
coords, feats, size, bs, _ = make_batch(1, 0)
md = scn.Metadata(3)
f = scn.ioLayers.InputLayerFunction.apply(3, md, size, coords, feats.to(dev), bs, 4)
t = scn.SparseConvNetTensor(f, md, size)
chans = [32, 48, 64, 80, 96, 112]
cur = t
out = []
for li, C in enumerate(chans):
    if li > 0:
        down = scn.Convolution(3, cur.features.shape[1], C, 2, 2, True).to(dev)
        with torch.no_grad():
            cur = down(cur)
    n = cur.metadata.level(cur.spatial_size).n
    conv = scn.SubmanifoldConvolution(3, C, C, 3, True).to(dev)
    x = scn.SparseConvNetTensor(Fn.tf32_exact(torch.randn(n, C, device=dev)), md, cur.spatial_size)
    with torch.no_grad():
        for _ in range(5): conv(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): conv(x)
        e1.record(); torch.cuda.synchronize()
    out.append("L%d N=%d C=%d tiles=%d: %.1f us" % (li, n, C, (n + 127) // 128, e0.elapsed_time(e1) / 50 * 1e3))
    cur = scn.SparseConvNetTensor(torch.randn(n, C, device=dev), md, cur.spatial_size)
print("NOSPLIT=%s | " % os.environ.get("SCN_CONV_NOSPLIT", "0") + " | ".join(out))
