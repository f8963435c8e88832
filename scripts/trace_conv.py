import os, sys, ctypes
os.environ["SCN_CONV_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from sparse_rcnn_b200 import scn, _lib
from sparse_rcnn_b200.scn import functions as Fn
from sparse_rcnn_b200.synthetic import make_batch
C = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0"); scn.set_precision("tf32")
coords, feats, size, bs, _ = make_batch(1, 0)
md = scn.Metadata(3); scn.ioLayers.InputLayerFunction.apply(3, md, size, coords, feats.to(dev), bs, 4)
n = md.level(size).n
conv = scn.SubmanifoldConvolution(3, C, C, 3, True).to(dev)
t = scn.SparseConvNetTensor(Fn.tf32_exact(torch.randn(n, C, device=dev)), md, size)
with torch.no_grad():
    for _ in range(3): conv(t)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 1024)()
dll = ctypes.CDLL(_lib.LIB_PATH)
print("rc", dll.scn_debug_trace(buf))
a = np.array(buf[:]).reshape(256, 4)
a = a[a[:, 0] > 0]
t0 = a[0, 0]
print("units traced", len(a))
print("unit | prod: wait_start wait_end (wait) | mma: wait_start wait_end (wait) | prod per-unit period | mma period")
for i in range(min(len(a), 140)):
    if i < 40 or i % 10 == 0:
        print("%4d | %8d %8d (%6d) | %8d %8d (%6d) | %6d | %6d" % (i, a[i,0]-t0, a[i,1]-t0, a[i,1]-a[i,0], a[i,2]-t0, a[i,3]-t0, a[i,3]-a[i,2],
              a[i,0]-a[i-1,0] if i else 0, a[i,3]-a[i-1,3] if i else 0))
pw = (a[:,1]-a[:,0]); mw = (a[:,3]-a[:,2])
print("mean producer wait-for-empty %.0f cyc, mean MMA wait-for-full %.0f cyc, mean unit period %.0f cyc" % (pw[8:].mean(), mw[8:].mean(), np.diff(a[8:,0]).mean()))
