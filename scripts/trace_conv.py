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
buf = (ctypes.c_longlong * 2048)()
dll = ctypes.CDLL(_lib.LIB_PATH)
print("rc", dll.scn_debug_trace(buf))
a = np.array(buf[:]).reshape(256, 8)
nz = int((a[:, 7] > 0).sum())
a = a[:nz]
t0 = a[a[:, 0] > 0][0, 0]
print("units with MMA trace", nz)
print("unit | producer(warp4 only: every 4th): t_start wait issueW ldgsts arrive | MMA: wait_start wait_end(+wait) done(+busy)")
for i in range(16, 60):
    r = a[i]
    if r[0] > 0:
        ps = "%7d w%5d W%5d L%5d a%4d" % (r[0]-t0, r[1]-r[0], r[2]-r[1], r[3]-r[2], r[4]-r[3])
    else:
        ps = " " * 36
    print("%3d | %s | %7d %7d (+%5d) %7d (+%5d)" % (i, ps, r[6]-t0, r[7]-t0, r[7]-r[6], r[5]-t0, r[5]-r[7]))
m = a[16:]
print("MMA: mean wait %.0f, mean busy %.0f, period %.0f" % ((m[:,7]-m[:,6]).mean(), (m[:,5]-m[:,7]).mean(), np.diff(m[:,7]).mean()))
pr = m[m[:,0] > 0]
print("producer warp 4: mean wait %.0f, weights %.0f, ldgsts %.0f, arrive %.0f, period(4 units) %.0f" % (
    (pr[:,1]-pr[:,0]).mean(), (pr[:,2]-pr[:,1]).mean(), (pr[:,3]-pr[:,2]).mean(), (pr[:,4]-pr[:,3]).mean(), np.diff(pr[:,0]).mean()))
