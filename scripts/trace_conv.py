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
a = a[a[:, 0] > 0]
print("units traced", len(a), "debug", os.environ.get("SCN_CONV_DEBUG", "0"))
d = a[8:]
print("producer (thread 0 of warp 4), mean cycles per unit:")
print("  wait empty        %7.0f" % (d[:,1]-d[:,0]).mean())
print("  weights issue     %7.0f" % (d[:,2]-d[:,1]).mean())
print("  8x LDGSTS issue   %7.0f" % (d[:,3]-d[:,2]).mean())
print("  arrive            %7.0f" % (d[:,4]-d[:,3]).mean())
print("  loop tail -> next %7.0f" % (d[1:,0]-d[:-1,4]).mean())
print("  unit period       %7.0f" % np.diff(d[:,0]).mean())
print("MMA thread: wait full %7.0f, period %7.0f" % ((d[:,7]-d[:,6]).mean(), np.diff(d[:,7]).mean()))
for i in range(20, 32):
    r = a[i]; print(i, [int(x - a[20,0]) for x in r])
