"""Timeline of CTA 0 of the tile-local kernel (variant built with -DSCN_TS_TRACE): per-unit stamps of the MMA warp and the
gather groups, per-tile stamps of loader / epilogue.  usage: SCN_B200_LIB=.../libscn_tstrace.so python scripts/trace_ts.py [C]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from sparse_rcnn_b200 import scn, _lib
from sparse_rcnn_b200.scn import functions as Fn
from sparse_rcnn_b200.scn.metadata import _stream
from sparse_rcnn_b200.synthetic import make_batch
dev = torch.device("cuda:0"); scn.set_precision("tf32")
C = int(sys.argv[1]) if len(sys.argv) > 1 else 32
coords, feats, size, bs, _ = make_batch(1, 0)
md = scn.Metadata(3)
scn.ioLayers.InputLayerFunction.apply(3, md, size, coords, feats.to(dev), bs, 4)
lvl = md.level(size); n = lvl.n; fmap = lvl.subm_map(3)
P = lambda t: t.data_ptr()
w = torch.randn(27, C, C, device=dev) * 0.05
img = torch.empty(_lib.LIB.load().scn_conv_weight_image_bytes(27, C, C), dtype=torch.uint8, device=dev)
s = _stream()
_lib.call("scn_conv_pack_weights", P(w), 27, C, C, 0, 0, P(img), s)
x = Fn.tf32_exact(torch.randn(n, C, device=dev)); out = torch.empty(n, C, device=dev)
for _ in range(3):
    _lib.call("scn_conv_fwd_tf32", P(x), C, C, n, P(fmap), n, 27, P(img), None, None, 0, None, 0, P(out), C, C, 0, s)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 16384)()
dll = ctypes.CDLL(_lib.LIB_PATH)
assert dll.scn_debug_ts_trace(buf) == 0
t = np.array(buf[:], dtype=np.int64)
t0 = t[16000]
mma = t[:4096].reshape(-1, 4)[:, :3]
nu = int((mma[:, 0] > 0).sum())
mma = mma[:nu] - t0
ga = t[4096:4096 + 4 * nu].reshape(-1, 4) - t0
print("units of CTA 0:", nu, " total cycles (first MMA stamp -> last):", int(mma[-1, 2] - mma[0, 0]))
d = np.diff(mma[:, 0])
print("MMA loop period per unit: mean %.0f median %.0f p90 %.0f" % (d.mean(), np.median(d), np.percentile(d, 90)))
print("MMA wait for a_full: mean %.0f median %.0f ; issue+commit+step: mean %.0f" % (
    (mma[:, 1] - mma[:, 0]).mean(), np.median(mma[:, 1] - mma[:, 0]), (mma[:, 2] - mma[:, 1]).mean()))
ok = ga[:, 0] > -10**12
print("gather (q=0 warp of each group) per unit: wait a_empty %.0f | LDS+STTM issue %.0f | wait::st+arrive %.0f | total %.0f" % (
    (ga[ok, 1] - ga[ok, 0]).mean(), (ga[ok, 2] - ga[ok, 1]).mean(), (ga[ok, 3] - ga[ok, 2]).mean(), (ga[ok, 3] - ga[ok, 0]).mean()))
print("gather arrive -> MMA sees it (mma after-wait - gather arrive): mean %.0f median %.0f" % (
    (mma[:, 1] - ga[:nu, 3]).mean(), np.median(mma[:, 1] - ga[:nu, 3])))
for g in range(4):
    own = ga[g::4]
    dd = np.diff(own[:, 0])
    print(" group %d: period between own units: mean %.0f ; busy %.0f" % (g, dd.mean(), (own[:, 3] - own[:, 0]).mean()))
ld = t[12000:12000 + 4 * 12].reshape(-1, 4)[:, :3] - t0
ep = t[14000:14000 + 4 * 12].reshape(-1, 4)[:, :3] - t0
hf = t[13000:13000 + 2 * 12].reshape(-1, 2) - t0
print("loader per tile (start, after halo_empty wait, copies issued):"); print(ld[:10])
print("gather waits halo_full (before, after):"); print(hf[:10])
print("epilogue per tile (start wait, acc_full seen, done):"); print(ep[:10])
print("first 40 units: MMA (top, ready, done) | gather (start, a_empty ok, sttm issued, arrived)")
for i in range(40):
    print(i, mma[i], ga[i])
