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
nb = int((mma[:, 0] > 0).sum())
mma = mma[:nb] - t0
ga = t[4096:4096 + 16 * nb].reshape(nb, 4, 4) - t0      # [batch, group, stamp]
print("batches of CTA 0:", nb, " cycles first->last MMA stamp:", int(mma[-1, 2] - mma[0, 0]))
d = np.diff(mma[:, 0])
print("MMA period per batch: mean %.0f median %.0f p90 %.0f | wait b_full mean %.0f median %.0f | issue+commit mean %.0f" % (
    d.mean(), np.median(d), np.percentile(d, 90), (mma[:, 1] - mma[:, 0]).mean(), np.median(mma[:, 1] - mma[:, 0]),
    (mma[:, 2] - mma[:, 1]).mean()))
ok = ga[:, :, 0] > -10**12
for g in range(4):
    x = ga[:, g][ok[:, g]]
    print(" group %d (units %d): wait b_empty %.0f | LDS+STTM issue %.0f | wait::st+arrive %.0f | period %.0f" % (
        g, len(x), (x[:, 1] - x[:, 0]).mean(), (x[:, 2] - x[:, 1]).mean(), (x[:, 3] - x[:, 2]).mean(), np.diff(x[:, 0]).mean()))
last = np.where(ok, ga[:, :, 3], -10**15).max(1)
print("last gather arrival of a batch -> MMA past its wait: mean %.0f median %.0f" % ((mma[:, 1] - last).mean(), np.median(mma[:, 1] - last)))
ld = t[12000:12000 + 4 * 10].reshape(-1, 4)[:, :3] - t0
ep = t[14000:14000 + 4 * 10].reshape(-1, 4)[:, :3] - t0
hf = t[13000:13000 + 2 * 10].reshape(-1, 2) - t0
print("loader per tile (start, after halo_empty wait, copies issued):"); print(ld[:9])
print("gather g0 waits halo_full (before, after):"); print(hf[:9])
print("epilogue per tile (start wait, acc_full seen, done):"); print(ep[:9])
print("first 24 batches: MMA (top, ready, done) | last gather arrival | group0 (start, b_empty ok, sttm, arrived)")
for i in range(24):
    print(i, mma[i], int(last[i]), ga[i, 0])
