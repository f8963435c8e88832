"""Timeline of CTA 0 of one small-level k_conv_tc launch (needs the `trace` variant library: SCN_EXP_TRACE)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sparse_rcnn_b200 import scn, _lib
from sparse_rcnn_b200.scn import functions as Fn
from sparse_rcnn_b200.scn.metadata import _stream
from sparse_rcnn_b200.synthetic import make_batch
dev = torch.device("cuda:0"); scn.set_precision("tf32")
level = int(sys.argv[1]) if len(sys.argv) > 1 else 5
coords, feats, size, bs, _ = make_batch(1, 0)
md = scn.Metadata(3)
f = scn.ioLayers.InputLayerFunction.apply(3, md, size, coords, feats.to(dev), bs, 4)
cur = scn.SparseConvNetTensor(f, md, size)
chans = [32, 48, 64, 80, 96, 112]
for li, C in enumerate(chans[:level + 1]):
    if li > 0:
        down = scn.Convolution(3, cur.features.shape[1], C, 2, 2, True).to(dev)
        with torch.no_grad(): cur = down(cur)
    n = cur.metadata.level(cur.spatial_size).n
    cur = scn.SparseConvNetTensor(torch.randn(n, C, device=dev), md, cur.spatial_size)
lvl = cur.metadata.level(cur.spatial_size); n = lvl.n; C = chans[level]
fmap = lvl.subm_map(3); P = lambda t: t.data_ptr()
w = torch.randn(27, C, C, device=dev) * 0.05
img = torch.empty(_lib.LIB.load().scn_conv_weight_image_bytes(27, C, C), dtype=torch.uint8, device=dev)
s = _stream()
_lib.call("scn_conv_pack_weights", P(w), 27, C, C, 0, 0, P(img), s)
x = Fn.tf32_exact(torch.randn(n, C, device=dev)); out = torch.empty(n, C, device=dev)
names = ["entry", "setup done", "idx loaded", "first copies issued", "first stage full", "tile committed", "accf seen",
         "reds issued", "fence done", "ticket done", "final done", "exit"]
buf = (ctypes.c_ulonglong * 32)()
cta = (ctypes.c_ulonglong * 1024)()
for rep in range(4):
    _lib.call("scn_conv_fwd_tf32", P(x), C, C, n, P(fmap), n, 27, P(img), None, None, 0, None, 0, P(out), C, C, 0, s)
    _lib.LIB.load().scn_debug_read_trace(buf)
    t0 = buf[0]
    print("L%d N=%d C=%d rep %d: " % (level, n, C, rep) + " | ".join("%s %.1f" % (nm, (buf[i] - t0) / 1e3) for i, nm in enumerate(names)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.call("scn_conv_fwd_tf32", P(x), C, C, n, P(fmap), n, 27, P(img), None, None, 0, None, 0, P(out), C, C, 0, s)
    e1.record(); torch.cuda.synchronize()
    _lib.LIB.load().scn_debug_read_cta_times(cta)
    last = max(cta[2 * i] for i in range(512))
    live = [i for i in range(512) if cta[2 * i + 1] > cta[2 * i] > last - 100000]      # this launch only
    ent = [cta[2 * i] for i in live]
    ext = [cta[2 * i + 1] for i in live]
    lo = min(ent)
    print("  CTAs %d: entries %.1f..%.1f us, exits %.1f..%.1f us; event time %.1f us" % (
        len(ent), 0.0, (max(ent) - lo) / 1e3, (min(ext) - lo) / 1e3, (max(ext) - lo) / 1e3, e0.elapsed_time(e1) * 1e3))
