"""GPU time of the two tensor-core kernels at every level of the backbone pyramid, launched back to back through the raw C
ABI (ctypes, ~3 us of host time per launch) so that short kernels are not host bound.  usage: time_raw.py [levels, e.g. 0,2]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sparse_rcnn_b200 import scn, _lib
from sparse_rcnn_b200.scn import functions as Fn
from sparse_rcnn_b200.scn.metadata import _stream
from sparse_rcnn_b200.synthetic import make_batch
dev = torch.device("cuda:0"); scn.set_precision("tf32")
coords, feats, size, bs, _ = make_batch(1, 0)
md = scn.Metadata(3)
f = scn.ioLayers.InputLayerFunction.apply(3, md, size, coords, feats.to(dev), bs, 4)
cur = scn.SparseConvNetTensor(f, md, size)
chans = [32, 48, 64, 80, 96, 112]
want = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else list(range(6))
P = lambda t: t.data_ptr()
def timed(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
tot_f = tot_w = 0.0
for li, C in enumerate(chans):
    if li > 0:
        down = scn.Convolution(3, cur.features.shape[1], C, 2, 2, True).to(dev)
        with torch.no_grad():
            cur = down(cur)
    lvl = cur.metadata.level(cur.spatial_size)
    n = lvl.n
    if li in want:
        fmap = lvl.subm_map(3)
        w = torch.randn(27, C, C, device=dev) * 0.05
        img = torch.empty(_lib.LIB.load().scn_conv_weight_image_bytes(27, C, C), dtype=torch.uint8, device=dev)
        s = _stream()
        _lib.call("scn_conv_pack_weights", P(w), 27, C, C, 0, 0, P(img), s)
        x = Fn.tf32_exact(torch.randn(n, C, device=dev)); go = Fn.tf32_exact(torch.randn(n, C, device=dev))
        out = torch.empty(n, C, device=dev); gw = torch.zeros(27, C, C, device=dev); gb = torch.zeros(C, device=dev)
        tf = timed(lambda: _lib.call("scn_conv_fwd_tf32", P(x), C, C, n, P(fmap), n, 27, P(img), None, None, 0, None, 0,
                                     P(out), C, C, 0, s))
        tw = timed(lambda: _lib.call("scn_conv_bwd_weight", P(x), C, C, P(fmap), n, 27, P(go), C, C, P(gw), P(gb), 1, s))
        tot_f += tf; tot_w += tw
        print("L%d N=%d C=%d tiles=%d: fwd %.1f us  wgrad %.1f us" % (li, n, C, (n + 127) // 128, tf, tw), flush=True)
    cur = scn.SparseConvNetTensor(torch.randn(n, C, device=dev), md, cur.spatial_size)
print("sum: fwd %.1f us  wgrad %.1f us  (lib=%s ctas=%s)" % (tot_f, tot_w, os.path.basename(_lib.LIB_PATH), os.environ.get("SCN_CONV_CTAS", "-")))
