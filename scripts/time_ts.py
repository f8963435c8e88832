"""Forward time of one SubM 3^3 layer on the bench scene at levels 0 / 1 under SCN_TS_DBG switches (tile-local kernel)."""
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
P = lambda t: t.data_ptr()
def timed(fn, n=40):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
lvl = md.level(size)
n = lvl.n
fmap = lvl.subm_map(3)
chans = [int(c) for c in sys.argv[1].split(",")] if len(sys.argv) > 1 else [32]
variants = sys.argv[2].split(",") if len(sys.argv) > 2 else ["0"]
for C in chans:
    w = torch.randn(27, C, C, device=dev) * 0.05
    img = torch.empty(_lib.LIB.load().scn_conv_weight_image_bytes(27, C, C), dtype=torch.uint8, device=dev)
    s = _stream()
    _lib.call("scn_conv_pack_weights", P(w), 27, C, C, 0, 0, P(img), s)
    x = Fn.tf32_exact(torch.randn(n, C, device=dev))
    out = torch.empty(n, C, device=dev)
    res = []
    for v in variants:
        os.environ["SCN_TS_DBG"] = v
        t = timed(lambda: _lib.call("scn_conv_fwd_tf32", P(x), C, C, n, P(fmap), n, 27, P(img), None, None, 0, None, 0, P(out), C, C, 0, s))
        res.append("dbg=%s %.1f us" % (v, t))
    print("N=%d C=%d: " % (n, C) + " | ".join(res), flush=True)
