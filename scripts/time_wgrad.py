"""Weight-gradient time of one SubM 3^3 layer on the bench scene at level 0: tile-local kernel vs conv_wgrad_tc.cu."""
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
scn.ioLayers.InputLayerFunction.apply(3, md, size, coords, feats.to(dev), bs, 4)
P = lambda t: t.data_ptr()
lvl = md.level(size); n = lvl.n; m = lvl.subm_map(3); s = _stream()
for C in [int(c) for c in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["32", "48"])]:
    x = Fn.tf32_exact(torch.randn(n, C, device=dev)); go = Fn.tf32_exact(torch.randn(n, C, device=dev))
    gw = torch.zeros(27, C, C, device=dev); gb = torch.zeros(C, device=dev)
    res = []
    for ts in ("1", "0"):
        os.environ["SCN_WGRAD_TS"] = ts
        fn = lambda: _lib.call("scn_conv_bwd_weight", P(x), C, C, P(m), n, 27, P(go), C, C, P(gw), P(gb), 1, s)
        for _ in range(5): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(40): fn()
        e1.record(); torch.cuda.synchronize()
        res.append("ts=%s %.1f us" % (ts, e0.elapsed_time(e1) / 40 * 1e3))
    print("N=%d C=%d: " % (n, C) + " | ".join(res), flush=True)
