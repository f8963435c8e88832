"""First-contact GPU diagnostics: staged checks with lots of output (not a test, not a bench)."""
import os
import subprocess
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np
import torch


def stage(name):
    print("\n=== %s ===" % name, flush=True)


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main(which):
    import scn_oracle as O
    from scn_oracle import rules as R
    from sparse_rcnn_b200 import scn, _lib
    from sparse_rcnn_b200.synthetic import make_batch
    from tests.util import make_pair, random_scene, rel_err, copy_params
    dev = torch.device("cuda:0")
    print(torch.cuda.get_device_name(0), "sm_count", _lib.raw("scn_device_sm_count")(), flush=True)

    if which in ("all", "rules"):
        stage("rulebook small")
        coords, feats, size = random_scene(0)
        to, tg = make_pair(scn, coords, feats, size, dev)
        print("N oracle", to.features.shape, "gpu", tg.features.shape)
        print("locations equal", torch.equal(tg.get_spatial_locations(), to.get_spatial_locations()))
        print("feat err", rel_err(tg.features, to.features))
        ref = R.rules_to_map(to.metadata.subm_rules(size, 3), to.features.shape[0])
        got = tg.metadata.level(size).subm_map(3).cpu().numpy()
        print("subm map equal", np.array_equal(ref, got), (ref != got).sum())
        ok, rules, parent, off = to.metadata.conv_rules(size, 2, 2)
        r = tg.metadata.strided_rules(size, 2, 2)
        print("strided parent equal", np.array_equal(r.parent_row.cpu().numpy(), parent),
              "cmap equal", np.array_equal(r.cmap.cpu().numpy(), R.rules_to_map(rules, to.metadata.grids[ok].n)))

    if which in ("all", "fp32"):
        stage("fp32 conv small")
        scn.set_precision("fp32")
        for cin, cout in [(5, 7), (32, 32), (48, 80)]:
            coords, feats, size = random_scene(1, channels=cin)
            to, tg = make_pair(scn, coords, feats, size, dev)
            lo = O.SubmanifoldConvolution(3, cin, cout, 3, True)
            lo.bias.data.normal_()
            lg = copy_params(lo, scn.SubmanifoldConvolution(3, cin, cout, 3, True), dev)
            xo = to.features.clone().requires_grad_(True)
            xg = tg.features.clone().requires_grad_(True)
            yo = lo(O.SparseConvNetTensor(xo, to.metadata, size)).features
            yg = lg(scn.SparseConvNetTensor(xg, tg.metadata, size)).features
            g = torch.randn_like(yo)
            yo.backward(g), yg.backward(g.to(dev))
            print(cin, cout, "fwd", rel_err(yg, yo), "dx", rel_err(xg.grad, xo.grad), "dw",
                  rel_err(lg.weight.grad, lo.weight.grad), "db", rel_err(lg.bias.grad, lo.bias.grad), flush=True)

    if which in ("all", "tf32"):
        stage("tf32 tcgen05 conv small")
        scn.set_precision("tf32")
        for cin, cout in [(32, 32), (5, 7), (48, 80), (22, 22), (128, 128), (16, 256)]:
            coords, feats, size = random_scene(2, channels=cin)
            to, tg = make_pair(scn, coords, feats, size, dev)
            lo = O.SubmanifoldConvolution(3, cin, cout, 3, True)
            lo.bias.data.normal_()
            lg = copy_params(lo, scn.SubmanifoldConvolution(3, cin, cout, 3, True), dev)
            with torch.no_grad():
                yo = lo(to).features
                yg = lg(tg).features
            torch.cuda.synchronize()
            e = rel_err(yg, yo)
            print(cin, cout, "n", yo.shape[0], "fwd rel err", e, flush=True)
            if e > 5e-3:
                d = (yg.cpu() - yo).abs()
                print("  worst rows", d.max(1).values.topk(5).indices.tolist(), "worst cols",
                      d.max(0).values.topk(min(5, cout)).indices.tolist())
                print("  yo[0,:8]", yo[0, :8].tolist(), "\n  yg[0,:8]", yg[0, :8].cpu().tolist())

    if which in ("all", "perf"):
        stage("full-size single layer timing")
        coords, feats, size, bs, _ = make_batch(1, 0)
        md = scn.Metadata(3)
        t0 = time.time()
        f6 = scn.ioLayers.InputLayerFunction.apply(3, md, size, coords, feats.to(dev), bs, 4)
        torch.cuda.synchronize()
        print("input layer (first call) %.1f ms" % ((time.time() - t0) * 1e3), "N", f6.shape[0], "P", len(coords))
        lvl = md.level(size)
        t0 = time.time()
        m = lvl.subm_map(3)
        torch.cuda.synchronize()
        print("subm map %.2f ms" % ((time.time() - t0) * 1e3), "k_mean", float((m >= 0).sum()) / lvl.n)
        n = lvl.n
        for c in (32, 64, 128):
            conv = scn.SubmanifoldConvolution(3, c, c, 3, True).to(dev)
            x = scn.SparseConvNetTensor(torch.randn(n, c, device=dev), md, size)
            with torch.no_grad():
                for prec in ("fp32", "tf32"):
                    scn.set_precision(prec)
                    ms = timed(lambda: conv(x))
                    pairs = float((m >= 0).sum())
                    print("C=%d %s: %.3f ms  %.1f TFLOP/s (useful)" % (c, prec, ms, 2 * pairs * c * c / ms / 1e9), flush=True)
                scn.set_precision("fp32")
                ref = conv(x).features
                scn.set_precision("tf32")
                got = conv(x).features
                print("   tf32 vs fp32 rel err", rel_err(got, ref))


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "driver"
    if which == "driver":
        for w, to in (("rules", 300), ("fp32", 300), ("tf32", 300), ("perf", 300)):
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), w], timeout=to)
                print("[stage %s exit %d]" % (w, r.returncode), flush=True)
            except subprocess.TimeoutExpired:
                print("[stage %s TIMED OUT]" % w, flush=True)
    else:
        try:
            main(which)
        except Exception:
            traceback.print_exc()
            sys.exit(1)
