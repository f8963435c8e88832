"""BASELINE config 3: single SubM 3^3 layer sweep, C in {16..256} x N in {50k..2M}; TF32 tcgen05 vs exact fp32.
Prints a markdown table (ms, useful TFLOP/s, algorithmic GB/s, rel. error of TF32 vs fp32)."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sparse_rcnn_b200 import scn
from sparse_rcnn_b200.scn import functions as Fn
from sparse_rcnn_b200.synthetic import make_batch

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n


print("| N active | k_mean | C | tf32 ms | fp32 ms | tf32 useful TFLOP/s | tf32 alg. GB/s | frac of 6551.7 | rel err tf32 vs fp32 |")
print("|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for target in (50_000, 150_000, 500_000, 2_000_000):
    scale = math.sqrt(target / 167_000.0)
    side = int(math.ceil(256 * max(scale, 0.6) / 32) * 32)
    size = (side, side, max(side // 2, 32))
    coords, feats, sz, bs, _ = make_batch(1, 0, spatial_size=size, scale=scale, room_offset=(16, 16, 4))
    md = scn.Metadata(3)
    scn.ioLayers.InputLayerFunction.apply(3, md, sz, coords, feats.to(dev), bs, 4)
    lvl = md.level(sz)
    n = lvl.n
    m = lvl.subm_map(3)
    pairs = float((m >= 0).sum())
    for C in (16, 32, 64, 128, 256):
        torch.manual_seed(0)
        conv = scn.SubmanifoldConvolution(3, C, C, 3, True).to(dev)
        x = Fn.tf32_exact(torch.randn(n, C, device=dev))
        t = scn.SparseConvNetTensor(x, md, sz)
        with torch.no_grad():
            scn.set_precision("tf32")
            ms_t = timed(lambda: conv(t))
            yt = conv(t).features
            scn.set_precision("fp32")
            ms_f = timed(lambda: conv(t), n=3)
            yf = conv(t).features
            err = float((yt - yf).abs().max() / yf.abs().max())
        alg = 2 * n * C * 4 + 27 * C * C * 4 + 4 * 27 * n
        print("| %d | %.1f | %d | %.3f | %.3f | %.1f | %.0f | %.3f | %.1e |" % (
            n, pairs / n, C, ms_t, ms_f, 2 * pairs * C * C / ms_t / 1e9, alg / ms_t / 1e6, alg / ms_t / 1e6 / 6551.7, err), flush=True)
    del md, lvl, m
    torch.cuda.empty_cache()
