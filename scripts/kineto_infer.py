"""GPU timeline of the sparse inference pass (backbone + seg + class net + mask net on 256 boxes)."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from sparse_rcnn_b200 import pipeline, scn
from sparse_rcnn_b200.synthetic import make_boxes
dev = torch.device("cuda:0"); scn.set_precision("tf32")
inf = pipeline.SparseInference(dev)
data, _ = bench.make_inputs(0)
boxes = make_boxes(data[0], 256, 7)
pdata = (data[0].pin_memory(), data[1].pin_memory(), data[2], data[3], data[4])
for _ in range(4): inf(pdata, boxes)
torch.cuda.synchronize()
NS = 5
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(NS): inf(pdata, boxes)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ks = sorted((e.time_range.start, e.time_range.end, e.name) for e in evs if e.time_range.end > e.time_range.start)
t0, t1 = ks[0][0], max(k[1] for k in ks)
busy, cs, ce = 0, ks[0][0], ks[0][1]
for s, e, _ in ks[1:]:
    if s > ce: busy += ce - cs; cs, ce = s, e
    else: ce = max(ce, e)
busy += ce - cs
print("span %.2f ms/scene, GPU busy %.2f ms/scene (%.0f%%), kernels/scene %d" % ((t1 - t0) / NS / 1e3, busy / NS / 1e3, 100.0 * busy / (t1 - t0), len(ks) // NS))
agg = collections.defaultdict(lambda: [0, 0.0])
for s, e, n in ks:
    n = n.split("(")[0][:60]; agg[n][0] += 1; agg[n][1] += (e - s)
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:14]:
    print("%-62s n/scene=%5.1f  %7.3f ms/scene  avg %6.1f us" % (n, c / NS, t / NS / 1e3, t / c))
gaps = collections.defaultdict(lambda: [0, 0.0]); end = ks[0][1]
for s, e, n in ks[1:]:
    if s - end > 5:
        nm = n.split("(")[0][:50]; gaps[nm][0] += 1; gaps[nm][1] += s - end
    end = max(end, e)
print("idle gaps > 5 us by following kernel:")
for n, (c, t) in sorted(gaps.items(), key=lambda x: -x[1][1])[:10]:
    print("  %-52s n/scene=%5.1f  %7.3f ms/scene" % (n, c / NS, t / NS / 1e3))
