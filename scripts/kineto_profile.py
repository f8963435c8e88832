"""In-situ GPU timeline of the training step via torch.profiler (CUPTI): GPU busy time vs wall time, per-kernel totals."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from sparse_rcnn_b200 import pipeline, scn
dev = torch.device("cuda:0"); scn.set_precision("tf32")
tr = pipeline.BackboneTrainer(dev)
data, labels = bench.make_inputs(0)
data = (data[0].to(dev), data[1].to(dev), data[2], data[3], data[4]); labels = labels.to(dev)
for _ in range(5): tr.step(data, labels)
torch.cuda.synchronize()
NS = 5
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(NS): tr.step(data, labels)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ks = [(e.time_range.start, e.time_range.end, e.name) for e in evs if e.time_range.end > e.time_range.start]
ks.sort()
t0, t1 = ks[0][0], max(k[1] for k in ks)
busy, cur_s, cur_e = 0, ks[0][0], ks[0][1]
for s, e, _ in ks[1:]:
    if s > cur_e:
        busy += cur_e - cur_s; cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
print("span %.2f ms/step, GPU busy %.2f ms/step (%.0f%%), kernels/step %d" % ((t1 - t0) / NS / 1e3, busy / NS / 1e3, 100.0 * busy / (t1 - t0), len(ks) // NS))
agg = collections.defaultdict(lambda: [0, 0.0])
for s, e, n in ks:
    n = n.split("(")[0][:60]; agg[n][0] += 1; agg[n][1] += (e - s)
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:16]:
    print("%-62s n/step=%5.1f  %7.3f ms/step  avg %6.1f us" % (n, c / NS, t / NS / 1e3, t / c))
# idle-gap analysis: where the GPU waits for the host (gap > 5 us), attributed to the kernel that follows the gap
gaps = collections.defaultdict(lambda: [0, 0.0])
end = ks[0][1]
big = []
for s, e, n in ks[1:]:
    if s - end > 5:
        nm = n.split("(")[0][:50]
        gaps[nm][0] += 1; gaps[nm][1] += s - end
        big.append((s - end, (s - t0) / 1e3, nm))
    end = max(end, e)
print("idle gaps > 5 us by following kernel (per step):")
for n, (c, t) in sorted(gaps.items(), key=lambda x: -x[1][1])[:14]:
    print("  %-52s n/step=%5.1f  %7.3f ms/step" % (n, c / NS, t / NS / 1e3))
print("largest gaps (us @ ms since start): " + ", ".join("%.0f@%.1f %s" % (g, at, n[:24]) for g, at, n in sorted(big, reverse=True)[:15]))
