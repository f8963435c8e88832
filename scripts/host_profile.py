"""Host-side profile of the training step: CPU time per step (no sync) vs GPU time, and cProfile top entries."""
import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from sparse_rcnn_b200 import pipeline, scn
dev = torch.device("cuda:0"); scn.set_precision("tf32")
tr = pipeline.BackboneTrainer(dev)
data, labels = bench.make_inputs(0)
data = (data[0].to(dev), data[1].to(dev), data[2], data[3], data[4]); labels = labels.to(dev)
for _ in range(5): tr.step(data, labels)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(10): tr.step(data, labels)
t_cpu = time.perf_counter() - t0
e1.record(); torch.cuda.synchronize()
print("per step: CPU issue time %.2f ms, GPU elapsed %.2f ms" % (t_cpu * 100, e0.elapsed_time(e1) / 10))
pr = cProfile.Profile(); pr.enable()
for _ in range(5): tr.step(data, labels)
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue()[:6000])
