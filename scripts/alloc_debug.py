import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
from sparse_rcnn_b200 import pipeline, scn
dev = torch.device("cuda:0"); scn.set_precision("tf32")
tr = pipeline.BackboneTrainer(dev)
inputs = []
for i in range(4):
    d, l = bench.make_inputs(1000 * 0 + i)
    inputs.append(((d[0].to(dev), d[1].to(dev), d[2], d[3], d[4]), l.to(dev)))
for i in range(8): tr.step(*inputs[i % 4])
torch.cuda.synchronize()
def stats():
    s = torch.cuda.memory_stats()
    return {k: s[k] for k in ("num_device_alloc", "num_device_free", "num_alloc_retries", "num_sync_all_streams", "segment.all.allocated", "segment.all.freed", "allocation.all.allocated")}
a = stats(); t0 = time.perf_counter()
for i in range(20): tr.step(*inputs[i % 4])
torch.cuda.synchronize(); dt = time.perf_counter() - t0
b = stats()
print("per step: %.2f ms;" % (dt * 50), {k: (b[k] - a[k]) / 20 for k in a})
print("reserved %.1f MB allocated peak %.1f MB" % (torch.cuda.memory_reserved() / 2**20, torch.cuda.max_memory_allocated() / 2**20))
# same scene every step
for i in range(5): tr.step(*inputs[0])
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(20): tr.step(*inputs[0])
torch.cuda.synchronize(); print("same scene every step: %.2f ms/step" % ((time.perf_counter() - t0) * 50))
