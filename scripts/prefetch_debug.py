import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
from sparse_rcnn_b200 import pipeline, scn
from sparse_rcnn_b200.scn import metadata as M
if len(sys.argv) > 1: sys.setswitchinterval(float(sys.argv[1]))
dev = torch.device("cuda:0"); scn.set_precision("tf32")
tr = pipeline.BackboneTrainer(dev)
inputs = []
for i in range(4):
    d, l = bench.make_inputs(i)
    c0 = d[0].to(dev) if os.environ.get('DEVCOORDS') else d[0]
    inputs.append(((c0, d[1].to(dev), d[2], d[3], d[4]), l.to(dev)))
T = {"take": 0.0, "rec": 0.0, "build": 0.0}
orig_take = M.GeometryPrefetcher.take
def take(self, coords):
    item = self.pending.pop(id(coords), None)
    if item is None: return None
    t0 = time.perf_counter(); md = item[0].result(); t1 = time.perf_counter()
    cur = torch.cuda.current_stream(self.device); cur.wait_event(md._ready)
    n = 0
    for t in M._walk_tensors(md, set()):
        if t.is_cuda: t.record_stream(cur); n += 1
    t2 = time.perf_counter(); T["take"] += t1 - t0; T["rec"] += t2 - t1; T["n"] = n
    return md
M.GeometryPrefetcher.take = take
orig_build = M.GeometryPrefetcher._build
def build(self, *a):
    t0 = time.perf_counter(); r = orig_build(self, *a); T["build"] += time.perf_counter() - t0; return r
M.GeometryPrefetcher._build = build
for pf in (False, True):
    for i in range(5): tr.step(*inputs[i % 4], next_data=inputs[(i + 1) % 4][0] if pf else None)
    torch.cuda.synchronize(); T.update(take=0.0, rec=0.0, build=0.0)
    t0 = time.perf_counter()
    for i in range(20): tr.step(*inputs[i % 4], next_data=inputs[(i + 1) % 4][0] if pf else None)
    torch.cuda.synchronize()
    print("prefetch=%s: %.2f ms/step; per step: wait for build %.2f ms, record_stream %.2f ms (%s tensors), build (worker wall) %.2f ms" % (
        pf, (time.perf_counter() - t0) * 50, T["take"] * 50, T["rec"] * 50, T.get("n"), T["build"] * 50))
