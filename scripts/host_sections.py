"""Host wall time of the sections of BackboneTrainer.step in the bench's e2e configuration (pinned inputs, staging, late
build): where the training thread spends its time and how long it blocks in the late build's row-count round trips."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from sparse_rcnn_b200 import pipeline, scn
from sparse_rcnn_b200.scn import metadata
dev = torch.device("cuda:0"); scn.set_precision("tf32")
tr = pipeline.BackboneTrainer(dev)
host = [bench.balanced_inputs(0, i) for i in range(4)]
pinned = [((d[0].pin_memory(), d[1].pin_memory(), d[2], d[3], d[4]), l.pin_memory()) for d, l in host]
tr.stage_uploads = True
tr.build_late = os.environ.get("BUILD_LATE", "1") == "1"
T = {}
def wrap(obj, name, key):
    f = getattr(obj, name)
    def g(*a, **k):
        t = time.perf_counter()
        try:
            return f(*a, **k)
        finally:
            T[key] = T.get(key, 0.0) + time.perf_counter() - t
    setattr(obj, name, g)
wrap(tr, "stage", "stage (3 H2D enqueues)")
wrap(tr, "prefetch", "late build (incl. round trips)")
wrap(tr.buckets, "zero", "buckets.zero")
wrap(tr.buckets, "finish", "buckets.finish")
wrap(tr.optimizer, "step", "optimizer.step")
wrap(scn.functions, "pack_all", "pack_all")
wrap(tr.backbone, "forward", "backbone forward enqueue")
orig_sync = torch.cuda.Event.synchronize
def ev_sync(self):
    t = time.perf_counter(); orig_sync(self); T["event.synchronize (round trips, loss)"] = T.get("event.synchronize (round trips, loss)", 0.0) + time.perf_counter() - t
torch.cuda.Event.synchronize = ev_sync
N = 30
def loop(n):
    slots = [torch.zeros(1).pin_memory() for _ in range(2)]; evs = [torch.cuda.Event() for _ in range(2)]; pending = None
    for i in range(n):
        loss = tr.step(*pinned[i % 4], next_batch=pinned[(i + 1) % 4])
        slots[i & 1].copy_(loss.reshape(1), non_blocking=True); evs[i & 1].record()
        if pending is not None: evs[pending].synchronize()
        pending = i & 1
loop(8); torch.cuda.synchronize(); T.clear()
t0 = time.perf_counter(); loop(N); torch.cuda.synchronize(); wall = time.perf_counter() - t0
print("BUILD_LATE=%s: %.3f ms per step wall" % (os.environ.get("BUILD_LATE", "1"), wall / N * 1e3))
for k, v in sorted(T.items(), key=lambda x: -x[1]): print("   %-44s %.3f ms per step" % (k, v / N * 1e3))
