import gc, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
from sparse_rcnn_b200 import pipeline, scn
from sparse_rcnn_b200.synthetic import make_boxes
dev = torch.device("cuda:0"); scn.set_precision("tf32")
inf = pipeline.SparseInference(dev)
host = [bench.make_inputs(i) for i in range(4)]
pinned = [(d[0].pin_memory(), d[1].pin_memory(), d[2], d[3], d[4]) for d, _ in host]
boxes = [make_boxes(d[0], 256, 7 + i) for i, (d, _) in enumerate(host)]
for i in range(8): inf(pinned[i % 4], boxes[i % 4])
torch.cuda.synchronize()
def st():
    s = torch.cuda.memory_stats(); return s["num_device_alloc"], s["num_device_free"], s["num_alloc_retries"]
times = []
a = st(); g0 = gc.get_count(); gs0 = [x["collections"] for x in gc.get_stats()]
for i in range(40):
    t0 = time.perf_counter()
    r = inf(pinned[i % 4], boxes[i % 4]); r["mpn_class"].argmax(1).cpu()
    times.append((time.perf_counter() - t0) * 1e3)
b = st(); gs1 = [x["collections"] for x in gc.get_stats()]
print("per-scene ms: min %.2f median %.2f max %.2f; >15 ms: %s" % (min(times), sorted(times)[20], max(times), [round(t, 1) for t in times if t > 15]))
print("cudaMalloc/free/retries during loop:", [y - x for x, y in zip(a, b)], "gc collections per generation:", [y - x for x, y in zip(gs0, gs1)])
print("reserved %.0f MB" % (torch.cuda.memory_reserved() / 2**20))
