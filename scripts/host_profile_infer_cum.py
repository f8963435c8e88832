"""Host-side cumulative profile of one sparse inference scene from raw proposals (which component holds the GIL how long)."""
import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from sparse_rcnn_b200 import pipeline, scn
from sparse_rcnn_b200.synthetic import make_rpn_outputs
dev = torch.device("cuda:0"); scn.set_precision("tf32")
inf = pipeline.SparseInference(dev)
data, _ = bench.make_inputs(0)
pdata = (data[0].pin_memory(), data[1].pin_memory(), data[2], data[3], data[4])
rpn = tuple(t.pin_memory() for t in make_rpn_outputs(data[0], 30000, 256, 7))
for _ in range(6): inf(pdata, rpn=rpn)
torch.cuda.synchronize()
# wall time of the sections, GPU drained before each scene so that waits inside are real waits
T = {}
def wrap(obj, name, key):
    f = getattr(obj, name)
    def g(*a, **k):
        t = time.perf_counter()
        try: return f(*a, **k)
        finally: T[key] = T.get(key, 0.0) + time.perf_counter() - t
    setattr(obj, name, g)
for name in ("backbone", "seg", "class_network", "mask_network", "roi_selector"):
    m = getattr(inf, name); wrap(m, "forward", name)
N = 10
t0 = time.perf_counter()
for _ in range(N): inf(pdata, rpn=rpn)
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
print("per scene: host issue %.2f ms; sections (host wall incl. their round trips): %s" % (
    t_issue / N * 1e3, ", ".join("%s %.2f" % (k, v / N * 1e3) for k, v in T.items())))
pr = cProfile.Profile(); pr.enable()
for _ in range(5): inf(pdata, rpn=rpn)
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(60); print(s.getvalue()[:9000])
