import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
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
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(10): inf(pdata, boxes)
t_cpu = time.perf_counter() - t0
e1.record(); torch.cuda.synchronize()
print("inference per scene: CPU issue %.2f ms, GPU elapsed %.2f ms" % (t_cpu * 100, e0.elapsed_time(e1) / 10))
pr = cProfile.Profile(); pr.enable()
for _ in range(5): inf(pdata, boxes)
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[:4500])
