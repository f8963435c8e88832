"""Sparse inference throughput (bench.py's inference object: raw RPN outputs -> selection -> class + mask networks) against
the number of host threads of SparseInference.run_many.  python scripts/infer_workers.py [workers ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from sparse_rcnn_b200 import pipeline, scn
from sparse_rcnn_b200.synthetic import make_rpn_outputs
dev = torch.device("cuda:0"); scn.set_precision("tf32")
if os.environ.get("SCN_SWITCH_INTERVAL"): sys.setswitchinterval(float(os.environ["SCN_SWITCH_INTERVAL"]))
host = [bench.balanced_inputs(0, i) for i in range(4)]
pinned = [(d[0].pin_memory(), d[1].pin_memory(), d[2], d[3], d[4]) for d, _ in host]
rpn = [tuple(t.pin_memory() for t in make_rpn_outputs(d[0], 30000, 256, 7 + i)) for i, (d, _) in enumerate(host)]
infer = pipeline.SparseInference(dev)
N = 32
sc, bx = [pinned[i % 4] for i in range(N)], [rpn[i % 4] for i in range(N)]
consume = lambda i, res: (int(res["mpn_mask"].shape[0]), res["mpn_class"].argmax(1).cpu(), len(res["roi_index"][0]))
for i in range(8): infer(pinned[i % 4], rpn=rpn[i % 4])
for nw in [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4]:
    for _ in range(2): infer.run_many(sc, rpn=bx, workers=nw, consume=consume)
    torch.cuda.synchronize()
    res = []
    for rep in range(3):
        m0 = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); infer.run_many(sc, rpn=bx, workers=nw, consume=consume); e1.record(); torch.cuda.synchronize()
        res.append("%.2f ms/scene (%d cudaMalloc)" % (e0.elapsed_time(e1) / N, torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - m0))
    print("workers=%d PDL=%s: %s" % (nw, os.environ.get("SCN_PDL", "1"), ", ".join(res)), flush=True)
