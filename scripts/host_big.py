import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
from sparse_rcnn_b200 import pipeline, scn
dev = torch.device("cuda:0"); scn.set_precision("tf32")
tr = pipeline.BackboneTrainer(dev)
inputs = []
for i in range(4):
    d, l = bench.make_inputs(i)
    inputs.append(((d[0].to(dev), d[1].to(dev), d[2], d[3], d[4]), l.to(dev)))
ahead = bool(os.environ.get("AHEAD"))
tr.build_ahead = ahead
for i in range(6): tr.step(*inputs[i % 4], next_batch=inputs[(i + 1) % 4] if ahead else None)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for i in range(10): tr.step(*inputs[i % 4], next_batch=inputs[(i + 1) % 4] if ahead else None)
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14); print(s.getvalue()[:3500])
