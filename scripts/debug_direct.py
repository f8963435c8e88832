import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
from sparse_rcnn_b200 import pipeline, scn
def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
dev = torch.device("cuda:0")
scn.set_precision(sys.argv[1] if len(sys.argv) > 1 else "tf32")
data, labels = bench.make_inputs(0, scene_kw=bench.CPU_SAMPLE)
def run(direct, steps=2):
    tr = pipeline.BackboneTrainer(dev, seed=3)
    if not direct:
        for p in tr.parameters(): p._scn_grad_hook = None
    tr.optimizer = torch.optim.SGD(tr.parameters(), lr=0.0)
    for _ in range(steps): loss = tr.step(data, labels)
    return {n: p.grad.detach().clone() for n, p in tr.named_parameters()}, float(loss)
for a, b, tag in ((True, False, "direct vs regular"), (False, False, "regular vs regular"), (True, True, "direct vs direct")):
    ga, la = run(a); gb, lb = run(b)
    errs = sorted(((rel_err(ga[n], gb[n]), n, tuple(ga[n].shape)) for n in ga), reverse=True)
    print(tag, "loss", la, lb, "worst:", [(round(e, 6), n, s) for e, n, s in errs[:4]], flush=True)
ga, _ = run(True, 1); gb, _ = run(False, 1)
errs = sorted(((rel_err(ga[n], gb[n]), n) for n in ga), reverse=True)
print("1 step direct vs regular", [(round(e, 6), n) for e, n in errs[:4]])
