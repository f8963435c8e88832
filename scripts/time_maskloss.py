"""Per-box mask loss at the inference/training size of the shipped config (256 boxes, ~615k (box, point) rows): this
library's MaskLoss (one segment kernel each way) vs the reference algorithm (loss.py:284-318: one BCE launch per box) on the
same GPU, forward + backward."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from sparse_rcnn_b200 import losses
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
counts = torch.randint(100, 4800, (256,), generator=g).tolist()
outs = [[(torch.randn(n, generator=g) * 3).to(dev).requires_grad_() for n in counts]]
tgts = [[(torch.rand(n, generator=g) > 0.5).to(dev) for n in counts]]
cls = [torch.randint(0, 18, (256,), generator=g).to(dev)]
w = (torch.arange(18, dtype=torch.float32) % 5 + 0.5).to(dev)


def ref_loss():
    losses_ = torch.stack([F.binary_cross_entropy_with_logits(o, t.float()) for o, t in zip(outs[0], tgts[0])])
    valid = ~torch.isnan(losses_)
    mw = w[cls[0][valid]]
    return (losses_[valid] * mw).sum() / mw.sum()


ml = losses.MaskLoss(class_weights=w).to(dev)
for name, fn in (("this library", lambda: ml(outs, tgts, cls)), ("reference algorithm", ref_loss)):
    for _ in range(3):
        fn().backward()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        l = fn(); l.backward()
    torch.cuda.synchronize()
    print("%-20s fwd+bwd %.2f ms  (loss %.6f, %d rows, 256 boxes)" % (name, (time.perf_counter() - t0) * 100, float(l), sum(counts)))
