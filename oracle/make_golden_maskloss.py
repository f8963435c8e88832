"""Golden vectors for the per-box mask loss, produced by the UNMODIFIED reference `ndsis.modules.loss.MaskLoss`
on the CPU (loss value + gradient w.r.t. every box's logits).  Build container only:  python oracle/make_golden_maskloss.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, "/root/reference")
import scn_oracle
sys.modules["sparseconvnet"] = scn_oracle          # ndsis.modules.loss imports model.py, which imports sparseconvnet
import torch
from ndsis.modules.loss import MaskLoss             # noqa: E402
from sparse_rcnn_b200.synthetic import make_mask_loss_case      # noqa: E402

cases = []
for seed, bps, weighted, empty_every in [(0, [7, 12], False, 5), (1, [30, 1, 25], True, 4), (2, [3], True, 0),
                                         (3, [4, 4], False, 1), (4, [], False, 5)]:
    outs, tgts, cls = make_mask_loss_case(seed, bps, empty_every=empty_every)
    for s in outs:
        for m in s:
            m.requires_grad_()
    w = (torch.arange(18, dtype=torch.float32) % 5 + 0.5) if weighted else None
    ml = MaskLoss(class_weights=w)
    loss = ml(outs, tgts, cls) if bps else ml([], [], [torch.zeros(0, dtype=torch.long)])
    flat = [m for s in outs for m in s]
    grads = []
    if loss.requires_grad:
        loss.backward()
        grads = [m.grad.clone() if m.grad is not None else torch.zeros_like(m) for m in flat]
    cases.append(dict(seed=seed, boxes_per_sample=bps, weighted=weighted, empty_every=empty_every, loss=float(loss),
                      grads=grads))
    print("case", seed, "loss", float(loss), "boxes", sum(bps), "elements", sum(len(m) for m in flat))
torch.save(dict(cases=cases), os.path.join(ROOT, "tests", "golden", "maskloss.pt"))
print("wrote tests/golden/maskloss.pt", os.path.getsize(os.path.join(ROOT, "tests", "golden", "maskloss.pt")), "bytes")
