"""CPU restatement of the reference's per-box mask loss (TEST INFRASTRUCTURE ONLY).

Follows ndsis/modules/loss.py:284-318 (MaskLoss._single_sample_loss / forward): one mean BCE-with-logits per box, NaN
(empty boxes) filtered, class-weighted average, `default_loss` (0) when nothing is valid.  PINNED by
tests/golden/maskloss.pt, the outputs of the unmodified reference (oracle/make_golden_maskloss.py)."""
import torch


def box_means(logits_list, target_list):
    """fp64 accumulation of torch's stable formulation; NaN for an empty box (mean of nothing)."""
    out = []
    for x, t in zip(logits_list, target_list):
        if len(x) == 0:
            out.append(float("nan"))
            continue
        x64, t64 = x.double(), t.double()
        m = (-x64).clamp(min=0)
        out.append(float(((1 - t64) * x64 + m + torch.log(torch.exp(-m) + torch.exp(-x64 - m))).mean()))
    return torch.tensor(out, dtype=torch.float64)


def mask_loss(masks_output, mask_target, class_target, class_weights=None):
    outs = [m for s in masks_output for m in s]
    tgts = [m for s in mask_target for m in s]
    if not outs:
        return 0.0
    means = box_means(outs, tgts)
    valid = ~torch.isnan(means)
    if not bool(valid.any()):
        return 0.0
    cls = torch.cat(class_target)[valid]
    if class_weights is None:
        return float(means[valid].mean())
    w = class_weights.double()[cls]
    return float((means[valid] * w).sum() / w.sum())
