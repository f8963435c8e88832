"""Proposal selection (top-k + NMS) at the reference's shipped sizes (run.py:848-853): this library vs the reference
algorithm (utils/bbox.py:713-759 restated with torch ops on the SAME GPU, i.e. n x n IoU + a Python loop of n steps)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sparse_rcnn_b200 import proposal
from sparse_rcnn_b200.synthetic import make_proposals
dev = torch.device("cuda:0")


def ref_nms_torch(boxes, thr):            # boxes [B, n, 2, 3] on the GPU, the reference's vectorised loop
    start, end = boxes[:, :, 0], boxes[:, :, 1]
    area = (end - start).prod(-1)
    inter = (torch.min(end[:, :, None], end[:, None]) - torch.max(start[:, :, None], start[:, None])).clamp(min=0).prod(-1)
    over = inter / (area[:, :, None] + area[:, None] - inter)
    m = torch.tril(over > thr, diagonal=-1)
    keep = m.new_ones(boxes.shape[:2])
    for box in m.unbind(-1):
        keep &= ~box
        m &= ~box.unsqueeze(-2)
    return keep


for B, A, pre, post, thr in [(1, 30000, 1024, 256, 0.5), (8, 30000, 1024, 256, 0.5), (1, 30000, 1024, 32, 0.3)]:
    score, boxes = make_proposals(3, B, A)
    score, boxes = score.to(dev), boxes.to(dev)
    sel = proposal.ProposalSelector(pre, post, thr)
    for _ in range(3): sel(score, boxes)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): out = sel(score, boxes)
    torch.cuda.synchronize(); t_new = (time.perf_counter() - t0) / 20 * 1e3
    top, idx = torch.topk(score, pre, dim=1)
    sb = torch.gather(boxes, 1, idx.view(B, pre, 1, 1).expand(B, pre, 2, 3))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): k = proposal.nms3d(sb, thr, post)
    e1.record(); torch.cuda.synchronize()
    t_k = e0.elapsed_time(e1) / 20
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(2): kr = ref_nms_torch(sb, thr)
    torch.cuda.synchronize(); t_ref = (time.perf_counter() - t0) / 2 * 1e3
    same = bool(torch.equal(kr, k[0]))
    print("B=%d anchors=%d pre=%d post=%d thr=%.1f: selector %.3f ms (NMS kernels %.3f ms) | reference algorithm on the GPU %.1f ms"
          " (NMS only) | same keep mask: %s" % (B, A, pre, post, thr, t_new, t_k, t_ref, same))
