"""CPU restatement of the reference's proposal selection (TEST INFRASTRUCTURE ONLY).

Follows ndsis/modules/proposal_selector.py:60-89 (ProposalSelector.forward) and ndsis/utils/bbox.py:713-759
(non_maximum_supression: strict-lower-triangle "IoU > threshold" matrix, then every box that is still alive when its
turn comes suppresses the lower-scored boxes it overlaps) with the IoU of bbox.py:235-240.  PINNED: tests/golden/nms.pt
holds the outputs of the UNMODIFIED reference functions (oracle/make_golden_nms.py), and
tests/test_golden.py::test_nms_oracle_matches_reference compares this file against them bit for bit."""
import torch


def iou_matrix(boxes):
    """boxes [n, 2, 3] -> [n, n] fp32 IoU (bbox.py:598-620, 235-240)."""
    start, end = boxes[:, 0], boxes[:, 1]
    area = (end - start).prod(-1)
    max_start = torch.max(start[:, None], start[None])
    min_end = torch.min(end[:, None], end[None])
    inter = (min_end - max_start).clamp(min=0).prod(-1)
    return inter / (area[:, None] + area[None] - inter)


def nms(boxes, thresh):
    """boxes [n, 2, 3] sorted by descending score -> keep [n] bool (bbox.py:713-759, one sample)."""
    n = len(boxes)
    over = iou_matrix(boxes) > thresh
    keep = torch.ones(n, dtype=torch.bool)
    for j in range(n):
        if keep[j]:
            later = over[j].clone()
            later[:j + 1] = False
            keep &= ~later
    return keep


def select(rpn_score, rpn_bbox, pre, post, thresh):
    """ProposalSelector.forward on CPU tensors: lists (scores, boxes, indices) over the batch."""
    if pre > 0:
        score, idx = torch.topk(rpn_score, pre, dim=1, sorted=True)
    else:
        score, idx = torch.sort(rpn_score, dim=1, descending=True)
    out = ([], [], [])
    for b in range(len(rpn_bbox)):
        bb = rpn_bbox[b][idx[b]]
        k = nms(bb, thresh)
        out[0].append(score[b][k][:post]), out[1].append(bb[k][:post]), out[2].append(idx[b][k][:post])
    return out
