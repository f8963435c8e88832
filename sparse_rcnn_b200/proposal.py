"""Proposal selection on the device: host-side mirror of the reference's
`ndsis/modules/proposal_selector.py` with the same class names, constructor arguments, forward signatures and return
formats, running `scn_nms3d` instead of the n x n IoU matrix + Python loop of n iterations
(`ndsis/utils/bbox.py:713-759`) and without the mid-forward `indices.cpu()` synchronisation (`proposal_selector.py:67`).
One host round trip remains: the per-sample survivor counts that size the returned lists."""
import torch
import torch.nn as nn

from . import _lib
from .scn.metadata import _ptr, _stream


def nms3d(boxes, thresh, max_keep=None):
    """boxes [B, n, 2, 3] fp32 CUDA, every sample sorted by descending score.
    -> (keep [B, n] bool == reference non_maximum_supression(boxes, thresh), keep_idx [B, max_keep] int32, counts [B] int32)."""
    if not boxes.is_cuda:
        raise RuntimeError("sparse_rcnn_b200.proposal needs CUDA tensors (there is no CPU fallback)")
    if boxes.dim() != 4 or boxes.shape[-2:] != (2, 3):
        raise RuntimeError("boxes must be [B, n, 2, 3] (start, stop)")
    B, n = boxes.shape[:2]
    max_keep = n if max_keep is None else int(max_keep)
    bx = boxes.detach().float().contiguous()
    dev = bx.device
    ws = torch.empty(int(_lib.raw("scn_nms3d_workspace_bytes")(B, n)), dtype=torch.uint8, device=dev)
    keep = torch.zeros((B, n), dtype=torch.uint8, device=dev)
    keep_idx = torch.zeros((B, max(max_keep, 1)), dtype=torch.int32, device=dev)
    counts = torch.zeros(B, dtype=torch.int32, device=dev)
    _lib.call("scn_nms3d", _ptr(bx), B, n, float(thresh), max_keep, _ptr(ws), _ptr(keep), _ptr(keep_idx), _ptr(counts),
              _stream())
    return keep.bool(), keep_idx, counts


class ProposalSelector(nn.Module):
    """reference ProposalSelector (proposal_selector.py:53-89): top-k by score, greedy NMS, keep the first
    num_keep_post_nms survivors.  Returns (scores, boxes, indices): lists over the batch; scores/boxes on the device,
    indices int64 on the CPU like the reference's."""

    def __init__(self, num_keep_pre_nms, num_keep_post_nms, thresh_nms):
        super().__init__()
        self.num_keep_pre_nms = num_keep_pre_nms
        self.num_keep_post_nms = num_keep_post_nms
        self.thresh_nms = thresh_nms

    def forward(self, rpn_score, rpn_bbox):
        if self.num_keep_pre_nms > 0:
            k = min(self.num_keep_pre_nms, rpn_score.shape[1])
            if k != self.num_keep_pre_nms:
                raise RuntimeError("selected index k out of range")      # torch.topk's error in the reference
            score, indices = torch.topk(rpn_score, k, dim=1, sorted=True)
        else:
            score, indices = torch.sort(rpn_score, dim=1, descending=True)
        B, n = score.shape
        bbox = torch.gather(rpn_bbox, 1, indices.view(B, n, 1, 1).expand(B, n, *rpn_bbox.shape[2:]))
        post = self.num_keep_post_nms
        _, keep_idx, counts = nms3d(bbox, self.thresh_nms, post)
        counts = counts.cpu().tolist()                                   # the one host round trip
        keep_idx = keep_idx.long()
        out_s, out_b, out_i = [], [], []
        for b in range(B):
            sel = keep_idx[b, :counts[b]]
            out_s.append(score[b, sel])
            out_b.append(bbox[b, sel])
            out_i.append(indices[b, sel])
        idx_cpu = torch.cat(out_i).cpu().split(counts) if B else ()
        return out_s, out_b, list(idx_cpu)


class RoiSelector(nn.Module):
    """reference RoiSelector (proposal_selector.py:25-50)."""

    def __init__(self, *args, detach=True, **kwargs):
        super().__init__()
        self.proposal_selector = ProposalSelector(*args, **kwargs)
        self.detach = detach

    def forward(self, rpn_bbox, rpn_score, anchor_description):
        if self.detach:
            rpn_bbox = rpn_bbox.detach()
            rpn_score = rpn_score.detach()
        roi_bbox_raw = anchor_description(rpn_bbox)
        roi_score_raw = torch.sigmoid(rpn_score)
        return self.proposal_selector(roi_score_raw, roi_bbox_raw)


def get_roi_selector(num_keep_pre_nms=1000, num_keep_post_nms=500, thresh_nms=0.5, val_num_keep_pre_nms=None,
                     val_num_keep_post_nms=None, val_thresh_nms=None):
    """reference get_roi_selector (proposal_selector.py:6-22); ConditionalStage = train/eval switch
    (ndsis/modules/custom_container.py)."""
    sel = RoiSelector(num_keep_pre_nms, num_keep_post_nms, thresh_nms)
    if val_num_keep_pre_nms:
        sel = _Conditional(sel, RoiSelector(val_num_keep_pre_nms, val_num_keep_post_nms, val_thresh_nms))
    return sel


class _Conditional(nn.Module):
    """reference ConditionalStage (custom_container.py:102-120): same attribute names => same state_dict keys."""

    def __init__(self, train_module, val_module):
        super().__init__()
        self.train_module, self.val_module = train_module, val_module

    def forward(self, *args, **kwargs):
        return (self.train_module if self.training else self.val_module)(*args, **kwargs)
