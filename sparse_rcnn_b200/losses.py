"""Loss-side per-box reductions on the device: host-side mirror of the reference's `MaskLoss`
(`ndsis/modules/loss.py:271-318`).  The reference calls `binary_cross_entropy_with_logits(reduction='mean')` once per box
(up to 256 tiny launches per step plus their backwards); here the boxes' logits are one concatenated tensor with a CSR of
box boundaries and `scn_segment_bce_fwd/bwd` produce every box's mean loss / gradient in one launch each.  Everything after
the per-box means (NaN filter for empty boxes, class-weighted average, the default loss) is the reference's arithmetic, kept
on the device (no boolean-index host sync)."""
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _lib
from .scn.metadata import _ptr, _stream


class SegmentBCEFunction(Function):
    @staticmethod
    def forward(ctx, logits, targets, seg_ptr):
        if not logits.is_cuda:
            raise RuntimeError("sparse_rcnn_b200.losses needs CUDA tensors (there is no CPU fallback)")
        x = logits.contiguous().float()
        t = targets.contiguous().to(torch.uint8)
        n_seg = seg_ptr.numel() - 1
        out = torch.empty(n_seg, dtype=torch.float32, device=x.device)
        _lib.call("scn_segment_bce_fwd", _ptr(x), _ptr(t), _ptr(seg_ptr), n_seg, _ptr(out), _stream())
        ctx.save_for_backward(x, t, seg_ptr)
        return out

    @staticmethod
    def backward(ctx, g):
        x, t, seg_ptr = ctx.saved_tensors
        gx = torch.zeros_like(x)
        # NaN means of empty boxes are filtered out downstream; their (empty) segments get no gradient anyway
        g = torch.nan_to_num(g.contiguous().float(), nan=0.0)
        _lib.call("scn_segment_bce_bwd", _ptr(x), _ptr(t), _ptr(seg_ptr), seg_ptr.numel() - 1, _ptr(g), _ptr(gx), _stream())
        return gx, None, None


def segment_bce_with_logits(logits, targets, counts):
    """Mean BCE-with-logits of every segment.  logits [M] fp32, targets [M] bool/uint8, counts: per-segment lengths (host
    ints, they are the crop's `bbox_sample_count`-style metadata).  -> [len(counts)] fp32, NaN for empty segments."""
    ptr = torch.zeros(len(counts) + 1, dtype=torch.int32)
    ptr[1:] = torch.cumsum(torch.as_tensor(counts, dtype=torch.int64), 0).to(torch.int32)
    return SegmentBCEFunction.apply(logits, targets, ptr.to(logits.device, non_blocking=True))


class MaskLoss(nn.Module):
    """reference MaskLoss (loss.py:271-318): same constructor, parameters (`class_weights`, `default_loss`) and forward
    signature: masks_output / mask_target are lists (samples) of lists (boxes) of [n_b] tensors, class_target a list
    (samples) of int64 [boxes]."""

    def __init__(self, class_weights=None, dtype=torch.get_default_dtype()):
        super().__init__()
        self.class_weights = None if class_weights is None else nn.Parameter(class_weights, requires_grad=False)
        self.default_loss = nn.Parameter(torch.zeros([], dtype=dtype), requires_grad=False)
        self.mask_loss_reduction = 'mean'
        self.weight_range = 16

    def forward(self, masks_output, mask_target, class_target):
        outs = [m for sample in masks_output for m in sample]
        tgts = [m for sample in mask_target for m in sample]
        if not outs:
            return self.default_loss
        return self.forward_flat(torch.cat(outs), torch.cat(tgts), [len(m) for m in outs], torch.cat(class_target))

    def forward_flat(self, logits, targets, counts, class_target_cat):
        """Same loss from the concatenated (box, point) layout `roi.SparseMaskLossSelector.flat` holds: logits [M], targets
        [M], per-box lengths (host ints), labels [boxes] -- no per-box Python, no list concatenation."""
        if not len(counts):
            return self.default_loss
        maskwise_loss = segment_bce_with_logits(logits, targets, counts)
        valid = ~torch.isnan(maskwise_loss)
        if self.class_weights is not None:
            w = self.class_weights[class_target_cat] * valid
        else:
            w = valid.to(maskwise_loss.dtype)
        wsum = w.sum()
        loss = (torch.where(valid, maskwise_loss, torch.zeros_like(maskwise_loss)) * w).sum() / wsum
        # no valid box at all -> the reference returns default_loss
        return torch.where(valid.any(), loss, self.default_loss.to(loss.dtype))
