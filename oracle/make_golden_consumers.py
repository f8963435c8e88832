"""Golden vectors for the CONSUMERS of the crop selection, produced by the UNMODIFIED reference classes (run in the build
container only; /root/reference is not on the GPU box):

    python oracle/make_golden_consumers.py   ->   tests/golden/consumers.pt

* `SparseMaskPredictor.forward`      /root/reference/ndsis/modules/model.py:826-882
* `SparseMaskLossSelector.forward`   /root/reference/ndsis/modules/model.py:1152-1227 (both branches: selection by overlap
  thresholds, and by a given class-selector description)
* `split_select_nd`                  /root/reference/ndsis/utils/basic_functions.py:177-216
The selection itself (`is_inside`, counts, splits) comes from the reference's own `roi_cut`."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, "/root/reference")

import scn_oracle

sys.modules["sparseconvnet"] = scn_oracle

import numpy as np
import torch

from sparse_rcnn_b200.synthetic import make_batch, make_boxes

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    from ndsis.modules.model import SparseMaskLossSelector, SparseMaskPredictor
    from ndsis.modules.roi_select_bbox_transform import BBoxTransformerSlice
    from ndsis.modules.roi_select_sparse import roi_cut
    from ndsis.utils.basic_functions import split_select_nd
    coords, feats, size, bs, splits = make_batch(3, 21, spatial_size=(32, 32, 16), room=(16, 16, 8), room_offset=(4, 4, 1),
                                                 n_furniture=1, density=0.5)
    boxes = make_boxes(coords, 5, 3, (32, 32, 16))
    for b in boxes:                                                   # small fixture: shrink the boxes around their centres
        c = (b[:, 0] + b[:, 1]) / 2
        b[:, 0], b[:, 1] = c - (c - b[:, 0]) * 0.5, c + (b[:, 1] - c) * 0.5
    boxes[1] = boxes[1][:2]                                           # ragged box counts
    boxes[2][1] = torch.tensor([[40., 40, 20], [45, 45, 25]])         # a box that contains no point
    bt, counts, assoc = BBoxTransformerSlice(clip=False, resize=None)(boxes, size)
    new_coords, _, is_inside = roi_cut(coords, feats, bt, assoc)
    g = torch.Generator().manual_seed(11)
    n_rows, n_cls = len(new_coords), 18
    scores = torch.randn(n_rows, n_cls, generator=g)
    BB = len(bt)
    cls = torch.randint(-1, n_cls, (BB,), generator=g)                # some invalid (negative) classes
    cls_hi = torch.randint(-1, n_cls + 2, (BB,), generator=g)         # ... and some beyond num_valid
    sel = (is_inside, counts, list(splits))
    out = dict(scene_seed=21, box_seed=3, boxes=boxes, counts=counts, splits=list(splits), scores=scores, cls=cls,
               inside_packed=np.packbits(is_inside.numpy(), axis=1), inside_shape=tuple(is_inside.shape))
    out["cls_hi"] = cls_hi
    for nv, c in ((0, cls), (n_cls, cls_hi)):
        masks = SparseMaskPredictor(num_valid=nv)(scores, sel, c)
        out["predictor_%d" % nv] = [m.clone() for m in masks]
    # block-diagonal split used by both consumers
    blocks = split_select_nd(is_inside, torch.tensor([counts, list(splits)]))
    out["blocks_sum"] = [int(b.sum()) for b in blocks]
    out["blocks_shape"] = [tuple(b.shape) for b in blocks]
    # loss selector
    n_gt = [3, 1, 4]
    gt_labels = [torch.randint(0, n_cls, (k,), generator=g) for k in n_gt]
    gt_masks = [torch.rand(k, p, generator=g) < 0.4 for k, p in zip(n_gt, splits)]
    max_ov = [torch.rand(c, generator=g) for c in counts]
    arg_ov = [torch.randint(0, k, (c,), generator=g) for k, c in zip(n_gt, counts)]
    tuples = [(None, None, m, a) for m, a in zip(max_ov, arg_ov)]
    assoc_given = [torch.randint(0, k, (c,), generator=g) for k, c in zip(n_gt, counts)]
    descr = [types.SimpleNamespace(gt_association=a) for a in assoc_given]
    out.update(gt_labels=gt_labels, gt_masks=gt_masks, max_ov=max_ov, arg_ov=arg_ov, assoc_given=assoc_given)
    selector = SparseMaskLossSelector(positive_threshold=0.5)
    for name, d, t in (("loss_by_overlap", None, tuples), ("loss_by_description", descr, None)):
        pred, gt, labels = selector(scores, sel, d, t, gt_labels, gt_masks)
        out[name] = dict(pred=[[m.clone() for m in s] for s in pred], gt=[[m.clone() for m in s] for s in gt],
                         labels=[l.clone() for l in labels])
        print(name, [len(s) for s in pred], [int(l.numel()) for l in labels])
    torch.save(out, os.path.join(OUT, "consumers.pt"))
    print("rows", n_rows, "boxes", BB, "blocks", out["blocks_shape"])


if __name__ == "__main__":
    main()
