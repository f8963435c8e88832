"""Per-proposal sparse crop ("mask crop"): host-side mirror of the reference's
ndsis/modules/roi_select_sparse.py + roi_select_bbox_transform.py on top of the crop kernels.

Reference behaviour restated (file:line in /root/reference):
  * BBoxTransformerSlice (roi_select_bbox_transform.py:87-97): cat boxes, optional divide by the
    feature stride, floor(start)/ceil(stop) -> int64 (utils/bbox.py:87-106), optional asymmetric
    clip start in [0,S-1], stop in [1,S] (utils/bbox.py:62-84); returns (boxes, counts, sample ids).
  * get_inside_indicator (roi_select_sparse.py:157-167): half-open integer box test AND sample match.
  * select_features / select_coords (:125-149): gather in (box, point) order; xyz stay ABSOLUTE,
    the batch column becomes the box index.
  * combiners (:55-122): Raw->Tensor uses InputLayer mode 4 with batch_size = #boxes; Tensor->Tensor
    uses mode 0.
The reference materialises a dense [BB, P] comparison, boolean-mask gathers and moves `is_inside`
and the cropped coords to the CPU (:180).  Here only the points of the box's own sample are tested,
the selection is compacted on the device (count -> scan -> ordered select) and the cropped keys feed
the device hash builder directly; `is_inside` is produced on the device (optionally copied to the
CPU to keep the reference's return format).
"""
import threading
import weakref

import torch
import torch.nn as nn

from . import _lib
from .scn.functions import Function
from .scn.metadata import _ptr, _stream, exclusive_scan, size_key

CROP_CHUNK = 2048      # SCN_CROP_CHUNK


def round_bbox(boxes):
    start, stop = boxes.unbind(-2)
    return torch.stack((start.floor(), stop.ceil()), dim=-2).long()


def clip_boxes_asymmetric(boxes, scene_shape):
    lo = boxes.new_tensor([[0], [1]])
    hi = torch.stack((scene_shape - 1, scene_shape))
    return torch.max(torch.min(boxes, hi), lo)


class BBoxTransformerSlice(nn.Module):
    def __init__(self, clip=False, resize=None):
        super().__init__()
        self.clip, self.resize = clip, resize

    def forward(self, bbox_batch, shape=None):
        counts = [len(b) for b in bbox_batch]
        raw = torch.cat(list(bbox_batch))
        if self.resize is not None:
            raw = raw / raw.new_tensor(self.resize)
        boxes = round_bbox(raw)
        if self.clip:
            boxes = clip_boxes_asymmetric(boxes, boxes.new_tensor(tuple(int(s) for s in shape)))
        assoc = torch.cat([torch.full((c,), i, dtype=torch.long) for i, c in enumerate(counts)]) \
            if counts else torch.zeros(0, dtype=torch.long)
        return boxes, counts, assoc


# ------------------------------------------------------------------------------- key cache
_TLS = threading.local()      # per host thread: SparseInference.run_many drives one scene stream per thread


def _key_cache():
    c = getattr(_TLS, "keys", None)
    if c is None:
        c = _TLS.keys = {}
    return c


def _packed_keys(coords, device):
    """Packed device keys of a raw [P,4] coordinate tensor (cached per tensor object)."""
    cache = _key_cache()
    k = id(coords)
    hit = cache.get(k)
    if hit is not None and hit[0]() is coords:
        return hit[1]
    P, ncol = coords.shape
    cdev = coords.to(device, non_blocking=True).contiguous()
    keys = torch.empty(P, dtype=torch.int64, device=device)
    err = torch.zeros(1, dtype=torch.int32, device=device)
    _lib.call("scn_pack_coords", _ptr(cdev), P, ncol, _ptr(keys), _ptr(err), _stream())
    if len(cache) > 64:
        cache.clear()
    cache[k] = (weakref.ref(coords), keys)
    return keys


def register_keys(coords, keys):
    """Let a crop reuse the keys the input layer already packed for the same coords tensor."""
    cache = _key_cache()
    if len(cache) > 64:
        cache.clear()
    cache[id(coords)] = (weakref.ref(coords), keys)


def clear_key_cache():
    _key_cache().clear()


class GatherRowsFunction(Function):
    """out[i] = x[idx[i]]; backward scatter-adds (boxes may overlap)."""

    @staticmethod
    def forward(ctx, x, idx):
        x = x.contiguous()
        n, C = idx.numel(), x.shape[1]
        out = torch.empty((n, C), dtype=torch.float32, device=x.device)
        _lib.call("scn_gather_rows", _ptr(x), x.stride(0), _ptr(idx), n, C, _ptr(out), C, _stream())
        ctx.idx, ctx.shape = idx, x.shape
        return out

    @staticmethod
    def backward(ctx, go):
        go = go.contiguous()
        gi = torch.zeros(ctx.shape, dtype=torch.float32, device=go.device)
        _lib.call("scn_scatter_add_rows", _ptr(go), _ptr(ctx.idx), ctx.idx.numel(), ctx.shape[1], _ptr(gi), _stream())
        return gi, None


class CropSelection:
    """Result of one crop: which source points fall into which box, in (box, point) order."""

    def __init__(self, sel_pt, new_keys, box_ptr, n_boxes, n_points, is_inside, bbox_sample_count, batch_splits):
        self.sel_pt, self.new_keys, self.box_ptr = sel_pt, new_keys, box_ptr
        self.n_boxes, self.n_points = n_boxes, n_points
        self._is_inside = is_inside
        self.bbox_sample_count, self.batch_splits = bbox_sample_count, batch_splits

    @property
    def total(self):
        return self.sel_pt.numel()

    def is_inside(self, cpu=False):
        """The reference's dense [n_boxes, n_points] bool matrix (roi_select_sparse.py:125-135).  The hot path works on the
        (box, point) CSR the crop kernel emits; the dense matrix (70 MB at 256 boxes x 273 k points) is only materialised
        when somebody asks for it -- here, from the CSR, once."""
        if self._is_inside is None:
            dev = self.sel_pt.device
            m = torch.zeros(self.n_boxes * self.n_points, dtype=torch.uint8, device=dev)
            if self.total:
                per_box = (self.box_ptr[1:] - self.box_ptr[:-1]).long()
                box_of = torch.repeat_interleave(torch.arange(self.n_boxes, device=dev), per_box, output_size=self.total)
                m[box_of * self.n_points + self.sel_pt.long()] = 1
            self._is_inside = m
        m = self._is_inside.view(self.n_boxes, self.n_points).bool()
        return m.cpu() if cpu else m

    def as_reference_tuple(self, cpu=True):
        """(is_inside, bbox_sample_count, batch_splits) -- the reference's `selection` triple."""
        return self.is_inside(cpu), self.bbox_sample_count, self.batch_splits

    def __iter__(self):
        return iter(self.as_reference_tuple())

    def __len__(self):
        return 3


def crop(point_keys, sample_ptr, max_sample_len, boxes, box_sample, want_is_inside=True):
    """point_keys int64[P] (device, grouped by sample), sample_ptr int32 [B+1] (device),
    boxes int64 [BB,2,3] (any device), box_sample int64 [BB].  One host sync (the selected count)."""
    dev = point_keys.device
    P, BB = point_keys.numel(), boxes.shape[0]
    s = _stream()
    b32 = boxes.to(device=dev, dtype=torch.int32).contiguous()
    bs32 = box_sample.to(device=dev, dtype=torch.int32).contiguous()
    n_chunks = max((max_sample_len + CROP_CHUNK - 1) // CROP_CHUNK, 1)
    counts = torch.empty(BB * n_chunks, dtype=torch.int32, device=dev)
    _lib.call("scn_crop_count", _ptr(point_keys), _ptr(sample_ptr), _ptr(b32), _ptr(bs32), BB, n_chunks,
              _ptr(counts), s)
    offsets = exclusive_scan(counts)
    total = int(offsets[-1].item()) if BB else 0
    sel_pt = torch.empty(total, dtype=torch.int32, device=dev)
    new_keys = torch.empty(total, dtype=torch.int64, device=dev)
    inside = torch.zeros(BB * P, dtype=torch.uint8, device=dev) if want_is_inside else None
    _lib.call("scn_crop_select", _ptr(point_keys), _ptr(sample_ptr), _ptr(b32), _ptr(bs32), BB, n_chunks,
              _ptr(offsets), P, _ptr(sel_pt), _ptr(new_keys), _ptr(inside), s)
    box_ptr = offsets[::n_chunks].contiguous() if BB else offsets          # [BB+1] first slot of every box
    return sel_pt, new_keys, box_ptr, total, inside


# ------------------------------------------------------------------------------- extractor / combiners
class RawScene:
    """extract() of the reference's Raw* combiners: scene = (coords, feats, spatial_size, ..., batch_splits)."""
    NEED_COORDS = True

    @staticmethod
    def extract(feature_map, device):
        coords, feats, spatial_size, *_, batch_splits = feature_map
        keys = _packed_keys(coords, device)
        splits = [int(v) for v in batch_splits]
        ptr = torch.tensor([0] + list(torch.tensor(splits).cumsum(0).tolist()), dtype=torch.int32).to(device)
        return keys, feats, spatial_size, splits, ptr, (max(splits) if splits else 0)


class TensorScene:
    """extract() of TensorToTensorFeatureExtractorCombiner (roi_select_sparse.py:100-110)."""
    NEED_COORDS = True

    @staticmethod
    def extract(feature_map, device):
        md = feature_map.metadata
        lvl = md.level(feature_map.spatial_size)
        ptr = lvl.batch_ptr(md.n_samples)
        splits = (ptr[1:] - ptr[:-1]).tolist()                 # host copy of B+1 ints (reference: bincount)
        return lvl.keys, feature_map.features, feature_map.spatial_size, splits, ptr, (max(splits) if splits else 0)


class SparseRoiCut(nn.Module):
    """SparseRoiCut (roi_select_sparse.py:29-52).  `scn` is the namespace used to build the cropped
    tensor; `raw_scene` selects RawToTensor (InputLayer mode 4) vs TensorToTensor (mode 0);
    `combine` in {'tensor','features','raw'}."""

    def __init__(self, scn, raw_scene=True, clip_boxes=False, resize_boxes=None, combine="tensor",
                 cpu_selection=False):
        super().__init__()
        self.scn = scn
        self.raw_scene = raw_scene
        self.bbox_transformer = BBoxTransformerSlice(clip=clip_boxes, resize=resize_boxes)
        self.combine, self.cpu_selection = combine, cpu_selection
        # True: the crop kernel also writes the dense [BB, P] byte matrix (zero fill + one byte per tested pair); False
        # (default): `selection.is_inside()` builds it from the (box, point) CSR on request
        self.dense_selection = False

    def forward(self, feature_map, bbox_batch):
        feats0 = feature_map[1] if self.raw_scene else feature_map.features
        dev = feats0.device
        keys, feats, spatial_size, splits, ptr, max_len = (RawScene if self.raw_scene else TensorScene).extract(
            feature_map, dev)
        boxes, counts, assoc = self.bbox_transformer(bbox_batch, spatial_size)
        sel_pt, new_keys, box_ptr, total, inside = crop(keys, ptr, max_len, boxes, assoc, want_is_inside=self.dense_selection)
        sel = CropSelection(sel_pt, new_keys, box_ptr, boxes.shape[0], keys.numel(), inside, counts, splits)
        new_feats = GatherRowsFunction.run(feats, sel_pt)
        out = combine_crop(self.scn, self.combine, new_keys, new_feats, spatial_size, boxes.shape[0],
                           mode=4 if self.raw_scene else 0)
        return out, sel


def combine_crop(scn, how, new_keys, new_feats, spatial_size, n_boxes, mode):
    if how == "features":
        return new_feats
    if how == "raw":
        return new_keys, new_feats, spatial_size, n_boxes
    md = scn.Metadata(3)
    size = torch.as_tensor(spatial_size, dtype=torch.long)
    fn = scn.ioLayers.InputLayerFunction
    f = getattr(fn, "run", fn.apply)(3, md, size, new_keys, new_feats, n_boxes, mode)
    return scn.SparseConvNetTensor(features=f, metadata=md, spatial_size=size)


class SparseRoiExtraCut(nn.Module):
    """SparseRoiExtraCut (roi_select_sparse.py:8-26): reuse a selection to gather another per-point
    feature tensor (RawToFeatures combiner => plain features)."""

    def forward(self, feature_map, selection):
        feats = feature_map[1]
        return GatherRowsFunction.run(feats, selection.sel_pt)


class SparseMaskPredictor(nn.Module):
    """Consumer of the crop selection (reference SparseMaskPredictor, model.py:859-882; format of
    `split_select_nd`, utils/basic_functions.py:177-216): for every box pick the logit column of its
    predicted class, sigmoid, and scatter back to a per-sample [n_boxes_i, P_i] point mask.  The reference
    splits the dense [BB, P] `is_inside` matrix block-diagonally on the CPU and loops over boxes in Python;
    here the (box, point) CSR produced by the crop kernel is used directly on the device."""

    def __init__(self, num_valid=0):
        super().__init__()
        self.num_valid = num_valid

    @torch.no_grad()
    def forward(self, mask_logits, selection, class_indices):
        """mask_logits [sum n_b, classes]; selection: CropSelection; class_indices int64 [BB] (device or CPU).
        Returns a list over samples of float masks [n_boxes_i, P_i]."""
        dev = mask_logits.device
        box_ptr = selection.box_ptr.long()
        n_rows = mask_logits.shape[0]
        cls = class_indices.to(dev).long()
        valid = cls >= 0
        if self.num_valid:
            valid &= cls < self.num_valid
        rows = torch.arange(n_rows, device=dev)
        box_of_row = torch.searchsorted(box_ptr, rows, right=True) - 1
        c = cls.clamp(min=0)[box_of_row]
        val = torch.sigmoid(mask_logits.detach()[rows, c]) * valid[box_of_row].to(mask_logits.dtype)
        counts, splits = selection.bbox_sample_count, selection.batch_splits
        pt = selection.sel_pt.long()
        out, b0, p0 = [], 0, 0
        for nb, npts in zip(counts, splits):
            m = torch.zeros((nb, npts), dtype=mask_logits.dtype, device=dev)
            if nb:
                lo, hi = int(box_ptr[b0]), int(box_ptr[b0 + nb])
                m[box_of_row[lo:hi] - b0, pt[lo:hi] - p0] = val[lo:hi]
            out.append(m)
            b0 += nb
            p0 += npts
        return out


# ------------------------------------------------------------------------------- loss-side consumers of the selection
def split_select_nd(tensor, splits_dims, dims=None):
    """Diagonal blocks of `tensor` cut by a table of splits (reference utils/basic_functions.py:177-216): column j of the
    table gives, per listed dimension, the length of block j; returns [tensor[block_j along every listed dim] for j]."""
    table = splits_dims if isinstance(splits_dims, torch.Tensor) else torch.stack(list(splits_dims))
    if table.dim() != 2:
        raise ValueError("splits must form a [n_dims, n_blocks] table")
    dims = list(range(table.shape[0])) if dims is None else list(dims)
    if len(dims) != table.shape[0]:
        raise ValueError("one row of splits per dimension")
    stops = table.cumsum(1)
    for d, row in zip(dims, stops):
        if tensor.shape[d] != (int(row[-1]) if row.numel() else 0):
            raise ValueError("splits of dimension %d do not sum to its length" % d)
    starts = (stops - table).tolist()
    stops = stops.tolist()
    out = []
    for j in range(table.shape[1]):
        index = [slice(None)] * tensor.dim()
        for k, d in enumerate(dims):
            index[d] = slice(starts[k][j], stops[k][j])
        out.append(tensor[tuple(index)])
    return out


class LossFilter(nn.Module):
    """reference LossFilter (model.py:1017-1032): keep boxes whose best overlap reaches the positive threshold (or falls
    below the negative one: those are associated with -1)."""

    def __init__(self, positive_threshold, negative_threshold=0):
        super().__init__()
        self.positive_threshold, self.negative_threshold = positive_threshold, negative_threshold

    def forward(self, max_overlap, argmax_overlap):
        keep = max_overlap >= self.positive_threshold
        assoc = argmax_overlap
        if self.negative_threshold:
            negative = max_overlap < self.negative_threshold
            keep = keep | negative
            assoc = torch.where(negative, torch.full_like(assoc, -1), assoc)
        return keep, assoc[keep]


def _nest(flat_list, counts):
    out, i = [], 0
    for c in counts:
        out.append(list(flat_list[i:i + c]))
        i += c
    return out


class SparseMaskLossSelector(nn.Module):
    """reference SparseMaskLossSelector (model.py:1152-1227): for every kept box the logit column of its ground-truth label
    at the box's points, and the ground-truth mask of its associated instance at the same points.  The reference splits the
    dense [BB, P] `is_inside` block-diagonally on the CPU and indexes per box in Python; here the (box, point) CSR of the
    crop (`CropSelection.box_ptr / sel_pt`) yields both as ONE gather each on the device.  Returns the reference's nested
    lists (views of the flat tensors); the flat form is kept in `.flat` = (logits [M], targets [M], per-box lengths,
    labels [boxes]) for `losses.MaskLoss.forward_flat`."""

    def __init__(self, positive_threshold):
        super().__init__()
        self.loss_filter = LossFilter(positive_threshold, 0)
        self.flat = None

    def forward(self, mask_scores, selection, class_selector_description_list, pred_gt_max_argmax_tuple_list,
                gt_labels_list, gt_masks_list):
        dev = mask_scores.device
        counts = [int(c) for c in selection.bbox_sample_count]
        splits = [int(v) for v in selection.batch_splits]
        if class_selector_description_list is None:
            kept = [self.loss_filter(mx, am) for _, _, mx, am in pred_gt_max_argmax_tuple_list]
            keep_list, assoc_list = [k for k, _ in kept], [a for _, a in kept]
        else:
            keep_list = [torch.ones(c, dtype=torch.bool) for c in counts]
            assoc_list = [d.gt_association for d in class_selector_description_list]
        selected_labels = [gl[a.to(gl.device)] for gl, a in zip(gt_labels_list, assoc_list)]
        kept_per_sample = [int(k.sum()) for k in keep_list]
        keep = torch.cat([k.to(dev) for k in keep_list]) if keep_list else torch.zeros(0, dtype=torch.bool, device=dev)
        if keep.numel() != sum(counts):
            raise RuntimeError("one overlap entry per box expected")
        box_ptr = selection.box_ptr.to(dev).long()
        boxes = keep.nonzero().squeeze(1)
        lens = (box_ptr[1:] - box_ptr[:-1])[boxes]
        lens_host = lens.tolist()
        nk, M = len(lens_host), sum(lens_host)
        seg = torch.repeat_interleave(torch.arange(nk, device=dev), lens, output_size=M)
        first = torch.cumsum(lens, 0) - lens
        rows = box_ptr[:-1][boxes][seg] + (torch.arange(M, device=dev) - first[seg])
        labels = torch.cat([l.to(dev) for l in selected_labels]).long() if selected_labels else boxes
        logits = mask_scores[rows, labels[seg]]
        # ground truth: sample of every kept box, its instance, the point's index inside the sample
        n_s = len(counts)
        sample_of_box = torch.repeat_interleave(torch.arange(n_s), torch.tensor(counts, dtype=torch.long)).to(dev)[boxes]
        p0 = torch.tensor([0] + splits[:-1], dtype=torch.long).cumsum(0).to(dev)
        psize = torch.tensor(splits, dtype=torch.long, device=dev)
        gsz = [int(m.shape[0]) * int(m.shape[1]) for m in gt_masks_list]
        gbase = torch.tensor([0] + gsz[:-1], dtype=torch.long).cumsum(0).to(dev)
        gflat = torch.cat([m.to(dev).reshape(-1) for m in gt_masks_list]) if gt_masks_list else keep[:0]
        assoc = torch.cat([a.to(dev) for a in assoc_list]).long() if assoc_list else boxes
        sb = sample_of_box[seg]
        local_pt = selection.sel_pt.to(dev).long()[rows] - p0[sb]
        targets = gflat[gbase[sb] + assoc[seg] * psize[sb] + local_pt]
        self.flat = (logits, targets, lens_host, labels)
        pred = _nest(torch.split(logits, lens_host), kept_per_sample)
        gt = _nest(torch.split(targets, lens_host), kept_per_sample)
        return pred, gt, selected_labels
