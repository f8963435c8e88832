"""Generate tests/golden/*.pt from the UNMODIFIED reference (/root/reference) -- run in the build
container only (the reference is not on the GPU box):

    python oracle/make_golden.py

1. crop_*.pt : inputs + outputs of the reference's own pure-torch `roi_cut` / `BBoxTransformerSlice`
   (importable without SparseConvNet once any `sparseconvnet` module is on the path).
2. ref_graph.pt : state_dict keys/shapes + outputs of the reference's FeatureExtractor,
   SegmentationNetwork, ClassNetwork and SparseMaskNetwork (ndsis/modules/model.py, built with the
   restated run.py configuration) running on the CPU oracle backend aliased as `sparseconvnet`.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, "/root/reference")

import scn_oracle

sys.modules["sparseconvnet"] = scn_oracle

import numpy as np
import torch

from sparse_rcnn_b200.synthetic import make_batch, make_boxes
from tests.util import reinit_by_name

OUT = os.path.join(ROOT, "tests", "golden")


def small_scene(n_scenes=2, seed=5):
    return make_batch(n_scenes, seed, spatial_size=(64, 64, 32), room=(44, 44, 22), room_offset=(8, 8, 2), n_furniture=3)


def tiny_scene(n_scenes=2, seed=5):
    return make_batch(n_scenes, seed, spatial_size=(32, 32, 16), room=(22, 22, 11), room_offset=(4, 4, 1),
                      n_furniture=2, density=1.2)


def golden_crop():
    from ndsis.modules.roi_select_sparse import roi_cut
    from ndsis.modules.roi_select_bbox_transform import BBoxTransformerSlice
    coords, feats, size, bs, splits = tiny_scene()
    boxes = make_boxes(coords, 6, 1, (32, 32, 16))
    boxes[1][0] = torch.tensor([[40., 40, 20], [45, 45, 25]])       # box containing no point
    for name, clip, resize in (("crop_raw", False, None), ("crop_stride4", True, [4, 4, 4])):
        c, f, s = coords, feats, size
        if resize is not None:       # level-2-like map: unique coarse coordinates per sample
            cc = coords.clone()
            cc[:, :3] //= 4
            _, idx = np.unique(scn_oracle.rules.pack_keys(cc.numpy()), return_index=True)
            idx = np.sort(idx)
            c, f, s = cc[idx], feats[idx], size // 4
        tr = BBoxTransformerSlice(clip=clip, resize=resize)
        bt, counts, assoc = tr(boxes, s)
        nc, nf, inside = roi_cut(c, f, bt, assoc)
        torch.save(dict(coords=c.to(torch.int16), feats=f, size=s, boxes=boxes, clip=clip, resize=resize, box_tensor=bt,
                        counts=counts, assoc=assoc, new_coords=nc.to(torch.int16), new_feats_sum=float(nf.double().sum()),
                        new_feats_head=nf[:32].clone(),
                        inside_packed=np.packbits(inside.numpy(), axis=1), inside_shape=tuple(inside.shape)),
                   os.path.join(OUT, name + ".pt"))
        print(name, "selected", len(nc), "of", tuple(inside.shape))


def reference_configs():
    from ndsis.modules.model import FeatureLevelDescriptor as FLD
    chans = [32, 48, 64, 80, 96, 112]
    desc = [FLD('B', 32, dict(stride=1, drop_input_relu=True))] + [FLD('B', c) for c in chans[1:]]
    fe = dict(num_dims=3, sparse=True, input_channels=6, network_description=desc, class_output_index=-4,
              num_dilations=5, num_units=2, bottleneck_divisor=0, stride=2, maxpool=False, relu_first=True,
              main_path_relu=False, bottleneck_groups=1, batchnorm=False, use_residuals=True, drop_input_relu=True)
    unet = dict(use_residuals=True, num_units=2, bottleneck_divisor=0, groups=1, main_path_relu=False,
                relu_first=True, batchnorm=False, concat=True, min_channels=16)
    common_class = dict(main_path_relu=False, relu_first=True, bottleneck_divisor=0, drop_input_relu=True,
                        make_dense=False, num_units=1)
    cls = dict(input_network_description=[FLD('B', 32, {**common_class, 'stride': 1})],
               output_network_description=[FLD('B', 64, {**common_class, 'stride': 2}),
                                           FLD('B', 128, {**common_class, 'stride': 2})],
               linear_channels=[64], num_classes=18, raw_scene=False, cut_shape=None,
               pooling_function_or_none=torch.mean, relu_after_pooling=True, positive_threshold=0.5,
               negative_threshold=0, selection_tuple=(32, 0, True))
    common_mask = dict(use_residuals=True, main_path_relu=False, relu_first=True, bottleneck_divisor=0,
                       drop_input_relu=True, make_dense=False)
    mask = dict(use_raw_features=True, use_unet_features=True, use_skip_features=False, internal_unet=True,
                unet_params=unet,
                input_network_description=[FLD('B', 16, {**common_mask, 'num_units': 2})],
                output_network_description=[FLD('I')] + [FLD('B', c, {**common_mask, 'num_units': 2, 'stride': 2})
                                                         for c in (32, 48, 64)],
                channel_list=[32, 18], positive_threshold=0.5, selection_tuple=(24, 0, True))
    return fe, unet, cls, mask


def golden_graph():
    from ndsis.modules.model import FeatureExtractor, SegmentationNetwork, ClassNetwork, SparseMaskNetwork
    fe_p, unet_p, cls_p, mask_p = reference_configs()
    torch.manual_seed(0)
    fe = FeatureExtractor(**fe_p, include_unet=True, unet_params=unet_p).eval()
    seg = SegmentationNetwork(3, True, fe.unet_strides, fe.unet_channels, 20).eval()
    cls = ClassNetwork(3, fe.class_sparse, fe.class_channels, fe.class_stride, **cls_p).eval()
    mask = SparseMaskNetwork(3, True, 6, fe.skip_connection_channels, fe.skip_connection_strides, fe.unet_channels,
                             fe.unet_strides, **mask_p).eval()
    for m in (fe, seg, cls, mask):
        reinit_by_name(m)
    data = small_scene(2, 9)
    boxes = make_boxes(data[0], 5, 2, (64, 64, 32))
    with torch.no_grad():
        scene_size, bs, anchors, class_map, inter, unet = fe(data)
        seg_out = seg(unet, data)
        cls_out, cls_sel, _ = cls(class_map, boxes, None)
        mask_out, mask_sel, _ = mask(data, inter, unet, boxes, None)
    g = dict(
        state=dict(fe={k: tuple(v.shape) for k, v in fe.state_dict().items()},
                   seg={k: tuple(v.shape) for k, v in seg.state_dict().items()},
                   cls={k: tuple(v.shape) for k, v in cls.state_dict().items()},
                   mask={k: tuple(v.shape) for k, v in mask.state_dict().items()}),
        weight_sums={n: float(sum(v.double().sum() for v in m.state_dict().values()))
                     for n, m in (("fe", fe), ("seg", seg), ("cls", cls), ("mask", mask))},
        scene_seed=9, box_seed=2,
        level_rows=[t.features.shape[0] for t in inter],
        unet_last_sum=float(unet[-1].features.double().sum()), unet_last_absmax=float(unet[-1].features.abs().max()),
        unet_last_head=unet[-1].features[:64].clone(),
        seg_head=seg_out[:64].clone(), seg_sum=float(seg_out.double().sum()),
        cls_out=cls_out.clone(), mask_rows=mask_out.shape[0], mask_head=mask_out[:64].clone(),
        mask_sum=float(mask_out.double().sum()),
        cls_inside_count=int(cls_sel[0].sum()), mask_inside_count=int(mask_sel[0].sum()))
    torch.save(g, os.path.join(OUT, "ref_graph.pt"))
    print("ref graph: levels", g["level_rows"], "mask rows", g["mask_rows"], "cls", tuple(cls_out.shape))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    golden_crop()
    golden_graph()
