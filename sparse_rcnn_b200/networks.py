"""Host-side mirror of the reference's sparse network assembly for the hot path.

/root/reference is not present on the GPU box, so the layer graphs that
`ndsis/modules/module_factory.py` + `model.py` build for the shipped sparse configuration
(scannet_config/run.py:515-815) are restated here, parameterised by the `scn` namespace
(the B200 backend `sparse_rcnn_b200.scn`, or any other object exposing the same names --
tests pass the CPU oracle).  Module nesting and attribute names follow the reference tree so
that `state_dict()` keys are identical (tests/test_golden.py::test_graph_mirror_has_reference_state_dict checks this against the
unmodified reference when it is available).

Graph facts restated (with the reference line that produces them):
  residual unit  : x + SubM3(ReLU(SubM3(ReLU(x))))          module_factory.py:127-183 (relu_first,
                   main_path_relu=False, bottleneck_divisor=0), shortcut = Identity | NetworkInNetwork
  encoder level  : [SubM 1^3 | Convolution 2^3/s2] + num_units residual units   :438-530
  decoder level  : ReLU, Deconvolution 2^3/s2, JoinTable(up, skip), NetworkInNetwork, units  :533-578
"""
import numpy as np
import torch
import torch.nn as nn


# ----------------------------------------------------------------------------- containers
class Interims(nn.Sequential):
    """Sequential that returns every intermediate result (reference SequentialInterims)."""

    def forward(self, x):
        outs = []
        for m in self._modules.values():
            x = m(x)
            outs.append(x)
        return outs


class SkipReunite(nn.Module):
    """One decoder level (reference SkipConnectionReuniter, custom_container.py:68-83)."""

    def __init__(self, input_stage, combiner, channel_changer, output_stage):
        super().__init__()
        self.input_stage, self.combiner = input_stage, combiner
        self.channel_changer, self.output_stage = channel_changer, output_stage

    def forward(self, x, skip):
        return self.output_stage(self.channel_changer(self.combiner([self.input_stage(x), skip])))


class ReuniteInterims(nn.Module):
    def __init__(self, *mods):
        super().__init__()
        self.module_list = nn.ModuleList(mods)

    def forward(self, x, skips):
        assert len(skips) == len(self.module_list)
        outs = []
        for skip, m in zip(skips, self.module_list):
            x = m(x, skip)
            outs.append(x)
        return outs


class ReuniteLast(nn.Module):
    """reference InverseSequentialInterims: only the final decoder output."""

    def __init__(self, *mods):
        super().__init__()
        self.module_list = nn.ModuleList(mods)

    def forward(self, x, skips):
        assert len(skips) == len(self.module_list)
        for skip, m in zip(skips, self.module_list):
            x = m(x, skip)
        return x


class UNet(nn.Module):
    def __init__(self, down, up):
        super().__init__()
        self.downsampling_layer, self.upsampling_layer = down, up

    def _executor(self, x):
        """Native executor program for this U-Net on the B200 backend (same rule as FeatureExtractor._executor)."""
        if not hasattr(getattr(x, "metadata", None), "prebuild"):
            return None
        from . import executor
        if not (executor.ENABLED["unet"] and FUSE["residual"]):
            return None
        if "_program" not in self.__dict__:
            self.__dict__["_program"] = executor.UNetProgram.compile(list(self.downsampling_layer),
                                                                     list(self.upsampling_layer.module_list))
        return self.__dict__["_program"]

    def forward(self, x):
        prog = self._executor(x)
        if prog is not None:
            inter, ups = prog.run(x)
            return ups if isinstance(self.upsampling_layer, ReuniteInterims) else ups[-1]
        *skip, out = self.downsampling_layer(x)
        return self.upsampling_layer(out, skip[::-1])


# ----------------------------------------------------------------------------- blocks
def _fuse_switch():
    from .scn import layers
    return layers.FUSE


FUSE = _fuse_switch()          # the B200 backend's switch (scn/layers.py): False runs the unfused scn.* module graph


def residual_unit(scn, cin, cout):
    """module_factory.py:127-183 (relu_first, main_path_relu=False, bottleneck_divisor=0).  On the B200 backend
    `scn.Sequential` recognises this module pattern itself and runs it fused -- also when the unmodified reference builds it."""
    shortcut = scn.Identity() if cin == cout else scn.NetworkInNetwork(cin, cout, True)
    inner = scn.Sequential(
        scn.ReLU(), scn.SubmanifoldConvolution(3, cin, cout, 3, True),
        scn.ReLU(), scn.SubmanifoldConvolution(3, cout, cout, 3, True))
    return scn.Sequential(scn.ConcatTable(shortcut, inner), scn.AddTable())


def unit_stage(scn, channels, num_units):
    return scn.Sequential(*[residual_unit(scn, channels, channels) for _ in range(num_units)])


def encoder_level(scn, cin, cout, stride, num_units):
    if stride == 1:
        entry = scn.Sequential(scn.SubmanifoldConvolution(3, cin, cout, 1, True))
    else:
        entry = scn.Sequential(scn.Convolution(3, cin, cout, (stride,) * 3, (stride,) * 3, True))
    return scn.Sequential(entry, unit_stage(scn, cout, num_units))


def decoder_level(scn, cin, skip, cout, stride, num_units):
    up = scn.Sequential(scn.ReLU(), scn.Deconvolution(3, cin, cout, (stride,) * 3, (stride,) * 3, True))
    changer = scn.NetworkInNetwork(cout + skip, cout, True)
    return SkipReunite(up, scn.JoinTable(), changer, unit_stage(scn, cout, num_units))


class InputStage(nn.Module):
    """reference SparseInputStage + CustomInputLayer(mode=4) (model.py:248-258,
    custom_operations.py:62-86)."""

    def __init__(self, scn, mode=4):
        super().__init__()
        self.scn, self.mode = scn, mode
        self.prefetcher = None      # optional scn.GeometryPrefetcher (B200 backend): geometry of this batch built ahead
        self.ready = {}             # id(coords) -> Metadata built by build_ahead()

    def build_ahead(self, data, n_levels, device):
        """Build the geometry (voxel hash, level pyramid, neighbour maps) of an UPCOMING batch now, on the current stream.
        Called between the forward and the backward of the current step, the builder's host round trips wait for the
        forward's kernels to drain (the GPU is busy) instead of idling the GPU at the head of the next step, and the next
        forward is issued without any synchronisation.  Same kernels, same order of results."""
        coords, _, spatial_size, batch_size = data[:4]
        coords = coords.long()
        if not len(coords) or id(coords) in self.ready or getattr(self.scn, "BACKEND", "") != "b200-cuda":
            return
        md = self.scn.Metadata(3)
        md.set_input(torch.as_tensor(spatial_size, dtype=torch.long), coords, batch_size, self.mode, device)
        md.prebuild(n_levels)
        md._prebuilt_for = coords
        self.ready[id(coords)] = md

    def forward(self, data):
        coords, feats, spatial_size, batch_size = data[:4]
        spatial_size = torch.as_tensor(spatial_size, dtype=torch.long)
        if not len(coords):
            return spatial_size, batch_size, None
        coords = coords.long()      # int64 already in the reference contract (data.py:95-98): the same tensor object
        md = self.ready.pop(id(coords), None)       # geometry built ahead of time for exactly this tensor (build_ahead)
        if md is None and self.prefetcher is not None:
            md = self.prefetcher.take(coords)
        if md is None:
            md = self.scn.Metadata(3)
        fn = self.scn.ioLayers.InputLayerFunction
        f = getattr(fn, "run", fn.apply)(3, md, spatial_size, coords, feats, batch_size, self.mode)
        t = self.scn.SparseConvNetTensor(features=f, metadata=md, spatial_size=spatial_size)
        return spatial_size, t.batch_size(), t


class FeatureExtractor(nn.Module):
    """Sparse U-Net feature extractor (reference FeatureExtractor, model.py:261-446, built by
    run.py:552-600 with include_unet=True).  encoder channels [32,48,64,80,96,112]."""

    def __init__(self, scn, input_channels=6, channels=(32, 48, 64, 80, 96, 112), num_units=2, include_unet=True,
                 class_output_index=-4, min_unet_channels=16):
        super().__init__()
        self.scn = scn
        self.input_channels = input_channels
        self.channels = list(channels)
        self.input_stage = InputStage(scn)
        levels, cin = [], input_channels
        for i, c in enumerate(channels):
            levels.append(encoder_level(scn, cin, c, 1 if i == 0 else 2, num_units))
            cin = c
        self.main_network = Interims(*levels)
        self.class_output_index = class_output_index
        self.include_unet = include_unet
        main_channels = [input_channels] + self.channels
        self.class_channels = main_channels[class_output_index]
        self.class_stride = np.array([2 ** max(len(main_channels) + class_output_index - 1, 0)] * 3) \
            if class_output_index < 0 else np.array([2 ** max(class_output_index - 1, 0)] * 3)
        if include_unet:
            rev = self.channels[::-1]
            ups, c = [], rev[0]
            for skip in rev[1:]:
                out = max(skip, min_unet_channels)
                ups.append(decoder_level(scn, c, skip, out, 2, num_units))
                c = out
            self.unet = ReuniteInterims(*ups)
            self.unet_channels = [max(s, min_unet_channels) for s in rev[1:]]
        else:
            self.unet = None

    def _executor(self):
        """Compiled layer table of main_network + unet for the native executor, or None (other backends, switched off,
        or a module tree the compiler does not recognise)."""
        if getattr(self.scn, "BACKEND", "") != "b200-cuda":
            return None
        from . import executor
        if not (executor.ENABLED["unet"] and FUSE["residual"]):
            return None
        if "_program" not in self.__dict__:
            ups = list(self.unet.module_list) if self.include_unet else []
            self.__dict__["_program"] = executor.UNetProgram.compile(list(self.main_network), ups)
        return self.__dict__["_program"]

    def forward(self, data):
        scene_size, batch_size, x = self.input_stage(data)
        if hasattr(x.metadata, "prebuild"):         # B200 backend: front-load the rulebook builder's host syncs
            x.metadata.prebuild(len(self.channels) - 1)
        prog = self._executor()
        if prog is not None:
            # B200 backend: encoder + decoder as ONE native call per direction (executor.py / csrc/unet_exec.cu); the same
            # kernels in the same order as the module graph below
            inter, unet_out = prog.run(x)
            if not self.include_unet:
                unet_out = None
        else:
            inter = self.main_network(x)
            unet_out = self.unet(inter[-1], inter[:-1][::-1]) if self.include_unet else None
        extended = [x] + inter
        class_out = extended[self.class_output_index]
        return scene_size, batch_size, [], class_out, inter, unet_out


class SegmentationNetwork(nn.Module):
    """reference SegmentationNetwork (model.py:449-467): SubM 1^3 C->classes + OutputLayer."""

    def __init__(self, scn, channels=32, num_classes=20):
        super().__init__()
        self.channel_changer = scn.SubmanifoldConvolution(3, channels, num_classes, 1, True)
        self.output_layer = scn.OutputLayer(3)

    def forward(self, unet_feature_maps, scene=None):
        return self.output_layer(self.channel_changer(unet_feature_maps[-1]))


def count_parameters(m):
    return sum(p.numel() for p in m.parameters())


# ----------------------------------------------------------------------------- heads on crops
class SparseGlobalPool(nn.Module):
    """Per-box mean over the rows of each sample (reference SparseGlobalPool / split_batch,
    custom_operations.py:24-59, which builds a [B, N] mask and loops in Python).  On the B200
    backend this is one segment-mean kernel over the batch-sorted rows."""

    def __init__(self, scn):
        super().__init__()
        self.scn = scn

    def forward(self, x):
        if getattr(self.scn, "BACKEND", "") == "b200-cuda":
            from .scn.functions import SegmentMeanFunction
            md = x.metadata
            n_seg = md.n_samples
            ptr = md.level(x.spatial_size).batch_ptr(n_seg)
            return SegmentMeanFunction.run(x.features, ptr, n_seg)
        # generic namespace (e.g. the CPU oracle in tests): reference algorithm
        b = x.get_spatial_locations()[:, -1]
        n = x.batch_size()
        out = x.features.new_zeros((n, x.features.shape[1]))
        cnt = torch.bincount(b, minlength=n).clamp(min=1).to(out.dtype)
        out.index_add_(0, b.to(x.features.device), x.features)
        return out / cnt[:, None].to(out.device)


def linear_stack(cin, channel_list, start_relu):
    """reference get_linear_layer_network (module_factory.py:676-693), end_relu=False."""
    layers = [nn.ReLU()] if start_relu else []
    layers.append(nn.Linear(cin, channel_list[0]))
    c = channel_list[0]
    for ch in channel_list[1:]:
        layers += [nn.ReLU(inplace=True), nn.Linear(c, ch)]
        c = ch
    return nn.Sequential(*layers)


class ClassNetwork(nn.Module):
    """Sparse branch of the reference ClassNetwork (model.py:470-569) as configured by
    run.py:640-718: level-2 map (64 ch, stride 4) -> [SubM1 64->32 + 1 unit] -> SparseRoiCut
    (Tensor->Tensor, boxes / stride, clipped) -> [Conv s2 32->64 + unit] -> [Conv s2 64->128 + unit]
    -> per-box mean -> ReLU, Linear 128->64, ReLU, Linear 64->classes."""

    def __init__(self, scn, roi_cut_factory, input_channels=64, stride=4, mid=32, out_channels=(64, 128),
                 linear_channels=(64,), num_classes=18):
        super().__init__()
        self.input_conv_layer = scn.Sequential(encoder_level(scn, input_channels, mid, 1, 1))
        self.roi_getter = roi_cut_factory(raw_scene=False, clip_boxes=True, resize_boxes=[stride] * 3)
        levels, c = [], mid
        for oc in out_channels:
            levels.append(encoder_level(scn, c, oc, 2, 1))
            c = oc
        self.output_conv_layer = scn.Sequential(*levels)
        self.vectorice_layer = SparseGlobalPool(scn)
        self.linear_layer = linear_stack(c, list(linear_channels) + [num_classes], start_relu=True)

    def forward(self, feature_map, roi_bbox):
        x = self.input_conv_layer(feature_map)
        boxes, selection = self.roi_getter(x, roi_bbox)
        if hasattr(boxes.metadata, "prebuild"):
            boxes.metadata.prebuild(len(self.output_conv_layer))
        y = self.output_conv_layer(boxes)
        return self.linear_layer(self.vectorice_layer(y)), selection


class SparseMaskNetwork(nn.Module):
    """reference SparseMaskNetwork (model.py:572-782) as configured by run.py:722-800
    (use_unet_features, use_raw_features, internal U-Net, no skip features):
      input convs on the last U-Net map (SubM1 32->16 + 2 units) -> OutputLayer to points, cat raw
      point features (16+C_in) -> SparseRoiCut on the raw scene with spatial size + 32 (mode 4,
      batch = boxes) -> internal U-Net 22->32->48->64->48->32->22 -> OutputLayer -> Linear 22->32->classes."""

    def __init__(self, scn, roi_cut_factory, input_channels=6, unet_channels=32, mid=16, levels=(32, 48, 64),
                 channel_list=(32, 18), extension=32, min_unet_channels=16):
        super().__init__()
        self.scn = scn
        self.extension = extension
        self.input_conv_layer = scn.Sequential(encoder_level(scn, unet_channels, mid, 1, 2))
        self.point_layer = scn.OutputLayer(3)
        self.output_roi_cut = roi_cut_factory(raw_scene=True, clip_boxes=False, resize_boxes=None)
        c0 = mid + input_channels
        down, c = [scn.Identity()], c0
        for ch in levels:
            down.append(encoder_level(scn, c, ch, 2, 2))
            c = ch
        chans = [c0] + list(levels)
        rev = chans[::-1]
        ups, c = [], rev[0]
        for skip in rev[1:]:
            out = max(skip, min_unet_channels)
            ups.append(decoder_level(scn, c, skip, out, 2, 2))
            c = out
        self.output_conv_layer = UNet(Interims(*down), ReuniteLast(*ups))
        self.final_point_layer = scn.OutputLayer(3)
        self.classes = channel_list[-1]
        self.linear_layer = linear_stack(c, list(channel_list), start_relu=False)

    def forward(self, scene, unet_feature_maps, roi_bbox):
        coords, feats, spatial_size, *other, batch_splits = scene
        x = self.input_conv_layer(unet_feature_maps[-1])
        point_feats = self.point_layer(x)
        combined = torch.cat((point_feats, feats.to(point_feats.device)), dim=-1)
        size = torch.as_tensor(spatial_size, dtype=torch.long) + self.extension
        new_scene = (coords, combined, size, *other, batch_splits)
        boxes_tensor, selection = self.output_roi_cut(new_scene, roi_bbox)
        if boxes_tensor.features.shape[0] == 0:
            return combined.new_zeros((0, self.classes)), selection
        if hasattr(boxes_tensor.metadata, "prebuild"):
            boxes_tensor.metadata.prebuild(len(self.output_conv_layer.downsampling_layer) - 1)
        y = self.output_conv_layer(boxes_tensor)
        return self.linear_layer(self.final_point_layer(y)), selection
