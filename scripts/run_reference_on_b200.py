"""The UNMODIFIED reference modules (ndsis/modules/model.py: FeatureExtractor, SegmentationNetwork, ClassNetwork,
SparseMaskNetwork) running on the B200 backend aliased as `sparseconvnet`, forward on CUDA on the
scene of tests/golden/ref_graph.pt (which holds the same modules' outputs on the CPU oracle).

    python scripts/run_reference_on_b200.py <dir that contains ndsis/> [out.pt]

The reference package is not part of this repository; tests/test_gpu_reference_dropin.py looks for it in /root/reference
(build container) and baseline/_ref (git-ignored copy that travels to the GPU box)."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), sys.argv[1]]

import torch

# the restated run.py configuration lives in oracle/make_golden.py (test infrastructure); importing it aliases the ORACLE as
# sparseconvnet, so the B200 backend is installed afterwards and before anything imports ndsis
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "oracle", "make_golden.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)
import sparse_rcnn_b200
scn = sparse_rcnn_b200.install_as_sparseconvnet()
assert "ndsis" not in sys.modules
from ndsis.modules.model import FeatureExtractor, SegmentationNetwork, ClassNetwork, SparseMaskNetwork
import ndsis.modules.module_factory as mf
assert mf.scn is scn, "the reference must be bound to the B200 backend"

from sparse_rcnn_b200 import _lib, networks
from sparse_rcnn_b200.scn import layers
from sparse_rcnn_b200.synthetic import make_boxes
from tests.util import reinit_by_name

dev = torch.device("cuda:0")
scn.set_precision(os.environ.get("SCN_PRECISION", "tf32"))
fe_p, unet_p, cls_p, mask_p = mg.reference_configs()
fe = FeatureExtractor(**fe_p, include_unet=True, unet_params=unet_p).eval()
seg = SegmentationNetwork(3, True, fe.unet_strides, fe.unet_channels, 20).eval()
cls = ClassNetwork(3, fe.class_sparse, fe.class_channels, fe.class_stride, **cls_p).eval()
mask = SparseMaskNetwork(3, True, 6, fe.skip_connection_channels, fe.skip_connection_strides, fe.unet_channels,
                         fe.unet_strides, **mask_p).eval()
for m in (fe, seg, cls, mask):
    reinit_by_name(m)
    m.to(dev)
n_fused = sum(1 for m in list(fe.modules()) + list(mask.modules()) + list(cls.modules())
              if isinstance(m, layers.Sequential) and m._plan() is not None)
g = torch.load(os.path.join(ROOT, "tests", "golden", "ref_graph.pt"), weights_only=False)
data = mg.small_scene(2, g["scene_seed"])
boxes = make_boxes(data[0], 5, g["box_seed"], (64, 64, 32))
coords, feats = data[0], data[1].to(dev)
gdata = (coords, feats) + tuple(data[2:])


def forward():
    with torch.no_grad():
        scene_size, bs, anchors, class_map, inter, unet = fe(gdata)
        seg_out = seg(unet, gdata)
        cls_out, cls_sel, _ = cls(class_map, boxes, None)
        mask_out, mask_sel, _ = mask(gdata, inter, unet, boxes, None)
    return inter, unet, seg_out, cls_out, cls_sel, mask_out, mask_sel


res = {}
forward()      # first touch packs the weight images (counted launches)
for fuse in (True, False):
    layers.FUSE["residual"] = fuse
    l0 = int(_lib.raw("scn_launch_count")())
    inter, unet, seg_out, cls_out, cls_sel, mask_out, mask_sel = forward()
    torch.cuda.synchronize()
    res[fuse] = dict(launches=int(_lib.raw("scn_launch_count")()) - l0, level_rows=[t.features.shape[0] for t in inter],
                     unet_last=unet[-1].features.cpu(), seg=seg_out.cpu(), cls=cls_out.cpu(), mask=mask_out.cpu(),
                     cls_inside=int(cls_sel[0].sum()), mask_inside=int(mask_sel[0].sum()),
                     locations=unet[-1].get_spatial_locations().clone())
layers.FUSE["residual"] = True

# the repository's own mirror of the same graph, same weights, same input
mfe = networks.FeatureExtractor(scn).eval()
reinit_by_name(mfe)
mfe.to(dev)
with torch.no_grad():
    m_unet = mfe(gdata)[5][-1].features.cpu()
out = dict(fused=res[True], plain=res[False], fused_sequentials=n_fused, mirror_unet_last=m_unet,
           state_keys={k: tuple(v.shape) for k, v in fe.state_dict().items()})
torch.save(out, sys.argv[2] if len(sys.argv) > 2 else "/tmp/reference_on_b200.pt")
print("OK fused_sequentials=%d launches fused=%d plain=%d" % (n_fused, res[True]["launches"], res[False]["launches"]))
