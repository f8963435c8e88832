"""GPU parity: mask-crop kernels vs the reference's own roi_cut (golden) and the sparse class /
mask networks vs the oracle."""
import os

import numpy as np
import pytest
import torch

import scn_oracle as O
from scn_oracle import roi_ref
from sparse_rcnn_b200 import networks
from sparse_rcnn_b200.synthetic import make_batch, make_boxes
from tests.util import reinit_by_name, rel_err

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["crop_raw", "crop_stride4"])
def test_crop_kernel_matches_reference_golden(cuda, name):
    from sparse_rcnn_b200 import roi, scn
    g = torch.load(os.path.join(G, name + ".pt"), weights_only=False)
    coords, feats = g["coords"].long(), g["feats"]
    splits = torch.bincount(coords[:, 3]).tolist()
    scene = (coords, feats.to(cuda), g["size"], len(splits), splits)
    cut = roi.SparseRoiCut(scn, raw_scene=True, clip_boxes=g["clip"], resize_boxes=g["resize"], combine="raw")
    (new_keys, new_feats, size, n_boxes), sel = cut(scene, g["boxes"])
    inside = sel.is_inside(cpu=True).numpy()
    assert np.array_equal(np.packbits(inside, axis=1), g["inside_packed"])           # bit-exact selection
    from sparse_rcnn_b200 import _lib
    loc = torch.empty((new_keys.numel(), 4), dtype=torch.int64, device=cuda)
    _lib.call("scn_unpack_keys", new_keys.data_ptr(), new_keys.numel(), loc.data_ptr(), 0)
    torch.cuda.synchronize()
    assert torch.equal(loc.cpu(), g["new_coords"].long())                             # (box, point) order, xyz absolute
    assert torch.equal(new_feats[:32].cpu(), g["new_feats_head"])
    assert n_boxes == len(g["assoc"])
    assert sel.bbox_sample_count == g["counts"]


def test_crop_empty_and_ragged(cuda):
    from sparse_rcnn_b200 import roi, scn
    coords, feats, size, bs, splits = make_batch(3, 2, spatial_size=(32, 32, 16), room=(22, 22, 11),
                                                 room_offset=(4, 4, 1), n_furniture=1, density=1.0)
    scene = (coords, feats.to(cuda), size, bs, splits)
    boxes = [torch.zeros(0, 2, 3), torch.tensor([[[0., 0, 0], [32, 32, 16]], [[100., 100, 100], [101, 101, 101]]]),
             torch.tensor([[[5.5, 5.5, 0.2], [9.1, 20.0, 8.0]]])]
    out, sel = roi.SparseRoiCut(scn, raw_scene=True)(scene, boxes)
    ref, rsel = roi_ref.OracleRoiCut(O, raw_scene=True)((coords, feats, size, bs, splits), boxes)
    assert torch.equal(sel.is_inside(cpu=True), rsel.is_inside())
    assert out.batch_size() == 3 == ref.batch_size()
    assert torch.equal(out.get_spatial_locations(), ref.get_spatial_locations())
    assert rel_err(out.features, ref.features) <= 1e-6
    # no boxes at all
    out0, sel0 = roi.SparseRoiCut(scn, raw_scene=True, combine="features")(scene, [torch.zeros(0, 2, 3)] * 3)
    assert out0.shape == (0, feats.shape[1]) and sel0.total == 0


@pytest.mark.parametrize("precision,tol", [("fp32", 5e-5), ("tf32", 3e-3)])
def test_class_and_mask_networks(cuda, precision, tol):
    from sparse_rcnn_b200 import roi, scn
    scn.set_precision(precision)
    ocut = lambda **kw: roi_ref.OracleRoiCut(O, **kw)
    gcut = lambda **kw: roi.SparseRoiCut(scn, **kw)
    nets_o = [networks.FeatureExtractor(O), networks.ClassNetwork(O, ocut), networks.SparseMaskNetwork(O, ocut)]
    nets_g = [networks.FeatureExtractor(scn), networks.ClassNetwork(scn, gcut), networks.SparseMaskNetwork(scn, gcut)]
    for a, b in zip(nets_o, nets_g):
        reinit_by_name(a).eval()
        b.load_state_dict(a.state_dict())
        b.to(cuda).eval()
    data = make_batch(2, 9, spatial_size=(64, 64, 32), room=(44, 44, 22), room_offset=(8, 8, 2), n_furniture=3)
    boxes = make_boxes(data[0], 5, 2, (64, 64, 32))
    gdata = (data[0], data[1].to(cuda), *data[2:])
    with torch.no_grad():
        oo, og = nets_o[0](data), nets_g[0](gdata)
        co, cso = nets_o[1](oo[3], boxes)
        cg, csg = nets_g[1](og[3], boxes)
        assert torch.equal(csg.is_inside(cpu=True), cso.is_inside())
        assert rel_err(cg, co) <= tol, rel_err(cg, co)
        mo, mso = nets_o[2](data, oo[5], boxes)
        mg, msg = nets_g[2](gdata, og[5], boxes)
        assert torch.equal(msg.is_inside(cpu=True), mso.is_inside())
        assert mg.shape == mo.shape and rel_err(mg, mo) <= tol, rel_err(mg, mo)


def test_global_pool_segment_mean(cuda):
    from sparse_rcnn_b200 import scn
    from tests.util import make_pair, random_scene
    coords, feats, size = random_scene(3, n_samples=4, channels=9)
    coords = coords[coords[:, 3] != 2]            # an empty sample in the middle
    feats = feats[: len(coords)]
    to, tg = make_pair(scn, coords, feats, size, cuda, batch_size=6)
    po, pg = networks.SparseGlobalPool(O)(to), networks.SparseGlobalPool(scn)(tg)
    assert pg.shape == (6, 9) and rel_err(pg, po) <= 1e-6
    x = tg.features.clone().requires_grad_(True)
    y = networks.SparseGlobalPool(scn)(scn.SparseConvNetTensor(x, tg.metadata, size))
    w = torch.randn_like(y)
    (y * w).sum().backward()
    xo = to.features.clone().requires_grad_(True)
    (networks.SparseGlobalPool(O)(O.SparseConvNetTensor(xo, to.metadata, size)) * w.cpu()).sum().backward()
    assert rel_err(x.grad, xo.grad) <= 1e-6


def test_mask_predictor_matches_reference_algorithm(cuda):
    """Device-side consumer of the crop CSR vs the reference SparseMaskPredictor algorithm (model.py:859-882)
    restated with the dense is_inside matrix on the CPU."""
    from sparse_rcnn_b200 import roi, scn
    coords, feats, size, bs, splits = make_batch(2, 4, spatial_size=(32, 32, 16), room=(22, 22, 11),
                                                 room_offset=(4, 4, 1), n_furniture=1, density=1.0)
    boxes = make_boxes(coords, 3, 5, (32, 32, 16))
    scene = (coords, feats.to(cuda), size, bs, splits)
    _, sel = roi.SparseRoiCut(scn, raw_scene=True, combine="features")(scene, boxes)
    torch.manual_seed(0)
    logits = torch.randn(sel.total, 18, device=cuda)
    cls = torch.tensor([2, 17, -1, 0, 5, 9])
    got = roi.SparseMaskPredictor(18)(logits, sel, cls)
    inside = sel.is_inside(cpu=True)
    lg = logits.cpu()
    row, b0, p0 = 0, 0, 0
    for s, (nb, npts) in enumerate(zip(sel.bbox_sample_count, sel.batch_splits)):
        ref = torch.zeros(nb, npts)
        for b in range(nb):
            m = inside[b0 + b, p0:p0 + npts]
            n = int(m.sum())
            c = int(cls[b0 + b])
            ref[b, m] = torch.sigmoid(lg[row:row + n, c]) if 0 <= c < 18 else 0.0
            row += n
        assert torch.allclose(got[s].cpu(), ref, atol=1e-6)
        b0 += nb
        p0 += npts
