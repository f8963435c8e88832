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
    assert sel._is_inside is None                                                     # built from the CSR on request ...
    inside = sel.is_inside(cpu=True).numpy()
    assert np.array_equal(np.packbits(inside, axis=1), g["inside_packed"])           # bit-exact selection
    cut.dense_selection = True                                                        # ... or written by the crop kernel
    _, sel2 = cut(scene, g["boxes"])
    assert sel2._is_inside is not None and np.array_equal(sel2.is_inside(cpu=True).numpy(), inside)
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
    from tests.util import canon_features
    (kg, fg_), (ko, fo_) = canon_features(out), canon_features(ref)
    assert np.array_equal(kg, ko) and rel_err(fg_, fo_) <= 1e-6       # same voxels, same features (any batch-sorted order)
    loc = out.get_spatial_locations()
    assert bool((loc[:-1, 3] <= loc[1:, 3]).all())
    # no boxes at all
    out0, sel0 = roi.SparseRoiCut(scn, raw_scene=True, combine="features")(scene, [torch.zeros(0, 2, 3)] * 3)
    assert out0.shape == (0, feats.shape[1]) and sel0.total == 0


def l2_err(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / b.norm().clamp(min=1e-30))


@pytest.mark.parametrize("precision,tol,gtol", [("fp32", 5e-5, 5e-3), ("tf32", 3e-3, 0.25), ("fp32 forward, tf32 backward", 5e-5, 2e-2)])
def test_class_and_mask_networks(cuda, precision, tol, gtol):
    """Forward AND backward: logits within the forward tolerance, every parameter gradient of the
    class network, the mask network and the backbone below them in L2 (ReLU masks flip on inputs straddling zero, see
    test_gpu_backbone.py), through the crop's scatter-add over overlapping boxes."""
    from sparse_rcnn_b200 import roi, scn
    fwd, bwd = ("fp32", "tf32") if "," in precision else (precision, precision)      # see test_gpu_baseline_size.py
    scn.set_precision(fwd)
    try:
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
        boxes[0][1] = boxes[0][0] + 1.5                              # heavily overlapping boxes: scatter-add in the crop backward
        gdata = (data[0], data[1].to(cuda), *data[2:])
        oo, og = nets_o[0](data), nets_g[0](gdata)
        co, cso = nets_o[1](oo[3], boxes)
        cg, csg = nets_g[1](og[3], boxes)
        # the class crop selects ROWS of the level-2 tensor: its is_inside columns follow the row order, compare canonically
        from tests.util import canon_order
        assert torch.equal(csg.is_inside(cpu=True)[:, canon_order(og[3].get_spatial_locations())],
                           cso.is_inside()[:, canon_order(oo[3].get_spatial_locations())])
        assert rel_err(cg, co) <= tol, rel_err(cg, co)
        mo, mso = nets_o[2](data, oo[5], boxes)
        mg, msg = nets_g[2](gdata, og[5], boxes)
        assert torch.equal(msg.is_inside(cpu=True), mso.is_inside())
        assert mg.shape == mo.shape and rel_err(mg, mo) <= tol, rel_err(mg, mo)
        gen = torch.Generator().manual_seed(3)
        wc, wm = torch.randn(co.shape, generator=gen), torch.randn(mo.shape, generator=gen) / mo.shape[0] ** 0.5
        ((co * wc).sum() + (mo * wm).sum()).backward()
        scn.set_precision(bwd)
        ((cg * wc.to(cuda)).sum() + (mg * wm.to(cuda)).sum()).backward()
        errs = {}
        for name, a, b in zip(("fe", "cls", "mask"), nets_o, nets_g):
            for (n, po), (_, pg) in zip(a.named_parameters(), b.named_parameters()):
                if po.grad is not None:
                    assert pg.grad is not None, (name, n)
                    errs[name + "." + n] = l2_err(pg.grad, po.grad)
        assert len(errs) > 150
        bad = {n: e for n, e in errs.items() if e > gtol}
        print("[%s] parameter-gradient L2: median %.2e max %.2e (%s)" % (
            precision, float(np.median(list(errs.values()))), max(errs.values()), max(errs, key=errs.get)))
        assert not bad, bad
    finally:
        scn.set_precision("tf32")


def test_crop_backward_with_overlapping_boxes_and_extra_cut(cuda):
    """SparseRoiCut backward = scatter-add over overlapping boxes (float atomics: compared at 1e-6, not bit-exact), and
    SparseRoiExtraCut (roi_select_sparse.py:8-26, RawToFeatures combiner) forward + backward, against the reference
    algorithm on the CPU (expand + boolean-mask gather)."""
    from sparse_rcnn_b200 import roi, scn
    coords, feats, size, bs, splits = make_batch(2, 6, spatial_size=(32, 32, 16), room=(22, 22, 11), room_offset=(4, 4, 1),
                                                 n_furniture=1, density=1.0)
    boxes = [torch.tensor([[[4., 4, 0], [20, 20, 9]], [[6., 6, 0], [22, 22, 10]], [[4., 4, 0], [20, 20, 9]]]),
             torch.tensor([[[0., 0, 0], [32, 32, 16]], [[10., 3, 0], [18, 30, 12]]])]
    fo = feats.clone().requires_grad_(True)
    fg = feats.to(cuda).requires_grad_(True)
    extra = torch.randn(len(coords), 5, generator=torch.Generator().manual_seed(1))
    eo, eg = extra.clone().requires_grad_(True), extra.to(cuda).requires_grad_(True)
    out_o, sel_o = roi_ref.OracleRoiCut(O, raw_scene=True, combine="features")((coords, fo, size, bs, splits), boxes)
    out_g, sel_g = roi.SparseRoiCut(scn, raw_scene=True, combine="features")((coords, fg, size, bs, splits), boxes)
    inside = sel_o.is_inside()
    assert torch.equal(sel_g.is_inside(cpu=True), inside)
    assert int(inside.sum(0).max()) >= 3                              # some point sits in three boxes
    assert torch.equal(out_g.detach().cpu(), out_o.detach())
    ex_o = eo[None].expand(len(inside), -1, -1)[inside]               # select_features (roi_select_sparse.py:125-133)
    ex_g = roi.SparseRoiExtraCut()((coords, eg, size, bs, splits), sel_g)
    assert torch.equal(ex_g.detach().cpu(), ex_o.detach())
    w1 = torch.randn(out_o.shape, generator=torch.Generator().manual_seed(2))
    w2 = torch.randn(ex_o.shape, generator=torch.Generator().manual_seed(3))
    ((out_o * w1).sum() + (ex_o * w2).sum()).backward()
    ((out_g * w1.to(cuda)).sum() + (ex_g * w2.to(cuda)).sum()).backward()
    assert rel_err(fg.grad, fo.grad) <= 1e-6 and rel_err(eg.grad, eo.grad) <= 1e-6


@pytest.mark.parametrize("branch", ["loss_by_overlap", "loss_by_description"])
def test_consumers_on_the_device_crop_match_reference_golden(cuda, branch):
    """SparseMaskPredictor / SparseMaskLossSelector fed by the DEVICE crop (CropSelection CSR) against goldens of the
    unmodified reference classes (oracle/make_golden_consumers.py); MaskLoss on the selector's flat form equals MaskLoss on
    the reference's nested lists."""
    import types
    from sparse_rcnn_b200 import losses, roi, scn
    g = torch.load(os.path.join(G, "consumers.pt"), weights_only=False)
    coords, feats, size, bs, splits = make_batch(3, g["scene_seed"], spatial_size=(32, 32, 16), room=(16, 16, 8),
                                                 room_offset=(4, 4, 1), n_furniture=1, density=0.5)
    _, sel = roi.SparseRoiCut(scn, raw_scene=True, combine="features")((coords, feats.to(cuda), size, bs, splits), g["boxes"])
    inside = sel.is_inside(cpu=True).numpy()
    assert np.array_equal(np.packbits(inside, axis=1), g["inside_packed"])
    scores = g["scores"].to(cuda)
    for nv, key in ((0, "cls"), (18, "cls_hi")):
        got = roi.SparseMaskPredictor(nv)(scores, sel, g[key])
        for a, b in zip(got, g["predictor_%d" % nv]):
            assert a.shape == b.shape and torch.allclose(a.cpu(), b, atol=1e-6)
    selector = roi.SparseMaskLossSelector(0.5)
    if branch == "loss_by_overlap":
        descr, tuples = None, [(None, None, m, a) for m, a in zip(g["max_ov"], g["arg_ov"])]
    else:
        descr, tuples = [types.SimpleNamespace(gt_association=a) for a in g["assoc_given"]], None
    sc = scores.clone().requires_grad_(True)
    pred, gt, labels = selector(sc, sel, descr, tuples, g["gt_labels"], g["gt_masks"])
    ref = g[branch]
    for sa, sb in zip(pred, ref["pred"]):
        assert len(sa) == len(sb)
        for a, b in zip(sa, sb):
            assert torch.equal(a.detach().cpu(), b)
    for sa, sb in zip(gt, ref["gt"]):
        for a, b in zip(sa, sb):
            assert torch.equal(a.cpu().bool(), b.bool())
    loss_fn = losses.MaskLoss().to(cuda)
    l_flat = loss_fn.forward_flat(*selector.flat[:3], selector.flat[3])
    l_list = loss_fn(pred, gt, [l.to(cuda) for l in labels])
    assert abs(float(l_flat) - float(l_list)) <= 1e-6 * abs(float(l_list))
    ref_loss = torch.stack([torch.nn.functional.binary_cross_entropy_with_logits(a, b.float())
                            for sa, sb in zip(ref["pred"], ref["gt"]) for a, b in zip(sa, sb) if a.numel()]).mean()
    assert abs(float(l_flat) - float(ref_loss)) <= 1e-5 * abs(float(ref_loss))
    l_flat.backward()
    assert sc.grad is not None and bool(torch.isfinite(sc.grad).all()) and float(sc.grad.abs().sum()) > 0


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
