"""CPU: the oracle (and the host-side graph mirror running on it) against golden vectors produced
by the UNMODIFIED reference (oracle/make_golden.py, run in the build container)."""
import os

import numpy as np
import pytest
import torch

import scn_oracle as O
from scn_oracle import roi_ref
from sparse_rcnn_b200 import networks
from sparse_rcnn_b200.synthetic import make_batch, make_boxes
from tests.util import reinit_by_name

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["crop_raw", "crop_stride4"])
def test_oracle_crop_matches_reference_roi_cut(name):
    g = torch.load(os.path.join(G, name + ".pt"), weights_only=False)
    coords, feats = g["coords"].long(), g["feats"]
    boxes, counts, assoc = roi_ref.transform_boxes(g["boxes"], g["size"], g["clip"], g["resize"])
    assert torch.equal(boxes, g["box_tensor"]) and counts == g["counts"] and torch.equal(assoc, g["assoc"])
    nc, nf, inside = roi_ref.roi_cut(coords, feats, boxes, assoc)
    assert torch.equal(nc, g["new_coords"].long())
    assert np.array_equal(np.packbits(inside.numpy(), axis=1), g["inside_packed"])
    assert tuple(inside.shape) == g["inside_shape"]
    assert torch.equal(nf[:32], g["new_feats_head"]) and abs(float(nf.double().sum()) - g["new_feats_sum"]) < 1e-6
    # product-side box transformer (pure host logic) agrees with the reference too
    from sparse_rcnn_b200.roi import BBoxTransformerSlice
    b2, c2, a2 = BBoxTransformerSlice(clip=g["clip"], resize=g["resize"])(g["boxes"], g["size"])
    assert torch.equal(b2, g["box_tensor"]) and c2 == g["counts"] and torch.equal(a2, g["assoc"])


@pytest.fixture(scope="module")
def graph():
    return torch.load(os.path.join(G, "ref_graph.pt"), weights_only=False)


def _nets():
    cut = lambda **kw: roi_ref.OracleRoiCut(O, **kw)
    fe = networks.FeatureExtractor(O)
    seg = networks.SegmentationNetwork(O)
    cls = networks.ClassNetwork(O, cut)
    mask = networks.SparseMaskNetwork(O, cut)
    for m in (fe, seg, cls, mask):
        reinit_by_name(m).eval()
    return fe, seg, cls, mask


def test_graph_mirror_has_reference_state_dict(graph):
    fe, seg, cls, mask = _nets()
    for name, m in (("fe", fe), ("seg", seg), ("cls", cls), ("mask", mask)):
        mine = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        assert mine == graph["state"][name], (name, set(mine) ^ set(graph["state"][name]))
        wsum = float(sum(v.double().sum() for v in m.state_dict().values()))
        assert abs(wsum - graph["weight_sums"][name]) < 1e-6 * max(1, abs(wsum))


def test_graph_mirror_reproduces_reference_outputs(graph):
    fe, seg, cls, mask = _nets()
    data = make_batch(2, graph["scene_seed"], spatial_size=(64, 64, 32), room=(44, 44, 22), room_offset=(8, 8, 2),
                      n_furniture=3)
    boxes = make_boxes(data[0], 5, graph["box_seed"], (64, 64, 32))
    with torch.no_grad():
        _, bs, _, class_map, inter, unet = fe(data)
        assert [t.features.shape[0] for t in inter] == graph["level_rows"]
        u = unet[-1].features
        assert torch.allclose(u[:64], graph["unet_last_head"], rtol=1e-4, atol=1e-4 * graph["unet_last_absmax"])
        assert abs(float(u.double().sum()) - graph["unet_last_sum"]) <= 1e-5 * u.numel() * graph["unet_last_absmax"]
        s = seg(unet)
        assert torch.allclose(s[:64], graph["seg_head"], rtol=1e-4, atol=1e-3)
        c, csel = cls(class_map, boxes)
        assert torch.allclose(c, graph["cls_out"], rtol=1e-4, atol=1e-3 * float(graph["cls_out"].abs().max()))
        assert int(csel.is_inside().sum()) == graph["cls_inside_count"]
        m, msel = mask(data, unet, boxes)
        assert m.shape[0] == graph["mask_rows"] and int(msel.is_inside().sum()) == graph["mask_inside_count"]
        assert torch.allclose(m[:64], graph["mask_head"], rtol=1e-4, atol=1e-3 * float(graph["mask_head"].abs().max()))


def test_reference_graph_live_when_available():
    """When /root/reference is present (build container), rebuild the reference modules on the oracle and
    compare state_dict keys live (guards against stale goldens)."""
    if not os.path.isdir("/root/reference/ndsis"):
        pytest.skip("reference checkout not present")
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys; sys.path[:0]=[%r, %r, '/root/reference']\n"
        "import scn_oracle; sys.modules['sparseconvnet']=scn_oracle\n"
        "import importlib.util, torch\n"
        "spec=importlib.util.spec_from_file_location('mg', %r); mg=importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)\n"
        "from ndsis.modules.model import FeatureExtractor\n"
        "from sparse_rcnn_b200 import networks\n"
        "fe_p, unet_p, _, _ = mg.reference_configs()\n"
        "ref = FeatureExtractor(**fe_p, include_unet=True, unet_params=unet_p)\n"
        "mine = networks.FeatureExtractor(scn_oracle)\n"
        "a={k:tuple(v.shape) for k,v in ref.state_dict().items()}; b={k:tuple(v.shape) for k,v in mine.state_dict().items()}\n"
        "assert a==b, set(a)^set(b)\nprint('OK', len(a))\n") % (root, os.path.join(root, "oracle"),
                                                                 os.path.join(root, "oracle", "make_golden.py"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OK 120" in r.stdout, r.stdout + r.stderr


# ----------------------------------------------------------------------------- proposal selection
def test_nms_oracle_matches_reference():
    """oracle/scn_oracle/nms_ref.py == the unmodified reference's ProposalSelector / non_maximum_supression
    (tests/golden/nms.pt, oracle/make_golden_nms.py), bit for bit."""
    import os
    import torch
    from scn_oracle import nms_ref
    from sparse_rcnn_b200.synthetic import make_proposals
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "nms.pt"))
    for c in gold["cases"]:
        score, boxes = make_proposals(c["seed"], c["B"], c["A"], clustered=c["clustered"])
        s, b, i = nms_ref.select(score, boxes, c["pre"], c["post"], c["thresh"])
        for k in range(c["B"]):
            assert torch.equal(i[k], c["indices"][k])
            assert torch.equal(s[k], c["scores"][k])
        if c["pre"] > 0:
            _, order = torch.topk(score, c["pre"], dim=1, sorted=True)
        else:
            _, order = torch.sort(score, dim=1, descending=True)
        for k in range(c["B"]):
            assert torch.equal(nms_ref.nms(boxes[k][order[k]], c["thresh"]), c["keep"][k])


# ----------------------------------------------------------------------------- per-box mask loss
def test_maskloss_oracle_matches_reference():
    """oracle/scn_oracle/maskloss_ref.py (fp64) vs the unmodified reference's MaskLoss (fp32): 1e-6 relative."""
    import os
    import torch
    from scn_oracle import maskloss_ref
    from sparse_rcnn_b200.synthetic import make_mask_loss_case
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "maskloss.pt"))
    for c in gold["cases"]:
        outs, tgts, cls = make_mask_loss_case(c["seed"], c["boxes_per_sample"], empty_every=c["empty_every"])
        w = (torch.arange(18, dtype=torch.float32) % 5 + 0.5) if c["weighted"] else None
        got = maskloss_ref.mask_loss(outs, tgts, cls, w)
        assert abs(got - c["loss"]) <= 1e-6 * max(abs(c["loss"]), 1.0), (got, c["loss"])


# ------------------------------------------------------------------ consumers of the crop selection
def _consumer_case():
    import types
    from sparse_rcnn_b200 import roi
    g = torch.load(os.path.join(G, "consumers.pt"), weights_only=False)
    inside = torch.from_numpy(np.unpackbits(g["inside_packed"], axis=1)[:, :g["inside_shape"][1]].astype(bool))
    box, pt = inside.nonzero(as_tuple=True)                       # row-major = (box, point) order of the crop
    box_ptr = torch.zeros(inside.shape[0] + 1, dtype=torch.int32)
    box_ptr[1:] = inside.sum(1).cumsum(0).to(torch.int32)
    sel = roi.CropSelection(pt.to(torch.int32), None, box_ptr, inside.shape[0], inside.shape[1],
                            inside.to(torch.uint8).reshape(-1), g["counts"], g["splits"])
    return g, inside, sel, types


def test_split_select_nd_matches_reference():
    from sparse_rcnn_b200 import roi
    g, inside, sel, _ = _consumer_case()
    blocks = roi.split_select_nd(inside, torch.tensor([g["counts"], g["splits"]]))
    assert [tuple(b.shape) for b in blocks] == g["blocks_shape"]
    assert [int(b.sum()) for b in blocks] == g["blocks_sum"]
    with pytest.raises(ValueError):
        roi.split_select_nd(inside, torch.tensor([g["counts"], [1, 2, 3]]))


@pytest.mark.parametrize("num_valid,cls_key", [(0, "cls"), (18, "cls_hi")])
def test_mask_predictor_matches_reference_golden(num_valid, cls_key):
    """Golden from the UNMODIFIED reference SparseMaskPredictor (model.py:826-882, oracle/make_golden_consumers.py)."""
    from sparse_rcnn_b200 import roi
    g, inside, sel, _ = _consumer_case()
    got = roi.SparseMaskPredictor(num_valid)(g["scores"], sel, g[cls_key])
    ref = g["predictor_%d" % num_valid]
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        assert a.shape == b.shape and torch.allclose(a, b, atol=1e-7)


@pytest.mark.parametrize("branch", ["loss_by_overlap", "loss_by_description"])
def test_mask_loss_selector_matches_reference_golden(branch):
    """Golden from the UNMODIFIED reference SparseMaskLossSelector (model.py:1152-1227), both branches; the flat form feeds
    losses.MaskLoss.forward_flat."""
    from sparse_rcnn_b200 import roi
    g, inside, sel, types = _consumer_case()
    selector = roi.SparseMaskLossSelector(0.5)
    if branch == "loss_by_overlap":
        descr, tuples = None, [(None, None, m, a) for m, a in zip(g["max_ov"], g["arg_ov"])]
    else:
        descr, tuples = [types.SimpleNamespace(gt_association=a) for a in g["assoc_given"]], None
    pred, gt, labels = selector(g["scores"], sel, descr, tuples, g["gt_labels"], g["gt_masks"])
    ref = g[branch]
    assert [len(s) for s in pred] == [len(s) for s in ref["pred"]]
    for sa, sb in zip(pred, ref["pred"]):
        for a, b in zip(sa, sb):
            assert torch.equal(a, b)
    for sa, sb in zip(gt, ref["gt"]):
        for a, b in zip(sa, sb):
            assert torch.equal(a.bool(), b.bool())
    for a, b in zip(labels, ref["labels"]):
        assert torch.equal(a, b)
    logits, targets, lens, lab = selector.flat
    assert logits.numel() == sum(lens) == targets.numel() and lab.numel() == len(lens)
