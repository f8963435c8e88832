"""CPU: the voxelisation / collation oracle (oracle/voxelize_oracle, numpy) against goldens produced by the unmodified
reference (`augment_coords`, `augment_features`, `collate_fn`; oracle/make_golden_voxelize.py): integer results bit-exact,
float results bit-exact as well (same fp32 operations in the same order)."""
import os

import numpy as np
import torch

import voxelize_oracle as V

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "voxelize.pt")


def run_case(case):
    coords_list, feats_list = [], []
    for s in case["samples"]:
        c, inside, cs = V.voxelize_sample(s["points"].numpy(), s["proj"].numpy(), s["offset"].numpy(), case["spatial_size"],
                                          shift=case["shift"])
        assert np.array_equal(inside, s["is_inside"].numpy())
        assert np.array_equal(c, s["coords"].numpy())
        # complete_shift as the reference reports it: minus the cut-out start (= +shift for fix_cut_out)
        want = s["complete_shift"].numpy()
        assert np.array_equal((cs + np.float32(case["shift"])).astype(np.float32), want)
        cshift = s["color_shift"].numpy() if s["color_shift"].ndim else None
        nshift = s["normal_shift"].numpy() if s["normal_shift"].ndim else None
        f = V.features_sample(inside, colors=s["colors"].numpy(), color_shift=cshift, normals=s["normals"].numpy(),
                              rotation=s["rotation"].numpy(), normal_shift=nshift)
        assert np.array_equal(f, s["features"].numpy())
        coords_list.append(c), feats_list.append(f)
    cb, fb, splits = V.collate(coords_list, feats_list)
    assert np.array_equal(cb, case["coords_batch"].numpy()) and np.array_equal(fb, case["features_batch"].numpy())
    assert splits == list(case["batch_splits"])


def test_oracle_reproduces_the_reference():
    cases = torch.load(GOLDEN)
    assert len(cases) == 6
    for case in cases[:3]:
        run_case(case)
    # the fully seeded case (every draw made by the reference): the oracle with the recorded projection, and the shift the
    # reference reports minus its own minimum gives back the drawn sub-pixel offset
    c = cases[3]
    at = 0
    for (pts, colors, normals), proj, shift, n in zip(c["inputs"], c["coords_projection"], c["coords_shift"], c["batch_splits"]):
        aug = V.project(pts.numpy(), proj.numpy())
        offset = shift.numpy() + aug.min(0)
        got, inside, _ = V.voxelize_sample(pts.numpy(), proj.numpy(), offset, c["spatial_size"], shift=0)
        want = c["coords_batch"][at:at + n, :3].numpy()
        # offset recovered through an fp32 subtraction: a coordinate within one ulp of an integer may land next door
        assert got.shape == want.shape and (got != want).any(1).mean() < 1e-3
        at += n


def test_start_positions_form_and_empty_sample():
    rng = np.random.default_rng(0)
    pts = rng.random((500, 3)).astype(np.float32) * 4
    proj = (np.eye(3) * 20).astype(np.float32)
    c0, in0, _ = V.voxelize_sample(pts, proj, [0.5, 0.5, 0.5], (32, 32, 32), shift=0)
    c1, in1, _ = V.voxelize_sample(pts, proj, [0.5, 0.5, 0.5], (32, 32, 32), start=(0, 0, 0))
    assert np.array_equal(c0, c1) and np.array_equal(in0, in1)
    c2, in2, _ = V.voxelize_sample(pts, proj, [0.5, 0.5, 0.5], (32, 32, 32), start=(10, 5, 0))
    assert in2.sum() < in0.sum() + 500 and (c2 >= 0).all() and (c2 < 32).all()
    e, ine, cs = V.voxelize_sample(np.zeros((0, 3), np.float32), proj, [0, 0, 0], (8, 8, 8), shift=0)
    assert e.shape == (0, 3) and ine.shape == (0,)


def test_random_cut_out_draws_like_the_reference():
    """sparse_rcnn_b200.voxelize.draw_random_cut (the host logic of the training loader's random cut-out; pure torch, runs on
    any device) and the oracle's random_cut against the unmodified reference function's outputs: same start positions, same
    points inside, same coordinates -- with the reference's RNG calls in the reference's order."""
    from sparse_rcnn_b200 import voxelize as Z
    cut = torch.load(GOLDEN)[5]["cut_cases"]
    assert len(cut) == 4
    branches = set()
    for c in cut:
        torch.manual_seed(c["seed"])
        start, inside = Z.draw_random_cut(c["disc"], c["size"], c["border"])
        assert torch.equal(start, c["start"]) and torch.equal(inside, c["is_inside"])
        assert torch.equal(c["disc"][inside] - start, c["coords"])
        # the oracle with the same draws
        torch.manual_seed(c["seed"])
        order = torch.multinomial(torch.ones(3), 3).tolist()
        draws = iter(lambda: (lambda lo, hi: int(torch.randint(lo, hi, ()))), None)
        s2, in2 = V.random_cut(c["disc"].numpy(), c["size"], c["border"], order, draws)
        assert np.array_equal(s2, c["start"].numpy()) and np.array_equal(in2, c["is_inside"].numpy())
        branches.add(bool(inside.all()))
    assert branches == {True, False}      # both the "extent fits" and the "draw and cut" branch were taken


def test_host_side_draws_and_label_gather_on_cpu():
    """Host logic of sparse_rcnn_b200.voxelize that needs no GPU: the first sample's distortion-matrix draws of the seeded
    golden (get_coord_distortion_matrix with theta / mirror left to the RNG) and the label gather with a label mapper."""
    from sparse_rcnn_b200 import voxelize as Z
    c = torch.load(GOLDEN)[3]
    torch.manual_seed(c["seed"])
    rot = Z.coord_distortion_matrix(torch.float32, c["sigma"], None, None)
    assert torch.equal(rot * c["scale"], c["coords_projection"][0])
    # every point kept (that case cuts nothing away): gt_segmentation == table[instance ids]
    n = [len(p[0]) for p in c["inputs"]]
    assert c["batch_splits"] == n
    ptr = [0, n[0], n[0] + n[1]]
    vox = dict(coords=torch.zeros(1), kept=torch.arange(ptr[-1], dtype=torch.int32))
    seg = Z.segmentation_labels_batch(vox, ptr, c["instance_ids"], c["semantic_instance_labels"], background_label=0)
    assert torch.equal(seg, c["gt_segmentation"])
    mapper = torch.arange(40) * 2
    seg2 = Z.segmentation_labels_batch(vox, ptr, c["instance_ids"], c["semantic_instance_labels"], background_label=-100, label_mapper=mapper)
    want = torch.where(c["gt_segmentation"] == 0, torch.tensor(-100), c["gt_segmentation"] * 2)
    # label 0 never occurs among the instance labels (drawn from 1..18), so a zero in the golden is the background
    assert torch.equal(seg2, want)
