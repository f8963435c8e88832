"""GPU parity: voxelisation + collation on the device (csrc/voxelize.cu, sparse_rcnn_b200/voxelize.py) against goldens from the
unmodified reference (tests/golden/voxelize.pt), against the numpy oracle on seeded inputs, and through properties at full
size.  Integer results bit-exact; float results bit-exact too (same fp32 operations in the same order)."""
import os

import numpy as np
import pytest
import torch

import voxelize_oracle as V

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "voxelize.pt")


def _cat(samples, key, dev):
    return torch.cat([s[key] for s in samples]).to(dev).contiguous()


def test_goldens_from_the_reference(cuda):
    from sparse_rcnn_b200 import voxelize as Z
    cases = torch.load(GOLDEN)
    for case in cases[:3]:
        ss = case["samples"]
        B = len(ss)
        ptr = [0]
        for s in ss:
            ptr.append(ptr[-1] + len(s["points"]))
        vox = Z.voxelize_batch(_cat(ss, "points", cuda), ptr, torch.stack([s["proj"] for s in ss]), torch.stack([s["offset"] for s in ss]),
                               case["spatial_size"], shift=case["shift"])
        assert vox["batch_splits"] == list(case["batch_splits"])
        assert torch.equal(vox["coords"].cpu(), case["coords_batch"])
        inside = torch.cat([s["is_inside"] for s in ss])
        assert torch.equal(vox["kept"].cpu().long(), inside.nonzero().flatten())
        for i, s in enumerate(ss):
            assert torch.equal(vox["complete_shift"][i], s["complete_shift"])
        shifts = lambda k: torch.stack([s[k] for s in ss]) if ss[0][k].ndim else None
        feats = Z.features_batch(vox, B, colors=_cat(ss, "colors", cuda), color_shift=shifts("color_shift"),
                                 normals=_cat(ss, "normals", cuda), rotation=torch.stack([s["rotation"] for s in ss]),
                                 normal_shift=shifts("normal_shift"))
        assert torch.equal(feats.cpu(), case["features_batch"])


def test_seeded_conversion_draws_like_the_reference(cuda):
    """Every random draw left to the code under test: same seed, same order of torch calls as convert_sample."""
    from sparse_rcnn_b200 import voxelize as Z
    c = torch.load(GOLDEN)[3]
    torch.manual_seed(c["seed"])
    data, aug, kept = Z.convert_and_collate(c["inputs"], spatial_size=c["spatial_size"], scale=c["scale"], shift=0,
                                            coord_noise_sigma=c["sigma"], color_noise_sigma=c["noise"], normal_noise_sigma=c["noise"],
                                            device=cuda)
    assert data[4] == list(c["batch_splits"]) and data[3] == 2
    assert torch.equal(data[0].cpu(), c["coords_batch"])
    assert torch.equal(data[1].cpu(), c["features_batch"])
    for a, proj, sh in zip(aug, c["coords_projection"], c["coords_shift"]):
        assert torch.equal(a["coords_projection"], proj) and torch.equal(a["coords_shift"], sh)
    # labels of the kept points (get_semantic_segmentation_labels + collate_fn's gt_segmentation)
    ptr = [0]
    for pts, _, _ in c["inputs"]:
        ptr.append(ptr[-1] + len(pts))
    vox = dict(coords=data[0], kept=kept)
    seg = Z.segmentation_labels_batch(vox, ptr, c["instance_ids"], c["semantic_instance_labels"], background_label=0)
    assert torch.equal(seg.cpu(), c["gt_segmentation"])


def test_training_loader_configuration(cuda):
    """The shipped TRAINING augmentation (scannet_config/run.py:971-984): random cut-out, random mirror / rotation / offset,
    coordinate noise, one colour-noise value per kept point -- every draw made by the code under test, sample after sample;
    golden from the unmodified convert_sample + collate_fn under the same seed."""
    from sparse_rcnn_b200 import voxelize as Z
    c = torch.load(GOLDEN)[4]
    torch.manual_seed(c["seed"])
    data, aug, kept = Z.convert_and_collate(c["inputs"], spatial_size=c["spatial_size"], scale=c["scale"], shift=None,
                                            coord_noise_sigma=0.1, color_noise_sigma=0.1, common_color_noise=False,
                                            normal_noise_sigma=0, common_normal_noise=False, device=cuda)
    assert data[4] == list(c["batch_splits"]) and 0 < sum(data[4]) < sum(len(p[0]) for p in c["inputs"])      # points were cut away
    assert torch.equal(data[0].cpu(), c["coords_batch"])
    assert torch.equal(data[1].cpu(), c["features_batch"])
    inside = torch.cat(c["remaining"])
    assert torch.equal(kept.cpu().long(), inside.nonzero().flatten())
    for a, sh in zip(aug, c["coords_shift"]):
        assert torch.equal(a["coords_shift"], sh)
    ptr = [0]
    for pts, _, _ in c["inputs"]:
        ptr.append(ptr[-1] + len(pts))
    seg = Z.segmentation_labels_batch(dict(coords=data[0], kept=kept), ptr, c["instance_ids"], c["semantic_instance_labels"], 0)
    assert torch.equal(seg.cpu(), c["gt_segmentation"])


@pytest.mark.parametrize("seed,sizes,start", [(0, [5000, 0, 3000, 1], None), (1, [257, 4097], (7, 3, 0)), (2, [1], None)])
def test_against_the_oracle(cuda, seed, sizes, start):
    """Ragged batches with empty and one-point samples, a drawn cut-out, ones column, no normals."""
    from sparse_rcnn_b200 import voxelize as Z
    rng = np.random.default_rng(seed)
    B = len(sizes)
    pts = [rng.random((n, 3)).astype(np.float32) * np.array([6, 5, 2.5], np.float32) for n in sizes]
    cols = [rng.standard_normal((n, 3)).astype(np.float32) for n in sizes]
    proj = (rng.standard_normal((B, 3, 3)) * 0.05 + np.eye(3) * 20).astype(np.float32)
    off = rng.random((B, 3)).astype(np.float32)
    cshift = rng.standard_normal((B, 3)).astype(np.float32)
    size = (96, 64, 48)
    st = None if start is None else np.tile(np.array(start), (B, 1))
    want_c, want_f = [], []
    for b in range(B):
        c, inside, _ = V.voxelize_sample(pts[b], proj[b], off[b], size, shift=2 if st is None else None, start=None if st is None else st[b])
        want_c.append(c)
        want_f.append(V.features_sample(inside, colors=cols[b], color_shift=cshift[b], use_ones=True))
    wc, wf, splits = V.collate(want_c, want_f)
    ptr = np.concatenate([[0], np.cumsum(sizes)]).tolist()
    dev_pts = torch.from_numpy(np.concatenate(pts)).to(cuda)
    vox = Z.voxelize_batch(dev_pts, ptr, torch.from_numpy(proj), torch.from_numpy(off), size, shift=2 if st is None else None, start=st)
    assert vox["batch_splits"] == splits
    assert np.array_equal(vox["coords"].cpu().numpy(), wc)
    f = Z.features_batch(vox, B, colors=torch.from_numpy(np.concatenate(cols)).to(cuda), color_shift=torch.from_numpy(cshift), use_ones=True)
    assert np.array_equal(f.cpu().numpy(), wf)


def test_full_size_properties_and_the_input_layer(cuda):
    """Eight bench-sized samples (2.2 M points): every coordinate inside the window, rows grouped by sample in input order,
    kept rows ascending, re-voxelising the kept points gives the same coordinates (idempotence), and the collated device
    tensor feeds scn.InputLayer directly -- the same active set as from the reference's host tensor."""
    from sparse_rcnn_b200 import scn, voxelize as Z
    g = torch.Generator().manual_seed(0)
    sizes = [273000 + 1000 * i for i in range(8)]
    pts = torch.rand(sum(sizes), 3, generator=g) * torch.tensor([5.2, 5.0, 2.6])
    pts[::2, 2] = 0.02
    ptr = np.concatenate([[0], np.cumsum(sizes)]).tolist()
    B, size = 8, (256, 256, 128)
    proj = torch.stack([Z.coord_distortion_matrix(torch.float32, 0.01, None, None) * 50.0 for _ in range(B)])
    off = torch.rand(B, 3)
    dp = pts.to(cuda)
    vox = Z.voxelize_batch(dp, ptr, proj, off, size, shift=0)
    c, kept = vox["coords"], vox["kept"].long()
    assert vox["n"] == sum(vox["batch_splits"]) and 0 < vox["n"] <= sum(sizes)
    assert bool((c[:, :3] >= 0).all()) and bool((c[:, :3] < torch.tensor(size, device=cuda)).all())
    assert bool((c[1:, 3] >= c[:-1, 3]).all()) and bool((kept[1:] > kept[:-1]).all())
    bounds = torch.tensor(ptr, device=cuda)
    assert torch.equal(torch.bucketize(kept, bounds, right=True) - 1, c[:, 3])
    # idempotence: the kept points alone, with the offset that reproduces the first run's shift
    sub_ptr = np.concatenate([[0], np.cumsum(vox["batch_splits"])]).tolist()
    again = Z.voxelize_batch(dp[kept].contiguous(), sub_ptr, proj, off, size, shift=0)
    same_min = torch.equal(again["complete_shift"], vox["complete_shift"])      # true unless a cut-away point held the minimum
    if same_min:
        assert torch.equal(again["coords"], c)
    # one sample through the input layer: device coords == host coords
    n0 = vox["batch_splits"][0]
    md_d, md_h = scn.Metadata(3), scn.Metadata(3)
    feats = torch.ones(n0, 1, device=cuda)
    size_t = torch.tensor(size)
    fd = scn.ioLayers.InputLayerFunction.apply(3, md_d, size_t, c[:n0], feats, 1, 4)
    fh = scn.ioLayers.InputLayerFunction.apply(3, md_h, size_t, c[:n0].cpu(), feats, 1, 4)
    assert fd.shape == fh.shape and torch.equal(md_d.level(size_t).keys, md_h.level(size_t).keys)
