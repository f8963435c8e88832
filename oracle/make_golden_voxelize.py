"""Goldens for the voxelisation / collation row from the UNMODIFIED reference (run in the build container,
where /root/reference exists):  python oracle/make_golden_voxelize.py  ->  tests/golden/voxelize.pt

Every case stores the inputs, the random draws the reference made (projection matrix, sub-pixel offset; the reference's
RNG calls are left as they are, seeded) and what `augment_coords` + `augment_features` + `collate_fn` returned."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")
from ndsis.data import sparse_augmentation as A      # noqa: E402
from ndsis.data.data import collate_fn                # noqa: E402


class _MaskSelfAssign:
    """torch >= 2.x refuses `t[t] = values` (the mask aliases the destination), which random_cut_out does
    (sparse_augmentation.py:76: `is_inside[is_inside] = remaining_inside`).  The reference SOURCE stays untouched: while its
    function runs, Tensor.__setitem__ clones a boolean mask that is the destination itself -- the semantics the line had
    under the torch version the reference was written for (mask read, then written)."""

    def __enter__(self):
        self.orig = torch.Tensor.__setitem__
        orig = self.orig

        def setitem(t, key, value):
            if key is t:
                key = key.clone()
            return orig(t, key, value)
        torch.Tensor.__setitem__ = setitem

    def __exit__(self, *a):
        torch.Tensor.__setitem__ = self.orig


def sample(seed, n, extent):
    g = torch.Generator().manual_seed(seed)
    pts = torch.rand(n, 3, generator=g) * torch.tensor(extent)
    # scan-like: half of the points on a floor plane
    pts[: n // 2, 2] = 0.05 * torch.rand(n // 2, generator=g)
    colors = torch.rand(n, 3, generator=g) * 2 - 1
    normals = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=1)
    return pts, colors, normals


def one_case(seed, sizes, spatial_size, shift, scale, theta, mirror, sigma, noise):
    torch.manual_seed(seed)
    out_samples, rec = [], []
    for i, n in enumerate(sizes):
        pts, colors, normals = sample(seed * 100 + i, n, (6.0, 5.0, 2.5))
        offset = torch.rand(3)
        coords, is_inside, size, aug, rotation = A.augment_coords(
            pts, scale=scale, spatial_size=spatial_size, max_empty_border_size_divisor=None, shift=shift,
            sub_pixel_offset=offset, coord_noise_sigma=sigma, theta=theta, mirror=mirror)
        feats, faug = A.augment_features(
            colors, normals, is_inside, coords, color_noise_sigma=noise, common_color_noise=True, normal_noise_sigma=noise,
            common_normal_noise=True, rotation=rotation, use_color=True, use_ones=False, use_normal=True)
        rec.append(dict(points=pts, colors=colors, normals=normals, offset=offset, proj=aug["coords_projection"],
                        rotation=rotation, coords=coords, is_inside=is_inside, complete_shift=aug["coords_shift"],
                        features=feats, color_shift=faug["color_shift"], normal_shift=faug["normals_shift"]))
        dummy = torch.zeros(0)
        out_samples.append(("s%d" % i, coords, feats, dummy, dummy, dummy, torch.zeros(len(coords), dtype=torch.long), {}, size))
    batch = collate_fn(out_samples)
    cb, fb, ss, bs, splits = batch["data"]
    return dict(samples=rec, spatial_size=list(spatial_size), shift=shift, coords_batch=cb, features_batch=fb,
                batch_spatial_size=ss, batch_size=bs, batch_splits=splits)


def seeded_case(seed, sizes, spatial_size, scale, sigma, noise):
    """convert_sample + collate_fn with EVERY draw left to the reference (distortion, mirror, theta, sub-pixel offset, noise
    vectors): pins the order of the draws in sparse_rcnn_b200/voxelize.py:convert_and_collate."""
    samples = []
    for i, n in enumerate(sizes):
        pts, colors, normals = sample(seed * 100 + i, n, (5.0, 4.0, 2.5))
        g = torch.Generator().manual_seed(seed * 1000 + i)
        n_inst = 5 + i
        samples.append(("s%d" % i, pts, colors, normals, torch.randint(0, n_inst + 1, (n,), generator=g),      # n_inst = no instance
                        torch.randint(1, 19, (n_inst,), generator=g)))
    torch.manual_seed(seed)
    conv = [A.convert_sample(
        smp, spatial_size=spatial_size, instance_cutoff_threshold=0.5, color_noise_sigma=noise, common_color_noise=True,
        normal_noise_sigma=noise, common_normal_noise=True, use_color=True, use_ones=False, use_normal=True,
        additional_bbox_pixel=0, background_label=0, scale=scale, max_empty_border_size_divisor=None, shift=0,
        sub_pixel_offset=None, coord_noise_sigma=sigma, theta=None, mirror=None) for smp in samples]
    batch = collate_fn(conv)
    cb, fb, ss, bs, splits = batch["data"]
    return dict(seed=seed, inputs=[(s[1], s[2], s[3]) for s in samples], spatial_size=list(spatial_size), scale=scale, sigma=sigma,
                noise=noise, coords_batch=cb, features_batch=fb, batch_splits=splits, gt_segmentation=batch["gt_segmentation"],
                instance_ids=[s[4] for s in samples], semantic_instance_labels=[s[5] for s in samples],
                coords_shift=[a["coords_shift"] for a in batch["augmentation"]],
                coords_projection=[a["coords_projection"] for a in batch["augmentation"]])


def train_config_case(seed, sizes, spatial_size, scale):
    """The shipped TRAINING loader's augmentation (scannet_config/run.py:971-984): random cut-out (shift=None), random mirror /
    rotation / sub-pixel offset, coordinate noise 0.1, per-point colour noise 0.1 -- every draw left to the reference."""
    samples = []
    for i, n in enumerate(sizes):
        pts, colors, normals = sample(seed * 100 + i, n, (5.0, 4.0, 2.5))
        g = torch.Generator().manual_seed(seed * 1000 + i)
        n_inst = 4 + i
        samples.append(("s%d" % i, pts, colors, normals, torch.randint(0, n_inst + 1, (n,), generator=g),
                        torch.randint(1, 19, (n_inst,), generator=g)))
    torch.manual_seed(seed)
    with _MaskSelfAssign():
      conv = [A.convert_sample(
        smp, spatial_size=spatial_size, instance_cutoff_threshold=0.5, color_noise_sigma=0.1, common_color_noise=False,
        normal_noise_sigma=0, common_normal_noise=False, use_color=True, use_ones=False, use_normal=True,
        additional_bbox_pixel=0, background_label=0, scale=scale, max_empty_border_size_divisor=None, shift=None,
        sub_pixel_offset=None, coord_noise_sigma=0.1, theta=None, mirror=None) for smp in samples]
    batch = collate_fn(conv)
    cb, fb, ss, bs, splits = batch["data"]
    return dict(seed=seed, inputs=[(s[1], s[2], s[3]) for s in samples], spatial_size=list(spatial_size), scale=scale,
                coords_batch=cb, features_batch=fb, batch_splits=splits, gt_segmentation=batch["gt_segmentation"],
                instance_ids=[s[4] for s in samples], semantic_instance_labels=[s[5] for s in samples],
                coords_shift=[a["coords_shift"] for a in batch["augmentation"]],
                remaining=[a["remaining_points"] for a in batch["augmentation"]])


def cut_cases():
    """random_cut_out on its own: discrete coordinates, window, border -> (start positions, is_inside), seeded."""
    out = []
    for seed, n, extent, size, border in [(0, 4000, (150, 120, 60), (96, 96, 128), (0, 0, 0)), (1, 3000, (300, 40, 200), (64, 64, 64), (8, 8, 8)),
                                          (2, 10, (5, 5, 5), (16, 16, 16), (0, 0, 0)), (3, 2000, (100, 100, 100), (32, 32, 32), (0, 0, 0))]:
        g = torch.Generator().manual_seed(seed)
        disc = (torch.rand(n, 3, generator=g) * torch.tensor(extent)).long()
        torch.manual_seed(100 + seed)
        with _MaskSelfAssign():
            start, inside, coords = A.random_cut_out(disc.clone(), torch.tensor(size), border)
        out.append(dict(seed=100 + seed, disc=disc, size=list(size), border=list(border), start=start, is_inside=inside, coords=coords))
    return out


if __name__ == "__main__":
    cases = [
        one_case(1, [4000, 2500], (256, 256, 128), 0, 50.0, 0.7, False, 0.0, 0.0),
        one_case(2, [3000, 1, 2000], (128, 96, 64), 3, 22.0, 2.1, True, 0.02, 0.05),      # points cut away, a 1-point sample
        one_case(3, [6000], (64, 64, 32), 0, 20.0, 4.0, False, 0.01, 0.0),
    ]
    cases.append(seeded_case(7, [3000, 2200], (192, 192, 96), 30.0, 0.01, 0.03))
    cases.append(train_config_case(11, [3500, 2600], (96, 96, 128), 30.0))
    cases.append(dict(cut_cases=cut_cases()))
    path = os.path.join(ROOT, "tests", "golden", "voxelize.pt")
    torch.save(cases, path)
    print(path, os.path.getsize(path), "bytes;", [c.get("batch_splits") for c in cases])
