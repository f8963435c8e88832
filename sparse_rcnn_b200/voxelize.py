"""Voxelisation + collation of a batch on the device: host mirror of the reference's per-sample conversion

    ndsis/data/sparse_augmentation.py:81-126  augment_coords   (project, shift to the origin, discretise, cut out)
    ndsis/data/sparse_augmentation.py:129-190 augment_features (rows that stay, normals rotated, common noise vectors)
    ndsis/data/data.py:88-115                 collate_fn       (batch-index column, concatenation, batch_splits)

for ALL samples of a batch in four kernel launches (csrc/voxelize.cu) instead of ~25 torch ops per sample.  The reference's
random draws are made HERE, on the host, with the reference's own torch calls in the reference's order (distortion matrix,
sub-pixel offset, noise vectors), so a seeded run reproduces `convert_sample` bit for bit.  The fixed cut-out (`shift`: the reference's
validation / overfit loaders, scannet_config/run.py:957-969,995-998) needs no data-dependent draw and converts a whole batch
in one pass; the random cut-out of the TRAINING loader (`shift=None`, run.py:971-984) draws its start positions from ranges
of the points that are still inside, dimension by dimension (sparse_augmentation.py:49-78), and per-point colour noise
draws one value per kept point: `convert_and_collate` then converts sample after sample (`draw_random_cut`: the reference's
RNG calls in the reference's order, the min / max of a dimension read back as the reference reads them).

Output follows the reference's collate contract: (coords int64 [P', 4] (x, y, z, sample), features fp32 [P', C],
spatial_size, batch_size, batch_splits) -- coords stay ON THE DEVICE (scn.InputLayer accepts them there; `.cpu()` gives the
reference's host tensor)."""
from math import pi

import torch

from . import _lib
from .scn.metadata import _ptr, _stream


def coord_distortion_matrix(dtype, coord_noise_sigma, theta, mirror):
    """The draws of get_coord_distortion_matrix (sparse_augmentation.py:9-38) in its order: randn(3,3), the mirror coin (if
    not fixed), theta (if not fixed; a list draws one entry).  Returns the almost-orthonormal 3x3 matrix."""
    m = torch.eye(3, dtype=dtype) + torch.randn((3, 3), dtype=dtype) * coord_noise_sigma
    m[0, 0] *= (torch.randint(0, 2, ()) * 2 - 1) if mirror is None else (-1 if mirror else 1)
    if theta is None:
        theta = torch.rand((), dtype=dtype) * 2 * pi
    else:
        theta = torch.tensor(theta, dtype=dtype)
        if theta.numel() > 1:
            theta = theta[torch.ones_like(theta).multinomial(1)[0]]
    c, s = torch.cos(theta), torch.sin(theta)
    return m @ torch.tensor([[c, s, 0.], [-s, c, 0.], [0., 0., 1.]])


_ACCEPT_ALL = (-(1 << 30), -(1 << 30), -(1 << 30), (1 << 31) - 1, (1 << 31) - 1, (1 << 31) - 1, 0, 0, 0)      # window: every point stays, unmoved


def draw_random_cut(disc, size, max_border=(0, 0, 0)):
    """The draws of random_cut_out (sparse_augmentation.py:49-78) on discrete coordinates `disc` int64 [P, 3] (any device):
    a random order of the dimensions (multinomial), then per dimension -- over the points that are still inside -- either
    the only possible start (the extent fits) or a `randint` start and the cut.  -> (start positions int64 [3] on the
    host, is_inside bool [P]); the coordinates that stay are disc[is_inside] - start.  Same torch RNG calls in the same order
    as the reference running on host tensors; the min / max of a dimension are the host round trips the reference makes too."""
    size = [int(v) for v in size]
    order = torch.multinomial(torch.ones(3), 3)
    start = torch.zeros(3, dtype=torch.long)
    inside = torch.ones(disc.shape[0], dtype=torch.bool, device=disc.device)
    n_inside = int(disc.shape[0])
    for dim in order.tolist():
        if not n_inside:
            break
        vals = disc[:, dim][inside]
        min_start = int(vals.min()) - int(max_border[dim])
        max_start = int(vals.max()) + 1 - size[dim] + int(max_border[dim])
        if max_start <= min_start:
            start[dim] = min_start
        else:
            start[dim] = torch.randint(min_start, max_start, ())
            rel = disc[:, dim] - int(start[dim])
            inside = inside & (rel >= 0) & (rel < size[dim])
            n_inside = int(inside.sum())
    return start, inside


def voxelize_batch(points, sample_ptr, proj, offset, spatial_size, shift=None, start=None, _window=None):
    """points fp32 [P, 3] (device, samples concatenated), sample_ptr [B + 1] (host ints), proj fp32 [B, 3, 3] = distortion *
    scale, offset fp32 [B, 3], spatial_size (3 ints), and either shift (int or [B, 3]: fix_cut_out) or start ([B, 3]: the
    start positions of a drawn cut-out).  -> dict(coords int64 [P', 4], kept int32 [P'] (input row of every output row),
    out_ptr int32 [B + 1] (device: batch_splits as offsets), complete_shift fp32 [B, 3] as the reference reports it)."""
    dev = points.device
    B = len(sample_ptr) - 1
    P = int(points.shape[0])
    assert points.dtype == torch.float32 and points.is_contiguous() and int(sample_ptr[-1]) == P and B >= 1
    size = [int(s) for s in spatial_size]
    win = torch.zeros((B, 9), dtype=torch.int32)
    win[:, 3:6] = torch.tensor(size, dtype=torch.int32)
    if _window is not None:
        win[:] = torch.tensor(_window, dtype=torch.int32)
        moved = torch.zeros((B, 3))
    elif (shift is None) == (start is None):
        raise ValueError("exactly one of shift (fixed cut-out) / start (drawn start positions) is needed")
    elif shift is not None:
        win[:, 6:9] = torch.as_tensor(shift, dtype=torch.int32).expand(B, 3) if not isinstance(shift, int) else int(shift)
        moved = -win[:, 6:9].float()
    else:
        st = torch.as_tensor(start, dtype=torch.int32).reshape(B, 3)
        win[:, 0:3], win[:, 6:9] = st, -st
        moved = st.float()
    host = torch.cat([torch.as_tensor(proj, dtype=torch.float32).reshape(B, 9), torch.as_tensor(offset, dtype=torch.float32).reshape(B, 3)], 1)
    params = host.to(dev, non_blocking=True)
    ints = torch.cat([torch.as_tensor(sample_ptr, dtype=torch.int32).reshape(-1), win.reshape(-1)]).to(dev, non_blocking=True)
    sp, window = ints[:B + 1], ints[B + 1:]
    pr, off = params[:, :9].contiguous(), params[:, 9:].contiguous()
    ws = torch.empty(int(_lib.raw("scn_voxelize_ws_bytes")(P, B)), dtype=torch.uint8, device=dev)
    coords = torch.empty((max(P, 1), 4), dtype=torch.int64, device=dev)
    kept = torch.empty(max(P, 1), dtype=torch.int32, device=dev)
    out_ptr = torch.empty(B + 1, dtype=torch.int32, device=dev)
    shift_out = torch.empty((B, 3), dtype=torch.float32, device=dev)
    _lib.call("scn_voxelize", _ptr(points), P, _ptr(sp), B, _ptr(pr), _ptr(off), _ptr(window), _ptr(ws), _ptr(coords), _ptr(kept),
              _ptr(out_ptr), _ptr(shift_out), _stream())
    splits_end = out_ptr.cpu()              # the one host read: how many points stayed (sizes the outputs)
    n = int(splits_end[-1])
    return dict(coords=coords[:n], kept=kept[:n], out_ptr=out_ptr, n=n,
                batch_splits=[int(b - a) for a, b in zip(splits_end[:-1], splits_end[1:])],
                complete_shift=shift_out.cpu() - moved)      # coords_shift of the reference: minus the cut-out start


def features_batch(vox, B, colors=None, color_shift=None, use_ones=False, normals=None, rotation=None, normal_shift=None):
    """[colours (+ common shift [B, 3]) | ones | normals @ rotation [B, 3, 3] (+ common shift)] of the kept points."""
    src = colors if colors is not None else normals
    dev = vox["coords"].device
    C = (3 if colors is not None else 0) + (1 if use_ones else 0) + (3 if normals is not None else 0)
    out = torch.empty((vox["n"], C), dtype=torch.float32, device=dev)
    up = lambda t: None if t is None else torch.as_tensor(t, dtype=torch.float32).reshape(B, -1).to(dev).contiguous()
    cs, rot, ns = up(color_shift), up(rotation), up(normal_shift)
    assert src is None or src.is_contiguous()
    _lib.call("scn_voxelize_features", _ptr(vox["kept"]), vox["n"], _ptr(vox["out_ptr"]), B, _ptr(colors), _ptr(cs), int(use_ones),
              _ptr(normals), _ptr(rot), _ptr(ns), _ptr(out), C, _stream())
    return out


def load_sample(scene_id, path):
    """data.py:6-8: the on-disk tuple (coords fp32 [P, 3], colors, normals, instance_ids, semantic_instance_labels_raw)
    written by the reference's preparation step, prefixed with the scene id."""
    return (scene_id,) + tuple(torch.load(path))


def segmentation_labels_batch(vox, sample_ptr, instance_ids, semantic_instance_labels, background_label, label_mapper=None):
    """get_semantic_segmentation_labels (sparse_augmentation.py:235-246) for the kept points of a whole batch: every sample's
    label table, optionally mapped, padded with the background label (the instance id one past the table = "no instance"),
    gathered through the sample's instance ids.  instance_ids: list of int64 [P_b]; semantic_instance_labels: list of int64
    [n_instances_b].  -> int64 [P'] on the device, in the collated row order (collate_fn's `gt_segmentation`)."""
    dev = vox["coords"].device
    tables, base, at = [], [], 0
    for lab in semantic_instance_labels:
        lab = torch.as_tensor(lab, dtype=torch.long)
        if label_mapper is not None:
            lab = torch.as_tensor(label_mapper, dtype=torch.long)[lab]
        tables.append(torch.nn.functional.pad(lab, (0, 1), value=background_label))
        base.append(at)
        at += len(lab) + 1
    table = torch.cat(tables).to(dev)
    ids = torch.cat([torch.as_tensor(i, dtype=torch.long) for i in instance_ids]).to(dev)
    kept = vox["kept"].long()
    sample_of = torch.bucketize(kept, torch.as_tensor(sample_ptr, dtype=torch.long, device=dev), right=True) - 1
    return table[torch.as_tensor(base, dtype=torch.long, device=dev)[sample_of] + ids[kept]]


def convert_and_collate(samples, *, spatial_size, scale, shift=0, start=None, coord_noise_sigma=0.0, theta=None, mirror=None,
                        sub_pixel_offset=None, color_noise_sigma=0.0, normal_noise_sigma=0.0, use_color=True, use_ones=False,
                        use_normal=True, common_color_noise=True, common_normal_noise=True, max_empty_border_size_divisor=None,
                        device="cuda"):
    """samples: list of (points [n, 3], colors [n, 3], normals [n, 3]) host or device fp32 tensors.  The per-sample draws
    follow convert_sample (sparse_augmentation.py:250-313): distortion matrix, sub-pixel offset, [random cut-out: dimension
    order and start positions], colour noise, normal noise -- sample by sample, as the reference's loop does.  -> (data
    5-tuple of collate_fn with coords / features on the device, augmentation list, is_inside-equivalent `kept` rows).

    Two modes.  With a fixed cut-out (`shift` / `start`) and common noise vectors no draw depends on the data: all draws are
    made first and the whole batch runs in one pass of the kernels.  The shipped TRAINING configuration
    (scannet_config/run.py:971-984: shift=None, per-point colour noise) draws from data-dependent ranges (random_cut_out) and
    draws one noise value per KEPT point, so the samples are converted one after the other, each with its own kernel passes,
    to keep the reference's order of RNG calls."""
    if (shift is None and start is None) or (color_noise_sigma and not common_color_noise) or (
            normal_noise_sigma and not common_normal_noise):
        return _convert_sequential(samples, spatial_size=spatial_size, scale=scale, shift=shift, start=start,
                                   coord_noise_sigma=coord_noise_sigma, theta=theta, mirror=mirror, sub_pixel_offset=sub_pixel_offset,
                                   color_noise_sigma=color_noise_sigma, normal_noise_sigma=normal_noise_sigma, use_color=use_color,
                                   use_ones=use_ones, use_normal=use_normal, common_color_noise=common_color_noise,
                                   common_normal_noise=common_normal_noise,
                                   max_empty_border_size_divisor=max_empty_border_size_divisor, device=device)
    B = len(samples)
    projs, rots, offs, cshift, nshift = [], [], [], [], []
    for pts, _, _ in samples:
        rot = coord_distortion_matrix(torch.float32, coord_noise_sigma, theta, mirror)
        rots.append(rot), projs.append(rot * scale)
        offs.append(torch.rand((3,), dtype=torch.float32) if sub_pixel_offset is None else torch.as_tensor(sub_pixel_offset, dtype=torch.float32))
        if use_color and color_noise_sigma:
            cshift.append(color_noise_sigma * torch.randn((3,), dtype=torch.float32))
        if use_normal and normal_noise_sigma:
            nshift.append(normal_noise_sigma * torch.randn((3,), dtype=torch.float32))
    dev = torch.device(device)
    cat = lambda k: torch.cat([s[k].to(dev, non_blocking=True) for s in samples]).contiguous()
    ptr = [0]
    for s in samples:
        ptr.append(ptr[-1] + len(s[0]))
    vox = voxelize_batch(cat(0), ptr, torch.stack(projs), torch.stack(offs), spatial_size, shift=None if start is not None else shift,
                         start=start)
    feats = features_batch(vox, B, colors=cat(1) if use_color else None, color_shift=torch.stack(cshift) if cshift else None,
                           use_ones=use_ones, normals=cat(2) if use_normal else None, rotation=torch.stack(rots) if use_normal else None,
                           normal_shift=torch.stack(nshift) if nshift else None)
    size = torch.tensor([int(s) for s in spatial_size], dtype=torch.long)
    data = (vox["coords"], feats, size, B, vox["batch_splits"])
    augmentation = [dict(coords_projection=projs[i], coords_shift=vox["complete_shift"][i],
                         **({"color_shift": cshift[i]} if cshift else {}), **({"normals_shift": nshift[i]} if nshift else {}))
                    for i in range(B)]
    return data, augmentation, vox["kept"]


def _convert_sequential(samples, *, spatial_size, scale, shift, start, coord_noise_sigma, theta, mirror, sub_pixel_offset,
                        color_noise_sigma, normal_noise_sigma, use_color, use_ones, use_normal, common_color_noise,
                        common_normal_noise, max_empty_border_size_divisor, device):
    """convert_sample for one sample after the other (data-dependent draws), collate_fn at the end."""
    dev = torch.device(device)
    size = [int(v) for v in spatial_size]
    border = [0, 0, 0] if max_empty_border_size_divisor is None else [v // max_empty_border_size_divisor for v in size]
    coords_l, feats_l, kept_l, aug_l, splits, at = [], [], [], [], [], 0
    for b, (pts, colors, normals) in enumerate(samples):
        rot = coord_distortion_matrix(torch.float32, coord_noise_sigma, theta, mirror)
        proj = rot * scale
        off = torch.rand((3,), dtype=torch.float32) if sub_pixel_offset is None else torch.as_tensor(sub_pixel_offset, dtype=torch.float32)
        p = pts.to(dev, non_blocking=True).contiguous()
        ptr = [0, len(p)]
        if shift is None and start is None:
            every = voxelize_batch(p, ptr, proj[None], off[None], size, _window=_ACCEPT_ALL)      # the discrete coordinates
            st, _ = draw_random_cut(every["coords"][:, :3], size, border)
            vox = voxelize_batch(p, ptr, proj[None], off[None], size, start=st[None])
        elif start is not None:
            vox = voxelize_batch(p, ptr, proj[None], off[None], size, start=torch.as_tensor(start).reshape(-1, 3)[b][None])
        else:
            vox = voxelize_batch(p, ptr, proj[None], off[None], size, shift=shift)
        n = vox["n"]
        aug = dict(coords_projection=proj, coords_shift=vox["complete_shift"][0])
        c_common = n_common = None
        if use_color and color_noise_sigma and common_color_noise:
            c_common = color_noise_sigma * torch.randn((3,), dtype=torch.float32)
        c_point = color_noise_sigma * torch.randn((n, 3), dtype=torch.float32) if (use_color and color_noise_sigma and not common_color_noise) else None
        if use_normal and normal_noise_sigma and common_normal_noise:
            n_common = normal_noise_sigma * torch.randn((3,), dtype=torch.float32)
        n_point = normal_noise_sigma * torch.randn((n, 3), dtype=torch.float32) if (use_normal and normal_noise_sigma and not common_normal_noise) else None
        f = features_batch(vox, 1, colors=colors.to(dev).contiguous() if use_color else None, color_shift=None if c_common is None else c_common[None],
                           use_ones=use_ones, normals=normals.to(dev).contiguous() if use_normal else None, rotation=rot[None] if use_normal else None,
                           normal_shift=None if n_common is None else n_common[None])
        if c_point is not None:
            f[:, 0:3] += c_point.to(dev)
            aug["color_shift"] = c_point
        elif c_common is not None:
            aug["color_shift"] = c_common
        if n_point is not None:
            c0 = (3 if use_color else 0) + (1 if use_ones else 0)
            f[:, c0:c0 + 3] += n_point.to(dev)
            aug["normals_shift"] = n_point
        elif n_common is not None:
            aug["normals_shift"] = n_common
        c = vox["coords"]
        c[:, 3] = b
        coords_l.append(c), feats_l.append(f), kept_l.append(vox["kept"] + at), aug_l.append(aug), splits.append(n)
        at += len(p)
    data = (torch.cat(coords_l), torch.cat(feats_l), torch.tensor(size, dtype=torch.long), len(samples), splits)
    return data, aug_l, torch.cat(kept_l)
