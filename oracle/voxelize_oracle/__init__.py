"""TEST INFRASTRUCTURE ONLY (never imported by the product): CPU restatement (numpy) of the deterministic core of the
reference's per-sample conversion and batch collation, the oracle for sparse_rcnn_b200/csrc/voxelize.cu.

Follows /root/reference/ndsis/data/sparse_augmentation.py:
  * augment_coords :81-126   project (coords @ (almost_orthonormal * scale)), shift = -min + sub_pixel_offset, `.long()`
  * fix_cut_out    :41-46    inside test on the UNMOVED discrete coordinate against [0, size), move by `shifts`
  * random_cut_out :49-78    (given its drawn start positions) inside = start <= c < start + size, move by -start
  * augment_single_feature / augment_features :129-190   rows that stay, normals @ rotation, optional common noise vector
and ndsis/data/data.py:88-115 (collate_fn): batch-index column, concatenation, batch_splits.

Pinned against goldens produced by the unmodified reference functions (oracle/make_golden_voxelize.py ->
tests/golden/voxelize.pt; tests/test_voxelize_oracle.py).  fp32 arithmetic restated exactly: torch's CPU matmul for a
[P,3] x [3,3] product is one rounded product followed by two fused multiply-adds per output (emulated in float64, which
holds an fp32 product exactly)."""
import numpy as np


def project(x, m):
    """x [P,3] f32, m [3,3] f32 -> x @ m as torch's CPU kernel rounds it."""
    x, m = np.asarray(x, np.float32), np.asarray(m, np.float32)
    x64, m64 = x.astype(np.float64), m.astype(np.float64)
    t = (x[:, 0:1] * m[0:1, :]).astype(np.float32)
    t = (x64[:, 1:2] * m64[1:2, :] + t.astype(np.float64)).astype(np.float32)      # fma: exact product, one rounding
    t = (x64[:, 2:3] * m64[2:3, :] + t.astype(np.float64)).astype(np.float32)
    return t


def voxelize_sample(points, proj, offset, size, shift=None, start=None):
    """-> (coords int64 [P', 3], is_inside bool [P], complete_shift f32 [3]).  Exactly one of shift (fix_cut_out) / start
    (random_cut_out with these start positions) is given."""
    aug = project(points, proj)
    if len(aug) == 0:
        return np.zeros((0, 3), np.int64), np.zeros(0, bool), np.zeros(3, np.float32)
    complete_shift = (-aug.min(0)).astype(np.float32) + np.asarray(offset, np.float32)
    disc = np.trunc((aug + complete_shift).astype(np.float32)).astype(np.int64)
    size = np.asarray(size, np.int64)
    if shift is not None:
        shift = np.broadcast_to(np.asarray(shift, np.int64), (3,))
        inside = ((disc >= 0) & (disc < size)).all(1)
        return (disc + shift)[inside], inside, complete_shift
    start = np.asarray(start, np.int64)
    rel = disc - start
    inside = ((rel >= 0) & (rel < size)).all(1)
    return rel[inside], inside, complete_shift


def features_sample(inside, colors=None, color_shift=None, use_ones=False, normals=None, rotation=None, normal_shift=None):
    n = int(inside.sum())
    parts = []
    if colors is not None:
        c = np.asarray(colors, np.float32)[inside]
        parts.append(c if color_shift is None else (c + np.asarray(color_shift, np.float32)).astype(np.float32))
    if use_ones:
        parts.append(np.ones((n, 1), np.float32))
    if normals is not None:
        v = np.asarray(normals, np.float32)[inside]
        if rotation is not None:
            v = project(v, rotation)
        parts.append(v if normal_shift is None else (v + np.asarray(normal_shift, np.float32)).astype(np.float32))
    return np.concatenate(parts, 1) if parts else np.zeros((n, 0), np.float32)


def collate(coords_list, features_list):
    """collate_fn: coords [sum P', 4] int64 with the sample index in the last column, features concatenated, batch_splits."""
    coords = np.concatenate([np.concatenate([c, np.full((len(c), 1), i, np.int64)], 1) for i, c in enumerate(coords_list)])
    return coords, np.concatenate(features_list), [len(c) for c in coords_list]


def random_cut(disc, size, border, order, draws):
    """random_cut_out :49-78 with its draws given: `order` = the drawn order of the dimensions, `draws` = an iterator of
    callables (lo, hi) -> int consumed only where the reference calls randint.  -> (start int64 [3], is_inside bool [P])."""
    disc = np.asarray(disc, np.int64)
    start = np.zeros(3, np.int64)
    inside = np.ones(len(disc), bool)
    for dim in order:
        if not inside.any():
            break
        vals = disc[inside, dim]
        lo = int(vals.min()) - int(border[dim])
        hi = int(vals.max()) + 1 - int(size[dim]) + int(border[dim])
        if hi <= lo:
            start[dim] = lo
        else:
            start[dim] = next(draws)(lo, hi)
            rel = disc[:, dim] - start[dim]
            inside &= (rel >= 0) & (rel < int(size[dim]))
    return start, inside
