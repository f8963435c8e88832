"""Synthetic "ScanNet-shaped" scenes.

Produces the 5-tuple the hot path consumes, in the format of the reference's
`collate_fn` (/root/reference ndsis/data/data.py:88-115):
``(coords [P,4] int64 CPU (x,y,z,b), feats [P,C] fp32, spatial_size [3] long,
batch_size, batch_splits)``.  A scene is a room shell (floor + 4 walls) plus
random furniture boxes (5 faces each) sampled as jittered surface points, so
that the active set has the 2-D-manifold neighbourhood statistics of a scan
(mean 3^3 occupancy ~10-18 instead of 27) and ~1.6-2 points per voxel.
"""
import numpy as np
import torch

# anchor shapes in metres (reference scannet_config/network.py:5-19); used for proposal boxes
ANCHORS_M = np.array([
    [0.3752, 0.3752, 0.4221], [0.6566, 0.6566, 0.5159], [0.6566, 0.6566, 0.9380],
    [0.4221, 0.4221, 1.6415], [0.1876, 1.3132, 1.0318], [0.3283, 0.9849, 1.8291],
    [0.7035, 1.5008, 0.8442], [1.3132, 0.1876, 1.0318], [0.9849, 0.3283, 1.7822],
    [1.5008, 0.7035, 0.8442], [0.8442, 2.1574, 0.3752], [2.1574, 0.8442, 0.3752],
    [2.4857, 1.1256, 1.0318], [1.1256, 2.4857, 1.0318]])
VOXEL_M = 0.0375       # reference scannet_config/run.py:357-370


def _face(rng, origin, u, v, normal, density, jitter):
    """Sample a rectangle origin + a*u + b*v (a,b in [0,1]) with `density` points per voxel^2."""
    area = np.linalg.norm(u) * np.linalg.norm(v)
    n = max(int(area * density), 1)
    a, b = rng.random(n), rng.random(n)
    p = origin[None] + a[:, None] * u[None] + b[:, None] * v[None]
    p = p + rng.normal(0, jitter, (n, 1)) * normal[None]
    return p, np.repeat(normal[None], n, 0)


def _box_faces(lo, hi, with_bottom):
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    d = hi - lo
    ex, ey, ez = np.array([d[0], 0, 0]), np.array([0, d[1], 0]), np.array([0, 0, d[2]])
    faces = [
        (lo, ey, ez, np.array([-1., 0, 0])), (lo + ex, ey, ez, np.array([1., 0, 0])),
        (lo, ex, ez, np.array([0, -1., 0])), (lo + ey, ex, ez, np.array([0, 1., 0])),
        (lo + ez, ex, ey, np.array([0, 0, 1.])),
    ]
    if with_bottom:
        faces.append((lo, ex, ey, np.array([0, 0, -1.])))
    return faces


def make_scene(seed=0, spatial_size=(256, 256, 128), room=(176, 176, 88), room_offset=(32, 32, 8),
               n_furniture=24, density=1.5, jitter=0.1, in_channels=6, scale=1.0):
    """One sample: (coords [P,3] int64, feats [P,C] fp32).  `scale` rescales the room (and
    the furniture) to steer the active-voxel count; scale=1 gives N ~ 150-180k."""
    rng = np.random.default_rng(seed)
    size = np.asarray(spatial_size)
    room = np.minimum(np.asarray(room, float) * scale, size - 2 * 4).astype(float)
    off = np.minimum(np.asarray(room_offset, float), size - room - 1)
    lo, hi = off, off + room
    pts, nrm = [], []
    # room shell: floor + 4 walls, normals pointing inwards
    ex, ey, ez = np.array([room[0], 0, 0]), np.array([0, room[1], 0]), np.array([0, 0, room[2]])
    shell = [(lo, ex, ey, np.array([0, 0, 1.])),
             (lo, ey, ez, np.array([1., 0, 0])), (lo + ex, ey, ez, np.array([-1., 0, 0])),
             (lo, ex, ez, np.array([0, 1., 0])), (lo + ey, ex, ez, np.array([0, -1., 0]))]
    for f in shell:
        p, n = _face(rng, *f, density, jitter)
        pts.append(p), nrm.append(n)
    for _ in range(n_furniture):
        e = rng.uniform(8, 48, 3) * scale
        e = np.minimum(e, room - 2)
        p0 = lo + np.array([rng.uniform(1, room[0] - e[0] - 1), rng.uniform(1, room[1] - e[1] - 1), 0.0])
        for f in _box_faces(p0, p0 + e, with_bottom=False):
            p, n = _face(rng, *f, density, jitter)
            pts.append(p), nrm.append(n)
    p = np.concatenate(pts)
    n = np.concatenate(nrm)
    c = np.floor(p).astype(np.int64)
    keep = ((c >= 0) & (c < size[None])).all(1)
    c, n = c[keep], n[keep]
    perm = rng.permutation(len(c))          # a scan has no spatial order
    c, n = c[perm], n[perm]
    rgb = rng.uniform(-1, 1, (len(c), 3))
    n = n + rng.normal(0, 0.05, n.shape)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    feats = np.concatenate([rgb, n], 1).astype(np.float32)
    if in_channels > 6:
        feats = np.concatenate([feats, np.ones((len(c), in_channels - 6), np.float32)], 1)
    feats = feats[:, :in_channels]
    return torch.from_numpy(c), torch.from_numpy(np.ascontiguousarray(feats))


def make_batch(n_scenes=1, seed=0, spatial_size=(256, 256, 128), **kw):
    """Batch in collate_fn format (data.py:95-107)."""
    coords, feats, splits = [], [], []
    for i in range(n_scenes):
        c, f = make_scene(seed + i, spatial_size, **kw)
        coords.append(torch.nn.functional.pad(c, (0, 1), value=i))
        feats.append(f)
        splits.append(len(c))
    return (torch.cat(coords), torch.cat(feats), torch.tensor(spatial_size, dtype=torch.long),
            n_scenes, splits)


def make_boxes(coords, n_boxes=256, seed=0, spatial_size=(256, 256, 128)):
    """Proposal-like boxes per sample: list over samples of [n_i, 2, 3] fp32 (start, stop) in voxels.
    Centres on random points, edges from the 14 anchor shapes x U[0.8,1.2]."""
    rng = np.random.default_rng(seed + 7919)
    coords = coords.numpy() if isinstance(coords, torch.Tensor) else coords
    n_samples = int(coords[:, 3].max()) + 1 if len(coords) else 1
    out = []
    for b in range(n_samples):
        c = coords[coords[:, 3] == b, :3]
        ctr = c[rng.integers(0, len(c), n_boxes)].astype(np.float64) + 0.5
        edge = ANCHORS_M[rng.integers(0, len(ANCHORS_M), n_boxes)] / VOXEL_M * rng.uniform(0.8, 1.2, (n_boxes, 1))
        box = np.stack([ctr - edge / 2, ctr + edge / 2], 1)
        box = np.clip(box, 0, np.asarray(spatial_size, float)[None, None])
        out.append(torch.from_numpy(box.astype(np.float32)))
    return out


def make_proposals(seed, B, A, scene=(256.0, 256.0, 128.0), clustered=True):
    """RPN-like proposals: A boxes per sample clustered around a few object centres (heavy overlaps) + scores in (0, 1).
    -> (score [B, A] fp32, boxes [B, A, 2, 3] fp32 (start, stop)).  CPU tensors, torch.Generator seeded."""
    import torch
    g = torch.Generator().manual_seed(seed)
    n_obj = 12
    centres = torch.rand(B, n_obj, 3, generator=g) * torch.tensor(scene)
    which = torch.randint(0, n_obj, (B, A), generator=g)
    c = torch.gather(centres, 1, which[..., None].expand(B, A, 3))
    if clustered:
        c = c + torch.randn(B, A, 3, generator=g) * 6.0
    else:
        c = torch.rand(B, A, 3, generator=g) * torch.tensor(scene)
    size = 8.0 + torch.rand(B, A, 3, generator=g) * 40.0
    boxes = torch.stack([c - size / 2, c + size / 2], dim=2)
    # unique scores per sample (a permutation), so that the descending order is the same on every device / library
    score = torch.stack([(torch.randperm(A, generator=g).float() + 0.5) / A for _ in range(B)])
    return score, boxes


def make_mask_loss_case(seed, boxes_per_sample, max_points=400, num_classes=18, empty_every=5):
    """Inputs of the reference's MaskLoss.forward: (masks_output, mask_target, class_target) = lists over samples of lists
    over boxes of [n_b] logits / bool targets, and int64 class ids per box.  Every `empty_every`-th box has no points."""
    import torch
    g = torch.Generator().manual_seed(seed)
    outs, tgts, cls = [], [], []
    k = 0
    for nb in boxes_per_sample:
        so, st = [], []
        for _ in range(nb):
            k += 1
            n = 0 if (empty_every and k % empty_every == 0) else int(torch.randint(1, max_points, (1,), generator=g))
            so.append(torch.randn(n, generator=g) * 4.0)
            st.append(torch.rand(n, generator=g) > 0.6)
        outs.append(so), tgts.append(st)
        cls.append(torch.randint(0, num_classes, (nb,), generator=g))
    return outs, tgts, cls


def make_rpn_outputs(coords, n_anchors=30000, n_objects=256, seed=0, spatial_size=(256, 256, 128), copies=4):
    """Raw region-proposal outputs of one batch as the RoiSelector receives them (model.py:990-1010): -> (score [B, A],
    bbox [B, A, 2, 3]) fp32 CPU.  `n_objects` object boxes per sample (make_boxes: centred on real points, anchor-shaped),
    each present `copies` times with a small jitter (the lower-scored copies overlap the best one with IoU > 0.5, so NMS
    removes them) and the highest scores; the remaining anchors are random boxes with low scores.  With 1024 -> NMS 0.5 ->
    256 selection this yields ~n_objects RoIs per sample, i.e. the mask / class workload of make_boxes."""
    import torch
    rng = np.random.default_rng(seed + 104729)
    obj = make_boxes(coords, n_objects, seed, spatial_size)
    hi = np.asarray(spatial_size, np.float32)
    scores, boxes = [], []
    for b in range(len(obj)):
        base = obj[b].numpy()
        size = base[:, 1] - base[:, 0]
        reps = [base]
        for _ in range(copies - 1):
            j = rng.uniform(-0.02, 0.02, base.shape).astype(np.float32) * size[:, None, :]
            reps.append(np.clip(base + j, 0, hi))
        top = np.concatenate(reps)
        n_fill = n_anchors - len(top)
        c = rng.uniform(0, 1, (n_fill, 3)).astype(np.float32) * hi
        e = rng.uniform(4, 40, (n_fill, 3)).astype(np.float32)
        fill = np.clip(np.stack([c - e / 2, c + e / 2], 1), 0, hi)
        bx = np.concatenate([top, fill]).astype(np.float32)
        # unique scores: copy k of object i scores above every copy k+1, all above the filler
        sc = np.empty(n_anchors, np.float32)
        order = rng.permutation(n_objects)
        for k in range(copies):
            sc[k * n_objects:(k + 1) * n_objects] = 1.0 - (k * n_objects + order + 0.5) / (2.0 * copies * n_objects)
        sc[len(top):] = 0.4 * (rng.permutation(n_fill) + 0.5) / n_fill
        perm = rng.permutation(n_anchors)
        scores.append(sc[perm]), boxes.append(bx[perm])
    return torch.from_numpy(np.stack(scores)), torch.from_numpy(np.stack(boxes))
