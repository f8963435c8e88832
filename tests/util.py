"""Shared helpers: build the same sparse tensor on the oracle (CPU) and on the B200 backend."""
import numpy as np
import torch

import scn_oracle as O
from scn_oracle import rules as R


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    if a.numel() == 0 and b.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-12))


def random_scene(seed, size=(24, 20, 16), n_samples=2, density=0.08, dup=1.6, channels=5):
    """Random clustered voxels with duplicates, grouped by sample like collate_fn."""
    rng = np.random.default_rng(seed)
    coords = []
    for b in range(n_samples):
        n = int(np.prod(size) * density)
        c = np.stack([rng.integers(0, s, n) for s in size], 1)
        # make surfaces-ish: snap one axis for half of the points
        c[: n // 2, 2] = size[2] // 3
        n_pts = int(n * dup)
        pick = rng.integers(0, n, n_pts)
        c = c[pick]
        coords.append(np.concatenate([c, np.full((n_pts, 1), b)], 1))
    coords = torch.from_numpy(np.concatenate(coords)).long()
    feats = torch.from_numpy(rng.standard_normal((len(coords), channels)).astype(np.float32))
    return coords, feats, torch.tensor(size, dtype=torch.long)


def make_pair(scn, coords, feats, size, device, mode=4, batch_size=0):
    """-> (oracle tensor, backend tensor) from the same points."""
    mo = O.Metadata(3)
    fo = O.ioLayers.InputLayerFunction.apply(3, mo, size, coords, feats, batch_size, mode)
    to = O.SparseConvNetTensor(fo, mo, size)
    mg = scn.Metadata(3)
    fg = scn.ioLayers.InputLayerFunction.apply(3, mg, size, coords, feats.to(device), batch_size, mode)
    tg = scn.SparseConvNetTensor(fg, mg, size)
    return to, tg


def copy_params(src, dst, device=None):
    sd = {k: v.clone() for k, v in src.state_dict().items()}
    dst.load_state_dict(sd)
    if device is not None:
        dst.to(device)
    return dst


def reinit_by_name(module, scale=1.0):
    """Deterministic, construction-order independent weights: every parameter is redrawn from a
    generator seeded by the crc32 of its state_dict key, keeping the std of its initialiser
    (biases: 0.05).  Lets goldens be regenerated without committing megabytes of weights.
    std depends only on the shape: sqrt(2 / (numel / shape[-1])) for matrices, 0.05 for vectors."""
    import zlib
    with torch.no_grad():
        for k, v in sorted(module.state_dict().items()):
            if not v.is_floating_point():
                continue
            g = torch.Generator().manual_seed(zlib.crc32(k.encode()))
            std = (2.0 / max(v.numel() // v.shape[-1], 1)) ** 0.5 if v.dim() >= 2 else 0.05
            v.copy_(torch.randn(v.shape, generator=g) * std * scale)
    return module


def canon_order(loc):
    """locations [n,4] (x,y,z,b) -> LongTensor of rows in canonical (b, x, y, z) order."""
    keys = R.pack_keys(loc.numpy() if isinstance(loc, torch.Tensor) else loc)
    return torch.from_numpy(np.argsort(keys, kind="stable"))


def canon_features(t):
    """(sorted keys, features in canonical row order) of a SparseConvNetTensor from either backend."""
    loc = t.get_spatial_locations()
    order = canon_order(loc)
    keys = np.sort(R.pack_keys(loc.numpy()))
    return keys, t.features.detach().cpu()[order]
