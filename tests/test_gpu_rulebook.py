"""GPU parity: rulebook builder vs the oracle -- bit-exact (integer / index work)."""
import numpy as np
import pytest
import torch

import scn_oracle as O
from scn_oracle import rules as R
from tests.util import make_pair, random_scene

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("first_rows")]


def _scn():
    from sparse_rcnn_b200 import scn
    return scn


@pytest.mark.parametrize("seed,mode", [(0, 4), (1, 4), (2, 3), (3, 0)])
def test_input_layer_rows_and_locations(cuda, seed, mode):
    scn = _scn()
    coords, feats, size = random_scene(seed, dup=1.0 if mode == 0 else 1.7)
    if mode == 0:   # unique coordinates required
        _, idx = np.unique(R.pack_keys(coords.numpy()), return_index=True)
        idx = np.sort(idx)
        coords, feats = coords[idx], feats[idx]
    to, tg = make_pair(scn, coords, feats, size, cuda, mode=mode)
    assert tg.metadata.level(size).n == to.metadata.grid(size).n
    assert torch.equal(tg.get_spatial_locations(), to.get_spatial_locations())       # first-appearance order
    assert np.array_equal(tg.metadata.point_row.cpu().numpy(), to.metadata.point_row)
    assert tg.batch_size() == to.batch_size()
    loc = tg.get_spatial_locations()
    assert bool((loc[:-1, 3] <= loc[1:, 3]).all())                                   # batch-sorted
    assert torch.allclose(tg.features.cpu(), to.features, rtol=1e-6, atol=1e-6)


def test_input_rule_csr_sorted(cuda):
    scn = _scn()
    coords, feats, size = random_scene(5, dup=3.0)
    _, tg = make_pair(scn, coords, feats, size, cuda)
    md = tg.metadata
    ptr, pts, pr = md.row_ptr.cpu().numpy(), md.row_pts.cpu().numpy(), md.point_row.cpu().numpy()
    assert ptr[0] == 0 and ptr[-1] == len(coords)
    for r in range(0, len(ptr) - 1, 7):
        seg = pts[ptr[r]:ptr[r + 1]]
        assert (np.diff(seg) > 0).all() and (pr[seg] == r).all()


@pytest.mark.parametrize("filter_size", [3, (3, 1, 3), 5])
def test_submanifold_map_bit_exact(cuda, filter_size):
    scn = _scn()
    coords, feats, size = random_scene(7)
    to, tg = make_pair(scn, coords, feats, size, cuda)
    ref = R.rules_to_map(to.metadata.subm_rules(size, filter_size), to.features.shape[0])
    got = tg.metadata.level(size).subm_map(filter_size).cpu().numpy()
    assert got.dtype == np.int32 and np.array_equal(got, ref)


def test_strided_pyramid_bit_exact(cuda):
    scn = _scn()
    coords, feats, size = random_scene(11, size=(32, 32, 16), density=0.05)
    to, tg = make_pair(scn, coords, feats, size, cuda)
    cur = size
    for _ in range(3):
        ok, rules, parent, off = to.metadata.conv_rules(cur, 2, 2)
        r = tg.metadata.strided_rules(cur, 2, 2)
        assert r.out_key == ok
        go, gg = to.metadata.grids[ok], tg.metadata.levels[ok]
        assert gg.n == go.n
        assert np.array_equal(gg.locations().numpy(), go.coords)          # same row numbering
        assert np.array_equal(r.parent_row.cpu().numpy(), parent)
        assert np.array_equal(r.cmap.cpu().numpy(), R.rules_to_map(rules, go.n))
        dref = R.rules_to_map([(p, i) for i, p in rules], len(parent))     # transposed rulebook
        assert np.array_equal(r.dmap.cpu().numpy(), dref)
        loc = gg.locations()
        assert bool((loc[:-1, 3] <= loc[1:, 3]).all())
        cur = torch.tensor(ok)


def test_odd_size_rejected(cuda):
    scn = _scn()
    coords, feats, size = random_scene(3, size=(9, 8, 8))
    _, tg = make_pair(scn, coords, feats, size, cuda)
    with pytest.raises(RuntimeError, match="incompatible"):
        tg.metadata.strided_rules(size, 2, 2)


def test_coordinate_range_checked(cuda):
    scn = _scn()
    coords = torch.tensor([[1, 2, 3, 0], [70000, 0, 0, 0]])
    with pytest.raises(RuntimeError, match="outside"):
        scn.ioLayers.InputLayerFunction.apply(3, scn.Metadata(3), torch.tensor([8, 8, 8]), coords,
                                              torch.zeros(2, 3, device=cuda), 0, 4)


def test_scan_large(cuda):
    from sparse_rcnn_b200.scn.metadata import exclusive_scan
    for n in (0, 1, 4095, 4096, 4097, 1_000_003, 20_000_000):
        x = torch.randint(0, 3, (n,), dtype=torch.int32, device=cuda)
        got = exclusive_scan(x)
        ref = torch.zeros(n + 1, dtype=torch.int64, device=cuda)
        ref[1:] = torch.cumsum(x.long(), 0)
        assert torch.equal(got.long(), ref), n


def test_full_size_scene_properties(cuda):
    """BASELINE-size scene: size-independent properties of the rulebook (no oracle at this size)."""
    scn = _scn()
    from sparse_rcnn_b200.synthetic import make_batch
    coords, feats, size, bs, splits = make_batch(1, 0)
    md = scn.Metadata(3)
    f = scn.ioLayers.InputLayerFunction.apply(3, md, size, coords, feats.to(cuda), bs, 4)
    lvl = md.level(size)
    keys = R.pack_keys(coords.numpy())
    uniq, first = np.unique(keys, return_index=True)
    assert lvl.n == len(uniq)
    loc = lvl.locations().numpy()
    assert np.array_equal(R.pack_keys(loc), keys[np.sort(first)])
    m = lvl.subm_map(3).cpu().numpy()
    n = lvl.n
    assert np.array_equal(m[13], np.arange(n))                      # centre offset = identity
    # symmetry: q = map[o][r] >= 0  <=>  map[26-o][q] == r ; per offset every input row appears at most once
    for o in (0, 5, 12, 20):
        r = np.nonzero(m[o] >= 0)[0]
        assert np.array_equal(m[26 - o][m[o][r]], r)
        assert len(np.unique(m[o][r])) == len(r)
    k_mean = (m >= 0).sum() / n
    assert 5 < k_mean < 27
