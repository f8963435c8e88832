"""CPU: pin the oracle against INDEPENDENT known answers -- dense torch ops.
SparseConvNet itself is absent (parity unpinned against it); these equivalences are what the
oracle's restatement of its semantics is anchored on."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import scn_oracle as O
from scn_oracle import rules as R
from tests.util import random_scene


def _tensor(seed, channels, size=(12, 10, 8), n_samples=2):
    coords, feats, size = random_scene(seed, size=size, n_samples=n_samples, density=0.15, channels=channels)
    md = O.Metadata(3)
    f = O.ioLayers.InputLayerFunction.apply(3, md, size, coords, feats, 0, 4)
    return O.SparseConvNetTensor(f, md, size), coords, feats


def _at(dense, loc):
    return dense[loc[:, 3], :, loc[:, 0], loc[:, 1], loc[:, 2]]


@pytest.mark.parametrize("fs", [3, 1, (3, 1, 3)])
def test_submanifold_equals_masked_conv3d(fs):
    torch.manual_seed(0)
    t, _, _ = _tensor(1, 5)
    conv = O.SubmanifoldConvolution(3, 5, 7, fs, True)
    conv.bias.data.normal_()
    y = conv(t)
    f3 = conv.filter_size
    W = conv.weight.view(*f3, 5, 7).permute(4, 3, 0, 1, 2)
    yd = F.conv3d(O.SparseToDense(3, 5)(t), W, conv.bias, padding=tuple(f // 2 for f in f3))
    assert torch.allclose(_at(yd, t.get_spatial_locations()), y.features, atol=1e-5)


def test_strided_conv_deconv_equal_dense():
    torch.manual_seed(1)
    t, _, _ = _tensor(2, 4)
    cv = O.Convolution(3, 4, 6, 2, 2, True)
    cv.bias.data.normal_()
    z = cv(t)
    d = O.SparseToDense(3, 4)(t)
    zd = F.conv3d(d, cv.weight.view(2, 2, 2, 4, 6).permute(4, 3, 0, 1, 2), None, stride=2)
    loc2 = z.get_spatial_locations()
    assert torch.allclose(_at(zd, loc2) + cv.bias, z.features, atol=1e-5)
    # active iff >= 1 active child
    occ = F.max_pool3d((d.abs().sum(1, keepdim=True) > 0).float(), 2)
    assert int(occ.sum()) == z.features.shape[0]
    dc = O.Deconvolution(3, 6, 4, 2, 2, True)
    dc.bias.data.normal_()
    u = dc(z)
    ud = F.conv_transpose3d(O.SparseToDense(3, 6)(z), dc.weight.view(2, 2, 2, 6, 4).permute(3, 4, 0, 1, 2), None, stride=2)
    assert torch.allclose(_at(ud, t.get_spatial_locations()) + dc.bias, u.features, atol=1e-5)
    assert torch.equal(u.get_spatial_locations(), t.get_spatial_locations())


def test_pooling_equals_dense():
    t, _, _ = _tensor(3, 3)
    loc = t.get_spatial_locations()
    d = O.SparseToDense(3, 3)(t)
    mask = torch.zeros(d.shape[0], 1, *d.shape[2:])
    mask[loc[:, 3], 0, loc[:, 0], loc[:, 1], loc[:, 2]] = 1
    mp = O.MaxPooling(3, 2, 2)(t)
    dense_max = F.max_pool3d(torch.where(mask.bool(), d, torch.full_like(d, -np.inf)), 2)
    assert torch.allclose(_at(dense_max, mp.get_spatial_locations()), mp.features)
    ap = O.AveragePooling(3, 2, 2)(t)
    assert torch.allclose(_at(F.avg_pool3d(d, 2), ap.get_spatial_locations()), ap.features, atol=1e-6)


def test_unpooling_equals_masked_nearest_upsample():
    """UnPooling(pool(x)) puts the pooled value on every ACTIVE site of the finer level: the dense equivalent is a nearest
    upsample masked to the active set; it is also the transpose of sum pooling (<unpool(c), y> == <c, 8 * avgpool(y)>)."""
    t, _, _ = _tensor(5, 4)
    ap = O.AveragePooling(3, 2, 2)(t)
    up = O.UnPooling(3, 2, 2)(ap)
    assert torch.equal(up.get_spatial_locations(), t.get_spatial_locations())
    dense = F.interpolate(O.SparseToDense(3, 4)(ap), scale_factor=2, mode="nearest")
    assert torch.allclose(_at(dense, t.get_spatial_locations()), up.features)
    c = torch.randn_like(ap.features)
    lhs = (O.UnPooling(3, 2, 2)(O.SparseConvNetTensor(c, ap.metadata, ap.spatial_size)).features * t.features).sum()
    assert torch.allclose(lhs, (c * ap.features * 8).sum(), rtol=1e-5)


def test_input_layer_mode4_is_unique_mean_in_first_appearance_order():
    t, coords, feats = _tensor(4, 6)
    keys = R.pack_keys(coords.numpy())
    uniq, first, inv = np.unique(keys, return_index=True, return_inverse=True)
    order = np.argsort(first)
    mean = torch.zeros(len(uniq), 6).index_add_(0, torch.from_numpy(inv.reshape(-1)), feats)
    mean /= torch.bincount(torch.from_numpy(inv.reshape(-1))).float()[:, None]
    assert torch.allclose(t.features, mean[order], atol=1e-6)
    assert np.array_equal(R.pack_keys(t.get_spatial_locations().numpy()), uniq[order])
    # output layer is the inverse rule: every point gets its voxel's row
    out = O.OutputLayer(3)(t)
    assert torch.allclose(out, t.features[torch.from_numpy(t.metadata.point_row)])


def test_autograd_of_oracle_conv_matches_dense():
    torch.manual_seed(5)
    t, _, _ = _tensor(5, 3)
    conv = O.SubmanifoldConvolution(3, 3, 4, 3, True)
    x = t.features.clone().requires_grad_(True)
    y = conv(O.SparseConvNetTensor(x, t.metadata, t.spatial_size)).features
    g = torch.randn_like(y)
    y.backward(g)
    loc = t.get_spatial_locations()
    x2 = t.features.clone().requires_grad_(True)
    dense = torch.zeros(2, *[int(s) for s in t.spatial_size], 3)
    dense = dense.index_put((loc[:, 3], loc[:, 0], loc[:, 1], loc[:, 2]), x2).permute(0, 4, 1, 2, 3)
    w2 = conv.weight.detach().clone().requires_grad_(True)
    yd = F.conv3d(dense, w2.view(3, 3, 3, 3, 4).permute(4, 3, 0, 1, 2), conv.bias.detach(), padding=1)
    _at(yd, loc).backward(g)
    assert torch.allclose(x.grad, x2.grad, atol=1e-5)
    assert torch.allclose(conv.weight.grad, w2.grad, atol=1e-4)


def test_rules_to_map_roundtrip_and_offset_order():
    t, _, _ = _tensor(6, 2)
    rules = t.metadata.subm_rules(t.spatial_size, 3)
    m = R.rules_to_map(rules, t.features.shape[0])
    loc = t.get_spatial_locations().numpy()
    # offset index = (dx*3+dy)*3+dz with z fastest: offset 14 is +1 in z
    r = np.nonzero(m[14] >= 0)[0]
    assert (loc[m[14][r], 2] - loc[r, 2] == 1).all() and (loc[m[14][r], :2] == loc[r, :2]).all()
    r = np.nonzero(m[22] >= 0)[0]   # +1 in x
    assert (loc[m[22][r], 0] - loc[r, 0] == 1).all()
    assert (m[13] == np.arange(m.shape[1])).all()


def test_ragged_batch_with_an_empty_sample_and_forced_batch_size():
    """Edge cases the reference relies on (roi_select_sparse.py:75-84,113-122: one sample per box, empty boxes allowed, the
    batch_size argument forces the sample count): a batch whose middle sample has no points keeps batch-sorted rows, the
    forced batch size is reported, and a submanifold convolution never mixes samples (same answer as per-sample runs)."""
    coords, feats, size = random_scene(9, size=(10, 8, 6), n_samples=3, density=0.2, channels=4)
    keep = coords[:, 3] != 1                                          # sample 1 becomes empty
    coords, feats = coords[keep], feats[keep]
    md = O.Metadata(3)
    f = O.ioLayers.InputLayerFunction.apply(3, md, size, coords, feats, 5, 4)       # 5 samples forced, 3 and 4 empty too
    t = O.SparseConvNetTensor(f, md, size)
    loc = t.get_spatial_locations()
    assert t.batch_size() == 5
    assert bool((loc[1:, 3] >= loc[:-1, 3]).all()) and set(loc[:, 3].tolist()) == {0, 2}
    torch.manual_seed(0)
    conv = O.SubmanifoldConvolution(3, 4, 5, 3, True)
    whole = conv(t).features
    for b in (0, 2):
        sel = coords[:, 3] == b
        mdb = O.Metadata(3)
        cb = coords[sel].clone()
        cb[:, 3] = 0
        fb = O.ioLayers.InputLayerFunction.apply(3, mdb, size, cb, feats[sel], 0, 4)
        part = conv(O.SparseConvNetTensor(fb, mdb, size)).features
        assert torch.allclose(whole[loc[:, 3] == b], part, atol=1e-6)


def test_strided_convolution_rejects_odd_sizes():
    """SparseConvNet asserts (out - 1) * stride + filter == in per dimension: an odd grid must raise, not
    silently drop a border."""
    coords, feats, size = random_scene(2, size=(9, 8, 6), n_samples=1, density=0.2, channels=3)
    md = O.Metadata(3)
    f = O.ioLayers.InputLayerFunction.apply(3, md, size, coords, feats, 0, 4)
    with pytest.raises(Exception):
        O.Convolution(3, 3, 4, 2, 2, False)(O.SparseConvNetTensor(f, md, size))
