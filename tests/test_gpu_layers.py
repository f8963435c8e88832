"""GPU parity: IO layers, pooling, sparse-to-dense, elementwise and BatchNorm vs the oracle."""
import pytest
import torch

import scn_oracle as O
from tests.util import copy_params, make_pair, random_scene, rel_err

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("first_rows")]


def _scn():
    from sparse_rcnn_b200 import scn
    scn.set_precision("fp32")
    return scn


def _fb(fo_fn, fg_fn, xo, xg, tol=1e-5):
    xo = xo.clone().requires_grad_(True)
    xg = xg.clone().requires_grad_(True)
    yo, yg = fo_fn(xo), fg_fn(xg)
    assert yo.shape == yg.shape
    assert rel_err(yg, yo) <= tol, ("fwd", rel_err(yg, yo))
    g = torch.randn_like(yo)
    yo.backward(g)
    yg.backward(g.to(yg.device))
    assert rel_err(xg.grad, xo.grad) <= tol, ("bwd", rel_err(xg.grad, xo.grad))


@pytest.mark.parametrize("mode", [1, 2, 3, 4])
def test_input_layer_modes_fwd_bwd(cuda, mode):
    scn = _scn()
    coords, feats, size = random_scene(20 + mode, dup=2.5, channels=6)

    def fo(x):
        return O.ioLayers.InputLayerFunction.apply(3, O.Metadata(3), size, coords, x, 0, mode)

    def fg(x):
        return scn.ioLayers.InputLayerFunction.apply(3, scn.Metadata(3), size, coords, x, 0, mode)

    _fb(fo, fg, feats, feats.to(cuda))


def test_output_layer_fwd_bwd(cuda):
    scn = _scn()
    coords, feats, size = random_scene(30, dup=2.0, channels=4)
    to, tg = make_pair(scn, coords, feats, size, cuda)
    x = torch.randn(to.features.shape[0], 9)
    _fb(lambda v: O.ioLayers.OutputLayerFunction.apply(3, to.metadata, v),
        lambda v: scn.ioLayers.OutputLayerFunction.apply(3, tg.metadata, v), x, x.to(cuda))
    out = scn.OutputLayer(3)(tg)
    assert out.shape[0] == len(coords)


@pytest.mark.parametrize("kind", ["max", "avg"])
def test_pooling_fwd_bwd(cuda, kind):
    scn = _scn()
    coords, feats, size = random_scene(40, channels=8)
    feats = torch.relu(feats)            # zeros => ties in max pooling (backward routes to every tie)
    to, tg = make_pair(scn, coords, feats, size, cuda)
    po = (O.MaxPooling if kind == "max" else O.AveragePooling)(3, 2, 2)
    pg = (scn.MaxPooling if kind == "max" else scn.AveragePooling)(3, 2, 2)
    _fb(lambda v: po(O.SparseConvNetTensor(v, to.metadata, size)).features,
        lambda v: pg(scn.SparseConvNetTensor(v, tg.metadata, size)).features, to.features, tg.features)


def test_unpooling_fwd_bwd(cuda):
    """scn.UnPooling (named by the north star; transpose of sum pooling) against the oracle, forward and backward, and the
    error when the finer level does not exist."""
    scn = _scn()
    coords, feats, size = random_scene(41, channels=8)
    to, tg = make_pair(scn, coords, feats, size, cuda)
    po, pg = O.AveragePooling(3, 2, 2)(to), scn.AveragePooling(3, 2, 2)(tg)
    _fb(lambda v: O.UnPooling(3, 2, 2)(O.SparseConvNetTensor(v, to.metadata, po.spatial_size)).features,
        lambda v: scn.UnPooling(3, 2, 2)(scn.SparseConvNetTensor(v, tg.metadata, pg.spatial_size)).features,
        po.features.detach(), pg.features.detach())
    up = scn.UnPooling(3, 2, 2)(pg)
    assert torch.equal(up.get_spatial_locations(), tg.get_spatial_locations())
    with pytest.raises(RuntimeError):
        scn.UnPooling(3, 2, 2)(tg)                     # nothing finer than the input level


def test_sparse_to_dense_fwd_bwd(cuda):
    scn = _scn()
    coords, feats, size = random_scene(50, channels=7, n_samples=3)
    to, tg = make_pair(scn, coords, feats, size, cuda)
    _fb(lambda v: O.SparseToDense(3, 7)(O.SparseConvNetTensor(v, to.metadata, size)),
        lambda v: scn.SparseToDense(3, 7)(scn.SparseConvNetTensor(v, tg.metadata, size)), to.features, tg.features, 0.0)
    # after a strided level
    co, cg = O.Convolution(3, 7, 7, 2, 2, False), None
    cg = copy_params(co, scn.Convolution(3, 7, 7, 2, 2, False), cuda)
    yo, yg = co(to), cg(tg)
    d_o, d_g = O.SparseToDense(3, 7)(yo), scn.SparseToDense(3, 7)(yg)
    assert d_o.shape == d_g.shape and rel_err(d_g, d_o) <= 1e-5


def test_relu_add_join(cuda):
    scn = _scn()
    coords, feats, size = random_scene(60, channels=10)
    to, tg = make_pair(scn, coords, feats, size, cuda)
    _fb(lambda v: O.ReLU()(O.SparseConvNetTensor(v, to.metadata, size)).features,
        lambda v: scn.ReLU()(scn.SparseConvNetTensor(v, tg.metadata, size)).features, to.features, tg.features, 0.0)
    T = lambda v: scn.SparseConvNetTensor(v, tg.metadata, size)
    a, b = tg.features, torch.randn_like(tg.features)
    assert torch.equal(scn.AddTable()([T(a), T(b)]).features, a + b)
    assert torch.equal(scn.JoinTable()([T(a), T(b)]).features, torch.cat([a, b], 1))


@pytest.mark.parametrize("leak", [0.0, 0.2])
@pytest.mark.parametrize("training", [True, False])
def test_batchnorm(cuda, leak, training):
    scn = _scn()
    coords, feats, size = random_scene(70, channels=12)
    to, tg = make_pair(scn, coords, feats, size, cuda)
    bo = O.BatchNormLeakyReLU(12, 1e-4, 0.9, leak) if leak else O.BatchNormReLU(12, 1e-4, 0.9)
    bo.weight.data.uniform_(0.5, 1.5), bo.bias.data.normal_()
    bo.running_mean.normal_(), bo.running_var.uniform_(0.5, 2)
    bg = scn.BatchNormLeakyReLU(12, 1e-4, 0.9, leak) if leak else scn.BatchNormReLU(12, 1e-4, 0.9)
    copy_params(bo, bg, cuda)
    bo.train(training), bg.train(training)
    _fb(lambda v: bo(O.SparseConvNetTensor(v, to.metadata, size)).features,
        lambda v: bg(scn.SparseConvNetTensor(v, tg.metadata, size)).features, to.features, tg.features, 2e-5)
    assert rel_err(bg.running_mean, bo.running_mean) <= 1e-5 and rel_err(bg.running_var, bo.running_var) <= 1e-5
    assert rel_err(bg.weight.grad, bo.weight.grad) <= 1e-4 and rel_err(bg.bias.grad, bo.bias.grad) <= 1e-4
