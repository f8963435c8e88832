"""GPU parity: convolution family (SubM / Convolution / Deconvolution / NiN) forward and backward
vs the oracle.  fp32 verification mode: rel <= 1e-5; TF32 tcgen05 path: rel <= 2e-3 (north_star)."""
import pytest
import torch

import scn_oracle as O
from tests.util import copy_params, make_pair, random_scene, rel_err

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("first_rows")]
TOL = {"fp32": 1e-5, "tf32": 2e-3}


def _scn(precision):
    from sparse_rcnn_b200 import scn
    scn.set_precision(precision)
    return scn


def _run(layer_o, layer_g, to, tg, tol):
    xo = to.features.clone().requires_grad_(True)
    xg = tg.features.clone().requires_grad_(True)
    yo = layer_o(O.SparseConvNetTensor(xo, to.metadata, to.spatial_size))
    yg = layer_g(type(tg)(xg, tg.metadata, tg.spatial_size))
    fo = yo.features if hasattr(yo, "features") else yo
    fg = yg.features if hasattr(yg, "features") else yg
    assert fo.shape == fg.shape
    assert rel_err(fg, fo) <= tol, ("fwd", rel_err(fg, fo))
    go = torch.randn_like(fo)
    fo.backward(go)
    fg.backward(go.to(fg.device))
    assert rel_err(xg.grad, xo.grad) <= tol, ("dx", rel_err(xg.grad, xo.grad))
    for (n, po), (_, pg) in zip(layer_o.named_parameters(), layer_g.named_parameters()):
        assert rel_err(pg.grad, po.grad) <= 5 * tol, (n, rel_err(pg.grad, po.grad))
    return yo, yg


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
@pytest.mark.parametrize("cin,cout,fs", [(5, 7, 3), (32, 32, 3), (6, 32, 1), (48, 80, 3), (22, 22, 3), (44, 24, 3),
                                         (128, 128, 3), (16, 256, 3), (7, 18, 3)])
def test_submanifold(cuda, precision, cin, cout, fs):
    scn = _scn(precision)
    torch.manual_seed(0)
    coords, feats, size = random_scene(cin * 31 + cout, channels=cin)
    to, tg = make_pair(scn, coords, feats, size, cuda)
    lo = O.SubmanifoldConvolution(3, cin, cout, fs, True)
    lo.bias.data.normal_()
    lg = copy_params(lo, scn.SubmanifoldConvolution(3, cin, cout, fs, True), cuda)
    _run(lo, lg, to, tg, TOL[precision])


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
@pytest.mark.parametrize("cin,cout", [(32, 48), (22, 32), (6, 16)])
def test_strided_conv_and_deconv(cuda, precision, cin, cout):
    scn = _scn(precision)
    torch.manual_seed(1)
    coords, feats, size = random_scene(cin + cout, size=(24, 20, 16), channels=cin)
    to, tg = make_pair(scn, coords, feats, size, cuda)
    lo = O.Convolution(3, cin, cout, 2, 2, True)
    lo.bias.data.normal_()
    lg = copy_params(lo, scn.Convolution(3, cin, cout, 2, 2, True), cuda)
    yo, yg = _run(lo, lg, to, tg, TOL[precision])
    assert tuple(yg.spatial_size.tolist()) == tuple(yo.spatial_size.tolist())
    do = O.Deconvolution(3, cout, cin, 2, 2, True)
    do.bias.data.normal_()
    dg = copy_params(do, scn.Deconvolution(3, cout, cin, 2, 2, True), cuda)
    yo2 = O.SparseConvNetTensor(yo.features.detach(), yo.metadata, yo.spatial_size)
    yg2 = type(yg)(yo.features.detach().to(cuda), yg.metadata, yg.spatial_size)
    _run(do, dg, yo2, yg2, TOL[precision])


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_network_in_network(cuda, precision):
    scn = _scn(precision)
    torch.manual_seed(2)
    coords, feats, size = random_scene(9, channels=44)
    to, tg = make_pair(scn, coords, feats, size, cuda)
    lo = O.NetworkInNetwork(44, 22, True)
    lo.bias.data.normal_()
    lg = copy_params(lo, scn.NetworkInNetwork(44, 22, True), cuda)
    _run(lo, lg, to, tg, TOL[precision])


def test_no_bias_and_empty_rows(cuda):
    scn = _scn("tf32")
    coords, feats, size = random_scene(4, channels=16)
    to, tg = make_pair(scn, coords, feats, size, cuda)
    lo = O.SubmanifoldConvolution(3, 16, 16, 3, False)
    lg = copy_params(lo, scn.SubmanifoldConvolution(3, 16, 16, 3, False), cuda)
    _run(lo, lg, to, tg, TOL["tf32"])


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_large_layer_linearity(cuda, precision):
    """BASELINE-size single layer (config 3): conv(a*x + y) == a*conv(x) + conv(y) and parity of the
    two precisions with each other (no oracle at this size)."""
    scn = _scn(precision)
    from sparse_rcnn_b200.synthetic import make_batch
    coords, feats, size, bs, _ = make_batch(1, 0)
    torch.manual_seed(3)
    md = scn.Metadata(3)
    f6 = scn.ioLayers.InputLayerFunction.apply(3, md, size, coords, feats.to(cuda), bs, 4)
    n = f6.shape[0]
    conv = scn.SubmanifoldConvolution(3, 32, 32, 3, False).to(cuda)
    x, y = torch.randn(n, 32, device=cuda), torch.randn(n, 32, device=cuda)
    T = lambda f: scn.SparseConvNetTensor(f, md, size)
    with torch.no_grad():
        cx, cy, cxy = conv(T(x)).features, conv(T(y)).features, conv(T(2.5 * x + y)).features
        assert rel_err(cxy, 2.5 * cx + cy) <= (1e-5 if precision == "fp32" else 4e-3)
        scn.set_precision("fp32")
        ref = conv(T(x)).features
        assert rel_err(cx, ref) <= TOL[precision]


def test_split_mode_fallback_equals_cluster_mode(cuda, monkeypatch):
    """Small levels: the cluster/DSMEM reduction (default) and the atomics fallback (SCN_CONV_NOCLUSTER=1: self-cleaning
    accumulation buffer + last-arriver epilogue) give the same result; the fallback leaves its buffer zeroed (second call)."""
    from sparse_rcnn_b200 import scn
    from sparse_rcnn_b200.synthetic import make_batch
    scn.set_precision("tf32")
    torch.manual_seed(0)
    coords, feats, size, bs, _ = make_batch(1, 3, spatial_size=(64, 64, 32), room=(40, 40, 20), room_offset=(8, 8, 4),
                                            n_furniture=3)
    md = scn.Metadata(3)
    f = scn.ioLayers.InputLayerFunction.apply(3, md, size, coords, feats.to(cuda), bs, 4)
    n = f.shape[0]
    assert n // 128 < 148                                            # few tiles => split-offset mode
    conv = scn.SubmanifoldConvolution(3, 48, 48, 3, True).to(cuda)
    x = scn.SparseConvNetTensor(torch.randn(n, 48, device=cuda), md, size)
    with torch.no_grad():
        a = conv(x).features.clone()
        monkeypatch.setenv("SCN_CONV_NOCLUSTER", "1")
        b1 = conv(x).features.clone()
        b2 = conv(x).features.clone()
        monkeypatch.delenv("SCN_CONV_NOCLUSTER")
        a2 = conv(x).features
    assert torch.equal(a, a2)                                        # cluster mode is deterministic
    assert rel_err(b1, a) < 1e-5 and rel_err(b2, a) < 1e-5           # fallback: same sums, atomics order differs


@pytest.mark.gpu
@pytest.mark.parametrize("channels", [32, 64, 48, 128, 22, 7])
def test_row_skipping_is_bit_exact(cuda, monkeypatch, channels):
    """Producers skip a gathered row that is inactive now and was inactive in the stage's previous use (conv_tc.cu);
    SCN_CONV_SKIP=0 fills every row as before.  Same arithmetic on the same operands => identical bits, on a level with
    many tiles per CTA (stages are reused hundreds of times) with full (32, 64, 128) and partial (48) channel blocks and the 8- / 4-byte copy paths (22, 7)."""
    from sparse_rcnn_b200 import scn
    from sparse_rcnn_b200.synthetic import make_batch
    scn.set_precision("tf32")
    torch.manual_seed(1)
    coords, feats, size, bs, _ = make_batch(1, 5, spatial_size=(256, 256, 128))
    md = scn.Metadata(3)
    f = scn.ioLayers.InputLayerFunction.apply(3, md, size, coords, feats.to(cuda), bs, 4)
    n = f.shape[0]
    assert n // 128 > 3 * 148                                        # several tiles per persistent CTA
    conv = scn.SubmanifoldConvolution(3, channels, channels, 3, True).to(cuda)
    x = scn.SparseConvNetTensor(torch.randn(n, channels, device=cuda), md, size)
    monkeypatch.setenv("SCN_CONV_TAILSPLIT", "0")                     # its atomics reorder fp32 sums; tested on its own below
    with torch.no_grad():
        monkeypatch.setenv("SCN_CONV_SKIP", "1")
        a = conv(x).features.clone()
        monkeypatch.setenv("SCN_CONV_SKIP", "0")
        b = conv(x).features.clone()
        monkeypatch.setenv("SCN_CONV_SKIP", "1")
        a2 = conv(x).features
    assert torch.equal(a, b) and torch.equal(a, a2)
    # backward: the input gradient runs the same kernel (bit-exact); the weight-gradient kernel skips rows the same way
    # but ends with atomics into gw, so its two runs agree to fp32 summation order
    go = torch.randn(n, channels, device=cuda)

    def grads():
        xin = x.features.clone().requires_grad_(True)
        conv.zero_grad()
        conv(scn.SparseConvNetTensor(xin, md, size)).features.backward(go)
        return xin.grad.clone(), conv.weight.grad.clone(), conv.bias.grad.clone()
    gx_a, gw_a, gb_a = grads()
    monkeypatch.setenv("SCN_CONV_SKIP", "0")
    gx_b, gw_b, gb_b = grads()
    monkeypatch.delenv("SCN_CONV_SKIP")
    assert torch.equal(gx_a, gx_b)
    assert rel_err(gw_a, gw_b) < 1e-5 and rel_err(gb_a, gb_b) < 1e-5


@pytest.mark.parametrize("channels", [32, 48, 128])
def test_tail_split_equals_persistent_grid(cuda, monkeypatch, channels):
    """Levels whose last wave of tiles is less than half full: the tail tiles are split into offset groups handed to the
    otherwise idle CTAs (conv_tc.cu, SCN_CONV_TAILSPLIT) and reduced with atomics + a last-arriver epilogue.  Same sums as
    whole tiles up to fp32 summation order; the accumulation buffer is left clean (second and third call)."""
    from sparse_rcnn_b200 import scn
    from sparse_rcnn_b200.synthetic import make_batch
    scn.set_precision("tf32")
    torch.manual_seed(2)
    coords, feats, size, bs, _ = make_batch(1, 5, spatial_size=(192, 192, 96), room=(128, 128, 64), room_offset=(32, 32, 8),
                                            n_furniture=12)
    md = scn.Metadata(3)
    f = scn.ioLayers.InputLayerFunction.apply(3, md, size, coords, feats.to(cuda), bs, 4)
    n = f.shape[0]
    tiles = (n + 127) // 128
    assert 0 < tiles % (148 * 3) <= 148 * 3 // 2 and 0 < tiles % (148 * 2) <= 148 and tiles > 148 * 3
    conv = scn.SubmanifoldConvolution(3, channels, channels, 3, True).to(cuda)
    x = scn.SparseConvNetTensor(torch.randn(n, channels, device=cuda), md, size)
    res = x.features.clone()
    with torch.no_grad():
        monkeypatch.setenv("SCN_CONV_TAILSPLIT", "0")
        ref = conv(x).features.clone()
        monkeypatch.setenv("SCN_CONV_TAILSPLIT", "1")
        for _ in range(3):
            a = conv(x).features
            assert rel_err(a, ref) < 1e-5, rel_err(a, ref)
            assert (a[-128 * 40:] - ref[-128 * 40:]).abs().max() < 1e-4 * ref.abs().max()      # the tail tiles themselves
    # backward through the same kernel (input gradient) with the switch on
    xin = res.requires_grad_(True)
    go = torch.randn(n, channels, device=cuda)
    conv(scn.SparseConvNetTensor(xin, md, size)).features.backward(go)
    g_on = xin.grad.clone()
    monkeypatch.setenv("SCN_CONV_TAILSPLIT", "0")
    xin.grad = None
    conv(scn.SparseConvNetTensor(xin, md, size)).features.backward(go)
    assert rel_err(g_on, xin.grad) < 1e-5


@pytest.mark.parametrize("cin,cout,epi2", [(32, 32, 9), (22, 22, 9), (7, 18, 8), (48, 80, 1), (6, 32, 9)])
@pytest.mark.parametrize("nocluster", ["0", "1"])
def test_second_output_of_the_epilogue(cuda, monkeypatch, cin, cout, epi2, nocluster):
    """scn_conv_fwd_tf32_dual: out2 == epi2(out) (ReLU = 1, ROUND = 8) bit for bit, `out` unchanged by the second store; odd
    widths take the scalar stores, the small scene the cluster reduction or (SCN_CONV_NOCLUSTER=1) the last-arriver pass."""
    from sparse_rcnn_b200 import _lib, scn
    from sparse_rcnn_b200.scn import functions as F
    scn.set_precision("tf32")
    monkeypatch.setenv("SCN_CONV_NOCLUSTER", nocluster)
    torch.manual_seed(cin + cout)
    coords, feats, size = random_scene(cin * 31 + cout, channels=cin)
    md = scn.Metadata(3)
    f = scn.ioLayers.InputLayerFunction.apply(3, md, torch.as_tensor(size), coords, feats.to(cuda), 0, 4)
    lvl = md.level(torch.as_tensor(size))
    n = f.shape[0]
    fmap = lvl.subm_map(3)
    w = torch.randn(27, cin, cout, device=cuda)
    bias = torch.randn(cout, device=cuda)
    res = torch.randn(n, cout, device=cuda)
    x = F.tf32_exact(torch.randn(n, cin, device=cuda))
    img = F._image(w, 27, cin, cout, 0, 0)
    ptr, s = F._ptr, F._stream()
    out = torch.empty(n, cout, device=cuda)
    _lib.call("scn_conv_fwd_tf32", ptr(x), cin, cin, n, ptr(fmap), n, 27, ptr(img), ptr(bias), ptr(res), cout, 0, 0, ptr(out), cout,
              cout, 2, s)
    o1, o2 = torch.full((n, cout), 7.0, device=cuda), torch.full((n, cout + 3), 7.0, device=cuda)
    _lib.call("scn_conv_fwd_tf32_dual", ptr(x), cin, cin, n, ptr(fmap), n, 27, ptr(img), ptr(bias), ptr(res), cout, 0, 0, ptr(o1), cout,
              cout, 2, ptr(o2), cout + 3, epi2, s)
    if nocluster == "0":
        assert torch.equal(o1, out)
    else:
        assert rel_err(o1, out) < 1e-5            # fp32 atomics
    want = o1.clamp_min(0) if epi2 & 1 else o1
    if epi2 & 8:
        want = F.tf32_exact(want.clone())
    assert torch.equal(o2[:, :cout], want)
    assert bool((o2[:, cout:] == 7.0).all())      # the padding columns of the wider second buffer are untouched
