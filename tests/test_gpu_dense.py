"""Dense stage behind the sparse backbone (sparse_rcnn_b200/dense.py + csrc/dense.cu): the reference's
get_dilation_network (module_factory.py:581-611) = SparseToDense + n x [Conv3d 3^3 padding 1 + ReLU], run on the gather-GEMM
kernels over the dense neighbour map.  The oracle for a DENSE convolution is torch's own conv3d in fp32 (CPU, double where
cheap): <= 1e-5 in fp32 mode, <= 2e-3 in TF32 mode."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as TF

from tests.util import rel_err

pytestmark = pytest.mark.gpu


def test_dense_map(cuda):
    from sparse_rcnn_b200.dense import DenseGrid
    for dil in (1, 2):
        g = DenseGrid(2, (5, 4, 3), cuda)
        m = g.map(dil).cpu().numpy()
        B, X, Y, Z = 2, 5, 4, 3
        want = np.full((27, B * X * Y * Z), -1, np.int32)
        idx = np.arange(B * X * Y * Z).reshape(B, X, Y, Z)
        o = 0
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    for b in range(B):
                        for x in range(X):
                            for y in range(Y):
                                for z in range(Z):
                                    qx, qy, qz = x + dx * dil, y + dy * dil, z + dz * dil
                                    if 0 <= qx < X and 0 <= qy < Y and 0 <= qz < Z:
                                        want[o, idx[b, x, y, z]] = idx[b, qx, qy, qz]
                    o += 1
        assert np.array_equal(m, want)
        assert np.array_equal(m[13], np.arange(B * X * Y * Z))


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("tf32", 2e-3)])
@pytest.mark.parametrize("cin,cout,dil", [(16, 32, 1), (64, 128, 1), (24, 24, 2)])
def test_dense_convolution_matches_conv3d(cuda, precision, tol, cin, cout, dil):
    from sparse_rcnn_b200 import scn
    from sparse_rcnn_b200.dense import DenseConvolution, DenseGrid
    scn.set_precision(precision)
    torch.manual_seed(cin + cout)
    B, X, Y, Z = 2, 12, 10, 8
    conv = DenseConvolution(cin, cout, dil, True, relu=True).to(cuda)
    conv.bias.data.normal_(0, 0.1)
    ref = nn.Conv3d(cin, cout, 3, padding=dil, dilation=dil)
    ref.load_state_dict({k: v.detach().cpu() for k, v in conv.state_dict().items()})
    x = torch.randn(B, cin, X, Y, Z)
    go = torch.randn(B, cout, X, Y, Z)
    xr = x.clone().double().requires_grad_(True)
    ref = ref.double()
    yr = TF.relu(ref(xr))
    yr.backward(go.double())
    grid = DenseGrid(B, (X, Y, Z), cuda)
    rows = x.permute(0, 2, 3, 4, 1).reshape(-1, cin).contiguous().to(cuda).requires_grad_(True)
    y = conv(rows, grid)
    y.backward(go.permute(0, 2, 3, 4, 1).reshape(-1, cout).contiguous().to(cuda))
    back = lambda t, c: t.detach().cpu().view(B, X, Y, Z, c).permute(0, 4, 1, 2, 3)
    assert rel_err(back(y, cout), yr) <= tol, rel_err(back(y, cout), yr)
    # a ReLU input that straddles zero within the tolerance flips its mask in TF32 mode: the gradients are held to the
    # TF32 bar in the L2 sense there, to 1e-5 in fp32 mode
    l2 = lambda a, b: float((a.double().cpu() - b.double()).norm() / b.double().norm())
    gtol = 1e-5 if precision == "fp32" else 2e-2
    assert l2(back(rows.grad, cin), xr.grad) <= gtol, l2(back(rows.grad, cin), xr.grad)
    assert l2(conv.weight.grad, ref.weight.grad) <= gtol, l2(conv.weight.grad, ref.weight.grad)
    assert l2(conv.bias.grad, ref.bias.grad) <= gtol
    scn.set_precision("tf32")


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("tf32", 2e-3)])
def test_dilation_network_matches_the_reference_graph(cuda, precision, tol):
    """SparseToDense + 3 x [Conv3d + ReLU] on a sparse level-2-like map: against scn.SparseToDense followed by torch's conv3d
    (the graph module_factory.get_dilation_network builds), same state_dict."""
    from sparse_rcnn_b200 import scn
    from sparse_rcnn_b200.dense import DilationNetwork
    from tests.util import make_pair, random_scene
    scn.set_precision(precision)
    torch.manual_seed(3)
    coords, feats, size = random_scene(5, channels=16, size=(16, 12, 8))
    _, t = make_pair(scn, coords, feats, size, cuda)
    net = DilationNetwork(scn, 16, 32, 3).to(cuda)
    assert list(net.state_dict()) == ["1.weight", "1.bias", "3.weight", "3.bias", "5.weight", "5.bias"]
    ref = nn.Sequential(nn.Identity(), nn.Conv3d(16, 32, 3, padding=1), nn.ReLU(), nn.Conv3d(32, 32, 3, padding=1), nn.ReLU(),
                        nn.Conv3d(32, 32, 3, padding=1), nn.ReLU())
    ref.load_state_dict({k: v.detach().cpu() for k, v in net.state_dict().items()})
    x = t.features.detach().clone().requires_grad_(True)
    out = net(scn.SparseConvNetTensor(x, t.metadata, size))
    dense = scn.SparseToDense(3, 16)(scn.SparseConvNetTensor(t.features.detach(), t.metadata, size)).cpu()
    want = ref.double()(dense.double())
    assert out.shape == want.shape == (2, 32, 16, 12, 8)
    assert rel_err(out, want) <= tol, rel_err(out, want)
    out.sum().backward()
    assert x.grad is not None and x.grad.shape == x.shape and float(x.grad.abs().sum()) > 0
    scn.set_precision("tf32")
