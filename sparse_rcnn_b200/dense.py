"""Dense stage behind the sparse backbone: host mirror of the reference's `get_dilation_network`
(ndsis/modules/module_factory.py:581-611) as configured for the region-proposal trunk -- `scn.SparseToDense`, then
`num_dilations` x [`nn.Conv3d(cin, cout, 3, padding=1)` + `nn.ReLU`] -- with the same module indices, parameter names and
shapes (`<2k+1>.weight [Cout, Cin, 3, 3, 3]`, `<2k+1>.bias [Cout]`), so the reference's state_dict loads.

The dense grid is kept channels-last (one row of C floats per cell, cells in (b, x, y, z) order): a dense 'same' convolution is
then the submanifold convolution over the trivial neighbour map `scn_dense_map` builds, and forward / input gradient / weight
gradient run on the same tcgen05 gather-GEMM kernels as the sparse layers, bias + ReLU (+ TF32 rounding) in the epilogue
(csrc/dense.cu).  The result is handed on as the [B, C, X, Y, Z] tensor the reference's anchor heads take
(anchor_network.py:127-219).  No CPU path.
"""
import torch
import torch.nn as nn

from . import _lib
from .scn import functions as F
from .scn.metadata import _ptr, _stream, size_key


class DenseGrid:
    """A dense [B, X, Y, Z] grid and its cached 3^3 neighbour maps (one per dilation)."""

    def __init__(self, batch, size, device):
        self.B = int(batch)
        self.X, self.Y, self.Z = (int(v) for v in size)
        self.device = device
        self.n = self.B * self.X * self.Y * self.Z
        self._maps = {}

    def map(self, dilation=1):
        m = self._maps.get(dilation)
        if m is None:
            m = torch.empty((27, self.n), dtype=torch.int32, device=self.device)
            _lib.call("scn_dense_map", self.B, self.X, self.Y, self.Z, int(dilation), _ptr(m), _stream())
            self._maps[dilation] = m
        return m


class SparseToDenseRowsFunction(F.Function):
    """scn.SparseToDense (module_factory.py:429-435) into channels-last rows; zero fill fused."""

    @staticmethod
    def forward(ctx, x, level, grid):
        x = F._check(x)
        C = x.shape[1]
        out = torch.empty((grid.n, C), dtype=torch.float32, device=x.device)
        _lib.call("scn_sparse_to_dense_rows_fwd", _ptr(x), C, _ptr(level.tab_keys), _ptr(level.tab_vals), level.cap, grid.B, grid.X,
                  grid.Y, grid.Z, _ptr(out), _stream())
        ctx.level, ctx.grid, ctx.n, ctx.C = level, grid, x.shape[0], C
        return out

    @staticmethod
    def backward(ctx, go):
        go = F._check(go)
        gi = torch.empty((ctx.n, ctx.C), dtype=torch.float32, device=go.device)
        g = ctx.grid
        _lib.call("scn_sparse_to_dense_rows_bwd", _ptr(go), _ptr(ctx.level.keys), ctx.n, ctx.C, g.X, g.Y, g.Z, _ptr(gi), _stream())
        return gi, None, None


class TransposeFunction(F.Function):
    """[B, rows, cols] -> [B, cols, rows] (dense rows <-> channels-first); its own inverse in the backward."""

    @staticmethod
    def forward(ctx, x, batches, rows, cols):
        x = F._check(x)
        out = torch.empty((batches, cols, rows), dtype=torch.float32, device=x.device)
        _lib.call("scn_transpose_batched", _ptr(x), batches, rows, cols, _ptr(out), _stream())
        ctx.dims = (batches, rows, cols)
        return out

    @staticmethod
    def backward(ctx, go):
        b, r, c = ctx.dims
        go = F._check(go)
        gi = torch.empty((b, r, c), dtype=torch.float32, device=go.device)
        _lib.call("scn_transpose_batched", _ptr(go), b, c, r, _ptr(gi), _stream())
        return gi, None, None, None


def _w27(weight):
    """nn.Conv3d weight [Cout, Cin, 3, 3, 3] -> gather-GEMM layout [27, Cin, Cout] (offsets last-dimension-fastest)."""
    cout, cin = weight.shape[:2]
    return weight.detach().permute(2, 3, 4, 1, 0).reshape(27, cin, cout).contiguous()


class DenseConvReLUFunction(F.Function):
    """relu(conv3d(x, w, b, padding=dilation, dilation=dilation)) on dense rows: ONE gather-GEMM launch forward (bias, ReLU and
    TF32 rounding in the epilogue), backward = ReLU mask pass + transposed gather-GEMM + weight / bias gradient."""

    @staticmethod
    def forward(ctx, x, weight, bias, grid, dilation, relu):
        x = F._check(x)
        cout, cin = weight.shape[:2]
        if x.shape != (grid.n, cin):
            raise RuntimeError("dense convolution expects [%d, %d] rows, got %s" % (grid.n, cin, tuple(x.shape)))
        w27 = _w27(weight)
        m = grid.map(dilation)
        out = F.conv_gemm(x, w27, 27, cin, cout, m, grid.n, bias=bias, relu=relu, round_out=relu)
        ctx.save_for_backward(x, w27, out)
        ctx.cfg = (grid, dilation, relu, cin, cout, bias is not None)
        return out

    @staticmethod
    def backward(ctx, go):
        x, w27, out = ctx.saved_tensors
        grid, dilation, relu, cin, cout, has_bias = ctx.cfg
        m = grid.map(dilation)
        go = F._check(go)
        tf32 = F.get_precision() == "tf32"
        if relu:
            g = torch.empty_like(go)
            _lib.call("scn_relu_bwd", _ptr(out), _ptr(go), _ptr(g), go.numel(), int(tf32), _stream())
            F._mark(g)
        else:
            g = F.tf32_exact(go) if tf32 else go
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = F.conv_gemm(g, w27, 27, cout, cin, m, grid.n, transpose=1, reverse=1)
        if ctx.needs_input_grad[1]:
            gw27 = F._wgrad(F.tf32_exact(x) if tf32 else x, m, g, 27, cin, cout, grid.n, w27)
            gw = gw27.reshape(3, 3, 3, cin, cout).permute(4, 3, 0, 1, 2).contiguous()
        if has_bias and ctx.needs_input_grad[2]:
            gb = F._bgrad(g, grid.n, cout)
        return gx, gw, gb, None, None, None


class DenseConvolution(nn.Module):
    """nn.Conv3d(cin, cout, 3, padding=dilation, dilation=dilation) on dense rows (parameters as nn.Conv3d's)."""

    def __init__(self, cin, cout, dilation=1, bias=True, relu=False):
        super().__init__()
        ref = nn.Conv3d(cin, cout, 3, padding=dilation, dilation=dilation, bias=bias)      # the reference's initialisation
        self.weight = ref.weight
        self.bias = ref.bias
        self.nIn, self.nOut, self.dilation, self.relu = cin, cout, dilation, relu

    def forward(self, rows, grid):
        return DenseConvReLUFunction.run(rows, self.weight, self.bias, grid, self.dilation, self.relu)


class DilationNetwork(nn.Sequential):
    """get_dilation_network(num_dims=3, sparse=True, cin, cout, num_dilations=n, make_dense=True, kernel_size=3):
    children 0 = SparseToDense, 2k+1 = convolution, 2k+2 = ReLU (fused into the convolution's epilogue here).
    forward(SparseConvNetTensor) -> [B, cout, X, Y, Z]."""

    def __init__(self, scn, cin, cout, num_dilations, dilation=1):
        layers = [scn.SparseToDense(3, cin)]
        c = cin
        for _ in range(num_dilations):
            layers += [DenseConvolution(c, cout, dilation, True, relu=True), nn.ReLU(inplace=True)]
            c = cout
        super().__init__(*layers)
        self.out_channels = c

    def forward(self, x):
        md = x.metadata
        size = size_key(x.spatial_size)
        grid = DenseGrid(md.n_samples, size, x.features.device)
        rows = SparseToDenseRowsFunction.run(x.features, md.level(x.spatial_size), grid)
        for m in self:
            if isinstance(m, DenseConvolution):
                rows = m(rows, grid)
        vol = grid.X * grid.Y * grid.Z
        out = TransposeFunction.run(rows.view(grid.B, vol, self.out_channels), grid.B, vol, self.out_channels)
        return out.view(grid.B, self.out_channels, grid.X, grid.Y, grid.Z)
