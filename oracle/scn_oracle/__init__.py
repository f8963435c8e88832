"""ORACLE (test infrastructure; only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference leg may import this).

`scn_oracle` is a pure CPU (numpy + torch-CPU) restatement of the part of the
`sparseconvnet` Python namespace that LeonhardFeiner/sparse_rcnn consumes
(SURVEY.md 8b lists the names; call sites: ndsis/modules/module_factory.py:5,
ndsis/modules/model.py:6, ndsis/modules/custom_operations.py:4,
ndsis/modules/roi_select_sparse.py:3).  It follows SparseConvNet's CPU backend:
hash-style rulebooks on the host (rules.py), then per kernel offset
`index_select -> matmul -> index_add_` in fp32 (upstream SCN/CPU/Convolution.cpp,
SURVEY.md A5).  It can be aliased as `sparseconvnet` so the UNMODIFIED
`ndsis.modules` run on it (oracle/make_golden.py does exactly that).

PARITY UNPINNED against SparseConvNet itself (not vendored, not installable
here); pinned against dense torch conv/pool equivalences and the reference's
own importable crop code.  See rules.py header.
"""
import math
import types

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from . import rules as R


# --------------------------------------------------------------------------- metadata
class Metadata:
    """Upstream `Metadata<d>`: per-scale grids + cached rulebooks."""

    def __init__(self, dimension=3):
        assert dimension == 3, "ndsis only uses 3-D (model.py:264, run.py:611)"
        self.dimension = dimension
        self.grids = {}          # tuple(spatial_size) -> R.Grid
        self.subm = {}           # (size, filter) -> rules
        self.conv = {}           # (in_size, filter, stride) -> (out_size, rules, parent, off)
        self.point_row = None
        self.n_samples = 0
        self.input_size = None

    @staticmethod
    def _key(size):
        return tuple(int(s) for s in np.asarray(size).reshape(-1))

    def set_input(self, spatial_size, coords, batch_size, mode):
        grid, point_row, n_samples = R.input_layer_rules(
            np.asarray(coords), batch_size, mode)
        self.input_size = self._key(spatial_size)
        self.grids[self.input_size] = grid
        self.point_row = point_row
        self.n_samples = n_samples
        self.mode = mode
        return grid.n

    def grid(self, spatial_size):
        return self.grids[self._key(spatial_size)]

    def subm_rules(self, spatial_size, filter_size):
        fk = self._key(np.broadcast_to(np.asarray(filter_size), (3,)))
        k = (self._key(spatial_size), fk)
        if k not in self.subm:
            self.subm[k] = R.submanifold_rules(self.grid(spatial_size), filter_size)
        return self.subm[k]

    def conv_rules(self, in_size, filter_size, stride):
        fk = self._key(np.broadcast_to(np.asarray(filter_size), (3,)))
        sk = self._key(np.broadcast_to(np.asarray(stride), (3,)))
        k = (self._key(in_size), fk, sk)
        if k not in self.conv:
            g_out, out_size, rules, parent, off = R.strided_rules(
                self.grid(in_size), np.asarray(in_size), filter_size, stride)
            ok = self._key(out_size)
            if ok in self.grids:
                # upstream inserts into the existing grid; same active set => same rows here
                assert self.grids[ok].n == g_out.n
            else:
                self.grids[ok] = g_out
            self.conv[k] = (ok, rules, parent, off)
        return self.conv[k]


class SparseConvNetTensor:
    """Upstream sparseConvNetTensor.py container."""

    def __init__(self, features=None, metadata=None, spatial_size=None):
        self.features = features
        self.metadata = metadata
        self.spatial_size = spatial_size

    def get_spatial_locations(self, spatial_size=None):
        if spatial_size is None:
            spatial_size = self.spatial_size
        return torch.from_numpy(self.metadata.grid(spatial_size).coords.copy())

    def batch_size(self):
        return self.metadata.n_samples

    def cpu(self):
        self.features = self.features.cpu()
        return self

    def __repr__(self):
        return "SparseConvNetTensor<%s, size=%s>" % (
            tuple(self.features.shape), tuple(np.asarray(self.spatial_size).tolist()))


def _idx(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64))


# --------------------------------------------------------------------------- functions
class InputLayerFunction(Function):
    """Upstream ioLayers.InputLayerFunction.  Modes: 0 copy
    (unique coords), 1 last, 2 first, 3 sum, 4 mean."""

    @staticmethod
    def forward(ctx, dimension, metadata, spatial_size, coords, input_features,
                batch_size, mode):
        n = metadata.set_input(spatial_size, coords.cpu().numpy(), batch_size, mode)
        pr = _idx(metadata.point_row)
        ctx.mode, ctx.n = mode, n
        f = input_features
        if mode == 0:
            ctx.save_for_backward(pr, None)
            return f.clone()
        counts = torch.bincount(pr, minlength=n).to(f.dtype)
        ctx.save_for_backward(pr, counts)
        if mode in (3, 4):
            out = f.new_zeros((n, f.shape[1])).index_add_(0, pr, f)
            if mode == 4:
                out = out / counts[:, None]
            return out
        # mode 1 (last) / 2 (first)
        order = torch.arange(len(pr))
        sel = torch.full((n,), -1 if mode == 1 else len(pr), dtype=torch.long)
        sel = sel.scatter_reduce(0, pr, order, "amax" if mode == 1 else "amin")
        ctx.sel = sel
        return f[sel]

    @staticmethod
    def backward(ctx, grad_out):
        pr, counts = ctx.saved_tensors
        if ctx.mode == 0:
            g = grad_out.clone()
        elif ctx.mode == 3:
            g = grad_out[pr]
        elif ctx.mode == 4:
            g = (grad_out / counts[:, None])[pr]
        else:
            g = grad_out.new_zeros((len(pr), grad_out.shape[1]))
            g[ctx.sel] = grad_out
        return None, None, None, None, g, None, None


class OutputLayerFunction(Function):
    """Upstream ioLayers.OutputLayerFunction: every original point receives
    its voxel's row (inverse of the input rule; backward accumulates per row).
    Call sites: custom_operations.py:7-10, model.py:461,576,600,643,658."""

    @staticmethod
    def forward(ctx, dimension, metadata, input_features):
        pr = _idx(metadata.point_row)
        ctx.n = input_features.shape[0]
        ctx.save_for_backward(pr)
        return input_features[pr]

    @staticmethod
    def backward(ctx, grad_out):
        pr, = ctx.saved_tensors
        g = grad_out.new_zeros((ctx.n, grad_out.shape[1])).index_add_(0, pr, grad_out)
        return None, None, g


def _rule_conv(x, w, bias, rules, n_out, swap=False):
    """SCN/CPU/Convolution.cpp: out = bias (or 0); for each offset
    out[rule.out] += in[rule.in] @ W[o]  (fp32, offsets in ascending order)."""
    cout = w.shape[-1]
    out = x.new_zeros((n_out, cout)) if bias is None else bias.expand(n_out, cout).clone()
    for o, (ri, ro) in enumerate(rules):
        if swap:
            ri, ro = ro, ri
        if len(ri):
            out.index_add_(0, _idx(ro), x[_idx(ri)] @ w[o])
    return out


class _RuleConvFunction(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, rules, n_out, swap):
        w = weight.reshape(weight.shape[0], weight.shape[-2], weight.shape[-1])
        ctx.rules, ctx.swap, ctx.has_bias = rules, swap, bias is not None
        ctx.save_for_backward(x, weight)
        return _rule_conv(x, w, bias, rules, n_out, swap)

    @staticmethod
    def backward(ctx, go):
        x, weight = ctx.saved_tensors
        w = weight.reshape(weight.shape[0], weight.shape[-2], weight.shape[-1])
        gx = torch.zeros_like(x)
        gw = torch.zeros_like(w)
        for o, (ri, ro) in enumerate(ctx.rules):
            if ctx.swap:
                ri, ro = ro, ri
            if len(ri):
                ri, ro = _idx(ri), _idx(ro)
                g = go[ro]
                gx.index_add_(0, ri, g @ w[o].t())
                gw[o] = x[ri].t() @ g
        gb = go.sum(0) if ctx.has_bias else None
        return gx, gw.reshape(weight.shape), gb, None, None, None


class _MaxPoolFunction(Function):
    """SCN MaxPooling: max over ACTIVE children; backward routes the gradient
    to every input equal to the output."""

    @staticmethod
    def forward(ctx, x, rules, n_out):
        out = x.new_full((n_out, x.shape[1]), -math.inf)
        for ri, ro in rules:
            if len(ri):
                ro_t = _idx(ro)
                out[ro_t] = torch.maximum(out[ro_t], x[_idx(ri)])
        ctx.rules = rules
        ctx.save_for_backward(x, out)
        return out

    @staticmethod
    def backward(ctx, go):
        x, out = ctx.saved_tensors
        gx = torch.zeros_like(x)
        for ri, ro in ctx.rules:
            if len(ri):
                ri_t, ro_t = _idx(ri), _idx(ro)
                gx[ri_t] = torch.where(x[ri_t] == out[ro_t], go[ro_t], gx[ri_t])
        return gx, None, None


# --------------------------------------------------------------------------- modules
class Sequential(nn.Sequential):
    def append(self, module):
        self.add_module(str(len(self._modules)), module)
        return self

    def input_spatial_size(self, out_size):
        for m in reversed(list(self._modules.values())):
            out_size = m.input_spatial_size(out_size)
        return out_size


class Identity(nn.Module):
    def forward(self, x):
        return x

    def input_spatial_size(self, out_size):
        return out_size


class ConcatTable(nn.Module):
    def __init__(self, *modules):
        super().__init__()
        for i, m in enumerate(modules):
            self.add_module(str(i), m)

    def append(self, module):
        self.add_module(str(len(self._modules)), module)
        return self

    def forward(self, x):
        return [m(x) for m in self._modules.values()]


def _like(x, features):
    return SparseConvNetTensor(features, x.metadata, x.spatial_size)


class AddTable(nn.Module):
    def forward(self, xs):
        return _like(xs[0], sum(x.features for x in xs))


class JoinTable(nn.Module):
    def forward(self, xs):
        return _like(xs[0], torch.cat([x.features for x in xs], 1))


class ReLU(nn.Module):
    def forward(self, x):
        return _like(x, torch.relu(x.features))


class BatchNormalization(nn.Module):
    """Upstream batchNormalization.py: per-channel statistics over
    the active rows, biased variance, eps inside the sqrt, running stats
    r <- momentum*r + (1-momentum)*batch (momentum is the KEEP factor), fused
    leaky ReLU."""

    def __init__(self, nPlanes, eps=1e-4, momentum=0.9, affine=True, leakiness=1):
        super().__init__()
        self.nPlanes, self.eps, self.momentum, self.leakiness = nPlanes, eps, momentum, leakiness
        self.register_buffer("running_mean", torch.zeros(nPlanes))
        self.register_buffer("running_var", torch.ones(nPlanes))
        if affine:
            self.weight = nn.Parameter(torch.ones(nPlanes))
            self.bias = nn.Parameter(torch.zeros(nPlanes))
        else:
            self.weight = self.bias = None

    def forward(self, x):
        f = x.features
        if self.training:
            mean = f.mean(0)
            var = f.var(0, unbiased=False)
            with torch.no_grad():
                self.running_mean.mul_(self.momentum).add_(mean, alpha=1 - self.momentum)
                self.running_var.mul_(self.momentum).add_(var, alpha=1 - self.momentum)
        else:
            mean, var = self.running_mean, self.running_var
        y = (f - mean) / torch.sqrt(var + self.eps)
        if self.weight is not None:
            y = y * self.weight + self.bias
        y = torch.where(y > 0, y, y * self.leakiness)
        return _like(x, y)


class BatchNormReLU(BatchNormalization):
    def __init__(self, nPlanes, eps=1e-4, momentum=0.9):
        super().__init__(nPlanes, eps, momentum, True, 0)


class BatchNormLeakyReLU(BatchNormalization):
    def __init__(self, nPlanes, eps=1e-4, momentum=0.9, leakiness=0.333):
        super().__init__(nPlanes, eps, momentum, True, leakiness)


def _triple(v):
    return tuple(int(a) for a in np.broadcast_to(np.asarray(v), (3,)))


def _conv_weight(volume, nin, nout, groups):
    assert groups == 1, "ndsis always builds groups=1 (module_factory.py:152,404-406)"
    std = math.sqrt(2.0 * groups / (nin * volume))       # SURVEY.md A5 initialiser
    return nn.Parameter(torch.empty(volume, groups, nin // groups, nout // groups).normal_(0, std))


class SubmanifoldConvolution(nn.Module):
    def __init__(self, dimension, nIn, nOut, filter_size, bias, groups=1):
        super().__init__()
        self.dimension, self.nIn, self.nOut = dimension, nIn, nOut
        self.filter_size = _triple(filter_size)
        self.weight = _conv_weight(int(np.prod(self.filter_size)), nIn, nOut, groups)
        self.bias = nn.Parameter(torch.zeros(nOut)) if bias else None

    def forward(self, x):
        assert x.features.shape[1] == self.nIn
        rules = x.metadata.subm_rules(x.spatial_size, self.filter_size)
        f = _RuleConvFunction.apply(x.features, self.weight, self.bias, rules,
                                    x.features.shape[0], False)
        return _like(x, f)

    def input_spatial_size(self, out_size):
        return out_size


class ValidConvolution(SubmanifoldConvolution):
    pass


class Convolution(nn.Module):
    def __init__(self, dimension, nIn, nOut, filter_size, filter_stride, bias, groups=1):
        super().__init__()
        self.dimension, self.nIn, self.nOut = dimension, nIn, nOut
        self.filter_size, self.filter_stride = _triple(filter_size), _triple(filter_stride)
        self.weight = _conv_weight(int(np.prod(self.filter_size)), nIn, nOut, groups)
        self.bias = nn.Parameter(torch.zeros(nOut)) if bias else None

    def forward(self, x):
        assert x.features.shape[1] == self.nIn
        out_key, rules, _, _ = x.metadata.conv_rules(
            x.spatial_size, self.filter_size, self.filter_stride)
        n_out = x.metadata.grids[out_key].n
        f = _RuleConvFunction.apply(x.features, self.weight, self.bias, rules, n_out, False)
        return SparseConvNetTensor(f, x.metadata, torch.tensor(out_key, dtype=torch.long))

    def input_spatial_size(self, out_size):
        return (out_size - 1) * torch.tensor(self.filter_stride) + torch.tensor(self.filter_size)


class Deconvolution(nn.Module):
    """Transpose of Convolution: reuses the matching conv rulebook with the
    columns swapped; its active output set is the pre-existing finer grid
    (SURVEY.md A4)."""

    def __init__(self, dimension, nIn, nOut, filter_size, filter_stride, bias, groups=1):
        super().__init__()
        self.dimension, self.nIn, self.nOut = dimension, nIn, nOut
        self.filter_size, self.filter_stride = _triple(filter_size), _triple(filter_stride)
        self.weight = _conv_weight(int(np.prod(self.filter_size)), nIn, nOut, groups)
        self.bias = nn.Parameter(torch.zeros(nOut)) if bias else None

    def forward(self, x):
        assert x.features.shape[1] == self.nIn
        in_size = torch.as_tensor(x.spatial_size)
        out_size = (in_size - 1) * torch.tensor(self.filter_stride) + torch.tensor(self.filter_size)
        md = x.metadata
        if md._key(out_size) not in md.grids:
            raise RuntimeError("Deconvolution needs the finer grid to exist in the Metadata")
        _, rules, _, _ = md.conv_rules(out_size, self.filter_size, self.filter_stride)
        n_out = md.grid(out_size).n
        f = _RuleConvFunction.apply(x.features, self.weight, self.bias, rules, n_out, True)
        return SparseConvNetTensor(f, md, out_size)


class NetworkInNetwork(nn.Module):
    def __init__(self, nIn, nOut, bias):
        super().__init__()
        self.nIn, self.nOut = nIn, nOut
        self.weight = nn.Parameter(torch.empty(nIn, nOut).normal_(0, math.sqrt(2.0 / nIn)))
        self.bias = nn.Parameter(torch.zeros(nOut)) if bias else None

    def forward(self, x):
        f = x.features @ self.weight
        if self.bias is not None:
            f = f + self.bias
        return _like(x, f)

    def input_spatial_size(self, out_size):
        return out_size


class MaxPooling(nn.Module):
    def __init__(self, dimension, pool_size, pool_stride, nFeaturesToDrop=0):
        super().__init__()
        self.pool_size, self.pool_stride = _triple(pool_size), _triple(pool_stride)

    def forward(self, x):
        out_key, rules, _, _ = x.metadata.conv_rules(x.spatial_size, self.pool_size, self.pool_stride)
        n_out = x.metadata.grids[out_key].n
        f = _MaxPoolFunction.apply(x.features, rules, n_out)
        return SparseConvNetTensor(f, x.metadata, torch.tensor(out_key, dtype=torch.long))


class AveragePooling(nn.Module):
    """Sum over active children divided by the pool VOLUME (not the count)."""

    def __init__(self, dimension, pool_size, pool_stride, nFeaturesToDrop=0):
        super().__init__()
        self.pool_size, self.pool_stride = _triple(pool_size), _triple(pool_stride)

    def forward(self, x):
        out_key, rules, parent, _ = x.metadata.conv_rules(x.spatial_size, self.pool_size, self.pool_stride)
        n_out = x.metadata.grids[out_key].n
        f = x.features.new_zeros((n_out, x.features.shape[1])).index_add_(
            0, _idx(parent), x.features) / float(np.prod(self.pool_size))
        return SparseConvNetTensor(f, x.metadata, torch.tensor(out_key, dtype=torch.long))


class UnPooling(nn.Module):
    """Inverse site map of the poolings: every active fine site gets its coarse parent's row (SparseConvNet UnPooling)."""

    def __init__(self, dimension, pool_size, pool_stride, nFeaturesToDrop=0):
        super().__init__()
        self.pool_size, self.pool_stride = _triple(pool_size), _triple(pool_stride)

    def forward(self, x):
        md = x.metadata
        coarse = torch.as_tensor(x.spatial_size)
        fine = (coarse - 1) * torch.tensor(self.pool_stride) + torch.tensor(self.pool_size)
        if md._key(fine) not in md.grids:
            raise RuntimeError("UnPooling needs the finer grid to exist in the Metadata")
        out_key, _, parent, _ = md.conv_rules(fine, self.pool_size, self.pool_stride)
        assert tuple(out_key) == tuple(int(v) for v in coarse)
        return SparseConvNetTensor(x.features[_idx(parent)], md, fine)


class SparseToDense(nn.Module):
    def __init__(self, dimension, nPlanes):
        super().__init__()
        self.nPlanes = nPlanes

    def forward(self, x):
        g = x.metadata.grid(x.spatial_size)
        size = [int(s) for s in np.asarray(x.spatial_size).reshape(-1)]
        c = _idx(g.coords)
        dense = x.features.new_zeros((x.metadata.n_samples, *size, x.features.shape[1]))
        dense = dense.index_put((c[:, 3], c[:, 0], c[:, 1], c[:, 2]), x.features)
        return dense.permute(0, 4, 1, 2, 3).contiguous()


class OutputLayer(nn.Module):
    def __init__(self, dimension):
        super().__init__()
        self.dimension = dimension

    def forward(self, x):
        return OutputLayerFunction.apply(self.dimension, x.metadata, x.features)


class InputLayer(nn.Module):
    def __init__(self, dimension, spatial_size, mode=3):
        super().__init__()
        self.dimension, self.mode = dimension, mode
        self.spatial_size = torch.as_tensor(spatial_size, dtype=torch.long)

    def forward(self, inp):
        coords, feats = inp[0], inp[1]
        bs = inp[2] if len(inp) > 2 else 0
        md = Metadata(self.dimension)
        f = InputLayerFunction.apply(self.dimension, md, self.spatial_size, coords.long(), feats, bs, self.mode)
        return SparseConvNetTensor(f, md, self.spatial_size)


ioLayers = types.SimpleNamespace(
    InputLayerFunction=InputLayerFunction, OutputLayerFunction=OutputLayerFunction,
    InputLayer=InputLayer, OutputLayer=OutputLayer)

BACKEND = "oracle-cpu"
