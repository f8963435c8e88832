"""Module classes of the `sparseconvnet` namespace consumed by ndsis/modules
(names, positional constructor signatures and parameter names/shapes as upstream, so that the
reference's module_factory.py / model.py build unchanged and state_dicts stay loadable:
SURVEY.md 8b).  Weight `[K^3, groups, Cin/g, Cout/g]`, bias `[Cout]`; NiN weight `[Cin, Cout]`;
BN `weight, bias, running_mean, running_var`.
"""
import math

import torch
import torch.nn as nn

from . import functions as F
from .metadata import Metadata, _triple, size_key


class SparseConvNetTensor:
    """Container (features, metadata, spatial_size); reference call sites:
    custom_operations.py:19-21,83-85; roi_select_sparse.py:83-84,103-109."""

    def __init__(self, features=None, metadata=None, spatial_size=None):
        self.features = features
        self.metadata = metadata
        self.spatial_size = spatial_size

    def get_spatial_locations(self, spatial_size=None):
        """int64 CPU [N,4] (x,y,z,b), row-aligned with `features` (cached)."""
        if spatial_size is None:
            spatial_size = self.spatial_size
        return self.metadata.level(spatial_size).locations()

    def batch_size(self):
        return self.metadata.n_samples

    def cuda(self):
        self.features = self.features.cuda()
        return self

    def cpu(self):
        self.features = self.features.cpu()
        return self

    def __repr__(self):
        return "SparseConvNetTensor<features=%s, spatial_size=%s>" % (
            tuple(self.features.shape), size_key(self.spatial_size))


def _like(x, features):
    return SparseConvNetTensor(features, x.metadata, x.spatial_size)


FUSE = {"residual": True}      # False: run the plain module graph (identical numerics in fp32 mode; tests compare both)


def _match_residual_unit(m):
    """(conv1, conv2) if `m` is the residual unit that the reference's module_factory.py:51-57,127-183 builds for the
    shipped sparse configuration (relu_first, no batch norm, identity shortcut):
        Sequential(ConcatTable(Identity, Sequential(ReLU, SubM 3^3 C->C, ReLU, SubM 3^3 C->C)), AddTable)
    else None.  Exact types only: a subclass may override forward."""
    if type(m) is not Sequential or len(m) != 2:
        return None
    table, add = m[0], m[1]
    if type(table) is not ConcatTable or type(add) is not AddTable or len(table._modules) != 2:
        return None
    shortcut, inner = table._modules["0"], table._modules["1"]
    if type(shortcut) is not Identity or type(inner) is not Sequential or len(inner) != 4:
        return None
    r1, c1, r2, c2 = inner[0], inner[1], inner[2], inner[3]
    if type(r1) is not ReLU or type(r2) is not ReLU:
        return None
    for c in (c1, c2):
        if type(c) is not SubmanifoldConvolution or c.filter_size != (3, 3, 3) or c.dimension != 3:
            return None
    if not (c1.nIn == c1.nOut == c2.nIn == c2.nOut) or (c1.bias is None) != (c2.bias is None):
        return None
    return c1, c2


class Sequential(nn.Sequential):
    """Runs its children in order -- except that maximal runs of residual units over one level (the reference's
    `unit_stage`, module_factory.py:438-578) execute as ONE autograd node, `functions.ResidualUnitFunction`: per unit one
    elementwise pass + two gather-GEMMs whose epilogues carry ReLU, residual add and TF32 rounding.  The module tree, the
    parameters and the state_dict keys are untouched, so the UNMODIFIED `ndsis.modules` graph gets the fused path; a
    Sequential that is itself one residual unit is handled the same way."""

    def append(self, module):
        self.add_module(str(len(self._modules)), module)
        return self

    def _plan(self):
        mods = tuple(self._modules.values())
        cached = self.__dict__.get("_scn_plan")
        if cached is not None and cached[0] == tuple(id(m) for m in mods):
            return cached[1]
        plan = []
        me = _match_residual_unit(self)
        if me is not None:
            plan.append(("units", [me]))
        else:
            for m in mods:
                u = _match_residual_unit(m)
                if u is None:
                    plan.append(("module", m))
                elif plan and plan[-1][0] == "units" and plan[-1][1][-1][1].nOut == u[0].nIn:
                    plan[-1][1].append(u)
                else:
                    plan.append(("units", [u]))
        if not any(kind == "units" for kind, _ in plan):
            plan = None
        self.__dict__["_scn_plan"] = (tuple(id(m) for m in mods), plan)
        return plan

    def forward(self, x):
        plan = self._plan() if FUSE["residual"] else None
        if plan is None:
            for m in self._modules.values():
                x = m(x)
            return x
        for kind, item in plan:
            if kind == "module":
                x = item(x)
                continue
            if not isinstance(x, SparseConvNetTensor) or not x.features.is_cuda:
                raise RuntimeError("sparse_rcnn_b200.scn modules need CUDA SparseConvNetTensors (there is no CPU fallback)")
            params = []
            for c1, c2 in item:
                params += [c1.weight, c1.bias, c2.weight, c2.bias]
            lvl = x.metadata.level(x.spatial_size)
            fmap = lvl.subm_map(item[0][0].filter_size, tile_book=False)
            lvl.ensure_tile_book(item[0][0].nIn)
            f = F.ResidualUnitFunction.run(x.features, fmap, lvl.n, *params)
            x = SparseConvNetTensor(f, x.metadata, x.spatial_size)
        return x

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


class AddTable(nn.Module):
    def forward(self, xs):
        f = xs[0].features
        for x in xs[1:]:
            f = F.AddFunction.run(f, x.features)
        return _like(xs[0], f)


class JoinTable(nn.Module):
    def forward(self, xs):
        return _like(xs[0], torch.cat([x.features for x in xs], 1))


class ReLU(nn.Module):
    def forward(self, x):
        return _like(x, F._mark(F.ReLUFunction.run(x.features)))

    def input_spatial_size(self, out_size):
        return out_size


class BatchNormalization(nn.Module):
    def __init__(self, nPlanes, eps=1e-4, momentum=0.9, affine=True, leakiness=1):
        super().__init__()
        self.nPlanes, self.eps, self.momentum, self.leakiness = nPlanes, eps, momentum, leakiness
        self.register_buffer("running_mean", torch.zeros(nPlanes))
        self.register_buffer("running_var", torch.ones(nPlanes))
        if affine:
            self.weight = nn.Parameter(torch.ones(nPlanes))
            self.bias = nn.Parameter(torch.zeros(nPlanes))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)

    def forward(self, x):
        f = F.BatchNormFunction.run(x.features, self.weight, self.bias, self.running_mean, self.running_var,
                                      self.eps, self.momentum, float(self.leakiness), self.training)
        return _like(x, f)

    def input_spatial_size(self, out_size):
        return out_size


class BatchNormReLU(BatchNormalization):
    def __init__(self, nPlanes, eps=1e-4, momentum=0.9):
        super().__init__(nPlanes, eps, momentum, True, 0)


class BatchNormLeakyReLU(BatchNormalization):
    def __init__(self, nPlanes, eps=1e-4, momentum=0.9, leakiness=0.333):
        super().__init__(nPlanes, eps, momentum, True, leakiness)


def _conv_weight(volume, nin, nout, groups):
    if groups != 1:
        raise RuntimeError("groups != 1 is not implemented (ndsis builds groups=1: module_factory.py:152,404-406)")
    std = math.sqrt(2.0 * groups / (nin * volume))
    return nn.Parameter(torch.empty(volume, groups, nin // groups, nout // groups).normal_(0, std))


class SubmanifoldConvolution(nn.Module):
    """module_factory.py:377-414.  Output sites == input sites."""

    def __init__(self, dimension, nIn, nOut, filter_size, bias, groups=1):
        super().__init__()
        self.dimension, self.nIn, self.nOut = dimension, nIn, nOut
        self.filter_size = _triple(filter_size)
        if any(f % 2 == 0 for f in self.filter_size):
            raise RuntimeError("SubmanifoldConvolution needs odd filter sizes")
        self.weight = _conv_weight(self.filter_size[0] * self.filter_size[1] * self.filter_size[2], nIn, nOut, groups)
        self.bias = nn.Parameter(torch.zeros(nOut)) if bias else None

    def forward(self, x):
        lvl = x.metadata.level(x.spatial_size)
        m = lvl.subm_map(self.filter_size, tile_book=False)
        if self.nIn == self.nOut and self.filter_size == (3, 3, 3):
            lvl.ensure_tile_book(self.nIn)
        f = F.ConvFunction.run(x.features, self.weight, self.bias, m, m, lvl.n, 1)
        return _like(x, f)

    def input_spatial_size(self, out_size):
        return out_size


ValidConvolution = SubmanifoldConvolution


class Convolution(nn.Module):
    """module_factory.py:221-241 (filter == stride)."""

    def __init__(self, dimension, nIn, nOut, filter_size, filter_stride, bias, groups=1):
        super().__init__()
        self.dimension, self.nIn, self.nOut = dimension, nIn, nOut
        self.filter_size, self.filter_stride = _triple(filter_size), _triple(filter_stride)
        self.weight = _conv_weight(self.filter_size[0] * self.filter_size[1] * self.filter_size[2], nIn, nOut, groups)
        self.bias = nn.Parameter(torch.zeros(nOut)) if bias else None

    def forward(self, x):
        md = x.metadata
        r = md.strided_rules(x.spatial_size, self.filter_size, self.filter_stride)
        n_out = md.levels[r.out_key].n
        f = F.ConvFunction.run(x.features, self.weight, self.bias, r.cmap, r.dmap, n_out, 0)
        return SparseConvNetTensor(f, md, r.out_size)

    def input_spatial_size(self, out_size):
        return (out_size - 1) * torch.tensor(self.filter_stride) + torch.tensor(self.filter_size)


class Deconvolution(nn.Module):
    """module_factory.py:244-271: transpose of Convolution; output active set = the finer grid
    that must already exist in the Metadata."""

    def __init__(self, dimension, nIn, nOut, filter_size, filter_stride, bias, groups=1):
        super().__init__()
        self.dimension, self.nIn, self.nOut = dimension, nIn, nOut
        self.filter_size, self.filter_stride = _triple(filter_size), _triple(filter_stride)
        self.weight = _conv_weight(self.filter_size[0] * self.filter_size[1] * self.filter_size[2], nIn, nOut, groups)
        self.bias = nn.Parameter(torch.zeros(nOut)) if bias else None

    def forward(self, x):
        md = x.metadata
        in_size = size_key(x.spatial_size)
        out_size = tuple((i - 1) * s + f for i, s, f in zip(in_size, self.filter_stride, self.filter_size))
        if out_size not in md.levels:
            raise RuntimeError("Deconvolution: no active set at spatial size %s in this Metadata" % (out_size,))
        r = md.strided_rules(out_size, self.filter_size, self.filter_stride)
        lvl = md.levels[out_size]
        if getattr(lvl, "size_tensor", None) is None:
            lvl.size_tensor = torch.tensor(out_size, dtype=torch.long)
        f = F.ConvFunction.run(x.features, self.weight, self.bias, r.dmap, r.cmap, lvl.n, 0)
        return SparseConvNetTensor(f, md, lvl.size_tensor)


class NetworkInNetwork(nn.Module):
    """module_factory.py:357-374: out = x @ W + b."""

    def __init__(self, nIn, nOut, bias):
        super().__init__()
        self.nIn, self.nOut = nIn, nOut
        self.weight = nn.Parameter(torch.empty(nIn, nOut).normal_(0, math.sqrt(2.0 / nIn)))
        self.bias = nn.Parameter(torch.zeros(nOut)) if bias else None

    def forward(self, x):
        f = F.ConvFunction.run(x.features, self.weight, self.bias, None, None, x.features.shape[0], 0)
        return _like(x, f)

    def input_spatial_size(self, out_size):
        return out_size


class _Pooling(nn.Module):
    IS_MAX = True

    def __init__(self, dimension, pool_size, pool_stride, nFeaturesToDrop=0):
        super().__init__()
        self.pool_size, self.pool_stride = _triple(pool_size), _triple(pool_stride)

    def forward(self, x):
        md = x.metadata
        r = md.strided_rules(x.spatial_size, self.pool_size, self.pool_stride)
        n_out = md.levels[r.out_key].n
        vol = self.pool_size[0] * self.pool_size[1] * self.pool_size[2]
        f = F.PoolFunction.run(x.features, r, n_out, self.IS_MAX, 1.0 / vol)
        return SparseConvNetTensor(f, md, r.out_size)


class MaxPooling(_Pooling):
    """module_factory.py:315-333: max over ACTIVE children."""
    IS_MAX = True


class AveragePooling(_Pooling):
    """module_factory.py:336-354: sum over active children / pool VOLUME."""
    IS_MAX = False


class UnPooling(nn.Module):
    """SparseConvNet UnPooling(dimension, pool_size, pool_stride): the inverse site map of MaxPooling/AveragePooling --
    every active site of the FINER level (which must already exist in the Metadata, as for Deconvolution) receives the
    feature row of its coarse parent.  Not used by the shipped ndsis configurations; named by the build's north star."""

    def __init__(self, dimension, pool_size, pool_stride, nFeaturesToDrop=0):
        super().__init__()
        self.pool_size, self.pool_stride = _triple(pool_size), _triple(pool_stride)

    def forward(self, x):
        md = x.metadata
        ck = size_key(x.spatial_size)
        fine = tuple((o - 1) * s + f for o, f, s in zip(ck, self.pool_size, self.pool_stride))
        if fine not in md.levels:
            raise RuntimeError("UnPooling needs the finer grid to exist in the Metadata")
        r = md.strided_rules(fine, self.pool_size, self.pool_stride)
        if r.out_key != ck:
            raise RuntimeError("UnPooling: spatial size %s is not the pooled size of %s" % (ck, fine))
        f = F.UnPoolFunction.run(x.features, r.parent_row, md.levels[ck].n)
        return SparseConvNetTensor(f, md, torch.tensor(fine, dtype=torch.long))


class SparseToDense(nn.Module):
    """module_factory.py:429-435 -> [B, C, X, Y, Z]."""

    def __init__(self, dimension, nPlanes):
        super().__init__()
        self.nPlanes = nPlanes

    def forward(self, x):
        md = x.metadata
        return F.SparseToDenseFunction.run(x.features, md.level(x.spatial_size), md.n_samples,
                                             size_key(x.spatial_size))


class OutputLayer(nn.Module):
    def __init__(self, dimension):
        super().__init__()
        self.dimension = dimension

    def forward(self, x):
        return F.OutputLayerFunction.run(self.dimension, x.metadata, x.features)


class InputLayer(nn.Module):
    def __init__(self, dimension, spatial_size, mode=3):
        super().__init__()
        self.dimension, self.mode = dimension, mode
        self.spatial_size = torch.as_tensor(spatial_size, dtype=torch.long)

    def forward(self, inp):
        coords, feats = inp[0], inp[1]
        bs = inp[2] if len(inp) > 2 else 0
        md = Metadata(self.dimension)
        f = F.InputLayerFunction.run(self.dimension, md, self.spatial_size, coords, feats, bs, self.mode)
        return SparseConvNetTensor(f, md, self.spatial_size)
