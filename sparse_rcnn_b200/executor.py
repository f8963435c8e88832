"""Host side of the native sparse U-Net executor (csrc/unet_exec.cu: scn_unet_plan / scn_unet_fwd / scn_unet_bwd).

`UNetProgram.compile(encoder, decoder)` reads the module tree that the reference assembles for its feature extractor
(module_factory.py:438-578, 789-830; run through model.py:414-446) -- a list of encoder levels
`Sequential(Sequential(entry convolution), Sequential(residual units))` and a list of decoder levels
`SkipConnectionReuniter(Sequential(ReLU, Deconvolution), JoinTable, NetworkInNetwork, Sequential(residual units))` --
into a LAYER TABLE of parameter / packed-image pointers.  Per scene, `run()` writes a GEOMETRY TABLE (rows and neighbour maps
of every level), allocates ONE activation arena and makes ONE C-ABI call; the backward is one more call (two when the
gradients are bucketed for an allreduce, so that the decoder's bucket is reduced while the encoder's backward runs).  The whole
encoder + decoder is one autograd node.  Same kernels in the same order as the module-by-module path, bit-identical results;
anything the compiler does not recognise leaves the module path in charge (`compile` returns None).
"""
import numpy as np
import torch

from . import _lib
from .scn import functions as F
from .scn import layers as L
from .scn.metadata import _ptr, _stream

ENABLED = {"unet": True}      # False: always run the module graph (tests compare both)
ENCODER_PHASE_SPLIT = 2       # bucketed backward: encoder levels >= this finish in the second call, the finer ones in the third


def _units_of(stage):
    """[(conv1, conv2), ...] if `stage` is a Sequential of residual units (possibly empty), else None."""
    if type(stage) is not L.Sequential:
        return None
    units = []
    for m in stage._modules.values():
        u = L._match_residual_unit(m)
        if u is None:
            return None
        units.append(u)
    return units


def _entry_of(entry):
    """The single convolution of an encoder level's entry Sequential: (kind, module) or None."""
    if type(entry) is L.Identity:
        return 0, None
    if type(entry) is not L.Sequential or len(entry) != 1:
        return None
    c = entry[0]
    if type(c) is L.SubmanifoldConvolution and c.filter_size in ((1, 1, 1), (3, 3, 3)):
        return 1, c
    if type(c) is L.Convolution and c.filter_size == (2, 2, 2) and c.filter_stride == (2, 2, 2):
        return 2, c
    return None


class UNetProgram:
    def __init__(self):
        self.enc, self.dec = [], []       # [(kind, conv or None, [(c1, c2), ...])], [(deconv, nin, units)]
        self.params = []                  # parameters in table order (None where a layer has no bias)
        self._table_key = None
        self._net = None

    # ------------------------------------------------------------------ compile
    @classmethod
    def compile(cls, encoder_levels, decoder_levels):
        """encoder_levels: modules in execution order; decoder_levels: modules with .input_stage / .combiner /
        .channel_changer / .output_stage (coarsest first), or an empty list for an encoder-only program."""
        prog = cls()
        for i, lev in enumerate(encoder_levels):
            if type(lev) is L.Identity:
                if i != 0:
                    return None
                prog.enc.append((0, None, []))
                continue
            if type(lev) is not L.Sequential or len(lev) != 2:
                return None
            entry, units = _entry_of(lev[0]), _units_of(lev[1])
            if entry is None or units is None or len(units) > 4:
                return None
            kind, conv = entry
            if (kind == 0) or (kind == 2 and i == 0):
                return None
            if units and units[0][0].nIn != conv.nOut:
                return None
            prog.enc.append((kind, conv, units))
        if not prog.enc or len(prog.enc) > 8:
            return None
        if decoder_levels and len(decoder_levels) != len(prog.enc) - 1:
            return None
        for lev in decoder_levels:
            up = getattr(lev, "input_stage", None)
            if type(up) is not L.Sequential or len(up) != 2 or type(up[0]) is not L.ReLU or type(up[1]) is not L.Deconvolution:
                return None
            d, nin, units = up[1], getattr(lev, "channel_changer", None), _units_of(getattr(lev, "output_stage", None))
            if d.filter_size != (2, 2, 2) or d.filter_stride != (2, 2, 2) or type(getattr(lev, "combiner", None)) is not L.JoinTable:
                return None
            if type(nin) is not L.NetworkInNetwork or units is None or len(units) > 4:
                return None
            if units and units[0][0].nIn != nin.nOut:
                return None
            prog.dec.append((d, nin, units))
        for kind, conv, units in prog.enc:
            if conv is not None:
                prog.params += [conv.weight, conv.bias]
            for c1, c2 in units:
                prog.params += [c1.weight, c1.bias, c2.weight, c2.bias]
        for d, nin, units in prog.dec:
            prog.params += [d.weight, d.bias, nin.weight, nin.bias]
            for c1, c2 in units:
                prog.params += [c1.weight, c1.bias, c2.weight, c2.bias]
        prog.weights = [p for p in prog.params if p is not None and p.dim() >= 2]
        n_enc_slots = sum((2 if conv is not None else 0) + 4 * len(units) for kind, conv, units in prog.enc)
        prog.n_encoder_params = sum(1 for p in prog.params[:n_enc_slots] if p is not None)
        return prog

    def encoder_params_below(self, level):
        """Number of (non-None) parameters of the encoder levels < level, in table order."""
        slots = sum((2 if conv is not None else 0) + 4 * len(units) for kind, conv, units in self.enc[:level])
        return sum(1 for p in self.params[:slots] if p is not None)

    def channels_in(self):
        kind, conv, _ = self.enc[0]
        return conv.nIn if conv is not None else None

    # ------------------------------------------------------------------ layer table
    def _images(self, w, K, cin, cout, submanifold, tf32, with_grad):
        if not tf32:
            return 0, 0
        f = F._image_entry(w, K, cin, cout, 0, 0)[0]
        b = F._image_entry(w, K, cout, cin, 1, 1 if submanifold else 0)[0] if with_grad else None
        return f.data_ptr(), (b.data_ptr() if b is not None else 0)

    def _table(self, tf32, with_grad, c_in):
        key = (tf32, with_grad, c_in, tuple(0 if p is None else p.data_ptr() for p in self.params))
        if key == self._table_key:
            return self._net
        P = lambda t: 0 if t is None else t.data_ptr()
        t = [len(self.enc)]

        def unit_rows(units):
            rows = []
            for c1, c2 in units:
                c = c1.nOut
                f1, b1 = self._images(c1.weight, 27, c, c, True, tf32, with_grad)
                f2, b2 = self._images(c2.weight, 27, c, c, True, tf32, with_grad)
                rows += [P(c1.weight), P(c1.bias), P(c2.weight), P(c2.bias), f1, f2, b1, b2]
            return rows

        for kind, conv, units in self.enc:
            if conv is None:
                t += [0, 1, c_in, c_in, 0, 0, 0, 0, 0]
                continue
            K = conv.weight.shape[0]
            f, b = self._images(conv.weight, K, conv.nIn, conv.nOut, kind == 1, tf32, with_grad)
            t += [kind, K, conv.nIn, conv.nOut, P(conv.weight), P(conv.bias), f, b, len(units)] + unit_rows(units)
        for d, nin, units in self.dec:
            f, b = self._images(d.weight, 8, d.nIn, d.nOut, False, tf32, with_grad)
            t += [8, d.nIn, d.nOut, P(d.weight), P(d.bias), f, b]
            f, b = self._images(nin.weight, 1, nin.nIn, nin.nOut, False, tf32, with_grad)
            t += [nin.nIn, nin.nOut, P(nin.weight), P(nin.bias), f, b, len(units)] + unit_rows(units)
        self._net = np.asarray(t, dtype=np.int64)
        self._table_key = key
        return self._net

    # ------------------------------------------------------------------ geometry table
    def _geometry(self, md, spatial_size):
        """The geometry table of one scene from its Metadata (levels and strided rules are cached there; built on demand)
        and the per-level spatial sizes.  A fresh array per call: inference runs scenes from several host threads."""
        g = np.zeros(8 * 4, dtype=np.int64)
        sizes = []
        size = spatial_size
        nl = len(self.enc)
        for i in range(nl):
            lvl = md.level(size)
            sizes.append(size)
            g[4 * i] = lvl.n
            needs_subm = bool(self.enc[i][2]) or (self.enc[i][0] == 1 and self.enc[i][1].filter_size == (3, 3, 3)) or \
                (i < nl - 1 and self.dec and bool(self.dec[nl - 2 - i][2]))
            m = lvl.subm_map(3, tile_book=False) if (needs_subm and lvl.n) else None
            if m is not None:      # tile book only where a layer of a width the tile-local kernels implement runs
                units = self.enc[i][2] or (self.dec[nl - 2 - i][2] if (i < nl - 1 and self.dec) else [])
                if units:
                    lvl.ensure_tile_book(units[0][0].nIn)
            g[4 * i + 1] = 0 if m is None else m.data_ptr()
            if i < nl - 1:
                r = md.strided_rules(size, 2, 2)
                g[4 * i + 2], g[4 * i + 3] = r.cmap.data_ptr(), r.dmap.data_ptr()
                size = r.out_size
            else:
                g[4 * i + 2] = g[4 * i + 3] = 0
        return g, sizes

    # ------------------------------------------------------------------ run
    def run(self, x):
        """x: SparseConvNetTensor at the finest level.  -> (encoder outputs [E_0 ..], decoder outputs [D_0 ..]) as
        SparseConvNetTensors (views into one arena)."""
        feats = F._check(x.features)
        md = x.metadata
        tf32 = F.get_precision() == "tf32"
        with_grad = torch.is_grad_enabled()
        c_in = feats.shape[1]
        if self.channels_in() is not None and self.channels_in() != c_in:
            raise RuntimeError("sparse U-Net expects %d input planes, got %d" % (self.channels_in(), c_in))
        geo, sizes = self._geometry(md, x.spatial_size)
        net = self._table(tf32, with_grad, c_in)
        if tf32:
            F.pack_all(self.weights)      # one launch for every stale packed image (none in steady-state inference)
        outs = UNetFunction.run(feats, self, net, geo, md, *[p for p in self.params if p is not None])
        nl = len(self.enc)
        T = L.SparseConvNetTensor
        enc = [T(outs[i], md, sizes[i]) for i in range(nl)]
        dec = [T(outs[nl + j], md, sizes[nl - 2 - j]) for j in range(len(self.dec))]
        return enc, dec


class UNetFunction(F.Function):
    """The whole encoder + decoder as one autograd node around scn_unet_fwd / scn_unet_bwd."""

    @staticmethod
    def forward(ctx, x, prog, net, geo, md, *params):
        nl, nd = len(prog.enc), len(prog.dec)
        plan = np.zeros(2 * 8 + 2, dtype=np.int64)
        _lib.call("scn_unet_plan", net.ctypes.data, geo.ctypes.data, plan.ctypes.data)
        offs = [int(v) for v in plan[:2 * nl + 1]]
        total, n_out = offs[2 * nl - 1], nl + nd
        arena = torch.empty(max(total, 1), dtype=torch.float32, device=x.device)
        tf32 = F.get_precision() == "tf32"
        _lib.call("scn_unet_fwd", net.ctypes.data, geo.ctypes.data, _ptr(x), _ptr(arena), nd, int(tf32), _stream())
        outs = []
        for k in range(n_out):
            lvl = k if k < nl else nl - 2 - (k - nl)
            n = int(geo[4 * lvl])
            if k < nl:
                kind, conv, units = prog.enc[k]
                c = conv.nOut if conv is not None else x.shape[1]
            else:
                c = prog.dec[k - nl][1].nOut
            outs.append(x if offs[k] < 0 else arena[offs[k]:offs[k] + n * c].view(n, c))
        ctx.prog, ctx.net, ctx.geo, ctx.md = prog, net, geo, md      # md keeps the maps alive
        ctx.bwd_floats = int(plan[2 * nl])
        ctx.save_for_backward(x, arena)
        ctx.set_materialize_grads(False)      # outputs nobody used arrive as None, not as zero tensors
        if offs[0] < 0:
            outs[0] = x.view_as(x)            # pass-through level 0: a view, never the input object itself
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        x, arena = ctx.saved_tensors
        prog, net, geo = ctx.prog, ctx.net, ctx.geo
        dev = arena.device
        # precision is read when the backward runs (like ConvFunction.backward): the layer table of THAT mode (tests run an
        # fp32 forward and a tf32 backward over the same saved activations), its transposed images packed
        tf32 = F.get_precision() == "tf32"
        net = prog._table(tf32, True, x.shape[1])
        if tf32:
            F.pack_all(prog.weights)
        barena = torch.empty(max(ctx.bwd_floats, 1), dtype=torch.float32, device=dev)
        keep = [None if g is None else F._check(g) for g in grads]
        seeds = np.asarray([0 if g is None else g.data_ptr() for g in keep], dtype=np.int64)
        params = [p for p in prog.params if p is not None]
        need = ctx.needs_input_grad[5:]
        direct = [F._direct_grad(p) if want else None for p, want in zip(params, need)]
        bucketed = all(d is not None for d, want in zip(direct, need) if want) and any(need)
        flat = None
        if bucketed:
            bufs = direct
        else:      # plain autograd: fresh zeroed gradients, returned to the engine
            sizes = [(p.numel() + 3) // 4 * 4 if want else 0 for p, want in zip(params, need)]
            flat = torch.zeros(max(sum(sizes), 1), dtype=torch.float32, device=dev)
            bufs, at = [], 0
            for p, want, sz in zip(params, need, sizes):
                bufs.append(flat[at:at + p.numel()].view_as(p) if want else None)
                at += sz
        pg, it = [], iter(bufs)
        for p in prog.params:      # table order, a zero where a layer has no bias or the gradient is not wanted
            b = None if p is None else next(it)
            pg.append(0 if b is None else b.data_ptr())
        pg = np.asarray(pg, dtype=np.int64)
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        args = (net.ctypes.data, geo.ctypes.data, _ptr(x), _ptr(arena), _ptr(barena), seeds.ctypes.data, pg.ctypes.data, _ptr(gx))
        s = _stream()
        if bucketed and prog.dec:
            # three calls: decoder | coarse encoder levels | fine encoder levels; after each one the hooks of its parameters
            # fire, so that a gradient bucket cut along these boundaries is all-reduced while the rest of the backward runs
            pw = list(zip(params, need))
            n_enc, split = prog.n_encoder_params, min(ENCODER_PHASE_SPLIT, len(prog.enc))
            n_fine = prog.encoder_params_below(split)

            def fire(lo, hi):
                for p, want in pw[lo:hi]:
                    if want:
                        p._scn_grad_hook(p)
            _lib.call("scn_unet_bwd", *args, 1 | (split << 8), int(tf32), s)
            fire(n_enc, len(pw))
            _lib.call("scn_unet_bwd", *args, 2 | (split << 8), int(tf32), s)
            fire(n_fine, n_enc)
            _lib.call("scn_unet_bwd", *args, 4 | (split << 8), int(tf32), s)
            fire(0, n_fine)
        else:
            _lib.call("scn_unet_bwd", *args, 7, int(tf32), s)
            if bucketed:
                for p, want in zip(params, need):
                    if want:
                        p._scn_grad_hook(p)
        pgrads = [None] * len(params) if bucketed else [b for b in bufs]
        return (gx, None, None, None, None, *pgrads)
