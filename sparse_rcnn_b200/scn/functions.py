"""autograd Functions of the sparse hot path; each forward/backward is one or two C-ABI calls.

Upstream equivalents are the `X_updateOutput / X_updateGradInput / X_backward` triplets of
SparseConvNet's pybind module; arithmetic follows SURVEY.md A5 (bias first, offsets accumulated
in fp32).  All tensors are fp32 CUDA; there is no CPU path here.
"""
import torch

from .. import _lib
from .metadata import _ptr, _stream

EPI_RELU, EPI_ADD, EPI_MASK, EPI_ROUND = 1, 2, 4, 8

_state = {"precision": "tf32"}


def set_precision(mode):
    """'tf32' (tcgen05 tensor cores, rel <= 2e-3) or 'fp32' (exact FFMA verification mode, <= 1e-5)."""
    if mode not in ("tf32", "fp32"):
        raise ValueError("precision must be 'tf32' or 'fp32'")
    _state["precision"] = mode


def get_precision():
    return _state["precision"]


class _NullCtx:
    """Stand-in for the autograd context when no graph is recorded."""
    needs_input_grad = ()

    def save_for_backward(self, *tensors):
        pass

    def set_materialize_grads(self, value):
        pass


class Function(torch.autograd.Function):
    """torch.autograd.Function plus `run`: the same as `apply` while gradients are recorded; under `torch.no_grad()`
    (inference) it calls `forward` directly -- `apply` costs ~10 us of bookkeeping per call even then, and the inference pass
    is host bound (68 calls per scene).  `apply` stays what the reference calls (ioLayers.*Function.apply)."""

    @classmethod
    def run(cls, *args):
        if torch.is_grad_enabled():
            return cls.apply(*args)
        return cls.forward(_NullCtx(), *args)


def _check(x):
    if not x.is_cuda:
        raise RuntimeError("sparse_rcnn_b200 ops need CUDA tensors (there is no CPU fallback)")
    if x.dtype != torch.float32:
        raise RuntimeError("features must be float32 (got %s)" % x.dtype)
    return x.contiguous()


def weights_changed():
    """Mark every cached packed weight image stale.  Called from a global optimizer step post-hook (below), so any
    `optimizer.step()` in the process is followed by a re-pack: version counters alone are not enough -- measured on
    torch 2.11, `Adam(fused=True)` updates parameters without touching `_version`, and an optimizer over the flat buffers of
    `parallel.GradientBuckets.flatten_parameters` never touches the module Parameters at all."""
    _state["generation"] += 1


_state["generation"] = 0
try:
    from torch.optim.optimizer import register_optimizer_step_post_hook as _register_step_hook
    _register_step_hook(lambda optimizer, args, kwargs: weights_changed())
except ImportError:      # older torch: trainers call weights_changed() themselves (pipeline.BackboneTrainer does anyway)
    pass


def _wver(weight):
    """Staleness key of a weight tensor: its own version counter and storage (load_state_dict, `.data =`), the version of
    the flat buffer it is a view of, and the global optimizer-step generation."""
    flat = getattr(weight, "_scn_flat", None)
    return (weight._version, weight.data_ptr(), -1 if flat is None else flat._version, _state["generation"])


def _image_entry(weight, K, cin, cout, transpose, reverse):
    """Persistent packed-image buffer of a weight tensor (one per orientation): [image, packed version, meta]."""
    cache = getattr(weight, "_scn_img", None)
    if cache is None:
        cache = {}
        try:
            weight._scn_img = cache
        except AttributeError:
            pass
    key = (transpose, reverse)
    hit = cache.get(key)
    if hit is None:
        nbytes = int(_lib.raw("scn_conv_weight_image_bytes")(K, cin, cout))
        hit = [torch.empty(nbytes, dtype=torch.uint8, device=weight.device), None, (K, cin, cout)]
        cache[key] = hit
    return hit


def _image(weight, K, cin, cout, transpose, reverse):
    """Packed (TF32-rounded, swizzled) weight image; re-packed when the tensor was modified since."""
    hit = _image_entry(weight, K, cin, cout, transpose, reverse)
    ver = _wver(weight)
    if hit[1] != ver:
        _lib.call("scn_conv_pack_weights", _ptr(weight), K, cin, cout, transpose, reverse, _ptr(hit[0]), _stream())
        hit[1] = ver
    return hit[0]


_pack_tables = {}


def pack_all(weights):
    """Re-pack every stale image of the given weight tensors in ONE launch (call after optimizer.step(): each layer
    otherwise re-packs both orientations lazily, ~120 launches per training step).  Images that were never used are
    not created here."""
    entries, rows = [], []
    for w in weights:
        cache = getattr(w, "_scn_img", None)
        if not cache:
            continue
        ver = _wver(w)
        for (transpose, reverse), hit in cache.items():
            if hit[1] != ver:
                entries.append((hit, ver))
                K, cin, cout = hit[2]
                rows.append((w.data_ptr(), hit[0].data_ptr(), K, cin, cout, transpose, reverse))
    if not rows:
        return 0
    key = tuple(r[:2] for r in rows)
    dev = entries[0][0][0].device
    table = _pack_tables.get(dev)
    if table is None or table[0] != key:
        table = (key, torch.tensor(rows, dtype=torch.int64).to(dev))
        _pack_tables[dev] = table
    _lib.call("scn_conv_pack_weights_multi", _ptr(table[1]), len(rows), _stream())
    for hit, ver in entries:
        hit[1] = ver
    return len(rows)


def prepack_forward(module):
    """Create and pack the FORWARD weight image of every convolution parameter under `module` on the current stream
    (one launch).  Inference from several host threads / streams (pipeline.SparseInference.run_many) calls this on the
    caller's stream first and makes the workers wait for it: a lazily created image is marked packed by the first thread
    that touches it BEFORE its pack kernel has run on that thread's stream, so a second stream could read it unpacked."""
    if _state["precision"] != "tf32":
        return 0
    weights = []
    for m in module.modules():
        w = getattr(m, "weight", None)
        if w is None or not w.is_cuda or not hasattr(m, "nOut") or not hasattr(m, "nIn"):
            continue
        K = w.shape[0] if w.dim() > 2 else 1
        cin, cout = w.shape[-2], w.shape[-1]
        if cout <= 256:
            _image_entry(w, K, cin, cout, 0, 0)
            weights.append(w)
    return pack_all(weights)


def _direct_grad(param):
    """Gradient buffer to accumulate into directly (parallel.GradientBuckets keeps .grad as zeroed views of flat buckets and
    registers `_scn_grad_hook`): saves a zero-fill and an AccumulateGrad add per parameter and step.  None -> regular path."""
    if param is None or getattr(param, "_scn_grad_hook", None) is None:
        return None
    g = param.grad
    if g is None or not g.is_contiguous() or g.dtype != torch.float32:
        return None
    return g


def conv_gemm(x, weight, K, cin, cout, fmap, n_out, bias=None, transpose=0, reverse=0, residual=None, relu=False,
              mask=None, round_out=False):
    """out[r] = epi(bias + sum_o x[fmap[o][r]] . W_eff[o]);  W_eff = weight (transposed / offset-reversed).
    cin/cout are the GEMM widths (after any transpose).  Epilogue order: bias, mask (mask > 0 ? v : 0), + residual,
    ReLU, TF32 rounding (tf32 mode only; the result is then marked as a valid tensor-core operand)."""
    out = torch.empty((n_out, cout), dtype=torch.float32, device=x.device)
    if n_out == 0:
        return out
    tf32 = _state["precision"] == "tf32" and cout <= 256
    epi = (EPI_RELU if relu else 0) | (EPI_ADD if residual is not None else 0) | (EPI_MASK if mask is not None else 0) | \
        (EPI_ROUND if (round_out and tf32) else 0)
    ld_res = residual.stride(0) if residual is not None else 0
    ld_mask = mask.stride(0) if mask is not None else 0
    s = _stream()
    if tf32:
        img = _image(weight, K, cin, cout, transpose, reverse)
        x = tf32_exact(x)           # cp.async gathers are truncated by the MMA: round the operand once
        _lib.call("scn_conv_fwd_tf32", _ptr(x), x.stride(0), cin, x.shape[0], _ptr(fmap), n_out, K, _ptr(img), _ptr(bias),
                  _ptr(residual), ld_res, _ptr(mask), ld_mask, _ptr(out), cout, cout, epi, s)
        if round_out:
            out._scn_tf32 = True
    else:
        _lib.call("scn_conv_fwd_fp32", _ptr(x), x.stride(0), cin, _ptr(fmap), n_out, K, _ptr(weight), transpose,
                  reverse, _ptr(bias), _ptr(residual), ld_res, _ptr(mask), ld_mask, _ptr(out), cout, cout, epi, s)
    return out


def tf32_exact(x):
    """tcgen05 kind::tf32 truncates its fp32 operands (a systematic shrink of ~3e-4 per layer).  The
    gathered operand is therefore required to be TF32-representable: tensors produced by kernels
    that already round (ReLU in tf32 mode) carry a mark, anything else gets one rounding pass."""
    if getattr(x, "_scn_tf32", False):
        return x
    y = torch.empty_like(x)
    _lib.call("scn_round_tf32", _ptr(x), _ptr(y), x.numel(), _stream())
    y._scn_tf32 = True
    return y


def _mark(t):
    if _state["precision"] == "tf32":
        t._scn_tf32 = True
    return t


def _image_state(weight, K, cin, cout, transpose, reverse):
    """(packed-image buffer, must it be re-packed); marks it packed (the C call that follows does the packing --
    callers only ask when that call really runs its kernels, i.e. not for empty inputs)."""
    hit = _image_entry(weight, K, cin, cout, transpose, reverse)
    ver = _wver(weight)
    stale = hit[1] != ver
    hit[1] = ver
    return hit[0], stale


class ConvFunction(Function):
    """Shared by SubmanifoldConvolution / Convolution / Deconvolution / NetworkInNetwork; each direction is ONE C-ABI call
    (scn_conv_layer_fwd / _bwd: operand rounding, stale-image packing, gather-GEMM(s), weight + bias gradient).

    fmap [K, n_out]: forward map; bmap [K, n_in]: map of the input gradient (for a submanifold
    conv bmap is fmap with the offsets reversed, expressed through `reverse_bwd`)."""

    @staticmethod
    def forward(ctx, x, weight, bias, fmap, bmap, n_out, reverse_bwd):
        x = _check(x)
        w = weight          # keep the Parameter object: the packed image is cached on it
        K, cin, cout = w.shape[0] if w.dim() > 2 else 1, w.shape[-2], w.shape[-1]
        if x.shape[1] != cin:
            raise RuntimeError("convolution expects %d input planes, got %d" % (cin, x.shape[1]))
        ctx.save_for_backward(x, weight)
        ctx.maps = (fmap, bmap)
        ctx.dims = (K, cin, cout, n_out, x.shape[0], reverse_bwd, bias is not None)
        ctx.bias_param = bias
        tf32 = _state["precision"] == "tf32" and cout <= 256
        out = torch.empty((n_out, cout), dtype=torch.float32, device=x.device)
        if n_out == 0:
            return out
        img = xr = None
        stale, exact = False, True
        if tf32:
            img, stale = _image_state(w, K, cin, cout, 0, 0)
            exact = bool(getattr(x, "_scn_tf32", False))
            if not exact:
                xr = torch.empty_like(x)
        _lib.call("scn_conv_layer_fwd", _ptr(x), x.stride(0), x.shape[0], cin, int(exact), _ptr(xr), _ptr(fmap), n_out, K,
                  _ptr(w), _ptr(img), int(stale), _ptr(bias), _ptr(out), cout, int(tf32), _stream())
        return out

    @staticmethod
    def backward(ctx, go):
        x, weight = ctx.saved_tensors
        fmap, bmap = ctx.maps
        K, cin, cout, n_out, n_in, reverse_bwd, has_bias = ctx.dims
        go = _check(go)
        w = weight
        gx = gw = gb = None
        tf32 = _state["precision"] == "tf32" and cin <= 256
        want_x = ctx.needs_input_grad[0]
        want_w = ctx.needs_input_grad[1]
        want_b = has_bias and ctx.needs_input_grad[2]
        dev = go.device
        if want_x:
            gx = torch.empty((n_in, cin), dtype=torch.float32, device=dev)
        b = ctx.bias_param
        dw = _direct_grad(w) if want_w else None
        db = _direct_grad(b) if want_b else None
        gw_buf = (dw if dw is not None else torch.zeros_like(w)) if want_w else None
        gb_buf = (db if db is not None else torch.zeros(cout, dtype=torch.float32, device=dev)) if want_b else None
        img = gr = None
        stale, exact = False, True
        if tf32:
            exact = bool(getattr(go, "_scn_tf32", False))
            if not exact and (want_x or want_w) and n_out:
                gr = torch.empty_like(go)
            if want_x:
                img, stale = _image_state(w, K, cout, cin, 1, reverse_bwd)
        _lib.call("scn_conv_layer_bwd", _ptr(go), n_out, cout, int(exact), _ptr(gr), _ptr(x), x.stride(0), n_in, cin,
                  _ptr(fmap), _ptr(bmap), K, _ptr(w), _ptr(img), int(stale), int(reverse_bwd), _ptr(gx), _ptr(gw_buf),
                  _ptr(gb_buf), int(tf32), _stream())
        if want_w:
            if dw is not None:
                w._scn_grad_hook(w)
            else:
                gw = gw_buf
        if want_b:
            if db is not None:
                b._scn_grad_hook(b)
            else:
                gb = gb_buf
        return gx, gw, gb, None, None, None, None


def _wgrad(x, fmap, go, K, cin, cout, n_out, like):
    gw = torch.zeros_like(like)
    if n_out:
        _lib.call("scn_conv_bwd_weight", _ptr(x), x.stride(0), cin, _ptr(fmap), n_out, K, _ptr(go), go.stride(0), cout,
                  _ptr(gw), None, 1 if _state["precision"] == "tf32" else 0, _stream())
    return gw


def _bgrad(go, n, cout):
    gb = torch.empty(cout, dtype=torch.float32, device=go.device)
    if n:
        _lib.call("scn_col_sum", _ptr(go), go.stride(0), n, cout, _ptr(gb), _stream())
    else:
        gb.zero_()
    return gb


def relu_round(x):
    """r = relu(x), additionally rounded to TF32 (and marked) in tf32 mode."""
    y = torch.empty_like(x)
    tf32 = _state["precision"] == "tf32"
    _lib.call("scn_relu_fwd", _ptr(x), _ptr(y), x.numel(), int(tf32), _stream())
    if tf32:
        y._scn_tf32 = True
    return y


def _image_buf(weight, K, cin, cout, key):
    """(packed-image buffer, must it be re-packed) for the fused unit: key 'f' = forward, 'b' = transposed + reversed."""
    tr = 0 if key == "f" else 1
    return _image_state(weight, K, cin, cout, tr, tr)


def _unit_forward(x, w1, b1, w2, b2, fmap, n, tf32):
    """One residual unit forward (one C-ABI call): returns (r = relu(x), h = relu(conv1(r)), y = x + conv2(h))."""
    K, c = w1.shape[0], w1.shape[-1]
    dev = x.device
    r = torch.empty((n, c), dtype=torch.float32, device=dev)
    h = torch.empty((n, c), dtype=torch.float32, device=dev)
    y = torch.empty((n, c), dtype=torch.float32, device=dev)
    if tf32:
        i1, s1 = _image_buf(w1, K, c, c, "f")
        i2, s2 = _image_buf(w2, K, c, c, "f")
    else:
        i1 = i2 = None
        s1 = s2 = False
    _lib.call("scn_residual_unit_fwd", _ptr(x), n, c, _ptr(fmap), K, _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), _ptr(i1),
              _ptr(i2), int(s1 or s2), _ptr(r), _ptr(h), _ptr(y), int(tf32), _stream())
    return r, h, y


def _unit_backward(gy, r, h, w1, b1, w2, b2, fmap, n, tf32, need_x, need_p):
    """One residual unit backward (one C-ABI call).  need_p: (w1, b1, w2, b2) wanted.  Returns (gx, [gw1, gb1, gw2, gb2]);
    gradients that went straight into the parameters' buckets come back as None."""
    K, c = w1.shape[0], w1.shape[-1]
    dev = gy.device
    new = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
    gyr = new(n, c) if tf32 else None
    gh = new(n, c)
    gx = new(n, c) if need_x else None
    wanted = [(w1, need_p[0]), (b1, b1 is not None and need_p[1]), (w2, need_p[2]), (b2, b2 is not None and need_p[3])]
    direct = [_direct_grad(p) if want else None for p, want in wanted]
    # all-or-nothing: the C call either accumulates into every parameter gradient or overwrites fresh buffers
    accumulate = all(d is not None for d, (p, want) in zip(direct, wanted) if want) and any(want for _, want in wanted)
    bufs = direct if accumulate else [(torch.empty_like(p) if want else None) for p, want in wanted]
    if tf32:
        i1, s1 = _image_buf(w1, K, c, c, "b")
        i2, s2 = _image_buf(w2, K, c, c, "b")
    else:
        i1 = i2 = None
        s1 = s2 = False
    _lib.call("scn_residual_unit_bwd", _ptr(gy), _ptr(r), _ptr(h), n, c, _ptr(fmap), K, _ptr(w1), _ptr(w2), _ptr(i1),
              _ptr(i2), int(s1 or s2), _ptr(gyr), _ptr(gh), _ptr(gx), _ptr(bufs[0]), _ptr(bufs[1]), _ptr(bufs[2]),
              _ptr(bufs[3]), int(accumulate), int(tf32), _stream())
    if accumulate:
        for p, want in wanted:
            if want:
                p._scn_grad_hook(p)
        bufs = [None, None, None, None]
    return gx, bufs


class ResidualUnitFunction(Function):
    """A chain of U residual units  y = x + conv2(relu(conv1(relu(x))))  over one neighbour map -- the `unit_stage` of the
    reference's sparse networks (module_factory.py:127-183 with relu_first, identity shortcut; :438-578 stacks num_units
    of them per level).  ReLU / residual add / TF32 rounding ride in the convolution epilogues (SCN_EPI_RELU|ROUND on conv1,
    SCN_EPI_ADD on conv2; backward: SCN_EPI_MASK|ROUND and SCN_EPI_MASK|ADD); each unit and direction is ONE C-ABI call and
    the whole chain is ONE autograd node, because the step is host bound (~45 us of Python per node and direction).
    apply(x, fmap, n, w1, b1, w2, b2[, w1', b1', w2', b2' ...])."""

    @staticmethod
    def forward(ctx, x, fmap, n, *params):
        x = _check(x)
        units = [params[i:i + 4] for i in range(0, len(params), 4)]
        c = units[0][0].shape[-1]
        tf32 = _state["precision"] == "tf32" and c <= 256
        saved = []
        for w1, b1, w2, b2 in units:
            r, h, x = _unit_forward(x, w1, b1, w2, b2, fmap, n, tf32)
            saved += [r, h]
        ctx.save_for_backward(*saved, *[u[0] for u in units], *[u[2] for u in units])
        ctx.cfg = (fmap, n, tf32, len(units))
        ctx.biases = [(u[1], u[3]) for u in units]
        return x

    @staticmethod
    def backward(ctx, gy):
        fmap, n, _, U = ctx.cfg
        t = ctx.saved_tensors
        rh, w1s, w2s = t[:2 * U], t[2 * U:3 * U], t[3 * U:]
        # precision is read when the backward runs (like ConvFunction.backward): tests isolate the backward kernels'
        # arithmetic from ReLU-mask flips by running an fp32 forward and a tf32 backward over the same saved activations
        tf32 = _state["precision"] == "tf32" and w1s[0].shape[-1] <= 256
        need = ctx.needs_input_grad
        g = _check(gy)
        grads = [None] * (4 * U)
        for u in range(U - 1, -1, -1):
            b1, b2 = ctx.biases[u]
            need_x = need[0] if u == 0 else True
            g, gp = _unit_backward(g, rh[2 * u], rh[2 * u + 1], w1s[u], b1, w2s[u], b2, fmap, n, tf32, need_x,
                                   need[3 + 4 * u:7 + 4 * u])
            grads[4 * u:4 * u + 4] = gp
        return (g, None, None, *grads)


class ReLUFunction(Function):
    @staticmethod
    def forward(ctx, x):
        x = _check(x)
        y = torch.empty_like(x)
        _lib.call("scn_relu_fwd", _ptr(x), _ptr(y), x.numel(), int(_state["precision"] == "tf32"), _stream())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, go):
        y, = ctx.saved_tensors
        go = _check(go)
        gi = torch.empty_like(go)
        _lib.call("scn_relu_bwd", _ptr(y), _ptr(go), _ptr(gi), go.numel(), int(_state["precision"] == "tf32"),
                  _stream())
        return _mark(gi)


class AddFunction(Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = _check(a), _check(b)
        if a.shape != b.shape:
            raise RuntimeError("AddTable: shape mismatch %s vs %s" % (tuple(a.shape), tuple(b.shape)))
        out = torch.empty_like(a)
        _lib.call("scn_add", _ptr(a), _ptr(b), _ptr(out), a.numel(), _stream())
        return out

    @staticmethod
    def backward(ctx, go):
        return go, go


class InputLayerFunction(Function):
    """scn.ioLayers.InputLayerFunction.apply(dimension, metadata, spatial_size, coords, features,
    batch_size, mode)  (custom_operations.py:74-82, roi_select_sparse.py:79-81,117-119)."""

    @staticmethod
    def forward(ctx, dimension, metadata, spatial_size, coords, input_features, batch_size, mode):
        f = _check(input_features)
        n = metadata.set_input(spatial_size, coords, batch_size, mode, f.device)
        if f.shape[0] != metadata.n_points:
            raise RuntimeError("InputLayer: %d coords but %d feature rows" % (metadata.n_points, f.shape[0]))
        ctx.md, ctx.mode, ctx.C = metadata, mode, f.shape[1]
        if mode == 0:
            return f.clone()
        out = torch.empty((n, f.shape[1]), dtype=torch.float32, device=f.device)
        _lib.call("scn_input_fwd", _ptr(f), f.stride(0), f.shape[1], _ptr(metadata.row_ptr), _ptr(metadata.row_pts),
                  n, mode, _ptr(out), _stream())
        return out

    @staticmethod
    def backward(ctx, go):
        md = ctx.md
        go = _check(go)
        if ctx.mode == 0:
            return None, None, None, None, go.clone(), None, None
        gf = torch.empty((md.n_points, ctx.C), dtype=torch.float32, device=go.device)
        _lib.call("scn_input_bwd", _ptr(go), ctx.C, _ptr(md.point_row), _ptr(md.row_ptr), _ptr(md.row_pts),
                  md.n_points, ctx.mode, _ptr(gf), _stream())
        return None, None, None, None, gf, None, None


class OutputLayerFunction(Function):
    """scn.ioLayers.OutputLayerFunction.apply(dimension, metadata, features)
    (custom_operations.py:7-10): every original point receives its voxel's row."""

    @staticmethod
    def forward(ctx, dimension, metadata, input_features):
        f = _check(input_features)
        ctx.md, ctx.n, ctx.C = metadata, f.shape[0], f.shape[1]
        out = torch.empty((metadata.n_points, f.shape[1]), dtype=torch.float32, device=f.device)
        _lib.call("scn_gather_rows", _ptr(f), f.stride(0), _ptr(metadata.point_row), metadata.n_points, f.shape[1],
                  _ptr(out), f.shape[1], _stream())
        return out

    @staticmethod
    def backward(ctx, go):
        md = ctx.md
        go = _check(go)
        row_ptr, row_pts = md.rule_csr()
        gi = torch.empty((ctx.n, ctx.C), dtype=torch.float32, device=go.device)
        _lib.call("scn_input_fwd", _ptr(go), go.stride(0), ctx.C, _ptr(row_ptr), _ptr(row_pts), ctx.n, 3, _ptr(gi),
                  _stream())
        return None, None, gi


class PoolFunction(Function):
    @staticmethod
    def forward(ctx, x, rules, n_out, is_max, inv_volume):
        x = _check(x)
        out = torch.empty((n_out, x.shape[1]), dtype=torch.float32, device=x.device)
        _lib.call("scn_pool_fwd", _ptr(x), x.shape[1], _ptr(rules.cmap), n_out, rules.K, int(is_max), inv_volume,
                  _ptr(out), _stream())
        ctx.rules, ctx.cfg = rules, (is_max, inv_volume)
        ctx.save_for_backward(x, out)
        return out

    @staticmethod
    def backward(ctx, go):
        x, out = ctx.saved_tensors
        go = _check(go)
        is_max, inv_volume = ctx.cfg
        gi = torch.empty_like(x)
        _lib.call("scn_pool_bwd", _ptr(x), _ptr(out), _ptr(go), x.shape[1], _ptr(ctx.rules.parent_row), x.shape[0],
                  int(is_max), inv_volume, _ptr(gi), _stream())
        return gi, None, None, None, None


class SparseToDenseFunction(Function):
    @staticmethod
    def forward(ctx, x, level, n_samples, size):
        x = _check(x)
        C = x.shape[1]
        out = torch.empty((n_samples, C, *size), dtype=torch.float32, device=x.device)
        _lib.call("scn_sparse_to_dense_fwd", _ptr(x), C, _ptr(level.tab_keys), _ptr(level.tab_vals), level.cap,
                  n_samples, size[0], size[1], size[2], _ptr(out), _stream())
        ctx.level, ctx.size, ctx.shape = level, size, x.shape
        return out

    @staticmethod
    def backward(ctx, go):
        go = _check(go)
        n, C = ctx.shape
        gi = torch.empty((n, C), dtype=torch.float32, device=go.device)
        _lib.call("scn_sparse_to_dense_bwd", _ptr(go), _ptr(ctx.level.keys), n, C, ctx.size[0], ctx.size[1],
                  ctx.size[2], _ptr(gi), _stream())
        return gi, None, None, None


class SegmentMeanFunction(Function):
    """Per-sample mean over batch-sorted rows (replaces split_batch + torch.mean,
    custom_operations.py:24-59, by one kernel)."""

    @staticmethod
    def forward(ctx, x, seg_ptr, n_seg):
        x = _check(x)
        out = torch.empty((n_seg, x.shape[1]), dtype=torch.float32, device=x.device)
        _lib.call("scn_segment_mean_fwd", _ptr(x), x.shape[1], _ptr(seg_ptr), n_seg, _ptr(out), _stream())
        ctx.seg, ctx.shape = (seg_ptr, n_seg), x.shape
        return out

    @staticmethod
    def backward(ctx, go):
        go = _check(go)
        seg_ptr, n_seg = ctx.seg
        gi = torch.zeros(ctx.shape, dtype=torch.float32, device=go.device)
        _lib.call("scn_segment_mean_bwd", _ptr(go), ctx.shape[1], _ptr(seg_ptr), n_seg, _ptr(gi), _stream())
        return gi, None, None


class BatchNormFunction(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, eps, momentum, leak, training):
        x = _check(x)
        n, C = x.shape
        s = _stream()
        if training:
            mean = torch.empty(C, dtype=torch.float32, device=x.device)
            var = torch.empty(C, dtype=torch.float32, device=x.device)
            _lib.call("scn_bn_stats", _ptr(x), n, C, _ptr(mean), _ptr(var), s)
            running_mean.mul_(momentum).add_(mean, alpha=1 - momentum)
            running_var.mul_(momentum).add_(var, alpha=1 - momentum)
        else:
            mean, var = running_mean, running_var
        y = torch.empty_like(x)
        w = weight.detach() if weight is not None else None
        b = bias.detach() if bias is not None else None
        _lib.call("scn_bn_apply", _ptr(x), n, C, _ptr(mean), _ptr(var), _ptr(w), _ptr(b), eps, leak, _ptr(y), s)
        ctx.save_for_backward(x, y, mean, var, weight)
        ctx.cfg = (eps, leak, training, weight is not None)
        return y

    @staticmethod
    def backward(ctx, go):
        x, y, mean, var, weight = ctx.saved_tensors
        eps, leak, training, affine = ctx.cfg
        go = _check(go)
        n, C = x.shape
        dx = torch.empty_like(x)
        tmp = torch.empty(2 * C, dtype=torch.float32, device=x.device)
        dg = torch.empty(C, dtype=torch.float32, device=x.device) if affine else None
        db = torch.empty(C, dtype=torch.float32, device=x.device) if affine else None
        _lib.call("scn_bn_bwd", _ptr(x), _ptr(y), _ptr(go), n, C, _ptr(mean), _ptr(var),
                  _ptr(weight.detach()) if affine else 0, eps, leak, int(training), _ptr(dx), _ptr(dg), _ptr(db),
                  _ptr(tmp), _stream())
        return dx, dg, db, None, None, None, None, None, None


class CrossEntropyFunction(Function):
    """nn.CrossEntropyLoss(weight, ignore_index, reduction='mean') of the segmentation head (ndsis/modules/loss.py:95-97)
    as two passes over the logits (torch's log_softmax + nll_loss pair costs 0.7 ms per step on 273k points)."""

    @staticmethod
    def forward(ctx, logits, labels, weight, ignore_index):
        logits = _check(logits)
        n, c = logits.shape
        if labels.dtype != torch.int64 or labels.shape != (n,) or not labels.is_cuda:
            raise RuntimeError("cross entropy expects int64 CUDA labels of shape [%d]" % n)
        if n == 0:
            raise RuntimeError("cross entropy needs at least one row")
        dev = logits.device
        lse = torch.empty(n, dtype=torch.float32, device=dev)
        rows = torch.empty((n, 2), dtype=torch.float32, device=dev)
        stats = torch.empty(2, dtype=torch.float32, device=dev)
        w = weight.contiguous() if weight is not None else None
        _lib.call("scn_cross_entropy_fwd", _ptr(logits), logits.stride(0), n, c, _ptr(labels), _ptr(w), int(ignore_index),
                  _ptr(lse), _ptr(rows), _ptr(stats), _stream())
        ctx.save_for_backward(logits, labels, lse, stats)
        ctx.cfg = (w, int(ignore_index))
        return stats[0] / stats[1]

    @staticmethod
    def backward(ctx, g):
        logits, labels, lse, stats = ctx.saved_tensors
        w, ignore_index = ctx.cfg
        n, c = logits.shape
        g = g.contiguous().float()
        dx = torch.empty((n, c), dtype=torch.float32, device=logits.device)
        _lib.call("scn_cross_entropy_bwd", _ptr(logits), logits.stride(0), n, c, _ptr(labels), _ptr(w), ignore_index,
                  _ptr(lse), _ptr(stats), _ptr(g), _ptr(dx), _stream())
        return dx, None, None, None


def cross_entropy(logits, labels, weight=None, ignore_index=-100):
    return CrossEntropyFunction.run(logits, labels, weight, ignore_index)


class UnPoolFunction(Function):
    """Transpose of sum-pooling: every fine row receives its coarse parent's row; backward sums the children."""

    @staticmethod
    def forward(ctx, x, parent_row, n_coarse):
        x = _check(x)
        n, C = parent_row.numel(), x.shape[1]
        out = torch.empty((n, C), dtype=torch.float32, device=x.device)
        _lib.call("scn_gather_rows", _ptr(x), x.stride(0), _ptr(parent_row), n, C, _ptr(out), C, _stream())
        ctx.parent_row, ctx.shape = parent_row, (n_coarse, C)
        return out

    @staticmethod
    def backward(ctx, go):
        go = _check(go)
        gi = torch.zeros(ctx.shape, dtype=torch.float32, device=go.device)
        _lib.call("scn_scatter_add_rows", _ptr(go), _ptr(ctx.parent_row), ctx.parent_row.numel(), ctx.shape[1], _ptr(gi),
                  _stream())
        return gi, None, None
