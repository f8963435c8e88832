"""Training-step plumbing on the GPU: one-launch weight packing, gradients accumulated straight into the flat buckets
(parallel.GradientBuckets) must equal the plain autograd path (ndsis/training/training.py:428-460 semantics)."""
import pytest
import torch

from tests.util import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


def test_pack_weights_multi_equals_single(cuda):
    from sparse_rcnn_b200 import _lib
    from sparse_rcnn_b200.scn.metadata import _stream
    torch.manual_seed(0)
    nbytes = _lib.raw("scn_conv_weight_image_bytes")
    shapes = [(27, 32, 32, 0, 0), (27, 48, 32, 1, 1), (8, 6, 16, 0, 0), (1, 112, 20, 1, 0), (27, 112, 112, 1, 1)]
    rows, single, multi, keep = [], [], [], []
    for K, cin, cout, tr, rev in shapes:
        a, b = (cout, cin) if tr else (cin, cout)        # storage layout of w: [K, Cin_w, Cout_w]
        w = torch.randn(K, a, b, device=cuda)
        i1 = torch.zeros(int(nbytes(K, cin, cout)), dtype=torch.uint8, device=cuda)
        i2 = torch.zeros_like(i1)
        _lib.call("scn_conv_pack_weights", w.data_ptr(), K, cin, cout, tr, rev, i1.data_ptr(), _stream())
        rows.append((w.data_ptr(), i2.data_ptr(), K, cin, cout, tr, rev))
        single.append(i1), multi.append(i2), keep.append(w)
    table = torch.tensor(rows, dtype=torch.int64).to(cuda)
    _lib.call("scn_conv_pack_weights_multi", table.data_ptr(), len(rows), _stream())
    for a, b in zip(single, multi):
        assert torch.equal(a, b)


def test_col_sum_add(cuda):
    from sparse_rcnn_b200 import _lib
    from sparse_rcnn_b200.scn.metadata import _stream
    torch.manual_seed(0)
    x = torch.randn(5000, 48, device=cuda)
    out = torch.full((48,), 3.0, device=cuda)
    _lib.call("scn_col_sum_add", x.data_ptr(), 48, 5000, 48, out.data_ptr(), _stream())
    assert rel_err(out, 3.0 + x.double().sum(0)) < 1e-5


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_direct_gradients_equal_autograd_accumulation(cuda, precision):
    """Same weights, same scene: gradients written straight into the buckets == gradients accumulated by autograd."""
    from sparse_rcnn_b200 import pipeline, scn
    import bench
    scn.set_precision(precision)
    try:
        data, labels = bench.make_inputs(0, scene_kw=bench.CPU_SAMPLE)
        grads = []
        for direct in (True, False):
            tr = pipeline.BackboneTrainer(cuda, seed=3)
            if not direct:
                for p in tr.parameters():
                    p._scn_grad_hook = None
            tr.optimizer = torch.optim.SGD(tr.parameters(), lr=0.0)        # keep the weights: compare two steps' grads
            for _ in range(2):                                             # second step exercises pack_all / stale images
                tr.step(data, labels)
            grads.append([p.grad.detach().clone() for p in tr.parameters()])
        if precision == "fp32":
            assert max(rel_err(a, b) for a, b in zip(*grads)) < 2e-5
        else:
            # tf32 forward is not bit-reproducible run to run at the small levels (split-offset partial sums are added with
            # fp32 atomics), which flips a few ReLU masks: two REGULAR runs differ by several percent in single deep-level
            # tensors, so compare the whole gradient in L2
            fa, fb = (torch.cat([g.flatten().double() for g in gs]) for gs in grads)
            assert float((fa - fb).norm() / fb.norm()) < 2e-2
        assert all(torch.isfinite(g).all() for g in grads[0])
    finally:
        scn.set_precision("tf32")


@pytest.mark.parametrize("weighted", [False, True])
def test_cross_entropy_matches_torch(cuda, weighted):
    """fp32 op: tolerance 1e-5 relative against nn.functional.cross_entropy (ndsis/modules/loss.py:95-97 semantics)."""
    from sparse_rcnn_b200.scn import functions as Fn
    torch.manual_seed(1)
    n, c = 10007, 20
    x = (torch.randn(n, c, device=cuda) * 3).requires_grad_()
    y = torch.randint(0, c, (n,), device=cuda)
    y[::17] = -100                                                  # ignored rows
    w = (torch.rand(c, device=cuda) + 0.5) if weighted else None
    ref = torch.nn.functional.cross_entropy(x, y, weight=w, ignore_index=-100)
    gref, = torch.autograd.grad(ref * 1.7, x)
    x2 = x.detach().clone().requires_grad_()
    out = Fn.cross_entropy(x2, y, w, -100)
    (out * 1.7).backward()
    assert abs(float(out) - float(ref)) <= 1e-5 * abs(float(ref))
    assert rel_err(x2.grad, gref) < 1e-5
    assert float(x2.grad[::17].abs().max()) == 0.0


def test_prefetched_geometry_equals_inline(cuda):
    """Rulebooks built one step ahead on the side stream (scn.GeometryPrefetcher) give the same losses as building them
    inside the step; a mismatching input is refused."""
    from sparse_rcnn_b200 import pipeline, scn
    import bench
    scn.set_precision("fp32")
    try:
        batches = [bench.make_inputs(s, scene_kw=bench.CPU_SAMPLE) for s in (0, 1, 2)]
        losses = []
        batches = batches + batches      # six steps: the prefetcher's arenas are recycled
        for prefetch in (False, True, "inline", "thread"):      # build_late: training thread / worker thread, recycled arenas
            tr = pipeline.BackboneTrainer(cuda, seed=5)
            tr.build_late = prefetch if isinstance(prefetch, str) else False
            out = []
            for i, (d, l) in enumerate(batches):
                nxt = batches[i + 1] if prefetch and i + 1 < len(batches) else None
                if isinstance(prefetch, str):
                    out.append(float(tr.step(d, l, next_batch=nxt)))
                    assert (tr.prefetcher is not None and len(tr.prefetcher.pending) == 1) or nxt is None
                    assert nxt is None or tr.prefetcher.last_arena_bytes > 0 or prefetch == "thread"
                else:
                    out.append(float(tr.step(d, l, next_data=nxt[0] if nxt else None)))
            losses.append(out)
            if prefetch:
                assert tr.prefetcher is not None and not tr.prefetcher.pending
                tr.prefetcher.shutdown()
        for other in losses[1:]:
            assert max(abs(a - b) / abs(a) for a, b in zip(losses[0], other)) < 1e-5, losses
        assert len(tr.prefetcher._free) <= 3 and tr.prefetcher.last_arena_bytes > 0      # arenas cycle, nothing leaks
        md = scn.Metadata(3)
        md._prebuilt_for = batches[0][0][0]
        with pytest.raises(RuntimeError):
            md.set_input(batches[1][0][2], batches[1][0][0], 1, 4, cuda)
    finally:
        scn.set_precision("tf32")


def test_staged_uploads_equal_inline(cuda):
    """Host->device copies issued one step ahead on the copy stream (BackboneTrainer.stage / next_batch) change nothing."""
    from sparse_rcnn_b200 import pipeline, scn
    from sparse_rcnn_b200.scn import metadata
    import bench
    scn.set_precision("fp32")
    try:
        host = [bench.make_inputs(s, scene_kw=bench.CPU_SAMPLE) for s in (0, 1, 2)]
        batches = [((d[0].pin_memory(), d[1].pin_memory(), d[2], d[3], d[4]), l.pin_memory()) for d, l in host]
        losses = []
        for stage in (False, True):
            tr = pipeline.BackboneTrainer(cuda, seed=5)
            tr.stage_uploads = True
            out = []
            for i, (d, l) in enumerate(batches):
                nb = batches[i + 1] if stage and i + 1 < len(batches) else None
                out.append(float(tr.step(d, l, next_batch=nb)))
            losses.append(out)
        assert not metadata._staged                                  # every staged tensor was picked up
        assert max(abs(a - b) / abs(a) for a, b in zip(*losses)) < 1e-5, losses      # sums contain fp32 atomics
    finally:
        scn.set_precision("tf32")


def test_staged_copy_that_is_never_taken_cannot_go_stale(cuda):
    """The staging buffers are recycled (two alternating arenas): a tensor that was staged but never picked up -- a warm-up
    loop that ends one batch early, as bench.py's does -- must not be handed out later, after its buffer has been rewritten
    (regression: the geometry worker packed garbage coordinates four steps later)."""
    from sparse_rcnn_b200.scn import metadata as M
    g = torch.Generator().manual_seed(0)
    host = [torch.randint(0, 1000, (5000 + 100 * i, 4), generator=g).pin_memory() for i in range(5)]      # three recycled buffers
    M.stage_to_device([host[0]], cuda)                  # staged, never taken
    for i in (1, 2, 3, 4, 1, 2):                        # its buffer is rewritten twice meanwhile
        M.stage_to_device([host[i]], cuda)
        assert torch.equal(M.take_staged(host[i]).cpu(), host[i])
    assert M.take_staged(host[0]) is None               # the stale entry is gone: the caller uploads it again
    M.stage_to_device([host[0]], cuda)
    assert torch.equal(M.take_staged(host[0]).cpu(), host[0])
    assert not M._staged


def test_geometry_built_ahead_equals_inline(cuda):
    """next_batch: rulebooks built between the previous step's forward and backward (InputStage.build_ahead) give the same
    losses as building them at the head of the step."""
    from sparse_rcnn_b200 import pipeline, scn
    import bench
    scn.set_precision("fp32")
    try:
        batches = [bench.make_inputs(s, scene_kw=bench.CPU_SAMPLE) for s in (0, 1, 2)]
        losses = []
        for ahead in (False, True):
            tr = pipeline.BackboneTrainer(cuda, seed=5)
            tr.build_ahead = ahead
            out = []
            for i, (d, l) in enumerate(batches):
                nb = batches[i + 1] if i + 1 < len(batches) else None
                out.append(float(tr.step(d, l, next_batch=nb)))
                assert len(tr.backbone.input_stage.ready) == (1 if ahead and nb is not None else 0)
            losses.append(out)
        assert max(abs(a - b) / abs(a) for a, b in zip(*losses)) < 1e-5, losses      # sums contain fp32 atomics
    finally:
        scn.set_precision("tf32")


def test_tf32_weight_images_follow_the_optimizer(cuda):
    """the optimizer updates the FLAT parameter buffers, so the packed TF32 weight images must be keyed on
    the flat buffers' version.  After two Adam steps with lr > 0 every cached image must equal a fresh pack of the current
    weights, and the TF32 losses must track the fp32 ones (they did not move off the initial weights before the fix)."""
    from sparse_rcnn_b200 import _lib, pipeline, scn
    from sparse_rcnn_b200.scn import functions as Fn
    from sparse_rcnn_b200.scn.metadata import _stream
    import bench
    data, labels = bench.make_inputs(0, scene_kw=bench.CPU_SAMPLE)
    losses = {}
    try:
        for precision in ("fp32", "tf32"):
            scn.set_precision(precision)
            tr = pipeline.BackboneTrainer(cuda, seed=3, lr=3e-3)
            losses[precision] = [float(tr.step(data, labels)) for _ in range(4)]
        Fn.pack_all(tr._weights)                                       # what the next step starts with
        checked = 0
        for w in tr._weights:
            for (transpose, reverse), hit in getattr(w, "_scn_img", {}).items():
                K, cin, cout = hit[2]
                fresh = torch.zeros_like(hit[0])
                _lib.call("scn_conv_pack_weights", w.data_ptr(), K, cin, cout, transpose, reverse, fresh.data_ptr(), _stream())
                assert torch.equal(fresh, hit[0]), (tuple(w.shape), transpose, reverse)
                checked += 1
        assert checked > 100
        a, b = losses["fp32"], losses["tf32"]
        assert a[-1] < 0.9 * a[0]                                      # the scene is being fitted
        assert max(abs(x - y) / abs(x) for x, y in zip(a, b)) < 2e-2, (a, b)
    finally:
        scn.set_precision("tf32")


def test_empty_input_does_not_leave_an_unpacked_image_marked_fresh(cuda):
    """an empty crop / batch runs no kernel; the packed image it would have used must not stay marked as
    packed for the current weight version.  Empty call first, then a real one, against untouched copies of the weights."""
    from sparse_rcnn_b200 import scn
    from sparse_rcnn_b200.scn.functions import ConvFunction, ResidualUnitFunction
    from tests.util import make_pair, random_scene
    scn.set_precision("tf32")
    torch.manual_seed(0)
    coords, feats, size = random_scene(3, channels=32, size=(28, 24, 16))
    _, tg = make_pair(scn, coords, feats, size, cuda)
    lvl = tg.metadata.level(size)
    m, n = lvl.subm_map(3), lvl.n
    ps = [torch.randn(27, 1, 32, 32, device=cuda) * 0.05, torch.randn(32, device=cuda),
          torch.randn(27, 1, 32, 32, device=cuda) * 0.05, torch.randn(32, device=cuda)]
    ps = [p.requires_grad_(True) for p in ps]
    x = tg.features.detach()
    go = torch.randn(n, 32, device=cuda)

    def run(params, xin):
        xin = xin.clone().requires_grad_(True)
        y = ResidualUnitFunction.apply(xin, m, xin.shape[0], *params)
        y.backward(go[: xin.shape[0]])
        return y.detach(), xin.grad

    run(ps, x[:0])                                                     # empty first: forward and backward
    y, gx = run(ps, x)
    y_ref, gx_ref = run([p.detach().clone().requires_grad_(True) for p in ps], x)
    assert torch.equal(y, y_ref) and torch.equal(gx, gx_ref)
    # plain convolution layer: an empty backward first (forward returns before touching the image)
    w, b = ps[0], ps[1]

    def conv(wt, bt, xin, n_out):
        xin = xin.clone().requires_grad_(True)
        y = ConvFunction.apply(xin, wt, bt, m, m, n_out, 1)
        y.backward(go[:n_out])
        return xin.grad
    w2, b2 = (torch.randn_like(w) * 0.05).requires_grad_(True), b.detach().clone().requires_grad_(True)
    conv(w2, b2, x[:0], 0)
    g1 = conv(w2, b2, x, n)
    g2 = conv(w2.detach().clone().requires_grad_(True), b2.detach().clone().requires_grad_(True), x, n)
    assert torch.equal(g1, g2)
