"""Native sparse U-Net executor (csrc/unet_exec.cu, sparse_rcnn_b200/executor.py) against the module-by-module path it
replaces (FeatureExtractor.forward, model.py:414-446): the same kernels in the same order, so activations are compared bit for
bit; parameter gradients go through the weight-gradient kernel's floating-point atomics and are held to 1e-5."""
import pytest
import torch

from sparse_rcnn_b200 import networks
from sparse_rcnn_b200.synthetic import make_batch
from tests.util import rel_err

pytestmark = pytest.mark.gpu


def _batch(n_scenes=2, seed=3):
    return make_batch(n_scenes, seed, spatial_size=(64, 64, 32), room=(44, 44, 22), room_offset=(8, 8, 2), n_furniture=4)


def _run(net, seg, data, cuda, extra_seed, with_executor):
    from sparse_rcnn_b200 import executor
    executor.ENABLED["unet"] = with_executor
    try:
        coords, feats, size, bs, splits = data
        f = feats.to(cuda).requires_grad_(True)
        net.zero_grad(), seg.zero_grad()
        out = net((coords, f, size, bs, splits))
        logits = seg(out[5])
        gen = torch.Generator().manual_seed(5)
        loss = (logits * torch.randn(logits.shape, generator=gen).to(cuda)).sum()
        if extra_seed:      # gradients entering at an encoder output and at an inner decoder output as well
            loss = loss + (out[4][2].features * torch.randn(out[4][2].features.shape, generator=gen).to(cuda)).sum()
            loss = loss + (out[5][1].features * torch.randn(out[5][1].features.shape, generator=gen).to(cuda)).sum()
        loss.backward()
        acts = [t.features.detach().clone() for t in out[4] + out[5]]
        grads = {n: p.grad.detach().clone() for n, p in list(net.named_parameters()) + list(seg.named_parameters())}
        return acts, logits.detach().clone(), f.grad.detach().clone(), grads, [tuple(t.spatial_size.tolist()) for t in out[4] + out[5]]
    finally:
        executor.ENABLED["unet"] = True


@pytest.mark.parametrize("precision", ["tf32", "fp32"])
@pytest.mark.parametrize("extra_seed", [False, True])
def test_executor_equals_module_graph(cuda, precision, extra_seed):
    from sparse_rcnn_b200 import _lib, scn
    scn.set_precision(precision)
    torch.manual_seed(0)
    net, seg = networks.FeatureExtractor(scn).to(cuda), networks.SegmentationNetwork(scn).to(cuda)
    assert net._executor() is not None
    data = _batch()
    _run(net, seg, data, cuda, extra_seed, True)           # first touch: packs the weight images
    l0 = int(_lib.raw("scn_launch_count")())
    a1, s1, gx1, g1, sz1 = _run(net, seg, data, cuda, extra_seed, True)
    l1 = int(_lib.raw("scn_launch_count")())
    a0, s0, gx0, g0, sz0 = _run(net, seg, data, cuda, extra_seed, False)
    l2 = int(_lib.raw("scn_launch_count")())
    assert sz1 == sz0 and len(a1) == len(a0) == 11
    for x, y in zip(a1, a0):
        assert x.shape == y.shape and torch.equal(x, y)     # same kernels, same order: identical bits
    assert torch.equal(s1, s0)
    # Gradients: one seed -> every sum has two terms (commutative), only the weight-gradient atomics differ.  With extra seeds
    # an encoder output sums THREE gradients, in another order than the autograd engine: last-bit differences, which the
    # TF32-rounding / ReLU-mask epilogues of the transposed convolutions amplify to the TF32 tolerance
    gtol = (2e-3 if precision == "tf32" else 1e-5) if extra_seed else 2e-5
    assert rel_err(gx1, gx0) <= (gtol if extra_seed else 1e-6), rel_err(gx1, gx0)
    assert set(g1) == set(g0)
    for n in g0:
        assert rel_err(g1[n], g0[n]) <= gtol, (n, rel_err(g1[n], g0[n]))
    print("kernel launches: executor %d, module graph %d" % (l1 - l0, l2 - l1))
    # (the executor's counted launches include the skip-gradient adds and the column copies of JoinTable and its backward --
    # own kernels since round 2b -- that the module graph does with uncounted torch kernels: 2 + 3 per decoder level)
    assert l1 - l0 <= l2 - l1 + (2 + 3) * 5


def test_executor_under_no_grad_and_threads(cuda):
    """Inference: no autograd graph, forward images only; two host threads share one compiled program."""
    import threading
    from sparse_rcnn_b200 import executor, scn
    scn.set_precision("tf32")
    torch.manual_seed(0)
    net = networks.FeatureExtractor(scn).to(cuda).eval()
    datas = [_batch(1, 3), _batch(2, 4)]
    want = []
    executor.ENABLED["unet"] = False
    with torch.no_grad():
        for d in datas:
            out = net((d[0], d[1].to(cuda), d[2], d[3], d[4]))
            want.append([t.features.clone() for t in out[4] + out[5]])
    executor.ENABLED["unet"] = True
    got = [None, None]

    def work(i):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s), torch.no_grad():
            for _ in range(3):
                d = datas[i]
                out = net((d[0], d[1].to(cuda), d[2], d[3], d[4]))
                got[i] = [t.features.clone() for t in out[4] + out[5]]
        s.synchronize()
    with torch.no_grad():
        d = datas[0]
        net((d[0], d[1].to(cuda), d[2], d[3], d[4]))        # weight images packed on the caller's stream first
    torch.cuda.synchronize()
    ts = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in ts], [t.join() for t in ts]
    for g, w in zip(got, want):
        for x, y in zip(g, w):
            assert torch.equal(x, y)


@pytest.mark.parametrize("precision", ["tf32", "fp32"])
def test_trainer_steps_with_and_without_executor(cuda, precision):
    """BackboneTrainer (gradient buckets, flat-parameter fused Adam): two steps with the executor == two steps without."""
    import bench
    from sparse_rcnn_b200 import executor, pipeline, scn
    scn.set_precision(precision)
    data, labels = bench.make_inputs(0, scene_kw=bench.CPU_SAMPLE)
    res = []
    for on in (True, False):
        executor.ENABLED["unet"] = on
        try:
            tr = pipeline.BackboneTrainer(cuda, seed=3)
            losses = [float(tr.step(data, labels)) for _ in range(3)]
            res.append((losses, [p.detach().clone() for p in tr.parameters()]))
        finally:
            executor.ENABLED["unet"] = True
    (l1, p1), (l0, p0) = res
    assert abs(l1[0] - l0[0]) <= 1e-6 * abs(l0[0])            # same forward
    for a, b in zip(l1, l0):
        assert abs(a - b) <= (2e-3 if precision == "tf32" else 1e-4) * abs(b), (l1, l0)
    assert l1[2] < l1[0]                                        # it trains


@pytest.mark.parametrize("extra_seed", [False, True])
def test_second_outputs_equal_separate_passes(cuda, monkeypatch, extra_seed):
    """scn_conv_fwd_tf32_dual inside the executor (the producing convolution also writes relu / round of its result for the
    next gather) against SCN_EXEC_DUAL=0 (one elementwise kernel per consumer) on the BASELINE scene: every epilogue flavour
    runs (tile-local kernel on levels 0-1, whole-tile persistent grid, tail split, cluster split on the small levels).
    Same values from the same fp32 results: activations and the input gradient are identical bits."""
    from sparse_rcnn_b200 import _lib, scn
    scn.set_precision("tf32")
    torch.manual_seed(0)
    net, seg = networks.FeatureExtractor(scn).to(cuda), networks.SegmentationNetwork(scn).to(cuda)
    data = make_batch(1, 0)
    monkeypatch.setenv("SCN_CONV_TAILSPLIT", "0")      # its fp32 atomics reorder sums between any two runs
    monkeypatch.setenv("SCN_EXEC_DUAL", "1")
    _run(net, seg, data, cuda, extra_seed, True)
    l0 = int(_lib.raw("scn_launch_count")())
    a1, s1, gx1, g1, _ = _run(net, seg, data, cuda, extra_seed, True)
    l1 = int(_lib.raw("scn_launch_count")())
    monkeypatch.setenv("SCN_EXEC_DUAL", "0")
    a0, s0, gx0, g0, _ = _run(net, seg, data, cuda, extra_seed, True)
    l2 = int(_lib.raw("scn_launch_count")())
    for x, y in zip(a1, a0):
        assert torch.equal(x, y)
    assert torch.equal(s1, s0) and torch.equal(gx1, gx0)
    for n in g0:      # k_conv_wgrad_tc ends in fp32 atomics
        assert rel_err(g1[n], g0[n]) <= 2e-5, (n, rel_err(g1[n], g0[n]))
    print("kernel launches: second outputs %d, separate passes %d" % (l1 - l0, l2 - l1))
    assert (l2 - l1) - (l1 - l0) >= 50
