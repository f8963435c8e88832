"""GPU parity of the DEFAULT row order (Morton, csrc/sort.cu) against the oracle through the canonical sort: the sort itself
(bit-exact against numpy), the input rule, the rulebooks of ragged multi-sample batches incl. empty samples, and one
convolution layer forward + backward."""
import numpy as np
import pytest
import torch

import scn_oracle as O
from scn_oracle import rules as R
from tests.util import make_pair, random_scene, rel_err
from tests.test_gpu_baseline_size import batch_sorted, canon, canon_map

pytestmark = pytest.mark.gpu


def _morton_np(c):
    def spread(v):
        v = v.astype(np.uint64) & np.uint64(0xFFFF)
        for sh, m in ((32, 0x1f00000000ffff), (16, 0x1f0000ff0000ff), (8, 0x100f00f00f00f00f), (4, 0x10c30c30c30c30c3),
                      (2, 0x1249249249249249)):
            v = (v | (v << np.uint64(sh))) & np.uint64(m)
        return v
    return (c[:, 3].astype(np.uint64) << np.uint64(48)) | (spread(c[:, 0]) << np.uint64(2)) | \
           (spread(c[:, 1]) << np.uint64(1)) | spread(c[:, 2])


@pytest.mark.parametrize("P,size,B", [(1, 8, 1), (2049, 64, 1), (70001, 256, 3), (300000, 300, 255), (50000, 65535, 40000)])
def test_morton_order_is_the_stable_sort(cuda, P, size, B):
    from sparse_rcnn_b200 import _lib
    from sparse_rcnn_b200.scn.metadata import _stream
    rng = np.random.default_rng(P)
    c = np.stack([rng.integers(0, size, P), rng.integers(0, size, P), rng.integers(0, max(size // 2, 1), P),
                  np.sort(rng.integers(0, B, P))], 1).astype(np.int64)
    c[P // 3:P // 3 + 50] = c[P // 3]                                 # duplicates: stability matters
    keys = torch.from_numpy(R.pack_keys(c).view(np.int64)).to(cuda)
    ws = torch.empty(int(_lib.raw("scn_morton_order_ws_bytes")(P)), dtype=torch.uint8, device=cuda)
    perm = torch.empty(P, dtype=torch.int32, device=cuda)
    skeys = torch.empty(P, dtype=torch.int64, device=cuda)
    cb, bb = max(size - 1, 1).bit_length(), (B - 1).bit_length()
    _lib.call("scn_morton_order", keys.data_ptr(), P, cb, bb, perm.data_ptr(), skeys.data_ptr(), ws.data_ptr(), _stream())
    ref = np.argsort(_morton_np(c), kind="stable")
    assert np.array_equal(perm.cpu().numpy(), ref)
    assert np.array_equal(skeys.cpu().numpy(), R.pack_keys(c).view(np.int64)[ref])


@pytest.mark.parametrize("seed,mode", [(0, 4), (1, 3), (2, 1), (3, 2)])
def test_rulebooks_in_morton_order(cuda, seed, mode):
    from sparse_rcnn_b200 import scn
    assert scn.get_row_order() == "morton"
    coords, feats, size = random_scene(seed, size=(32, 32, 16), n_samples=4, density=0.05, dup=1.7)
    coords = coords[coords[:, 3] != 1]                                # an empty sample in the middle
    feats = feats[: len(coords)]
    to, tg = make_pair(scn, coords, feats, size, cuda, mode=mode, batch_size=6)
    assert tg.metadata.row_order == "morton" and tg.batch_size() == to.batch_size() == 6
    loc = tg.get_spatial_locations()
    assert batch_sorted(loc)
    mk = _morton_np(loc.numpy())
    assert (mk[:-1] < mk[1:]).all()                                   # strictly increasing Morton keys
    kg, og, ig = canon(loc)
    ko, oo, io = canon(to.get_spatial_locations())
    assert np.array_equal(kg, ko)
    assert np.array_equal(ig[tg.metadata.point_row.cpu().numpy()], io[to.metadata.point_row])
    assert rel_err(tg.features.cpu()[torch.from_numpy(og)], to.features[torch.from_numpy(oo)]) <= 1e-6
    # input-rule CSR: ascending point indices per row
    md = tg.metadata
    ptr, pts, pr = md.row_ptr.cpu().numpy(), md.row_pts.cpu().numpy(), md.point_row.cpu().numpy()
    for r in range(0, len(ptr) - 1, 5):
        seg = pts[ptr[r]:ptr[r + 1]]
        assert len(seg) and (np.diff(seg) > 0).all() and (pr[seg] == r).all()
    # pyramid: every level stays Morton / batch sorted and the maps agree canonically
    cur, prev = tuple(size.tolist()), None
    for lvl in range(3):
        go, gg = to.metadata.grids[cur], tg.metadata.levels[cur]
        kg, og, ig = canon(gg.locations())
        ko, oo, io = canon(go.coords)
        assert np.array_equal(kg, ko)
        mk = _morton_np(gg.locations().numpy())
        assert (mk[:-1] < mk[1:]).all()
        assert np.array_equal(canon_map(gg.subm_map(3).cpu().numpy(), og, ig),
                              canon_map(R.rules_to_map(to.metadata.subm_rules(cur, 3), go.n), oo, io))
        if prev is not None:
            ps, pog, pig, poo, pio = prev
            _, rules, parent, _ = to.metadata.conv_rules(ps, 2, 2)
            r = tg.metadata.strided_rules(ps, 2, 2)
            assert np.array_equal(canon_map(r.cmap.cpu().numpy(), og, pig), canon_map(R.rules_to_map(rules, go.n), oo, pio))
        ok, _, _, _ = to.metadata.conv_rules(cur, 2, 2)
        tg.metadata.strided_rules(cur, 2, 2)
        prev, cur = (cur, og, ig, oo, io), ok


def test_mode0_keeps_input_order(cuda):
    from sparse_rcnn_b200 import scn
    coords, feats, size = random_scene(3, dup=1.0)
    _, idx = np.unique(R.pack_keys(coords.numpy()), return_index=True)
    idx = np.sort(idx)
    coords, feats = coords[idx], feats[idx]
    to, tg = make_pair(scn, coords, feats, size, cuda, mode=0)
    assert tg.metadata.row_order == "first"
    assert torch.equal(tg.get_spatial_locations(), coords) and torch.equal(tg.features.cpu(), feats)


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("tf32", 2e-3)])
def test_conv_layer_in_morton_order(cuda, precision, tol):
    from sparse_rcnn_b200 import scn
    scn.set_precision(precision)
    try:
        coords, feats, size = random_scene(5, size=(40, 40, 16), n_samples=2, density=0.06, channels=32)
        to, tg = make_pair(scn, coords, feats, size, cuda)
        _, og, ig = canon(tg.get_spatial_locations())
        _, oo, io = canon(to.get_spatial_locations())
        lo = O.SubmanifoldConvolution(3, 32, 48, 3, True)
        lo.bias.data.normal_()
        lg = scn.SubmanifoldConvolution(3, 32, 48, 3, True)
        lg.load_state_dict(lo.state_dict())
        lg.to(cuda)
        xo = to.features.clone().requires_grad_(True)
        xg = tg.features.clone().requires_grad_(True)
        yo = lo(O.SparseConvNetTensor(xo, to.metadata, size)).features
        yg = lg(scn.SparseConvNetTensor(xg, tg.metadata, size)).features
        og_t, oo_t = torch.from_numpy(og), torch.from_numpy(oo)
        assert rel_err(yg.detach().cpu()[og_t], yo.detach()[oo_t]) <= tol
        gc = torch.randn(yo.shape, generator=torch.Generator().manual_seed(1))      # canonical output gradient
        yo.backward(gc[torch.from_numpy(io)])
        yg.backward(gc[torch.from_numpy(ig)].to(cuda))
        assert rel_err(xg.grad.cpu()[og_t], xo.grad[oo_t]) <= tol
        assert rel_err(lg.weight.grad, lo.weight.grad) <= 5 * tol and rel_err(lg.bias.grad, lo.bias.grad) <= 5 * tol
    finally:
        scn.set_precision("tf32")


# ------------------------------------------------------------------ tile books + the tile-local tensor-memory kernel
@pytest.fixture(scope="module")
def bench_level():
    from sparse_rcnn_b200 import scn
    from sparse_rcnn_b200.synthetic import make_batch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    dev = torch.device("cuda:0")
    coords, feats, size, bs, _ = make_batch(1, 0)
    md = scn.Metadata(3)
    scn.ioLayers.InputLayerFunction.apply(3, md, size, coords, feats.to(dev), bs, 4)
    return md, size


def test_tile_book_reproduces_the_map(cuda, bench_level):
    """lmap / rows / nloc / umask of every tile against the neighbour map they were built from (integer work: exact)."""
    from sparse_rcnn_b200 import _lib
    md, size = bench_level
    lvl = md.level(size)
    m = lvl.subm_map(3)
    assert lvl.coherent and m.data_ptr() in lvl._books
    book = lvl._books[m.data_ptr()].cpu().numpy()
    n = lvl.n
    nt = (n + 127) // 128
    a256 = lambda v: (v + 255) // 256 * 256
    BLOB = 64 + 27 * 128 * 2
    o0 = a256(nt * BLOB)
    o1 = o0 + a256(nt * 512 * 4)
    o2 = o1 + a256(nt * 4)
    assert int(_lib.raw("scn_tile_book_bytes")(n)) == o2 + a256(nt * 4)
    blobs = book[:nt * BLOB].reshape(nt, BLOB)
    lmap = np.ascontiguousarray(blobs[:, 64:]).view(np.uint16).reshape(nt, 27, 128)
    rows = book[o0:o0 + nt * 512 * 4].view(np.int32).reshape(nt, 512)
    nloc = book[o1:o1 + nt * 4].view(np.int32)
    useq = book[o2:o2 + nt * 4].view(np.uint32)
    mp = np.full((27, nt * 128), -1, np.int32)
    mp[:, :n] = m.cpu().numpy()
    mp = mp.reshape(27, nt, 128).transpose(1, 0, 2)                   # [tile, offset, row]
    assert nloc.max() <= 512 and nloc.min() >= 1
    active = lmap != 0xFFFF
    assert np.array_equal(active, mp >= 0)
    assert not (lmap == 0xFFFE).any()                                 # Morton order: every halo set fits the book
    t_idx = np.broadcast_to(np.arange(nt)[:, None, None], lmap.shape)
    assert np.array_equal(rows[t_idx[active], lmap[active]], mp[active])
    assert (lmap[active] < nloc[t_idx[active]]).all()
    for t in (0, nt // 2, nt - 1):                                    # lists hold DISTINCT rows in ascending order
        assert len(np.unique(rows[t, :nloc[t]])) == nloc[t] == len(np.unique(mp[t][mp[t] >= 0]))
        assert (np.diff(rows[t, :nloc[t]]) > 0).all()
    # processing order: centre offset first, then the other active offsets ascending
    for t in (0, 1, nt // 3, nt - 1):
        act = np.nonzero(active[t].any(1))[0].tolist()
        want = [13] + [o for o in act if o != 13]
        assert blobs[t, 0] == len(want) and blobs[t, 1:1 + len(want)].tolist() == want
        # the same order as a bit sequence: bit 0 = centre, bits 1..13 = offsets 0..12, bits 14..26 = offsets 14..26
        bits = [b for b in range(27) if (int(useq[t]) >> b) & 1]
        assert [13 if b == 0 else (b - 1 if b <= 13 else b) for b in bits] == want
    print("halo rows per tile: mean %.0f max %d; empty units %.1f %%" % (
        nloc.mean(), nloc.max(), 100.0 * (1 - blobs[:, 0].astype(float).sum() / (27.0 * nt))))


@pytest.mark.parametrize("C", [32, 16, 48, 64, -48, -64])
def test_tile_local_kernel_equals_rule_kernel(cuda, bench_level, monkeypatch, C):
    if C < 0:      # streamed weights shared by the two CTAs of a cluster (multicast); opt-in variant
        monkeypatch.setenv("SCN_CONV_TS_CLUSTER", "2")
        C = -C
    else:
        monkeypatch.setenv("SCN_CONV_TS_CLUSTER", "1")
    """conv_ts.cu (halo set in shared memory, A operand in tensor memory) against conv_tc.cu (cp.async gather per offset) on
    the bench scene: same TF32 products accumulated in the same offset order; forward with every epilogue, input gradient."""
    from sparse_rcnn_b200 import networks, scn
    scn.set_precision("tf32")
    md, size = bench_level
    lvl = md.level(size)
    n = lvl.n
    torch.manual_seed(C)
    conv = scn.SubmanifoldConvolution(3, C, C, 3, True).to(cuda)
    conv.bias.data.normal_()
    x0 = torch.randn(n, C, device=cuda)
    go = torch.randn(n, C, device=cuda)
    unit = networks.residual_unit(scn, C, C).to(cuda)

    def run():
        x = x0.clone().requires_grad_(True)
        y = conv(scn.SparseConvNetTensor(x, md, size)).features
        y.backward(go)
        x2 = x0.clone().requires_grad_(True)
        unit.zero_grad()
        z = unit(scn.SparseConvNetTensor(x2, md, size)).features      # ReLU|ROUND, ADD epilogues; backward: MASK|ROUND, MASK|ADD
        z.backward(go)
        return [y.detach(), x.grad, z.detach(), x2.grad]
    monkeypatch.setenv("SCN_CONV_TS", "0")
    monkeypatch.setenv("SCN_CONV_TAILSPLIT", "0")
    ref = run()
    monkeypatch.setenv("SCN_CONV_TS", "1")
    from sparse_rcnn_b200 import _lib
    before = int(_lib.raw("scn_conv_ts_launch_count")())
    got = run()
    assert int(_lib.raw("scn_conv_ts_launch_count")()) == before + 6      # conv fwd + dx, unit: 2 fwd + 2 bwd
    got2 = run()
    # Same TF32 products; the tile-local kernel adds the centre offset first, so fp32 sums differ in their last bit.  Through
    # a residual unit that bit can flip the TF32 rounding (2^-11) of the inner activation: 1e-4 there, 1e-6 for a single layer.
    for a, b, c, name in zip(got, ref, got2, ("conv fwd", "conv dx", "unit fwd", "unit dx")):
        assert rel_err(a, b) <= {"unit fwd": 1e-4, "unit dx": 2e-3}.get(name, 1e-6), (name, rel_err(a, b))
        if name == "unit dx":      # ... and where that bit flips a ReLU mask, one gradient element toggles: rare
            assert ((a - b).abs() > 1e-5 * b.abs().max()).float().mean() < 2e-3
        assert torch.equal(a, c), name                                # deterministic


@pytest.mark.parametrize("C", [32, 48])
def test_tile_local_weight_gradient(cuda, bench_level, monkeypatch, C):
    """conv_wgrad_ts.cu (halo set in shared memory, A = [(offset, channel)] x [row] in tensor memory, per-CTA partial sums reduced
    in CTA order) against conv_wgrad_tc.cu (cp.async gather per offset, floating-point atomics) on the bench scene: the same
    TF32 products, another summation order -> 1e-5; and against itself: bit-reproducible."""
    from sparse_rcnn_b200 import _lib, scn
    from sparse_rcnn_b200.scn import functions as Fn
    from sparse_rcnn_b200.scn.metadata import _stream
    scn.set_precision("tf32")
    md, size = bench_level
    lvl = md.level(size)
    n, K = lvl.n, 27
    m = lvl.subm_map(3)
    torch.manual_seed(C)
    x = Fn.tf32_exact(torch.randn(n, C, device=cuda))
    go = Fn.tf32_exact(torch.randn(n, C, device=cuda))
    P = lambda t: t.data_ptr()

    def run(ts, fill=0.0):
        monkeypatch.setenv("SCN_WGRAD_TS", ("48" if C == 48 else "1") if ts else "0")      # C = 48 is opt-in (slower)
        gw = torch.full((K, C, C), fill, device=cuda)
        gb = torch.full((C,), fill, device=cuda)
        _lib.call("scn_conv_bwd_weight", P(x), C, C, P(m), n, K, P(go), C, C, P(gw), P(gb), 1, _stream())
        torch.cuda.synchronize()
        return gw, gb
    ref_w, ref_b = run(False)
    before = int(_lib.raw("scn_conv_wgrad_ts_launch_count")())
    w1, b1 = run(True)
    assert int(_lib.raw("scn_conv_wgrad_ts_launch_count")()) == before + 1
    w2, b2 = run(True)
    assert rel_err(w1, ref_w) <= 1e-5, rel_err(w1, ref_w)
    assert rel_err(b1, ref_b) <= 1e-5, rel_err(b1, ref_b)
    assert torch.equal(w1, w2) and torch.equal(b1, b2)              # deterministic
    w3, b3 = run(True, fill=1.0)                                     # gradients are ADDED to what the buffers hold
    assert rel_err(w3 - 1.0, w1) <= 1e-5 and rel_err(b3 - 1.0, b1) <= 1e-5
    # independent check of a few offsets in float64 (exact products of TF32-representable operands)
    mm = m.long()
    for o in (0, 13, 26):
        act = mm[o] >= 0
        want = x.double()[mm[o][act]].t() @ go.double()[act]
        assert rel_err(w1[o], want) <= 2e-5, (o, rel_err(w1[o], want))
