"""GPU parity AT THE CONFIGURATION THE BENCH TIMES ("next round" item 1a): the BASELINE scene
(256x256x128 grid, ~167k active voxels, ~273k points) against the oracle.

* every level of the pyramid: active sets, submanifold 3^3 neighbour maps, strided cmap / dmap -- BIT-EXACT;
* one SubM 32->32 layer at level 0 and one 48->48 layer at level 1 in the large-grid regime (persistent CTAs, with and
  without tail split / row skipping, and the tile-local tensor-memory kernel) -- forward, input gradient, weight gradient;
* the full FeatureExtractor + segmentation head forward and backward.

SURVEY 8c: rows may be numbered in any batch-sorted order, parity is "after canonical sort".  Everything here is therefore
compared through the canonical (b, x, y, z) permutation of both sides, so the same tests hold for first-appearance rows
(`scn.set_row_order('first')`, SparseConvNet's level-0 numbering) and for Morton rows (`'morton'`, the default)."""
import numpy as np
import pytest
import torch

import scn_oracle as O
from scn_oracle import rules as R
from sparse_rcnn_b200 import networks
from sparse_rcnn_b200.synthetic import make_batch
from tests.util import rel_err

pytestmark = pytest.mark.gpu
TOL = {"fp32": 1e-5, "tf32": 2e-3}


def canon(loc):
    """locations [n,4] (x,y,z,b) -> (sorted keys, order: canonical index -> row, inv: row -> canonical index)."""
    keys = R.pack_keys(loc.numpy() if isinstance(loc, torch.Tensor) else loc)
    order = np.argsort(keys, kind="stable")
    inv = np.empty(len(order), np.int64)
    inv[order] = np.arange(len(order))
    return keys[order], order, inv


def canon_map(m, col_order, val_inv):
    """map [K, n_cols] of rows (-1 = inactive) -> the same map with columns and values expressed canonically."""
    mm = np.asarray(m)[:, col_order].astype(np.int64)
    return np.where(mm >= 0, val_inv[np.maximum(mm, 0)], -1)


def batch_sorted(loc):
    return bool((loc[:-1, 3] <= loc[1:, 3]).all())


@pytest.fixture(scope="module")
def scene():
    return make_batch(1, 0)


@pytest.fixture(scope="module")
def pair(scene):
    """Oracle and CUDA metadata of the bench scene with all six levels built."""
    from sparse_rcnn_b200 import scn
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    dev = torch.device("cuda:0")
    coords, feats, size, bs, _ = scene
    mo = O.Metadata(3)
    fo = O.ioLayers.InputLayerFunction.apply(3, mo, size, coords, feats, bs, 4)
    mg = scn.Metadata(3)
    fg = scn.ioLayers.InputLayerFunction.apply(3, mg, size, coords, feats.to(dev), bs, 4)
    sizes = [tuple(size.tolist())]
    for _ in range(5):
        ok, _, _, _ = mo.conv_rules(sizes[-1], 2, 2)
        r = mg.strided_rules(sizes[-1], 2, 2)
        assert r.out_key == ok
        sizes.append(ok)
    return mo, fo, mg, fg, sizes


def test_pyramid_rulebooks_bit_exact_at_baseline_size(cuda, pair):
    mo, fo, mg, fg, sizes = pair
    prev = None
    for li, s in enumerate(sizes):
        go, gg = mo.grids[s], mg.levels[s]
        assert gg.n == go.n, (li, gg.n, go.n)
        loc_g = gg.locations()
        assert batch_sorted(loc_g)
        kg, og, ig = canon(loc_g)
        ko, oo, io = canon(go.coords)
        assert np.array_equal(kg, ko), "level %d: active sets differ" % li
        # submanifold 3^3 neighbour map, canonical columns and values
        ref = canon_map(R.rules_to_map(mo.subm_rules(s, 3), go.n), oo, io)
        got = canon_map(gg.subm_map(3).cpu().numpy(), og, ig)
        assert np.array_equal(got, ref), "level %d: subm map differs" % li
        if prev is not None:
            ps, (pog, pig), (poo, pio) = prev
            _, rules, parent, off = mo.conv_rules(ps, 2, 2)
            r = mg.strided_rules(ps, 2, 2)
            assert np.array_equal(canon_map(r.cmap.cpu().numpy(), og, pig),
                                  canon_map(R.rules_to_map(rules, go.n), oo, pio)), "level %d: cmap differs" % li
            dref = R.rules_to_map([(p, i) for i, p in rules], len(parent))
            assert np.array_equal(canon_map(r.dmap.cpu().numpy(), pog, ig), canon_map(dref, poo, io)), \
                "level %d: dmap differs" % li
            assert np.array_equal(ig[r.parent_row.cpu().numpy()][pog], io[parent][poo])
        prev = (s, (og, ig), (oo, io))
    # input rule: every point lands on the row of its voxel; mean features agree
    kg, og, ig = canon(mg.levels[sizes[0]].locations())
    ko, oo, io = canon(mo.grids[sizes[0]].coords)
    assert np.array_equal(ig[mg.point_row.cpu().numpy()], io[mo.point_row])
    assert rel_err(fg[torch.from_numpy(og).to(fg.device)], fo[torch.from_numpy(oo)]) <= 1e-6


def _layer_case(pair, cuda, level, C):
    """Canonically defined random input / output gradient for one SubM layer at `level`."""
    from sparse_rcnn_b200 import scn
    mo, fo, mg, fg, sizes = pair
    s = sizes[level]
    go, gg = mo.grids[s], mg.levels[s]
    _, og, ig = canon(gg.locations())
    _, oo, io = canon(go.coords)
    g = torch.Generator().manual_seed(100 * level + C)
    xc = torch.randn(go.n, C, generator=g)
    # TF32-representable inputs so that both precisions see the same operand
    xc = (xc.view(torch.int32) & ~0x1FFF).view(torch.float32)
    gc = torch.randn(go.n, C, generator=g)
    st = torch.tensor(s, dtype=torch.long)
    lo = O.SubmanifoldConvolution(3, C, C, 3, True)
    torch.manual_seed(level)
    lo.weight.data.normal_(0, (2.0 / (27 * C)) ** 0.5)
    lo.bias.data.normal_()
    lg = scn.SubmanifoldConvolution(3, C, C, 3, True)
    lg.load_state_dict(lo.state_dict())
    lg.to(cuda)
    xo = xc[torch.from_numpy(io)].clone().requires_grad_(True)
    yo = lo(O.SparseConvNetTensor(xo, mo, st)).features
    yo.backward(gc[torch.from_numpy(io)])
    ref = dict(y=yo.detach()[torch.from_numpy(oo)], gx=xo.grad[torch.from_numpy(oo)], gw=lo.weight.grad, gb=lo.bias.grad)
    xg0 = xc[torch.from_numpy(ig)].to(cuda)
    gg0 = gc[torch.from_numpy(ig)].to(cuda)
    og_t = torch.from_numpy(og).to(cuda)

    def run():
        lg.zero_grad()
        xg = xg0.clone().requires_grad_(True)
        yg = lg(scn.SparseConvNetTensor(xg, mg, st)).features
        yg.backward(gg0)
        return dict(y=yg.detach()[og_t], gx=xg.grad[og_t], gw=lg.weight.grad.clone(), gb=lg.bias.grad.clone())
    return ref, run, gg.n


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
@pytest.mark.parametrize("level,C", [(0, 32), (1, 48)])
def test_subm_layer_vs_oracle_in_the_large_grid_regime(cuda, pair, monkeypatch, precision, level, C):
    from sparse_rcnn_b200 import scn
    scn.set_precision(precision)
    try:
        ref, run, n = _layer_case(pair, cuda, level, C)
        assert (n + 127) // 128 > 148 * 3            # more tiles than resident CTAs: persistent-grid regime
        variants = [{}]
        if precision == "tf32":
            variants += [{"SCN_CONV_TS": "0"}, {"SCN_CONV_TS": "0", "SCN_CONV_TAILSPLIT": "0"},
                         {"SCN_CONV_TS": "0", "SCN_CONV_SKIP": "0"}]
        for env in variants:
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            got = run()
            for k in env:
                monkeypatch.delenv(k)
            tol = TOL[precision]
            assert rel_err(got["y"], ref["y"]) <= tol, (env, "fwd", rel_err(got["y"], ref["y"]))
            assert rel_err(got["gx"], ref["gx"]) <= tol, (env, "dx", rel_err(got["gx"], ref["gx"]))
            assert rel_err(got["gw"], ref["gw"]) <= 5 * tol, (env, "dw", rel_err(got["gw"], ref["gw"]))
            assert rel_err(got["gb"], ref["gb"]) <= 5 * tol, (env, "db", rel_err(got["gb"], ref["gb"]))
    finally:
        scn.set_precision("tf32")


# Per-parameter L2 bound of the whole-network gradient.  fp32: every forward op agrees to ~2e-6, the residual is ReLU masks
# that flip on inputs straddling zero (measured 2.5e-3 on bias gradients that sum 167k rows).  tf32: the same mechanism
# with 1e-3 forward differences.
GRAD_L2 = {"fp32": 5e-3, "mixed": 2e-2, "tf32": 0.2}


def l2_err(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / b.norm().clamp(min=1e-30))


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("tf32", 2e-3), ("fp32 forward, tf32 backward", 2e-5)])
def test_feature_extractor_vs_oracle_at_baseline_size(cuda, scene, precision, tol):
    """Whole sparse U-Net + segmentation head on the bench scene: every level's features (canonical order) and the per-point
    logits within the forward tolerance; parameter gradients per parameter in L2.

    Why the network GRADIENT is bounded in L2 and per mode (VERDICT r1 "next" 1e): a ReLU input that agrees to rounding but
    straddles zero flips its mask and that flip propagates through the 3^3 stencils.  The third mode separates the two
    effects by measurement: an fp32 forward gives (nearly) the oracle's masks, the TF32 tensor-core kernels then run the
    whole backward over those activations -- every parameter gradient must then agree to 2e-2, which is the bound a wrong
    dgrad / wgrad on any level would break.  With a TF32 forward the masks differ (forward rel err ~1.5e-3 over 44 ReLU
    layers) and the same kernels give ~7e-2 median; that mode keeps the round-1 bound."""
    from sparse_rcnn_b200 import scn
    fwd, bwd = ("fp32", "tf32") if "," in precision else (precision, precision)
    key = "mixed" if "," in precision else precision
    scn.set_precision(fwd)
    try:
        torch.manual_seed(0)
        ref = networks.FeatureExtractor(O)
        seg_o = networks.SegmentationNetwork(O)
        net = networks.FeatureExtractor(scn)
        seg_g = networks.SegmentationNetwork(scn)
        net.load_state_dict(ref.state_dict()), seg_g.load_state_dict(seg_o.state_dict())
        net.to(cuda), seg_g.to(cuda)
        coords, feats, size, bs, splits = scene
        out_o = ref((coords, feats, size, bs, splits))
        out_g = net((coords, feats.to(cuda), size, bs, splits))
        worst = 0.0
        for lo, lg in zip(out_o[4] + out_o[5], out_g[4] + out_g[5]):
            assert lo.features.shape == lg.features.shape
            ko, oo, _ = canon(lo.get_spatial_locations())
            kg, og, _ = canon(lg.get_spatial_locations())
            assert np.array_equal(ko, kg)
            e = rel_err(lg.features[torch.from_numpy(og).to(cuda)], lo.features[torch.from_numpy(oo)])
            worst = max(worst, e)
            assert e <= tol, (tuple(lo.features.shape), e)
        so, sg = seg_o(out_o[5]), seg_g(out_g[5])
        assert so.shape == sg.shape == (len(coords), 20)
        assert rel_err(sg, so) <= tol, rel_err(sg, so)
        g = torch.randn(so.shape, generator=torch.Generator().manual_seed(5)) / so.shape[0]
        so.backward(g)
        scn.set_precision(bwd)
        sg.backward(g.to(cuda))
        errs = {n: l2_err(pg.grad, po.grad) for (n, po), (_, pg) in zip(ref.named_parameters(), net.named_parameters())
                if po.grad is not None}
        bad = {n: e for n, e in errs.items() if e > GRAD_L2[key]}
        print("[%s] worst forward rel err %.2e; parameter-gradient L2: median %.2e max %.2e (%s)" % (
            precision, worst, float(np.median(list(errs.values()))), max(errs.values()), max(errs, key=errs.get)))
        assert not bad, bad
    finally:
        scn.set_precision("tf32")
