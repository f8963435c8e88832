"""The drop-in boundary exercised with the reference's OWN module tree (VERDICT r1 missing #5/#6):
the unmodified `ndsis.modules.model` classes are built on `sparse_rcnn_b200.scn` aliased as `sparseconvnet` and run forward
on CUDA; their outputs are compared with tests/golden/ref_graph.pt -- the same unmodified classes running on the CPU oracle
(oracle/make_golden.py).  The reference package is not part of this repository: it is looked for in /root/reference (build
container) and in baseline/_ref (a git-ignored copy that travels to the GPU box); without it the test is skipped."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_dir():
    for d in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isdir(os.path.join(d, "ndsis", "modules")):
            return d
    return None


def test_scn_sequential_recognises_the_reference_residual_units():
    """CPU: `scn.Sequential` plans fused execution for the module pattern of module_factory.py:51-57,127-183 and for
    nothing else; the module tree (state_dict keys) is untouched."""
    from sparse_rcnn_b200 import networks, scn
    from sparse_rcnn_b200.scn import layers
    stage = networks.unit_stage(scn, 32, 2)
    plan = stage._plan()
    assert plan is not None and len(plan) == 1 and plan[0][0] == "units" and len(plan[0][1]) == 2
    assert networks.residual_unit(scn, 32, 32)._plan()[0][0] == "units"          # a lone unit
    assert networks.residual_unit(scn, 64, 32)._plan() is None                    # NetworkInNetwork shortcut: not fused
    mixed = scn.Sequential(scn.SubmanifoldConvolution(3, 6, 32, 1, True), networks.residual_unit(scn, 32, 32),
                           networks.residual_unit(scn, 32, 32), scn.ReLU(), networks.residual_unit(scn, 32, 32))
    assert [k for k, _ in mixed._plan()] == ["module", "units", "module", "units"]
    assert len(mixed._plan()[1][1]) == 2
    bn = scn.Sequential(scn.ConcatTable(scn.Identity(), scn.Sequential(
        scn.BatchNormReLU(32), scn.SubmanifoldConvolution(3, 32, 32, 3, False),
        scn.BatchNormReLU(32), scn.SubmanifoldConvolution(3, 32, 32, 3, False))), scn.AddTable())
    assert bn._plan() is None                                                    # batch-norm variant: plain graph
    mixed.append(networks.residual_unit(scn, 32, 32))                             # the plan follows later edits
    assert len(mixed._plan()[-1][1]) == 2
    assert [k for k in stage.state_dict()] == ["0.0.1.1.weight", "0.0.1.1.bias", "0.0.1.3.weight", "0.0.1.3.bias",
                                               "1.0.1.1.weight", "1.0.1.1.bias", "1.0.1.3.weight", "1.0.1.3.bias"]
    assert layers.FUSE is networks.FUSE


def test_reference_modules_build_on_the_backend_with_fused_plans():
    """CPU, build container only: the unmodified FeatureExtractor constructs on the B200 namespace, has the mirror's
    state_dict, and every one of its unit stages is recognised by the fusion planner."""
    ref = reference_dir()
    if ref is None:
        pytest.skip("reference package not present")
    code = (
        "import sys, importlib.util; sys.path[:0]=[%r, %r, %r]\n"
        "spec=importlib.util.spec_from_file_location('mg', %r); mg=importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)\n"
        "import sparse_rcnn_b200; scn=sparse_rcnn_b200.install_as_sparseconvnet()\n"
        "from ndsis.modules.model import FeatureExtractor\n"
        "from sparse_rcnn_b200 import networks\n"
        "from sparse_rcnn_b200.scn import layers\n"
        "fe_p, unet_p, _, _ = mg.reference_configs()\n"
        "ref = FeatureExtractor(**fe_p, include_unet=True, unet_params=unet_p)\n"
        "mine = networks.FeatureExtractor(scn)\n"
        "a={k:tuple(v.shape) for k,v in ref.state_dict().items()}; b={k:tuple(v.shape) for k,v in mine.state_dict().items()}\n"
        "assert a==b, set(a)^set(b)\n"
        "count=lambda net: sum(len(item) for m in net.modules() if isinstance(m, layers.Sequential) and m._plan() for kind,item in m._plan() if kind=='units')\n"
        "assert count(ref)==count(mine)==44, (count(ref), count(mine))  # 22 units, each seen by its stage and by itself\n"
        "print('OK', len(a))\n") % (ROOT, os.path.join(ROOT, "oracle"), ref, os.path.join(ROOT, "oracle", "make_golden.py"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OK 120" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("tf32", 3e-3), ("fp32", 1e-4)])
def test_unmodified_reference_modules_forward_on_cuda(cuda, tmp_path, precision, tol):
    ref = reference_dir()
    if ref is None:
        pytest.skip("reference package not present (neither /root/reference nor baseline/_ref)")
    out = str(tmp_path / "ref_on_b200.pt")
    env = dict(os.environ, SCN_PRECISION=precision)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "run_reference_on_b200.py"), ref, out],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
    got = torch.load(out, weights_only=False)
    g = torch.load(os.path.join(ROOT, "tests", "golden", "ref_graph.pt"), weights_only=False)
    assert got["state_keys"] == g["state"]["fe"]
    assert got["fused_sequentials"] >= 11                      # 6 encoder + 5 decoder unit stages (+ heads)
    f, p = got["fused"], got["plain"]
    assert f["launches"] < p["launches"]                        # the fused path really ran, and is the default
    # the reference's own graph reaches the same kernels as the repository's mirror: identical bits
    assert torch.equal(f["unet_last"], got["mirror_unet_last"])
    # rows follow the Morton curve here and first appearance in the golden: compare through the canonical sort
    from scn_oracle import rules as R
    import numpy as np
    order = np.argsort(R.pack_keys(f["locations"].numpy()), kind="stable")
    for res in (f, p):
        assert res["level_rows"] == g["level_rows"]
        assert res["cls_inside"] == g["cls_inside_count"] and res["mask_inside"] == g["mask_inside_count"]
        assert res["mask"].shape[0] == g["mask_rows"]
        amax = g["unet_last_absmax"]
        u = res["unet_last"][order]
        assert abs(float(u.double().sum()) - g["unet_last_sum"]) <= tol * u.numel() ** 0.5 * amax * 4
        c = res["cls"]
        assert float((c - g["cls_out"]).abs().max()) <= tol * float(g["cls_out"].abs().max()), "class logits"
    # point-ordered outputs (segmentation, mask logits) do not depend on the row order
    assert float((f["seg"][:64] - g["seg_head"]).abs().max()) <= tol * max(1.0, float(g["seg_head"].abs().max()))
    assert float((f["mask"][:64] - g["mask_head"]).abs().max()) <= tol * max(1.0, float(g["mask_head"].abs().max()))
    if precision == "fp32":                                     # fused and plain graphs: same arithmetic in fp32 mode
        assert float((f["unet_last"] - p["unet_last"]).abs().max()) <= 1e-5 * g["unet_last_absmax"]
