"""Proposal selection on the GPU: bit-exact against the unmodified reference's outputs
(tests/golden/nms.pt) and against the pinned CPU restatement (oracle/scn_oracle/nms_ref.py)."""
import os

import pytest
import torch

from scn_oracle import nms_ref
from sparse_rcnn_b200.synthetic import make_proposals

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


def test_selector_matches_reference_goldens(cuda):
    from sparse_rcnn_b200 import proposal
    gold = torch.load(os.path.join(G, "nms.pt"))
    for c in gold["cases"]:
        score, boxes = make_proposals(c["seed"], c["B"], c["A"], clustered=c["clustered"])
        sel = proposal.ProposalSelector(c["pre"], c["post"], c["thresh"])
        s, b, i = sel(score.to(cuda), boxes.to(cuda))
        assert len(s) == len(b) == len(i) == c["B"]
        for k in range(c["B"]):
            assert not i[k].is_cuda and i[k].dtype == torch.int64          # reference returns CPU indices
            assert torch.equal(i[k], c["indices"][k])
            assert torch.equal(s[k].cpu(), c["scores"][k])
            assert torch.equal(b[k].cpu(), boxes[k][c["indices"][k]])
        # the raw NMS indicator on the reference's sorted order
        if c["pre"] > 0:
            _, order = torch.topk(score, c["pre"], dim=1, sorted=True)
        else:
            _, order = torch.sort(score, dim=1, descending=True)
        sorted_boxes = boxes[torch.arange(c["B"])[:, None], order]
        keep, keep_idx, counts = proposal.nms3d(sorted_boxes.to(cuda), c["thresh"], c["post"])
        assert torch.equal(keep.cpu(), c["keep"])
        assert counts.cpu().tolist() == [min(n, c["post"]) for n in c["n_kept"]]


@pytest.mark.parametrize("B,n,thr", [(1, 1, 0.5), (2, 31, 0.2), (1, 32, 0.2), (3, 33, 0.4), (1, 2000, 0.3), (2, 4096, 0.5)])
def test_nms_vs_restatement_edge_sizes(cuda, B, n, thr):
    """word boundaries, a single box, the global-memory path (n > 1264) and the maximum size"""
    from sparse_rcnn_b200 import proposal
    score, boxes = make_proposals(100 + n, B, n)
    order = torch.sort(score, dim=1, descending=True)[1]
    sb = boxes[torch.arange(B)[:, None], order]
    keep, keep_idx, counts = proposal.nms3d(sb.to(cuda), thr, n)
    for k in range(B):
        ref = nms_ref.nms(sb[k], thr)
        assert torch.equal(keep[k].cpu(), ref)
        cnt = int(counts[k])
        assert cnt == int(ref.sum())
        assert torch.equal(keep_idx[k, :cnt].cpu().long(), torch.nonzero(ref).flatten())


def test_nms_degenerate_and_empty(cuda):
    from sparse_rcnn_b200 import proposal
    # zero-volume boxes: IoU is NaN (0/0) in the reference arithmetic, never above the threshold -> all kept
    z = torch.zeros(1, 5, 2, 3)
    keep, _, counts = proposal.nms3d(z.to(cuda), 0.5, 5)
    assert keep.all() and int(counts[0]) == 5 and nms_ref.nms(z[0], 0.5).all()
    # identical boxes: only the first survives
    one = torch.tensor([[[0., 0, 0], [4, 4, 4]]]).repeat(7, 1, 1)[None]
    keep, keep_idx, counts = proposal.nms3d(one.to(cuda), 0.5, 7)
    assert keep[0].tolist() == [True] + [False] * 6 and int(counts[0]) == 1 and int(keep_idx[0, 0]) == 0
    # no proposals at all
    keep, keep_idx, counts = proposal.nms3d(torch.zeros(2, 0, 2, 3, device=cuda), 0.5, 4)
    assert keep.shape == (2, 0) and counts.tolist() == [0, 0]
    with pytest.raises(RuntimeError):
        proposal.nms3d(torch.zeros(1, 4, 2, 3), 0.5)                       # CPU tensor: no fallback


def test_nms_full_size_properties(cuda):
    """BASELINE size (8 samples x 1024 proposals): no two survivors overlap above the threshold, every suppressed box
    overlaps a better survivor, and NMS is idempotent on its own output."""
    from sparse_rcnn_b200 import proposal
    B, n, thr = 8, 1024, 0.5
    score, boxes = make_proposals(7, B, n)
    order = torch.sort(score, dim=1, descending=True)[1]
    sb = boxes[torch.arange(B)[:, None], order].to(cuda)
    keep, keep_idx, counts = proposal.nms3d(sb, thr, n)
    for k in range(B):
        iou = nms_ref.iou_matrix(sb[k].cpu())
        kk = keep[k].cpu()
        sub = iou[kk][:, kk]
        sub.fill_diagonal_(0)
        assert float(sub.max()) <= thr
        dead = torch.nonzero(~kk).flatten()
        better = torch.tril(iou > thr, diagonal=-1)                        # [i, j]: j better than i and overlapping
        assert bool((better[dead][:, kk].any(dim=1)).all())
        again, _, c2 = proposal.nms3d(sb[k][kk.to(cuda)][None], thr, n)
        assert bool(again.all()) and int(c2[0]) == int(kk.sum())


# ----------------------------------------------------------------------------- per-box mask loss
def test_mask_loss_matches_reference_goldens(cuda):
    """floating-point kernel: 1e-5 relative on the loss and on the gradients, against the unmodified reference's MaskLoss
    (tests/golden/maskloss.pt) incl. empty boxes, all-empty and no-box cases, with and without class weights."""
    from sparse_rcnn_b200 import losses
    from sparse_rcnn_b200.synthetic import make_mask_loss_case
    from tests.util import rel_err
    gold = torch.load(os.path.join(G, "maskloss.pt"))
    for c in gold["cases"]:
        outs, tgts, cls = make_mask_loss_case(c["seed"], c["boxes_per_sample"], empty_every=c["empty_every"])
        outs = [[m.to(cuda).requires_grad_() for m in s] for s in outs]
        tgts = [[m.to(cuda) for m in s] for s in tgts]
        cls = [t.to(cuda) for t in cls]
        w = (torch.arange(18, dtype=torch.float32) % 5 + 0.5).to(cuda) if c["weighted"] else None
        ml = losses.MaskLoss(class_weights=w).to(cuda)
        loss = ml(outs, tgts, cls) if c["boxes_per_sample"] else ml([], [], [torch.zeros(0, dtype=torch.long, device=cuda)])
        assert abs(float(loss) - c["loss"]) <= 1e-5 * max(abs(c["loss"]), 1.0), (float(loss), c["loss"])
        if c["grads"]:
            loss.backward()
            flat = [m for s in outs for m in s]
            for m, g in zip(flat, c["grads"]):
                got = m.grad if m.grad is not None else torch.zeros_like(m)
                assert got.shape == g.shape
                if len(g):
                    assert rel_err(got, g) < 1e-5


def test_segment_bce_full_size(cuda):
    """BASELINE size: 256 boxes, 615k (box, point) rows: equals torch's elementwise op reduced per box (1e-5), and the sum of
    the gradients of a box equals mean(sigmoid(x) - t) (linearity of the backward)."""
    from sparse_rcnn_b200 import losses
    g = torch.Generator().manual_seed(3)
    counts = torch.randint(0, 4800, (256,), generator=g).tolist()
    counts[10] = counts[200] = 0
    M = sum(counts)
    x = (torch.randn(M, generator=g) * 3).to(cuda).requires_grad_()
    t = (torch.rand(M, generator=g) > 0.5).to(cuda)
    means = losses.segment_bce_with_logits(x, t, counts)
    el = torch.nn.functional.binary_cross_entropy_with_logits(x.detach(), t.float(), reduction="none").double()
    ptr = torch.tensor([0] + counts).cumsum(0)
    ref = torch.stack([el[a:b].mean() if b > a else torch.tensor(float("nan"), device=cuda, dtype=torch.float64)
                       for a, b in zip(ptr[:-1].tolist(), ptr[1:].tolist())])
    ok = ~torch.isnan(ref)
    assert torch.equal(torch.isnan(means), ~ok)
    assert float(((means.double() - ref)[ok].abs() / ref[ok]).max()) < 1e-5
    torch.nan_to_num(means, nan=0.0).sum().backward()
    sg = (torch.sigmoid(x.detach().double()) - t.double())
    for a, b in list(zip(ptr[:-1].tolist(), ptr[1:].tolist()))[:40]:
        if b > a:
            assert abs(float(x.grad[a:b].double().sum()) - float(sg[a:b].mean())) < 1e-5
