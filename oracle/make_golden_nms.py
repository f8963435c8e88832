"""Golden vectors for the proposal-selection row, produced by the UNMODIFIED reference
(`ndsis.modules.proposal_selector.ProposalSelector`, `ndsis.utils.bbox.non_maximum_supression`) on the CPU.
Run in the build container (needs /root/reference):  python oracle/make_golden_nms.py  -> tests/golden/nms.pt"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)
from ndsis.modules.proposal_selector import ProposalSelector          # noqa: E402
from ndsis.utils.bbox import non_maximum_supression                   # noqa: E402


from sparse_rcnn_b200.synthetic import make_proposals as make_case      # noqa: E402  (seeded generator, ships)


cases = []
for seed, B, A, pre, post, thr in [(0, 2, 3000, 1024, 256, 0.5), (1, 3, 1500, 1024, 32, 0.3), (2, 1, 700, 512, 256, 0.5),
                                   (3, 2, 300, 0, 500, 0.5), (4, 1, 40, 33, 10, 0.1)]:
    score, boxes = make_case(seed, B, A, clustered=seed != 3)
    sel = ProposalSelector(pre, post, thr)
    s, b, i = sel(score, boxes)
    # the raw indicator of the reference's NMS on the sorted boxes (what scn_nms3d's `keep` must equal)
    if pre > 0:
        _, order = torch.topk(score, pre, dim=1, sorted=True)
    else:
        _, order = torch.sort(score, dim=1, descending=True)
    sorted_boxes = boxes[torch.arange(B)[:, None], order]
    keep = non_maximum_supression(sorted_boxes, thr)
    cases.append(dict(seed=seed, B=B, A=A, pre=pre, post=post, thresh=thr, clustered=seed != 3,
                      scores=[t.clone() for t in s], indices=[t.clone() for t in i], keep=keep.clone(),
                      n_kept=[int(k.sum()) for k in keep]))
    print("case", seed, "kept per sample", cases[-1]["n_kept"], "returned", [len(t) for t in s])
torch.save(dict(cases=cases, generator="sparse_rcnn_b200.synthetic.make_proposals"), os.path.join(ROOT, "tests", "golden", "nms.pt"))
print("wrote tests/golden/nms.pt", os.path.getsize(os.path.join(ROOT, "tests", "golden", "nms.pt")), "bytes")
