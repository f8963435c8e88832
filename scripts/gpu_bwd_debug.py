"""Locate the first backward mismatch between oracle and GPU backbone (fp32 mode)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import scn_oracle as O
from sparse_rcnn_b200 import networks, scn
from sparse_rcnn_b200.synthetic import make_batch
from tests.util import rel_err

dev = torch.device("cuda:0")
scn.set_precision(sys.argv[1] if len(sys.argv) > 1 else "fp32")
torch.manual_seed(0)
ref = networks.FeatureExtractor(O); net = networks.FeatureExtractor(scn)
seg_o = networks.SegmentationNetwork(O); seg_g = networks.SegmentationNetwork(scn)
net.load_state_dict(ref.state_dict()); seg_g.load_state_dict(seg_o.state_dict()); net.to(dev); seg_g.to(dev)
coords, feats, size, bs, splits = make_batch(2, 3, spatial_size=(64, 64, 32), room=(44, 44, 22), room_offset=(8, 8, 2), n_furniture=4)
fo = feats.clone().requires_grad_(True); fg = feats.to(dev).requires_grad_(True)

# record every intermediate feature tensor through module forward hooks
def instrument(model, store):
    def hook(mod, inp, out):
        f = out.features if hasattr(out, "features") else (out if isinstance(out, torch.Tensor) else None)
        if f is not None and f.requires_grad:
            f.retain_grad()
            store.append((type(mod).__name__, f))
    for m in model.modules():
        if type(m).__name__ in ("SubmanifoldConvolution", "Convolution", "Deconvolution", "NetworkInNetwork", "ReLU", "AddTable", "JoinTable"):
            m.register_forward_hook(hook)
so_, sg_ = [], []
instrument(ref, so_); instrument(net, sg_)
out_o = ref((coords, fo, size, bs, splits)); out_g = net((coords, fg, size, bs, splits))
so, sg = seg_o(out_o[5]), seg_g(out_g[5])
g = torch.randn_like(so)
so.backward(g); sg.backward(g.to(dev))
print("n records", len(so_), len(sg_))
for i, ((n1, a), (n2, b)) in enumerate(zip(so_, sg_)):
    e_f = rel_err(b, a)
    e_g = rel_err(b.grad, a.grad) if a.grad is not None and b.grad is not None else -1
    flag = "  <<<<" if e_g > 1e-4 else ""
    print("%3d %-24s %-14s fwd %.2e grad %.2e%s" % (i, n1, tuple(a.shape), e_f, e_g, flag))
print("input grad", rel_err(fg.grad, fo.grad))
bad = (fg.grad.cpu() - fo.grad).abs().max(1).values
print("rows with err > 1e-3:", int((bad > 1e-3).sum()), "of", len(bad), "first bad rows", (bad > 1e-3).nonzero()[:10].flatten().tolist())
for (n, po), (_, pg) in zip(ref.named_parameters(), net.named_parameters()):
    e = rel_err(pg.grad, po.grad)
    if e > 1e-4: print("param", n, e)
