"""Forward and backward of the sparse U-Net executor on the bench scene, CUDA events around the calls (no rulebook work, no
loss, no optimizer): what a switch (SCN_EXEC_DUAL, SCN_EXEC_SIDE, SCN_PDL ...) changes in each direction."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sparse_rcnn_b200 import networks, scn
from sparse_rcnn_b200.synthetic import make_batch
dev = torch.device("cuda:0"); scn.set_precision("tf32")
torch.manual_seed(0)
net = networks.FeatureExtractor(scn).to(dev)
coords, feats, size, bs, splits = make_batch(1, 0)
md = scn.Metadata(3)
size_t = torch.as_tensor(size, dtype=torch.long)
f = scn.ioLayers.InputLayerFunction.apply(3, md, size_t, coords, feats.to(dev), bs, 4)
md.prebuild(5, book_channels=[32, 48, 64, 80, 96, 112])
ex = net._executor()
R = int(os.environ.get("REPS", "20"))
x0 = scn.SparseConvNetTensor(f, md, size_t)
with torch.no_grad():
    for _ in range(3): ex.run(x0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(R): ex.run(x0)
    e1.record(); torch.cuda.synchronize()
fw = e0.elapsed_time(e1) / R
bws = []
for i in range(R + 3):
    x = scn.SparseConvNetTensor(f.clone().requires_grad_(True), md, size_t)
    enc, dec = ex.run(x)
    y = dec[-1].features
    g = torch.ones_like(y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); y.backward(g); e1.record(); torch.cuda.synchronize()
    if i >= 3: bws.append(e0.elapsed_time(e1))
bws.sort()
print("DUAL=%s SIDE=%s PDL=%s: forward %.3f ms (mean of %d back to back), backward median %.3f ms (min %.3f)" % (
    os.environ.get("SCN_EXEC_DUAL", "1"), os.environ.get("SCN_EXEC_SIDE", "1"), os.environ.get("SCN_PDL", "1"), fw, R, bws[len(bws) // 2], bws[0]))
