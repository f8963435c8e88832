"""Measured BEFORE building a bf16 operand path (missing #4): what operand precision does to the sparse U-Net on
the BASELINE scene.  Runs the CPU oracle's FeatureExtractor + segmentation head (fp32 accumulation everywhere) with the
gathered features and the weights of every convolution rounded to fp32 (reference) / TF32 (10-bit mantissa, round to nearest,
what the tcgen05 kind::tf32 path computes after scn_round_tf32) / bf16 (7-bit mantissa, round to nearest even), and reports
the relative error (max |a - b| / max |b|, the tests' rel_err) of every encoder / decoder output and the logits, plus the
gradient of the loss wrt the input features.  CPU only: python tests/tools/measure_bf16.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import scn_oracle as O
import bench
from sparse_rcnn_b200 import networks

def round_tf32(x):
    i = x.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF            # round to nearest (ties away), 13 mantissa bits dropped
    return i.view(torch.float32)
ROUND = {"fp32": lambda x: x, "tf32": round_tf32, "bf16": lambda x: x.bfloat16().float()}
mode = ["fp32"]
orig = O._rule_conv
def rule_conv(x, w, bias, rules, n_out, swap=False):
    r = ROUND[mode[0]]
    return orig(r(x), r(w), bias, rules, n_out, swap)
O._rule_conv = rule_conv
def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())
torch.set_num_threads(os.cpu_count() or 1)
torch.manual_seed(0)
net, seg = networks.FeatureExtractor(O), networks.SegmentationNetwork(O)
data, labels = bench.make_inputs(0)
res = {}
for m in ("fp32", "tf32", "bf16"):
    mode[0] = m
    x = data[1].clone().requires_grad_(True)
    out = net((data[0], x, data[2], data[3], data[4]))
    logits = seg(out[5])
    loss = torch.nn.functional.cross_entropy(logits, labels)
    loss.backward()
    res[m] = ([t.features.detach() for t in out[4]] + [t.features.detach() for t in out[5]], logits.detach(), x.grad.detach(), float(loss))
ref = res["fp32"]
print("N = %d active voxels; %d tensors compared" % (ref[0][0].shape[0], len(ref[0])))
for m in ("tf32", "bf16"):
    acts, lg, gx, loss = res[m]
    errs = [rel(a, b) for a, b in zip(acts, ref[0])]
    print("%s operands: encoder outputs %s | decoder outputs %s | logits %.2e | d loss / d input %.2e | loss %.6f vs %.6f" % (
        m, " ".join("%.1e" % e for e in errs[:6]), " ".join("%.1e" % e for e in errs[6:]), rel(lg, ref[1]), rel(gx, ref[2]), loss, ref[3]))
