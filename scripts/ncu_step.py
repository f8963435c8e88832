"""One training step + one inference scene of the bench workload between cudaProfilerStart/Stop, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum ...
-> per-launch duration and DRAM/L2 bytes of EVERY kernel of the step (profiles/r2_c_kernel_table.md is made from the csv by
scripts/ncu_table.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from sparse_rcnn_b200 import pipeline, scn
from sparse_rcnn_b200.synthetic import make_boxes
dev = torch.device("cuda:0")
scn.set_precision("tf32")
trainer = pipeline.BackboneTrainer(dev, distributed=False)
host = [bench.make_inputs(i) for i in range(2)]
res = [((d[0].to(dev), d[1].to(dev), d[2], d[3], d[4]), l.to(dev)) for d, l in host]
infer = pipeline.SparseInference(dev)
boxes = [make_boxes(d[0], 256, 7 + i) for i, (d, _) in enumerate(host)]
for i in range(4):
    trainer.step(*res[i % 2])
    infer(res[i % 2][0], boxes[i % 2])
torch.cuda.synchronize()
mode = sys.argv[1] if len(sys.argv) > 1 else "both"
torch.cuda.cudart().cudaProfilerStart()
if mode in ("both", "train"):
    trainer.step(*res[0])
if mode in ("both", "infer"):
    infer(res[0][0], boxes[0])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done", trainer.last_active)
