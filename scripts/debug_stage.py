import os, sys, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from sparse_rcnn_b200 import pipeline, scn
from sparse_rcnn_b200.scn import metadata as M
dev = torch.device("cuda:0"); scn.set_precision("tf32")
tr = pipeline.BackboneTrainer(dev)
host = [bench.balanced_inputs(0, i) for i in range(4)]
pinned = [((d[0].pin_memory(), d[1].pin_memory(), d[2], d[3], d[4]), l.pin_memory()) for d, l in host]
orig = M.take_staged
def checked(t):
    d = orig(t)
    if d is not None:
        torch.cuda.current_stream().synchronize()
        ok = torch.equal(d.cpu(), t)
        if not ok:
            print("MISMATCH in take_staged: shape %s dtype %s thread %s step %d ptr %x" % (tuple(t.shape), t.dtype, threading.current_thread().name, STEP[0], d.data_ptr()), flush=True)
    return d
M.take_staged = checked
pipeline.take_staged = checked
STEP = [0]
tr.stage_uploads = True; tr.build_late = "thread"
for i in range(4): tr.step(*pinned[i])
for i in range(40):
    STEP[0] = i
    tr.step(*pinned[i % 4], next_batch=pinned[(i + 1) % 4])
torch.cuda.synchronize()
print("done; ring:", [(a.buf.numel(), a.buf.data_ptr()) for a in M._stage_rings[dev][0] if a is not None])
