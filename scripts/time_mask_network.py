"""BASELINE config 4: the sparse Mask Network batched over 256 variable-shape proposals of one scene (sparse crop + mask
convolutions), timed alone with CUDA events on the bench scene, next to the other parts of the sparse inference pass."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from sparse_rcnn_b200 import pipeline, roi, scn, _lib
from sparse_rcnn_b200.synthetic import make_boxes

dev = torch.device("cuda:0"); scn.set_precision("tf32")
inf = pipeline.SparseInference(dev)
data, _ = bench.make_inputs(0)
boxes = make_boxes(data[0], 256, 7)
ddata = pipeline._to_device(data, dev)


def parts():
    roi.clear_key_cache()
    out = inf.backbone(ddata)
    scene_size, batch_size, _, class_map, inter, unet = out
    roi.register_keys(ddata[0], inter[0].metadata.point_keys)
    return class_map, unet


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    l0 = _lib.raw("scn_launch_count")()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n, (_lib.raw("scn_launch_count")() - l0) // n


with torch.no_grad():
    class_map, unet = parts()
    res = inf.mask_network(ddata, unet, boxes)
    pts = int(res[0].shape[0])
    rows = [("backbone (rulebooks + U-Net)", timed(lambda: parts())),
            ("segmentation head", timed(lambda: inf.seg(unet))),
            ("class network (crop at stride 4 + 2 levels + pool + MLP), 256 boxes", timed(lambda: inf.class_network(class_map, boxes))),
            ("MASK NETWORK (crop + U-Net 22-32-48-64 + MLP), 256 boxes, %d (box, point) rows" % pts,
             timed(lambda: inf.mask_network(ddata, unet, boxes)))]
print("| part | GPU-elapsed ms | host wall ms | scn launches |")
print("|---|---:|---:|---:|")
for name, (g, h, l) in rows:
    print("| %s | %.2f | %.2f | %d |" % (name, g, h, l))
print("mask network: %.1f M (box, point) rows/s" % (pts / rows[-1][1][0] / 1e3))
