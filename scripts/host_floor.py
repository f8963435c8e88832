"""Host issue floor of the training step: the same step on a tiny scene (GPU work negligible) is pure host time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
from sparse_rcnn_b200 import pipeline, scn
dev = torch.device("cuda:0"); scn.set_precision("tf32")
tr = pipeline.BackboneTrainer(dev)
for name, kw in (("tiny 64x64x32", dict(spatial_size=(64, 64, 32), room=(40, 40, 20), room_offset=(8, 8, 4), n_furniture=3)),
                 ("bench scene", bench.SCENE)):
    data, labels = bench.make_inputs(0, scene_kw=kw)
    data = (data[0], data[1].to(dev), data[2], data[3], data[4]); labels = labels.to(dev)
    for _ in range(5): tr.step(data, labels)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): tr.step(data, labels)
    t_cpu = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print("%s: N=%d  CPU issue %.2f ms/step, wall %.2f ms/step" % (name, tr.last_active, t_cpu * 50, t_all * 50))
if len(sys.argv) > 1:
    import cProfile, pstats, io
    data, labels = bench.make_inputs(0, scene_kw=dict(spatial_size=(64, 64, 32), room=(40, 40, 20), room_offset=(8, 8, 4), n_furniture=3))
    data = (data[0], data[1].to(dev), data[2], data[3], data[4]); labels = labels.to(dev)
    for _ in range(3): tr.step(data, labels)
    pr = cProfile.Profile(); pr.enable()
    for _ in range(10): tr.step(data, labels)
    torch.cuda.synchronize(); pr.disable()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats(sys.argv[1]).print_stats(45); print(s.getvalue()[:9000])
