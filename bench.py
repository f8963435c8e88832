#!/usr/bin/env python
"""bench.py -- headline benchmark of the sparse hot path (contract in the task description).

metric  : active voxels/s through the sparse U-Net backbone, forward + backward
          (BASELINE.json metric (i); workload = BASELINE configs[0]/[1] scene: one synthetic
          ScanNet-sized scene per GPU per step, ~167k active voxels, 6 input channels).
step    : rulebook build (GPU hash + neighbour maps) -> backbone fwd -> seg-head cross entropy ->
          bwd -> [gradient allreduce, N>1] -> Adam.  A fresh Metadata is built every step, like
          the reference (custom_operations.py:70).
value   : coords/features already resident in HBM when the timed region starts.
e2e     : same step through the public API with HOST (pinned) buffers: H2D of coords+features+
          labels and D2H of the loss inside the timed region, every step.
roofline: dominant kernel = tcgen05 TF32 gather-GEMM (k_conv_tc) on the level-0 SubM 3^3 32->32
          layer, timed alone with CUDA events and an L2 flush between launches.
cpu_baseline / --impl reference: the CPU oracle (SparseConvNet-CPU-style restatement; SparseConvNet
          itself is not installable here) on the host cores, bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch

METRIC = "active_voxels_per_sec_sparse_backbone_fwd_bwd"
UNIT = "voxels/s"
SCENE = dict(spatial_size=(256, 256, 128))
WORKLOAD = ("sparse U-Net feature extractor (6->[32,48,64,80,96,112] + U-Net decoder, 44 SubM 3^3 convs) "
            "fwd+bwd+Adam on 1 synthetic ScanNet-sized scene/GPU/step (256x256x128 grid, ~167k active "
            "voxels, ~273k points, 6 ch) = BASELINE configs[0]/[1] scene")
CPU_SAMPLE = dict(spatial_size=(128, 128, 64), room=(88, 88, 44), room_offset=(16, 16, 4), n_furniture=8)


def env_int(name, default):
    return int(os.environ.get(name, default))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_inputs(seed, in_channels=6, scene_kw=SCENE, num_classes=20):
    from sparse_rcnn_b200.synthetic import make_batch
    data = make_batch(1, seed, in_channels=in_channels, **scene_kw)
    g = torch.Generator().manual_seed(seed)
    labels = torch.randint(0, num_classes, (len(data[0]),), generator=g)
    return data, labels


def active_voxels(data):
    c = data[0].numpy()
    key = (c[:, 3].astype("int64") << 48) | (c[:, 0].astype("int64") << 32) | (c[:, 1].astype("int64") << 16) | c[:, 2].astype("int64")
    import numpy as np
    return int(np.unique(key).size)


def balanced_inputs(rank, i, tol=0.01, tries=40):
    """Scene i of rank `rank`: rank 0 uses seed i; every other rank draws its OWN scenes (different geometry) but keeps only
    one whose active-voxel count is within 1 % of rank 0's scene i, so that at N > 1 the max-over-ranks step time measures
    the collective and not which rank happened to draw the largest scene."""
    ref = make_inputs(i)
    if rank == 0:
        return ref
    want = active_voxels(ref[0])
    best, best_err = None, None
    for k in range(tries):
        cand = make_inputs(1000 * rank + i + 100 * k)
        err = abs(active_voxels(cand[0]) - want) / want
        if best is None or err < best_err:
            best, best_err = cand, err
        if err <= tol:
            break
    return best


def run_cpu_oracle(steps, warmup, sample_kw=SCENE):
    """Backbone fwd+bwd on the CPU oracle (the SparseConvNet-CPU-style path); returns (voxels/s, info)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import scn_oracle as O
    from sparse_rcnn_b200 import networks
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    net = networks.FeatureExtractor(O)
    seg = networks.SegmentationNetwork(O)
    data, labels = make_inputs(0, scene_kw=sample_kw)
    opt = torch.optim.Adam(list(net.parameters()) + list(seg.parameters()), lr=1e-3)
    times, n_active = [], 0
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        out = net(data)
        loss = torch.nn.functional.cross_entropy(seg(out[5]), labels)
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        n_active = out[4][0].features.shape[0]
        if i >= warmup:
            times.append(dt)
    t = sum(times) / len(times)
    info = {"cores": cores, "kind": "port",
            "sample": "the bench scene itself (%s grid, N=%d active voxels), fwd+bwd+Adam incl. rulebook build, %d step(s)" % (
                "x".join(map(str, sample_kw["spatial_size"])), n_active, steps)}
    return n_active / t, t, info


def reference_arm(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    steps = min(args.steps, 3)
    warm = min(args.warmup, 1)
    v, t, info = run_cpu_oracle(steps, warm)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "parallelism": "host cores (torch intra-op threads), rank 0 only",
                       "precision": "fp32",
                       "note": "SparseConvNet is not vendored/installable (unpinned external dependency); this is the "
                               "repo's CPU restatement of its algorithm on the host cores"},
            "cpu_baseline": dict(info, value=v, unit=UNIT),
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def time_kernels(dev, md_size, n_iter=20):
    """Live roofline measurements on the level-0 geometry of a bench scene: the dominant kernel (SubM 3^3 32->32 forward,
    k_conv_ts), its weight gradient (k_conv_wgrad_tc) and the rulebook kernel (k_subm_map).  Each kernel is launched through
    its C-ABI entry point directly (no Python layer code between the events), n_iter times with an L2 flush (256 MB memset)
    in front of every launch; all event pairs are queued before the one synchronisation, and the flush kernel in front of
    each pair gives the host ~80 us to enqueue the next launch, so host jitter is not inside any pair."""
    from sparse_rcnn_b200 import scn, _lib
    from sparse_rcnn_b200.scn import functions as Fn
    from sparse_rcnn_b200.scn.metadata import _stream
    md, size = md_size
    lvl = md.level(size)
    n, C, K = lvl.n, 32, 27
    P = lambda t: t.data_ptr()
    s = _stream()
    m = lvl.subm_map(3)
    pairs = int((m >= 0).sum().item())
    w = torch.randn(K, C, C, device=dev) * 0.05
    img = torch.empty(int(_lib.raw("scn_conv_weight_image_bytes")(K, C, C)), dtype=torch.uint8, device=dev)
    _lib.call("scn_conv_pack_weights", P(w), K, C, C, 0, 0, P(img), s)
    x = Fn.tf32_exact(torch.randn(n, C, device=dev))
    go = Fn.tf32_exact(torch.randn(n, C, device=dev))
    out = torch.empty(n, C, device=dev)
    gw, gb = torch.zeros(K, C, C, device=dev), torch.zeros(C, device=dev)
    m2 = torch.empty_like(m)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    ts0 = int(_lib.raw("scn_conv_ts_launch_count")())
    wts0 = int(_lib.raw("scn_conv_wgrad_ts_launch_count")())
    calls = {
        "conv": lambda: _lib.call("scn_conv_fwd_tf32", P(x), C, C, n, P(m), n, K, P(img), None, None, 0, None, 0, P(out), C, C, 0, s),
        "wgrad": lambda: _lib.call("scn_conv_bwd_weight", P(x), C, C, P(m), n, K, P(go), C, C, P(gw), P(gb), 1, s),
        "rulebook": lambda: _lib.call("scn_subm_map", P(lvl.keys), n, P(lvl.tab_keys), P(lvl.tab_vals), lvl.cap, 3, 3, 3, P(m2), s),
    }
    ms = {}
    for name, fn in calls.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        evs = []
        for _ in range(n_iter):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        ms[name] = sum(a.elapsed_time(b) for a, b in evs) / n_iter
    used_ts = (int(_lib.raw("scn_conv_ts_launch_count")()) > ts0, int(_lib.raw("scn_conv_wgrad_ts_launch_count")()) > wts0)
    assert torch.equal(m2, m)
    # algorithmic bytes, SURVEY.md 8d with s = 4 (fp32 storage)
    alg = {"conv": n * C * 4 + n * C * 4 + K * C * C * 4 + 4 * K * n,
           "wgrad": n * C * 4 + n * C * 4 + K * C * C * 4 + 4 * K * n,
           "rulebook": 8 * n + K * n * (8 + 4)}
    return ms, alg, 2.0 * pairs * C * C, n, pairs, used_ts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="tf32", choices=["tf32", "fp32"])
    ap.add_argument("--infer-workers", type=int, default=3,
                    help="host threads (one CUDA stream each) of the sparse inference measurement; 1 = one scene at a time")
    ap.add_argument("--sync-loss", action="store_true",
                    help="e2e: read every step's loss with a blocking .item() right after the step instead of one step late")
    ap.add_argument("--build-ahead", action="store_true",
                    help="build each step's rulebooks between the previous step's forward and backward (the next batch is "
                         "known one step ahead, as with a DataLoader).  Off by default: measured 7.89 vs 7.80 ms -- the "
                         "host is busy, not blocked, so moving its synchronisation points buys nothing")
    ap.add_argument("--no-stage", dest="stage", action="store_false",
                    help="do NOT upload each batch one step ahead on a copy stream (BackboneTrainer.stage).  Staging is on by "
                         "default since the native executor freed the host (e2e 6.4 -> 5.8 ms); every step still uploads "
                         "exactly one batch from pinned host memory inside the timed region")
    ap.add_argument("--prefetch", action="store_true",
                    help="build each step's rulebooks one step ahead on a side stream from a worker thread "
                         "(scn.GeometryPrefetcher).  Off by default: with the native executor it reaches 5.3 ms per step in "
                         "good runs but is not stable (3 runs: 6.5 / 6.4 / 7.0 ms, e2e 6.1 / 5.8 / 20.9 ms -- the worker shares "
                         "the GIL and the launch queue with the training thread); without it 6.22 +- 0.01 ms "
                         "(profiles/r2_e_executor.md)")
    ap.add_argument("--geometry-ahead", default="thread", choices=["off", "inline", "thread"],
                    help="build the following batch's rulebooks one step ahead on a high-priority side stream into a recycled "
                         "arena (the batch is known one step ahead, as with a DataLoader; every step still builds exactly one "
                         "geometry inside the timed loop).  thread: a worker thread takes the builder's row-count round trips; "
                         "inline: the training thread does, after enqueueing the step; off: at the head of its own step")
    ap.add_argument("--geometry-at", default="end", choices=["start", "forward", "end"],
                    help="where in the step the worker thread is handed the next batch")
    ap.set_defaults(stage=True)
    args = ap.parse_args()
    if os.environ.get("SCN_SWITCH_INTERVAL"):      # experiment: GIL hand-off latency between the training thread and --prefetch's worker
        sys.setswitchinterval(float(os.environ["SCN_SWITCH_INTERVAL"]))
    if args.impl == "reference":
        return reference_arm(args)

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback; use --impl reference for the CPU arm)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from sparse_rcnn_b200 import _lib, pipeline, scn
    scn.set_precision(args.precision)
    trainer = pipeline.BackboneTrainer(dev, distributed=world > 1)
    if world > 1:       # identical replicas
        for p in trainer.parameters():
            dist.broadcast(p.data, 0)
    W = max(args.warmup, 3)
    K = args.steps

    # weak scaling: every rank owns its own scene(s); a different scene per step so nothing is cached
    n_distinct = 4
    host = [balanced_inputs(rank, i) for i in range(n_distinct)]
    pinned = [((d[0].pin_memory(), d[1].pin_memory(), d[2], d[3], d[4]), l.pin_memory()) for d, l in host]
    resident = [((d[0].to(dev), d[1].to(dev), d[2], d[3], d[4]), l.to(dev)) for d, l in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(inputs, read_loss):
        # --prefetch: the following batch is known (as with a DataLoader), so its rulebooks are built on a side stream
        # while this step runs; every step's geometry is still built inside the loop, one step ahead
        nxt = (lambda i: inputs[(i + 1) % n_distinct][0]) if args.prefetch else (lambda i: None)
        # the following batch is known (DataLoader): its rulebooks are built between this step's forward and backward
        # (--no-build-ahead: at the head of its own step); every step still builds exactly one geometry and uploads exactly
        # one batch inside the loop (h2d_bytes_per_step)
        trainer.stage_uploads, trainer.build_ahead = args.stage, args.build_ahead
        trainer.build_late = False if (args.prefetch or args.geometry_ahead == "off") else args.geometry_ahead
        trainer.build_late_at = args.geometry_at
        nb = (lambda i: inputs[(i + 1) % n_distinct]) if (args.stage or args.build_ahead or trainer.build_late) else (lambda i: None)
        # every distinct scene twice before the warm-up proper: first-touch costs of a new geometry (caching-allocator growth,
        # dynamic-smem attributes, the recycled arenas of the prefetcher / staging ring) must not land in the timed region
        for i in range(2 * n_distinct):
            trainer.step(*inputs[i % n_distinct])
        for i in range(W):
            trainer.step(*inputs[i % n_distinct], next_data=nxt(i), next_batch=nb(i))
        # Untimed settle loop behind the W warm-up steps: groups of n_distinct steps until two consecutive groups take the same
        # wall time to 3 % (or 10 groups).  The one-step-ahead pipeline needs the host to run ahead of the GPU; right after
        # process start (or after another process has loaded the host) the first dozens of steps can be slower.
        # `settle_steps` in the JSON line says how many were run; the timed region below is unchanged.
        prev, settle = None, 0
        for grp in range(10):
            torch.cuda.synchronize()
            t_g = time.perf_counter()
            for i in range(n_distinct):
                trainer.step(*inputs[(W + settle + i) % n_distinct], next_data=nxt(W + settle + i), next_batch=nb(W + settle + i))
            torch.cuda.synchronize()
            t_g = time.perf_counter() - t_g
            settle += n_distinct
            done = prev is not None and abs(t_g - prev) <= 0.03 * prev
            if world > 1:      # every rank must run the same number of steps (each one holds an allreduce): stop together
                flag = torch.tensor([1.0 if done else 0.0], device=dev)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                done = bool(flag.item() > 0.5)
            if done:
                break
            prev = t_g
        diag.setdefault("settle_steps", []).append(settle)
        # the timed loop starts at scene 0: restart the one-step-ahead hand-over there
        trainer.step(*inputs[(W + settle) % n_distinct], next_data=nxt(n_distinct - 1), next_batch=nb(n_distinct - 1))
        barrier()
        launches0 = _lib.raw("scn_launch_count")()
        mallocs0 = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
        step_wall = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        voxels, t0 = 0, time.perf_counter()
        e0.record()
        # e2e: every step's loss is copied device->host (4 bytes, pinned) and read by the host.  Like a training loop that
        # logs its loss, the host reads step i's value after it has queued step i+1 (--sync-loss: right away, which stalls
        # the host until the whole step has drained and leaves the GPU idle while the next step is being issued)
        slots = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
        evs = [torch.cuda.Event() for _ in range(2)]
        pending, loss_sum = None, 0.0
        for i in range(K):
            loss = trainer.step(*inputs[i % n_distinct], next_data=nxt(i), next_batch=nb(i))
            if read_loss:
                if args.sync_loss:
                    loss_sum += float(loss.item())
                else:
                    slots[i & 1].copy_(loss.reshape(1), non_blocking=True)
                    evs[i & 1].record()
                    if pending is not None:
                        evs[pending].synchronize()
                        loss_sum += float(slots[pending][0])
                    pending = i & 1
            voxels += trainer.last_active
            step_wall.append(time.perf_counter())
        if read_loss and pending is not None:
            evs[pending].synchronize()
            loss_sum += float(slots[pending][0])
        if read_loss and not (loss_sum == loss_sum):
            raise RuntimeError("non-finite loss in the timed region")
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        launches = _lib.raw("scn_launch_count")() - launches0
        # diagnostics of the timed region: cudaMalloc calls of the caching allocator (a steady-state step makes none) and the
        # host-side duration of the slowest step
        diag["device_allocs"] = diag.get("device_allocs", 0) + torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - mallocs0
        gaps = [b - a for a, b in zip([t0] + step_wall[:-1], step_wall)]
        diag.setdefault("slowest_step_host_ms", []).append(round(max(gaps) * 1e3, 3) if gaps else None)
        t = torch.tensor([ms, float(voxels), wall * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            tm = t.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ts = t.clone()
            dist.all_reduce(ts, op=dist.ReduceOp.SUM)
            ms, voxels, wall_ms = float(tm[0]), float(ts[1]), float(tm[2])
        else:
            wall_ms = wall * 1e3
        return max(ms, 1e-9), voxels, launches, wall_ms

    diag = {}
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, vox_dev, launches, _ = timed(resident, read_loss=False)
    ms_e2e, vox_e2e, _, _ = timed(pinned, read_loss=True)
    clocks = sampler.stop() if rank == 0 else None

    roof = roof_w = roof_r = None
    if rank == 0:
        data, _ = resident[0]
        md = scn.Metadata(3)
        scn.ioLayers.InputLayerFunction.apply(3, md, data[2], data[0], data[1], data[3], 4)
        kms, alg, flops, n0, pairs, used_ts = time_kernels(dev, (md, data[2]))
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peaks = json.load(open(pk))
        peak = float(peaks.get("hbm_gbs", 6650.0))
        # dram__bytes_read.sum + dram__bytes_write.sum per launch of these kernels on this layer, from the committed ncu
        # capture (profiles/roofline_traffic.json says which report each number comes from)
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))

        def roofline(key, kernel):
            achieved = alg[key] / (kms[key] * 1e-3) / 1e9
            return {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak,
                    "peak_source": "measured" if peaks else "fallback", "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic.get(key, {}).get("bytes_per_launch"), "ms_per_launch": kms[key],
                    "algorithmic_bytes": alg[key], "l2": "flushed before every launch"}
        roof = roofline("conv", "%s SubM 3^3 32->32 forward, N=%d, %d pairs" % (
            "k_conv_ts<32> (tile-local, A operand in tensor memory)" if used_ts[0] else "k_conv_tc<4>", n0, pairs))
        roof["tflops_useful"] = flops / (kms["conv"] * 1e-3) / 1e12
        roof_w = roofline("wgrad", "%s weight + bias gradient of the same layer" % (
            "k_wgrad_ts<32> + k_wgrad_ts_reduce<32> (tile-local, deterministic)" if used_ts[1] else "k_conv_wgrad_tc<4>"))
        roof_w["tflops_useful"] = flops / (kms["wgrad"] * 1e-3) / 1e12
        roof_r = roofline("rulebook", "k_subm_map 3^3 neighbour map of level 0 (13 N hash probes)")

    # metric part (ii): sparse inference scenes/s (backbone + segmentation + class network + sparse mask network on
    # 256 proposal boxes per scene), every rank runs its own scenes (no collective), pinned host inputs.
    # The RoIs come from RAW region-proposal outputs (30k anchors per scene, pinned host tensors): top-1024 by score,
    # 3-D NMS at 0.5 (scn_nms3d), first 256 survivors -- proposal.ProposalSelector inside SparseInference; only the dense
    # trunk that would produce those scores is not part of the pass.
    from sparse_rcnn_b200.synthetic import make_rpn_outputs
    infer = pipeline.SparseInference(dev)
    rpn = [tuple(t.pin_memory() for t in make_rpn_outputs(d[0], 30000, 256, 7 + i)) for i, (d, _) in enumerate(host)]
    n_inf = max(K, 8)
    n_inf += n_inf % 2
    workers = max(1, args.infer_workers)
    consume = lambda i, res: (int(res["mpn_mask"].shape[0]), res["mpn_class"].argmax(1).cpu(),     # D2H of the class decision
                              len(res["roi_index"][0]))
    seq = lambda n: ([pinned[i % n_distinct][0] for i in range(n)], [rpn[i % n_distinct] for i in range(n)])
    for i in range(max(W, 2 * n_distinct)):              # every distinct scene twice: the caching allocator has to settle
                                                         # on the inference pass's tensor sizes (a cudaMalloc costs ~10 ms)
        infer(pinned[i % n_distinct][0], rpn=rpn[i % n_distinct])

    def timed_inference(nw):
        sc, bx = seq(n_inf)
        for _ in range(2 if nw > 1 else 0):              # the allocator pools are per stream: settle the workers' too
            infer.run_many(sc, rpn=bx, workers=nw, consume=consume)
        barrier()
        i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.raw("scn_launch_count")()
        i0.record()
        out = infer.run_many(sc, rpn=bx, workers=nw, consume=consume)
        i1.record()
        barrier()
        ms = i0.elapsed_time(i1)
        if world > 1:
            tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            ms = float(tmax[0])
        return ms, sum(o[0] for o in out), _lib.raw("scn_launch_count")() - l0, sum(o[2] for o in out)

    ser_ms, mask_pts, inf_launches, n_rois = timed_inference(1)
    inf_ms = ser_ms
    if workers > 1:
        inf_ms, mask_pts, inf_launches, n_rois = timed_inference(workers)
    inference = {"scenes_per_sec": world * n_inf / (inf_ms * 1e-3), "ms_per_scene": inf_ms / n_inf, "boxes_per_scene": n_rois / n_inf,
                 "proposals_per_scene": "30000 anchors -> top 1024 -> NMS 0.5 -> <= 256",
                 "mask_points_per_scene": mask_pts // n_inf, "scn_launches_per_scene": int(inf_launches // n_inf),
                 "host_threads": workers, "ms_per_scene_one_thread": ser_ms / n_inf,
                 "scope": "backbone + segmentation + proposal selection (top-k, 3-D NMS kernel) + class net + mask net from raw RPN "
                          "scores/boxes in pinned host memory; the dense RPN trunk that would emit those scores is not built; "
                          "scenes are independent, each host thread drives its own stream (SparseInference.run_many)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, t, info = run_cpu_oracle(1, 1)
        cpu = dict(info, value=v, unit=UNIT)

    if rank == 0:
        P = len(host[0][0][0])
        h2d = P * 4 * 8 + P * 6 * 4 + P * 8
        line = {
            "metric": METRIC, "value": vox_dev / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32" if args.precision == "tf32" else "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "parallelism": "dp%d (one scene per rank, NCCL gradient allreduce)" % world if world > 1 else "single GPU",
                       "l2": "a different scene every step (4 distinct, inputs+activations > L2 over a step)",
                       "precision": args.precision},
            "e2e": {"value": vox_e2e / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "loss_read": "blocking .item() per step" if args.sync_loss else
                    "every step, async copy into pinned memory, read by the host one step late",
                    "ms_per_step": ms_e2e / K},
            "gpu_launches": int(launches), "timed_region": diag, "clocks": clocks, "roofline": roof, "roofline_wgrad": roof_w,
            "roofline_rulebook": roof_r, "inference": inference,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
