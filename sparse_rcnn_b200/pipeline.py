"""Public entry points of the hot path: a training step of the sparse backbone and the sparse
inference pass (backbone + segmentation + class network + sparse mask network).

These are the calls bench.py and __graft_entry__.smoke() time / exercise.  Inputs follow the
reference's `collate_fn` 5-tuple (ndsis/data/data.py:88-115): coords stay int64 on the HOST,
features may be host (pinned) or device tensors; the host->device copies happen inside.
"""
import torch
import torch.nn as nn

from . import executor, networks, proposal, roi, scn
from .parallel import GradientBuckets
from .scn.metadata import stage_to_device, take_staged


def _dev(t, device):
    d = take_staged(t)
    return d if d is not None else t.to(device, non_blocking=True)


def _to_device(data, device):
    coords, feats, size, bs, splits = data
    return coords, _dev(feats, device), size, bs, splits


class BackboneTrainer(nn.Module):
    """fwd + point-wise cross-entropy on the segmentation head + bwd (+ gradient allreduce) + Adam,
    i.e. the sparse part of ndsis/training/training.py:428-460 for the shipped sparse U-Net."""

    def __init__(self, device, in_channels=6, num_classes=20, lr=1e-3, distributed=False, seed=0):
        super().__init__()
        torch.manual_seed(seed)
        self.device = device
        self.backbone = networks.FeatureExtractor(scn, input_channels=in_channels)
        self.seg = networks.SegmentationNetwork(scn, 32, num_classes)
        self.to(device)
        # gradient buckets cut where the backward pass completes them: segmentation head + decoder, coarse encoder levels
        # (most of the parameters), fine encoder levels (most of the time) -- the executor fires each group's hooks at once
        enc = list(self.backbone.main_network)
        split = executor.ENCODER_PHASE_SPLIT
        groups = [list(self.seg.parameters()) + list(self.backbone.unet.parameters()),
                  [p for lev in enc[split:] for p in lev.parameters()],
                  [p for lev in enc[:split] for p in lev.parameters()]]
        self.buckets = GradientBuckets(list(self.parameters()), groups=groups)
        self.buckets.enabled = True
        # torch.optim.Adam as in the reference (ndsis/training/training.py:386); the fused implementation is the same
        # update in one multi-tensor kernel (the foreach path costs ~1.2 ms of host time per step, and the step is host bound)
        # ... over the two flat parameter buffers (every parameter is a view into them): same element-wise update, no
        # per-parameter host work (torch's Adam spends ~0.4 ms per step walking 150 parameters)
        self.optimizer = torch.optim.Adam(self.buckets.flatten_parameters(), lr=lr, fused=True)
        self.distributed = distributed
        self._weights = [p for p in self.parameters() if p.dim() >= 2]
        self.prefetcher = None
        self.build_ahead = False       # next_batch: build its rulebooks between this step's forward and backward (no gain)
        self.stage_uploads = False     # next_batch: also issue its host->device copies now (measured: no gain)
        # next_batch: build its rulebooks one step ahead on a high-priority side stream, into a recycled arena
        # (scn.GeometryPrefetcher).  "thread": a worker thread issues the builder's kernels and takes its row-count round
        # trips (blocked ~3 ms per step while the GPU is busy with this step); "inline": the training thread does, after it
        # has enqueued the whole step.  build_late_at: where in the step the build is started (start | forward | end).
        self.build_late = False
        self.build_late_at = "end"

    def prefetch(self, data, threaded=True):
        """Start building the geometry (voxel hash, level pyramid, neighbour maps) of an upcoming batch on a side stream;
        the step() that later receives the same coords tensor picks it up (scn.GeometryPrefetcher).  threaded=False: built
        by the calling thread, inline (see GeometryPrefetcher)."""
        if self.prefetcher is None:
            self.prefetcher = scn.GeometryPrefetcher(self.device, n_levels=len(self.backbone.channels) - 1, threaded=threaded,
                                                     book_channels=list(self.backbone.channels))
            self.backbone.input_stage.prefetcher = self.prefetcher
        coords, _, size, bs = data[:4]
        self.prefetcher.submit(coords.long(), torch.as_tensor(size, dtype=torch.long), bs)

    def stage(self, data, labels=None):
        """Start the host->device copies (coords, features, labels; pinned host tensors) of an upcoming batch on the copy
        stream; the step() that later receives the same tensor objects uses the staged copies."""
        stage_to_device([data[0], data[1], labels], self.device)

    def step(self, data, labels, next_data=None, next_batch=None):
        """data: collate_fn 5-tuple, labels int64 [P] (host or device).  Returns the loss (device scalar).
        next_batch = (data, labels) of the following step, if already known (as from a DataLoader): its rulebooks are built
        between this step's forward and backward (`build_ahead`), optionally its host->device copies are issued now
        (`stage_uploads`).  next_data: build its rulebooks in a worker thread instead (scn.GeometryPrefetcher, opt-in)."""
        if next_data is not None:
            self.prefetch(next_data)
        ahead = self.build_late if next_batch is not None else False
        at = "end" if ahead == "inline" else self.build_late_at
        if ahead and at == "start":
            self.prefetch(next_batch[0], threaded=True)
        data = _to_device(data, self.device)
        labels = _dev(labels, self.device)
        if next_batch is not None and self.stage_uploads:
            self.stage(*next_batch)
        self.buckets.zero()
        scn.functions.pack_all(self._weights)      # one launch: every packed weight image the optimizer made stale
        out = self.backbone(data)
        logits = self.seg(out[5])
        loss = scn.functions.cross_entropy(logits, labels)      # nn.CrossEntropyLoss semantics (loss.py:95-97)
        if ahead and at == "forward":
            self.prefetch(next_batch[0], threaded=True)
        if next_batch is not None and self.build_ahead:
            # the following batch's rulebooks, between this step's forward and backward: their host round trips wait
            # while the GPU drains the forward instead of idling it at the head of the next step
            self.backbone.input_stage.build_ahead(next_batch[0], len(self.backbone.channels) - 1, self.device)
        loss.backward()
        self.buckets.finish()
        self.optimizer.step()
        scn.functions.weights_changed()            # packed TF32 weight images are stale now (re-packed by the next pack_all)
        if ahead and at == "end":
            # the whole step is enqueued: the following batch's rulebooks now, on the high-priority side stream -- the six
            # row-count round trips wait for that stream only, while the GPU still has this step's backward to run
            self.prefetch(next_batch[0], threaded=ahead != "inline")
        self.last_active = out[4][0].features.shape[0]
        return loss.detach()


class SparseInference(nn.Module):
    """Sparse inference pass on given proposal boxes (the dense RPN trunk / NMS that would produce
    them is out of scope): feature extractor -> segmentation logits per point,
    class logits per box, mask logits per (box, point)."""

    def __init__(self, device, in_channels=6, num_seg_classes=20, num_classes=18, seed=0, num_keep_pre_nms=1024,
                 num_keep_post_nms=256, thresh_nms=0.5):
        super().__init__()
        torch.manual_seed(seed)
        self.device = device
        cut = lambda **kw: roi.SparseRoiCut(scn, **kw)
        self.backbone = networks.FeatureExtractor(scn, input_channels=in_channels)
        self.seg = networks.SegmentationNetwork(scn, 32, num_seg_classes)
        self.class_network = networks.ClassNetwork(scn, cut, input_channels=64, stride=4, num_classes=num_classes)
        self.mask_network = networks.SparseMaskNetwork(scn, cut, input_channels=in_channels,
                                                       channel_list=(32, num_classes))
        # proposal selection of the shipped configuration (run.py: 1024 highest scores -> NMS 0.5 -> 256 RoIs)
        self.roi_selector = proposal.ProposalSelector(num_keep_pre_nms, num_keep_post_nms, thresh_nms)
        self.to(device)
        self.eval()

    @torch.no_grad()
    def forward(self, data, boxes=None, rpn=None):
        """boxes: proposal boxes per sample (list of [n_i, 2, 3], voxel units) -- or rpn = (score [B, A], bbox [B, A, 2, 3]):
        the raw proposal scores / boxes a region-proposal head would emit (model.py:990-1010 hands them to the RoiSelector),
        reduced on the device by `proposal.ProposalSelector` (top-k, 3-D NMS kernel, first num_keep_post_nms survivors)."""
        data = _to_device(data, self.device)
        roi.clear_key_cache()
        selected = None
        if boxes is None:
            if rpn is None:
                raise ValueError("SparseInference needs proposal boxes or raw RPN outputs")
            score, bbox = (t.to(self.device, non_blocking=True) for t in rpn)
            selected = self.roi_selector(score, bbox)              # (scores, boxes, indices) lists over the batch
            boxes = selected[1]
        scene_size, batch_size, _, class_map, inter, unet = self.backbone(data)
        roi.register_keys(data[0], inter[0].metadata.point_keys)     # packed once per forward
        seg = self.seg(unet)
        cls, cls_sel = self.class_network(class_map, boxes)
        mask, mask_sel = self.mask_network(data, unet, boxes)
        out = dict(segmentation=seg, mpn_class=cls, mpn_mask=mask, class_selection=cls_sel, mask_selection=mask_sel,
                   n_active=inter[0].features.shape[0])
        if selected is not None:
            out["roi_score"], out["roi_bbox"], out["roi_index"] = selected
        return out

    def _geometry_prefetcher(self):
        pf = getattr(self, "_prefetcher", None)
        if pf is None:
            ch = list(self.backbone.channels)
            pf = self._prefetcher = scn.GeometryPrefetcher(self.device, n_levels=len(ch) - 1, book_channels=ch)
            self.backbone.input_stage.prefetcher = pf
        return pf

    def _submit_geometry(self, pf, scene):
        coords, _, size, bs = scene[:4]
        pf.submit(coords.long(), torch.as_tensor(size, dtype=torch.long), bs, again=True)

    def run_many(self, scenes, boxes=None, workers=2, consume=None, rpn=None, geometry_ahead=False):
        """Inference over independent scenes from `workers` host threads, one CUDA stream each (scene i -> worker
        i mod workers).  Why: one scene's pass is a ping-pong between host and GPU -- 16 host reads of row counts (every
        level of three pyramids and two crops) during which the host waits for the GPU, each followed by a stretch in which
        the GPU waits for the host to issue again (`scripts/host_profile_infer.py`: ~3 of 11 ms per scene blocked in
        `.item()`).  Those reads release the GIL, so a second thread issues its scene's kernels meanwhile.  Scenes are
        independent (no collective), the library keeps its split-mode workspace per stream, crop key caches are
        per thread, packed weight images are read-only here.
        `consume(i, result)` runs on the worker's stream right after scene i (e.g. the device->host read of a decision) and
        its return value replaces the result; returned tensors are `record_stream`-ed for the caller's stream."""
        import threading
        n = len(scenes)
        # geometry_ahead: the BACKBONE geometry of the scenes ahead (voxel hash, six-level pyramid, maps, tile books: 6 of a
        # scene's 16 host round trips) is built by the prefetcher's thread on its own stream into recycled arenas; every
        # worker hands in scene i + workers before it runs scene i.  Same geometry, same results.  OFF by default -- measured
        # slower (profiles/r2_h section 8: 8.4 -> 9.5 ms per scene with one worker, 7.2 -> 7.6-7.9 with three): the pass is
        # bound by Python under the GIL, which a further thread does not shorten, and the builder's kernels then queue
        # behind other scenes' persistent convolution CTAs.
        pf = self._geometry_prefetcher() if (geometry_ahead and n > 1) else None
        if workers <= 1 or n <= 1:
            out = []
            try:
                if pf is not None:
                    self._submit_geometry(pf, scenes[0])
                for i in range(n):
                    if pf is not None and i + 1 < n:
                        self._submit_geometry(pf, scenes[i + 1])
                    r = self(scenes[i], None if boxes is None else boxes[i], rpn=None if rpn is None else rpn[i])
                    out.append(consume(i, r) if consume is not None else r)
            finally:
                if pf is not None:
                    pf.drain()
            return out
        workers = min(workers, n)
        if pf is not None:
            for i in range(workers):
                self._submit_geometry(pf, scenes[i])
        streams = getattr(self, "_streams", None)
        if streams is None or len(streams) < workers:
            streams = self._streams = [torch.cuda.Stream(device=self.device) for _ in range(workers)]
        caller = torch.cuda.current_stream(self.device)
        # every forward weight image exists and is packed on the caller's stream before a worker can look at it
        scn.functions.prepack_forward(self)
        ready = torch.cuda.Event()
        ready.record(caller)
        done = [torch.cuda.Event() for _ in range(workers)]
        results, errors = [None] * n, []

        def work(k):
            try:
                torch.cuda.set_device(self.device)
                with torch.cuda.stream(streams[k]):
                    streams[k].wait_event(ready)      # weights / packed images written on the caller's stream
                    for i in range(k, n, workers):
                        if pf is not None and i + workers < n:
                            self._submit_geometry(pf, scenes[i + workers])
                        r = self(scenes[i], None if boxes is None else boxes[i], rpn=None if rpn is None else rpn[i])
                        results[i] = consume(i, r) if consume is not None else r
                    done[k].record(streams[k])
            except BaseException as e:      # surfaced in the calling thread
                errors.append(e)

        threads = [threading.Thread(target=work, args=(k,), name="scn-infer-%d" % k) for k in range(workers)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if pf is not None:
            pf.drain()
        if errors:
            raise errors[0]
        for d in done:
            caller.wait_event(d)

        def mark(o):
            if isinstance(o, torch.Tensor) and o.is_cuda:
                o.record_stream(caller)
            elif isinstance(o, dict):
                for v in o.values():
                    mark(v)
            elif isinstance(o, (list, tuple)):
                for v in o:
                    mark(v)
        mark(results)
        return results
