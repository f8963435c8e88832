"""Multi-GPU plumbing for the two ways the path shards.

* Inference: scenes are independent (own hash grids) -> scene i goes to rank i mod world,
  NO collective (`shard_scenes`).
* Training: data parallel over the batch with ONE collective, the gradient allreduce.  The
  reference has no distributed code at all (single process, scannet_config/run.py:498-499);
  the hook point is right before `optimizer.step()` (ndsis/training/training.py:458-460).
  Gradients live as views into two flat fp32 buckets (decoder-side parameters finish first in
  backward), each bucket is all-reduced asynchronously as soon as its last gradient has been
  accumulated, so the first allreduce overlaps the rest of backward.  NVSwitch gives every GPU
  full bandwidth to every peer, so buckets are sized for launch latency/overlap, not link count.
"""
import torch
import torch.distributed as dist


def shard_scenes(n_scenes, rank, world):
    """Indices of the scenes this rank owns (round robin; weak scaling keeps len() fixed)."""
    return list(range(rank, n_scenes, world))


class GradientBuckets:
    """Flat gradient buckets with overlapped asynchronous allreduce (mean)."""

    def __init__(self, params, n_buckets=2, process_group=None, groups=None):
        """params: parameters in registration order, cut into `n_buckets` equal-sized buckets in reverse order -- or
        `groups`: explicit buckets, given in the order in which the backward pass completes them (the native U-Net executor
        finishes decoder / coarse encoder levels / fine encoder levels in three calls and fires each group's hooks at once,
        so buckets cut along those boundaries start their allreduce as early as possible)."""
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        # NCCL averages inside the collective; gloo (CPU tests) has no AVG: sum, then divide
        self.average_in_collective = dist.is_initialized() and dist.get_backend(process_group) == "nccl"
        if groups is not None:
            self.buckets = [[p for p in g if p.requires_grad] for g in groups]
            self.buckets = [b for b in self.buckets if b]
            seen = set()
            for b in self.buckets:
                for p in b:
                    if id(p) in seen:
                        raise ValueError("a parameter appears in two gradient buckets")
                    seen.add(id(p))
            missing = [p for p in params if p.requires_grad and id(p) not in seen]
            if missing:
                self.buckets.append(missing)
        else:
            params = [p for p in params if p.requires_grad]
            # reverse registration order ~ order in which backward produces gradients
            order = list(reversed(params))
            total = sum(p.numel() for p in order)
            target = (total + n_buckets - 1) // n_buckets
            self.buckets, cur, cur_n = [], [], 0
            for p in order:
                cur.append(p)
                cur_n += p.numel()
                if cur_n >= target and len(self.buckets) < n_buckets - 1:
                    self.buckets.append(cur)
                    cur, cur_n = [], 0
            if cur:
                self.buckets.append(cur)
        self.flats, self._pending, self._handles = [], [], []
        self._fired = set()
        self._launched = []
        self.offsets = []
        for bi, bucket in enumerate(self.buckets):
            # every view starts on a 16-byte boundary (vector reductions / loads in the kernels); the pad elements stay 0
            offs, n = [], 0
            for p in bucket:
                offs.append(n)
                n += (p.numel() + 3) // 4 * 4
            self.offsets.append(offs)
            flat = torch.zeros(n, dtype=bucket[0].dtype, device=bucket[0].device)
            for p, off in zip(bucket, offs):
                p.grad = flat[off:off + p.numel()].view_as(p)      # autograd accumulates in place
                hook = self._make_hook(bi)
                p.register_post_accumulate_grad_hook(hook)
                # the scn autograd Functions accumulate parameter gradients straight into these views and call the hook
                # themselves (no zero-fill + AccumulateGrad add per parameter); torch.autograd.grad() on such parameters
                # is not supported while the buckets exist
                p._scn_grad_hook = hook
            self.flats.append(flat)
            self._pending.append(len(bucket))
            self._launched.append(False)
        self.enabled = True

    def _make_hook(self, bi):
        def hook(param):
            if not self.enabled or id(param) in self._fired:
                return
            self._fired.add(id(param))
            self._pending[bi] -= 1
            if self._pending[bi] == 0:
                self._launch_ready()
        return hook

    def _launch_ready(self):
        """Collectives are issued in FIXED bucket order on every rank (bucket i only after buckets 0..i-1), whatever the
        order in which the buckets complete locally: a rank whose class / mask head saw no RoI completes its buckets in a
        different order than its peers, and mismatched all_reduce sequences hang or corrupt (same rule as torch DDP)."""
        for bi in range(len(self.buckets)):
            if self._launched[bi]:
                continue
            if self._pending[bi] > 0:
                break
            self._launch(bi)

    def _launch(self, bi):
        self._launched[bi] = True
        if self.world > 1:
            op = dist.ReduceOp.AVG if self.average_in_collective else dist.ReduceOp.SUM
            self._handles.append(dist.all_reduce(self.flats[bi], op=op, group=self.group, async_op=True))

    def flatten_parameters(self):
        """Make every parameter a view of one flat buffer per bucket (same order as its gradient view) and return the flat
        buffers as leaf Parameters whose .grad are the gradient buckets.  An optimizer over THESE (two tensors instead of
        ~150) does the same element-wise update with none of the per-parameter host work, which matters because the step is
        host bound.  Modules keep their Parameter objects; `p.data` now aliases the flat storage."""
        flat_params = []
        for bucket, offs, gflat in zip(self.buckets, self.offsets, self.flats):
            data = torch.zeros(gflat.numel(), dtype=bucket[0].dtype, device=bucket[0].device)
            for p, off in zip(bucket, offs):
                view = data[off:off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
            fp = torch.nn.Parameter(data)
            for p in bucket:
                # the optimizer mutates `fp`, so fp's version counter is what tells a cached derivative of p (the packed
                # tensor-core weight images, scn/functions.py:_wver) that p has changed; p._version itself never moves
                p._scn_flat = fp
            fp.grad = gflat
            flat_params.append(fp)
        self.flat_params = flat_params
        return flat_params

    def zero(self):
        """Zero the flat buckets (instead of optimizer.zero_grad(set_to_none=True), which would
        detach the gradient views)."""
        for f in self.flats:
            f.zero_()
        self._pending = [len(b) for b in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._handles = []
        self._fired = set()

    def finish(self):
        """Call after backward, before optimizer.step(): launches buckets whose hooks did not all
        fire (unused parameters), waits for the collectives and turns the sums into means."""
        for bi in range(len(self.buckets)):      # unused parameters: whatever is left, still in bucket order
            self._pending[bi] = 0
            if not self._launched[bi]:
                self._launch(bi)
        for h in self._handles:
            h.wait()      # NCCL: the current stream waits for the collective's stream, the host does not block
        self._handles = []
        if self.world > 1 and not self.average_in_collective:
            for f in self.flats:
                f.div_(self.world)

    def nbytes(self):
        return sum(f.numel() * f.element_size() for f in self.flats)
