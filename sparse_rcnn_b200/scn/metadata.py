"""Device-resident Metadata: the B200 replacement of SparseConvNet's `Metadata<d>`.

Reference usage: one fresh `scn.Metadata(dim)` per input layer / crop
(/root/reference ndsis/modules/custom_operations.py:70, roi_select_sparse.py:77,115),
shared by every SparseConvNetTensor derived from it.  Upstream keeps per-scale CPU
hash maps and std::vector rulebooks that are re-uploaded on every convolution call;
here every scale is a `Level` living in HBM: packed row keys, an open-addressing
table key->row, and cached output-stationary neighbour maps.  All buffers are torch
tensors (caching allocator); the C ABI only sees raw pointers.
"""
import os

import numpy as np
import threading

import torch

from .. import _lib


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    """cudaStream_t of torch's current stream.  torch.cuda.current_stream() costs ~10 us of Python per call
    (device-index resolution, env lookups); the raw accessor is ~0.3 us and this is called for every kernel."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


_KEY_ATTR = "_scn_size_key"

_order = {"rows": "morton"}


def set_row_order(order):
    """'morton' (default): rows of every level follow the Morton curve of (b, x, y, z) -- spatially coherent tiles, the
    layout the tile-local convolution kernel needs (csrc/sort.cu, conv_ts.cu).  'first': SparseConvNet's numbering, level-0
    rows in order of first appearance in the point list (Metadata/InputLayer.h).  Either way rows are grouped by ascending
    sample index (roi_select_sparse.py:106-107) and every result is identical after the canonical (b, x, y, z) sort
    (SURVEY 8c).  Mode-0 input layers (rows = input order by contract) are never reordered."""
    if order not in ("morton", "first"):
        raise ValueError("row order must be 'morton' or 'first'")
    _order["rows"] = order


def get_row_order():
    return _order["rows"]


def size_key(size):
    """Hashable key of a spatial size; cached on the tensor object (the same size tensor travels through a level)."""
    if isinstance(size, tuple):
        return size
    if isinstance(size, torch.Tensor):
        k = getattr(size, _KEY_ATTR, None)
        if k is None:
            k = tuple(int(s) for s in size.reshape(-1).tolist())
            try:
                setattr(size, _KEY_ATTR, k)
            except AttributeError:
                pass
        return k
    return tuple(int(s) for s in np.asarray(size).reshape(-1))


def _triple(v):
    if isinstance(v, tuple) and len(v) == 3:
        return v
    if isinstance(v, int):
        return (v, v, v)
    return tuple(int(a) for a in np.broadcast_to(np.asarray(v), (3,)))


# ---------------------------------------------------------------- buffers of the rulebook builder
# Geometry built ahead on a side stream (GeometryPrefetcher) must not go through the caching allocator: blocks allocated under
# the side stream and used under the training stream need record_stream, which delays their reuse, and with a different scene
# every step the pools never settle -- measured 4-13 cudaMalloc calls inside 20 timed steps and single steps of 10-100 ms.
# While a thread has an arena installed, every buffer of the builder is a 256-byte aligned slice of that one pre-allocated
# block; without one (the default) buffers come from torch as before.
_arena_tls = threading.local()


class Arena:
    def __init__(self, nbytes, device):
        self.buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        self.off = 0
        self.event = None      # recorded by the last user when it let go of the arena

    def take(self, shape, dtype):
        if isinstance(shape, int):
            shape = (shape,)
        n = 1
        for d in shape:
            n *= int(d)
        nbytes = n * dtype.itemsize
        start = (self.off + 255) // 256 * 256
        if start + nbytes > self.buf.numel():
            return None
        self.off = start + nbytes
        return self.buf[start:start + nbytes].view(dtype).view(shape)


def _new(shape, dtype, device):
    a = getattr(_arena_tls, "arena", None)
    if a is not None:
        t = a.take(shape, dtype)
        if t is not None:
            return t
    return torch.empty(shape, dtype=dtype, device=device)


def exclusive_scan(x):
    """int32 [n] -> int32 [n+1] exclusive prefix sum (own kernel, no CUB)."""
    n = x.numel()
    out = _new(n + 1, torch.int32, x.device)
    tmp = _new(int(_lib.raw("scn_scan_tmp_elems")(n)), torch.int32, x.device)
    _lib.call("scn_exclusive_scan", _ptr(x), _ptr(out), n, _ptr(tmp), _stream())
    return out


class Level:
    """One spatial scale: rows <-> voxel coordinates."""

    def __init__(self, keys, tab_keys, tab_vals, cap, n):
        self.keys = keys              # int64 storage of uint64 packed (b,x,y,z), [n]
        self.tab_keys = tab_keys
        self.tab_vals = tab_vals
        self.cap = cap
        self.n = n
        self.subm = {}                # filter triple -> int32 [K, n] (None for 1x1x1 = identity)
        self._locations = None
        self._batch_ptr = {}
        self.coherent = False         # rows follow the Morton curve (set by the builders): tile books are worth building
        self._books = {}              # map data_ptr -> tile book buffer (attached in the library, detached in __del__)

    def __del__(self):
        try:
            for ptr, book in self._books.items():      # only if the registry still holds THIS book (arena addresses recur)
                _lib.call("scn_tile_book_detach_if", ptr, _ptr(book))
        except Exception:      # interpreter shutdown
            pass

    def locations(self):
        """int64 CPU [n, 4] (x,y,z,b), row aligned (reference contract: custom_operations.py:26-32)."""
        if self._locations is None:
            out = torch.empty((self.n, 4), dtype=torch.int64, device=self.keys.device)
            _lib.call("scn_unpack_keys", _ptr(self.keys), self.n, _ptr(out), _stream())
            self._locations = out.cpu()
        return self._locations

    def batch_ptr(self, n_seg):
        """int32 [n_seg+1]: first row of every sample (rows are batch-sorted)."""
        if n_seg not in self._batch_ptr:
            p = torch.empty(n_seg + 1, dtype=torch.int32, device=self.keys.device)
            _lib.call("scn_batch_offsets", _ptr(self.keys), self.n, n_seg, _ptr(p), _stream())
            self._batch_ptr[n_seg] = p
        return self._batch_ptr[n_seg]

    def subm_map(self, filter_size, tile_book=True):
        """Neighbour map [K, n] of an odd filter.  tile_book=True also builds the tile book of a 3^3 map right away (direct
        users of the C ABI); the module / executor paths pass False and call `ensure_tile_book(channels)` when a layer that
        the tile-local kernels can take actually runs on this level (a 22-channel mask-network level never needs one)."""
        f = _triple(filter_size)
        if f not in self.subm:
            if f == (1, 1, 1):
                self.subm[f] = None
            else:
                K = f[0] * f[1] * f[2]
                m = _new((K, self.n), torch.int32, self.keys.device)
                _lib.call("scn_subm_map", _ptr(self.keys), self.n, _ptr(self.tab_keys), _ptr(self.tab_vals),
                          self.cap, f[0], f[1], f[2], _ptr(m), _stream())
                self.subm[f] = m
        if (tile_book or _EAGER_BOOKS) and f == (3, 3, 3):
            self.ensure_tile_book()
        return self.subm[f]

    def ensure_tile_book(self, channels=None):
        """Tile book of the 3^3 map for the tile-local kernels (csrc/conv_ts.cu, conv_wgrad_ts.cu): halo lists + 16-bit local
        maps, built once per level -- only for Morton-ordered levels with enough tiles, and (when the caller names its layer
        width) only for widths those kernels implement."""
        m = self.subm.get((3, 3, 3))
        if m is None or not self.coherent or self.n < TILE_BOOK_MIN_ROWS or m.data_ptr() in self._books:
            return
        if channels is not None and channels not in TILE_LOCAL_CHANNELS:
            return
        book = _new(int(_lib.raw("scn_tile_book_bytes")(self.n)), torch.uint8, m.device)
        _lib.call("scn_tile_book_build", _ptr(m), self.n, 27, _ptr(book), _stream())
        _lib.call("scn_tile_book_attach", _ptr(m), _ptr(book), self.n)
        self._books[m.data_ptr()] = book


TILE_BOOK_MIN_ROWS = 128 * 148 * 2      # levels with fewer tiles run conv_tc.cu's cluster-split mode anyway
TILE_LOCAL_CHANNELS = (16, 32, 48, 64)    # C -> C layers csrc/conv_ts.cu implements
_EAGER_BOOKS = os.environ.get("SCN_TILE_BOOK_EAGER", "0") == "1"      # A/B switch: a book for every large Morton level


def build_level(keys, morton_bits=None):
    """keys int64[P] (packed) -> (Level, point_row int32 [P]).  Rows are numbered by first appearance (SparseConvNet
    InputLayer.h) over the point list -- or, with morton_bits = (coordinate bits, batch bits), over the point list
    sorted by (b, Morton(x, y, z)), i.e. in Morton order of the voxels.  One host sync (the active count)."""
    dev = keys.device
    P = keys.numel()
    perm = None
    if morton_bits is not None and P > 1:
        ws = _new(int(_lib.raw("scn_morton_order_ws_bytes")(P)), torch.uint8, dev)
        perm = _new(P, torch.int32, dev)
        skeys = _new(P, torch.int64, dev)
        _lib.call("scn_morton_order", _ptr(keys), P, int(morton_bits[0]), int(morton_bits[1]), _ptr(perm), _ptr(skeys),
                  _ptr(ws), _stream())
        keys = skeys
    cap = 64
    while cap < 2 * P:
        cap <<= 1
    tab_keys = _new(cap, torch.int64, dev)
    tab_vals = _new(cap, torch.int32, dev)
    s = _stream()
    first = _new(max(P, 1), torch.int32, dev)
    rank = _new(P + 1, torch.int32, dev)
    tmp = _new(int(_lib.raw("scn_scan_tmp_elems")(P)), torch.int32, dev)
    _lib.call("scn_level_count", _ptr(keys), P, _ptr(tab_keys), _ptr(tab_vals), cap, _ptr(first), _ptr(rank), _ptr(tmp), s)
    n = int(rank[P].item())
    point_row = _new(P, torch.int32, dev)
    row_keys = _new(n, torch.int64, dev)
    _lib.call("scn_level_finish", _ptr(keys), P, _ptr(tab_keys), _ptr(tab_vals), cap, _ptr(rank), _ptr(point_row),
              _ptr(row_keys), s)
    if perm is not None:      # back to the caller's point order
        sorted_rows, point_row = point_row, _new(P, torch.int32, dev)
        _lib.call("scn_scatter_i32", _ptr(sorted_rows), _ptr(perm), P, _ptr(point_row), s)
    level = Level(row_keys, tab_keys, tab_vals, cap, n)
    level.coherent = perm is not None
    return level, point_row


class Strided:
    """Rulebook of a (filter == stride) convolution between two levels."""

    def __init__(self, out_key, cmap, dmap, parent_row, K):
        self.out_key, self.cmap, self.dmap, self.parent_row, self.K = out_key, cmap, dmap, parent_row, K
        self.out_size = torch.tensor(out_key, dtype=torch.long)      # one size tensor per level (carries its key)
        setattr(self.out_size, _KEY_ATTR, tuple(out_key))


class Metadata:
    def __init__(self, dimension=3):
        if dimension != 3:
            raise RuntimeError("sparse_rcnn_b200 implements the 3-D path only (ndsis uses num_dims=3)")
        self.dimension = dimension
        self.levels = {}
        self.strided = {}
        self.point_row = None
        self.row_ptr = None
        self.row_pts = None
        self.n_points = 0
        self.n_samples = 0
        self.mode = None
        self.input_size = None
        self._prebuilt_for = None

    # ---------------------------------------------------------------- input layer rule
    def set_input(self, spatial_size, coords, batch_size, mode, device):
        """coords: int64 [P, 3|4] on CPU or device, or already-packed keys (int64 [P], device) when
        `coords.dim() == 1` (device-side crop path).  Returns the number of active rows."""
        if self._prebuilt_for is not None:
            # geometry built ahead of time by a GeometryPrefetcher for exactly this coords tensor
            if self._prebuilt_for is not coords or mode != self.mode or size_key(spatial_size) != self.input_size:
                raise RuntimeError("prefetched Metadata used with a different input than it was built for")
            return self.levels[self.input_size].n
        s = _stream()
        if coords.dim() == 1:
            keys = coords
            # device-side crop path: batch ids are box ids < batch_size (no host sync needed)
            max_b = int(batch_size) - 1 if int(batch_size) > 0 else None
        else:
            if coords.dtype != torch.int64:
                coords = coords.long()
            P, ncol = coords.shape
            if ncol not in (3, 4):
                raise RuntimeError("coords must be [P, 3] or [P, 4] (x,y,z[,b])")
            max_b = 0
            if ncol == 4 and P:
                max_b = int(coords[:, 3].max())          # CPU in the reference contract => no device sync
            cdev = take_staged(coords)             # uploaded ahead of time on the copy stream (stage_to_device)?
            if cdev is None:
                cdev = coords.to(device, non_blocking=True)
            cdev = cdev.contiguous()
            keys = _new(P, torch.int64, device)
            err = _new(1, torch.int32, device).zero_()
            _lib.call("scn_pack_coords", _ptr(cdev), P, ncol, _ptr(keys), _ptr(err), s)
            self._err = err
        morton_bits = None
        if mode != 0 and _order["rows"] == "morton":
            smax = max(size_key(spatial_size))
            nb = max(int(batch_size) - 1, max_b if max_b is not None else 65535, 0)      # unknown sample count: all 16 bits
            morton_bits = (min(max(int(smax) - 1, 1).bit_length(), 16), nb.bit_length())
        level, point_row = build_level(keys, morton_bits)
        self.row_order = "morton" if morton_bits is not None else "first"
        if coords.dim() != 1 and int(self._err.item()):
            raise RuntimeError("InputLayer: coordinate outside [0, 65534]")
        if max_b is None:
            max_b = int((keys.max().item() >> 48) & 0xFFFF) if keys.numel() else 0
        P = keys.numel()
        if mode == 0 and level.n != P:
            raise RuntimeError("InputLayer mode 0 requires unique coordinates (%d points, %d voxels)" % (P, level.n))
        self.input_size = size_key(spatial_size)
        self.levels[self.input_size] = level
        self.point_row, self.n_points, self.mode = point_row, P, mode
        self.point_keys = keys
        self.n_samples = max(int(batch_size), max_b + 1, 1)
        if mode != 0:
            n = level.n
            cnt = _new(max(n, 1), torch.int32, device)
            self.row_ptr = _new(n + 1, torch.int32, device)
            self.row_pts = _new(P, torch.int32, device)
            tmp = _new(int(_lib.raw("scn_scan_tmp_elems")(n)), torch.int32, device)
            _lib.call("scn_input_rule", _ptr(point_row), P, n, _ptr(cnt), _ptr(self.row_ptr), _ptr(self.row_pts), _ptr(tmp), s)
        return level.n

    def level(self, spatial_size):
        k = size_key(spatial_size)
        if k not in self.levels:
            raise RuntimeError("Metadata has no active set at spatial size %s" % (k,))
        return self.levels[k]

    def rule_csr(self):
        """(row_ptr, row_pts) of the input rule; built lazily for mode 0 (identity)."""
        if self.row_ptr is None:
            dev = self.point_row.device
            self.row_ptr = torch.arange(self.n_points + 1, dtype=torch.int32, device=dev)
            self.row_pts = torch.arange(self.n_points, dtype=torch.int32, device=dev)
        return self.row_ptr, self.row_pts

    def prebuild(self, n_levels, filter_size=2, stride=2, subm_filter=3, book_channels=None):
        """Build `n_levels` strided levels below the input level (and the submanifold maps of every level) NOW.
        Each new level costs one host sync (its active-row count sizes the buffers).  Done lazily, those syncs land
        in the middle of the network and drain a full launch queue every time; done here, right after the input
        layer, they cost ~30 us each and everything after runs asynchronously.  Purely an ordering change: the
        same cached rulebooks are produced.  book_channels: layer width per level (level 0 first) -- the tile book of a level
        whose width the tile-local kernels implement is built here too (a prefetched geometry is then complete; otherwise
        the first layer that runs on the level builds it)."""
        size = self.input_size
        for _ in range(n_levels + 1):
            lvl = self.levels[size]

            def own_maps(lvl=lvl, i=_):      # this level's neighbour map (+ tile book): independent of the next level's count
                if subm_filter and lvl.n:
                    lvl.subm_map(subm_filter, tile_book=False)
                    if book_channels is not None and i < len(book_channels) and _triple(subm_filter) == (3, 3, 3):
                        lvl.ensure_tile_book(book_channels[i])
            if _ == n_levels or any(s % 2 for s in size) or min(size) < 2:
                own_maps()
                break
            done = []
            try:
                r = self.strided_rules(size, filter_size, stride, between=lambda: (own_maps(), done.append(1)))
            except RuntimeError:
                if not done:
                    own_maps()
                break
            size = r.out_key
        return self

    # ---------------------------------------------------------------- strided rulebooks
    def strided_rules(self, in_size, filter_size, stride, between=None):
        """`between`: optional callable that enqueues GPU work independent of the new level (run while the host waits
        for the level's row count; always called exactly once if given)."""
        f, st = _triple(filter_size), _triple(stride)
        ik = size_key(in_size)
        key = (ik, f, st)
        if key in self.strided:
            if between is not None:
                between()
            return self.strided[key]
        if f != st:
            raise RuntimeError("only filter_size == filter_stride is implemented (the form ndsis uses)")
        out_size = tuple((i - a) // b + 1 for i, a, b in zip(ik, f, st))
        if any((o - 1) * b + a != i for o, a, b, i in zip(out_size, f, st, ik)):
            raise RuntimeError("spatial size %s incompatible with filter %s / stride %s" % (ik, f, st))
        lin = self.level(ik)
        dev = lin.keys.device
        s = _stream()
        K = f[0] * f[1] * f[2]
        pkeys = _new(lin.n, torch.int64, dev)
        offs = _new(lin.n, torch.int32, dev)
        if out_size in self.levels:
            if between is not None:
                between()
            _lib.call("scn_stride_keys", _ptr(lin.keys), lin.n, st[0], st[1], st[2], _ptr(pkeys), _ptr(offs), s)
            lout = self.levels[out_size]
            parent_row = _new(lin.n, torch.int32, dev)
            _lib.call("scn_hash_lookup", _ptr(pkeys), lin.n, _ptr(lout.tab_keys), _ptr(lout.tab_vals), lout.cap,
                      _ptr(parent_row), s)
            cmap = _new((K, lout.n), torch.int32, dev).fill_(-1)
            dmap = _new((K, lin.n), torch.int32, dev)
            _lib.call("scn_strided_maps", _ptr(parent_row), _ptr(offs), lin.n, lout.n, K, _ptr(cmap), _ptr(dmap), s)
        else:
            # new coarse level: two C calls around the one host round trip (its active-row count)
            P = lin.n
            cap = 64
            while cap < 2 * P:
                cap <<= 1
            tab_keys = _new(cap, torch.int64, dev)
            tab_vals = _new(cap, torch.int32, dev)
            first = _new(max(P, 1), torch.int32, dev)
            rank = _new(P + 1, torch.int32, dev)
            tmp = _new(int(_lib.raw("scn_scan_tmp_elems")(P)), torch.int32, dev)
            _lib.call("scn_strided_level_count", _ptr(lin.keys), P, st[0], st[1], st[2], _ptr(pkeys), _ptr(offs),
                      _ptr(tab_keys), _ptr(tab_vals), cap, _ptr(first), _ptr(rank), _ptr(tmp), s)
            # The coarse level's row count sizes its buffers: one host round trip.  The count is copied to pinned memory
            # and an EVENT is recorded right behind the copy; `between` (prebuild: this level's neighbour map and tile
            # book, which do not depend on the count) is enqueued before the host waits, so the host waits for the count
            # kernels only and issues the finish kernels while that independent work still runs -- the round trip no
            # longer idles the GPU (kineto: ~30 us of idle per level before, profiles/r2_e_executor.md)
            slot = _count_slot(dev)
            slot[0].copy_(rank[P:P + 1], non_blocking=True)
            slot[1].record(torch.cuda.current_stream(dev))
            if between is not None:
                between()
            slot[1].synchronize()
            n = int(slot[0][0])
            parent_row = _new(P, torch.int32, dev)
            row_keys = _new(n, torch.int64, dev)
            cmap = _new((K, n), torch.int32, dev)
            dmap = _new((K, P), torch.int32, dev)
            _lib.call("scn_strided_level_finish", _ptr(pkeys), _ptr(offs), P, _ptr(tab_keys), _ptr(tab_vals), cap,
                      _ptr(rank), _ptr(parent_row), _ptr(row_keys), n, K, _ptr(cmap), _ptr(dmap), s)
            lout = Level(row_keys, tab_keys, tab_vals, cap, n)
            lout.coherent = lin.coherent      # coarse rows are numbered by their first fine row: a Morton code's parent is its prefix
            self.levels[out_size] = lout
        r = Strided(out_size, cmap, dmap, parent_row, K)
        self.strided[key] = r
        return r


# ---------------------------------------------------------------- pinned slot + event for row-count round trips
_count_slots = threading.local()      # per host thread (inference runs scenes from several threads / streams)


def _count_slot(device):
    slots = getattr(_count_slots, "by_device", None)
    if slots is None:
        slots = _count_slots.by_device = {}
    key = str(device)
    if key not in slots:
        slots[key] = (torch.zeros(1, dtype=torch.int32).pin_memory(), torch.cuda.Event())
    return slots[key]


# ---------------------------------------------------------------- host -> device staging one step ahead
_copy_streams = {}
_staged = {}


def stage_to_device(tensors, device):
    """Start the host->device copies of an UPCOMING batch's tensors (pinned host memory) on a side stream, so that they
    overlap the current step instead of sitting at the head of the next one.  Plain asynchronous copies issued by the
    calling thread: no worker thread, no host synchronisation.  `take_staged(t)` later returns the device copy.
    The copies land in one of three recycled device buffers (grow-only `Arena`s, round robin), not in fresh allocations:
    blocks allocated under the copy stream and used under the training stream would need `record_stream`, and with batches
    of a different size every step that kept the caching allocator calling cudaMalloc inside steps (10-30 ms each).  A
    buffer is rewritten three calls later; the copy stream first waits for what the calling stream had enqueued at the
    PREVIOUS call -- by then every kernel that read the buffer's previous contents (the step after the one that staged
    them) was enqueued -- and not for the step the caller is in the middle of (waiting for "now" delayed the copy, and the
    geometry built from it, by a step: e2e 5.2 -> 5.45 ms)."""
    device = torch.device(device)
    cs = _copy_streams.get(device)
    if cs is None:
        cs = _copy_streams[device] = torch.cuda.Stream(device)
    todo = [t for t in tensors if isinstance(t, torch.Tensor) and not t.is_cuda and id(t) not in _staged]
    if not todo:
        return
    ring = _stage_rings.setdefault(device, [[None] * 3, 0, [[], [], []], None])
    need = sum((t.numel() * t.element_size() + 255) // 256 * 256 for t in todo) + 256
    slot = ring[1] % 3
    ring[1] += 1
    arena = ring[0][slot]
    for k in ring[2][slot]:      # copies staged into THIS buffer and never taken die with its contents
        hit = _staged.get(k)
        if hit is not None and arena is not None and hit[1].untyped_storage().data_ptr() == arena.buf.untyped_storage().data_ptr():
            del _staged[k]
    ring[2][slot] = [id(t) for t in todo]
    if arena is None or arena.buf.numel() < need:
        arena = ring[0][slot] = Arena((need + (1 << 24) - 1) >> 24 << 24, device)      # 16 MB granules, grow only
    arena.off = 0
    ready, ring[3] = ring[3], torch.cuda.Event()
    ring[3].record(torch.cuda.current_stream(device))      # waited for by the NEXT call
    with torch.cuda.stream(cs):
        if ready is not None:
            cs.wait_event(ready)
        for t in todo:
            d = arena.take(tuple(t.shape), t.dtype)
            d.copy_(t, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
            _staged[id(t)] = (t, d, ev)


_stage_rings = {}


def take_staged(t):
    """Device copy of host tensor `t` if it was staged (ordered after the copy on the current stream), else None."""
    item = _staged.pop(id(t), None) if isinstance(t, torch.Tensor) else None
    if item is None or item[0] is not t:
        return None
    _, d, ev = item
    cur = torch.cuda.current_stream(d.device)
    cur.wait_event(ev)      # d is a slice of a recycled staging buffer: nothing for the allocator to track
    return d


def _walk_tensors(obj, seen):
    """Every torch tensor reachable from a Metadata (levels, maps, rules, caches)."""
    if id(obj) in seen:
        return
    seen.add(id(obj))
    if isinstance(obj, torch.Tensor):
        yield obj
    elif isinstance(obj, dict):
        for v in obj.values():
            yield from _walk_tensors(v, seen)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            yield from _walk_tensors(v, seen)
    elif isinstance(obj, (Level, Strided, Metadata)):
        yield from _walk_tensors(vars(obj), seen)


class GeometryPrefetcher:
    """Builds the geometry of the NEXT batch (voxel hash, every level of the pyramid, neighbour maps) while the current
    step runs: a worker thread issues the rulebook kernels on a high-priority side stream and takes their host round
    trips (one active-row count per level), so the training thread never waits for the GPU.  Same kernels, same results;
    only WHEN they run changes -- the role a DataLoader worker plays for the reference's CPU-side preprocessing
    (ndsis/data/data.py:88-115).  Without it the step has a bubble at its start: the host waits for the previous step
    to drain before it can read the first count, then the GPU waits for the host (profiles/r1_i_host_bound.md)."""

    def __init__(self, device, n_levels, mode=4, dimension=3, threaded=True, book_channels=None):
        """threaded=False: `submit` builds inline on the side stream from the CALLING thread (no worker, no GIL hand-offs).
        Called at the end of a training step -- after the step's forward, backward and optimizer kernels are enqueued --
        the builder's host round trips wait only for the side stream's own kernels while the GPU works through the queued
        step, and the next step starts without the host <-> GPU ping-pong of six row-count reads at its head."""
        self.device = torch.device(device)
        self.n_levels, self.mode, self.dimension = n_levels, mode, dimension
        self.book_channels = book_channels
        self.use_arena = os.environ.get("SCN_GEOMETRY_ARENA", "1") != "0"
        self._free, self._lock, self.last_arena_bytes = [], threading.Lock(), 0
        self.stream = torch.cuda.Stream(self.device, priority=-1)
        self.pool = None
        if threaded:
            import concurrent.futures
            self.pool = concurrent.futures.ThreadPoolExecutor(max_workers=1, thread_name_prefix="scn-geometry")
        self.pending = {}

    ARENA_BYTES_PER_POINT = 1024      # measured ~420 bytes per point on scan-like scenes (all levels, maps, tile books)

    def _acquire(self, n_points):
        need = max(int(n_points), 1) * self.ARENA_BYTES_PER_POINT
        need = (need + (1 << 25) - 1) >> 25 << 25      # 32 MB granules: scenes of similar size share arenas
        with self._lock:
            for i, a in enumerate(self._free):
                if a.buf.numel() >= need:
                    return self._free.pop(i)
        return Arena(need, self.device)

    def _release(self, arena, event):
        arena.event, arena.off = event, 0
        with self._lock:
            self._free.append(arena)

    def _build(self, coords, spatial_size, batch_size):
        torch.cuda.set_device(self.device)
        with torch.cuda.stream(self.stream):
            arena = self._acquire(len(coords)) if self.use_arena else None
            if arena is not None and arena.event is not None:
                self.stream.wait_event(arena.event)      # its previous geometry's last consumer kernels (device-side wait)
            _arena_tls.arena = arena
            try:
                md = Metadata(self.dimension)
                md.set_input(spatial_size, coords, batch_size, self.mode, self.device)
                md.prebuild(self.n_levels, book_channels=self.book_channels)
            except BaseException:
                if arena is not None:      # a build that fails (bad coordinates) gives its arena back
                    back = torch.cuda.Event()
                    back.record(self.stream)
                    self._release(arena, back)
                raise
            finally:
                _arena_tls.arena = None
            ev = torch.cuda.Event()
            ev.record(self.stream)
        md._prebuilt_for = coords
        md._ready = ev
        if arena is not None:
            md._lease = _Lease(self, arena)
            self.last_arena_bytes = arena.off
        return md

    def submit(self, coords, spatial_size, batch_size, again=False):
        """Start building the geometry of `coords` (a later InputLayer call with this very tensor picks it up).  A tensor
        that is already pending is not built twice unless `again` (inference over a scene list that repeats tensors: one
        geometry per occurrence, handed out in submission order)."""
        if not len(coords) or (id(coords) in self.pending and not again):
            return
        if self.pool is None:
            item = (_Done(self._build(coords, spatial_size, batch_size)), coords)
        else:
            item = (self.pool.submit(self._build, coords, spatial_size, batch_size), coords)
        with self._lock:
            self.pending.setdefault(id(coords), []).append(item)

    def take(self, coords):
        """The prefetched Metadata for `coords`, ordered after its build on the current stream; None if not submitted."""
        with self._lock:
            items = self.pending.get(id(coords))
            item = items.pop(0) if items else None
            if items is not None and not items:
                del self.pending[id(coords)]
        if item is None:
            return None
        md = item[0].result()
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(md._ready)
        lease = getattr(md, "_lease", None)
        base = lease.arena.buf.untyped_storage().data_ptr() if lease is not None else None
        for t in _walk_tensors(md, set()):
            if t.is_cuda and t.untyped_storage().data_ptr() != base:      # slices of the arena are not the allocator's business
                t.record_stream(cur)      # allocated on the side stream, used (and later freed) under the training stream
        return md

    def drain(self):
        """Forget geometries that were submitted and never taken (after an aborted run); waits for builds in flight."""
        with self._lock:
            items = [it for lst in self.pending.values() for it in lst]
            self.pending.clear()
        for fut, _ in items:
            try:
                fut.result()
            except Exception:
                pass

    def shutdown(self):
        if self.pool is not None:
            self.pool.shutdown(wait=True)
        self.pending.clear()


class _Lease:
    """Returns a geometry's arena to its prefetcher when the Metadata dies.  By then every kernel that reads the geometry has
    been enqueued (the autograd graph that held it is gone): an event on the releasing thread's current stream orders the
    arena's next build behind them."""

    def __init__(self, owner, arena):
        self.owner, self.arena = owner, arena

    def __del__(self):
        try:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.owner.device))
            self.owner._release(self.arena, ev)
        except Exception:      # interpreter shutdown
            pass


class _Done:
    """A finished result with the Future interface `take` uses."""

    def __init__(self, value):
        self.value = value

    def result(self):
        return self.value
