"""CPU: the rulebook builder's arena (scn/metadata.py: Arena, _new) -- the bump allocator behind geometry built one step
ahead (no caching-allocator traffic on the side stream)."""
import threading

import torch

from sparse_rcnn_b200.scn import metadata as M


def test_slices_are_aligned_disjoint_and_typed():
    a = M.Arena(1 << 16, "cpu")
    t1 = a.take((3, 5), torch.int32)
    t2 = a.take(7, torch.int64)
    t3 = a.take((0, 4), torch.float32)
    t4 = a.take(100, torch.uint8)
    assert t1.shape == (3, 5) and t1.dtype == torch.int32 and t2.shape == (7,) and t3.numel() == 0
    base = a.buf.data_ptr()
    for t in (t1, t2, t4):
        assert (t.data_ptr() - base) % 256 == 0
    t1.fill_(-1), t2.fill_(7), t4.fill_(9)
    assert bool((t1 == -1).all()) and bool((t2 == 7).all()) and bool((t4 == 9).all())      # no overlap
    assert all(t.untyped_storage().data_ptr() == a.buf.untyped_storage().data_ptr() for t in (t1, t2, t4))
    assert a.take(1 << 16, torch.uint8) is None                                              # exhausted: caller falls back


def test_new_uses_the_thread_local_arena_only_where_installed():
    a = M.Arena(4096, "cpu")
    M._arena_tls.arena = a
    try:
        x = M._new(16, torch.int32, "cpu")
        big = M._new(1 << 20, torch.int32, "cpu")      # does not fit: a regular tensor
    finally:
        M._arena_tls.arena = None
    assert x.untyped_storage().data_ptr() == a.buf.untyped_storage().data_ptr()
    assert big.untyped_storage().data_ptr() != a.buf.untyped_storage().data_ptr() and big.numel() == 1 << 20
    y = M._new(16, torch.int32, "cpu")
    assert y.untyped_storage().data_ptr() != a.buf.untyped_storage().data_ptr()
    seen = []
    M._arena_tls.arena = a
    try:
        th = threading.Thread(target=lambda: seen.append(getattr(M._arena_tls, "arena", None)))
        th.start(), th.join()
    finally:
        M._arena_tls.arena = None
    assert seen == [None]                                   # another thread never sees this thread's arena
