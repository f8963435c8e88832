"""ORACLE (test infrastructure, never shipped / never on the product path).

CPU restatement of SparseConvNet's *rulebook* semantics in numpy.

SparseConvNet (facebookresearch/SparseConvNet, unpinned master, installed by
`git clone` + `develop.sh` per /root/reference README.md:23-32) is NOT vendored
under /root/reference and is not installable here, so every function below
restates the upstream algorithm from its published behaviour (SURVEY.md
Appendix A1-A4) and is anchored on the reference's own call sites:

  * input layer ........ ndsis/modules/custom_operations.py:67-86 (mode 4),
                         ndsis/modules/roi_select_sparse.py:75-84,113-122 (mode 4 / 0)
  * locations .......... ndsis/modules/custom_operations.py:26-32,
                         ndsis/modules/roi_select_sparse.py:103-109 (batch-sorted assert)
  * submanifold rules .. ndsis/modules/module_factory.py:377-414
  * strided rules ...... ndsis/modules/module_factory.py:221-271,315-354

PARITY UNPINNED against SparseConvNet itself (no golden vectors exist in the
reference); the restatement is pinned instead against
independent dense `torch.nn.functional.conv3d` equivalences
(tests/test_oracle_dense_equivalence.py) and the reference's importable
`roi_cut` (tests/golden/).

Row-order contract: level-0 rows are numbered in order of first
appearance of the voxel in the input point list (upstream InputLayer.h).  At
strided levels upstream numbers rows in hash-iteration order, which is
implementation-defined; this oracle (and the CUDA path) number a coarse site
by the first fine *row* that maps onto it, which keeps every level grouped by
ascending batch index as ndsis/modules/roi_select_sparse.py:106-107 requires.
"""
import numpy as np

COORD_BITS = 16
COORD_MAX = (1 << COORD_BITS) - 1


def pack_keys(coords):
    """(x,y,z,b) int64 [N,4] -> uint64 key b<<48 | x<<32 | y<<16 | z.

    Sorting by key sorts by (b, x, y, z), the canonical order of SURVEY.md 8c.
    """
    c = np.asarray(coords, dtype=np.int64)
    if c.size and (c.min() < 0 or c.max() > COORD_MAX):
        raise RuntimeError("coordinate out of range [0, 65535]")
    c = c.astype(np.uint64)
    return (c[:, 3] << np.uint64(48)) | (c[:, 0] << np.uint64(32)) | \
           (c[:, 1] << np.uint64(16)) | c[:, 2]


def unpack_keys(keys):
    k = np.asarray(keys, dtype=np.uint64)
    m = np.uint64(COORD_MAX)
    out = np.empty((len(k), 4), dtype=np.int64)
    out[:, 0] = (k >> np.uint64(32)) & m
    out[:, 1] = (k >> np.uint64(16)) & m
    out[:, 2] = k & m
    out[:, 3] = (k >> np.uint64(48)) & m
    return out


class Grid:
    """One spatial scale of a Metadata: coordinate -> row lookup (upstream
    `SparseGrids`: per-sample hash map Point->row with global row ids)."""

    def __init__(self, coords):
        self.coords = np.ascontiguousarray(coords, dtype=np.int64)   # [N,4] x,y,z,b  (row order)
        self.keys = pack_keys(self.coords)
        self.order = np.argsort(self.keys, kind="stable")
        self.sorted_keys = self.keys[self.order]

    @property
    def n(self):
        return len(self.coords)

    def lookup(self, coords):
        """rows for (x,y,z,b) queries, -1 where inactive / out of range."""
        c = np.asarray(coords, dtype=np.int64)
        ok = ((c >= 0) & (c <= COORD_MAX)).all(1)
        q = pack_keys(np.where(ok[:, None], c, 0))
        if self.n == 0:
            return np.full(len(c), -1, dtype=np.int64)
        pos = np.searchsorted(self.sorted_keys, q)
        pos = np.minimum(pos, self.n - 1)
        hit = ok & (self.sorted_keys[pos] == q)
        return np.where(hit, self.order[pos], -1).astype(np.int64)


def input_layer_rules(coords, batch_size, mode):
    """Upstream Metadata/InputLayer.h `inputLayerRules`.

    coords: int64 [P, 4] (x,y,z,b)  (a [P,3] array means a single sample).
    Returns (grid, point_row[P], n_samples).
      mode 0: every point is its own row (coords guaranteed unique), rows = input order.
      mode>0: duplicates share a row; rows numbered by first appearance.
    """
    c = np.asarray(coords, dtype=np.int64)
    if c.ndim != 2:
        raise RuntimeError("coords must be [P, dim(+1)]")
    if c.shape[1] == 3:
        c = np.concatenate([c, np.zeros((len(c), 1), np.int64)], 1)
    n_samples = max(int(batch_size), int(c[:, 3].max()) + 1 if len(c) else 0, 1)
    if mode == 0:
        return Grid(c), np.arange(len(c), dtype=np.int64), n_samples
    keys = pack_keys(c)
    _, first, inverse = np.unique(keys, return_index=True, return_inverse=True)
    rank = np.empty(len(first), dtype=np.int64)
    rank[np.argsort(first, kind="stable")] = np.arange(len(first))
    point_row = rank[inverse.reshape(-1)]
    first_sorted = np.sort(first)
    return Grid(c[first_sorted]), point_row, n_samples


def filter_offsets(filter_size):
    """All offsets of a K^3 box, last dimension fastest (upstream
    SubmanifoldConvolutionRules.h / RectangularRegions.h iteration order)."""
    fs = np.broadcast_to(np.asarray(filter_size, dtype=np.int64), (3,))
    g = np.stack(np.meshgrid(*[np.arange(f) for f in fs], indexing="ij"), -1)
    return g.reshape(-1, 3), fs


def submanifold_rules(grid, filter_size):
    """Upstream SubmanifoldConvolutionRules.h: for every active
    output site and every offset o (last dim fastest) emit (in_row, out_row)
    when the neighbour is active.  Returned as a list over offsets of
    (in_rows, out_rows) int64 arrays, ordered by ascending out_row."""
    offs, fs = filter_offsets(filter_size)
    rules = []
    rows = np.arange(grid.n, dtype=np.int64)
    for o in offs:
        q = grid.coords.copy()
        q[:, :3] += o - fs // 2
        nb = grid.lookup(q)
        m = nb >= 0
        rules.append((nb[m], rows[m]))
    return rules


def strided_rules(grid_in, in_size, filter_size, stride):
    """Upstream ConvolutionRules.h for filter == stride (the
    only strided form ndsis uses: module_factory.py:221-271,315-354).

    out_site = p // stride, offset index = row-major index of (p - out*stride)
    in the filter box (last dim fastest).  Coarse rows are numbered by first
    appearance while scanning the fine rows in order.
    Returns (grid_out, out_size, rules[offset] = (in_rows, out_rows), parent[N_in], off[N_in])."""
    fs = np.broadcast_to(np.asarray(filter_size, dtype=np.int64), (3,))
    st = np.broadcast_to(np.asarray(stride, dtype=np.int64), (3,))
    if not (fs == st).all():
        raise RuntimeError("oracle restates filter_size == filter_stride only")
    in_size = np.asarray(in_size, dtype=np.int64)
    out_size = (in_size - fs) // st + 1
    if ((out_size - 1) * st + fs != in_size).any():
        # upstream scn.Convolution asserts this per dimension
        raise RuntimeError("input size %s not compatible with filter/stride" % (in_size,))
    pc = grid_in.coords.copy()
    pc[:, :3] //= st
    rem = grid_in.coords[:, :3] - pc[:, :3] * st
    off = (rem[:, 0] * fs[1] + rem[:, 1]) * fs[2] + rem[:, 2]
    grid_out, parent, _ = input_layer_rules(pc, 0, 4)
    nk = int(fs.prod())
    rows = np.arange(grid_in.n, dtype=np.int64)
    rules = []
    for o in range(nk):
        m = off == o
        i, p = rows[m], parent[m]
        s = np.argsort(p, kind="stable")
        rules.append((i[s], p[s]))
    return grid_out, out_size, rules, parent, off


def rules_to_map(rules, n_out):
    """Rule lists -> output-stationary neighbour map [K, n_out] int32 (-1 =
    inactive): the canonical, order-independent form compared bit-exactly with
    the CUDA rulebook builder."""
    m = np.full((len(rules), n_out), -1, dtype=np.int32)
    for o, (i, p) in enumerate(rules):
        m[o, p] = i
    return m
