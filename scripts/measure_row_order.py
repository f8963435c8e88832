"""Row-order statistics of the bench scene (CPU, numpy): for a 128-row output tile of a SubM 3^3 layer under a given
row order -- (a) fraction of (tile, offset) units with NO active pair (skippable), (b) distinct input rows a tile
touches over all 27 offsets (the tile's "halo set"), (c) fill density of non-empty units.  Orders: first appearance
(SparseConvNet / round 1), Morton (b, interleave(x,y,z)), and 4^3-block Morton.  Recorded in profiles/r2_a_row_order.md."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from sparse_rcnn_b200.synthetic import make_batch


def part1by2(v):
    v = v.astype(np.uint64) & 0x1fffff
    v = (v | (v << 32)) & 0x1f00000000ffff
    v = (v | (v << 16)) & 0x1f0000ff0000ff
    v = (v | (v << 8)) & 0x100f00f00f00f00f
    v = (v | (v << 4)) & 0x10c30c30c30c30c3
    v = (v | (v << 2)) & 0x1249249249249249
    return v


def morton(c):
    return (part1by2(c[:, 0]) << 2) | (part1by2(c[:, 1]) << 1) | part1by2(c[:, 2])


def stats(c, name, tile=128):
    n = len(c)
    key = (c[:, 0].astype(np.int64) << 40) | (c[:, 1].astype(np.int64) << 20) | c[:, 2].astype(np.int64)
    order = np.argsort(key)
    skey = key[order]
    maps = np.full((27, n), -1, np.int64)
    o = 0
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dz in (-1, 0, 1):
                q = c + np.array([dx, dy, dz])
                qk = (q[:, 0].astype(np.int64) << 40) | (q[:, 1].astype(np.int64) << 20) | q[:, 2].astype(np.int64)
                pos = np.searchsorted(skey, qk)
                pos[pos >= n] = n - 1
                hit = skey[pos] == qk
                maps[o, hit] = order[pos[hit]]
                o += 1
    nt = (n + tile - 1) // tile
    empty = 0
    halo = []
    dens = []
    cnt_hist = np.zeros(28, int)
    for t in range(nt):
        m = maps[:, t * tile:(t + 1) * tile]
        act = (m >= 0)
        per = act.sum(1)
        empty += (per == 0).sum()
        cnt_hist[(per > 0).sum()] += 1
        dens.append(per[per > 0].mean() / m.shape[1])
        halo.append(len(np.unique(m[m >= 0])))
    halo = np.array(halo)
    print("%-22s N=%d tiles=%d kbar=%.2f | empty units %.1f%% | fill of non-empty units %.1f%% | halo rows/tile: mean %.0f p50 %d p90 %d p99 %d max %d"
          % (name, n, nt, (maps >= 0).sum() / n, 100.0 * empty / (27 * nt), 100 * np.mean(dens), halo.mean(),
             np.percentile(halo, 50), np.percentile(halo, 90), np.percentile(halo, 99), halo.max()))
    return maps


def level_coords(c):
    u = np.unique(c // 2, axis=0)
    return u


if __name__ == "__main__":
    coords = make_batch(1, 0)[0].numpy()[:, :3]
    _, first = np.unique((coords[:, 0] << 40) | (coords[:, 1] << 20) | coords[:, 2], return_index=True)
    c0 = coords[np.sort(first)]          # first-appearance order
    for lvl in range(3):
        print("--- level %d" % lvl)
        stats(c0, "first appearance" if lvl == 0 else "first fine row")
        cm = c0[np.argsort(morton(c0), kind="stable")]
        stats(cm, "morton")
        blk = c0 // 4
        k2 = (morton(blk) << np.uint64(6)) | ((c0[:, 2] % 4).astype(np.uint64) << np.uint64(4)) | ((c0[:, 0] % 4).astype(np.uint64) << np.uint64(2)) | (c0[:, 1] % 4).astype(np.uint64)
        stats(c0[np.argsort(k2, kind="stable")], "4^3 block, z-major in")
        lex = c0[np.lexsort((c0[:, 2], c0[:, 1], c0[:, 0]))]
        stats(lex, "lexicographic x,y,z")
        # next level keeps the order of the first fine row that maps onto each coarse voxel
        p = c0 // 2
        pk = (p[:, 0] << 40) | (p[:, 1] << 20) | p[:, 2]
        _, f = np.unique(pk, return_index=True)
        c0 = p[np.sort(f)]
