"""CPU checks of the bounding-box tree behind CVTArchive's search for measure_dim > 3 (csrc/cvt_tree.cu).

The tree is built on the host (qd_cvt_tree_build_host, the analogue of the KD-tree the reference builds in
CVTArchive.__init__, ribs/archives/_cvt_archive.py:410-429), so its structure can be checked without a GPU:
every centroid sits in exactly one leaf, every box encloses what is below it, and a NumPy walk of the structure
with the kernel's visiting rule returns np.argmin of the exact distances (_cvt_archive.py:626-632).
"""
import ctypes as C

import numpy as np
import pytest

LEAF = FAN = 32
MAXLEV = 7


def build(cen):
    from pyribs_b200 import _lib

    L = _lib.lib()
    cen = np.ascontiguousarray(cen)
    n, d = cen.shape
    code = _lib.QD_F64 if cen.dtype == np.float64 else _lib.QD_F32
    nbytes = int(L.qd_cvt_tree_bytes(n, d))
    assert nbytes > 0
    blob = np.zeros(nbytes, dtype=np.uint8)
    _lib.check(L.qd_cvt_tree_build_host(cen.ctypes.data_as(C.c_void_p), code, n, d, blob.ctypes.data_as(C.c_void_p)))
    return parse(blob, n, d)


def parse(blob, n, d):
    hdr = blob[:256]
    Cn = int(hdr[0:8].view(np.int64)[0])
    dd, nleaves, nlev, _ = (int(v) for v in hdr[8:24].view(np.int32))
    nl = hdr[24:24 + 4 * (MAXLEV + 1)].view(np.int32)
    off_box = hdr[56:56 + 8 * (MAXLEV + 1)].view(np.uint64)
    off_pts, off_ids, total = (int(v) for v in hdr[120:144].view(np.uint64))
    assert (Cn, dd) == (n, d) and total == blob.size and nleaves == -(-n // LEAF)
    pts = blob[off_pts:off_pts + nleaves * d * LEAF * 8].view(np.float64).reshape(nleaves, d, LEAF)
    ids = blob[off_ids:off_ids + nleaves * LEAF * 4].view(np.int32).reshape(nleaves, LEAF)
    boxes = []
    dp = (d + 1) // 2  # fp32 data is stored as pairs of coordinates: [group][lo | hi][pair][child][2]
    for l in range(nlev - 1):
        g = int(nl[l + 1])
        o = int(off_box[l])
        b = blob[o:o + g * 2 * dp * FAN * 8].view(np.float32).reshape(g, 2, dp, FAN, 2)
        b = np.ascontiguousarray(b.transpose(0, 1, 2, 4, 3)).reshape(g, 2, 2 * dp, FAN)
        if d & 1:  # the partner of an odd last coordinate: [0, 0] for real children, empty like the rest otherwise
            pad = b[:, :, d, :]
            real = np.isfinite(b[:, 0, 0, :])
            assert np.all(pad[:, 0][real] == 0) and np.all(pad[:, 1][real] == 0)
            assert np.all(np.isposinf(pad[:, 0][~real])) and np.all(np.isneginf(pad[:, 1][~real]))
        boxes.append(np.ascontiguousarray(b[:, :, :d, :]))
    return dict(pts=pts, ids=ids, boxes=boxes, nlev=nlev, n=[int(v) for v in nl[:nlev]])


def walk(tree, x):
    """The kernel's traversal with exact fp64 bounds: best-first over a node's children, stop at the first
    child whose bound exceeds the best squared distance so far; (distance, id) lexicographic minimum."""
    best, bi, visited = np.inf, np.iinfo(np.int32).max, 0

    def leaf(j):
        nonlocal best, bi, visited
        visited += 1
        diff = x[:, None] - tree["pts"][j]
        with np.errstate(invalid="ignore", over="ignore"):
            dd = np.sum(np.square(np.ascontiguousarray(diff.T)), axis=1)
        for t in np.argsort(dd, kind="stable"):
            if dd[t] < best or (dd[t] == best and tree["ids"][j, t] < bi):
                best, bi = dd[t], int(tree["ids"][j, t])

    def node(level, j):
        b = tree["boxes"][level - 1][j].astype(np.float64)
        gap = np.maximum(np.maximum(b[0] - x[:, None], x[:, None] - b[1]), 0.0)
        lb = np.sum(gap * gap, axis=0)
        for c in np.argsort(lb, kind="stable"):
            if not np.isfinite(lb[c]) or lb[c] > best:
                break
            (leaf if level == 1 else lambda k: node(level - 1, k))(j * FAN + c)

    if tree["nlev"] == 1:
        leaf(0)
    else:
        node(tree["nlev"] - 1, 0)
    return bi, best, visited


@pytest.mark.parametrize("n,d,dtype", [(1, 4, np.float64), (31, 5, np.float64), (32, 8, np.float64),
                                        (33, 8, np.float32), (1024, 8, np.float64), (1025, 6, np.float64),
                                        (40_000, 8, np.float64), (5000, 16, np.float32)])
def test_tree_structure(n, d, dtype):
    rng = np.random.default_rng(n + d)
    cen = rng.uniform(-1, 1, (n, d)).astype(dtype)
    t = build(cen)
    ids = t["ids"].reshape(-1)
    real = ids[ids != np.iinfo(np.int32).max]
    assert np.array_equal(np.sort(real), np.arange(n)), "every centroid in exactly one leaf slot"
    assert np.all(ids[:n] != np.iinfo(np.int32).max) and np.all(ids[n:] == np.iinfo(np.int32).max)
    flat = t["pts"].transpose(0, 2, 1).reshape(-1, d)
    assert np.array_equal(flat[:n], cen[real].astype(np.float64))
    assert np.all(np.isposinf(flat[n:]))
    # level-0 boxes enclose the leaf's points; higher boxes enclose their children
    if t["nlev"] > 1:
        b0 = t["boxes"][0].astype(np.float64)
        for j in range(t["n"][0]):
            m = t["ids"][j] != np.iinfo(np.int32).max
            p = t["pts"][j][:, m]
            lo, hi = b0[j // FAN, 0, :, j % FAN], b0[j // FAN, 1, :, j % FAN]
            assert np.all(lo <= p.min(1)) and np.all(hi >= p.max(1))
            assert np.all(p.min(1) - lo <= np.abs(lo) * 2e-7 + 1e-37), "boxes are tight to one fp32 ulp"
        for l in range(1, t["nlev"] - 1):
            child, par = t["boxes"][l - 1], t["boxes"][l]
            for j in range(t["n"][l]):
                lo, hi = par[j // FAN, 0, :, j % FAN], par[j // FAN, 1, :, j % FAN]
                assert np.array_equal(lo, child[j, 0].min(axis=1)) and np.array_equal(hi, child[j, 1].max(axis=1))
        # slots past the end of a level are empty boxes (never visited)
        for l in range(t["nlev"] - 1):
            b = t["boxes"][l].transpose(0, 3, 1, 2).reshape(-1, 2, d)[t["n"][l]:]
            assert np.all(np.isposinf(b[:, 0])) and np.all(np.isneginf(b[:, 1]))


def test_tree_walk_equals_argmin():
    rng = np.random.default_rng(3)
    n, d = 20_000, 8
    cen = rng.uniform(-1, 1, (n, d))
    cen[500:520] = cen[100:120]          # duplicated centroids: exact ties, the lower index must win
    t = build(cen)
    q = rng.uniform(-1.2, 1.2, (200, d))
    q[:20] = cen[100:120]                # distance 0 to two centroids each
    q[20:30] = 0.5 * (cen[7] + cen[8])   # (near-)ties between two centroids
    visited = []
    for x in q:
        bi, best, v = walk(t, x)
        dd = np.sum(np.square(x[None] - cen), axis=1)
        assert bi == int(np.argmin(dd)) and best == dd.min()
        visited.append(v)
    assert np.mean(visited) < 0.15 * (n / LEAF), "the boxes prune: a query visits a small part of the leaves"
