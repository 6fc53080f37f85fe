"""World-size-2 gloo test (CPU) of the host-side logic of the owner-computes (cell-sharded) store
(pyribs_b200/archives/_shard_util.py): cell ranges, the collective that gathers the shards'
running statistics, and the stitching of `data()` / `retrieve()` answers reproduce what one process holding the whole
archive reports. The device side of the same path is tests/test_gpu_multi.py::test_two_gpu_owner_computes_store."""
import os
import socket
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _whole_archive(rng, cells):
    """A fake archive state: which cells are occupied, in which order they were inserted, and their rows."""
    occ = rng.permutation(cells)[: cells // 3]  # insertion order
    obj = rng.standard_normal(len(occ))
    obj[7] = obj[19] = obj.max() + 1.0  # the best objective sits in two cells (in different shards for world = 2)
    sol = rng.standard_normal((len(occ), 3))
    return occ.astype(np.int32), obj, sol


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pyribs_b200.archives import _shard_util as U

    cells = 1001
    occ, obj, sol = _whole_archive(np.random.default_rng(5), cells)
    lo, hi = U.shard_cell_range(cells, world, rank)
    mine = (occ >= lo) & (occ < hi)
    ok, why = True, []

    def check(c, msg):
        nonlocal ok
        if not c:
            ok = False
            why.append(msg)

    # running statistics: four doubles per rank, gathered, merged
    local_obj = obj[mine]
    rows = U.all_gather_rows((mine.sum(), local_obj.sum(), local_obj.max() if mine.any() else -np.inf,
                              1.0 if mine.any() else 0.0), world)
    n, osum, omax, owner = U.merge_shard_stats(rows, np.float64)
    check(n == len(occ), "elite count")
    check(abs(osum - obj.sum()) <= 1e-12 * abs(obj).sum(), "objective sum")
    check(omax == obj.max(), "best objective")
    best_cells = occ[obj == obj.max()]
    def rank_of(c):
        return next(r for r in range(world) if U.shard_cell_range(cells, world, r)[0] <= c < U.shard_cell_range(cells, world, r)[1])

    check(owner == min(rank_of(c) for c in best_cells), "owner = lowest rank holding the best objective")
    # data(): the shards one after the other, archive-wide indices
    local = {"index": occ[mine], "objective": obj[mine], "solution": sol[mine]}
    parts = [None] * world
    dist.all_gather_object(parts, local)
    d = U.stitch_shard_data(parts)
    order_d, order_w = np.argsort(d["index"]), np.argsort(occ)
    check(np.array_equal(d["index"][order_d], occ[order_w]) and np.array_equal(d["objective"][order_d], obj[order_w])
          and np.array_equal(d["solution"][order_d], sol[order_w]), "data() as a set")
    ranks_of = np.array([rank_of(c) for c in d["index"]])
    check(np.all(np.diff(ranks_of) >= 0), "data() is rank-ordered")
    check(np.array_equal(d["index"][ranks_of == rank], occ[mine]), "insertion order kept within a shard")
    # retrieve(): every row from the rank that owns its cell
    probe = np.random.default_rng(9).integers(0, cells, 200)
    dense_obj = np.full(cells, np.nan)
    dense_obj[occ] = obj
    dense_occ = np.zeros(cells, dtype=bool)
    dense_occ[occ] = True
    own = (probe >= lo) & (probe < hi)
    local_rows = {"objective": np.where(own, dense_obj[probe], -123.0)}  # rows of foreign cells hold junk
    parts = [None] * world
    dist.all_gather_object(parts, (own, dense_occ[probe] & own, local_rows))
    occupied, out = U.stitch_shard_retrieve(probe, parts)
    check(np.array_equal(occupied, dense_occ[probe]), "retrieve occupied")
    check(np.array_equal(out["objective"], dense_obj[probe], equal_nan=True), "retrieve rows")
    check(np.array_equal(out["index"], probe.astype(np.int32)), "retrieve index")
    q.put((rank, ok, why, (lo, hi)))
    dist.destroy_process_group()


def test_shard_store_host_logic_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res
    assert sorted(r[3] for r in res) == [(0, 501), (501, 1001)]


def test_shard_cell_range_covers_every_cell_once():
    from pyribs_b200.archives._shard_util import merge_shard_stats, shard_cell_range

    for cells, world in ((1, 1), (10, 3), (1_000_000, 8), (7, 8), (250_000, 4)):
        edges = [shard_cell_range(cells, world, r) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == cells
        assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
    assert merge_shard_stats([[0, 0, -np.inf, 0], [0, 0, -np.inf, 0]], np.float64) == (0, 0.0, None, -1)
    assert merge_shard_stats([[2, 3.0, 2.5, 1], [1, 2.5, 2.5, 1]], np.float64) == (3, 5.5, 2.5, 0)
