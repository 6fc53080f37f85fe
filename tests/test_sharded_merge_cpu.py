"""World-size-2 gloo test (CPU) of the host-side logic of the centroid-sharded CVT search:
shard bounds + the two-step all-reduce merge reproduce np.argmin over the full centroid set,
including exact ties across shards."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pyribs_b200.archives._cvt_archive import merge_sharded_argmin, shard_bounds

    rng = np.random.default_rng(0)
    cen = rng.integers(-4, 5, (301, 3)).astype(np.float64) / 4.0  # lattice => many exact ties across shards
    meas = rng.integers(-8, 9, (500, 3)).astype(np.float64) / 8.0
    lo, hi = shard_bounds(len(cen), world, rank)
    dd = ((meas[:, None, :] - cen[None, lo:hi, :]) ** 2).sum(-1)
    idx = torch.from_numpy((dd.argmin(1) + lo).astype(np.int32))
    d_local = torch.from_numpy(dd.min(1))

    def mask(d_loc, d_min, ix):
        ix[d_loc != d_min] = 2**31 - 1

    idx, d_min = merge_sharded_argmin(idx, d_local, None, mask)
    full = ((meas[:, None, :] - cen[None]) ** 2).sum(-1)
    ok = bool(np.array_equal(idx.numpy(), full.argmin(1).astype(np.int32))) and bool(
        np.array_equal(d_min.numpy(), full.min(1)))
    q.put((rank, ok, (lo, hi)))
    dist.destroy_process_group()


def test_sharded_merge_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    bounds = sorted(b for _, _, b in res)
    assert bounds == [(0, 151), (151, 301)]
