"""Two-GPU tests (NCCL): data-parallel insertion and centroid-sharded CVT search must reproduce
the single-archive oracle on the global batch. Skipped on boxes with fewer than 2 GPUs."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import qd_oracle as O

    from pyribs_b200.archives import CVTArchive, GridArchive

    rng = np.random.default_rng(7)
    n, Bl = 6, 5000
    ok = True
    # (1) data-parallel GridArchive (CMA-MAE) vs oracle on the concatenated batch
    gpu = GridArchive(solution_dim=n, dims=(30, 30), ranges=[(-1, 1)] * 2, learning_rate=0.05, threshold_min=0.0,
                      data_parallel=True)
    ora = O.make_grid_oracle(solution_dim=n, dims=(30, 30), ranges=[(-1, 1)] * 2, learning_rate=0.05,
                             threshold_min=0.0)
    cen = rng.uniform(-1, 1, (4000, 3))
    cvt = CVTArchive(solution_dim=n, centroids=cen, ranges=[(-1, 1)] * 3, shard_centroids=True, search_method="tensor")
    cvt_dp = CVTArchive(solution_dim=n, centroids=cen, ranges=[(-1, 1)] * 3, data_parallel=True)
    cvt_both = CVTArchive(solution_dim=n, centroids=cen, ranges=[(-1, 1)] * 3, data_parallel=True, shard_centroids=True,
                          search_method="tensor")
    ora_cvt = O.make_cvt_oracle(solution_dim=n, centroids=cen)
    for step in range(3):
        g = dict(solution=rng.standard_normal((world * Bl, n)), objective=rng.uniform(0, 10, world * Bl) + step,
                 measures=rng.normal(0, 0.3, (world * Bl, 2)).clip(-1, 1))
        sl = slice(rank * Bl, (rank + 1) * Bl)
        out = gpu.add(g["solution"][sl], g["objective"][sl], g["measures"][sl])
        ref = ora.add(**g)
        ok &= bool(np.array_equal(out["status"], ref["status"][sl]))
        ok &= bool(np.allclose(out["value"], ref["value"][sl], rtol=1e-9, atol=0))
        ok &= len(gpu) == len(ora)
        ok &= bool(np.array_equal(gpu._store.occupied_list, ora.store.occupied_list[: len(ora)]))
        occ = ora.store.occupied
        ok &= bool(np.allclose(gpu._store._fields["threshold"].cpu().numpy()[occ], ora.store.fields["threshold"][occ],
                               rtol=1e-9))
        # (2) centroid-sharded CVT: every rank passes the FULL batch, each searches its shard
        m3 = rng.uniform(-1.1, 1.1, (world * Bl, 3))
        g3 = dict(solution=g["solution"], objective=g["objective"], measures=m3)
        out3 = cvt.add(**g3)
        ref3 = ora_cvt.add(**g3)
        ok &= bool(np.array_equal(out3["status"], ref3["status"]))
        ok &= bool(np.array_equal(cvt.index_of(m3), ora_cvt.index_fn(m3)))
        # (3) data-parallel CVT: each rank passes its slice
        out4 = cvt_dp.add(g3["solution"][sl], g3["objective"][sl], m3[sl])
        ok &= bool(np.array_equal(out4["status"], ref3["status"][sl]))
        ok &= len(cvt_dp) == len(ora_cvt)
        # (4) both: rows sharded over the ranks AND centroids sharded over the ranks
        out5 = cvt_both.add(g3["solution"][sl], g3["objective"][sl], m3[sl])
        ok &= bool(np.array_equal(out5["status"], ref3["status"][sl]))
        ok &= len(cvt_both) == len(ora_cvt)
        ok &= bool(np.array_equal(cvt_both._store.occupied_list, ora_cvt.store.occupied_list[: len(ora_cvt)]))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_two_gpu_insertion():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def _emitter_worker(rank, world, port, q):
    """Emitters sharded over the ranks: every rank owns E/world emitters and a replica of
    a data-parallel archive; `add` all-gathers the ranks' batches. With restart rule "basic" nothing depends
    on rank-local randomness, so the replicas must equal a single-process run of all emitters."""
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from pyribs_b200.archives import GridArchive
    from pyribs_b200.emitters import EvolutionStrategyEmitter
    from pyribs_b200.schedulers import Scheduler

    n, E, lam = 12, 6, 10
    mb = n / 2 * 5.12

    def sphere(sol):
        worst = (5.12 + 2.048) ** 2 * n
        obj = (worst - np.sum((sol - 2.048) ** 2, axis=1)) / worst * 100
        c = np.clip(sol, -5.12, 5.12)
        return obj, np.stack((c[:, : n // 2].sum(1), c[:, n // 2:].sum(1)), axis=1)

    def run(data_parallel, emitter_ids):
        a = GridArchive(solution_dim=n, dims=(30, 30), ranges=[(-mb, mb)] * 2, learning_rate=0.05, threshold_min=0.0,
                        data_parallel=data_parallel)
        seeds = np.random.SeedSequence(5).spawn(E)
        ems = [EvolutionStrategyEmitter(a, x0=np.zeros(n), sigma0=0.5, batch_size=lam, seed=seeds[i], ranker="imp",
                                        selection_rule="mu", restart_rule="basic") for i in emitter_ids]
        s = Scheduler(a, ems)
        for _ in range(25):
            s.tell(*sphere(s.ask()))
        return a

    per = E // world
    sharded = run(True, range(rank * per, (rank + 1) * per))
    whole = run(False, range(E))
    ds, dw = sharded.data(), whole.data()
    ok = len(sharded) == len(whole) and all(np.array_equal(ds[k], dw[k]) for k in ds)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_gpu_sharded_emitters():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_emitter_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def _shard_tie_worker(rank, world, port, q):
    """Centroid-sharded search at 8-D with exact ties ACROSS the shard boundary: duplicated centroids whose copies
    live on different ranks and lattice points equidistant from centroids of both shards. The two all-reduces
    (distance, then index among the ranks holding the minimum) must return the lowest GLOBAL index, like
    np.argmin over the unsharded array (_cvt_archive.py:626-632)."""
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from pyribs_b200.archives import CVTArchive

    rng = np.random.default_rng(11)
    C, B, d = 40_000, 30_000, 8
    cen = rng.uniform(-1, 1, (C, d))
    cen[C // 2: C // 2 + 60] = cen[:60]                       # copies on the other rank
    cen[C // 2 + 100: C // 2 + 164] = np.round(cen[100:164] * 4) / 4
    cen[200:264] = cen[C // 2 + 100: C // 2 + 164]            # lattice points present in both shards
    meas = rng.uniform(-1.1, 1.1, (B, d))
    meas[:60] = cen[:60]
    meas[60:124] = cen[200:264]
    meas[124:188] = cen[200:264] + 0.125                       # exactly representable offsets: ties in fp64
    ok = True
    want = None
    for method in ("auto", "tensor", "scan"):
        sh = CVTArchive(solution_dim=1, centroids=cen, ranges=[(-1, 1)] * d, shard_centroids=True, search_method=method)
        got = sh.index_of(meas)
        if want is None:
            one = CVTArchive(solution_dim=1, centroids=cen, ranges=[(-1, 1)] * d, search_method="scan")
            want = one.index_of(meas)
            sub = np.r_[0:188, rng.choice(B, 300, replace=False)]
            dd = ((meas[sub, None, :] - cen[None, :, :]) ** 2).sum(-1)
            ok &= bool(np.array_equal(want[sub], dd.argmin(1)))
            ok &= bool(np.array_equal(want[:60], np.arange(60)))
            ok &= bool(np.array_equal(want[60:124], np.arange(200, 264)))
        ok &= bool(np.array_equal(got, want))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_gpu_sharded_search_cross_shard_ties():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_tie_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def _owned_worker(rank, world, port, q):
    """Owner-computes store: cells sharded over the ranks by contiguous range, every rank
    commits only the rows of its own cells. Against a single-process archive fed the global batches: status / value
    of the rank's rows bit-equal, the same elites (as a set -- the sharded list is shard after shard), the same
    statistics and best elite, collective retrieve / sample_elites, and a batch with a NaN objective dropped by
    every rank."""
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from pyribs_b200.archives import CVTArchive, GridArchive

    ok, why = True, []

    def check(cond, msg):
        nonlocal ok
        if not cond:
            ok = False
            why.append(msg)

    rng = np.random.default_rng(11)
    cen = rng.uniform(-1, 1, (3001, 6))
    makers = {
        "grid_mae": lambda **kw: GridArchive(solution_dim=5, dims=(37, 41), ranges=[(-1, 1)] * 2, learning_rate=0.1,
                                             threshold_min=-1.0, extra_fields={"tag": ((), np.int32)}, seed=7, **kw),
        "cvt6": lambda **kw: CVTArchive(solution_dim=5, centroids=cen, ranges=[(-1, 1)] * 6, seed=7, **kw),
    }
    for name, make in makers.items():
        sharded, whole = make(data_parallel=True, shard_store=True), make()
        d = whole.measure_dim
        Bl = 4000
        for step in range(4):
            full = dict(solution=rng.standard_normal((world * Bl, 5)), objective=rng.uniform(0, 10, world * Bl) + step,
                        measures=rng.uniform(-1, 1, (world * Bl, d)))
            if step == 2:  # exact ties inside one cell, rows on different ranks: first in the GLOBAL batch wins
                full["measures"][Bl - 5:Bl + 5] = full["measures"][0]
                full["objective"][Bl - 5:Bl + 5] = 99.0
            if name == "grid_mae":
                full["tag"] = rng.integers(0, 1000, world * Bl).astype(np.int32)
            mine = {k: v[rank * Bl:(rank + 1) * Bl] for k, v in full.items()}
            if step % 2:  # CUDA tensors in (asynchronous) and NumPy in
                got = sharded.add(**{k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in mine.items()})
                got = {k: v.cpu().numpy() for k, v in got.items()}
            else:
                got = sharded.add(**mine)
            ref = whole.add(**full)
            check(np.array_equal(got["status"], ref["status"][rank * Bl:(rank + 1) * Bl]), f"{name} s{step} status")
            check(np.array_equal(got["value"], ref["value"][rank * Bl:(rank + 1) * Bl]), f"{name} s{step} value")
        check(len(sharded) == len(whole) and sharded.cells == whole.cells, f"{name} len / cells")
        ds, dw = sharded.data(), whole.data()
        o1, o2 = np.argsort(ds["index"]), np.argsort(dw["index"])
        for k in dw:
            check(np.array_equal(ds[k][o1], dw[k][o2]), f"{name} data[{k}]")
        lo, hi = sharded._cell_lo, sharded._cell_hi
        check(len(sharded._store) == int(np.count_nonzero((dw["index"] >= lo) & (dw["index"] < hi))), f"{name} shard size")
        s1, s2 = sharded.stats, whole.stats
        check(s1.num_elites == s2.num_elites and s1.obj_max == s2.obj_max and s1.coverage == s2.coverage, f"{name} stats")
        check(abs(float(s1.qd_score) - float(s2.qd_score)) <= 1e-9 * abs(float(s2.qd_score)), f"{name} qd_score")
        b1, b2 = sharded.best_elite, whole.best_elite
        check(all(np.array_equal(b1[k], b2[k]) for k in b2), f"{name} best_elite")
        probe = rng.uniform(-1, 1, (64, d))
        occ1, r1 = sharded.retrieve(probe)
        occ2, r2 = whole.retrieve(probe)
        check(np.array_equal(occ1, occ2) and all(np.array_equal(r1[k], r2[k], equal_nan=True) for k in r2),
              f"{name} retrieve")
        e = sharded.sample_elites(8)
        check(set(e["index"].tolist()) <= set(dw["index"].tolist()) and e["solution"].shape == (8, 5), f"{name} sample")
        bad = {k: v.copy() for k, v in mine.items()}
        bad["objective"][3] = np.nan if rank == 1 else 1.0  # only rank 1's slice is bad: every rank must drop the batch
        n_before = len(sharded)
        try:
            sharded.add(**bad)
            check(False, f"{name} NaN batch accepted")
        except ValueError:
            pass
        check(len(sharded) == n_before, f"{name} store touched by a dropped batch")
    q.put((rank, ok, why))
    dist.destroy_process_group()


def test_two_gpu_owner_computes_store():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_owned_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res
