"""GPU parity tests at the BASELINE.json shapes themselves (configs[2], [3], [4]) and regressions for the state
hand-over between the device-resident store and the host mirror (async adds, retessellate, checkpoint / resume).

Tolerances: indices, statuses, winners, occupied_list order bit-exact; values / thresholds / ES state 1e-6
relative (north_star), observed ~1e-12."""
import os
import pickle
import sys

import numpy as np
import pytest

from golden_util import assert_close, check_archive_dump, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import qd_oracle as O  # noqa: E402

from test_gpu_archives import gpu_dump  # noqa: E402

pytestmark = pytest.mark.gpu
RTOL = 1e-6


# ---- configs[4]: CVTArchive 1M centroids, 8-D ------------------------------------------------------------------
@pytest.mark.parametrize("method", ["auto", "tensor"])
def test_c5_shape_index_golden(method):
    """Cell indices of 20k seeded queries against the seeded 1M x 8 centroid set: the reference's KD-tree
    (tests/golden/c5_index.npz, oracle/gen_golden.py gen_c5_index) vs the box-tree and the tcgen05 search."""
    from pyribs_b200.archives import CVTArchive

    g = load_golden("c5_index")
    C, d, Q = int(g["C"]), int(g["d"]), int(g["Q"])
    cen = np.random.default_rng(int(g["centroids"])).uniform(-1, 1, (C, d))
    q = np.random.default_rng(int(g["queries"])).uniform(-1, 1, (Q, d))
    assert cen.sum() == float(g["checksum_centroids"]) and q.sum() == float(g["checksum_queries"]), \
        "NumPy regenerated different inputs than the golden was made from"
    a = CVTArchive(solution_dim=1, centroids=cen, ranges=[(-1, 1)] * d, search_method=method)
    assert a._pick_method(Q, C) == (3 if method == "auto" else 0)
    np.testing.assert_array_equal(a.index_of(q), g["index"])
    assert a._overflow_fallbacks == 0
    if method == "auto":  # a large batch takes the home-leaf-sorted walk; same answer in any order
        big = np.tile(q, (2, 1))[:30_000]
        np.testing.assert_array_equal(a.index_of(big), np.tile(g["index"], 2)[:30_000])


def test_c5_shape_add_vs_oracle():
    """Two 200k-row adds into the 1M-centroid archive against the oracle (SciPy KD-tree + the reference's add
    algebra): statuses bit-exact, values 1e-6, same cells occupied in the same order."""
    from pyribs_b200.archives import CVTArchive

    C, d, n, B = 1_000_000, 8, 10, 200_000
    cen = np.random.default_rng(0).uniform(-1, 1, (C, d))
    rng = np.random.default_rng(5)
    gpu = CVTArchive(solution_dim=n, centroids=cen, ranges=[(-1, 1)] * d)
    ora = O.make_cvt_oracle(solution_dim=n, centroids=cen, workers=-1)
    for step in range(2):
        b = dict(solution=rng.uniform(-1, 1, (B, n)), objective=rng.standard_normal(B) + 0.1 * step,
                 measures=rng.uniform(-1, 1, (B, d)))
        if step == 1:
            b["measures"][:50_000] = prev[:50_000] + rng.normal(0, 1e-3, (50_000, d))  # revisit occupied cells
        prev = b["measures"]
        g, o = gpu.add(**b), ora.add(**b)
        np.testing.assert_array_equal(g["status"], o["status"])
        assert_close(g["value"], o["value"], RTOL, f"step {step} value")
    assert len(gpu) == len(ora)
    np.testing.assert_array_equal(gpu._store.occupied_list, ora.store.occupied_list[: len(ora)])
    assert_close(gpu.stats.qd_score, ora.objective_sum, RTOL, "qd_score")


# ---- configs[2]: CMA-MAE GridArchive 500x500, 100 emitters x 512 ---------------------------------------------------
def test_c3_shape_add_vs_oracle():
    from pyribs_b200.archives import GridArchive

    rng = np.random.default_rng(33)
    E, lam, n = 100, 512, 100
    kw = dict(solution_dim=n, dims=(500, 500), ranges=[(-1, 1)] * 2, learning_rate=0.01, threshold_min=0.0)
    gpu, ora = GridArchive(**kw), O.make_grid_oracle(**kw)
    uniq = []
    for step in range(3):
        mu = rng.normal(0, 0.05, (E, 1, 2))
        meas = np.clip(mu + rng.normal(0, 0.02, (E, lam, 2)), -1, 1).reshape(E * lam, 2)
        b = dict(solution=rng.standard_normal((E * lam, n)), objective=rng.uniform(0, 100, E * lam) + 10.0 * step,
                 measures=meas)
        g, o = gpu.add(**b), ora.add(**b)
        np.testing.assert_array_equal(g["status"], o["status"], err_msg=f"step {step}")
        assert_close(g["value"], o["value"], RTOL, f"step {step} value")
        uniq.append(len(np.unique(ora.index_fn(meas))))
        assert len(gpu) == len(ora)
        np.testing.assert_array_equal(gpu._store.occupied_list, ora.store.occupied_list[: len(ora)])
        occ = ora.store.occupied
        assert_close(gpu._store._fields["threshold"].cpu().numpy()[occ], ora.store.fields["threshold"][occ], RTOL,
                     f"step {step} thresholds")
        np.testing.assert_array_equal(gpu._store._fields["objective"].cpu().numpy()[occ],
                                      ora.store.fields["objective"][occ])
        np.testing.assert_array_equal(gpu._store._fields["solution"].cpu().numpy()[occ],
                                      ora.store.fields["solution"][occ])
    assert all(2000 < u < 12000 for u in uniq), uniq  # ~10x same-cell collisions per batch, as configs[2] describes


# ---- configs[3]: sep-CMA-ES / LM-MA-ES, n = 10 000, batch 4096 -------------------------------------------------------
@pytest.mark.parametrize("es_name,kw", [("sep_cma_es", {}), ("lm_ma_es", {}), ("lm_ma_es", {"n_vectors": 20})])
def test_c4_shape_es_vs_oracle(es_name, kw):
    from pyribs_b200.emitters.opt import _get_es

    n, lam, mu = 10_000, 4096, 2048
    gpu = _get_es(es_name, sigma0=0.5, solution_dim=n, batch_size=lam, seed=1, dtype=np.float64,
                  lower_bounds=np.full(n, -np.inf), upper_bounds=np.full(n, np.inf), **kw)
    ora = O.ORACLE_ES[es_name](0.5, n, lam, seed=1, **kw)
    x0 = np.random.default_rng(0).standard_normal(n) * 0.1
    gpu.reset(x0), ora.reset(x0)
    for g in range(2):  # the second LM-MA ask applies the first direction vector
        z = np.random.default_rng(100 + g).standard_normal((lam, n))
        so = ora.ask(z)
        sg = gpu.ask(z=z)
        assert_close(sg, so, RTOL, f"{es_name} g{g} solutions")
        rank = np.argsort(np.sum((so - 1.0) ** 2, axis=1))
        ora.tell(rank, mu), gpu.tell(rank, None, mu)
        assert_close(gpu.mean, ora.mean, RTOL, f"{es_name} g{g} mean")
        assert_close(gpu.sigma, ora.sigma, RTOL, f"{es_name} g{g} sigma")
        assert_close(gpu.ps, ora.ps, RTOL, f"{es_name} g{g} ps")
        if es_name == "sep_cma_es":
            assert_close(gpu.pc, ora.pc, RTOL, f"g{g} pc")
            assert_close(gpu.cov, ora.C, RTOL, f"g{g} C")
    x = gpu.ask()  # the library's own stream at this size; the returned array is a read-only view
    assert x.shape == (lam, n) and not x.flags.writeable and np.all(np.isfinite(x))


# ---- device-resident state vs host mirror (async adds) ---------------------------------------------------------------
def test_async_add_then_data_then_stats():
    """add(CUDA tensors) only enqueues; whichever read comes first (data() here, not len()) must adopt the
    device-side state -- stats / best_elite afterwards are those of the batch."""
    import torch

    from pyribs_b200.archives import GridArchive

    rng = np.random.default_rng(4)
    kw = dict(solution_dim=3, dims=(20, 20), ranges=[(-1, 1)] * 2)
    gpu, ora = GridArchive(**kw), O.make_grid_oracle(**kw)
    b = dict(solution=rng.standard_normal((500, 3)), objective=rng.uniform(0, 5, 500), measures=rng.uniform(-1, 1, (500, 2)))
    gpu.add(**{k: torch.from_numpy(v).cuda() for k, v in b.items()})
    ora.add(**b)
    d = gpu.data()
    assert len(d["index"]) == len(ora)
    assert gpu.stats.num_elites == len(ora)
    assert_close(gpu.stats.qd_score, ora.objective_sum, 1e-12, "qd_score")
    assert gpu.best_elite is not None and gpu.best_elite["objective"] == b["objective"].max()
    # a batch the device drops (NaN objective) surfaces as the reference's ValueError at the first read
    bad = {k: torch.from_numpy(v).cuda() for k, v in b.items()}
    bad["objective"][7] = float("nan")
    gpu.add(**bad)
    with pytest.raises(ValueError, match="objective must be finite"):
        gpu.data()
    assert gpu.stats.num_elites == len(ora)  # nothing of the dropped batch was committed


@pytest.mark.parametrize("name", ["one_add", "three_adds"])
def test_retessellate_golden(name):
    """add -> retessellate -> best_elite / stats / data, then another add: the reference's trajectory
    (_grid_archive.py:926-976; tests/golden/retessellate.npz)."""
    from pyribs_b200.archives import GridArchive

    c = load_golden("retessellate")[name]
    a = GridArchive(solution_dim=3, dims=[10, 20], ranges=[(-1, 1), (-2, 2)])
    a.pipeline_commit = name == "three_adds"  # the replaced store must keep joining the side-stream row copy
    for k in sorted(c["in"]):
        a.add(**c["in"][k])
    check_archive_dump(gpu_dump(a), c["before"], RTOL, f"{name} before")
    a.retessellate([5, 8])
    assert a.best_elite is not None
    check_archive_dump(gpu_dump(a), c["after"], RTOL, f"{name} after")
    np.testing.assert_array_equal(a.data("index"), c["data_index"])
    np.testing.assert_array_equal(a.data("objective"), c["data_objective"])
    info = a.add(**c["again"]["in"])
    np.testing.assert_array_equal(info["status"], c["again"]["status"])
    assert_close(info["value"], c["again"]["value"], RTOL, "value after retessellate")
    check_archive_dump(gpu_dump(a), c["again"]["after"], RTOL, f"{name} again")


# ---- checkpoint / resume -------------------------------------------------------------------------------------
def _archive(kind):
    from pyribs_b200.archives import CVTArchive, GridArchive

    if kind == "grid_mae":
        return GridArchive(solution_dim=4, dims=(30, 30), ranges=[(-1, 1)] * 2, learning_rate=0.1, threshold_min=0.0,
                           extra_fields={"meta": ((2,), np.int32)}, seed=3)
    if kind == "grid_me":
        return GridArchive(solution_dim=4, dims=(30, 30), ranges=[(-1, 1)] * 2, seed=3)
    d = 2 if kind == "cvt2" else 6
    return CVTArchive(solution_dim=4, centroids=np.random.default_rng(1).uniform(-1, 1, (3000, d)),
                      ranges=[(-1, 1)] * d, seed=3)


@pytest.mark.parametrize("kind", ["grid_mae", "grid_me", "cvt2", "cvt6"])
def test_archive_pickle_round_trip(kind):
    rng = np.random.default_rng(8)
    a = _archive(kind)
    d = a.measure_dim

    def batch(j):
        b = dict(solution=rng.standard_normal((2000, 4)), objective=rng.uniform(0, 10, 2000) + (3 - j),
                 measures=rng.uniform(-1, 1, (2000, d)))
        if kind == "grid_mae":
            b["meta"] = rng.integers(0, 100, (2000, 2)).astype(np.int32)
        return b

    empty = pickle.loads(pickle.dumps(a))
    assert len(empty) == 0 and empty.best_elite is None and empty.stats.num_elites == 0
    bs = [batch(j) for j in range(4)]
    for b in bs[:2]:
        a.add(**b), empty.add(**b)
    twin = pickle.loads(pickle.dumps(a))
    for x in (twin, empty):
        assert len(x) == len(a)
        for f in ("num_elites", "coverage", "qd_score", "norm_qd_score", "obj_max", "obj_mean"):
            assert getattr(x.stats, f) == getattr(a.stats, f), f
        for k, v in a.best_elite.items():
            np.testing.assert_array_equal(x.best_elite[k], v)
    for b in bs[2:]:  # continue adding: the restored archive behaves like the one that never left the device
        ra, rt = a.add(**b), twin.add(**b)
        np.testing.assert_array_equal(ra["status"], rt["status"])
        np.testing.assert_array_equal(ra["value"], rt["value"])
    da, dt = a.data(), twin.data()
    for k in da:
        np.testing.assert_array_equal(da[k], dt[k], err_msg=k)
    assert a.stats.qd_score == twin.stats.qd_score and a.best_elite["objective"] == twin.best_elite["objective"]
    np.testing.assert_array_equal(a.sample_elites(5)["index"], twin.sample_elites(5)["index"])  # same generator state


@pytest.mark.parametrize("es_name", ["sep_cma_es", "lm_ma_es", "cma_es", "openai_es"])
def test_es_pickle_round_trip(es_name):
    from pyribs_b200.emitters.opt import _get_es

    n, lam = 40, 16
    es = _get_es(es_name, sigma0=0.5, solution_dim=n, batch_size=lam, seed=5, dtype=np.float64,
                 lower_bounds=np.full(n, -np.inf), upper_bounds=np.full(n, np.inf))
    es.reset(np.zeros(n))

    def generation(e):
        x = np.array(e.ask())
        fit = -np.sum((x - 1.0) ** 2, axis=1)
        order = np.argsort(-fit)
        e.tell(order, fit[order], lam // 2)
        return x

    for _ in range(3):
        generation(es)
    twin = pickle.loads(pickle.dumps(es))
    for _ in range(3):  # same state, same Philox counter -> the same samples from here on
        np.testing.assert_array_equal(generation(es), generation(twin))


def test_emitter_and_scheduler_pickle():
    from pyribs_b200.archives import GridArchive
    from pyribs_b200.emitters import EvolutionStrategyEmitter
    from pyribs_b200.schedulers import Scheduler

    def make():
        a = GridArchive(solution_dim=6, dims=(20, 20), ranges=[(-3, 3)] * 2, seed=1)
        ems = [EvolutionStrategyEmitter(a, x0=np.zeros(6), sigma0=0.5, batch_size=8, seed=s, es="sep_cma_es")
               for s in (1, 2)]
        return Scheduler(a, ems)

    def it(s):
        x = s.ask()
        s.tell(-np.sum(x**2, axis=1), x[:, :2])
        return np.array(x)

    s = make()
    for _ in range(3):
        it(s)
    twin = pickle.loads(pickle.dumps(s))
    for _ in range(3):
        np.testing.assert_array_equal(it(s), it(twin))
    assert s.archive.stats.qd_score == twin.archive.stats.qd_score


# ---- near-ties: the winner of a cell is the first row holding the EXACT maximum -----------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("B", [64, 3000, 300_000])
def test_near_tie_objectives_resolve_exactly(B, dtype):
    """The insert kernel picks a cell's winner through a key truncated to make room for the row number
    (csrc/insert.cu `pick`); rows of one cell whose objectives differ only in the bits below the truncation, and
    exact ties, must still resolve like the reference (highest objective, first in batch: _grid_archive.py:689-696).
    B = 64 is list mode (cells >> rows), the others cell mode with 6 / 12 / 19 row bits."""
    from pyribs_b200.archives import GridArchive

    rng = np.random.default_rng(B)
    kw = dict(solution_dim=2, dims=(30, 30), ranges=[(-1, 1)] * 2, dtype=dtype)
    for mae in (False, True):
        kw2 = dict(kw, learning_rate=0.1, threshold_min=-1.0) if mae else kw
        gpu, ora = GridArchive(**kw2), O.make_grid_oracle(**kw2)
        meas = rng.uniform(-1, 1, (B, 2))
        base = rng.choice([1.0, 2.5, -3.0, 1e-3, 7e5], B)
        ulps = rng.integers(0, 4, B)  # 0..3 ulps above the base value: ties and near-ties galore
        obj = base.astype(dtype)
        for _ in range(3):
            obj = np.where(ulps > 0, np.nextafter(obj, dtype(np.inf)), obj)
            ulps = ulps - 1
        b = dict(solution=rng.standard_normal((B, 2)).astype(dtype), objective=obj.astype(dtype),
                 measures=meas.astype(dtype))
        for step in range(2):
            g, o = gpu.add(**b), ora.add(**b)
            np.testing.assert_array_equal(g["status"], o["status"])
            assert_close(g["value"], o["value"], RTOL, "value")
            occ = ora.store.occupied
            np.testing.assert_array_equal(gpu._store._fields["objective"].cpu().numpy()[occ],
                                          ora.store.fields["objective"][occ])
            # the solution row identifies WHICH of the tied rows won
            np.testing.assert_array_equal(gpu._store._fields["solution"].cpu().numpy()[occ],
                                          ora.store.fields["solution"][occ])
            np.testing.assert_array_equal(gpu._store.occupied_list, ora.store.occupied_list[: len(ora)])
            b = dict(b, objective=np.nextafter(b["objective"], dtype(np.inf)))  # everything improves by one ulp


# ---- lean host routes equal the general one -------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["grid_mae", "grid_me", "cvt2", "cvt6"])
@pytest.mark.parametrize("B", [1, 100, 3000, 20_000])
def test_lean_add_routes_equal_the_general_route(kind, B):
    """CUDA tensors through `_add_device`, NumPy batches through `_add_small`, and both through the general route
    (`_dev_fast = _small_adds = False`): same status / value bits, same store."""
    import torch

    rng = np.random.default_rng(B)
    archives = [_archive(kind) for _ in range(4)]
    archives[2]._dev_fast = False
    archives[3]._small_adds = False
    d = archives[0].measure_dim
    for step in range(3):
        b = dict(solution=rng.standard_normal((B, 4)), objective=rng.uniform(0, 10, B) + step,
                 measures=rng.uniform(-1, 1, (B, d)))
        if kind == "grid_mae":
            b["meta"] = rng.integers(0, 100, (B, 2)).astype(np.int32)
        dev = {k: torch.from_numpy(v).cuda() for k, v in b.items()}
        outs = [archives[0].add(**dev), archives[1].add(**b), archives[2].add(**dev), archives[3].add(**b)]
        outs = [{k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in o.items()} for o in outs]
        for o in outs[1:]:
            np.testing.assert_array_equal(o["status"], outs[0]["status"])
            np.testing.assert_array_equal(o["value"], outs[0]["value"])
    ref = archives[0].data()
    for a in archives[1:]:
        da = a.data()
        for k in ref:
            np.testing.assert_array_equal(da[k], ref[k], err_msg=k)
        assert a.stats.qd_score == archives[0].stats.qd_score
    # what falls back: wrong dtype / non-contiguous tensors still work (general route), wrong shapes still raise
    a = archives[0]
    b32 = {k: (v.float() if v.dtype == torch.float64 else v) for k, v in dev.items()}
    a.add(**b32)
    with pytest.raises(ValueError):
        a.add(**dict(dev, objective=dev["objective"][:-1]))
    assert len(a) > 0


@pytest.mark.parametrize("pinned", [False, True])
def test_chunked_host_search_equals_device_batch(pinned):
    """Host batches of >= 2^18 rows overlap the transfer of their measures with the search, piece by piece
    (DeviceArchive._search_chunked): same cells, statuses and values as the same batch handed over as CUDA tensors."""
    import torch

    from pyribs_b200.archives import CVTArchive

    rng = np.random.default_rng(21)
    cen = rng.uniform(-1, 1, (20_000, 6))
    a, b = (CVTArchive(solution_dim=3, centroids=cen, ranges=[(-1, 1)] * 6) for _ in range(2))
    B = 300_001
    for step in range(2):
        batch = dict(solution=rng.standard_normal((B, 3)), objective=rng.standard_normal(B) + step,
                     measures=rng.uniform(-1.1, 1.1, (B, 6)))
        host = {k: (torch.from_numpy(v).pin_memory().numpy() if pinned else v) for k, v in batch.items()}
        ra = a.add(**host)
        rb = b.add(**{k: torch.from_numpy(v).cuda() for k, v in batch.items()})
        np.testing.assert_array_equal(ra["status"], rb["status"].cpu().numpy())
        np.testing.assert_array_equal(ra["value"], rb["value"].cpu().numpy())
    da, db = a.data(), b.data()
    for k in da:
        np.testing.assert_array_equal(da[k], db[k], err_msg=k)
    bad = dict(host, measures=host["measures"].copy())
    bad["measures"][B - 7, 2] = np.inf  # in the last piece
    with pytest.raises(ValueError, match="measures must be finite"):
        a.add(**bad)


def test_batch_of_2_pow_24_rows_uses_the_separate_row_counter():
    """From 2^24 rows on the per-cell row count no longer fits the low bits of the fixed-point sum word and lives
    in its own array (csrc/insert.cu FxConst::b == 0, Scratch::cnt); the row number in `pick` takes 25 bits. Same
    statuses, values, thresholds and elites as the oracle."""
    from pyribs_b200.archives import GridArchive

    B = (1 << 24) + 3
    rng = np.random.default_rng(24)
    kw = dict(solution_dim=1, dims=(16, 16), ranges=[(-1, 1)] * 2, learning_rate=0.5, threshold_min=-1.0)
    gpu, ora = GridArchive(**kw), O.make_grid_oracle(**kw)
    for step in range(2):
        b = dict(solution=rng.standard_normal((B, 1)), objective=rng.uniform(0, 1, B) + 0.25 * step,
                 measures=rng.uniform(-1, 1, (B, 2)))
        b["objective"][B - 1] = b["objective"][5] = 7.0  # a tie between the first rows and the very last one
        b["measures"][B - 1] = b["measures"][5]
        g, o = gpu.add(**b), ora.add(**b)
        np.testing.assert_array_equal(g["status"], o["status"])
        assert_close(g["value"], o["value"], RTOL, f"step {step} value")
        occ = ora.store.occupied
        assert_close(gpu._store._fields["threshold"].cpu().numpy()[occ], ora.store.fields["threshold"][occ], RTOL,
                     f"step {step} thresholds")
        np.testing.assert_array_equal(gpu._store._fields["solution"].cpu().numpy()[occ], ora.store.fields["solution"][occ])
    assert len(gpu) == len(ora) == 256
