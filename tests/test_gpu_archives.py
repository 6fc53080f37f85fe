"""GPU parity tests: the CUDA path (through the C ABI) vs golden vectors and the oracle."""
import os
import sys

import numpy as np
import pytest

from golden_util import assert_close, check_archive_dump, load_golden, steps_of

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import qd_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu

RTOL = 1e-6  # north_star: objectives / thresholds / add values within 1e-6 relative (fp64)
RTOL_F32 = 2e-5  # float32 archives: a few float32 ulps


def gpu_dump(a):
    st = a._store
    occ = st.occupied.copy()
    d = {"occupied": occ, "occupied_list": np.array(st.occupied_list), "n": len(st)}
    for name, t in st._fields.items():
        x = t.cpu().numpy().copy()
        x[~occ] = 0
        d["field_" + name] = x
    s = a.stats
    d["stats"] = {k: getattr(s, k) for k in ("num_elites", "coverage", "qd_score", "norm_qd_score", "obj_max",
                                             "obj_mean")}
    if a.best_elite is not None:
        d["best_elite"] = a.best_elite
    return d


def make_gpu_archive(cfg, **kw):
    from pyribs_b200.archives import CVTArchive, GridArchive

    dtype = np.float32 if int(cfg["f32"]) else np.float64
    lr = None if np.isnan(cfg["lr"]) else float(cfg["lr"])
    args = dict(learning_rate=lr, threshold_min=float(cfg["tmin"]), dtype=dtype,
                qd_score_offset=float(cfg.get("offset", 0.0)))
    if int(cfg.get("extras", 0)):
        args["extra_fields"] = {"meta": ((3,), np.int32), "score": ((), np.float32)}
    if int(cfg["kind"]) == 0:
        return GridArchive(solution_dim=int(cfg["n"]), dims=[int(v) for v in cfg["dims"]],
                           ranges=[tuple(r) for r in cfg["ranges"]], **args)
    d = cfg["centroids"].shape[1]
    return CVTArchive(solution_dim=int(cfg["n"]), centroids=cfg["centroids"], ranges=[(-1, 1)] * d, **args, **kw)


def test_grid_index_golden():
    from pyribs_b200.archives import GridArchive

    g = load_golden("grid_index")
    for name, c in g.items():
        dt = c["measures"].dtype
        a = GridArchive(solution_dim=1, dims=[int(v) for v in c["dims"]], ranges=[tuple(r) for r in c["ranges"]],
                        dtype=dt)
        np.testing.assert_array_equal(a.index_of(c["measures"]), c["index"], err_msg=name)


@pytest.mark.parametrize("method", ["scan", "tensor", "bucket", "tree"])
def test_cvt_index_golden(method):
    from pyribs_b200.archives import CVTArchive

    g = load_golden("cvt_index")
    for name, c in g.items():
        dt = c["measures"].dtype
        d = c["centroids"].shape[1]
        if method == "bucket" and d > 3:
            continue
        a = CVTArchive(solution_dim=1, centroids=c["centroids"], ranges=[(-1, 1)] * d, dtype=dt, search_method=method)
        try:
            idx = a.index_of(c["measures"])
        except Exception as e:  # noqa: BLE001
            if "not built" in str(e):
                pytest.skip("tensor-core search not built yet")
            raise
        assert a._overflow_fallbacks == 0
        if "index" in c:
            np.testing.assert_array_equal(idx, c["index"], err_msg=f"{name} ({method}) vs KD-tree")
        if "index_brute" in c:
            np.testing.assert_array_equal(idx, c["index_brute"], err_msg=f"{name} ({method}) vs brute force")


@pytest.mark.parametrize("name", sorted(load_golden("add_traj").keys()))
def test_add_golden(name):
    case = load_golden("add_traj")[name]
    cfg = case["cfg"]
    a = make_gpu_archive(cfg)
    rtol = RTOL_F32 if int(cfg["f32"]) else RTOL
    single = int(cfg.get("single", 0))
    for j, s in enumerate(steps_of(case)):
        b = s["in"]
        if single:
            infos = [a.add_single(**{k: v[i] for k, v in b.items()}) for i in range(len(b["objective"]))]
            status = np.array([x["status"] for x in infos], dtype=np.int32)
            value = np.array([x["value"] for x in infos])
        else:
            info = a.add(**b)
            status, value = info["status"], info["value"]
            assert status.dtype == np.int32 and value.dtype == b["objective"].dtype
        np.testing.assert_array_equal(status, s["status"], err_msg=f"{name} s{j} status")
        assert_close(value, s["value"], rtol, f"{name} s{j} value")
        check_archive_dump(gpu_dump(a), s["after"], rtol, f"{name} s{j}")


@pytest.mark.parametrize("case", ["uniform", "clustered", "degenerate_axis", "one_centroid", "lattice_ties", "f32"])
@pytest.mark.parametrize("d", [1, 2, 3])
def test_cvt_bucket_equals_scan(case, d):
    """The bucket-grid search (measure_dim <= 3) must return exactly what the fp64 scan returns:
    empty buckets, clustered centroids, a zero-extent axis, queries outside the bounding box,
    exact ties on a lattice (lowest index wins) and float32 archives."""
    from pyribs_b200.archives import CVTArchive

    rng = np.random.default_rng(d * 100 + len(case))
    C, B = 5000, 20000
    dt = np.float32 if case == "f32" else np.float64
    if case == "clustered":
        cen = np.concatenate([rng.normal(m, 0.01, (C // 5, d)) for m in rng.uniform(-1, 1, (5, d))])
    elif case == "one_centroid":
        cen = rng.uniform(-1, 1, (1, d))
    elif case == "lattice_ties":
        side = int(round(C ** (1.0 / d)))
        cen = np.stack(np.meshgrid(*[np.arange(side) / 4.0] * d, indexing="ij"), -1).reshape(-1, d)
        cen = np.concatenate([cen, cen[::7]])  # duplicates: the first copy must win
    else:
        cen = rng.uniform(-1, 1, (C, d))
    if case == "degenerate_axis":
        cen[:, -1] = 0.25
    span = np.abs(cen).max() + 0.1
    meas = rng.uniform(-1.2, 1.2, (B, d)) * span
    if case == "lattice_ties":
        meas[: B // 2] = np.round(meas[: B // 2] * 8) / 8  # many exactly equidistant queries
    meas[:20] *= 50.0
    meas[20:25] *= 1e6
    cen, meas = cen.astype(dt), meas.astype(dt)
    out = {}
    for method in ("scan", "bucket"):
        a = CVTArchive(solution_dim=1, centroids=cen, ranges=[(-1, 1)] * d, search_method=method, dtype=dt)
        out[method] = a.index_of(meas)
    np.testing.assert_array_equal(out["bucket"], out["scan"])
    if dt == np.float64:
        sub = rng.choice(B, 300, replace=False)
        dd = ((meas[sub, None, :] - cen[None, :, :]) ** 2).sum(-1)
        np.testing.assert_array_equal(out["bucket"][sub], dd.argmin(1))


@pytest.mark.parametrize("case", ["uniform", "clustered", "degenerate_axis", "one_centroid", "few", "lattice_ties", "f32",
                                  "all_equal"])
@pytest.mark.parametrize("d", [2, 4, 8, 9, 20])
def test_cvt_tree_equals_scan(case, d):
    """The bounding-box tree search (csrc/cvt_tree.cu, the default for 3 < measure_dim <= 16) must return exactly
    what the fp64 scan returns: clustered centroids, a zero-extent axis, a single leaf, queries far outside the
    boxes (beyond fp32 range too), exact ties on a lattice and between duplicates (lowest index wins), float32
    archives, compile-time (2, 4, 8) and run-time (9, 20) measure dimensions."""
    from pyribs_b200.archives import CVTArchive

    rng = np.random.default_rng(d * 100 + len(case))
    C, B = 6000, 8000
    dt = np.float32 if case == "f32" else np.float64
    if case == "clustered":
        cen = np.concatenate([rng.normal(m, 0.01, (C // 5, d)) for m in rng.uniform(-1, 1, (5, d))])
    elif case == "one_centroid":
        cen = rng.uniform(-1, 1, (1, d))
    elif case == "few":
        cen = rng.uniform(-1, 1, (33, d))
    elif case == "all_equal":
        cen = np.tile(rng.uniform(-1, 1, (1, d)), (700, 1))
    elif case == "lattice_ties":
        side = max(2, int(round(C ** (1.0 / d))))
        cen = np.stack(np.meshgrid(*[np.arange(side) / 4.0] * d, indexing="ij"), -1).reshape(-1, d)[:20000]
        cen = np.concatenate([cen, cen[::7]])  # duplicates: the first copy must win
    else:
        cen = rng.uniform(-1, 1, (C, d))
    if case == "degenerate_axis":
        cen[:, -1] = 0.25
    span = np.abs(cen).max() + 0.1
    meas = rng.uniform(-1.2, 1.2, (B, d)) * span
    if case == "lattice_ties":
        meas[: B // 2] = np.round(meas[: B // 2] * 8) / 8  # many exactly equidistant queries
    meas[:20] *= 50.0
    meas[20:25] *= 1e6
    if dt == np.float64:
        meas[25:28] *= 1e40  # finite in fp64, beyond fp32: the box bounds saturate and the walk degrades to a scan
    meas[30:40] = cen[:10] if len(cen) >= 10 else cen[0]
    cen, meas = cen.astype(dt), meas.astype(dt)
    out = {}
    for method in ("scan", "tree"):
        a = CVTArchive(solution_dim=1, centroids=cen, ranges=[(-1, 1)] * d, search_method=method, dtype=dt)
        out[method] = a.index_of(meas)
    np.testing.assert_array_equal(out["tree"], out["scan"])
    if dt == np.float64:
        sub = rng.choice(B, 200, replace=False)
        dd = ((meas[sub, None, :] - cen[None, :, :]) ** 2).sum(-1)
        np.testing.assert_array_equal(out["tree"][sub], dd.argmin(1))
    a = CVTArchive(solution_dim=1, centroids=cen, ranges=[(-1, 1)] * d, dtype=dt)
    assert a._pick_method(B, len(cen)) == (2 if d <= 3 else 3 if d <= 10 else (0 if B * len(cen) >= 1 << 26 else 1))
    bad = meas.copy()
    bad[5, 0] = np.nan
    with pytest.raises(ValueError, match="measures must be finite"):
        CVTArchive(solution_dim=1, centroids=cen, ranges=[(-1, 1)] * d, search_method="tree", dtype=dt).index_of(bad)


def _random_batches(rng, steps, B, n, md, scale):
    for step in range(steps):
        yield dict(solution=rng.standard_normal((B, n)), objective=rng.normal(step, 3, B),
                   measures=rng.uniform(-1.3, 1.3, (B, md)) * scale)


@pytest.mark.parametrize("mode", ["me", "mae"])
@pytest.mark.parametrize("kind", ["grid", "cvt"])
def test_add_vs_oracle_random(kind, mode):
    """Fresh seeded inputs at sizes the oracle finishes in seconds (includes heavy collisions)."""
    from pyribs_b200.archives import CVTArchive, GridArchive

    rng = np.random.default_rng(11)
    lr, tmin = (None, -np.inf) if mode == "me" else (0.02, -1.5)
    n = 16
    if kind == "grid":
        gpu = GridArchive(solution_dim=n, dims=(40, 60), ranges=[(-1, 1), (-2, 2)], learning_rate=lr,
                          threshold_min=tmin, seed=5)
        ora = O.make_grid_oracle(solution_dim=n, dims=(40, 60), ranges=[(-1, 1), (-2, 2)], learning_rate=lr,
                                 threshold_min=tmin, seed=5)
        md, scale = 2, np.array([1.0, 2.0])
    else:
        cen = rng.uniform(-1, 1, (3000, 4))
        gpu = CVTArchive(solution_dim=n, centroids=cen, ranges=[(-1, 1)] * 4, learning_rate=lr, threshold_min=tmin,
                         seed=5)
        ora = O.make_cvt_oracle(solution_dim=n, centroids=cen, learning_rate=lr, threshold_min=tmin, seed=5)
        md, scale = 4, np.ones(4)
    for j, b in enumerate(_random_batches(rng, 4, 20000, n, md, scale)):
        g = gpu.add(**b)
        o = ora.add(**b)
        np.testing.assert_array_equal(g["status"], o["status"], err_msg=f"step {j}")
        assert_close(g["value"], o["value"], RTOL, f"step {j} value")
        nocc = len(ora)
        assert len(gpu) == nocc
        np.testing.assert_array_equal(gpu._store.occupied_list, ora.store.occupied_list[:nocc])
        occ = ora.store.occupied
        np.testing.assert_array_equal(gpu._store.occupied, occ)
        for name in ("solution", "objective", "measures"):
            np.testing.assert_array_equal(gpu._store._fields[name].cpu().numpy()[occ], ora.store.fields[name][occ])
        assert_close(gpu._store._fields["threshold"].cpu().numpy()[occ], ora.store.fields["threshold"][occ], RTOL)
        so = ora.stats()
        assert_close(gpu.stats.qd_score, so["qd_score"], RTOL, "qd_score")
        assert gpu.stats.obj_max == so["obj_max"]
        e1, e2 = gpu.sample_elites(7), ora.sample_elites(7)
        np.testing.assert_array_equal(e1["solution"], e2["solution"])
        np.testing.assert_array_equal(e1["index"], e2["index"])


def test_add_is_deterministic():
    """Two identical archives fed the same batches end bit-identical (binned sums)."""
    from pyribs_b200.archives import GridArchive

    rng = np.random.default_rng(3)
    batches = [dict(solution=rng.standard_normal((30000, 8)), objective=rng.uniform(0, 100, 30000),
                    measures=rng.normal(0, 0.1, (30000, 2))) for _ in range(3)]
    outs = []
    for _ in range(2):
        a = GridArchive(solution_dim=8, dims=(50, 50), ranges=[(-1, 1)] * 2, learning_rate=0.01, threshold_min=0.0)
        for b in batches:
            a.add(**b)
        outs.append((a._store._fields["threshold"].cpu().numpy().copy(), a._store.occupied, float(a.stats.qd_score)))
    occ = outs[0][1]
    np.testing.assert_array_equal(outs[0][0][occ], outs[1][0][occ])
    assert outs[0][2] == outs[1][2]


def test_error_contract():
    from pyribs_b200.archives import GridArchive

    a = GridArchive(solution_dim=3, dims=[10, 20], ranges=[(-1, 1), (-2, 2)])
    a.add([[1, 2, 3]], [1.0], [[0.25, 0.25]])
    before = gpu_dump(a)
    with pytest.raises(ValueError):
        a.add([[1, 2, 3]], [np.inf], [[0.0, 0.0]])
    with pytest.raises(ValueError):
        a.add([[1, 2, 3], [1, 2, 3]], [5.0, 6.0], [[0.0, np.nan], [0.5, 0.5]])
    with pytest.raises(ValueError):
        a.add([[1, 2, 3]], [1.0, 2.0], [[0.0, 0.0]])
    with pytest.raises(ValueError):
        a.add([[1, 2]], [1.0], [[0.0, 0.0]])
    with pytest.raises(ValueError):
        a.index_of([[np.nan, 0.0]])
    after = gpu_dump(a)
    np.testing.assert_array_equal(before["occupied"], after["occupied"])
    np.testing.assert_array_equal(before["field_objective"], after["field_objective"])
    assert len(a) == 1
    # the store still works after the aborted batches (scratch left clean)
    info = a.add([[4, 5, 6], [7, 8, 9]], [2.0, 3.0], [[0.25, 0.25], [0.25, 0.25]])
    np.testing.assert_array_equal(info["status"], [1, 1])
    assert a.best_elite["objective"] == 3.0
    np.testing.assert_array_equal(a.best_elite["solution"], [7, 8, 9])
    with pytest.raises(IndexError):
        GridArchive(solution_dim=3, dims=[4], ranges=[(-1, 1)]).sample_elites(1)
    e = GridArchive(solution_dim=3, dims=[4], ranges=[(-1, 1)])
    out = e.add(np.zeros((0, 3)), np.zeros(0), np.zeros((0, 1)))
    assert out["status"].shape == (0,) and out["value"].shape == (0,)


def test_device_tensor_inputs_stay_on_device():
    import torch

    from pyribs_b200.archives import GridArchive

    a = GridArchive(solution_dim=4, dims=[8, 8], ranges=[(-1, 1)] * 2)
    g = torch.Generator(device="cuda").manual_seed(0)
    sol = torch.randn(1000, 4, device="cuda", dtype=torch.float64, generator=g)
    obj = torch.randn(1000, device="cuda", dtype=torch.float64, generator=g)
    meas = torch.rand(1000, 2, device="cuda", dtype=torch.float64, generator=g) * 2 - 1
    info = a.add(sol, obj, meas)
    assert info["status"].is_cuda and info["value"].is_cuda
    b = GridArchive(solution_dim=4, dims=[8, 8], ranges=[(-1, 1)] * 2)
    info2 = b.add(sol.cpu().numpy(), obj.cpu().numpy(), meas.cpu().numpy())
    np.testing.assert_array_equal(info["status"].cpu().numpy(), info2["status"])
    np.testing.assert_array_equal(a.data("objective"), b.data("objective"))


def test_retrieve_and_iteration():
    from pyribs_b200.archives import GridArchive

    a = GridArchive(solution_dim=3, dims=[10, 20], ranges=[(-1, 1), (-2, 2)])
    a.add([[1, 2, 3], [4, 5, 6]], [1.0, 2.0], [[0.25, 0.25], [-0.5, -0.5]])
    occ, data = a.retrieve([[0.25, 0.25], [0.9, 0.9]])
    np.testing.assert_array_equal(occ, [True, False])
    assert data["objective"][0] == 1.0 and np.isnan(data["objective"][1]) and data["index"][1] == -1
    rows = list(a)
    assert len(rows) == 2 and {float(r["objective"]) for r in rows} == {1.0, 2.0}
    it = iter(a)
    next(it)
    a.add([[1, 2, 3]], [9.0], [[0.0, 0.0]])
    with pytest.raises(RuntimeError):
        next(it)
    df = a.data(return_type="pandas")
    assert list(df.columns)[:4] == ["solution_0", "solution_1", "solution_2", "objective"]
    a.clear()
    assert len(a) == 0 and a.stats.obj_max is None


def test_device_adds_are_asynchronous_and_state_lives_on_device():
    """CUDA-tensor insertions are only enqueued: len / stats / best_elite / occupied_list read the
    device-side running state back lazily and must equal what the same NumPy batches give (which
    wait after every call); a non-finite batch is dropped on the device and surfaces as the
    reference's ValueError at the next read, leaving the archive untouched."""
    import torch

    from pyribs_b200.archives import CVTArchive, GridArchive

    rng = np.random.default_rng(7)
    cen = rng.uniform(-1, 1, (500, 2))
    for make in (lambda: GridArchive(solution_dim=5, dims=[20, 20], ranges=[(-1, 1)] * 2, learning_rate=0.1,
                                     threshold_min=-1.0),
                 lambda: GridArchive(solution_dim=5, dims=[20, 20], ranges=[(-1, 1)] * 2),
                 lambda: CVTArchive(solution_dim=5, centroids=cen, ranges=[(-1, 1)] * 2)):
        a, b = make(), make()
        batches = [dict(solution=rng.standard_normal((3000, 5)), objective=rng.normal(j, 2.0, 3000),
                        measures=rng.uniform(-1.1, 1.1, (3000, 2))) for j in range(6)]
        outs = []
        for bt in batches:  # no host read in between
            outs.append(a.add(*(torch.from_numpy(bt[k]).cuda() for k in ("solution", "objective", "measures"))))
        assert a._store._pending
        for bt, o in zip(batches, outs):
            r = b.add(**bt)
            np.testing.assert_array_equal(o["status"].cpu().numpy(), r["status"])
            np.testing.assert_array_equal(o["value"].cpu().numpy(), r["value"])
        assert len(a) == len(b) and not a._store._pending
        sa, sb = a.stats, b.stats
        assert sa.num_elites == sb.num_elites and sa.qd_score == sb.qd_score and sa.obj_max == sb.obj_max
        assert sa.obj_mean == sb.obj_mean
        for k in a.best_elite:
            np.testing.assert_array_equal(a.best_elite[k], b.best_elite[k])
        np.testing.assert_array_equal(a._store.occupied_list, b._store.occupied_list)
        da, db = a.data(), b.data()
        for k in da:
            np.testing.assert_array_equal(da[k], db[k])
        # a bad batch between good ones: dropped, reported once, later batches still land
        bad = {k: torch.from_numpy(v).cuda() for k, v in batches[0].items()}
        bad["objective"] = bad["objective"].clone()
        bad["objective"][17] = float("nan")
        n0 = len(a)
        a.add(bad["solution"], bad["objective"], bad["measures"])
        good = dict(batches[5])
        good["objective"] = good["objective"] + 50.0
        a.add(*(torch.from_numpy(good[k]).cuda() for k in ("solution", "objective", "measures")))
        with pytest.raises(ValueError, match="objective"):
            len(a)
        b.add(**good)
        assert len(a) == len(b) >= n0
        np.testing.assert_array_equal(a.data("objective"), b.data("objective"))
        a.clear()
        assert len(a) == 0 and a.stats.obj_max is None and a.best_elite is None
        a.add(**batches[0])
        c = make()
        c.add(**batches[0])
        np.testing.assert_array_equal(a.data("threshold"), c.data("threshold"))
        assert a.stats.qd_score == c.stats.qd_score


@pytest.mark.parametrize("shape", [(200_000, 10_000, 2), (60_000, 100_000, 8), (30_000, 50_000, 5), (3_000, 300_000, 3),
                                   (20_000, 20_000, 13), (20_000, 20_000, 24)])
def test_cvt_tensor_equals_scan(shape):
    """The tcgen05 filter + fp64 rescoring must be bit-identical to the fp64 scan, including
    far-out-of-range queries (per-row rescaling / outlier path) and duplicated centroids."""
    import torch

    from pyribs_b200.archives import CVTArchive

    B, C, d = shape
    rng = np.random.default_rng(B + C + d)
    cen = rng.uniform(-1, 1, (C, d)) * rng.choice([1.0, 3.0, 0.01])
    cen[C // 2: C // 2 + 100] = cen[:100]  # exact duplicates -> lowest index must win
    meas = rng.uniform(-1.2, 1.2, (B, d)) * np.abs(cen).max()
    meas[:50] *= 1e3
    meas[50:60] *= 1e9
    meas[60:70] = cen[C // 2: C // 2 + 10]
    out = {}
    for method in ("scan", "tensor", "tree") + (("bucket",) if d <= 3 else ()):
        a = CVTArchive(solution_dim=1, centroids=cen, ranges=[(-1, 1)] * d, search_method=method)
        out[method] = a.index_of(meas)
        assert a._overflow_fallbacks == 0
    np.testing.assert_array_equal(out["tensor"], out["scan"])
    np.testing.assert_array_equal(out["tree"], out["scan"])
    if "bucket" in out:
        np.testing.assert_array_equal(out["bucket"], out["scan"])
    assert np.all(out["scan"][60:70] == np.arange(10))
    # spot-check against an fp64 NumPy brute force
    sub = rng.choice(B, 200, replace=False)
    dd = ((meas[sub, None, :] - cen[None, :, :]) ** 2).sum(-1)
    np.testing.assert_array_equal(out["tensor"][sub], dd.argmin(1))


def test_cma_mae_batch_beyond_packed_count_range():
    """Batches of >= 2^24 rows take the separate-counter variant of the fixed-point cell sums
    (csrc/insert.cu make_fx): 16.8M rows into 64 cells, checked against NumPy's sequential per-cell
    sums (np.bincount adds in batch order, like the reference's numba loop) and first-wins argmax."""
    import torch

    from pyribs_b200.archives import GridArchive

    B, lr = (1 << 24) + 3, 1e-6
    rng = np.random.default_rng(5)
    meas = rng.uniform(-1, 1, (B, 2))
    obj = np.round(rng.uniform(0.5, 100, B), 3)  # rounded: exact ties inside cells do occur
    a = GridArchive(solution_dim=1, dims=(8, 8), ranges=[(-1, 1)] * 2, learning_rate=lr, threshold_min=0.0)
    sol = np.arange(B, dtype=np.float64)[:, None]
    out = a.add(torch.from_numpy(sol).cuda(), torch.from_numpy(obj).cuda(), torch.from_numpy(meas).cuda())
    idx = a.index_of(meas)
    assert len(a) == 64
    assert bool((out["status"] == 2).all().item())
    np.testing.assert_array_equal(out["value"].cpu().numpy(), obj)  # threshold_min = 0
    k = np.bincount(idx, minlength=64).astype(np.float64)
    S = np.bincount(idx, weights=obj, minlength=64)
    ratio = (1.0 - lr) ** k
    thr = ratio * 0.0 + (S / k) * (1.0 - ratio)
    d = a.data()
    order = np.argsort(d["index"])
    np.testing.assert_allclose(d["threshold"][order], thr, rtol=1e-9)
    best = np.full(64, -np.inf)
    np.maximum.at(best, idx, obj)
    np.testing.assert_array_equal(d["objective"][order], best)
    first = np.full(64, B, dtype=np.int64)
    cand = np.nonzero(obj == best[idx])[0]
    np.minimum.at(first, idx[cand], cand)
    np.testing.assert_array_equal(d["solution"][order, 0], first.astype(np.float64))


def test_zero_copy_rows_from_pinned_host_arrays():
    """When at most cells << batch rows can win, add() of page-locked NumPy arrays does not ship the solution /
    extra-field rows: the insert kernel reads the winners' rows straight from host memory. Same archive as the
    DMA path and as device inputs."""
    import torch

    from pyribs_b200.archives import CVTArchive, GridArchive

    rng = np.random.default_rng(3)
    cen = rng.uniform(-1, 1, (300, 2))
    for make in (lambda: CVTArchive(solution_dim=7, centroids=cen, ranges=[(-1, 1)] * 2, extra_fields={"tag": ((), np.int32)}),
                 lambda: GridArchive(solution_dim=7, dims=[12, 12], ranges=[(-1, 1)] * 2, learning_rate=0.2,
                                     threshold_min=-2.0, extra_fields={"tag": ((), np.int32)})):
        a, b, c = make(), make(), make()
        b._zero_copy_rows = False
        a._small_adds = b._small_adds = False  # (c: batches of this size are staged in one page-locked block, _add_small)
        for j in range(4):
            B = 5000
            pin = dict(solution=torch.from_numpy(rng.standard_normal((B, 7))).pin_memory(),
                       objective=torch.from_numpy(rng.normal(j, 1.0, B)).pin_memory(),
                       measures=torch.from_numpy(rng.uniform(-1, 1, (B, 2))).pin_memory(),
                       tag=torch.from_numpy(rng.integers(0, 1000, B).astype(np.int32)).pin_memory())
            host = {k: v.numpy() for k, v in pin.items()}
            ra, rb, rc = a.add(**host), b.add(**host), c.add(**host)
            assert a._last_zero_copy and not b._last_zero_copy and not c._last_zero_copy
            np.testing.assert_array_equal(ra["status"], rb["status"])
            np.testing.assert_array_equal(ra["value"], rb["value"])
            np.testing.assert_array_equal(rc["status"], rb["status"])
            np.testing.assert_array_equal(rc["value"], rb["value"])
        da, db, dc = a.data(), b.data(), c.data()
        for k in da:
            np.testing.assert_array_equal(da[k], db[k], err_msg=k)
            np.testing.assert_array_equal(dc[k], db[k], err_msg=k)
        assert a._zero_copy_bytes > 0
        for k in a.best_elite:
            np.testing.assert_array_equal(a.best_elite[k], b.best_elite[k])


@pytest.mark.parametrize("kind", ["grid", "cvt"])
@pytest.mark.parametrize("pinned", [False, True])
def test_result_archive_shares_the_index_of_the_main_archive(kind, pinned):
    """Scheduler(archive, emitters, result_archive=...) inserts every batch into two archives of the same
    geometry (CMA-MAE: _scheduler.py:260-264). The second insertion reuses the cell indices of the first;
    the result archive must equal one that indexed the batch itself."""
    import torch

    from pyribs_b200.archives import CVTArchive, GridArchive
    from pyribs_b200.schedulers import Scheduler

    rng = np.random.default_rng(17)
    cen = rng.uniform(-1, 1, (400, 2))

    def make(lr):
        kw = dict(learning_rate=lr, threshold_min=0.0) if lr is not None else {}
        if kind == "grid":
            return GridArchive(solution_dim=6, dims=[15, 15], ranges=[(-1, 1)] * 2, **kw)
        return CVTArchive(solution_dim=6, centroids=cen, ranges=[(-1, 1)] * 2, **kw)

    main, shared, alone = make(0.05), make(None), make(None)
    assert shared.same_geometry(main) and not shared.same_geometry(GridArchive(solution_dim=6, dims=[15, 16],
                                                                               ranges=[(-1, 1)] * 2))

    class _Stub:  # Scheduler only needs something emitter-shaped to be constructed
        pass

    sched = Scheduler(main, [_Stub()], result_archive=shared, pool_emitters=False)
    for j in range(4):
        B = 3000
        b = dict(solution=rng.standard_normal((B, 6)), objective=rng.uniform(0, 10, B) + j,
                 measures=rng.uniform(-1.1, 1.1, (B, 2)))
        if pinned:
            keep = {k: torch.from_numpy(v).pin_memory() for k, v in b.items()}
            b = {k: v.numpy() for k, v in keep.items()}
        sched._add_to_archives(dict(b))
        assert shared._last_idx is not None and main._last_idx is not None
        alone.add(**b)
    ds, da = shared.data(), alone.data()
    assert len(shared) == len(alone) > 100
    for k in ds:
        np.testing.assert_array_equal(ds[k], da[k], err_msg=k)
    assert shared.stats.qd_score == alone.stats.qd_score


@pytest.mark.parametrize("kind", ["grid_me", "grid_mae", "cvt_mae", "grid_list_mode"])
def test_pipelined_commit_equals_default(kind):
    """`archive.pipeline_commit = True` runs the row copy of add() k on a side stream next to the row phases of
    add() k+1 (QD_INSERT_ROWS / QD_INSERT_COPY, two winner-list buffers). Everything observable must equal the
    default single-launch mode: status / value of every call, stats, best elite, the stored rows -- also when the
    batch tensors are dropped right after add() and when reads are interleaved."""
    import torch

    from pyribs_b200.archives import CVTArchive, GridArchive

    rng = np.random.default_rng(31)
    n = 24
    cen = rng.uniform(-1, 1, (700, 2))

    def make():
        if kind == "grid_me":
            return GridArchive(solution_dim=n, dims=[40, 30], ranges=[(-1, 1)] * 2, seed=3,
                               extra_fields={"tag": ((2,), np.int32)})
        if kind == "grid_mae":
            return GridArchive(solution_dim=n, dims=[40, 30], ranges=[(-1, 1)] * 2, learning_rate=0.05, threshold_min=-1.0,
                               seed=3)
        if kind == "grid_list_mode":  # cells >> rows: the compact winner list path
            return GridArchive(solution_dim=n, dims=[400, 300], ranges=[(-1, 1)] * 2, learning_rate=0.2, threshold_min=0.0,
                               seed=3)
        return CVTArchive(solution_dim=n, centroids=cen, ranges=[(-1, 1)] * 2, learning_rate=0.1, threshold_min=0.0, seed=3)

    a, b = make(), make()
    b.pipeline_commit = True
    assert b.pipeline_commit and not a.pipeline_commit
    B = 5000
    for j in range(9):
        batch = dict(solution=rng.standard_normal((B, n)), objective=rng.uniform(0, 10, B) + 0.7 * j,
                     measures=rng.uniform(-1.05, 1.05, (B, 2)))
        if kind == "grid_me":
            batch["tag"] = rng.integers(0, 1000, (B, 2)).astype(np.int32)
        da = {k: torch.from_numpy(v).cuda() for k, v in batch.items()}
        db = {k: torch.from_numpy(v).cuda() for k, v in batch.items()}
        ra, rb = a.add(**da), b.add(**db)
        del db  # freed while the side-stream copy may still be reading it
        junk = torch.full((B, n), float("nan"), dtype=torch.float64, device="cuda")  # would land in the freed block
        assert torch.equal(ra["status"], rb["status"]) and torch.equal(ra["value"], rb["value"])
        del junk
        if j in (3, 6):  # interleaved reads wait for the copy
            assert len(a) == len(b)
            xa, xb = a.data(), b.data()
            for k in xa:
                np.testing.assert_array_equal(xa[k], xb[k], err_msg=f"{kind} add {j} {k}")
            ea, eb = a.sample_elites(5), b.sample_elites(5)
            for k in ea:
                np.testing.assert_array_equal(ea[k], eb[k])
    assert len(a) == len(b) > 100
    sa, sb = a.stats, b.stats
    assert (sa.num_elites, sa.qd_score, sa.obj_max, sa.obj_mean) == (sb.num_elites, sb.qd_score, sb.obj_max, sb.obj_mean)
    for k in a.best_elite:
        np.testing.assert_array_equal(a.best_elite[k], b.best_elite[k], err_msg=f"best elite {k}")
    xa, xb = a.data(), b.data()
    for k in xa:
        np.testing.assert_array_equal(xa[k], xb[k], err_msg=f"{kind} final {k}")
    # back to the default mode, and a host batch after pipelined adds
    b.pipeline_commit = False
    h = dict(solution=rng.standard_normal((300, n)), objective=rng.uniform(20, 30, 300), measures=rng.uniform(-1, 1, (300, 2)))
    if kind == "grid_me":
        h["tag"] = np.zeros((300, 2), dtype=np.int32)
    ra, rb = a.add(**h), b.add(**h)
    np.testing.assert_array_equal(ra["status"], rb["status"])
    for k in xa:
        np.testing.assert_array_equal(a.data()[k], b.data()[k])
