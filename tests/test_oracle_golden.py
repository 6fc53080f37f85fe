"""Pins oracle/qd_oracle.py to the reference: golden vectors (always) + live reference
(when /root/reference is importable). CPU only."""
import os
import sys

import numpy as np
import pytest

from golden_util import assert_close, check_archive_dump, gens_of, load_golden, steps_of

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import qd_oracle as O  # noqa: E402

EXACT = 1e-12  # oracle vs reference: same arithmetic, same order


def oracle_dump(a):
    st = a.store
    occ = st.occupied.copy()
    d = {"occupied": occ, "occupied_list": st.occupied_list.copy(), "n": st.n_occupied, "stats": a.stats()}
    for name, arr in st.fields.items():
        x = arr.copy()
        x[~occ] = 0
        d["field_" + name] = x
    if a.best_elite is not None:
        d["best_elite"] = a.best_elite
    return d


def make_oracle_archive(cfg):
    dtype = np.float32 if int(cfg["f32"]) else np.float64
    lr = None if np.isnan(cfg["lr"]) else float(cfg["lr"])
    kw = dict(learning_rate=lr, threshold_min=float(cfg["tmin"]), dtype=dtype,
              qd_score_offset=float(cfg.get("offset", 0.0)))
    if int(cfg.get("extras", 0)):
        kw["extra_fields"] = {"meta": ((3,), np.int32), "score": ((), np.float32)}
    if int(cfg["kind"]) == 0:
        return O.make_grid_oracle(solution_dim=int(cfg["n"]), dims=cfg["dims"], ranges=cfg["ranges"], **kw)
    return O.make_cvt_oracle(solution_dim=int(cfg["n"]), centroids=cfg["centroids"], **kw)


def test_grid_index_golden():
    g = load_golden("grid_index")
    for name, c in g.items():
        dt = c["measures"].dtype
        lo = c["ranges"][:, 0].astype(dt)
        hi = c["ranges"][:, 1].astype(dt)
        idx = O.grid_index_of(c["measures"], c["dims"], lo, hi - lo, np.asarray(1e-6, dtype=dt))
        np.testing.assert_array_equal(idx, c["index"], err_msg=name)


def test_cvt_index_golden():
    g = load_golden("cvt_index")
    for name, c in g.items():
        if "index" in c:
            idx = O.CentroidIndex(c["centroids"]).index_of(c["measures"])
            np.testing.assert_array_equal(idx, c["index"], err_msg=name)
        if "index_brute" in c:
            idx = O.CentroidIndex(c["centroids"], method="brute_force").index_of(c["measures"])
            np.testing.assert_array_equal(idx, c["index_brute"], err_msg=name + " brute")


@pytest.mark.parametrize("name", sorted(load_golden("add_traj").keys()))
def test_add_golden(name):
    case = load_golden("add_traj")[name]
    a = make_oracle_archive(case["cfg"])
    single = int(case["cfg"].get("single", 0))
    for j, s in enumerate(steps_of(case)):
        b = s["in"]
        if single:
            infos = [a.add_single(**{k: v[i] for k, v in b.items()}) for i in range(len(b["objective"]))]
            status = np.array([x["status"] for x in infos], dtype=np.int32)
            value = np.array([x["value"] for x in infos])
        else:
            info = a.add(**b)
            status, value = info["status"], info["value"]
        np.testing.assert_array_equal(status, s["status"], err_msg=f"{name} s{j} status")
        assert_close(value, s["value"], EXACT, f"{name} s{j} value")
        check_archive_dump(oracle_dump(a), s["after"], EXACT, f"{name} s{j}")


@pytest.mark.parametrize("name", sorted(load_golden("es_traces").keys()))
def test_es_golden(name):
    case = load_golden("es_traces")[name]
    cfg = case["cfg"]
    es_name = name.split("_n")[0]
    kw = {}
    if es_name == "lm_ma_es" and int(cfg["n_vectors"]) > 0:
        kw["n_vectors"] = int(cfg["n_vectors"])
    es = O.ORACLE_ES[es_name](float(cfg["sigma0"]), int(cfg["n"]), int(cfg["lam"]), seed=1, **kw)
    es.reset(cfg["x0"])
    for gi, g in enumerate(gens_of(case)):
        z = np.random.default_rng(int(g["zseed"])).standard_normal((int(cfg["lam"]), int(cfg["n"])))
        sol = es.ask(z)
        if "solutions" in g:
            assert_close(sol, g["solutions"], 1e-10, f"{name} g{gi} solutions")
        es.tell(g["ranking"], int(g["mu"]))
        fit = -np.sum((sol - 1.0) ** 2, axis=1)
        assert int(es.check_stop(fit[g["ranking"]])) == int(g["stop"])
        for key, attr in [("mean", "mean"), ("sigma", "sigma"), ("ps", "ps"), ("pc", "pc"), ("C", "C"),
                          ("invsqrt", "invsqrt"), ("eigvals", "eigvals"), ("m", "m")]:
            if key in g and hasattr(es, attr):
                assert_close(getattr(es, attr), g[key], 1e-9, f"{name} g{gi} {key}")


@pytest.mark.parametrize("name", sorted(load_golden("openai_es").keys()))
def test_openai_es_golden(name):
    """OpenAI-ES + Adam (ribs/emitters/opt/_openai_es.py, _adam_opt.py) with the reference's noise injected."""
    case = load_golden("openai_es")[name]
    cfg = case["cfg"]
    n, lam, mirror = int(cfg["n"]), int(cfg["lam"]), bool(int(cfg["mirror"]))
    es = O.OracleOpenAIES(float(cfg["sigma0"]), n, lam, seed=1, mirror_sampling=mirror, lr=float(cfg["lr"]),
                          l2_coeff=float(cfg["l2"]), beta1=float(cfg["beta1"]))
    es.reset(cfg["x0"])
    for gi, g in enumerate(gens_of(case)):
        z = np.random.default_rng(int(g["zseed"])).standard_normal((lam // 2 if mirror else lam, n))
        sol = es.ask(z)
        np.testing.assert_array_equal(sol, g["solutions"], err_msg=f"{name} g{gi} solutions")
        es.tell(g["ranking"])
        assert_close(es.theta, g["theta"], EXACT, f"{name} g{gi} theta")
        assert_close(es.m, g["m"], EXACT, f"{name} g{gi} m")
        assert_close(es.v, g["v"], EXACT, f"{name} g{gi} v")
        assert_close(es.last_update_ratio, g["ratio"], 1e-10, f"{name} g{gi} ratio")
        fit = -np.sum((sol - 1.0) ** 2, axis=1)
        assert int(es.check_stop(fit[g["ranking"]])) == int(g["stop"])


def test_rankers_golden():
    g = load_golden("rankers")
    for tag, c in g.items():
        idx, _ = O.rank_one_key(c["value"])
        np.testing.assert_array_equal(idx, c["imp"]["indices"])
        idx, _ = O.rank_one_key(c["objective"])
        np.testing.assert_array_equal(idx, c["obj"]["indices"])
        idx, rv = O.rank_two_stage(c["status"], c["value"])
        np.testing.assert_array_equal(idx, c["2imp"]["indices"])
        np.testing.assert_array_equal(rv, c["2imp"]["values"])
        idx, _ = O.rank_two_stage(c["status"], c["objective"])
        np.testing.assert_array_equal(idx, c["2obj"]["indices"])
        proj = c["measures"] @ c["rd"]["dir"]
        np.testing.assert_array_equal(O.rank_one_key(proj)[0], c["rd"]["indices"])
        proj = c["measures"] @ c["2rd"]["dir"]
        np.testing.assert_array_equal(O.rank_two_stage(c["status"], proj)[0], c["2rd"]["indices"])


@pytest.mark.parametrize("tag", ["f64", "f64_bounds", "f32"])
def test_variation_emitters_golden(tag):
    """GaussianEmitter / IsoLineEmitter asks (ribs/emitters/_gaussian_emitter.py:139-166,
    _iso_line_emitter.py:156-201): the oracle replays the reference's parent draws (archive seed 77) and
    normals; the asked solutions are bit-identical."""
    case = load_golden("emitters")[tag]
    cfg = case["cfg"]
    dtype = np.float32 if int(cfg["f32"]) else np.float64
    n, B = int(cfg["n"]), int(cfg["B"])
    a = O.make_grid_oracle(solution_dim=n, dims=(12, 12), ranges=[(-1, 1)] * 2, dtype=dtype, seed=77)
    lo, hi = ((np.full(n, -0.8, dtype=dtype), np.full(n, 0.9, dtype=dtype)) if int(cfg["bounded"])
              else (-np.inf, np.inf))
    ge = O.OracleGaussianEmitter(a, cfg["sigma"], cfg["gx0"], B, lo, hi, dtype=dtype)
    ie = O.OracleIsoLineEmitter(a, float(cfg["iso_sigma"]), float(cfg["line_sigma"]), cfg["ix0"], B, lo, hi,
                                dtype=dtype)
    for j, s in enumerate(steps_of(case)):
        if j == 1:
            a.add(**case["fill"])
        xg = ge.ask(s["gauss_z"])
        xi = ie.ask(s["iso_z"], s["line_z"])
        assert xg.dtype == dtype and xi.dtype == dtype
        np.testing.assert_array_equal(xg, s["gauss_x"], err_msg=f"{tag} s{j} gaussian")
        np.testing.assert_array_equal(xi, s["iso_x"], err_msg=f"{tag} s{j} isoline")


# ---- live reference (authoring container only) ----------------------------------------
def _import_reference():
    ref = os.environ.get("QD_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "ribs")):
        pytest.skip("reference tree not present")
    for p in (os.path.join(ROOT, "oracle", "shim"), ref):
        if p not in sys.path:
            sys.path.insert(0, p)
    import ribs  # noqa: F401
    from ribs.archives import CVTArchive, GridArchive

    return GridArchive, CVTArchive


def test_live_reference_add_random():
    GridArchive, CVTArchive = _import_reference()
    rng = np.random.default_rng(0)
    for trial in range(6):
        lr = [None, 0.01, 0.5][trial % 3]
        tmin = -np.inf if lr is None else float(rng.uniform(-5, 5))
        if trial < 3:
            ref = GridArchive(solution_dim=7, dims=(25, 40), ranges=[(-1, 1), (-2, 2)], learning_rate=lr,
                              threshold_min=tmin, seed=3)
            ora = O.make_grid_oracle(solution_dim=7, dims=(25, 40), ranges=[(-1, 1), (-2, 2)], learning_rate=lr,
                                     threshold_min=tmin, seed=3)
            md = 2
        else:
            cen = rng.uniform(-1, 1, (300, 3))
            ref = CVTArchive(solution_dim=7, centroids=cen, ranges=[(-1, 1)] * 3, learning_rate=lr,
                             threshold_min=tmin, seed=3)
            ora = O.make_cvt_oracle(solution_dim=7, centroids=cen, learning_rate=lr, threshold_min=tmin, seed=3)
            md = 3
        for step in range(4):
            B = 2000
            b = dict(solution=rng.standard_normal((B, 7)), objective=rng.normal(step, 3, B),
                     measures=rng.uniform(-1.3, 1.3, (B, md)) * ([1, 2] if md == 2 else [1, 1, 1]))
            r = ref.add(**b)
            o = ora.add(**b)
            np.testing.assert_array_equal(r["status"], o["status"])
            np.testing.assert_array_equal(r["value"], o["value"])
            st = ref._store
            n = len(st)
            assert n == len(ora)
            np.testing.assert_array_equal(st.occupied_list, ora.store.occupied_list[:n])
            occ = np.array(st.occupied)
            for name, arr in st._fields.items():
                np.testing.assert_array_equal(np.array(arr)[occ], ora.store.fields[name][occ])
            assert ref.stats.qd_score == ora.stats()["qd_score"]
            assert ref.stats.obj_max == ora.stats()["obj_max"]
            e1 = ref.sample_elites(5)
            e2 = ora.sample_elites(5)
            np.testing.assert_array_equal(e1["solution"], e2["solution"])


def test_live_reference_openai_es_and_emitters():
    """Fresh configurations (not the committed goldens): the oracle's OpenAI-ES + Adam and its variation emitters
    against the reference imported here, driven by the same generators."""
    _import_reference()
    from ribs.archives import GridArchive
    from ribs.emitters import GaussianEmitter, IsoLineEmitter
    from ribs.emitters.opt import _get_es

    rng = np.random.default_rng(12)
    for mirror, lam, adam in ((True, 10, {"lr": 0.03, "beta2": 0.99}), (False, 7, {"l2_coeff": 0.05, "epsilon": 1e-6})):
        n = 23
        ref = _get_es("openai_es", sigma0=0.4, solution_dim=n, batch_size=lam, seed=3, dtype=np.float64,
                      lower_bounds=-np.inf, upper_bounds=np.inf, mirror_sampling=mirror, **adam)
        ora = O.OracleOpenAIES(0.4, n, lam, seed=3, mirror_sampling=mirror, **adam)
        x0 = rng.standard_normal(n)
        ref.reset(x0), ora.reset(x0)
        for g in range(8):  # same seed, same generator calls: the oracle draws the very normals the reference draws
            sr, so = np.array(ref.ask()), ora.ask()
            np.testing.assert_array_equal(sr, so)
            fit = -np.sum((sr + 0.3) ** 2, axis=1)
            ranking = np.flip(np.argsort(fit))
            ref.tell(ranking, fit, lam // 2), ora.tell(ranking)
            np.testing.assert_array_equal(ref.adam_opt.theta, ora.theta)
            assert ref.last_update_ratio == ora.last_update_ratio
            assert ref.check_stop(fit[ranking]) == ora.check_stop(fit[ranking])
    # variation emitters: same archive seed -> same parents, same emitter seed -> same noise
    n, B = 5, 33
    fill = dict(solution=rng.standard_normal((400, n)), objective=rng.standard_normal(400),
                measures=rng.uniform(-1, 1, (400, 2)))
    ra = GridArchive(solution_dim=n, dims=(9, 9), ranges=[(-1, 1)] * 2, seed=4)
    oa = O.make_grid_oracle(solution_dim=n, dims=(9, 9), ranges=[(-1, 1)] * 2, seed=4)
    lo, hi = np.full(n, -0.7), np.full(n, 1.1)
    rg = GaussianEmitter(ra, sigma=0.2, x0=np.zeros(n), lower_bounds=lo, upper_bounds=hi, batch_size=B, seed=8)
    og = O.OracleGaussianEmitter(oa, 0.2, np.zeros(n), B, lo, hi, seed=8)
    ri = IsoLineEmitter(ra, iso_sigma=0.05, line_sigma=0.3, x0=np.zeros(n), lower_bounds=lo, upper_bounds=hi,
                        batch_size=B, seed=9)
    oi = O.OracleIsoLineEmitter(oa, 0.05, 0.3, np.zeros(n), B, lo, hi, seed=9)
    for step in range(3):
        if step == 1:
            ra.add(**fill), oa.add(**fill)
        np.testing.assert_array_equal(rg.ask(), og.ask())
        np.testing.assert_array_equal(ri.ask(), oi.ask())
