"""Generates tests/golden/*.npz from the UNMODIFIED reference (test infrastructure).

Run in the authoring container only (`/root/reference` is absent on the GPU box):

    python oracle/gen_golden.py

The reference is imported from /root/reference through oracle/shim/ (array_api_compat ->
SciPy's vendored copy; numpy_groupies.aggregate_nb -> numba restatement of len/sum/
argmax). Every array written here is an output of the
reference's own code on the recorded inputs.
"""

from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("QD_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "shim"))
sys.path.insert(0, REF)

import numpy as np  # noqa: E402

import ribs  # noqa: E402
from ribs.archives import CVTArchive, GridArchive  # noqa: E402
from ribs.emitters.opt import _get_es  # noqa: E402
from ribs.emitters.rankers import _get_ranker  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def save(name, d):
    flat = {}

    def rec(prefix, v):
        if isinstance(v, dict):
            for k, x in v.items():
                rec(f"{prefix}/{k}" if prefix else str(k), x)
        else:
            flat[prefix] = np.asarray(v)

    rec("", d)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **flat)
    print(f"{name}: {len(flat)} arrays, {os.path.getsize(path) / 1e3:.1f} kB")


# ---------------------------------------------------------------------------------------
# 1. GridArchive.index_of
# ---------------------------------------------------------------------------------------
def adversarial_measures(rng, dims, ranges, dtype, n_rand):
    lo = np.array([r[0] for r in ranges], dtype=np.float64)
    hi = np.array([r[1] for r in ranges], dtype=np.float64)
    d = len(dims)
    pts = [rng.uniform(lo - 0.1 * (hi - lo), hi + 0.1 * (hi - lo), (n_rand, d))]
    # exact boundaries and +-1 ulp around them, per dimension, other dims random
    for k in range(d):
        edges = np.linspace(lo[k], hi[k], dims[k] + 1).astype(dtype)
        for e in (edges, np.nextafter(edges, dtype(np.inf)), np.nextafter(edges, dtype(-np.inf))):
            m = rng.uniform(lo, hi, (len(e), d))
            m[:, k] = e
            pts.append(m)
    big = np.array([1e300 if dtype == np.float64 else 1e38, 2.0**31, 2.0**31 * (hi[0] - lo[0]),
                    2.0**33, 1e10, 3e9])
    for v in big:
        for sgn in (1.0, -1.0):
            m = rng.uniform(lo, hi, (d, d))
            m[np.arange(d), np.arange(d)] = sgn * v
            pts.append(m)
    return np.concatenate(pts).astype(dtype)


def gen_grid_index():
    rng = np.random.default_rng(1234)
    out = {}
    configs = [
        ((10, 20), [(-1, 1), (-2, 2)]),
        ((100, 100), [(-1, 1), (-1, 1)]),
        ((500, 500), [(-1, 1), (-1, 1)]),
        ((100, 100), [(-256, 256), (-256, 256)]),
        ((7, 5, 3, 4, 6), [(-1, 1), (0, 1), (-3, 2), (10, 20), (-1e-3, 1e-3)]),
        ((64,), [(-1, 1)]),
    ]
    for ci, (dims, ranges) in enumerate(configs):
        for dt in (np.float64, np.float32):
            a = GridArchive(solution_dim=1, dims=dims, ranges=ranges, dtype=dt)
            m = adversarial_measures(rng, dims, ranges, dt, 4000)
            with np.errstate(all="ignore"):
                idx = a.index_of(m)
            out[f"c{ci}_{np.dtype(dt).name}"] = {
                "dims": np.array(dims), "ranges": np.array(ranges, dtype=np.float64),
                "measures": m, "index": idx,
            }
    # the 13 probes of tests/archives/grid_archive_test.py:668-697
    for dt in (np.float64, np.float32):
        a = GridArchive(solution_dim=0, dims=[10], ranges=[(0, 0.1)], epsilon=1e-6, dtype=dt)
        m = np.array([[-0.01], [0.0], [0.01], [0.0100001], [0.05], [0.0500001], [0.09],
                      [0.0900001], [0.1], [0.11], [0.03], [0.07], [0.0999999]], dtype=dt)
        out[f"probe_{np.dtype(dt).name}"] = {
            "dims": np.array([10]), "ranges": np.array([(0, 0.1)]), "measures": m, "index": a.index_of(m)}
    save("grid_index", out)


# ---------------------------------------------------------------------------------------
# 2. CVTArchive.index_of
# ---------------------------------------------------------------------------------------
def gen_cvt_index():
    rng = np.random.default_rng(99)
    out = {}
    for name, C, d, B in [("d2", 10_000, 2, 8_000), ("d8", 12_000, 8, 3_000), ("d3", 777, 3, 3_001),
                          ("d1", 50, 1, 500), ("d16", 2_000, 16, 1_000)]:
        cen = rng.uniform(-1, 1, (C, d))
        m = rng.uniform(-1.1, 1.1, (B, d))
        a = CVTArchive(solution_dim=1, centroids=cen, ranges=[(-1, 1)] * d)
        out[name] = {"centroids": cen, "measures": m, "index": a.index_of(m)}
        if C * B <= 4e7:
            b = CVTArchive(solution_dim=1, centroids=cen, ranges=[(-1, 1)] * d,
                           nearest_neighbors="brute_force", chunk_size=512)
            out[name]["index_brute"] = b.index_of(m)
    # float32 archive
    cen = rng.uniform(-1, 1, (3000, 2)).astype(np.float32)
    m = rng.uniform(-1, 1, (4000, 2)).astype(np.float32)
    a = CVTArchive(solution_dim=1, centroids=cen, ranges=[(-1, 1)] * 2, dtype=np.float32)
    out["d2_f32"] = {"centroids": cen, "measures": m, "index": a.index_of(m)}
    # constructed exact ties vs brute force (first-min wins): duplicated centroids and
    # exact midpoints on a dyadic lattice.
    cen = rng.integers(-8, 9, (400, 2)).astype(np.float64) / 8.0
    cen = np.concatenate([cen, cen[:50]])  # duplicates
    m = np.concatenate([(cen[:200] + cen[200:400]) / 2.0, cen[:100], rng.integers(-16, 17, (500, 2)) / 16.0])
    b = CVTArchive(solution_dim=1, centroids=cen, ranges=[(-1, 1)] * 2, nearest_neighbors="brute_force")
    out["ties"] = {"centroids": cen, "measures": m, "index_brute": b.index_of(m)}
    # 4-centroid fixture of tests/archives/conftest.py
    cen = np.array([[0.5, 0.5], [-0.5, 0.5], [-0.5, -0.5], [0.5, -0.5]])
    m = rng.uniform(-1, 1, (64, 2))
    a = CVTArchive(solution_dim=3, centroids=cen, ranges=[(-1, 1), (-1, 1)])
    out["fixture4"] = {"centroids": cen, "measures": m, "index": a.index_of(m)}
    save("cvt_index", out)


# ---------------------------------------------------------------------------------------
# 3. add trajectories
# ---------------------------------------------------------------------------------------
def dump_archive(a):
    st = a._store
    n = len(st)
    occ = np.array(st.occupied)
    d = {"occupied": occ, "occupied_list": np.array(st.occupied_list), "n": n}
    for name, arr in st._fields.items():
        x = np.array(arr)
        x[~occ] = 0  # unoccupied rows are np.empty garbage
        d["field_" + name] = x
    s = a.stats
    d["stats"] = {
        "num_elites": s.num_elites, "coverage": s.coverage, "qd_score": s.qd_score,
        "norm_qd_score": s.norm_qd_score,
        "obj_max": np.nan if s.obj_max is None else s.obj_max,
        "obj_mean": np.nan if s.obj_mean is None else s.obj_mean,
    }
    if a.best_elite is not None:
        d["best_elite"] = {k: v for k, v in a.best_elite.items()}
    return d


def run_traj(make, batches, mode="batch"):
    a = make()
    steps = {}
    for j, b in enumerate(batches):
        if mode == "batch":
            info = a.add(**b)
        else:
            infos = [a.add_single(**{k: v[i] for k, v in b.items()}) for i in range(len(b["objective"]))]
            info = {"status": np.array([x["status"] for x in infos], dtype=np.int32),
                    "value": np.array([x["value"] for x in infos])}
        steps[f"s{j}"] = {"in": b, "status": info["status"], "value": info["value"], "after": dump_archive(a)}
    return steps


def clustered(rng, E, lam, sd_c=0.05, sd=0.02):
    mu = rng.normal(0, sd_c, (E, 1, 2))
    return np.clip(mu + rng.normal(0, sd, (E, lam, 2)), -1, 1).reshape(E * lam, 2)


def gen_add():
    rng = np.random.default_rng(2024)
    out = {}

    def batch(B, n, meas, obj=None, dtype=np.float64, extras=None):
        b = {"solution": rng.standard_normal((B, n)).astype(dtype),
             "objective": (rng.uniform(0, 100, B) if obj is None else obj).astype(dtype),
             "measures": meas.astype(dtype)}
        if extras:
            b.update(extras)
        return b

    # (a) CMA-MAE on 50x50, B=4096, 3 calls with rising objectives, 5 learning rates
    for lr in (0.0, 0.001, 0.01, 0.1, 1.0):
        bs = [batch(1024, 4, clustered(rng, 4, 256, 0.15, 0.06), rng.uniform(0, 100, 1024) + 15 * j) for j in range(3)]
        mk = lambda lr=lr: GridArchive(solution_dim=4, dims=(32, 32), ranges=[(-1, 1)] * 2,  # noqa: E731
                                       learning_rate=lr, threshold_min=0.0)
        out[f"mae_lr{lr}"] = {"cfg": {"kind": 0, "dims": [32, 32], "ranges": [(-1, 1)] * 2, "lr": lr,
                                      "tmin": 0.0, "n": 4, "f32": 0}, **run_traj(mk, bs)}
    # (b) CMA-ME (defaults) with ties in objective (integers) and many collisions
    bs = [batch(1500, 4, rng.uniform(-1.2, 1.2, (1500, 2)), rng.integers(0, 20, 1500).astype(float) + 3 * j)
          for j in range(3)]
    mk = lambda: GridArchive(solution_dim=4, dims=(20, 20), ranges=[(-1, 1)] * 2)  # noqa: E731
    out["me_ties"] = {"cfg": {"kind": 0, "dims": [20, 20], "ranges": [(-1, 1)] * 2, "lr": np.nan,
                              "tmin": -np.inf, "n": 4, "f32": 0}, **run_traj(mk, bs)}
    # (c) negative threshold_min, objectives straddling it
    bs = [batch(1000, 3, rng.uniform(-1, 1, (1000, 2)), rng.normal(-10, 8, 1000)) for _ in range(3)]
    mk = lambda: GridArchive(solution_dim=3, dims=(10, 20), ranges=[(-1, 1), (-2, 2)],  # noqa: E731
                             learning_rate=0.1, threshold_min=-10.0, qd_score_offset=-30.0)
    out["mae_negmin"] = {"cfg": {"kind": 0, "dims": [10, 20], "ranges": [(-1, 1), (-2, 2)], "lr": 0.1,
                                 "tmin": -10.0, "n": 3, "f32": 0, "offset": -30.0}, **run_traj(mk, bs)}
    # (d) float32 archive
    bs = [batch(1500, 5, rng.uniform(-1, 1, (1500, 2)), rng.uniform(0, 50, 1500) + 10 * j, np.float32)
          for j in range(2)]
    mk = lambda: GridArchive(solution_dim=5, dims=(30, 30), ranges=[(-1, 1)] * 2, learning_rate=0.05,  # noqa: E731
                             threshold_min=1.0, dtype=np.float32)
    out["mae_f32"] = {"cfg": {"kind": 0, "dims": [30, 30], "ranges": [(-1, 1)] * 2, "lr": 0.05,
                              "tmin": 1.0, "n": 5, "f32": 1}, **run_traj(mk, bs)}
    bs = [batch(1500, 5, rng.uniform(-1, 1, (1500, 2)), rng.integers(0, 9, 1500).astype(float) + 2 * j, np.float32)
          for j in range(2)]
    mk = lambda: GridArchive(solution_dim=5, dims=(30, 30), ranges=[(-1, 1)] * 2, dtype=np.float32)  # noqa: E731
    out["me_f32"] = {"cfg": {"kind": 0, "dims": [30, 30], "ranges": [(-1, 1)] * 2, "lr": np.nan,
                             "tmin": -np.inf, "n": 5, "f32": 1}, **run_traj(mk, bs)}
    # (e) extra numeric fields of mixed shapes/dtypes
    ex = lambda B: {"meta": rng.integers(0, 1000, (B, 3)).astype(np.int32),  # noqa: E731
                    "score": rng.standard_normal(B).astype(np.float32)}
    bs = [batch(1500, 4, rng.uniform(-1, 1, (1500, 2)), extras=ex(1500)) for _ in range(2)]
    mk = lambda: GridArchive(solution_dim=4, dims=(16, 16), ranges=[(-1, 1)] * 2,  # noqa: E731
                             extra_fields={"meta": ((3,), np.int32), "score": ((), np.float32)})
    out["me_extras"] = {"cfg": {"kind": 0, "dims": [16, 16], "ranges": [(-1, 1)] * 2, "lr": np.nan,
                                "tmin": -np.inf, "n": 4, "f32": 0, "extras": 1}, **run_traj(mk, bs)}
    # (f) CVT archives (CMA-ME and CMA-MAE)
    cen = rng.uniform(-1, 1, (500, 2))
    bs = [batch(1500, 5, rng.uniform(-1, 1, (1500, 2)), rng.standard_normal(1500) + j) for j in range(3)]
    mk = lambda: CVTArchive(solution_dim=5, centroids=cen, ranges=[(-1, 1)] * 2)  # noqa: E731
    out["cvt_me"] = {"cfg": {"kind": 1, "centroids": cen, "lr": np.nan, "tmin": -np.inf, "n": 5, "f32": 0},
                     **run_traj(mk, bs)}
    cen8 = rng.uniform(-1, 1, (600, 8))
    bs = [batch(1500, 5, rng.uniform(-1, 1, (1500, 8)), rng.uniform(0, 10, 1500) + 2 * j) for j in range(3)]
    mk = lambda: CVTArchive(solution_dim=5, centroids=cen8, ranges=[(-1, 1)] * 8,  # noqa: E731
                            learning_rate=0.2, threshold_min=0.5)
    out["cvt_mae"] = {"cfg": {"kind": 1, "centroids": cen8, "lr": 0.2, "tmin": 0.5, "n": 5, "f32": 0},
                      **run_traj(mk, bs)}
    # (g) add_single trajectories (scalar rule)
    bs = [batch(300, 3, rng.uniform(-1, 1, (300, 2)), rng.uniform(0, 10, 300))]
    mk = lambda: GridArchive(solution_dim=3, dims=(6, 6), ranges=[(-1, 1)] * 2, learning_rate=0.3,  # noqa: E731
                             threshold_min=2.0)
    out["single_mae"] = {"cfg": {"kind": 0, "dims": [6, 6], "ranges": [(-1, 1)] * 2, "lr": 0.3, "tmin": 2.0,
                                 "n": 3, "f32": 0, "single": 1}, **run_traj(mk, bs, mode="single")}
    mk = lambda: GridArchive(solution_dim=3, dims=(6, 6), ranges=[(-1, 1)] * 2)  # noqa: E731
    out["single_me"] = {"cfg": {"kind": 0, "dims": [6, 6], "ranges": [(-1, 1)] * 2, "lr": np.nan,
                                "tmin": -np.inf, "n": 3, "f32": 0, "single": 1}, **run_traj(mk, bs, mode="single")}
    # (h) KATs of tests/archives/grid_archive_test.py:327-596 on the (10,20) fixture
    kat = lambda sol, obj, meas: {"solution": np.array(sol, float), "objective": np.array(obj, float),  # noqa: E731
                                  "measures": np.array(meas, float)}
    mk = lambda: GridArchive(solution_dim=3, dims=[10, 20], ranges=[(-1, 1), (-2, 2)])  # noqa: E731
    bs = [kat([[1, 2, 3]], [1.0], [[0.25, 0.25]]),
          kat([[1, 2, 3]] * 6, [0.0, 0.0, 2.0, 2.0, 0.5, 3.0], [[0, 0], [0.25, 0.25], [0.25, 0.25], [0, 0],
                                                                  [0.25, 0.25], [0.25, 0.25]]),
          kat([[1, 2, 3], [4, 5, 6], [7, 8, 9]], [5.0, 5.0, 5.0], [[0.7, 0.7]] * 3),
          kat(np.zeros((0, 3)), np.zeros(0), np.zeros((0, 2)))]
    out["kat_me"] = {"cfg": {"kind": 0, "dims": [10, 20], "ranges": [(-1, 1), (-2, 2)], "lr": np.nan,
                             "tmin": -np.inf, "n": 3, "f32": 0}, **run_traj(mk, bs)}
    mk = lambda: GridArchive(solution_dim=3, dims=[10, 20], ranges=[(-1, 1), (-2, 2)],  # noqa: E731
                             learning_rate=0.1, threshold_min=-10.0)
    bs = [kat([[1, 2, 3]] * 4, [-20.0, -10.0, 0.0, 10.0], [[0, 0]] * 4),
          kat([[1, 2, 3]] * 4, [-20.0, 5.0, 5.0, 1.0], [[0, 0], [0, 0], [0.5, 0.5], [0, 0]])]
    out["kat_mae"] = {"cfg": {"kind": 0, "dims": [10, 20], "ranges": [(-1, 1), (-2, 2)], "lr": 0.1,
                              "tmin": -10.0, "n": 3, "f32": 0}, **run_traj(mk, bs)}
    save("add_traj", out)


# ---------------------------------------------------------------------------------------
# 4. ES traces with injected noise
# ---------------------------------------------------------------------------------------
class InjectedRNG:
    """Stands in for np.random.Generator inside the reference ES: unit normals come from
    a recorded array (Generator.normal(loc, scale) == loc + scale * standard_normal)."""

    def __init__(self):
        self.z = None

    def normal(self, loc, scale, size):
        assert self.z is not None and tuple(size) == self.z.shape
        z, self.z = self.z, None
        return loc + scale * z

    def standard_normal(self, size, dtype=None):
        assert self.z is not None and tuple(size) == self.z.shape
        z, self.z = self.z, None
        return z


def es_state(es, name):
    d = {"mean": es.mean, "sigma": es.sigma}
    if name == "cma_es":
        d.update(pc=es.pc, ps=es.ps, C=es.cov.cov, eigvals=es.cov.eigenvalues, invsqrt=es.cov.invsqrt,
                 B=es.cov.eigenbasis, cond=es.cov.condition_number, current_eval=es.current_eval,
                 updated_eval=es.cov.updated_eval)
    elif name == "sep_cma_es":
        d.update(pc=es.pc, ps=es.ps, C=es.cov.cov, current_eval=es.current_eval)
    else:
        d.update(ps=es.ps, m=es.m, gens=es.current_gens)
    return {k: np.array(v) for k, v in d.items()}


def gen_es():
    rng = np.random.default_rng(77)
    out = {}
    for name in ("cma_es", "sep_cma_es", "lm_ma_es"):
        for (n, lam) in [(10, 8), (100, 36), (256, 64)]:
            for rule in ("mu", "filter"):
                if name == "lm_ma_es" and lam > n:
                    continue
                if n == 256 and rule == "filter":
                    continue
                kw = {"n_vectors": 5} if (name == "lm_ma_es" and n == 100) else {}
                es = _get_es(name, sigma0=0.5, solution_dim=n, batch_size=lam, seed=1, dtype=np.float64,
                             lower_bounds=np.full(n, -np.inf), upper_bounds=np.full(n, np.inf), **kw)
                es._rng = InjectedRNG()
                x0 = rng.standard_normal(n)
                es.reset(x0)
                G = {10: 20, 100: 12, 256: 4}[n]
                case = {"cfg": {"n": n, "lam": lam, "sigma0": 0.5, "x0": x0, "filter": int(rule == "filter"),
                                "n_vectors": kw.get("n_vectors", -1)}}
                for g in range(G):
                    zseed = int(rng.integers(0, 2**31))
                    z = np.random.default_rng(zseed).standard_normal((lam, n))
                    es._rng.z = z
                    sol = np.array(es.ask())
                    fit = -np.sum((sol - 1.0) ** 2, axis=1)  # shifted sphere
                    ranking = np.flip(np.argsort(fit))
                    mu = lam // 2 if rule == "mu" else int(rng.integers(0, lam + 1))
                    if g == 3:
                        mu = 0  # exercises the early return
                    es.tell(ranking, fit, mu)
                    stop = es.check_stop(fit[ranking])
                    st = es_state(es, name)
                    heavy = (n <= 10) or (g in (0, G // 2, G - 1) and n <= 100) or g == G - 1
                    if not heavy:
                        for k in ("C", "invsqrt", "B", "m"):
                            if k in st and st[k].ndim == 2:
                                del st[k]
                    case[f"g{g}"] = {"zseed": zseed, "ranking": ranking, "mu": mu, "stop": int(stop), **st}
                    if heavy:
                        case[f"g{g}"]["solutions"] = sol
                        if n > 10:
                            case[f"g{g}"].pop("B", None)
                out[f"{name}_n{n}_l{lam}_{rule}"] = case
    save("es_traces", out)


# ---------------------------------------------------------------------------------------
# 5. rankers
# ---------------------------------------------------------------------------------------
def gen_rankers():
    rng = np.random.default_rng(5)
    out = {}
    a = GridArchive(solution_dim=2, dims=(10, 10), ranges=[(-1, 1), (-2, 2)])
    for tag, B, ties in [("a", 36, False), ("b", 512, False), ("c", 40, True), ("d", 9, True)]:
        status = rng.integers(0, 3, B).astype(np.int32)
        value = rng.integers(-3, 4, B).astype(float) if ties else rng.standard_normal(B)
        objective = rng.integers(0, 5, B).astype(float) if ties else rng.standard_normal(B)
        measures = rng.uniform(-1, 1, (B, 2))
        data = {"solution": np.zeros((B, 2)), "objective": objective, "measures": measures}
        info = {"status": status, "value": value}
        case = {"status": status, "value": value, "objective": objective, "measures": measures}
        for r in ("imp", "2imp", "obj", "2obj", "rd", "2rd"):
            rk = _get_ranker(r, 11)
            rk.reset(None, a)
            idx, vals = rk.rank(None, a, data, info)
            case[r] = {"indices": idx, "values": vals}
            if r in ("rd", "2rd"):
                case[r]["dir"] = rk.target_measure_dir
        out[tag] = case
    save("rankers", out)


# ---------------------------------------------------------------------------------------
# 6. GaussianEmitter / IsoLineEmitter asks (MAP-Elites variation operators)
# ---------------------------------------------------------------------------------------
def gen_emitters():
    from ribs.emitters import GaussianEmitter, IsoLineEmitter

    out = {}
    for tag, dtype, bounded in (("f64", np.float64, False), ("f64_bounds", np.float64, True), ("f32", np.float32, False)):
        rng = np.random.default_rng(21)
        n, B = 7, 48
        arch = GridArchive(solution_dim=n, dims=(12, 12), ranges=[(-1, 1)] * 2, seed=77, dtype=dtype)
        fill = dict(solution=rng.standard_normal((600, n)).astype(dtype), objective=rng.standard_normal(600).astype(dtype),
                    measures=rng.uniform(-1, 1, (600, 2)).astype(dtype))
        lo = np.full(n, -0.8, dtype=dtype) if bounded else None
        hi = np.full(n, 0.9, dtype=dtype) if bounded else None
        sig = np.linspace(0.05, 0.3, n).astype(dtype)
        ge = GaussianEmitter(arch, sigma=sig, x0=np.full(n, 0.25), lower_bounds=lo, upper_bounds=hi, batch_size=B, seed=5)
        ie = IsoLineEmitter(arch, iso_sigma=0.02, line_sigma=0.3, x0=np.full(n, -0.1), lower_bounds=lo, upper_bounds=hi,
                            batch_size=B, seed=6)
        g_rng, i_rng = np.random.default_rng(5), np.random.default_rng(6)  # replay the emitters' normal streams
        case = {"cfg": {"n": n, "B": B, "f32": int(dtype == np.float32), "bounded": int(bounded), "sigma": sig,
                        "iso_sigma": 0.02, "line_sigma": 0.3, "gx0": np.full(n, 0.25), "ix0": np.full(n, -0.1)},
                "fill": fill}
        for step in range(3):  # step 0: empty archive (parents = x0); then two asks on the filled archive
            if step == 1:
                arch.add(**fill)
            zg = g_rng.standard_normal((B, n))
            xg = ge.ask()
            zi, zl = i_rng.standard_normal((B, n)), i_rng.standard_normal((B, 1))
            xi = ie.ask()
            case[f"s{step}"] = {"gauss_z": zg, "gauss_x": xg, "iso_z": zi, "line_z": zl, "iso_x": xi}
        out[tag] = case
    save("emitters", out)


# ---------------------------------------------------------------------------------------
# 6b. OpenAI-ES traces with injected noise (mirror sampling on / off, Adam options)
# ---------------------------------------------------------------------------------------
def gen_openai_es():
    rng = np.random.default_rng(91)
    out = {}
    for tag, n, lam, mirror, adam in (("mirror_n10", 10, 8, True, {}), ("mirror_n300", 300, 64, True, {"lr": 0.05}),
                                      ("plain_n40", 40, 15, False, {"lr": 0.02, "l2_coeff": 0.01, "beta1": 0.8})):
        es = _get_es("openai_es", sigma0=0.3, solution_dim=n, batch_size=lam, seed=1, dtype=np.float64,
                     lower_bounds=-np.inf, upper_bounds=np.inf, mirror_sampling=mirror, **adam)
        es._rng = InjectedRNG()
        x0 = rng.standard_normal(n)
        es.reset(x0)
        case = {"cfg": {"n": n, "lam": lam, "sigma0": 0.3, "x0": x0, "mirror": int(mirror), "lr": adam.get("lr", 0.001),
                        "l2": adam.get("l2_coeff", 0.0), "beta1": adam.get("beta1", 0.9)}}
        for g in range(12):
            zseed = int(rng.integers(0, 2**31))
            z = np.random.default_rng(zseed).standard_normal((lam // 2 if mirror else lam, n))
            es._rng.z = z
            sol = np.array(es.ask())
            fit = -np.sum((sol - 1.0) ** 2, axis=1)
            ranking = np.flip(np.argsort(fit))
            es.tell(ranking, fit, lam // 2)
            case[f"g{g}"] = {"zseed": zseed, "ranking": ranking, "solutions": sol, "theta": np.array(es.adam_opt.theta),
                             "m": np.array(es.adam_opt._m), "v": np.array(es.adam_opt._v),
                             "ratio": float(es.last_update_ratio), "stop": int(es.check_stop(fit[ranking]))}
        out[tag] = case
    save("openai_es", out)


# ---------------------------------------------------------------------------------------
# 6c. CategoricalArchive trajectories (string categories, numeric solutions)
# ---------------------------------------------------------------------------------------
def gen_categorical():
    from ribs.archives import CategoricalArchive

    cats = [["A", "B", "C"], ["One", "Two", "Three", "Four"]]
    out = {}
    for tag, kw in (("me", {}), ("mae", {"learning_rate": 0.2, "threshold_min": -1.0})):
        rng = np.random.default_rng(5)
        a = CategoricalArchive(solution_dim=3, categories=cats, seed=9, **kw)
        case = {}
        for j in range(5):
            B = 40
            m = np.stack([rng.choice(cats[0], B), rng.choice(cats[1], B)], axis=1)
            b = {"solution": rng.standard_normal((B, 3)), "objective": rng.uniform(-2, 4, B) + 0.5 * j}
            info = a.add(b["solution"], b["objective"], m)
            d = a.data()
            order = np.argsort(d["index"])
            st = a.stats
            case[f"s{j}"] = {"solution": b["solution"], "objective": b["objective"], "measures": m.astype("U5"),
                             "status": info["status"], "value": info["value"],
                             "index": d["index"][order], "d_solution": d["solution"][order],
                             "d_objective": d["objective"][order], "d_threshold": d["threshold"][order],
                             "d_measures": d["measures"][order].astype("U5"),
                             "qd_score": float(st.qd_score), "obj_max": float(st.obj_max), "n": len(a),
                             "best_solution": a.best_elite["solution"], "best_measures": a.best_elite["measures"].astype("U5"),
                             "index_of": a.index_of(m)}
        single = a.add_single(np.ones(3), 100.0, ["B", "Four"])
        case["single"] = {"status": np.asarray(single["status"]), "value": np.asarray(single["value"]),
                          "best_measures": a.best_elite["measures"].astype("U5")}
        out[tag] = case
    save("categorical", out)


# ---------------------------------------------------------------------------------------
# 7. BanditScheduler (UCB1 emitter selection) driven by deterministic stub emitters
# ---------------------------------------------------------------------------------------
def gen_bandit():
    import warnings

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bandit_stubs as S
    from ribs.schedulers import BanditScheduler

    out = {}
    for case, reselect, k in S.CASES:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            masks, success, selection = S.run_bandit(BanditScheduler, case, reselect, k)
        out[f"{case}_{reselect}_{k}"] = {"active": masks, "success": success, "selection": selection}
    save("bandit", out)


# ---------------------------------------------------------------------------------------
# 10. BASELINE.json configs[4] shape: 1M centroids, 8-D -- cell indices of 20k seeded queries
#     (inputs are regenerated from the seeds by the test; only the indices are stored)
# ---------------------------------------------------------------------------------------
C5_SEEDS = {"centroids": 0, "queries": 1, "C": 1_000_000, "d": 8, "Q": 20_000}


def gen_c5_index():
    cen = np.random.default_rng(C5_SEEDS["centroids"]).uniform(-1, 1, (C5_SEEDS["C"], C5_SEEDS["d"]))
    q = np.random.default_rng(C5_SEEDS["queries"]).uniform(-1, 1, (C5_SEEDS["Q"], C5_SEEDS["d"]))
    a = CVTArchive(solution_dim=1, centroids=cen, ranges=[(-1, 1)] * C5_SEEDS["d"],
                   kdtree_query_kwargs={"workers": -1})
    idx = a.index_of(q)
    save("c5_index", {"index": idx.astype(np.int32), "checksum_centroids": np.float64(cen.sum()),
                      "checksum_queries": np.float64(q.sum()), **{k: np.int64(v) for k, v in C5_SEEDS.items()}})


# ---------------------------------------------------------------------------------------
# 11. GridArchive.retessellate (_grid_archive.py:926-976): add -> retessellate -> reads
# ---------------------------------------------------------------------------------------
def gen_retessellate():
    rng = np.random.default_rng(77)
    out = {}
    for name, n_adds in (("one_add", 1), ("three_adds", 3)):
        a = GridArchive(solution_dim=3, dims=[10, 20], ranges=[(-1, 1), (-2, 2)])
        ins = []
        for j in range(n_adds):
            b = {"solution": rng.standard_normal((400, 3)), "objective": rng.uniform(0, 10, 400) - 2.0 * j,
                 "measures": rng.uniform(-1.1, 1.1, (400, 2)) * np.array([1.0, 2.0])}
            a.add(**b)
            ins.append(b)
        before = dump_archive(a)
        a.retessellate([5, 8])
        out[name] = {"in": {f"b{j}": b for j, b in enumerate(ins)}, "before": before, "after": dump_archive(a),
                     "data_index": a.data("index"), "data_objective": a.data("objective")}
        b = {"solution": rng.standard_normal((100, 3)), "objective": rng.uniform(0, 12, 100),
             "measures": rng.uniform(-1, 1, (100, 2))}
        info = a.add(**b)
        out[name]["again"] = {"in": b, "status": info["status"], "value": info["value"], "after": dump_archive(a)}
    save("retessellate", out)


ALL = {"grid_index": gen_grid_index, "cvt_index": gen_cvt_index, "add": gen_add, "es": gen_es, "rankers": gen_rankers,
       "emitters": gen_emitters, "openai_es": gen_openai_es, "categorical": gen_categorical, "bandit": gen_bandit,
       "c5_index": gen_c5_index, "retessellate": gen_retessellate}

if __name__ == "__main__":
    print("reference", ribs.__version__, "from", os.path.dirname(ribs.__file__))
    for name in (sys.argv[1:] or list(ALL)):  # `python oracle/gen_golden.py c5_index retessellate` regenerates two files
        ALL[name]()
