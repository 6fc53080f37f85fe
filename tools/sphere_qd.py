"""End-to-end check against the reference's published results (docs/examples/sphere.md: mean +- std over 20 trials,
10 000 iterations, 100-dimensional sphere with linear projection, examples/sphere.py): runs the same algorithm
configurations on this framework and prints QD score / coverage next to the published numbers.

    python tools/sphere_qd.py [ALGORITHM ...] [--iters 10000] [--seed 1]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

PUBLISHED = {  # docs/examples/sphere.md
    "map_elites": (416_111.64, 2_681.37, 50.72, 0.40), "cvt_map_elites": (416_350.11, 3_942.42, 50.67, 0.49),
    "line_map_elites": (491_013.92, 1_118.74, 60.44, 0.17),
    "me_map_elites": (533_477.98, 15_041.68, 65.28, 2.05), "cma_me_imp": (456_678.01, 2_882.10, 55.84, 0.41),
    "cma_me_imp_mu": (498_464.35, 2_214.86, 61.11, 0.34), "cma_me_basic": (398_192.00, 10_174.36, 47.04, 1.30),
    "cma_me_rd": (458_403.60, 2_739.42, 56.21, 0.42), "cma_mae": (633_368.55, 1_504.43, 81.02, 0.27),
}


def sphere(sol):
    """examples/sphere.py:960-1017 (objective and measures; the gradients are not needed here)."""
    dim = sol.shape[1]
    shift = 5.12 * 0.4
    worst = (-5.12 - shift) ** 2 * dim
    raw = np.sum(np.square(sol - shift), axis=1)
    obj = (raw - worst) / (0.0 - worst) * 100
    clipped = sol.copy()
    mask = (clipped < -5.12) | (clipped > 5.12)
    clipped[mask] = 5.12 / clipped[mask]
    meas = np.stack((clipped[:, : dim // 2].sum(1), clipped[:, dim // 2:].sum(1)), axis=1)
    return obj, meas


def build(algorithm, seed, dim=100):
    """examples/sphere.py CONFIG + create_scheduler (:1020-1150)."""
    from pyribs_b200.archives import CVTArchive, GridArchive
    from pyribs_b200.emitters import EvolutionStrategyEmitter, GaussianEmitter, IsoLineEmitter
    from pyribs_b200.schedulers import BanditScheduler, Scheduler

    mb = dim / 2 * 5.12
    bounds = [(-mb, mb)] * 2
    es = lambda **kw: (EvolutionStrategyEmitter, dict(sigma0=0.5, batch_size=36, **kw), 15)  # noqa: E731
    cfg = {
        "map_elites": dict(emitters=[(GaussianEmitter, dict(sigma=0.5, batch_size=540), 1)]),
        "cvt_map_elites": dict(cvt=True, emitters=[(GaussianEmitter, dict(sigma=0.5, batch_size=540), 1)]),
        "line_map_elites": dict(emitters=[(IsoLineEmitter, dict(iso_sigma=0.5, line_sigma=0.2, batch_size=540), 1)]),
        "me_map_elites": dict(emitters=[es(ranker="obj"), es(ranker="2rd"), es(ranker="2imp"),
                                        (IsoLineEmitter, dict(iso_sigma=0.01, line_sigma=0.1, batch_size=36), 15)],
                              bandit=dict(num_active=15, reselect="terminated")),
        "cma_me_imp": dict(emitters=[es(ranker="2imp", selection_rule="filter", restart_rule="no_improvement")]),
        "cma_me_imp_mu": dict(emitters=[es(ranker="2imp", selection_rule="mu", restart_rule="no_improvement")]),
        "cma_me_basic": dict(emitters=[es(ranker="2obj", selection_rule="mu", restart_rule="basic")]),
        "cma_me_rd": dict(emitters=[es(ranker="2rd", selection_rule="filter", restart_rule="no_improvement")]),
        "cma_mae": dict(archive=dict(threshold_min=0, learning_rate=0.01), result=True,
                        emitters=[es(ranker="imp", selection_rule="mu", restart_rule="basic")]),
    }[algorithm]
    result = GridArchive(solution_dim=dim, dims=(100, 100), ranges=bounds, seed=seed) if cfg.get("result") else None
    if cfg.get("cvt"):  # centroids=10000: k-means over 100 000 uniform samples, as the reference's constructor does
        archive = CVTArchive(solution_dim=dim, centroids=10_000, ranges=bounds, seed=seed)
    else:
        archive = GridArchive(solution_dim=dim, dims=(100, 100), ranges=bounds, seed=seed, **cfg.get("archive", {}))
    ss = np.random.SeedSequence(seed)
    emitters = []
    for cls, kw, count in cfg["emitters"]:
        emitters += [cls(archive, x0=np.zeros(dim), seed=s, **kw) for s in ss.spawn(count)]
    if "bandit" in cfg:
        return BanditScheduler(archive, emitters, result, **cfg["bandit"])
    return Scheduler(archive, emitters, result)


def run(algorithm, iters, seed):
    sched = build(algorithm, seed)
    t0 = time.perf_counter()
    for _ in range(iters):
        sol = sched.ask()
        obj, meas = sphere(np.asarray(sol))
        sched.tell(obj, meas)
    dt = time.perf_counter() - t0
    st = sched.result_archive.stats
    pub = PUBLISHED[algorithm]
    out = {"algorithm": algorithm, "iterations": iters, "seed": seed, "qd_score": float(st.qd_score),
           "coverage_pct": float(st.coverage) * 100, "it_per_s": iters / dt,
           "published_qd_score": f"{pub[0]:,.2f} +- {pub[1]:,.2f}", "published_coverage_pct": f"{pub[2]:.2f} +- {pub[3]:.2f}"}
    if iters == 10_000:
        out["qd_score_sigmas_from_published_mean"] = (out["qd_score"] - pub[0]) / pub[1]
        out["coverage_sigmas_from_published_mean"] = (out["coverage_pct"] - pub[2]) / pub[3]
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("algorithms", nargs="*", default=["cma_me_imp"])
    ap.add_argument("--iters", type=int, default=10_000)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    for alg in a.algorithms:
        run(alg, a.iters, a.seed)
