"""BASELINE.json configs[0] / configs[2]: sphere CMA-ME (GridArchive 100x100, 15 emitters x 36) or, with
`c3`, CMA-MAE lr=0.01 (GridArchive 500x500, 100 emitters x 512, ranker imp, selection mu, restart basic). Times ask + evaluate + tell
per iteration for this repo (GPU) and for the CPU oracle port of the same loop."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))

def sphere(sol):
    n = sol.shape[1]; best = n * 5.12 ** 2; worst = 0  # examples/sphere.py:960-1017 (shifted sphere, linear projection)
    shift = 5.12 * 0.4
    obj = (best - np.sum((sol - shift) ** 2, axis=1)) / best * 100  # normalised to ~[.., 100]
    c = np.clip(sol, -5.12, 5.12)
    meas = np.stack((c[:, : n // 2].sum(1), c[:, n // 2:].sum(1)), axis=1)
    return obj, meas

def sphere_torch(sol):
    import torch
    n = sol.shape[1]
    worst = (5.12 + 2.048) ** 2 * n
    obj = (worst - ((sol - 2.048) ** 2).sum(1)) / worst * 100
    c = sol.clamp(-5.12, 5.12)
    return obj, torch.stack((c[:, : n // 2].sum(1), c[:, n // 2:].sum(1)), dim=1)


def run_gpu(iters, n=100, E=15, lam=36, c3=False, device_eval=False, pooled=True):
    from pyribs_b200.archives import GridArchive
    from pyribs_b200.emitters import EvolutionStrategyEmitter
    from pyribs_b200.schedulers import Scheduler
    mb = n / 2 * 5.12
    if c3:
        a = GridArchive(solution_dim=n, dims=(500, 500), ranges=[(-mb, mb)] * 2, learning_rate=0.01, threshold_min=0.0, seed=1)
    else:
        a = GridArchive(solution_dim=n, dims=(100, 100), ranges=[(-mb, mb)] * 2, seed=1)
    seeds = np.random.SeedSequence(1).spawn(E)
    kw = dict(ranker="imp", selection_rule="mu", restart_rule="basic") if c3 else dict(
        ranker="2imp", selection_rule="filter", restart_rule="no_improvement")
    ems = [EvolutionStrategyEmitter(a, x0=np.zeros(n), sigma0=0.5, batch_size=lam, seed=s, **kw) for s in seeds]
    s = Scheduler(a, ems, pool_emitters=pooled)
    if device_eval:
        import torch
        for _ in range(10):
            s.tell(*sphere_torch(s.ask(return_device=True)))
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(iters):
            s.tell(*sphere_torch(s.ask(return_device=True)))
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        return dict(it_per_s=iters / dt, solutions_per_s=iters * E * lam / dt, elites=len(a), qd=float(a.stats.qd_score),
                    restarts=sum(e.restarts for e in ems), evaluator="torch sphere on the GPU (ask/tell exchange CUDA tensors)")
    for _ in range(10):
        sol = s.ask(); s.tell(*sphere(sol))
    t0 = time.perf_counter(); t_ask = t_tell = 0.0
    for _ in range(iters):
        t1 = time.perf_counter(); sol = s.ask(); t2 = time.perf_counter()
        o, m = sphere(sol); t3 = time.perf_counter(); s.tell(o, m); t4 = time.perf_counter()
        t_ask += t2 - t1; t_tell += t4 - t3
    dt = time.perf_counter() - t0
    return dict(it_per_s=iters / dt, solutions_per_s=iters * E * lam / dt, ask_ms=t_ask / iters * 1e3,
                tell_ms=t_tell / iters * 1e3, elites=len(a), qd=float(a.stats.qd_score), restarts=sum(e.restarts for e in ems))

def run_cpu(iters, n=100, E=15, lam=36, c3=False):
    import qd_oracle as O
    mb = n / 2 * 5.12
    a = (O.make_grid_oracle(solution_dim=n, dims=(500, 500), ranges=[(-mb, mb)] * 2, learning_rate=0.01, threshold_min=0.0, seed=1)
         if c3 else O.make_grid_oracle(solution_dim=n, dims=(100, 100), ranges=[(-mb, mb)] * 2, seed=1))
    es = [O.OracleCMAES(0.5, n, lam, seed=i) for i in range(E)]
    for e in es: e.reset(np.zeros(n))
    def step():
        sols = [e.ask() for e in es]; sol = np.concatenate(sols); o, m = sphere(sol)
        info = a.add(sol, o, m)
        for i, e in enumerate(es):
            sl = slice(i * lam, (i + 1) * lam)
            if c3:
                idx, rv = O.rank_one_key(info["value"][sl]); new = lam // 2
                e.tell(idx, new)
                if e.check_stop(rv[idx]):
                    e.reset(a.sample_elites(1)["solution"][0])
                continue
            idx, rv = O.rank_two_stage(info["status"][sl], info["value"][sl])
            new = int(info["status"][sl].astype(bool).sum())
            e.tell(idx, new)
            if e.check_stop(rv[idx]) or new == 0:
                e.reset(a.sample_elites(1)["solution"][0])
    for _ in range(10): step()
    t0 = time.perf_counter()
    for _ in range(iters): step()
    dt = time.perf_counter() - t0
    return dict(it_per_s=iters / dt, solutions_per_s=iters * E * lam / dt, elites=len(a))

def run_reference(iters, n=100, E=15, lam=36, c3=False):
    """The same loop on the unmodified reference (ribs from baseline/_ref or /root/reference, through oracle/shim)."""
    for ref in (os.environ.get("QD_REFERENCE", "/root/reference"), os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isdir(os.path.join(ref, "ribs")):
            break
    else:
        return {"unavailable": "no reference package on this box"}
    for p in (os.path.join(ROOT, "oracle", "shim"), ref):
        if p not in sys.path:
            sys.path.insert(0, p)
    from ribs.archives import GridArchive
    from ribs.emitters import EvolutionStrategyEmitter
    from ribs.schedulers import Scheduler
    mb = n / 2 * 5.12
    if c3:
        a = GridArchive(solution_dim=n, dims=(500, 500), ranges=[(-mb, mb)] * 2, learning_rate=0.01, threshold_min=0.0, seed=1)
    else:
        a = GridArchive(solution_dim=n, dims=(100, 100), ranges=[(-mb, mb)] * 2, seed=1)
    kw = dict(ranker="imp", selection_rule="mu", restart_rule="basic") if c3 else dict(
        ranker="2imp", selection_rule="filter", restart_rule="no_improvement")
    ems = [EvolutionStrategyEmitter(a, x0=np.zeros(n), sigma0=0.5, batch_size=lam, seed=i, **kw) for i in range(E)]
    s = Scheduler(a, ems)
    for _ in range(3):
        sol = s.ask(); s.tell(*sphere(sol))
    t0 = time.perf_counter()
    for _ in range(iters):
        sol = s.ask(); s.tell(*sphere(sol))
    dt = time.perf_counter() - t0
    return dict(it_per_s=iters / dt, solutions_per_s=iters * E * lam / dt, elites=len(a),
                what="ribs.schedulers.Scheduler + ribs GridArchive + ribs EvolutionStrategyEmitter (cma_es), unmodified")


if __name__ == "__main__":
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    if len(sys.argv) > 2 and sys.argv[2] == "c3":
        print(json.dumps({"case": "C3 sphere CMA-MAE 100x512 n=100, grid 500x500", "gpu": run_gpu(iters, E=100, lam=512, c3=True),
                          "gpu_device_eval": run_gpu(iters, E=100, lam=512, c3=True, device_eval=True),
                          "gpu_unpooled": run_gpu(max(3, iters // 5), E=100, lam=512, c3=True, pooled=False),
                          "cpu_oracle_port": run_cpu(max(3, iters // 5), E=100, lam=512, c3=True),
                          "reference": run_reference(max(3, iters // 10), E=100, lam=512, c3=True), "cores": os.cpu_count()}))
    else:
        print(json.dumps({"case": "C1 sphere CMA-ME 15x36 n=100", "gpu": run_gpu(iters), "gpu_device_eval": run_gpu(iters, device_eval=True),
                          "gpu_unpooled": run_gpu(iters, pooled=False), "cpu_oracle_port": run_cpu(iters),
                          "reference": run_reference(iters), "cores": os.cpu_count()}))
