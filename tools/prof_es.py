"""Runs a few generations of the configs[3] strategies (n=10 000, lambda=4096) so that an ncu launch list
(`ncu --metrics gpu__time_duration.sum`) shows where ask / tell time goes. Usage: prof_es.py [es ...]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyribs_b200.emitters.opt import _get_es

def run(es_name, n=10_000, lam=4096, gens=6, **kw):
    es = _get_es(es_name, sigma0=0.5, solution_dim=n, batch_size=lam, seed=1, dtype=np.float64,
                 lower_bounds=np.full(n, -np.inf), upper_bounds=np.full(n, np.inf), **kw)
    es.reset(np.zeros(n))
    ta, tt = [], []
    for g in range(gens):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        X = es.ask(return_device=True)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        fit = -(X - 1.0).pow(2).sum(1)
        rank = torch.argsort(fit, descending=True)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        es.tell(rank, None, lam // 2)
        torch.cuda.synchronize(); t3 = time.perf_counter()
        ta.append((t1 - t0) * 1e3); tt.append((t3 - t2) * 1e3)
    print(json.dumps({"es": es_name, **kw, "ask_ms": [round(x, 3) for x in ta], "tell_ms": [round(x, 3) for x in tt]}), flush=True)

if __name__ == "__main__":
    which = sys.argv[1:] or ["sep", "lm20", "lm"]
    if "sep" in which: run("sep_cma_es")
    if "lm20" in which: run("lm_ma_es", gens=24, n_vectors=20)
    if "lm" in which: run("lm_ma_es", gens=8)
