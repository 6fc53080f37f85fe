"""Host-side cost of ask + tell at configs[3] scale (the GPU work is ~0.16 ms per sep-CMA generation): cProfile of
200 generations with a fixed device ranking. Usage: es_host_profile.py [sep_cma_es|lm_ma_es]"""
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pyribs_b200.emitters.opt import _get_es


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "sep_cma_es"
    n, lam = 10_000, 4096
    kw = {"n_vectors": 20} if name == "lm_ma_es" else {}
    es = _get_es(name, sigma0=0.5, solution_dim=n, batch_size=lam, seed=1, dtype=np.float64,
                 lower_bounds=np.full(n, -np.inf), upper_bounds=np.full(n, np.inf), **kw)
    es.reset(np.zeros(n))
    rank = torch.randperm(lam, device="cuda")

    def gens(k):
        for _ in range(k):
            es.ask(return_device=True)
            es.tell(rank, None, lam // 2)

    gens(30)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    gens(200)
    t1 = time.perf_counter()  # host time to enqueue (the GPU may lag behind)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name}: host {1e6 * (t1 - t0) / 200:.1f} us per generation to enqueue, {1e6 * (t2 - t0) / 200:.1f} us wall")
    pr = cProfile.Profile()
    pr.enable()
    gens(200)
    pr.disable()
    torch.cuda.synchronize()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(32)
    print(s.getvalue()[:6000])


if __name__ == "__main__":
    main()
