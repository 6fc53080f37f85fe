"""Copy-phase ceiling of the insert kernel: 250k winner rows of 800 B with sequential / permuted
sources and destinations (phase timers of qd_insert_phase_times)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyribs_b200.archives import GridArchive
from tools.e2e_breakdown import phase_times

rng = np.random.default_rng(0)
G, n = 500, 100
N = G * G
centers = (np.stack(np.meshgrid(np.arange(G), np.arange(G), indexing="ij"), -1).reshape(-1, 2) + 0.5) / G * 2 - 1

def run(name, cells, B_extra=0):
    a = GridArchive(solution_dim=n, dims=(G, G), ranges=[(-1, 1)] * 2)
    meas = centers[cells]
    obj = np.ones(len(cells))
    if B_extra:  # losers interleaved: winners are 1 row in (1 + B_extra / len(cells))
        k = B_extra // len(cells)
        meas = np.repeat(meas, k + 1, axis=0)
        obj = np.tile(np.concatenate([[1.0], np.zeros(k)]), len(cells))
    B = len(obj)
    sol = torch.randn(B, n, dtype=torch.float64, device="cuda")
    res = []
    for rep in range(3):
        a.clear()
        a.add(sol, torch.from_numpy(obj).cuda(), torch.from_numpy(meas).cuda())
        ph, tot = phase_times(a)
        assert len(a) == len(set(cells.tolist()))
        res.append(round(ph["copy"], 1))
    W = len(a)
    print(json.dumps({"case": name, "B": B, "winners": W, "copy_us": res, "GBps": round(W * 1632 / min(res) / 1e3, 1)}))

run("src sequential, dst sequential", np.arange(N))
run("src sequential, dst = cell-sorted from permuted rows", rng.permutation(N))
run("src 1-in-8 rows, dst sequential", np.arange(N), B_extra=7 * N)
run("src 1-in-8 rows of permuted cells", rng.permutation(N), B_extra=7 * N)
