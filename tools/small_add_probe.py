"""Latency of small host-side add() calls (NumPy in, NumPy out): the reference's benchmark protocol
(benchmarks/cvt_add.py:79-86, 100 rows per add) and a few neighbours, with a cProfile of the 100-row CVT case."""
import cProfile, io, json, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyribs_b200.archives import CVTArchive, GridArchive

rng = np.random.default_rng(0)

def run(a, batches, n=1000):
    for j in range(50): a.add(**batches[j % len(batches)])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for j in range(n): a.add(**batches[j % len(batches)])
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6

def batches(B, n, d, k=8):
    return [dict(solution=rng.uniform(-1, 1, (B, n)), objective=rng.standard_normal(B) + 0.01 * j,
                 measures=rng.uniform(-1, 1, (B, d))) for j in range(k)]

out = {}
for cells in (10, 1000, 10_000):
    for B in (1, 100, 1000, 2048, 4096):
        a = CVTArchive(solution_dim=10, centroids=rng.uniform(-1, 1, (cells, 2)), ranges=[(-1, 1)] * 2)
        bs = batches(B, 10, 2)
        out[f"cvt2d_{cells}cells_B{B}_us"] = round(run(a, bs), 1)
        a._small_adds = False
        out[f"cvt2d_{cells}cells_B{B}_us_general_path"] = round(run(a, bs, 300), 1)
for B in (100, 540, 2048):
    a = GridArchive(solution_dim=100, dims=(100, 100), ranges=[(-1, 1)] * 2)
    bs = batches(B, 100, 2)
    out[f"grid100x100_n100_B{B}_us"] = round(run(a, bs), 1)
    a._small_adds = False
    out[f"grid100x100_n100_B{B}_us_general_path"] = round(run(a, bs, 300), 1)
a = CVTArchive(solution_dim=10, centroids=rng.uniform(-1, 1, (1000, 8)), ranges=[(-1, 1)] * 8)
out["cvt8d_1000cells_B100_us"] = round(run(a, batches(100, 10, 8)), 1)
print(json.dumps(out))
a = CVTArchive(solution_dim=10, centroids=rng.uniform(-1, 1, (1000, 2)), ranges=[(-1, 1)] * 2)
bs = batches(100, 10, 2)
run(a, bs, 200)
pr = cProfile.Profile(); pr.enable()
for j in range(2000): a.add(**bs[j % 8])
pr.disable(); s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[:4500])
