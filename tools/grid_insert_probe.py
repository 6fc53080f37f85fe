"""GridArchive insertion at the size bench.py's roofline line uses (500x500 cells, CMA-MAE lr=0.01, n=100, 2M-row
batches of uniform measures): ms per add over 30 adds and the phase timers of the last one. Optional argv[1] =
"pipe" turns on the pipelined commit."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyribs_b200.archives import GridArchive
from tools.e2e_breakdown import phase_times

dev = torch.device("cuda")
Bg, n = 2_000_000, 100
g = torch.Generator(device=dev); g.manual_seed(0)
a = GridArchive(solution_dim=n, dims=(500, 500), ranges=[(-1, 1)] * 2, learning_rate=0.01, threshold_min=0.0)
a.pipeline_commit = len(sys.argv) > 1 and sys.argv[1] == "pipe"
bs = [dict(solution=torch.randn(Bg, n, dtype=torch.float64, device=dev, generator=g),
           objective=torch.rand(Bg, dtype=torch.float64, device=dev, generator=g) * 100 + 5.0 * j,
           measures=torch.rand(Bg, 2, dtype=torch.float64, device=dev, generator=g) * 2 - 1) for j in range(3)]
for j in range(5):
    a.add(**bs[j % 3])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 30
e0.record()
for j in range(K):
    a.add(**bs[j % 3])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
ph, tot = phase_times(a)
W = int(a._store._last.n_winners) if hasattr(a._store, "_last") and a._store._last is not None else -1
alg = Bg * (16 + 8 + 4 + 4 + 8) + 248604 * 2 * (800 + 16) + 248604 * 3 * 8
print(json.dumps({"ms_per_add": ms, "GBps_alg": alg / ms / 1e6, "frac_of_6549": alg / ms / 1e6 / 6549.1,
                  "phases_us": {k: round(v, 1) for k, v in ph.items()}, "kernel_total_us": round(tot, 1), "len": len(a),
                  "pipelined": a.pipeline_commit}))
