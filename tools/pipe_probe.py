"""Timeline of pipelined GridArchive adds (archive.pipeline_commit): start / end of the row-phase launch (caller's
stream) and of the row-copy launch (side stream) of consecutive calls, relative to the first event."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyribs_b200.archives import GridArchive

Bg, n = 2_000_000, 100
a = GridArchive(solution_dim=n, dims=(500, 500), ranges=[(-1, 1)] * 2, learning_rate=0.01, threshold_min=0.0)
a.pipeline_commit = True
g = torch.Generator(device="cuda").manual_seed(1)
bs = [dict(solution=torch.randn(Bg, n, dtype=torch.float64, device="cuda", generator=g),
           objective=torch.rand(Bg, dtype=torch.float64, device="cuda", generator=g) * 100 + 5.0 * j,
           measures=torch.rand(Bg, 2, dtype=torch.float64, device="cuda", generator=g) * 2 - 1) for j in range(3)]
for j in range(4):
    a.add(**bs[j % 3])
torch.cuda.synchronize()
a._pipe_probe = []
t0 = torch.cuda.Event(enable_timing=True); t0.record()
for j in range(8):
    a.add(**bs[j % 3])
a.join()
torch.cuda.synchronize()
for k, ev in enumerate(a._pipe_probe):
    print(json.dumps({"call": k, "rows_start_us": round(t0.elapsed_time(ev[0]) * 1e3, 1), "rows_end_us": round(t0.elapsed_time(ev[1]) * 1e3, 1),
                      "copy_start_us": round(t0.elapsed_time(ev[2]) * 1e3, 1), "copy_end_us": round(t0.elapsed_time(ev[3]) * 1e3, 1)}))
