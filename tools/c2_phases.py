"""Per-phase device time of the insert kernel on configs[1] (CVT 10k centroids, 100k-row batches, device tensors)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
import e2e_breakdown as E
from pyribs_b200.archives import CVTArchive
rng = np.random.default_rng(0)
cen = rng.uniform(-1, 1, (10_000, 2)); B = 100_000
a = CVTArchive(solution_dim=10, centroids=cen, ranges=[(-1, 1)] * 2)
dev = [dict(solution=torch.from_numpy(rng.uniform(-1, 1, (B, 10))).cuda(), objective=torch.from_numpy(rng.standard_normal(B) + 0.05 * j).cuda(),
            measures=torch.from_numpy(rng.uniform(-1, 1, (B, 2))).cuda()) for j in range(12)]
rows = []
for j in range(12):
    a.add(**dev[j]); torch.cuda.synchronize()
    ph, tot = E.phase_times(a); rows.append({"total_us": round(tot, 1), **{k: round(v, 1) for k, v in ph.items()}})
for r in rows[4:]: print(json.dumps(r))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for j in range(200): a.add(**dev[j % 12])
e1.record(); torch.cuda.synchronize(); print(json.dumps({"us_per_add_device": e0.elapsed_time(e1) / 200 * 1e3}))
