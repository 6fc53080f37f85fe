"""Steady-state large CMA-MAE GridArchive insertion for ncu: 500x500 cells, 2M rows x n=100."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyribs_b200.archives import GridArchive
rng = np.random.default_rng(0); B, n = 2_000_000, 100
a = GridArchive(solution_dim=n, dims=(500, 500), ranges=[(-1, 1)] * 2, learning_rate=0.01, threshold_min=0.0)
bs = [dict(solution=torch.randn(B, n, dtype=torch.float64, device="cuda"), objective=torch.from_numpy(rng.uniform(0, 100, B) + 5.0 * j).cuda(),
           measures=torch.from_numpy(rng.uniform(-1, 1, (B, 2))).cuda()) for j in range(3)]
for j in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    a.add(**bs[j % 3])
print(len(a))
