"""Host-side overhead of add(): tiny batches, device tensors in (so GPU work is negligible)."""
import cProfile, pstats, os, sys, time, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyribs_b200.archives import GridArchive, CVTArchive

def run(a, b, n=2000):
    for _ in range(50): a.add(**b)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): a.add(**b)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6

rng = np.random.default_rng(0)
B, n = 540, 100
a = GridArchive(solution_dim=n, dims=(100, 100), ranges=[(-256, 256)] * 2)
dev = dict(solution=torch.randn(B, n, dtype=torch.float64, device="cuda"), objective=torch.randn(B, dtype=torch.float64, device="cuda"),
           measures=torch.rand(B, 2, dtype=torch.float64, device="cuda") * 400 - 200)
host = {k: v.cpu().numpy() for k, v in dev.items()}
print(f"grid add, device in : {run(a, dev):7.1f} us/call")
print(f"grid add, numpy in  : {run(a, host):7.1f} us/call")
cen = rng.uniform(-1, 1, (1000, 2))
c = CVTArchive(solution_dim=n, centroids=cen, ranges=[(-1, 1)] * 2)
devc = dict(dev, measures=torch.rand(B, 2, dtype=torch.float64, device="cuda") * 2 - 1)
print(f"cvt  add, device in : {run(c, devc):7.1f} us/call")
pr = cProfile.Profile(); pr.enable()
for _ in range(2000): a.add(**dev)
pr.disable(); s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:5000])
