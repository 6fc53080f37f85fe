import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
for n in (100, 256, 1000):
    A = torch.randn(n, n, dtype=torch.float64, device="cuda"); A = A @ A.T + n * torch.eye(n, dtype=torch.float64, device="cuda")
    print(f"n={n}: torch.linalg.eigh gpu {t(lambda: torch.linalg.eigh(A)):.3f} ms", end="; ")
    Ab = A.unsqueeze(0).repeat(15, 1, 1).contiguous()
    print(f"batched x15 {t(lambda: torch.linalg.eigh(Ab), 10):.3f} ms", end="; ")
    if n <= 256:
        Ab = A.unsqueeze(0).repeat(100, 1, 1).contiguous()
        print(f"batched x100 {t(lambda: torch.linalg.eigh(Ab), 5):.3f} ms", end="; ")
    An = A.cpu().numpy(); t0 = time.perf_counter()
    for _ in range(20): np.linalg.eigh(An)
    print(f"numpy cpu {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms")
from pyribs_b200.emitters.opt import CMAEvolutionStrategy
es = CMAEvolutionStrategy(0.5, 100, 36, seed=1); es.reset(np.zeros(100))
print("ask (no refresh) ms:", t(lambda: es.ask(), 100))
print("ask device  ms:", t(lambda: es.ask(return_device=True), 100))
print("forced refresh ms:", t(lambda: es.update_eigensystem(force=True), 30))
sol = es.ask(); fit = -np.sum(sol ** 2, 1); rk = np.flip(np.argsort(fit))
print("tell ms:", t(lambda: es.tell(rk, fit, 18), 100))
