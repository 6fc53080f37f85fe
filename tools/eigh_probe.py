"""Jacobi eigensolver on CMA-ES-like covariance chains: time and sweeps per refresh (15 matrices, n = 100 / 36 / 150)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyribs_b200 import _lib
L = _lib.lib(); sp = torch.cuda.current_stream().cuda_stream
out = []
for n in (100, 36, 150):
    batch = 15
    rng = np.random.default_rng(0)
    C = np.stack([np.eye(n) for _ in range(batch)])
    Cd = torch.from_numpy(C).cuda(); eig = torch.empty(batch, n, dtype=torch.float64, device="cuda")
    V = torch.empty(batch, n, n, dtype=torch.float64, device="cuda"); sw = torch.zeros(batch, dtype=torch.int32, device="cuda")
    c1cmu = 1.7e-3
    for refresh in range(6):
        for it in range(3):
            y = rng.standard_normal((batch, n, 18)) * (1.0 + 0.3 * refresh)
            C = (1 - c1cmu * 18) * C + c1cmu * np.einsum("bik,bjk->bij", y, y)
        C = np.maximum(C, C.transpose(0, 2, 1))
        Cd.copy_(torch.from_numpy(C))
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); _lib.check(L.qd_sym_eigh_batched(n, batch, Cd.data_ptr(), eig.data_ptr(), V.data_ptr(), sw.data_ptr(), sp)); e1.record()
        torch.cuda.synchronize()
        Vh = V.cpu().numpy(); e = eig.cpu().numpy()
        err = max(np.abs((Vh[b] * e[b]) @ Vh[b].T - C[b]).max() for b in range(batch))
        orth = max(np.abs(Vh[b].T @ Vh[b] - np.eye(n)).max() for b in range(batch))
        out.append({"n": n, "refresh": refresh, "ms": round(e0.elapsed_time(e1), 3), "sweeps": sw.cpu().tolist()[:4],
                    "recon_err": float(err), "orth_err": float(orth)})
print(json.dumps(out))
