"""Diagnostics for the tcgen05 CVT search: dumps raw scores / margins / candidate counts."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyribs_b200 import _lib
from pyribs_b200.archives import CVTArchive

def run(B, C, d, seed=0):
    rng = np.random.default_rng(seed)
    cen = rng.uniform(-1, 1, (C, d))
    meas = rng.uniform(-1, 1, (B, d))
    a = CVTArchive(solution_dim=1, centroids=cen, ranges=[(-1, 1)] * d, search_method="tensor")
    ld = 256 * ((C + 255) // 256)
    scores = torch.full((B, ld), float("nan"), device="cuda", dtype=torch.float32)
    eps = torch.zeros(B, device="cuda", dtype=torch.float32)
    _lib.lib().qd_cvt_debug_buffers(scores.data_ptr(), eps.data_ptr())
    flags = torch.zeros(1, dtype=torch.int32, device="cuda")
    m = torch.from_numpy(meas).cuda()
    idx = a._index_of_device(m, flags.data_ptr())
    torch.cuda.synchronize()
    _lib.lib().qd_cvt_debug_buffers(None, None)
    sc = a._cvt_scratch
    rc = sc[:4096].view(torch.int32).cpu().numpy()
    nout = int(sc[4096:4100].view(torch.int32).item())
    exact = (cen ** 2).sum(1)[None, :] - 2 * meas @ cen.T
    got = scores.cpu().numpy()[:, :C].astype(np.float64)
    # scores are in the scaled domain: s^2 * t * exact ; with |c|<=1 s=1 (maybe 1 or 2), t=1
    ratio = np.nanmedian(got / exact)
    err = np.abs(got - exact * ratio)
    print(f"B={B} C={C} d={d}: flags={int(flags.item())} cands={rc.sum()} ({rc.sum()/B:.1f}/query) maxregion={rc.max()} "
          f"outliers={nout} nan={np.isnan(got).sum()} scale_ratio={ratio:.4f} max|err|={np.nanmax(err):.3e} "
          f"eps[min,max]=({eps.min().item():.3e},{eps.max().item():.3e})")
    ref = np.argmin(((meas[:, None, :] - cen[None]) ** 2).sum(-1), 1) if B * C < 5e7 else None
    if ref is not None:
        print("   mismatches vs numpy:", int((idx.cpu().numpy() != ref).sum()))
    return got, exact

if __name__ == "__main__":
    run(256, 512, 2)
    run(1000, 3000, 2)
    run(8000, 10000, 2)
    run(4096, 4096, 8)
    run(4096, 4096, 13)


def golden_overflow():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from golden_util import load_golden
    g = load_golden("cvt_index")
    for name, c in g.items():
        d = c["centroids"].shape[1]
        a = CVTArchive(solution_dim=1, centroids=c["centroids"], ranges=[(-1, 1)] * d, dtype=c["measures"].dtype,
                       search_method="tensor")
        flags = torch.zeros(1, dtype=torch.int32, device="cuda")
        m = torch.from_numpy(c["measures"]).cuda()
        a._index_of_device(m, flags.data_ptr())
        torch.cuda.synchronize()
        rc = a._cvt_scratch[:4096].view(torch.int32).cpu().numpy()
        print(name, c["measures"].shape, c["centroids"].shape, "flags", int(flags.item()), "cands", rc.sum(), "max", rc.max())


if __name__ == "__main__":
    golden_overflow()
