"""Where the time of one add() goes: pinned H2D bandwidth of the box, host-side cost of the NumPy
path, and the per-phase device time of the persistent insert kernel (qd_insert_phase_times)."""
import ctypes as C, cProfile, io, json, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyribs_b200 import _lib
from pyribs_b200.archives import CVTArchive, GridArchive
from pyribs_b200.archives._device_store import current_stream_ptr

PH = ["absmax", "classify", "mark(list mode)", "select(list mode)", "update(+resolve)", "best/count+copy", "barrier", "publish"]

def phase_times(a):
    out = (C.c_uint64 * 10)()
    _lib.check(_lib.lib().qd_insert_phase_times(a._store._scratch.data_ptr(), a._store.capacity, out, current_stream_ptr()))
    t = list(out)
    ph = {PH[i]: (t[i + 1] - t[i]) / 1e3 for i in range(8)}
    if t[9] > t[4]:
        ph["update: CTA 0 done after"] = (t[9] - t[4]) / 1e3
    return ph, (t[8] - t[0]) / 1e3

def h2d_bw():
    res = {}
    for mb in (2.4, 8, 64):
        n = int(mb * 1e6)
        h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
        for _ in range(3): d.copy_(h, non_blocking=True)
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): d.copy_(h, non_blocking=True)
        e1.record(); torch.cuda.synchronize()
        res[f"h2d_{mb}MB_GBps"] = n * 10 / e0.elapsed_time(e1) / 1e6
        e0.record()
        for _ in range(10): h.copy_(d, non_blocking=True)
        e1.record(); torch.cuda.synchronize()
        res[f"d2h_{mb}MB_GBps"] = n * 10 / e0.elapsed_time(e1) / 1e6
    return res

def c2():
    rng = np.random.default_rng(0)
    cen = rng.uniform(-1, 1, (10_000, 2)); B = 100_000
    a = CVTArchive(solution_dim=10, centroids=cen, ranges=[(-1, 1)] * 2)
    pin = [dict(solution=torch.from_numpy(rng.uniform(-1, 1, (B, 10))).pin_memory(), objective=torch.from_numpy(rng.standard_normal(B) + 0.05 * j).pin_memory(),
                measures=torch.from_numpy(rng.uniform(-1, 1, (B, 2))).pin_memory()) for j in range(8)]
    host = [{k: v.numpy() for k, v in b.items()} for b in pin]
    dev = [{k: v.cuda() for k, v in b.items()} for b in pin]
    for j in range(8): a.add(**host[j])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for j in range(200): a.add(**host[j % 8])
    torch.cuda.synchronize(); e2e_us = (time.perf_counter() - t0) / 200 * 1e6
    ph, tot = phase_times(a)
    a._sync_device_adds = True
    for j in range(8): a.add(**dev[j])
    t0 = time.perf_counter()
    for j in range(200): a.add(**dev[j % 8])
    torch.cuda.synchronize(); dev_sync_us = (time.perf_counter() - t0) / 200 * 1e6
    a._sync_device_adds = False
    t0 = time.perf_counter()
    for j in range(200): a.add(**dev[j % 8])
    t_enq = (time.perf_counter() - t0) / 200 * 1e6
    torch.cuda.synchronize(); dev_async_us = (time.perf_counter() - t0) / 200 * 1e6
    ph2, tot2 = phase_times(a)
    print(json.dumps({"case": "C2 add", "e2e_us": e2e_us, "device_in_sync_us": dev_sync_us, "device_in_async_us": dev_async_us,
                      "host_enqueue_us": t_enq, "insert_phases_us_host_batch": ph, "insert_phases_us_device_batch": ph2, "insert_total_us": tot2}))
    # search kernel: measures in HBM vs streamed from page-locked host memory; DMA of the same arrays
    flags = torch.zeros(1, dtype=torch.int32, device="cuda")
    def ev_time(fn, n=50):
        for _ in range(5): fn()
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3
    md = torch.empty_like(dev[0]["measures"]); od = torch.empty_like(dev[0]["objective"])
    print(json.dumps({"search_us_measures_in_hbm": ev_time(lambda: a._index_of_device(dev[0]["measures"], flags.data_ptr())),
                      "search_us_measures_in_pinned_host": ev_time(lambda: a._index_of_device(pin[0]["measures"], flags.data_ptr())),
                      "dma_measures_us": ev_time(lambda: md.copy_(pin[0]["measures"], non_blocking=True)),
                      "dma_objective_us": ev_time(lambda: od.copy_(pin[0]["objective"], non_blocking=True))}))
    for label, bs in (("numpy", host), ("device", dev)):
        pr = cProfile.Profile(); pr.enable()
        for j in range(300): a.add(**bs[j % 8])
        pr.disable(); torch.cuda.synchronize(); s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
        print("== cProfile", label); print(s.getvalue()[:4200])

def grid_large():
    rng = np.random.default_rng(0); B, n = 2_000_000, 100
    for lr, tmin, name in ((0.01, 0.0, "cma_mae"), (None, -np.inf, "cma_me")):
        a = GridArchive(solution_dim=n, dims=(500, 500), ranges=[(-1, 1)] * 2, learning_rate=lr, threshold_min=tmin)
        bs = [dict(solution=torch.randn(B, n, dtype=torch.float64, device="cuda"), objective=torch.from_numpy(rng.uniform(0, 100, B) + 5.0 * j).cuda(),
                   measures=torch.from_numpy(rng.uniform(-1, 1, (B, 2))).cuda()) for j in range(3)]
        rows = []
        for j in range(6):
            a.add(**bs[j % 3]); ph, tot = phase_times(a); rows.append((ph, tot))
        print(json.dumps({"case": f"grid 500x500 B=2M n=100 {name}", "calls": [{"total_us": t, **p} for p, t in rows[2:]]}))

if __name__ == "__main__":
    print(json.dumps(h2d_bw()))
    c2()
    grid_large()
