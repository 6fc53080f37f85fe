"""What bounds the NumPy-in / NumPy-out add() of configs[1] (100k rows, 10k centroids): PCIe bandwidth of the box,
search with measures in HBM vs page-locked host memory, DMA times, and the end-to-end step."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import e2e_breakdown as E

print(json.dumps(E.h2d_bw()))
rng = np.random.default_rng(0)
cen = rng.uniform(-1, 1, (10_000, 2)); B = 100_000
from pyribs_b200.archives import CVTArchive
a = CVTArchive(solution_dim=10, centroids=cen, ranges=[(-1, 1)] * 2)
pin = [dict(solution=torch.from_numpy(rng.uniform(-1, 1, (B, 10))).pin_memory(), objective=torch.from_numpy(rng.standard_normal(B) + 0.05 * j).pin_memory(),
            measures=torch.from_numpy(rng.uniform(-1, 1, (B, 2))).pin_memory()) for j in range(8)]
host = [{k: v.numpy() for k, v in b.items()} for b in pin]
dev = [{k: v.cuda() for k, v in b.items()} for b in pin]
for j in range(8): a.add(**host[j])
torch.cuda.synchronize(); t0 = time.perf_counter()
for j in range(300): a.add(**host[j % 8])
torch.cuda.synchronize(); print(json.dumps({"e2e_us": (time.perf_counter() - t0) / 300 * 1e6, "zero_copy": bool(a._last_zero_copy)}))
flags = torch.zeros(1, dtype=torch.int32, device="cuda")
def ev_time(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3
md = torch.empty_like(dev[0]["measures"]); od = torch.empty_like(dev[0]["objective"]); sd = torch.empty_like(dev[0]["solution"])
st = torch.empty(B, dtype=torch.int32).pin_memory(); std = torch.empty(B, dtype=torch.int32, device="cuda")
print(json.dumps({"search_us_measures_in_hbm": ev_time(lambda: a._index_of_device(dev[0]["measures"], flags.data_ptr())),
                  "search_us_measures_in_pinned_host": ev_time(lambda: a._index_of_device(pin[0]["measures"], flags.data_ptr())),
                  "dma_measures_us": ev_time(lambda: md.copy_(pin[0]["measures"], non_blocking=True)),
                  "dma_objective_us": ev_time(lambda: od.copy_(pin[0]["objective"], non_blocking=True)),
                  "dma_solution_us": ev_time(lambda: sd.copy_(pin[0]["solution"], non_blocking=True)),
                  "dma_status_d2h_us": ev_time(lambda: st.copy_(std, non_blocking=True))}))
# one-call latency pieces: empty kernel + sync, and the device-batch add with sync
a._sync_device_adds = True
t0 = time.perf_counter()
for j in range(300): a.add(**dev[j % 8])
torch.cuda.synchronize(); print(json.dumps({"device_in_sync_us": (time.perf_counter() - t0) / 300 * 1e6}))
x = torch.zeros(1, device="cuda")
t0 = time.perf_counter()
for j in range(300):
    x.add_(1); torch.cuda.synchronize()
print(json.dumps({"tiny_kernel_plus_sync_us": (time.perf_counter() - t0) / 300 * 1e6}))
