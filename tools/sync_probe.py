"""How long does the host wait for a tiny kernel on this box: blocking synchronize vs polling."""
import time, json, ctypes, torch
x = torch.zeros(1, device="cuda"); s = torch.cuda.current_stream()
def run(wait, n=300):
    for _ in range(20):
        x.add_(1); wait()
    t0 = time.perf_counter()
    for _ in range(n):
        x.add_(1); wait()
    return (time.perf_counter() - t0) / n * 1e6
ev = torch.cuda.Event()
def poll_stream():
    while not s.query(): pass
def poll_event():
    ev.record()
    while not ev.query(): pass
rt = ctypes.CDLL("libcudart.so.12")
raw = ctypes.c_void_p(s.cuda_stream)
def raw_sync():
    rt.cudaStreamSynchronize(raw)
def raw_poll():
    while rt.cudaStreamQuery(raw) != 0: pass
out = {"torch_synchronize_us": run(torch.cuda.synchronize), "stream_synchronize_us": run(s.synchronize), "stream_query_poll_us": run(poll_stream),
       "event_query_poll_us": run(poll_event), "raw_cudaStreamSynchronize_us": run(raw_sync), "raw_cudaStreamQuery_poll_us": run(raw_poll)}
# launch only (no wait), then one sync
t0 = time.perf_counter()
for _ in range(300): x.add_(1)
t1 = time.perf_counter(); torch.cuda.synchronize()
out["launch_only_us"] = (t1 - t0) / 300 * 1e6
# pinned flag written by a D2H copy, host spins on memory
flag = torch.zeros(1, dtype=torch.int64).pin_memory(); d = torch.zeros(1, dtype=torch.int64, device="cuda")
import numpy as np
fl = flag.numpy()
def spin_flag_factory():
    state = {"k": 0}
    def wait():
        state["k"] += 1
        d.fill_(state["k"]); flag.copy_(d, non_blocking=True)
        while fl[0] != state["k"]: pass
    return wait
out["d2h_flag_spin_us"] = run(spin_flag_factory())
print(json.dumps(out))
