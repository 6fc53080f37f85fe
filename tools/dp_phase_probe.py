"""Where the non-search part of a data-parallel add() goes (headline shape, batch rows sharded over the ranks):
CUDA-event times of search / push / barrier / insert on every rank plus the insert kernel's phase timers.
torchrun --nproc-per-node N tools/dp_phase_probe.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
from pyribs_b200 import _lib
from pyribs_b200.archives import CVTArchive
from tools.e2e_breakdown import phase_times

B, C, d, n = 1_000_000, 1_000_000, 8, 10
Bl = B // world
cen = np.random.default_rng(0).uniform(-1, 1, (C, d))
OWNED = len(sys.argv) > 1 and sys.argv[1] == "owned"  # owner-computes store instead of the replicated one
a = CVTArchive(solution_dim=n, centroids=cen, ranges=[(-1, 1)] * d, data_parallel=True, shard_store=OWNED)
rng = np.random.default_rng(1)
bs = []
for j in range(3):
    full = dict(solution=rng.uniform(-1, 1, (B, n)), objective=rng.standard_normal(B), measures=rng.uniform(-1, 1, (B, d)))
    bs.append({k: torch.from_numpy(np.ascontiguousarray(v[rank * Bl:(rank + 1) * Bl])).to(dev) for k, v in full.items()})

# instrument the sub-steps of add() by wrapping the archive's own helpers
ev = {}
def timed(name, fn):
    def w(*args, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(*args, **kw); e1.record()
        ev.setdefault(name, []).append((e0, e1))
        return out
    return w
a._index_of_device = timed("search", a._index_of_device)
a._symm_exchange = timed("exchange (2 pushes + barrier + flags)", a._symm_exchange)
L = _lib.lib()
class Wrap:  # the archive calls _lib.lib().qd_insert_sharded_rows(...): hand it a timed wrapper
    def __init__(self, L):
        self._L = L
        self._ins = timed("insert", L.qd_insert_sharded_rows)
        self._own = timed("insert", L.qd_insert_owned_cells)
    def __getattr__(self, k):
        return self._ins if k == "qd_insert_sharded_rows" else self._own if k == "qd_insert_owned_cells" else getattr(self._L, k)
_lib._lib = Wrap(L)
for j in range(6):
    a.add(bs[j % 3]["solution"], bs[j % 3]["objective"] + 0.05 * j, bs[j % 3]["measures"])
dist.barrier(); torch.cuda.synchronize()
ev.clear()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 12
t0.record()
for j in range(6, 6 + K):
    a.add(bs[j % 3]["solution"], bs[j % 3]["objective"] + 0.05 * j, bs[j % 3]["measures"])
t1.record(); torch.cuda.synchronize()
out = {"rank": rank, "world": world, "store": "owner computes" if OWNED else "replicated", "ms_per_add": t0.elapsed_time(t1) / K}
for k, v in ev.items():
    out[k + " ms"] = round(float(np.mean([e0.elapsed_time(e1) for e0, e1 in v])), 4)
ph, tot = phase_times(a)
out["insert_phases_us"] = {k: round(v, 1) for k, v in ph.items()}
out["insert_kernel_us"] = round(tot, 1)
a._sync(); out["winners_last_add_on_this_rank"] = int(a._store._last.n_winners)
print(json.dumps(out), flush=True)
dist.destroy_process_group()
