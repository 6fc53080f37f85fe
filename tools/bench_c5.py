"""BASELINE.json configs[4]: CVTArchive 1M centroids, 8-D measures, batch add of 1M solutions,
centroids sharded across the ranks (torchrun --nproc-per-node N tools/bench_c5.py).
Prints one JSON line on rank 0; the CPU reference (oracle port, SciPy KD-tree on all host cores) is
timed on a bounded sample of the same workload."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--centroids", type=int, default=1_000_000)
    ap.add_argument("--batch", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=8)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--cpu-sample", type=int, default=100_000)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from pyribs_b200 import _lib
    from pyribs_b200.archives import CVTArchive

    C, B, d, n = args.centroids, args.batch, args.dim, 10
    cen = np.random.default_rng(0).uniform(-1, 1, (C, d))
    rng = np.random.default_rng(1)
    batches = [dict(solution=rng.uniform(-1, 1, (B, n)), objective=rng.standard_normal(B) + 0.1 * j,
                    measures=rng.uniform(-1, 1, (B, d))) for j in range(2)]
    # rows are sharded over the ranks too (each rank uploads / owns B/world rows; measures, objective and
    # rows are all-gathered over NVLink), so host->device traffic per rank is 1/world of the batch
    a = CVTArchive(solution_dim=n, centroids=cen, ranges=[(-1, 1)] * d, shard_centroids=world > 1,
                   data_parallel=world > 1, search_method="tensor")
    Bl = B // world
    batches = [{k: v[rank * Bl:(rank + 1) * Bl] for k, v in b.items()} for b in batches] if world > 1 else batches
    devb = [{k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in b.items()} for b in batches]

    def step(j):
        b = devb[j % 2]
        return a.add(b["solution"], b["objective"], b["measures"])

    for j in range(2):
        out = step(j)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for j in range(args.steps):
        step(j)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # end to end: host batch in, NumPy status/value out
    a.clear()
    hb = [{k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory().numpy() for k, v in b.items()} for b in batches]
    a.add(**hb[0])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for j in range(3):
        a.add(**hb[(j + 1) % 2])
    e2e = (time.perf_counter() - t0) / 3
    if rank == 0:
        import qd_oracle as O

        S = args.cpu_sample
        ora = O.make_cvt_oracle(solution_dim=n, centroids=cen, workers=-1)
        S = min(S, len(batches[0]["objective"]))
        sb = {k: v[:S] for k, v in batches[0].items()}
        ora.add(**{k: v[:1000] for k, v in sb.items()})
        t0 = time.perf_counter()
        ref = ora.add(**sb)
        cpu_s = time.perf_counter() - t0
        # parity spot check on the sample: same indices as the KD-tree
        line = {
            "config": f"CVTArchive {C} centroids, {d}-D, batch add of {B}, centroids sharded over {world} GPU(s)",
            "n_gpus": world, "ms_per_add": ms, "insertions_per_s": B / ms * 1e3,
            "e2e_ms_per_add": e2e * 1e3, "e2e_insertions_per_s": B / e2e,
            "cpu_reference": {"insertions_per_s": S / cpu_s, "cores": os.cpu_count(), "sample_rows": S,
                              "kind": "oracle port, KD-tree workers=-1"},
            "speedup_device": (B / ms * 1e3) / (S / cpu_s), "speedup_e2e": (B / e2e) / (S / cpu_s),
            "gpu_launches_total": _lib.launch_count(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
