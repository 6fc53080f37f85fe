"""Benchmark of the hot path: batched CVTArchive insertion (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one `archive.add` of a 100 000-row synthetic batch (10-D solutions, 2-D
measures) into a 10 000-centroid CVTArchive. Prints ONE JSON line (rank 0).

  value      insertions/s with the batches already resident in HBM (CUDA tensors in,
             CUDA tensors out), CUDA-event timed, max over ranks
  e2e        the same metric through the NumPy-in / NumPy-out API: pinned-host -> device copy
             of the batch and device -> host copy of status/value inside the timed region
  roofline   the nearest-centroid kernel timed alone with CUDA events
  cpu_baseline  the oracle port (NumPy + SciPy KD-tree, the reference's algorithm) on the
             box's host cores, bounded sample

`--impl reference` times only the CPU oracle port on the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = {
    "workload": "cvt_add: CVTArchive 10k centroids, 2-D measures, batch add of 100k solutions (solution_dim=10)",
    "centroids": 10_000,
    "measure_dim": 2,
    "solution_dim": 10,
    "batch": 100_000,
}
N_RESIDENT = 16  # distinct batches cycled through: 16 x 10.4 MB = 166 MB > 126 MB L2


def make_inputs(seed, n_batches, B=None):
    B = B or WORKLOAD["batch"]
    rng = np.random.default_rng(seed)
    cen = np.random.default_rng(0).uniform(-1, 1, (WORKLOAD["centroids"], WORKLOAD["measure_dim"]))
    batches = []
    for j in range(n_batches):
        batches.append(dict(
            solution=rng.uniform(-1, 1, (B, WORKLOAD["solution_dim"])),
            # rising objectives: a realistic share of every batch keeps inserting
            objective=rng.standard_normal(B) + 0.05 * j,
            measures=rng.uniform(-1, 1, (B, WORKLOAD["measure_dim"])),
        ))
    return cen, batches


class ClockSampler:
    """Samples SM clocks / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index=0):
        self.samples = []
        self.index = index
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)
        self.max_mhz = None

    def _run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                    "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                    "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            while not self._stop.is_set():
                mhz = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.samples.append((mhz, [k for k, b in bits.items() if r & b]))
                self._stop.wait(0.005)
        except Exception as e:  # noqa: BLE001
            self.samples.append((None, [f"nvml unavailable: {type(e).__name__}"]))

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=5)

    def summary(self):
        sm = sorted(s[0] for s in self.samples if s[0] is not None)
        reasons = sorted({r for s in self.samples for r in s[1]})
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": reasons or ["unsampled"]}
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm)}


def cpu_reference(cen, batches, steps, warmup, workers):
    """The oracle port (reference algorithm on the CPU). Returns (insertions/s, seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import qd_oracle as O

    a = O.make_cvt_oracle(solution_dim=WORKLOAD["solution_dim"], centroids=cen, workers=workers)
    for j in range(warmup):
        a.add(**batches[j % len(batches)])
    t0 = time.perf_counter()
    for j in range(steps):
        a.add(**batches[(warmup + j) % len(batches)])
    dt = time.perf_counter() - t0
    return steps * WORKLOAD["batch"] / dt, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    cen, batches = make_inputs(42, 4)
    cores = os.cpu_count() or 1
    val, dt = cpu_reference(cen, batches, args.steps, args.warmup, workers=-1)
    line = {
        "impl": "reference", "metric": "archive insertions/sec (batch add)", "value": val, "unit": "insertions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(WORKLOAD),
        "cpu_baseline": {"value": val, "unit": "insertions/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} full 100k-row adds, oracle port, KD-tree workers=-1"},
        "e2e": {"value": val, "unit": "insertions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--search", default="auto", choices=["auto", "tensor", "scan", "bucket"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from pyribs_b200 import _lib
    from pyribs_b200.archives import CVTArchive

    B = WORKLOAD["batch"]
    # weak scaling: every rank contributes its own 100k-row slice of a (world x 100k)-row global batch;
    # indices are found locally, (index, objective, rows) are all-gathered over NCCL and every rank
    # commits the global batch to its replica
    cen, batches = make_inputs(42 + rank, N_RESIDENT)
    archive = CVTArchive(solution_dim=WORKLOAD["solution_dim"], centroids=cen,
                         ranges=[(-1, 1)] * WORKLOAD["measure_dim"], search_method=args.search,
                         data_parallel=world > 1)
    dev_batches = [{k: torch.from_numpy(v).to(dev) for k, v in b.items()} for b in batches]
    pinned = [{k: torch.from_numpy(v).pin_memory() for k, v in b.items()} for b in batches]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for j in range(warmup):
            fn(j)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _lib.launch_count()
        e0.record()
        for j in range(steps):
            fn(warmup + j)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = _lib.launch_count() - n0
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    # ---- device-resident throughput -------------------------------------------------------
    def step_dev(j):
        b = dev_batches[j % N_RESIDENT]
        archive.add(b["solution"], b["objective"], b["measures"])

    with ClockSampler(local_rank) as clk:
        ms_dev, launches = timed(step_dev, args.steps, args.warmup)
    clocks = clk.summary()

    # ---- end to end: pinned host arrays in, NumPy status/value out ---------------------------
    archive.clear()
    host_batches = [{k: v.numpy() for k, v in b.items()} for b in pinned]

    def step_e2e(j):
        b = host_batches[j % N_RESIDENT]
        out = archive.add(b["solution"], b["objective"], b["measures"])
        assert out["status"].shape[0] == B

    ms_e2e, _ = timed(step_e2e, args.steps, args.warmup)
    h2d = sum(v.nbytes for v in batches[0].values())
    d2h = B * (4 + 8) + 56

    # ---- dominant kernel alone: nearest-centroid search -------------------------------------
    flags = torch.zeros(1, dtype=torch.int32, device=dev)

    def step_search(j):
        archive._index_of_device(dev_batches[j % N_RESIDENT]["measures"], flags.data_ptr())

    ms_k, k_launches = timed(step_search, args.steps, args.warmup)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops", 1590.0))
    flops = 2.0 * B * WORKLOAD["centroids"] * WORKLOAD["measure_dim"]
    ach = flops / (ms_k / args.steps * 1e-3) / 1e12
    # dram bytes of one cvt_tc_kernel launch from the committed `ncu --set full` capture
    # (profiles/r01_top_kernels_ncu_full.txt: 2.375 MB read + 768 B written)
    ncu_traffic = 2_375_424 + 768
    pairs = float(B) * WORKLOAD["centroids"]
    roofline = {
        "kernel": "cvt nearest-centroid search (cvt_tc_kernel + rescoring)", "bound": "tensor", "achieved": ach,
        "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": ncu_traffic,
        "note": "d=2: 4 GFLOP of useful contraction per launch; the kernel is bound by the TMEM->register epilogue "
                "that must inspect every one of the B*C accumulators (measured tcgen05.ld ceiling 96 fp32/clk/SM, "
                "tools/tmem_bw.cu), not by the tensor pipe",
        "pairs_per_s": pairs / (ms_k / args.steps * 1e-3),
        "tmem_read_ceiling_pairs_per_s": 96 * 148 * 1.965e9,
        "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops, burst)" if peaks else "fallback 1590",
        "flops_per_launch": flops, "ms_per_search": ms_k / args.steps, "launches_per_search": k_launches / args.steps,
    }

    line = {
        "metric": "archive insertions/sec (batch add)",
        "value": world * B * args.steps / (ms_dev * 1e-3),
        "unit": "insertions/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {**WORKLOAD, "global_batch": world * B, "search": args.search,
                   "parallelism": f"dp{world}: batch-sharded search + all-gather + replicated insert" if world > 1
                   else "single GPU",
                   "l2": f"{N_RESIDENT} resident batches cycled (166 MB > 126 MB L2)"},
        "clocks": clocks,
        "e2e": {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": "insertions/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "roofline": roofline,
    }
    if rank == 0:
        cores = os.cpu_count() or 1
        hb = [{k: v for k, v in b.items()} for b in batches[:4]]
        v1, t1 = cpu_reference(cen, hb, 150, 1, workers=1)
        vN, tN = cpu_reference(cen, hb, 1000, 2, workers=-1)
        best = max(v1, vN)
        line["cpu_baseline"] = {
            "value": best, "unit": "insertions/s", "cores": cores if vN >= v1 else 1, "kind": "port",
            "sample": f"oracle port (NumPy + SciPy KD-tree): 150 adds of 100k rows with workers=1 "
                      f"({v1:.3g}/s, {t1:.1f}s) and 1000 adds with workers=-1 on {cores} cores ({vN:.3g}/s, {tN:.1f}s)",
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
