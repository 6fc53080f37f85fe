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
OBJ_DRIFT = 0.05  # objective offset per step: every step keeps inserting (a stationary stream would stop
                  # improving the archive after the first pass over the resident batches)
MAX_OBJ_STEPS = 2048  # distinct per-step objective vectors kept resident (0.8 MB each)


def make_inputs(seed, n_batches, B=None):
    B = B or WORKLOAD["batch"]
    rng = np.random.default_rng(seed)
    cen = np.random.default_rng(0).uniform(-1, 1, (WORKLOAD["centroids"], WORKLOAD["measure_dim"]))
    batches = []
    for j in range(n_batches):
        batches.append(dict(
            solution=rng.uniform(-1, 1, (B, WORKLOAD["solution_dim"])),
            objective=rng.standard_normal(B),  # + OBJ_DRIFT * step at use
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


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures
C_SIZEOF_RESULT = 120  # sizeof(qd_add_result)
NCU_TRAFFIC = {"cvt_bucket_kernel": 1_833_984, "insert_kernel_grid_2M": 311_130_624 + 205_896_704}


def grid_insert_roofline(torch, dev, peak_hbm, pipelined=False):
    """GridArchive insertion against the HBM roofline (north_star: >= 60 % of HBM peak at 1 GPU): CMA-MAE
    lr=0.01, 500x500 cells (the grid of configs[2]), n=100, 2M-row batches of uniform measures -- the size at
    which the insertion is throughput- rather than latency-bound. Algorithmic bytes per SURVEY 8(d):
    45 B/row + 1 648 B per winner."""
    from pyribs_b200.archives import GridArchive

    Bg, n = 2_000_000, 100
    a = GridArchive(solution_dim=n, dims=(500, 500), ranges=[(-1, 1)] * 2, learning_rate=0.01, threshold_min=0.0)
    a.pipeline_commit = pipelined
    g = torch.Generator(device=dev).manual_seed(1)
    bs = [dict(solution=torch.randn(Bg, n, dtype=torch.float64, device=dev, generator=g),
               objective=torch.rand(Bg, dtype=torch.float64, device=dev, generator=g) * 100 + 5.0 * j,
               measures=torch.rand(Bg, 2, dtype=torch.float64, device=dev, generator=g) * 2 - 1) for j in range(3)]
    for j in range(3):
        a.add(**bs[j])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 12
    e0.record()
    for j in range(iters):
        a.add(**bs[j % 3])  # 3 x 1.65 GB of inputs: nothing survives in the 126 MB L2 between calls
    a.join()  # pipelined: the last row copy belongs to the timed region
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    st = a.add(**bs[0])["status"]
    W = int(torch.unique(a.index_of(bs[0]["measures"])[st != 0]).numel())
    alg = Bg * 45 + W * (2 * n * 8 + 2 * 2 * 8 + 2 * 8)
    return {"kernel": "insert_kernel<double, double, 2, fused grid index> (GridArchive.add, "
                      + ("row phases and row copy as two launches on two streams: the copy of add k overlaps the row "
                         "phases of add k+1; archive.pipeline_commit = True)" if pipelined else "one launch per add)"),
            "workload": "GridArchive 500x500, CMA-MAE lr=0.01, n=100, 2M-row batches, uniform measures"
                        + (", back-to-back adds" if pipelined else ""),
            "bound": "hbm", "achieved": alg / ms / 1e6, "peak": peak_hbm, "unit": "GB/s", "frac": alg / ms / 1e6 / peak_hbm,
            "ms_per_add": ms, "insertions_per_s": Bg / ms * 1e3, "winners_per_add": W, "algorithmic_bytes": alg,
            "traffic": NCU_TRAFFIC.get("insert_kernel_grid_2M"),
            "note": "row passes are bound by scattered L1/L2 accesses (threshold gather + 3 atomics per row, ~6 L1tex "
                    "wavefront-cycles per row), the winner-row copy by HBM (TMA bulk, 5.3-5.9 TB/s); DESIGN.md 4.2"}


def es_roofline(torch, dev, peak_hbm):
    """BASELINE.json's second metric (emitter ask+tell solutions/sec) on configs[3]: sep-CMA-ES and LM-MA-ES,
    solution_dim 10 000, batch 4096, mu = 2048, device-resident (ask hands out the CUDA tensor, the ranking is a
    CUDA tensor), on-device ziggurat normals. CUDA events on the caller's stream, which waits for the strategy's
    own stream on both sides; median of 10 generations after 12 (LM-MA: all 20 direction vectors in use).
    Algorithmic bytes per SURVEY 8(d): sep ask lam*n*8 + 3*n*8, tell mu*n*8 + 8*n*8; LM-MA ask 2*lam*n*8 +
    m*n*8 (+ 4*lam*n*m flops on the fp64 tensor pipe), tell 2*mu*n*8 + 2*m*n*8."""
    import numpy as np

    from pyribs_b200.emitters.opt import _get_es

    n, lam, mu = 10_000, 4096, 2048
    out = []
    for es_name, kw, warm in (("sep_cma_es", {}, 4), ("lm_ma_es", {"n_vectors": 20}, 22)):
        es = _get_es(es_name, sigma0=0.5, solution_dim=n, batch_size=lam, seed=1, dtype=np.float64,
                     lower_bounds=np.full(n, -np.inf), upper_bounds=np.full(n, np.inf), **kw)
        es.reset(np.zeros(n))
        cur = torch.cuda.current_stream()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        t_ask, t_tell = [], []
        for g in range(warm + 10):
            torch.cuda.synchronize()
            ev[0].record()
            X = es.ask(return_device=True)  # the caller's stream waits for the strategy's stream
            ev[1].record()
            rank = torch.argsort((X - 1.0).pow(2).sum(1))  # sphere objective, best first
            es._es_stream().wait_stream(cur)
            ev[2].record(es._es_stream())
            es.tell(rank, None, mu)
            ev[3].record(es._es_stream())
            torch.cuda.synchronize()
            if g >= warm:
                t_ask.append(ev[0].elapsed_time(ev[1]))
                t_tell.append(ev[2].elapsed_time(ev[3]))
        a, t = float(np.median(t_ask)), float(np.median(t_tell))
        m = kw.get("n_vectors", 0)
        ask_b = lam * n * 8 + 3 * n * 8 if es_name == "sep_cma_es" else 2 * lam * n * 8 + m * n * 8
        tell_b = mu * n * 8 + 8 * n * 8 if es_name == "sep_cma_es" else 2 * mu * n * 8 + 2 * m * n * 8
        e = {"kernel": {"sep_cma_es": "sep_ask_rng_kernel (Philox + ziggurat fused with the transform) / "
                                      "wsum_partial_kernel + sep_tell_*",
                        "lm_ma_es": "fill_normal_zig + tall_dots (fp64 mma) + lmma_coef + lmma_out (fp64 mma) / "
                                    "wsum_partial_kernel x2 + lmma_tell_*"}[es_name] + ", csrc/es_stream.cu, es.cu",
             "workload": f"{es_name} n={n} lambda={lam} mu={mu}" + (f" n_vectors={m}" if m else ""),
             "metric": "emitter ask+tell solutions/sec", "solutions_per_s": lam / ((a + t) * 1e-3),
             "bound": "hbm", "achieved": (ask_b + tell_b) / ((a + t) * 1e-3) / 1e9, "peak": peak_hbm, "unit": "GB/s",
             "frac": (ask_b + tell_b) / ((a + t) * 1e-3) / 1e9 / peak_hbm,
             "ask_ms": a, "tell_ms": t, "ask_GBps": ask_b / a / 1e6, "tell_GBps": tell_b / t / 1e6,
             "algorithmic_bytes": ask_b + tell_b}
        if m:
            e["ask_fp64_TFLOPs"] = 4.0 * lam * n * m / (a * 1e-3) / 1e12
        out.append(e)
        del es, X
        torch.cuda.empty_cache()
    return out


def cpu_reference(cen, batches, steps, warmup, workers):
    """The oracle port (reference algorithm on the CPU). Returns (insertions/s, seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import qd_oracle as O

    a = O.make_cvt_oracle(solution_dim=WORKLOAD["solution_dim"], centroids=cen, workers=workers)

    def add(j):  # same drifting objectives as the GPU arm (the vector add is <1 % of a step)
        b = batches[j % len(batches)]
        a.add(solution=b["solution"], objective=b["objective"] + OBJ_DRIFT * j, measures=b["measures"])

    for j in range(warmup):
        add(j)
    t0 = time.perf_counter()
    for j in range(steps):
        add(warmup + j)
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
    ap.add_argument("--no-grid", action="store_true", help="skip the GridArchive 2M-row insertion roofline")
    ap.add_argument("--no-es", action="store_true", help="skip the configs[3] ask/tell measurement")
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
    # one objective vector per step (drifting upwards), resident on the device and in page-locked host memory
    n_obj = min(MAX_OBJ_STEPS, args.steps + args.warmup)
    obj_host = torch.empty((n_obj, B), dtype=torch.float64).pin_memory()
    for j in range(n_obj):
        obj_host[j] = torch.from_numpy(batches[j % N_RESIDENT]["objective"]) + OBJ_DRIFT * j
    obj_dev = obj_host.to(dev)
    obj_np = obj_host.numpy()

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
        archive.add(b["solution"], obj_dev[j % n_obj], b["measures"])

    with ClockSampler(local_rank) as clk:
        ms_dev, launches = timed(step_dev, args.steps, args.warmup)
    clocks = clk.summary()

    # ---- end to end: pinned host arrays in, NumPy status/value out ---------------------------
    archive.clear()
    host_batches = [{k: v.numpy() for k, v in b.items()} for b in pinned]

    def step_e2e(j):
        b = host_batches[j % N_RESIDENT]
        out = archive.add(b["solution"], obj_np[j % n_obj], b["measures"])
        assert out["status"].shape[0] == B

    zc0 = archive._zero_copy_bytes
    ms_e2e, _ = timed(step_e2e, args.steps, args.warmup)
    # bytes that crossed PCIe per step, counted from what was copied: measures + objective by DMA; the
    # solution rows either by DMA (all of them) or, for page-locked inputs when at most cells << batch rows
    # can win, only the winners' rows, read by the insert kernel straight from host memory (zero-copy)
    zc = (archive._zero_copy_bytes - zc0) / (args.steps + args.warmup)
    e2e_zero_copy = bool(archive._last_zero_copy)
    h2d = (batches[0]["measures"].nbytes + batches[0]["objective"].nbytes
           + (zc if e2e_zero_copy else batches[0]["solution"].nbytes))
    d2h = B * (4 + 8) + C_SIZEOF_RESULT

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
    peak_hbm = float(peaks.get("hbm_gbs", 6650.0))
    C_, d_ = WORKLOAD["centroids"], WORKLOAD["measure_dim"]
    flops = 2.0 * B * C_ * d_  # the contraction an exhaustive search performs
    s_k = ms_k / args.steps * 1e-3
    ach = flops / s_k / 1e12
    method = {0: "tensor", 1: "scan", 2: "bucket"}[archive._pick_method(B, C_)]
    if method == "bucket":
        kernel = "cvt_bucket_kernel<double, 2> (exact nearest centroid through a bucket grid, csrc/cvt_bucket.cu)"
        note = ("achieved = algorithmic flops of SURVEY 8(d) (2*B*C*d, what an exhaustive search computes) / measured "
                "kernel time. The bucket grid evaluates ~20 of the 10 000 centroids per query in fp64, so this is work "
                "avoided, not tensor-pipe utilisation; the kernel itself is a latency-bound gather over an L1/L2-resident "
                "220 KB structure (hbm view below). The tcgen05 search (--search tensor, the path of the 8-D / 1M-centroid "
                "configuration) measures 23.7 TFLOP/s algorithmic on this shape (profiles/r01_bench_line_n1.json).")
        # dram bytes of one launch: profiles/r01_cvt_bucket_insert_ncu_full.txt
        ncu_traffic = NCU_TRAFFIC.get("cvt_bucket_kernel")
    else:
        kernel = "cvt nearest-centroid search (cvt_tc_kernel + rescoring)" if method == "tensor" else "cvt_scan_fp64_kernel"
        note = ("d=2: 4 GFLOP of useful contraction per launch; the tcgen05 kernel is bound by the TMEM->register epilogue "
                "that must inspect every one of the B*C accumulators (measured tcgen05.ld ceiling 96 fp32/clk/SM, "
                "tools/tmem_bw.cu), not by the tensor pipe")
        ncu_traffic = 2_375_424 + 768 if method == "tensor" else None
    alg_bytes = (B + C_) * d_ * 8 + 4 * B
    roofline = {
        "kernel": kernel, "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
        "traffic": ncu_traffic, "note": note,
        "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops, burst)" if peaks else "fallback 1590",
        "flops_per_launch": flops, "ms_per_search": ms_k / args.steps, "launches_per_search": k_launches / args.steps,
        "queries_per_s": B / s_k,
        "hbm_view": {"algorithmic_bytes": alg_bytes, "achieved_GBps": alg_bytes / s_k / 1e9, "peak_GBps": peak_hbm,
                     "frac": alg_bytes / s_k / 1e9 / peak_hbm},
    }

    # ---- the other kernel of the step (insert_kernel) on this workload, from its own phase timers ----
    secondary = []
    try:
        import ctypes as C_t

        from pyribs_b200.archives._device_store import current_stream_ptr

        archive.clear()
        for j in range(8):
            step_dev(j)
        st = archive.add(dev_batches[8]["solution"], obj_dev[8 % n_obj], dev_batches[8]["measures"])["status"]
        t = (C_t.c_uint64 * 10)()
        _lib.check(_lib.lib().qd_insert_phase_times(archive._store._scratch.data_ptr(), archive._store.capacity, t,
                                                    current_stream_ptr()))
        ins_us = (t[8] - t[0]) / 1e3
        idx4 = archive.index_of(dev_batches[8]["measures"])
        W = int(torch.unique(idx4[st != 0]).numel())
        n_ = WORKLOAD["solution_dim"]
        ins_bytes = B * (4 + 8) + B * (8 + 1) + B * (4 + 8) + W * (2 * n_ * 8 + 2 * d_ * 8) + W * 16
        secondary.append({"kernel": "insert_kernel (persistent cooperative, csrc/insert.cu) on this workload",
                          "bound": "hbm", "achieved": ins_bytes / ins_us / 1e3, "peak": peak_hbm, "unit": "GB/s",
                          "frac": ins_bytes / ins_us / 1e3 / peak_hbm, "us_per_launch": ins_us, "winners": W,
                          "algorithmic_bytes": ins_bytes,
                          "note": "100k rows are far too few to fill a B200: 8 phases + grid barriers in 6-8 us"})
    except Exception as e:  # noqa: BLE001
        secondary.append({"kernel": "insert_kernel", "error": f"{type(e).__name__}: {e}"})
    if world == 1 and rank == 0 and not args.no_grid:
        try:
            secondary.append(grid_insert_roofline(torch, dev, peak_hbm))
            torch.cuda.empty_cache()
            secondary.append(grid_insert_roofline(torch, dev, peak_hbm, pipelined=True))
        except Exception as e:  # noqa: BLE001
            secondary.append({"kernel": "insert_kernel (GridArchive 2M rows)", "error": f"{type(e).__name__}: {e}"})
    if world == 1 and rank == 0 and not args.no_es:
        try:
            secondary.extend(es_roofline(torch, dev, peak_hbm))
        except Exception as e:  # noqa: BLE001
            secondary.append({"kernel": "evolution strategies (configs[3])", "error": f"{type(e).__name__}: {e}"})
    roofline["secondary"] = secondary

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
                   "l2": f"{N_RESIDENT} resident batches cycled (166 MB > 126 MB L2)",
                   "objective": f"standard normal + {OBJ_DRIFT} per step (every step keeps inserting)"},
        "clocks": clocks,
        "e2e": {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": "insertions/s",
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps,
                "transfers": "measures streamed by the search kernel and winners' rows gathered by the insert kernel "
                             "straight from the page-locked inputs (zero-copy over PCIe), objective by DMA, status / "
                             "value written by the kernel into page-locked result buffers"
                if e2e_zero_copy else "whole batch by DMA (H2D), status / value by DMA (D2H)"},
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
