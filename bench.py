"""Benchmark of the hot path: batched CVTArchive insertion at the size `north_star` sets its target on
(BASELINE.json configs[4]): 1M centroids, 8-D measures, batch add of 1M solutions.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    torchrun --nproc-per-node N bench.py --gpus N ...          (one rank per GPU, NCCL)

One "step" = one `archive.add` of a 1M-row synthetic batch (10-D solutions, 8-D measures). Prints ONE JSON line
(rank 0).

  value      insertions/s with the batches already resident in HBM (CUDA tensors in, CUDA tensors out),
             CUDA-event timed, max over ranks
  e2e        the same metric through the NumPy-in / NumPy-out API: host -> device copy of the batch (page-locked
             inputs) and device -> host copy of status / value inside the timed region; `e2e.pageable` repeats it
             with ordinary (pageable) NumPy arrays, which is what a reference user passes
  roofline   the nearest-centroid kernel timed alone with CUDA events; the GridArchive insertion (the kernel of
             this path that runs in the HBM-bandwidth regime), configs[1] and the ES streams under `secondary`
  cpu_baseline  the UNMODIFIED reference (`ribs` from baseline/_ref or /root/reference through oracle/shim) on the
             box's host cores, same inputs, full 1M-row adds; its status / value are compared with the GPU's

N > 1 is STRONG scaling of the same 1M-row add: every rank takes B/N rows of the batch, finds their cells against
the full centroid set, the ranks exchange (cell, objective) over NVLink peer memory, and the insertion gathers the
winners' rows from their owners. The centroid-sharded search `north_star` describes (local argmin + all-reduce-min
of distances, then of indices) is measured beside it (`alt`).

`--impl reference` times only the reference on the same workload (rank 0; other ranks exit).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = {
    "workload": "cvt_add_1M: CVTArchive 1M centroids, 8-D measures, batch add of 1M solutions (solution_dim=10) "
                "[BASELINE.json configs[4]]",
    "centroids": 1_000_000,
    "measure_dim": 8,
    "solution_dim": 10,
    "batch": 1_000_000,
}
N_RESIDENT = 3      # distinct batches cycled through: 3 x 152 MB = 456 MB > 126 MB L2
OBJ_DRIFT = 0.05    # objective offset per step: every step keeps inserting
METRIC = "archive insertions/sec (batch add)"


OWNER_TEXT = ("strong scaling over {world} GPUs: batch rows sharded (B/{world} per rank), search index replicated, "
              "(cell, objective) exchanged over NVLink peer memory (12 B per row), STORE sharded by cell range -- every "
              "rank classifies / commits only the rows of its own cells, pulls only its own winners' rows from their "
              "owners and pushes status / value back to the rows' ranks (owner computes)")
REPLICATED_TEXT = ("strong scaling over {world} GPUs: batch rows sharded (B/{world} per rank), search index replicated, "
                   "(cell, objective) exchanged over NVLink peer memory, every rank commits the global batch to its own "
                   "replica of the store, winners' rows gathered from their owners by the insert kernel")


def make_config(world):
    cfg = dict(WORKLOAD)
    cfg["global_batch"] = WORKLOAD["batch"]
    cfg["parallelism"] = "single GPU" if world == 1 else OWNER_TEXT.format(world=world)
    cfg["l2"] = f"{N_RESIDENT} resident batches cycled (456 MB > 126 MB L2)"
    cfg["objective"] = f"standard normal + {OBJ_DRIFT} per step (every step keeps inserting)"
    return cfg


def make_centroids():
    return np.random.default_rng(0).uniform(-1, 1, (WORKLOAD["centroids"], WORKLOAD["measure_dim"]))


def make_batches(n_batches, B=None, seed=1):
    B = B or WORKLOAD["batch"]
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_batches):
        out.append(dict(solution=rng.uniform(-1, 1, (B, WORKLOAD["solution_dim"])),
                        objective=rng.standard_normal(B),
                        measures=rng.uniform(-1, 1, (B, WORKLOAD["measure_dim"]))))
    return out


# --------------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clocks / throttle reasons through NVML: once before, every 5 ms during, once after."""

    def __init__(self, index=0):
        self.samples = []
        self.index = index
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)
        self.max_mhz = None
        self._h = None
        self._nv = None
        self._bits = {}
        try:
            import pynvml as nv

            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self._h = nv.nvmlDeviceGetHandleByIndex(phys)
            self._nv = nv
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM))
            self._bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                          "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                          "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                          "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        except Exception as e:  # noqa: BLE001
            self.samples.append((None, [f"nvml unavailable: {type(e).__name__}"]))

    def sample(self):
        if self._h is None:
            return
        nv = self._nv
        try:
            mhz = float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
            r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
            self.samples.append((mhz, [k for k, b in self._bits.items() if r & b]))
        except Exception as e:  # noqa: BLE001
            self.samples.append((None, [f"nvml error: {type(e).__name__}"]))

    def _run(self):
        while not self._stop.is_set():
            self.sample()
            self._stop.wait(0.005)

    def __enter__(self):
        self.sample()
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=5)
        self.sample()

    def summary(self):
        sm = sorted(s[0] for s in self.samples if s[0] is not None)
        reasons = sorted({r for s in self.samples for r in s[1]})
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": reasons or ["unsampled"]}
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------------------------------------------------
# the reference on the host cores
# --------------------------------------------------------------------------------------------------------
def load_reference():
    """Imports the UNMODIFIED reference (`ribs`): baseline/_ref (pip --target install, travels to the GPU box) or
    /root/reference, with oracle/shim for the two pure-Python dependencies this image lacks (array_api_compat ->
    SciPy's vendored copy, numpy_groupies -> numba restatement of the three reductions the path uses). Returns
    (module, where) or (None, why)."""
    shim = os.path.join(ROOT, "oracle", "shim")
    for where in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(where, "ribs")):
            for p in (shim, where):
                if p not in sys.path:
                    sys.path.insert(0, p)
            try:
                import ribs  # noqa: F401

                return ribs, where
            except Exception as e:  # noqa: BLE001
                return None, f"import ribs from {where} failed: {type(e).__name__}: {e}"
    return None, "no baseline/_ref and no /root/reference"


class CpuArm:
    """The reference's CVTArchive (or, if it cannot be imported, the oracle port) on the workload of this file."""

    def __init__(self, cen, workers=-1, centroids_kw=None):
        ribs, where = load_reference()
        self.kind = "reference" if ribs is not None else "port"
        self.where = where
        n, d = WORKLOAD["solution_dim"], cen.shape[1]
        t0 = time.perf_counter()
        if ribs is not None:
            from ribs.archives import CVTArchive

            self.archive = CVTArchive(solution_dim=n, centroids=cen, ranges=[(-1, 1)] * d,
                                      kdtree_query_kwargs={"workers": workers}, **(centroids_kw or {}))
        else:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import qd_oracle as O

            self.archive = O.make_cvt_oracle(solution_dim=n, centroids=cen, workers=workers)
        self.build_s = time.perf_counter() - t0
        self.workers = workers

    def add(self, b, drift=0.0, rows=None):
        sl = slice(None) if rows is None else slice(0, rows)
        return self.archive.add(solution=b["solution"][sl], objective=b["objective"][sl] + drift,
                                measures=b["measures"][sl])

    def describe(self):
        if self.kind == "reference":
            return f"unmodified ribs.archives.CVTArchive.add from {self.where} (scipy KD-tree, workers={self.workers})"
        return f"oracle port ({self.where}); NumPy + SciPy KD-tree, workers={self.workers}"


def run_reference(args, rank, world):
    """`--impl reference`: the reference alone, every step a 1M-row add (shortened only if the whole run would
    exceed ~4 minutes on this box)."""
    if rank != 0:
        return
    cen = make_centroids()
    batches = make_batches(2)
    cores = os.cpu_count() or 1
    arm = CpuArm(cen)
    B = WORKLOAD["batch"]
    arm.add(batches[0], rows=2000)  # numba compilation of the group reductions
    arm.archive.clear()
    t0 = time.perf_counter()
    arm.add(batches[0])
    t_first = time.perf_counter() - t0
    total = args.steps + args.warmup
    rows = B
    if t_first * total > 240.0:
        rows = max(B // 8, int(B * 240.0 / (t_first * total)) // 1000 * 1000)
    for j in range(1, args.warmup):
        arm.add(batches[j % 2], OBJ_DRIFT * j, rows)
    t0 = time.perf_counter()
    for j in range(args.steps):
        arm.add(batches[(args.warmup + j) % 2], OBJ_DRIFT * (args.warmup + j), rows)
    dt = time.perf_counter() - t0
    val = args.steps * rows / dt
    sample = (f"{args.steps} adds of {rows} rows into the 1M-centroid archive"
              + ("" if rows == B else f" (row sample of the 1M-row batch: a full add takes {t_first:.1f} s here)")
              + f"; {arm.describe()}; KD-tree build {arm.build_s:.1f} s not timed")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "insertions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": make_config(args.gpus),
        "cpu_baseline": {"value": val, "unit": "insertions/s", "cores": cores, "kind": arm.kind, "sample": sample},
        "e2e": {"value": val, "unit": "insertions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------------
# secondary measurements (N = 1)
# --------------------------------------------------------------------------------------------------------
C_SIZEOF_RESULT = 120  # sizeof(qd_add_result)
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures
NCU_TRAFFIC = {"insert_kernel_grid_2M": 315_400_704 + 206_669_312}  # dram read + write, profiles/r02j_grid_insert_2M_ncu_full.txt


def grid_insert_roofline(torch, dev, peak_hbm, Bg=2_000_000, n=100, iters=12):
    """GridArchive insertion against the HBM roofline (north_star: >= 60 % of HBM peak at 1 GPU): CMA-MAE
    lr=0.01, 500x500 cells (the grid of configs[2]), n=100, 2M-row batches of uniform measures -- the size at
    which the insertion is throughput- rather than latency-bound. Algorithmic bytes per SURVEY 8(d):
    45 B/row + 1 648 B per winner."""
    from pyribs_b200.archives import GridArchive

    a = GridArchive(solution_dim=n, dims=(500, 500), ranges=[(-1, 1)] * 2, learning_rate=0.01, threshold_min=0.0)
    g = torch.Generator(device=dev).manual_seed(1)
    bs = [dict(solution=torch.randn(Bg, n, dtype=torch.float64, device=dev, generator=g),
               objective=torch.rand(Bg, dtype=torch.float64, device=dev, generator=g) * 100 + 5.0 * j,
               measures=torch.rand(Bg, 2, dtype=torch.float64, device=dev, generator=g) * 2 - 1) for j in range(3)]
    for j in range(3):
        a.add(**bs[j])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for j in range(iters):
        a.add(**bs[j % 3])  # 3 x 1.65 GB of inputs (n = 100): nothing survives in the 126 MB L2 between calls
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    st = a.add(**bs[0])["status"]
    W = int(torch.unique(a.index_of(bs[0]["measures"])[st != 0]).numel())
    alg = Bg * 45 + W * (2 * n * 8 + 2 * 2 * 8 + 2 * 8)
    return {"kernel": "insert_kernel<double, double, 2, fused grid index> (GridArchive.add, one launch per add)",
            "workload": f"GridArchive 500x500, CMA-MAE lr=0.01, n={n}, {Bg // 1000}k-row batches, uniform measures",
            "bound": "hbm", "achieved": alg / ms / 1e6, "peak": peak_hbm, "unit": "GB/s", "frac": alg / ms / 1e6 / peak_hbm,
            "ms_per_add": ms, "insertions_per_s": Bg / ms * 1e3, "winners_per_add": W, "algorithmic_bytes": alg,
            "traffic": NCU_TRAFFIC.get("insert_kernel_grid_2M") if (Bg, n) == (2_000_000, 100) else None}


def es_lines(torch, dev, peak_hbm, cpu=True):
    """BASELINE.json's second metric (emitter ask+tell solutions/sec) on configs[3]: sep-CMA-ES and LM-MA-ES,
    solution_dim 10 000, batch 4096, mu = 2048. Device-resident (ask hands out the CUDA tensor, the ranking is a CUDA
    tensor), e2e (ask returns NumPy, the ranking comes from NumPy) and the reference's own strategy classes on the
    host cores, three generations each. Algorithmic bytes per SURVEY 8(d)."""
    from pyribs_b200.emitters.opt import _get_es

    n, lam, mu = 10_000, 4096, 2048
    out = []
    ribs, _ = load_reference() if cpu else (None, None)
    for es_name, kw, warm in (("sep_cma_es", {}, 4), ("lm_ma_es", {"n_vectors": 20}, 22), ("lm_ma_es", {}, 6)):
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
        # steady state: generations back to back, no synchronisation in between (the host enqueues generation g + 1
        # while the GPU runs generation g, as it does in a loop with a GPU evaluator); the ranking is a fixed device
        # permutation -- the evaluation is not part of this metric -- so the time is ask + tell and nothing else
        rank_fixed = torch.randperm(lam, device=dev)
        pe = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        # (with the default n_vectors = batch size the LM-MA ask grows by one direction vector per generation until
        # 4096 are in use: that line is quoted over generations warm + 10 .. warm + 30 and says so)
        K = 40 if (es_name == "sep_cma_es" or "n_vectors" in kw) else 10
        for rep in range(2):  # (the first pass warms the loop up)
            torch.cuda.synchronize()
            pe[0].record()
            for g in range(K):
                X = es.ask(return_device=True)
                es.tell(rank_fixed, None, mu)
            cur.wait_stream(es._es_stream())
            pe[1].record()
            torch.cuda.synchronize()
        steady_ms = pe[0].elapsed_time(pe[1]) / K
        # end to end: NumPy out of ask (D2H of the lam x n samples), NumPy ranking into tell
        e2e = []
        for g in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            Xh = es.ask()
            t1 = time.perf_counter()
            rank_h = np.argsort(((Xh - 1.0) ** 2).sum(1))
            t2 = time.perf_counter()
            es.tell(rank_h, None, mu)
            torch.cuda.synchronize()
            t3 = time.perf_counter()
            e2e.append((t1 - t0) + (t3 - t2))
        e2e_s = float(np.median(e2e[1:]))
        m = kw.get("n_vectors", lam if es_name == "lm_ma_es" else 0)
        m_eff = min(m, warm + 10) if m else 0
        ask_b = lam * n * 8 + 3 * n * 8 if es_name == "sep_cma_es" else 2 * lam * n * 8 + m_eff * n * 8
        tell_b = mu * n * 8 + 8 * n * 8 if es_name == "sep_cma_es" else 2 * mu * n * 8 + 2 * m * n * 8
        e = {"kernel": {"sep_cma_es": "sep_ask_rng_kernel (Philox + ziggurat fused with the transform) / "
                                      "wsum_partial_kernel + sep_tell_*",
                        "lm_ma_es": "fill_normal_zig + tall_dots (fp64 mma) + lmma_coef + lmma_out (fp64 mma) / "
                                    "wsum_partial_kernel x2 + lmma_tell_*"}[es_name] + ", csrc/es_stream.cu, es.cu",
             "workload": f"{es_name} n={n} lambda={lam} mu={mu}" + (f" n_vectors={m}" if m else ""),
             "metric": "emitter ask+tell solutions/sec", "solutions_per_s": lam / (steady_ms * 1e-3),
             "ms_per_generation": steady_ms,
             "timing": f"solutions_per_s / ms_per_generation / achieved / frac: {K} generations back to back (ask + tell, "
                       "fixed device ranking, no synchronisation between generations); ask_ms / tell_ms: one call from "
                       "an idle GPU, host launch latency included (solutions_per_s_single_calls is their sum's rate)",
             "solutions_per_s_single_calls": lam / ((a + t) * 1e-3),
             "bound": "hbm", "achieved": (ask_b + tell_b) / (steady_ms * 1e-3) / 1e9, "peak": peak_hbm, "unit": "GB/s",
             "frac": (ask_b + tell_b) / (steady_ms * 1e-3) / 1e9 / peak_hbm,
             "ask_ms": a, "tell_ms": t, "ask_GBps": ask_b / a / 1e6, "tell_GBps": tell_b / t / 1e6,
             "algorithmic_bytes": ask_b + tell_b,
             "e2e": {"value": lam / e2e_s, "unit": "solutions/s", "d2h_bytes_per_step": lam * n * 8,
                     "h2d_bytes_per_step": lam * 8, "ms_per_generation": e2e_s * 1e3}}
        if es_name == "lm_ma_es":
            e["ask_fp64_TFLOPs"] = 4.0 * lam * n * m_eff / (a * 1e-3) / 1e12
        del es, X
        torch.cuda.empty_cache()
        if ribs is not None:
            try:
                from ribs.emitters.opt import _get_es as ref_get_es

                ref = ref_get_es(es_name, sigma0=0.5, solution_dim=n, batch_size=lam, seed=1, dtype=np.float64,
                                 lower_bounds=-np.inf, upper_bounds=np.inf, **kw)
                ref.reset(np.zeros(n))
                ts = []
                for g in range(3):
                    t0 = time.perf_counter()
                    Xr = ref.ask()
                    t1 = time.perf_counter()
                    fit = -((Xr - 1.0) ** 2).sum(1)
                    order = np.argsort(-fit)
                    t2 = time.perf_counter()
                    ref.tell(order, fit[order], mu)
                    t3 = time.perf_counter()
                    ts.append((t1 - t0) + (t3 - t2))
                cpu_s = float(np.median(ts))
                e["cpu_baseline"] = {"value": lam / cpu_s, "unit": "solutions/s", "cores": os.cpu_count() or 1,
                                     "kind": "reference",
                                     "sample": f"ribs.emitters.opt {es_name} ask+tell, 3 generations, median "
                                               f"{cpu_s:.2f} s per generation (the reference pins BLAS to 1 thread "
                                               "inside ask/tell)"}
            except Exception as ex:  # noqa: BLE001
                e["cpu_baseline"] = {"error": f"{type(ex).__name__}: {ex}"}
        out.append(e)
    return out


def c2_line(torch, dev, steps, timed_fn_factory):
    """BASELINE.json configs[1] (last round's headline): CVTArchive 10k centroids, 2-D, 100k-row adds."""
    from pyribs_b200.archives import CVTArchive

    B, C, d, n = 100_000, 10_000, 2, 10
    cen = np.random.default_rng(0).uniform(-1, 1, (C, d))
    rng = np.random.default_rng(42)
    nb = 16
    hb = [dict(solution=rng.uniform(-1, 1, (B, n)), objective=rng.standard_normal(B),
               measures=rng.uniform(-1, 1, (B, d))) for _ in range(nb)]
    a = CVTArchive(solution_dim=n, centroids=cen, ranges=[(-1, 1)] * d)
    db = [{k: torch.from_numpy(v).to(dev) for k, v in b.items()} for b in hb]
    pb = [{k: torch.from_numpy(v).pin_memory().numpy() for k, v in b.items()} for b in hb]
    timed = timed_fn_factory()
    ms_dev, _ = timed(lambda j: a.add(db[j % nb]["solution"], db[j % nb]["objective"] + 0.05 * j,
                                      db[j % nb]["measures"]), steps, 5)
    a.clear()
    ms_pin, _ = timed(lambda j: a.add(pb[j % nb]["solution"], pb[j % nb]["objective"], pb[j % nb]["measures"]),
                      steps, 5)
    a.clear()
    ms_page, _ = timed(lambda j: a.add(hb[j % nb]["solution"], hb[j % nb]["objective"], hb[j % nb]["measures"]),
                       steps, 5)
    out = {"workload": "configs[1]: CVTArchive 10k centroids, 2-D, batch add of 100k (bucket-grid search + insert)",
           "insertions_per_s": B * steps / (ms_dev * 1e-3), "ms_per_add": ms_dev / steps,
           "e2e_pinned_insertions_per_s": B * steps / (ms_pin * 1e-3), "e2e_pinned_ms_per_add": ms_pin / steps,
           "e2e_pageable_insertions_per_s": B * steps / (ms_page * 1e-3), "e2e_pageable_ms_per_add": ms_page / steps}
    try:
        arm = CpuArm(cen)
        arm.add(hb[0], rows=2000)
        arm.archive.clear()
        t0 = time.perf_counter()
        k = 30
        for j in range(k):
            arm.add(hb[j % nb], 0.05 * j)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": k * B / dt, "unit": "insertions/s", "cores": os.cpu_count() or 1,
                               "kind": arm.kind, "sample": f"{k} adds of 100k rows; {arm.describe()}"}
    except Exception as ex:  # noqa: BLE001
        out["cpu_baseline"] = {"error": f"{type(ex).__name__}: {ex}"}
    return out


def published_protocol(torch, dev):
    """The reference's only published benchmark (benchmarks/cvt_add.py:79-86): 1 000 adds of 100 2-D rows into
    CVTArchives of 10 ... 1 000 cells; seconds for the 100k insertions. Both arms get the same centroid arrays
    (uniform random: the script's k-means centroids need scikit-learn's RNG) and NumPy inputs."""
    from pyribs_b200.archives import CVTArchive

    rng = np.random.default_rng(7)
    nb, bs = 1000, 100
    sol = rng.uniform(-1, 1, (nb, bs, 10))
    obj = rng.standard_normal((nb, bs))
    meas = rng.uniform(-1, 1, (nb, bs, 2))
    ribs, _ = load_reference()
    rows = []
    for cells in (10, 50, 100, 500, 1000):
        cen = np.random.default_rng(cells).uniform(-1, 1, (cells, 2))
        a = CVTArchive(solution_dim=10, centroids=cen, ranges=[(-1, 1)] * 2)
        for j in range(20):
            a.add(sol[j], obj[j], meas[j])
        a.clear()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for j in range(nb):
            a.add(sol[j], obj[j], meas[j])
        torch.cuda.synchronize()
        t_gpu = time.perf_counter() - t0
        row = {"cells": cells, "gpu_s": t_gpu}
        if ribs is not None:
            from ribs.archives import CVTArchive as RefCVT

            for nn in ("scipy_kd_tree", "brute_force"):
                r = RefCVT(solution_dim=10, centroids=cen, ranges=[(-1, 1)] * 2, nearest_neighbors=nn)
                for j in range(5):
                    r.add(sol[j], obj[j], meas[j])
                r.clear()
                t0 = time.perf_counter()
                for j in range(nb):
                    r.add(sol[j], obj[j], meas[j])
                row["ref_" + nn + "_s"] = time.perf_counter() - t0
        rows.append(row)
    return {"workload": "benchmarks/cvt_add.py protocol: 1 000 adds x 100 rows (2-D), NumPy in / NumPy out, seconds "
                        "per 100k insertions", "rows": rows}


# --------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--search", default="auto", choices=["auto", "tensor", "scan", "tree"])
    ap.add_argument("--no-secondary", action="store_true", help="skip configs[1], the published protocol, the "
                    "GridArchive roofline and the ES lines (N = 1 only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the reference arm inside this run")
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

    B, C, d, n = WORKLOAD["batch"], WORKLOAD["centroids"], WORKLOAD["measure_dim"], WORKLOAD["solution_dim"]
    assert B % world == 0
    Bl = B // world
    lo, hi = rank * Bl, (rank + 1) * Bl
    cen = make_centroids()
    batches = make_batches(N_RESIDENT)              # the global batches (every rank derives the same ones)
    local = [{k: np.ascontiguousarray(v[lo:hi]) for k, v in b.items()} for b in batches]
    if rank != 0 or world > 1:
        full0 = batches[0] if rank == 0 else None
        batches = None
    t0 = time.perf_counter()
    archive = CVTArchive(solution_dim=n, centroids=cen, ranges=[(-1, 1)] * d, search_method=args.search,
                         data_parallel=world > 1)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    dev_batches = [{k: torch.from_numpy(v).to(dev) for k, v in b.items()} for b in local]
    pinned = [{k: torch.from_numpy(v).pin_memory() for k, v in b.items()} for b in local]
    host_batches = [{k: v.numpy() for k, v in b.items()} for b in pinned]

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

    # ---- in-run parity check: cell indices against SciPy's KD-tree (what the reference calls) on a row sample --
    from scipy.spatial import cKDTree

    ns = max(20_000 // world, 2048)
    sub = np.random.default_rng(100 + rank).choice(Bl, ns, replace=False)
    kd_idx = cKDTree(cen).query(local[0]["measures"][sub], workers=-1)[1]
    got = archive.index_of(local[0]["measures"][sub])
    mism = int(np.count_nonzero(got != kd_idx))
    if world > 1:
        t = torch.tensor([mism], device=dev, dtype=torch.int64)
        dist.all_reduce(t)
        mism = int(t.item())
    if mism:
        raise SystemExit(f"parity check failed: {mism} of {ns * world} sampled cell indices differ from the KD-tree")

    # ---- device-resident throughput -------------------------------------------------------------------------
    first = []

    def step_dev(j):
        b = dev_batches[j % N_RESIDENT]
        out = archive.add(b["solution"], b["objective"] + OBJ_DRIFT * j, b["measures"])
        if j < 2:
            first.append({k: v.cpu().numpy() for k, v in out.items()})

    with ClockSampler(local_rank) as clk:
        ms_dev, launches = timed(step_dev, args.steps, args.warmup)
    clocks = clk.summary()

    # ---- end to end: host arrays in, NumPy status/value out ---------------------------------------------------
    archive.clear()

    def step_e2e(j):
        b = host_batches[j % N_RESIDENT]
        out = archive.add(b["solution"], b["objective"], b["measures"])
        assert out["status"].shape[0] == Bl

    e2e_steps = min(args.steps, 10)
    ms_e2e, _ = timed(step_e2e, e2e_steps, 3)
    archive.clear()

    def step_page(j):
        b = local[j % N_RESIDENT]
        archive.add(b["solution"], b["objective"], b["measures"])

    ms_page, _ = timed(step_page, e2e_steps, 2)
    h2d = sum(v.nbytes for v in local[0].values())
    d2h = Bl * (4 + 8) + C_SIZEOF_RESULT

    # ---- dominant kernel alone: nearest-centroid search ---------------------------------------------------------
    flags = torch.zeros(1, dtype=torch.int32, device=dev)

    def step_search(j):
        archive._index_of_device(dev_batches[j % N_RESIDENT]["measures"], flags.data_ptr())

    ms_k, k_launches = timed(step_search, args.steps, args.warmup)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_hbm = float(peaks.get("hbm_gbs", 6650.0))
    peak_tf = float(peaks.get("bf16_tflops", 1590.0))
    s_k = ms_k / args.steps * 1e-3
    method = {0: "tensor", 1: "scan", 2: "bucket", 3: "tree"}[archive._pick_method(Bl, C)]
    alg_bytes = (Bl + C) * d * 8 + 4 * Bl
    roofline = {
        "kernel": {"tree": "cvt_tree_kernel<double, 8> (bounding-box tree in KD order, one warp per query, exact fp64 "
                           "leaves; csrc/cvt_tree.cu)",
                   "tensor": "cvt_tc_kernel (tcgen05 filter) + rescore_kernel (csrc/cvt_tc.cu)",
                   "scan": "cvt_scan_fp64_kernel", "bucket": "cvt_bucket_kernel"}[method],
        "bound": "hbm", "achieved": alg_bytes / s_k / 1e9, "peak": peak_hbm, "unit": "GB/s",
        "frac": alg_bytes / s_k / 1e9 / peak_hbm, "traffic": None,
        "algorithmic_bytes": alg_bytes, "ms_per_search": ms_k / args.steps,
        "launches_per_search": k_launches / args.steps, "queries_per_s": Bl / s_k,
        "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650",
        "note": "algorithmic bytes of SURVEY 8(d) for index_of, (B + C) d s + 4 B, over the measured kernel time. The "
                "search is neither HBM- nor tensor-bound, so the frac against HBM is small by construction: a query walks "
                "~16 nodes and ~44 leaves of an L2-resident 104 MB tree in ~5 300 warp instructions; the sorted-walk "
                "kernel issues on 76 % of its cycles at 50 % occupancy (ncu: profiles/r02j_cvt_tree_c5_ncu_full.txt, "
                "DESIGN.md 4.3b) -- it is instruction-issue-bound. The exhaustive tcgen05 contraction it replaces on "
                "this shape ran at 164 TFLOP/s = 10 % of the bf16 peak and took 13x longer (profiles/r02_cvt_tc_*). The "
                "bandwidth-regime roofline of this line is the GridArchive insertion under `secondary`.",
        "exhaustive_equivalent_TFLOPs": 2.0 * Bl * C * d / s_k / 1e12, "bf16_peak_TFLOPs": peak_tf,
    }

    alt = None
    if world > 1:
        # ---- the centroid-sharded search of north_star beside it: local argmin + two all-reduces ------------
        try:
            g_meas = [torch.from_numpy(np.ascontiguousarray(b_["measures"])).to(dev) for b_ in
                      make_batches(1)] if False else None
            sh = CVTArchive(solution_dim=n, centroids=cen, ranges=[(-1, 1)] * d, search_method=args.search,
                            shard_centroids=True, data_parallel=True)

            def step_sh(j):
                b = dev_batches[j % N_RESIDENT]
                sh.add(b["solution"], b["objective"] + OBJ_DRIFT * j, b["measures"])

            ms_sh, _ = timed(step_sh, min(args.steps, 10), 3)
            gm = sh._all_gather_many([dev_batches[0]["measures"]])[0]
            ms_loc, _ = timed(lambda j: sh._local_search(gm, flags.data_ptr(), True), 10, 3)
            ms_all, _ = timed(lambda j: sh._index_of_device(gm, flags.data_ptr()), 10, 3)
            alt = {"parallelism": f"centroids sharded over {world} GPUs (C/{world} per rank): every rank searches all "
                                  "1M rows in its shard, all-reduce(MIN) of the exact distances, mask, "
                                  "all-reduce(MIN) of the indices; replicated insert",
                   "insertions_per_s": B * min(args.steps, 10) / (ms_sh * 1e-3), "ms_per_add": ms_sh / min(args.steps, 10),
                   "local_search_ms": ms_loc / 10, "search_plus_allreduces_ms": ms_all / 10,
                   "allreduce_share_of_add": max(0.0, (ms_all - ms_loc) / 10) / (ms_sh / min(args.steps, 10))}
            del sh, gm
        except Exception as e:  # noqa: BLE001
            alt = {"error": f"{type(e).__name__}: {e}"}

    if world > 1:
        # ---- owner-computes store beside the replicated one: cells sharded by contiguous range, every rank commits
        # only the rows of its own cells (DeviceArchive._add_owned) -----------------------------
        try:
            own = CVTArchive(solution_dim=n, centroids=cen, ranges=[(-1, 1)] * d, search_method=args.search,
                             data_parallel=True, shard_store=True)

            first_own = []

            def step_own(j):
                b = dev_batches[j % N_RESIDENT]
                out = own.add(b["solution"], b["objective"] + OBJ_DRIFT * j, b["measures"])
                if j < 2:
                    first_own.append({k: v.cpu().numpy() for k, v in out.items()})

            ms_own, launches_own = timed(step_own, args.steps, args.warmup)
            # the two stores must tell every row the same thing (first two adds into the empty archives)
            mism = sum(int(np.count_nonzero(x["status"] != y["status"])) + int(np.count_nonzero(x["value"] != y["value"]))
                       for x, y in zip(first, first_own))
            t = torch.tensor([mism], device=dev, dtype=torch.int64)
            dist.all_reduce(t)
            if int(t.item()):
                raise SystemExit(f"parity check failed: owner-computes and replicated store disagree on {int(t.item())} "
                                 "status / value entries")
            n_own = len(own)
            own.clear()

            def step_own_e2e(j):
                b = host_batches[j % N_RESIDENT]
                own.add(b["solution"], b["objective"], b["measures"])

            ms_own_e2e, _ = timed(step_own_e2e, e2e_steps, 3)
            alt["owner_computes_store"] = {
                "parallelism": f"as the headline, but the STORE is sharded by cell range over the {world} ranks: every "
                               "rank classifies / commits only the rows of its own cells and pulls only its own "
                               "winners' rows; status / value return by a reverse push over NVLink",
                "insertions_per_s": B * args.steps / (ms_own * 1e-3), "ms_per_add": ms_own / args.steps,
                "e2e_insertions_per_s": B * e2e_steps / (ms_own_e2e * 1e-3), "e2e_ms_per_add": ms_own_e2e / e2e_steps,
                "elites_after_timed_adds": n_own, "gpu_launches": int(launches_own),
                "parity": "status / value of the first two adds bit-equal to the replicated store's on every rank"}
            del own
        except Exception as e:  # noqa: BLE001
            alt["owner_computes_store"] = {"error": f"{type(e).__name__}: {e}"}

    secondary = []
    cpu = None
    if world == 1 and rank == 0:
        if not args.no_cpu:
            # ---- the reference on the host cores: same centroids, same first batches, full 1M-row adds ---------
            arm = CpuArm(cen)
            arm.add(batches[0], rows=2000)
            arm.archive.clear()
            ts, outs = [], []
            for j in range(3):
                t0 = time.perf_counter()
                outs.append(arm.add(batches[j % N_RESIDENT], OBJ_DRIFT * j))
                ts.append(time.perf_counter() - t0)
            cpu_s = float(np.median(ts))
            for j in range(2):  # status bit-exact, value within 1e-6 relative (north_star)
                if not np.array_equal(outs[j]["status"], first[j]["status"]):
                    raise SystemExit(f"parity check failed: add {j} statuses differ from the reference")
                if not np.allclose(outs[j]["value"], first[j]["value"], rtol=1e-6, atol=0):
                    raise SystemExit(f"parity check failed: add {j} values differ from the reference")
            cpu = {"value": B / cpu_s, "unit": "insertions/s", "cores": os.cpu_count() or 1, "kind": arm.kind,
                   "sample": f"3 full 1M-row adds (median {cpu_s:.2f} s; first, into the empty archive, {ts[0]:.2f} s); "
                             f"{arm.describe()}; KD-tree build {arm.build_s:.1f} s not timed; status / value of the "
                             "first two adds equal the GPU's (status bit-exact, value 1e-6)"}
            del arm
        if not args.no_secondary:
            del dev_batches, pinned, host_batches
            torch.cuda.empty_cache()

            def factory():
                return timed

            for name, fn in (("configs[1]", lambda: c2_line(torch, dev, 200, factory)),
                             ("published protocol", lambda: published_protocol(torch, dev)),
                             ("grid insert", lambda: grid_insert_roofline(torch, dev, peak_hbm)),
                             # the same archive with 10x longer solution rows: the row copy (HBM-bound) is 97 % of the
                             # bytes instead of 81 %, i.e. how the kernel does where the roofline is the bound
                             ("grid insert, n=1000", lambda: grid_insert_roofline(torch, dev, peak_hbm, 1_000_000, 1000, 8))):
                try:
                    secondary.append(fn())
                    torch.cuda.empty_cache()
                except Exception as e:  # noqa: BLE001
                    secondary.append({"kernel": name, "error": f"{type(e).__name__}: {e}"})
            try:
                secondary.extend(es_lines(torch, dev, peak_hbm, cpu=not args.no_cpu))
            except Exception as e:  # noqa: BLE001
                secondary.append({"kernel": "evolution strategies (configs[3])", "error": f"{type(e).__name__}: {e}"})
    roofline["secondary"] = secondary

    line = {
        "metric": METRIC,
        "value": B * args.steps / (ms_dev * 1e-3),
        "unit": "insertions/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps,
        "higher_is_better": True,
        "scaling": "strong",  # the global batch (1M rows) is fixed; --gpus N splits it
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": make_config(world),
        "search": method,
        "index_build_s": build_s,
        "parity": f"{ns * world} sampled cell indices equal SciPy's KD-tree"
                  + ("; status / value of two adds equal the reference's" if cpu else ""),
        "clocks": clocks,
        "e2e": {"value": B * e2e_steps / (ms_e2e * 1e-3), "unit": "insertions/s",
                "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": d2h * world,
                "ms_per_step": ms_e2e / e2e_steps, "inputs": "page-locked NumPy arrays",
                "pageable": {"value": B * e2e_steps / (ms_page * 1e-3), "ms_per_step": ms_page / e2e_steps,
                             "inputs": "ordinary (pageable) NumPy arrays, staged through the archive's pinned "
                                       "buffers"}},
        "gpu_launches": int(launches),
        "roofline": roofline,
    }
    if alt is not None:
        own = alt.get("owner_computes_store") or {}
        if "insertions_per_s" in own:
            # N > 1: the owner-computes store is the configuration of record (per-rank insert work and row traffic 1/N);
            # the replicated store measured in the same run moves to `alt`
            alt["replicated_store"] = {
                "parallelism": REPLICATED_TEXT.format(world=world), "insertions_per_s": line["value"],
                "ms_per_add": line["ms_per_step"], "e2e_insertions_per_s": line["e2e"]["value"],
                "e2e_ms_per_add": line["e2e"]["ms_per_step"], "e2e_pageable": line["e2e"]["pageable"],
                "gpu_launches": line["gpu_launches"]}
            line["value"], line["ms_per_step"] = own["insertions_per_s"], own["ms_per_add"]
            line["gpu_launches"] = own["gpu_launches"]
            line["e2e"] = {"value": own["e2e_insertions_per_s"], "unit": "insertions/s",
                           "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": d2h * world,
                           "ms_per_step": own["e2e_ms_per_add"], "inputs": "page-locked NumPy arrays"}
            line["parity"] += "; owner-computes store bit-equal to the replicated store on the first two adds"
            del alt["owner_computes_store"]
        else:  # the owner-computes store could not be timed: the replicated store's numbers stand
            line["config"]["parallelism"] = REPLICATED_TEXT.format(world=world)
        line["alt"] = alt
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
