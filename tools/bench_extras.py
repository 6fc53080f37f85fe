"""Secondary measurements (not the driver's bench line): GridArchive insertion against the HBM
roofline (C3 and a large batch), ES ask/tell at C4, and the CVT search at C5 per-rank scale.
Prints one JSON object per line."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyribs_b200 import _lib  # noqa: E402
from pyribs_b200.archives import CVTArchive, GridArchive  # noqa: E402

PEAKS = {}
try:
    PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass
HBM = float(PEAKS.get("hbm_gbs", 6650.0))


def cuda_time(fn, iters, warmup=3):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def grid_case(name, B, n, dims, lr, tmin, clustered, nb):
    rng = np.random.default_rng(0)
    a = GridArchive(solution_dim=n, dims=dims, ranges=[(-1, 1)] * 2, learning_rate=lr, threshold_min=tmin)
    batches = []
    for j in range(nb):
        if clustered:
            E = 100
            mu = rng.normal(0, 0.05, (E, 1, 2))
            meas = np.clip(mu + rng.normal(0, 0.02, (E, B // E, 2)), -1, 1).reshape(-1, 2)
        else:
            meas = rng.uniform(-1, 1, (B, 2))
        batches.append(dict(solution=torch.randn(B, n, dtype=torch.float64, device="cuda"),
                            objective=torch.from_numpy(rng.uniform(0, 100, B) + 5.0 * j).cuda(),
                            measures=torch.from_numpy(meas).cuda()))
    winners = []

    def step(i):
        b = batches[i % nb]
        before = a._store._updates[0]
        a.add(b["solution"], b["objective"], b["measures"])

    ms = cuda_time(step, nb * 2, warmup=nb)
    # algorithmic bytes: 45 B/row + (2*n*8 + 2*2*8 + 2*8) B per winner; winners of one
    # steady-state call counted on the host from a replay
    b = batches[0]
    idx = a.index_of(b["measures"]).cpu().numpy()
    st = a.add(b["solution"], b["objective"], b["measures"])["status"].cpu().numpy()
    W = len(np.unique(idx[st != 0]))
    alg = B * 45 + W * (2 * n * 8 + 2 * 2 * 8 + 2 * 8)
    out = {"case": name, "B": B, "n": n, "dims": list(dims), "mode": "cma_mae" if lr is not None else "cma_me",
           "ms_per_add": ms, "insertions_per_s": B / ms * 1e3, "winners_per_add": W, "algorithmic_bytes": alg,
           "achieved_GBps": alg / ms / 1e6, "hbm_peak_GBps": HBM, "frac_of_hbm": alg / ms / 1e6 / HBM,
           "note": "whole add() incl. host sync; device tensors in"}
    print(json.dumps(out), flush=True)


def es_case(name, es_name, n, lam, gens, **kw):
    from pyribs_b200.emitters.opt import _get_es

    es = _get_es(es_name, sigma0=0.5, solution_dim=n, batch_size=lam, seed=1, dtype=np.float64,
                 lower_bounds=np.full(n, -np.inf), upper_bounds=np.full(n, np.inf), **kw)
    es.reset(np.zeros(n))
    t_ask, t_tell = [], []
    for g in range(gens):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        X = es.ask(return_device=True)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        fit = -(X - 1.0).pow(2).sum(1)
        rank = torch.argsort(fit, descending=True)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        es.tell(rank, None, lam // 2)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        t_ask.append(t1 - t0)
        t_tell.append(t3 - t2)
    a, t = np.median(t_ask[3:]), np.median(t_tell[3:])
    s = 8
    print(json.dumps({"case": name, "es": es_name, "n": n, "lambda": lam, "ask_ms": a * 1e3, "tell_ms": t * 1e3,
                      "ask_ms_last": t_ask[-1] * 1e3, "solutions_per_s": lam / (a + t),
                      "ask_GBps_alg": lam * n * s / a / 1e9, "tell_GBps_alg": (lam // 2) * n * s / t / 1e9,
                      "note": "device-resident (ask returns the CUDA tensor), on-device Philox normals"}), flush=True)


def cvt_case(name, B, C, d, method, iters=3):
    rng = np.random.default_rng(0)
    cen = rng.uniform(-1, 1, (C, d))
    a = CVTArchive(solution_dim=1, centroids=cen, ranges=[(-1, 1)] * d, search_method=method)
    m = torch.from_numpy(rng.uniform(-1, 1, (B, d))).cuda()
    flags = torch.zeros(1, dtype=torch.int32, device="cuda")
    ms = cuda_time(lambda i: a._index_of_device(m, flags.data_ptr()), iters, warmup=1)
    print(json.dumps({"case": name, "B": B, "C": C, "d": d, "method": method, "ms_per_search": ms,
                      "pairs_per_s": B * C / ms * 1e3, "queries_per_s": B / ms * 1e3,
                      "TFLOPs_alg": 2.0 * B * C * d / ms / 1e9, "flags": int(flags.item())}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="grid,es,cvt")
    args = ap.parse_args()
    what = args.what.split(",")
    if "grid" in what:
        grid_case("C3 CMA-MAE 500x500 B=51200 n=100", 51200, 100, (500, 500), 0.01, 0.0, True, 6)
        grid_case("large CMA-ME 500x500 B=2M n=100 uniform", 2_000_000, 100, (500, 500), None, -np.inf, False, 3)
        grid_case("large CMA-MAE 500x500 B=2M n=100 uniform", 2_000_000, 100, (500, 500), 0.01, 0.0, False, 3)
    if "gridlarge" in what:
        grid_case("large CMA-MAE 500x500 B=2M n=100 uniform", 2_000_000, 100, (500, 500), 0.01, 0.0, False, 3)
    if "es" in what:
        es_case("C4 sep-CMA n=10000 lam=4096", "sep_cma_es", 10_000, 4096, 12)
        es_case("C4 LM-MA n=10000 lam=4096 n_vectors=20", "lm_ma_es", 10_000, 4096, 30, n_vectors=20)
        es_case("C4 LM-MA n=10000 lam=4096 default vectors", "lm_ma_es", 10_000, 4096, 12)
        es_case("C1 CMA-ES n=100 lam=36", "cma_es", 100, 36, 30)
        es_case("CMA-ES n=1000 lam=36", "cma_es", 1000, 36, 12)
    if "tree" in what:
        cvt_case("C5 full (1M x 1M, 8-D) box tree", 1_000_000, 1_000_000, 8, "tree", 3)
        cvt_case("C5 per-rank shard (1M x 125k, 8-D) box tree", 1_000_000, 125_000, 8, "tree", 3)
        cvt_case("C5 batch shard (125k x 1M, 8-D) box tree", 125_000, 1_000_000, 8, "tree", 3)
        cvt_case("1M x 1M, 4-D box tree", 1_000_000, 1_000_000, 4, "tree", 3)
        cvt_case("200k x 1M, 12-D box tree", 200_000, 1_000_000, 12, "tree", 2)
        cvt_case("100k x 200k, 16-D box tree", 100_000, 200_000, 16, "tree", 2)
        cvt_case("100k x 200k, 16-D tensor", 100_000, 200_000, 16, "tensor", 2)
        cvt_case("51200 x 1M, 8-D box tree", 51_200, 1_000_000, 8, "tree", 3)
    if "tree5" in what:
        cvt_case("C5 full (1M x 1M, 8-D) box tree", 1_000_000, 1_000_000, 8, "tree", 3)
    if "cvt5" in what:
        cvt_case("C5 per-rank shard (1M x 125k, 8-D)", 1_000_000, 125_000, 8, "tensor", 2)
    if "cvt" in what:
        cvt_case("C2 search", 100_000, 10_000, 2, "tensor", 10)
        cvt_case("C5 per-rank shard (1M x 125k, 8-D)", 1_000_000, 125_000, 8, "tensor", 2)
        cvt_case("C5 fp64 scan sample (100k x 125k, 8-D)", 100_000, 125_000, 8, "scan", 1)
