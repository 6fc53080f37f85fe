"""Times the ziggurat generator kernels alone (CUDA events, 20 launches after 5 warm-ups) at configs[3] scale:
qd_fill_normal_zig (41 M normals), qd_sep_ask_rng (4096 x 10 000, with and without z_out / row_ids), and the
strategies' own ask and tell. Usage: zig_probe.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pyribs_b200 import _lib


def timed(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3  # us


def timed_wall(fn, iters=20, warm=5):
    import time

    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e6  # us


def main():
    L, s = _lib.lib(), torch.cuda.current_stream().cuda_stream
    lam, n = 4096, 10_000
    dev = dict(dtype=torch.float64, device="cuda")
    out = torch.empty(lam * n, **dev)
    mean, C = torch.zeros(n, **dev), torch.ones(n, **dev)
    status = torch.zeros(8, **dev)
    status[0] = 0.5
    X, z = torch.empty(lam, n, **dev), torch.empty(lam, n, **dev)
    ids = torch.randperm(lam, device="cuda").int()
    res = {}
    res["fill_us"] = timed(lambda: _lib.check(L.qd_fill_normal_zig(out.data_ptr(), lam * n, 7, 0, s)))
    res["sep_ask_us"] = timed(lambda: _lib.check(L.qd_sep_ask_rng(
        lam, n, None, mean.data_ptr(), C.data_ptr(), status.data_ptr(), 7, 0, X.data_ptr(), None, s)))
    res["sep_ask_ids_us"] = timed(lambda: _lib.check(L.qd_sep_ask_rng(
        lam, n, ids.data_ptr(), mean.data_ptr(), C.data_ptr(), status.data_ptr(), 7, 0, X.data_ptr(), None, s)))
    res["sep_ask_zout_us"] = timed(lambda: _lib.check(L.qd_sep_ask_rng(
        lam, n, None, mean.data_ptr(), C.data_ptr(), status.data_ptr(), 7, 0, X.data_ptr(), z.data_ptr(), s)))
    nbytes = lam * n * 8
    res["fill_TBps"] = nbytes / res["fill_us"] / 1e6
    res["sep_ask_TBps"] = nbytes / res["sep_ask_us"] / 1e6
    # the generator's law, at the scale this launch gives for free
    x = out[: 8_000_000].cpu().numpy()
    res["mean"], res["var"] = float(x.mean()), float(x.var())
    res["tail_beyond_R"] = float(np.mean(np.abs(out.cpu().numpy()) > 4.385945034871305)), 1.1548328e-05  # got, expected
    print(json.dumps(res), flush=True)
    from pyribs_b200.emitters.opt import _get_es

    for name, kw in (("sep_cma_es", {}), ("lm_ma_es", {"n_vectors": 20}), ("lm_ma_es", {})):
        es = _get_es(name, sigma0=0.5, solution_dim=n, batch_size=lam, seed=1, dtype=np.float64,
                     lower_bounds=np.full(n, -np.inf), upper_bounds=np.full(n, np.inf), **kw)
        es.reset(np.zeros(n))
        rank = torch.randperm(lam, device="cuda")
        for _ in range(24):  # LM-MA uses min(generations, n_vectors) direction vectors: let all 20 come into play
            es.ask(return_device=True)
            es.tell(rank, None, lam // 2)
        ask = timed(lambda: es.ask(return_device=True), iters=10, warm=2)
        tell = timed_wall(lambda: es.tell(rank, None, lam // 2), iters=10, warm=2)  # (tell runs on the strategy's stream)
        print(json.dumps({"es": name, **kw, "ask_us": ask, "tell_us": tell}), flush=True)


if __name__ == "__main__":
    main()
