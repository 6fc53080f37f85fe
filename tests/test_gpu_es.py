"""GPU parity tests of the evolution strategies against the reference's golden traces.

Unit normals are injected (the reference's PCG64 stream cannot be replayed on a GPU); for
CMA-ES the eigensystem of the CPU oracle is injected whenever the oracle refreshes it, because
eigenvector signs / degenerate-subspace bases differ between LAPACK and cuSOLVER.
Tolerance: 1e-6 relative (north_star), observed ~1e-12."""
import os
import sys

import numpy as np
import pytest

from golden_util import assert_close, gens_of, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import qd_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu
RTOL = 1e-6


def _make(es_name, cfg):
    from pyribs_b200.emitters.opt import _get_es

    kw = {}
    if es_name == "lm_ma_es" and int(cfg["n_vectors"]) > 0:
        kw["n_vectors"] = int(cfg["n_vectors"])
    n, lam = int(cfg["n"]), int(cfg["lam"])
    gpu = _get_es(es_name, sigma0=float(cfg["sigma0"]), solution_dim=n, batch_size=lam, seed=1, dtype=np.float64,
                  lower_bounds=np.full(n, -np.inf), upper_bounds=np.full(n, np.inf), **kw)
    ora = O.ORACLE_ES[es_name](float(cfg["sigma0"]), n, lam, seed=1, **kw)
    gpu.reset(cfg["x0"])
    ora.reset(cfg["x0"])
    return gpu, ora


@pytest.mark.parametrize("name", sorted(load_golden("es_traces").keys()))
def test_es_trace(name):
    case = load_golden("es_traces")[name]
    cfg = case["cfg"]
    es_name = name.split("_n")[0]
    gpu, ora = _make(es_name, cfg)
    n, lam = int(cfg["n"]), int(cfg["lam"])
    for gi, g in enumerate(gens_of(case)):
        z = np.random.default_rng(int(g["zseed"])).standard_normal((lam, n))
        before = getattr(ora, "updated_eval", None)
        sol_o = ora.ask(z)
        if es_name == "cma_es" and ora.updated_eval != before:
            gpu.set_eigensystem(ora.eigvals, ora.B)
        sol_g = gpu.ask(z=z)
        assert not sol_g.flags.writeable
        assert_close(sol_g, sol_o, 1e-9, f"{name} g{gi} solutions vs oracle")
        if "solutions" in g:
            assert_close(sol_g, g["solutions"], RTOL, f"{name} g{gi} solutions vs golden")
        ora.tell(g["ranking"], int(g["mu"]))
        gpu.tell(g["ranking"], None, int(g["mu"]))
        fit = -np.sum((sol_o - 1.0) ** 2, axis=1)
        assert int(gpu.check_stop(fit[g["ranking"]])) == int(g["stop"])
        assert_close(gpu.mean, g["mean"], RTOL, f"{name} g{gi} mean")
        assert_close(gpu.sigma, g["sigma"], RTOL, f"{name} g{gi} sigma")
        assert_close(gpu.ps, g["ps"], RTOL, f"{name} g{gi} ps")
        if "pc" in g:
            assert_close(gpu.pc, g["pc"], RTOL, f"{name} g{gi} pc")
        if "C" in g:
            assert_close(gpu.cov, g["C"], RTOL, f"{name} g{gi} C")
        if "m" in g:
            assert_close(gpu.m, g["m"], RTOL, f"{name} g{gi} m")
        if "invsqrt" in g and es_name == "cma_es":
            assert_close(gpu.invsqrt, g["invsqrt"], RTOL, f"{name} g{gi} invsqrt")


@pytest.mark.parametrize("es_name", ["sep_cma_es", "cma_es", "lm_ma_es"])
def test_es_stream_is_ordered_behind_the_callers_stream(es_name):
    """A strategy works on its own CUDA stream; its state is initialised, and CUDA-tensor rankings are produced, on
    the caller's. With the caller's stream kept busy (a 20 ms spin kernel) the strategy must still see both: the
    first generation equals the oracle's, and a tell fed a ranking computed behind a second spin moves the mean
    exactly as the same ranking handed over as a NumPy array does."""
    import torch

    from pyribs_b200.emitters.opt import _get_es

    n, lam = 12, 10
    x0 = np.linspace(-1, 1, n)
    rng = np.random.default_rng(5)
    z = rng.standard_normal((lam, n))
    ora = O.ORACLE_ES[es_name](0.7, n, lam, seed=1)
    ora.reset(x0)
    want = ora.ask(z)

    def build():
        torch.cuda._sleep(40_000_000)  # ~20 ms of the caller's stream ahead of the constructor's fills
        es = _get_es(es_name, sigma0=0.7, solution_dim=n, batch_size=lam, seed=1, dtype=np.float64,
                     lower_bounds=np.full(n, -np.inf), upper_bounds=np.full(n, np.inf))
        es.reset(x0)
        return es

    a, b = build(), build()
    assert_close(a.ask(z=z), want, 1e-9, "first generation behind a busy caller stream")
    assert_close(b.ask(z=z), want, 1e-9, "first generation behind a busy caller stream")
    fit = torch.from_numpy(-np.sum((want - 0.3) ** 2, axis=1)).cuda()
    torch.cuda._sleep(40_000_000)
    rank_d = torch.argsort(fit, descending=True)  # still queued behind the spin when tell() returns
    a.tell(rank_d, None, lam // 2)
    b.tell(rank_d.cpu().numpy(), None, lam // 2)
    np.testing.assert_array_equal(a.mean, b.mean)
    assert not np.array_equal(a.mean, x0)


def test_cma_own_eigensystem_is_consistent():
    """The cuSOLVER path (no injection): C = B L B^T, C^-1/2 matches NumPy on the same C."""
    from pyribs_b200.emitters.opt import CMAEvolutionStrategy

    rng = np.random.default_rng(0)
    n, lam = 60, 24
    es = CMAEvolutionStrategy(0.7, n, lam, seed=3)
    es.reset(rng.standard_normal(n))
    for g in range(12):
        sol = es.ask()
        fit = -np.sum(sol**2, axis=1)
        es.tell(np.flip(np.argsort(fit)), fit, lam // 2)
    es.update_eigensystem(force=True)
    Cm = es.cov
    np.testing.assert_array_equal(Cm, Cm.T)
    lam_np, B_np = np.linalg.eigh(Cm)
    assert_close(np.sort(es.eigenvalues), np.sort(np.abs(lam_np)), 1e-9, "eigenvalues")
    inv = (B_np * (1 / np.sqrt(np.abs(lam_np)))) @ B_np.T
    assert_close(es.invsqrt, np.maximum(inv, inv.T), 1e-8, "invsqrt")
    assert es.sigma > 0 and np.all(np.isfinite(es.mean))


def test_philox_normals():
    import torch

    from pyribs_b200 import _lib

    L = _lib.lib()
    out = torch.empty(2_000_001, dtype=torch.float64, device="cuda")
    _lib.check(L.qd_fill_normal(out.data_ptr(), out.numel(), 1234, 0, torch.cuda.current_stream().cuda_stream))
    x = out.cpu().numpy()
    assert abs(x.mean()) < 5e-3 and abs(x.var() - 1) < 5e-3
    assert abs(((x - x.mean()) ** 4).mean() / x.var() ** 2 - 3) < 3e-2
    out2 = torch.empty_like(out)
    _lib.check(L.qd_fill_normal(out2.data_ptr(), out2.numel(), 1234, 0, torch.cuda.current_stream().cuda_stream))
    assert torch.equal(out, out2)
    _lib.check(L.qd_fill_normal(out2.data_ptr(), out2.numel(), 1235, 0, torch.cuda.current_stream().cuda_stream))
    assert not torch.equal(out, out2)


def test_bounds_are_respected():
    from pyribs_b200.emitters.opt import SeparableCMAEvolutionStrategy

    n = 12
    es = SeparableCMAEvolutionStrategy(0.6, n, 64, seed=0, lower_bounds=np.full(n, -1.0), upper_bounds=np.full(n, 1.0))
    es.reset(np.zeros(n))
    sol = es.ask()
    assert sol.shape == (64, n) and np.all(sol >= -1.0) and np.all(sol <= 1.0)


@pytest.mark.parametrize("es", ["cma_es", "sep_cma_es", "lm_ma_es"])
def test_sphere_loop_improves(es):
    """End to end: Scheduler + 3 emitters + GridArchive on the sphere benchmark (the reference's
    own smoke test, tests/emitters/evolution_strategy_emitter_test.py:84-96, plus a progress check)."""
    from pyribs_b200.archives import GridArchive
    from pyribs_b200.emitters import EvolutionStrategyEmitter
    from pyribs_b200.schedulers import Scheduler

    n = 20
    archive = GridArchive(solution_dim=n, dims=(20, 20), ranges=[(-51.2, 51.2)] * 2, seed=1)
    emitters = [EvolutionStrategyEmitter(archive, x0=np.zeros(n), sigma0=0.5, ranker="2imp", es=es, batch_size=18,
                                         seed=s) for s in range(3)]
    sched = Scheduler(archive, emitters)
    qd = []
    for it in range(30):
        sol = sched.ask()
        assert sol.shape == (54, n)
        obj = -np.sum((sol - 2.048) ** 2, axis=1)
        clipped = np.clip(sol, -5.12, 5.12)
        meas = np.stack((clipped[:, : n // 2].sum(1), clipped[:, n // 2:].sum(1)), axis=1)
        sched.tell(obj, meas)
        qd.append(archive.stats.num_elites)
    assert qd[-1] > qd[0]
    assert all(e.itrs == 30 for e in emitters)


@pytest.mark.parametrize("n", [1, 2, 7, 36, 100, 101, 150])
def test_batched_jacobi_eigh(n):
    """Own eigensolver vs NumPy/LAPACK: eigenvalues, orthogonality and reconstruction."""
    import torch

    from pyribs_b200 import _lib

    rng = np.random.default_rng(n)
    batch = 5
    mats = []
    for b in range(batch):
        G = rng.standard_normal((n, n))
        A = G @ G.T / n + np.diag(rng.uniform(0, 10.0 ** rng.integers(-3, 4), n))
        if b == 0:
            A = np.eye(n)  # fully degenerate (the ES start state)
        mats.append((A + A.T) / 2)
    A = np.stack(mats)
    Ad = torch.from_numpy(A).cuda()
    eig = torch.empty(batch, n, dtype=torch.float64, device="cuda")
    V = torch.empty(batch, n, n, dtype=torch.float64, device="cuda")
    sweeps = torch.zeros(batch, dtype=torch.int32, device="cuda")
    _lib.check(_lib.lib().qd_sym_eigh_batched(n, batch, Ad.data_ptr(), eig.data_ptr(), V.data_ptr(), sweeps.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream))
    eig, V = eig.cpu().numpy(), V.cpu().numpy()
    assert int(sweeps.max()) < 40
    for b in range(batch):
        ref = np.linalg.eigvalsh(A[b])
        scale = np.abs(ref).max()
        assert np.all(np.diff(eig[b]) >= 0)
        assert np.max(np.abs(eig[b] - ref)) <= 1e-12 * scale
        assert np.max(np.abs(V[b].T @ V[b] - np.eye(n))) <= 1e-12
        assert np.max(np.abs((V[b] * eig[b]) @ V[b].T - A[b])) <= 1e-12 * scale


def _sphere(sol):
    n = sol.shape[1]
    best, worst = 0.0, (5.12 + 2.048) ** 2 * n
    obj = (worst - np.sum((sol - 2.048) ** 2, axis=1)) / (worst - best) * 100
    c = np.clip(sol, -5.12, 5.12)
    return obj, np.stack((c[:, : n // 2].sum(1), c[:, n // 2:].sum(1)), axis=1)


@pytest.mark.parametrize("mode", ["cma_me", "cma_mae", "obj"])
def test_pooled_scheduler_matches_emitter_by_emitter(mode):
    """The batched CMA-ES pool (one launch sequence for all emitters) must reproduce the
    emitter-by-emitter run exactly: same Philox streams, same kernels' arithmetic, same ranking and
    restart order -- identical archives, restart counts and strategy state after 40 iterations."""
    from pyribs_b200.archives import GridArchive
    from pyribs_b200.emitters import EvolutionStrategyEmitter
    from pyribs_b200.schedulers import Scheduler

    n, E, lam = 20, 6, 14
    mb = n / 2 * 5.12

    def build(pooled):
        if mode == "cma_mae":
            a = GridArchive(solution_dim=n, dims=(40, 40), ranges=[(-mb, mb)] * 2, learning_rate=0.01,
                            threshold_min=0.0, seed=3)
            kw = dict(ranker="imp", selection_rule="mu", restart_rule="basic")
        elif mode == "obj":
            a = GridArchive(solution_dim=n, dims=(40, 40), ranges=[(-mb, mb)] * 2, seed=3)
            kw = dict(ranker="2obj", selection_rule="filter", restart_rule=7)
        else:
            a = GridArchive(solution_dim=n, dims=(40, 40), ranges=[(-mb, mb)] * 2, seed=3)
            kw = dict(ranker="2imp", selection_rule="filter", restart_rule="no_improvement")
        seeds = np.random.SeedSequence(11).spawn(E)
        ems = [EvolutionStrategyEmitter(a, x0=np.zeros(n), sigma0=0.5, batch_size=lam, seed=s, **kw) for s in seeds]
        s = Scheduler(a, ems, pool_emitters=pooled)
        assert (s._pool is not None) == pooled
        return a, ems, s

    (a1, e1, s1), (a2, e2, s2) = build(False), build(True)
    for it in range(40):
        x1, x2 = s1.ask(), s2.ask()
        np.testing.assert_array_equal(x1, x2, err_msg=f"iteration {it}: samples differ")
        s1.tell(*_sphere(x1))
        s2.tell(*_sphere(x2))
    assert len(a1) == len(a2) and len(a1) > 20
    d1, d2 = a1.data(), a2.data()
    for k in d1:
        np.testing.assert_array_equal(d1[k], d2[k], err_msg=k)
    assert [e.restarts for e in e1] == [e.restarts for e in e2]
    assert [e.itrs for e in e1] == [e.itrs for e in e2]
    for a, b in zip(e1, e2):
        np.testing.assert_array_equal(a._opt.mean, b._opt.mean)
        # the emitter-by-emitter run has already symmetrised C for its prefetched next ask (_cma_es.py:67)
        ca, cb = a._opt.cov, b._opt.cov
        np.testing.assert_array_equal(np.maximum(ca, ca.T), np.maximum(cb, cb.T))
        assert a._opt.sigma == b._opt.sigma
    if mode == "cma_me":
        assert sum(e.restarts for e in e1) > 0  # the restart path was exercised
    # device-evaluator variant of the same loop: CUDA tensors out of ask, into tell
    import torch

    xd = s2.ask(return_device=True)
    assert xd.is_cuda and tuple(xd.shape) == (E * lam, n)
    o, m = _sphere(xd.cpu().numpy())
    s2.tell(torch.from_numpy(o).cuda(), torch.from_numpy(m).cuda())
    s1.tell(*_sphere(s1.ask()))
    np.testing.assert_array_equal(a1.data("objective"), a2.data("objective"))


# ---- streaming strategies at scale: ziggurat normals, fused sep-CMA ask, LM-MA ask as tall-skinny products ----
def _zig_fill(count, seed, offset=0):
    import torch

    from pyribs_b200 import _lib

    out = torch.empty(count, dtype=torch.float64, device="cuda")
    _lib.check(_lib.lib().qd_fill_normal_zig(out.data_ptr(), count, seed, offset,
                                             torch.cuda.current_stream().cuda_stream))
    return out


def test_ziggurat_normals():
    """qd_fill_normal_zig: a standard normal to the resolution 4M samples can see (moments, a Kolmogorov
    distance, the wedge / tail regions the slow path produces), reproducible, and counter-addressed:
    out[4q..4q+3] depend on (seed, offset + q) only."""
    import torch
    from scipy import stats

    x = _zig_fill(4_000_003, 77).cpu().numpy()
    assert np.all(np.isfinite(x))
    assert abs(x.mean()) < 2.5e-3 and abs(x.var() - 1) < 4e-3
    assert abs(stats.skew(x)) < 6e-3 and abs(stats.kurtosis(x)) < 1.5e-2
    assert stats.kstest(x[:1_000_000], "norm").statistic < 2.0e-3
    for t in (1.0, 2.0, 3.0, 4.0, 4.385945034871305):  # two-sided tail masses, incl. beyond the base layer's edge
        p, got = 2 * stats.norm.sf(t), np.mean(np.abs(x) > t)
        assert abs(got - p) < 6 * np.sqrt(p / x.size) + 1e-7, (t, got, p)
    assert abs(np.mean(x > 0) - 0.5) < 1.5e-3
    assert np.unique(x).size > 0.999 * x.size  # 53-bit variates: no visible lattice
    a = _zig_fill(4096, 5)
    assert torch.equal(a, _zig_fill(4096, 5)) and not torch.equal(a, _zig_fill(4096, 6))
    assert torch.equal(a[400:1000], _zig_fill(600, 5, offset=100))  # quad 100 onwards
    assert torch.equal(_zig_fill(4001, 5)[:4000], _zig_fill(4000, 5))  # ragged tail does not disturb the rest


@pytest.mark.parametrize("n,lam", [(1000, 64), (10, 7), (37, 50)])
def test_sep_ask_fused_generator(n, lam):
    """qd_sep_ask_rng == qd_sep_ask applied to the normals it reports (bit for bit), and for n % 4 == 0 those
    are the stream qd_fill_normal_zig produces from the same counters; row_ids redirect rows."""
    import torch

    from pyribs_b200 import _lib

    L, s = _lib.lib(), torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(0)
    dev = dict(dtype=torch.float64, device="cuda")
    mean = torch.from_numpy(rng.standard_normal(n)).cuda()
    C = torch.from_numpy(rng.uniform(0.1, 3.0, n)).cuda()
    status = torch.zeros(8, **dev)
    status[0] = 0.37
    X, z, X2 = torch.zeros(lam, n, **dev), torch.zeros(lam, n, **dev), torch.zeros(lam, n, **dev)
    _lib.check(L.qd_sep_ask_rng(lam, n, None, mean.data_ptr(), C.data_ptr(), status.data_ptr(), 99, 1000, X.data_ptr(),
                                z.data_ptr(), s))
    _lib.check(L.qd_sep_ask(z.data_ptr(), lam, n, None, mean.data_ptr(), C.data_ptr(), status.data_ptr(),
                            X2.data_ptr(), s))
    assert torch.equal(X, X2)
    ref = np.sqrt(C.cpu().numpy()) * (0.37 * z.cpu().numpy()) + mean.cpu().numpy()  # _sep_cma_es.py:160-179
    np.testing.assert_array_equal(X.cpu().numpy(), ref)
    if n % 4 == 0:
        assert torch.equal(z.reshape(-1), _zig_fill(lam * n, 99, offset=1000))
    zz = z.cpu().numpy()
    assert abs(zz.mean()) < 5 / np.sqrt(zz.size) and abs(zz.var() - 1) < 8 / np.sqrt(zz.size)
    # without z_out, and with redirected rows
    ids = torch.from_numpy(rng.permutation(lam).astype(np.int32)).cuda()
    X3 = torch.zeros(lam, n, **dev)
    _lib.check(L.qd_sep_ask_rng(lam, n, ids.data_ptr(), mean.data_ptr(), C.data_ptr(), status.data_ptr(), 99, 1000,
                                X3.data_ptr(), None, s))
    assert torch.equal(X3[ids.long()], X)


def test_sep_cma_default_ask_is_reproducible_and_normal():
    from pyribs_b200.emitters.opt import SeparableCMAEvolutionStrategy

    def run(seed):
        es = SeparableCMAEvolutionStrategy(0.5, 2000, 128, seed=seed)
        es.reset(np.full(2000, 0.25))
        a = es.ask()
        es.tell(np.arange(128), None, 64)
        return a, es.ask()

    (a1, b1), (a2, b2), (a3, _) = run(4), run(4), run(5)
    np.testing.assert_array_equal(a1, a2)
    np.testing.assert_array_equal(b1, b2)
    assert not np.array_equal(a1, a3) and not np.array_equal(a1, b1)
    zz = (a1 - 0.25) / 0.5
    assert abs(zz.mean()) < 0.01 and abs(zz.var() - 1) < 0.02


@pytest.mark.parametrize("n,lam,nv,gens", [(1003, 77, 37, 40), (256, 33, None, 6), (64, 5, 3, 5), (2048, 128, 70, 72)])
def test_lmma_ask_matches_the_sequential_transforms(n, lam, nv, gens):
    """LM-MA ask as Z M^T, M M^T, forward substitution and one output pass against the oracle's rank-one
    transforms applied one after the other (_lm_ma_es.py:136-144), with ragged sizes and more direction vectors
    than one register group; the strategies are driven in lockstep so the vectors are realistic."""
    from pyribs_b200.emitters.opt import LMMAEvolutionStrategy

    kw = {} if nv is None else {"n_vectors": nv}
    gpu = LMMAEvolutionStrategy(0.8, n, lam, seed=2, **kw)
    ora = O.OracleLMMAES(0.8, n, lam, seed=2, **kw)
    x0 = np.linspace(-1, 1, n)
    gpu.reset(x0), ora.reset(x0)
    # the reference's cd (1 / (1.5^j n)) makes the transforms nearly the identity; scale it up so that an error in
    # the coefficients would be visible at the tolerance
    boost = np.minimum(0.3, ora.cd * n * 0.2)
    ora.cd = boost
    gpu._cd_d.copy_(__import__("torch").from_numpy(boost))
    rng = np.random.default_rng(n + lam)
    for g in range(gens):
        z = rng.standard_normal((lam, n))
        so, sg = ora.ask(z), gpu.ask(z=z)
        assert_close(sg, so, 1e-9, f"gen {g} solutions")
        if g == 0:
            np.testing.assert_array_equal(sg, so)  # no transform yet: mean + sigma z, bit for bit
        rank = np.argsort(-(-np.sum((so - 0.5) ** 2, axis=1)))[::-1].copy()
        ora.tell(rank, lam // 2), gpu.tell(rank, None, lam // 2)
    assert_close(gpu.m, ora.m, 1e-9, "direction vectors")
    assert_close(gpu.mean, ora.mean, 1e-9, "mean")


def test_lmma_default_ask_blocks_and_streams():
    """The default LM-MA ask generates its normals inside the call, row block by row block on two side streams
    (blocks of 40 MB of z set through qd_lmma_set_block_bytes: three blocks here). z must be the flat ziggurat stream of the strategy's counters,
    and the solutions must equal, bit for bit, the single-stream ask applied to that z afterwards."""
    import torch

    from pyribs_b200 import _lib
    from pyribs_b200.emitters.opt import LMMAEvolutionStrategy

    n, lam = 8192, 1280
    prev = _lib.lib().qd_lmma_set_block_bytes(40 << 20)
    try:
        _lmma_blocks_body(n, lam)
    finally:
        _lib.lib().qd_lmma_set_block_bytes(prev)


def _lmma_blocks_body(n, lam):
    import torch

    from pyribs_b200 import _lib
    from pyribs_b200.emitters.opt import LMMAEvolutionStrategy

    es = LMMAEvolutionStrategy(0.3, n, lam, seed=7, n_vectors=10)
    es.reset(np.zeros(n))
    for g in range(3):
        off = es._rng_offset
        X = es.ask(return_device=True)
        torch.cuda.synchronize()
        z = es._z
        assert es._rng_offset == off + lam * n // 4
        assert torch.equal(z.reshape(-1), _zig_fill(lam * n, es._seed, offset=off))
        # replay with the normals injected: same kernels, one block at a time on the caller's stream
        X2 = torch.empty_like(X)
        itrs = min(es.current_gens, es.n_vectors)
        ws = torch.empty(int(_lib.lib().qd_lmma_ask_workspace_doubles(lam, n, itrs)), dtype=torch.float64, device="cuda")
        with torch.cuda.stream(es._es_stream()):
            _lib.check(_lib.lib().qd_lmma_ask(z.data_ptr(), lam, n, None, es._mean.data_ptr(), es._status.data_ptr(),
                                              es._m.data_ptr(), es._cd_d.data_ptr(), itrs, X2.data_ptr(), ws.data_ptr(),
                                              es._es_stream().cuda_stream))
        torch.cuda.synchronize()
        assert torch.equal(X, X2)
        fit = -(X - 0.1).pow(2).sum(1)
        es.tell(torch.argsort(fit, descending=True), None, lam // 2)
    assert es.sigma > 0 and np.all(np.isfinite(es.mean)) and np.any(es.m != 0)


def test_lmma_bounds_with_fused_generator():
    from pyribs_b200.emitters.opt import LMMAEvolutionStrategy

    n = 40
    es = LMMAEvolutionStrategy(0.8, n, 24, seed=3, lower_bounds=np.full(n, -2.0), upper_bounds=np.full(n, 2.0))
    es.reset(np.zeros(n))
    for g in range(3):
        sol = es.ask()
        assert sol.shape == (24, n) and np.all(sol >= -2.0) and np.all(sol <= 2.0)
        # the retained normals belong to the accepted samples: generation 0 has no transform, x = sigma z
        if g == 0:
            np.testing.assert_array_equal(sol, 0.8 * es._z.cpu().numpy())
        es.tell(np.argsort(np.sum(sol ** 2, axis=1)), None, 12)


# ---- OpenAI-ES ------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(load_golden("openai_es").keys()))
def test_openai_es_trace(name):
    """ribs/emitters/opt/_openai_es.py + AdamOpt against the reference's traces (noise injected: the half batch
    under mirror sampling). Solutions are bit-identical (theta + sigma0 * z); theta / m / v within 1e-6 relative
    (observed ~1e-13: the gradient's row sum is blocked differently from NumPy's)."""
    from pyribs_b200.emitters.opt import _get_es

    case = load_golden("openai_es")[name]
    cfg = case["cfg"]
    n, lam, mirror = int(cfg["n"]), int(cfg["lam"]), bool(int(cfg["mirror"]))
    es = _get_es("openai_es", sigma0=float(cfg["sigma0"]), solution_dim=n, batch_size=lam, seed=1, dtype=np.float64,
                 lower_bounds=-np.inf, upper_bounds=np.inf, mirror_sampling=mirror, lr=float(cfg["lr"]),
                 l2_coeff=float(cfg["l2"]), beta1=float(cfg["beta1"]))
    es.reset(cfg["x0"])
    for gi, g in enumerate(gens_of(case)):
        z = np.random.default_rng(int(g["zseed"])).standard_normal((lam // 2 if mirror else lam, n))
        sol = es.ask(z=z)
        if gi == 0:
            np.testing.assert_array_equal(sol, g["solutions"], err_msg=f"{name} g0 solutions")
        assert_close(sol, g["solutions"], RTOL, f"{name} g{gi} solutions")
        es.tell(g["ranking"], None, lam // 2)
        assert_close(es.theta, g["theta"], RTOL, f"{name} g{gi} theta")
        assert_close(es._to_host(es._m), g["m"], RTOL, f"{name} g{gi} m")
        assert_close(es._to_host(es._v), g["v"], RTOL, f"{name} g{gi} v")
        fit = -np.sum((np.asarray(g["solutions"]) - 1.0) ** 2, axis=1)
        assert int(es.check_stop(fit[g["ranking"]])) == int(g["stop"])
        assert_close(es.last_update_ratio, g["ratio"], 1e-6, f"{name} g{gi} ratio")


def test_openai_es_default_ask_and_emitter():
    """Own normal stream: the two halves of a mirrored batch are reflections of each other about theta; the
    contract checks of the constructor (_openai_es.py:80-111); and an EvolutionStrategyEmitter(es="openai_es") makes
    progress on the sphere function."""
    from pyribs_b200.archives import GridArchive
    from pyribs_b200.emitters import EvolutionStrategyEmitter
    from pyribs_b200.emitters.opt import OpenAIEvolutionStrategy
    from pyribs_b200.schedulers import Scheduler

    n = 50
    es = OpenAIEvolutionStrategy(0.2, n, 32, seed=5)
    es.reset(np.full(n, 0.5))
    x = es.ask()
    assert x.shape == (32, n)
    np.testing.assert_allclose(x[:16] - 0.5, -(x[16:] - 0.5), rtol=0, atol=1e-15)
    np.testing.assert_array_equal(es.noise[:16], -es.noise[16:])
    np.testing.assert_array_equal(x, 0.5 + 0.2 * es.noise)  # theta + sigma0 * noise, bit for bit
    zz = (x[:16] - 0.5) / 0.2
    assert abs(zz.mean()) < 0.15 and abs(zz.var() - 1) < 0.2
    es2 = OpenAIEvolutionStrategy(0.2, n, 32, seed=5)
    es2.reset(np.full(n, 0.5))
    np.testing.assert_array_equal(es2.ask(), x)
    assert OpenAIEvolutionStrategy(0.1, 100).batch_size % 2 == 0
    with pytest.raises(ValueError):
        OpenAIEvolutionStrategy(0.1, n, 7)
    with pytest.raises(ValueError):
        OpenAIEvolutionStrategy(0.1, n, 8, lower_bounds=-1.0, upper_bounds=1.0)
    b = OpenAIEvolutionStrategy(0.3, n, 9, mirror_sampling=False, lower_bounds=-1.0, upper_bounds=1.0, seed=2)
    b.reset(np.zeros(n))
    xb = b.ask()
    assert xb.shape == (9, n) and np.all(np.abs(xb) <= 1.0)
    a = GridArchive(solution_dim=n, dims=(30, 30), ranges=[(-25.6, 25.6)] * 2, seed=1)
    ems = [EvolutionStrategyEmitter(a, x0=np.zeros(n), sigma0=0.3, es="openai_es", es_kwargs={"lr": 0.05},
                                    ranker="2imp", batch_size=24, seed=s) for s in (1, 2)]
    sched = Scheduler(a, ems)
    best = []
    for it in range(40):
        sol = sched.ask()
        obj = -np.sum((sol - 2.0) ** 2, axis=1)
        sched.tell(obj, np.stack([sol[:, : n // 2].sum(1), sol[:, n // 2:].sum(1)], axis=1))
        best.append(a.stats.obj_max)
    assert best[-1] > best[0] and len(a) > 10
