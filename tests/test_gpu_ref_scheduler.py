"""The drop-in claim itself: the REFERENCE's own `ribs.schedulers.Scheduler` / `BanditScheduler`
(ribs/schedulers/_scheduler.py:180-213, 250-296, 361-412; _bandit_scheduler.py:205-427), unmodified, drive this
repo's archives and emitters. The reference package is the offline install under baseline/_ref (it travels to the
GPU box) or the source tree in the authoring container; the two pure-Python packages it imports besides
NumPy/SciPy come from oracle/shim. Skipped where neither exists.

What is compared: the trajectory under the reference's scheduler equals, bit for bit, the trajectory of this
repo's stand-in scheduler (pyribs_b200/schedulers) on twins built from the same seeds -- i.e. the classes behave
the same behind either scheduler, which is all "drop-in behind the Scheduler" can mean when the sample streams
(Philox here, PCG64 in the reference) differ by construction."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _ref_schedulers():
    for ref in (os.environ.get("QD_REFERENCE", "/root/reference"), os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isdir(os.path.join(ref, "ribs")):
            break
    else:
        pytest.skip("reference package not present (neither /root/reference nor baseline/_ref)")
    for p in (os.path.join(ROOT, "oracle", "shim"), ref):
        if p not in sys.path:
            sys.path.insert(0, p)
    import ribs.schedulers as S

    assert "pyribs_b200" not in S.__file__
    return S


def _sphere(x):
    x = np.asarray(x)
    return -np.sum((x - 0.4) ** 2, axis=1), np.clip(x[:, :2], -1, 1)


def _make(kind, es="cma_es"):
    from pyribs_b200.archives import CVTArchive, GridArchive
    from pyribs_b200.emitters import EvolutionStrategyEmitter, GaussianEmitter

    n = 16
    if kind == "grid":
        a = GridArchive(solution_dim=n, dims=(20, 20), ranges=[(-1, 1)] * 2, seed=1)
        r = None
    elif kind == "mae":
        a = GridArchive(solution_dim=n, dims=(20, 20), ranges=[(-1, 1)] * 2, learning_rate=0.1, threshold_min=-50.0,
                        seed=1)
        r = GridArchive(solution_dim=n, dims=(20, 20), ranges=[(-1, 1)] * 2, seed=2)
    else:
        cen = np.random.default_rng(3).uniform(-1, 1, (300, 2))
        a = CVTArchive(solution_dim=n, centroids=cen, ranges=[(-1, 1)] * 2, seed=1)
        r = None
    ems = [EvolutionStrategyEmitter(a, x0=np.zeros(n), sigma0=0.3, batch_size=16, seed=10 + s, es=es,
                                    ranker="imp" if kind == "mae" else "2imp") for s in range(3)]
    ems.append(GaussianEmitter(a, sigma=0.2, x0=np.zeros(n), batch_size=8, seed=99))
    return a, ems, r


def _dump(a):
    d = a.data()
    order = np.argsort(d["index"])
    return {k: np.asarray(v)[order] for k, v in d.items()}


@pytest.mark.parametrize("kind,es", [("grid", "cma_es"), ("mae", "sep_cma_es"), ("cvt", "lm_ma_es")])
def test_reference_scheduler_drives_these_classes(kind, es):
    S = _ref_schedulers()
    from pyribs_b200.schedulers import Scheduler as Standin

    a1, e1, r1 = _make(kind, es)
    a2, e2, r2 = _make(kind, es)
    ref = S.Scheduler(a1, e1, result_archive=r1)
    own = Standin(a2, e2, result_archive=r2, pool_emitters=False)
    for it in range(12):
        x1, x2 = ref.ask(), own.ask()
        np.testing.assert_array_equal(np.asarray(x1), np.asarray(x2), err_msg=f"ask {it}")
        assert np.asarray(x1).shape == (3 * 16 + 8, 16)
        o, m = _sphere(x1)
        ref.tell(o, m)
        own.tell(o, m)
    d1, d2 = _dump(a1), _dump(a2)
    assert len(d1["index"]) > 20
    for k in d1:
        np.testing.assert_array_equal(d1[k], d2[k], err_msg=k)
    assert a1.stats.qd_score == a2.stats.qd_score and a1.stats.num_elites == a2.stats.num_elites
    if r1 is not None:
        for k, v in _dump(r1).items():
            np.testing.assert_array_equal(v, _dump(r2)[k], err_msg=f"result archive {k}")
    # the pooled fast path of the stand-in (one launch sequence for all strategies) is the same trajectory again
    if es != "lm_ma_es":
        a3, e3, r3 = _make(kind, es)
        pooled = Standin(a3, e3[:3], result_archive=r3)
        a4, e4, r4 = _make(kind, es)
        ref2 = S.Scheduler(a4, e4[:3], result_archive=r4)
        for it in range(6):
            x3, x4 = np.asarray(pooled.ask()), np.asarray(ref2.ask())
            np.testing.assert_array_equal(x3, x4, err_msg=f"pooled ask {it}")
            o, m = _sphere(x3)
            pooled.tell(o, m), ref2.tell(o, m)
        assert a3.stats.qd_score == a4.stats.qd_score


def test_reference_scheduler_add_mode_single():
    S = _ref_schedulers()
    a1, e1, _ = _make("grid")
    a2, e2, _ = _make("grid")
    single, batch = S.Scheduler(a1, e1, add_mode="single"), S.Scheduler(a2, e2, add_mode="batch")
    x1, x2 = single.ask(), batch.ask()
    np.testing.assert_array_equal(np.asarray(x1), np.asarray(x2))
    o, m = _sphere(x1)
    single.tell(o, m), batch.tell(o, m)
    # one generation, CMA-ME: inserting one by one leaves the same elites as the batch rule
    d1, d2 = _dump(a1), _dump(a2)
    for k in ("index", "objective", "solution"):
        np.testing.assert_array_equal(d1[k], d2[k], err_msg=k)


def test_reference_bandit_scheduler_drives_these_classes():
    S = _ref_schedulers()
    from pyribs_b200.schedulers import BanditScheduler as Standin

    def make():
        from pyribs_b200.archives import GridArchive
        from pyribs_b200.emitters import EvolutionStrategyEmitter, IsoLineEmitter

        a = GridArchive(solution_dim=8, dims=(15, 15), ranges=[(-1, 1)] * 2, seed=4)
        pool = [EvolutionStrategyEmitter(a, x0=np.zeros(8), sigma0=0.3, batch_size=10, seed=20 + s, es="sep_cma_es")
                for s in range(4)]
        pool += [IsoLineEmitter(a, x0=np.zeros(8), batch_size=10, seed=40 + s) for s in range(2)]
        return a, pool

    a1, p1 = make()
    a2, p2 = make()
    ref = S.BanditScheduler(a1, p1, num_active=3, reselect="terminated")
    own = Standin(a2, p2, num_active=3, reselect="terminated")
    for it in range(15):
        x1, x2 = np.asarray(ref.ask()), np.asarray(own.ask())
        np.testing.assert_array_equal(x1, x2, err_msg=f"ask {it}")
        np.testing.assert_array_equal(np.asarray(ref.active), np.asarray(own.active))
        o = -np.sum((x1 - 0.3) ** 2, axis=1)
        m = np.clip(x1[:, :2], -1, 1)
        ref.tell(o, m), own.tell(o, m)
    assert a1.stats.qd_score == a2.stats.qd_score and len(a1) == len(a2) > 10


@pytest.mark.parametrize("es_name", ["cma_es", "sep_cma_es", "lm_ma_es"])
def test_reference_emitter_over_device_es(es_name):
    """The reference's own EvolutionStrategyEmitter (ribs/emitters/_evolution_strategy_emitter.py:108-265) with a
    device strategy of this repo passed as the `es=` callable, inside the reference's Scheduler, over this repo's
    archive: same trajectory as this repo's stand-in emitter from the same seeds."""
    S = _ref_schedulers()
    import ribs.emitters as RE

    from pyribs_b200.archives import GridArchive
    from pyribs_b200.emitters import EvolutionStrategyEmitter as Standin
    from pyribs_b200.emitters import opt
    from pyribs_b200.schedulers import Scheduler as OwnScheduler

    opt.register_with_reference()
    klass = opt._NAME_TO_ES_MAP[es_name]
    n = 12

    def make(emitter_cls, es):
        a = GridArchive(solution_dim=n, dims=(20, 20), ranges=[(-1, 1)] * 2, seed=5)
        return a, [emitter_cls(a, x0=np.full(n, 0.1), sigma0=0.25, batch_size=12, seed=7 + s, es=es, ranker="2imp",
                               restart_rule="no_improvement") for s in range(2)]

    a1, e1 = make(RE.EvolutionStrategyEmitter, klass)
    a2, e2 = make(Standin, es_name)
    assert "pyribs_b200" not in type(e1[0]).__module__ and isinstance(e1[0]._opt, klass)
    ref, own = S.Scheduler(a1, e1), OwnScheduler(a2, e2, pool_emitters=False)
    for it in range(15):
        x1, x2 = np.asarray(ref.ask()), np.asarray(own.ask())
        np.testing.assert_array_equal(x1, x2, err_msg=f"ask {it}")
        o, m = _sphere(x1)
        ref.tell(o, m), own.tell(o, m)
    assert a1.stats.qd_score == a2.stats.qd_score and len(a1) == len(a2) > 10
    assert [e.restarts for e in e1] == [e.restarts for e in e2]
