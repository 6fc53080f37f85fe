"""GPU parity tests of the MAP-Elites variation emitters (csrc/emit.cu through qd_emit_variation) against the
golden vectors generated from the reference (oracle/gen_golden.py gen_emitters) and the oracle."""
import os
import sys

import numpy as np
import pytest

from golden_util import load_golden, steps_of

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import qd_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["f64", "f64_bounds", "f32"])
def test_variation_emitters_golden(tag):
    """ribs/emitters/_gaussian_emitter.py:139-166, _iso_line_emitter.py:156-201. The archive draws the parents
    with its own generator (seed 77, as the reference run did) and the kernel gathers their rows on the
    device; with the reference's normals injected the asked solutions are bit-identical."""
    from pyribs_b200.archives import GridArchive
    from pyribs_b200.emitters import GaussianEmitter, IsoLineEmitter

    case = load_golden("emitters")[tag]
    cfg = case["cfg"]
    dtype = np.float32 if int(cfg["f32"]) else np.float64
    n, B = int(cfg["n"]), int(cfg["B"])
    a = GridArchive(solution_dim=n, dims=(12, 12), ranges=[(-1, 1)] * 2, seed=77, dtype=dtype)
    lo = np.full(n, -0.8, dtype=dtype) if int(cfg["bounded"]) else None
    hi = np.full(n, 0.9, dtype=dtype) if int(cfg["bounded"]) else None
    ge = GaussianEmitter(a, sigma=cfg["sigma"], x0=cfg["gx0"], lower_bounds=lo, upper_bounds=hi, batch_size=B, seed=5)
    ie = IsoLineEmitter(a, iso_sigma=float(cfg["iso_sigma"]), line_sigma=float(cfg["line_sigma"]), x0=cfg["ix0"],
                        lower_bounds=lo, upper_bounds=hi, batch_size=B, seed=6)
    for j, s in enumerate(steps_of(case)):
        if j == 1:
            a.add(**case["fill"])
        xg = ge.ask(z=s["gauss_z"])
        xi = ie.ask(z=s["iso_z"], line_z=s["line_z"])
        assert xg.dtype == dtype and xg.shape == (B, n) and xi.dtype == dtype and xi.shape == (B, n)
        np.testing.assert_array_equal(xg, s["gauss_x"], err_msg=f"{tag} s{j} gaussian")
        np.testing.assert_array_equal(xi, s["iso_x"], err_msg=f"{tag} s{j} isoline")


def test_variation_emitters_vs_oracle_large():
    """A batch far beyond the golden's size: same parents (both archives seeded alike), same normals."""
    from pyribs_b200.archives import GridArchive
    from pyribs_b200.emitters import GaussianEmitter, IsoLineEmitter

    rng = np.random.default_rng(3)
    n, B = 33, 5000
    fill = dict(solution=rng.standard_normal((20000, n)), objective=rng.standard_normal(20000),
                measures=rng.uniform(-1, 1, (20000, 2)))
    g = GridArchive(solution_dim=n, dims=(40, 40), ranges=[(-1, 1)] * 2, seed=9)
    o = O.make_grid_oracle(solution_dim=n, dims=(40, 40), ranges=[(-1, 1)] * 2, seed=9)
    g.add(**fill), o.add(**fill)
    lo, hi = np.full(n, -1.5), np.linspace(0.5, 2.0, n)
    ge = GaussianEmitter(g, sigma=0.3, x0=np.zeros(n), lower_bounds=lo, upper_bounds=hi, batch_size=B)
    oe = O.OracleGaussianEmitter(o, 0.3, np.zeros(n), B, lo, hi)
    ie = IsoLineEmitter(g, iso_sigma=0.05, line_sigma=0.4, x0=np.zeros(n), bounds=list(zip(lo, hi)), batch_size=B)
    oi = O.OracleIsoLineEmitter(o, 0.05, 0.4, np.zeros(n), B, lo, hi)
    for _ in range(2):
        z, lz = rng.standard_normal((B, n)), rng.standard_normal((B, 1))
        np.testing.assert_array_equal(ge.ask(z=z), oe.ask(z))
        np.testing.assert_array_equal(ie.ask(z=z, line_z=lz), oi.ask(z, lz))


def test_variation_emitters_interface():
    """Constructor contract of the reference (errors, initial_solutions, properties; _gaussian_emitter.py:65-137,
    _iso_line_emitter.py:71-154) and the library's own normal stream (seeded, reproducible)."""
    from pyribs_b200.archives import GridArchive
    from pyribs_b200.emitters import GaussianEmitter, IsoLineEmitter

    a = GridArchive(solution_dim=4, dims=(10, 10), ranges=[(-1, 1)] * 2, seed=1)
    with pytest.raises(ValueError):
        GaussianEmitter(a, sigma=0.1)
    with pytest.raises(ValueError):
        GaussianEmitter(a, sigma=0.1, x0=np.zeros(4), initial_solutions=np.zeros((2, 4)))
    with pytest.raises(ValueError):
        IsoLineEmitter(a, x0=np.zeros(5))
    with pytest.raises(ValueError):
        IsoLineEmitter(a, x0=np.zeros(4), bounds=[(-1, 1)] * 4, lower_bounds=np.zeros(4))
    init = np.array([[0.0, 1.0, 2.0, 3.0], [4.0, 5.0, 6.0, 7.0]])
    e = GaussianEmitter(a, sigma=0.1, initial_solutions=init, bounds=[(-1, 5)] * 4, batch_size=8, seed=3)
    np.testing.assert_array_equal(e.ask(), np.clip(init, -1, 5))  # empty archive: the initial solutions, clipped
    assert e.x0 is None and e.batch_size == 8 and e.sigma == 0.1
    np.testing.assert_array_equal(e.lower_bounds, np.full(4, -1.0))
    GaussianEmitter(a, sigma=0.1, x0=np.zeros(4), seed=np.random.SeedSequence(3).spawn(1)[0]).ask()  # as sphere.py seeds them
    e1 = IsoLineEmitter(a, x0=np.ones(4), batch_size=16, seed=11)
    e2 = IsoLineEmitter(a, x0=np.ones(4), batch_size=16, seed=11)
    x1, x2 = e1.ask(), e2.ask()
    np.testing.assert_array_equal(x1, x2)
    assert x1.shape == (16, 4) and np.all(np.abs(x1 - 1.0) < 0.2) and np.std(x1) > 0  # x0 + iso noise (directions = 0)
    assert not np.array_equal(e1.ask(), x1)  # the stream advances
    e1.tell(x1, np.zeros(16), np.zeros((16, 2)), {"status": np.zeros(16), "value": np.zeros(16)})


def test_map_elites_runs_on_the_scheduler():
    """MAP-Elites with Iso+Line on the sphere function (the reference's `map_elites` example configuration,
    examples/sphere.py): coverage and QD score grow."""
    from pyribs_b200.archives import GridArchive
    from pyribs_b200.emitters import IsoLineEmitter
    from pyribs_b200.schedulers import Scheduler

    n = 20
    a = GridArchive(solution_dim=n, dims=(50, 50), ranges=[(-51.2, 51.2)] * 2, seed=2)
    ems = [IsoLineEmitter(a, x0=np.zeros(n), iso_sigma=0.1, line_sigma=0.2, batch_size=64, seed=s) for s in (1, 2, 3)]
    sched = Scheduler(a, ems)
    hist = []
    for it in range(40):
        x = sched.ask()
        assert x.shape == (192, n)
        shift = 5.12 * 0.4
        obj = -np.sum((x - shift) ** 2, axis=1)
        clipped = np.where(np.abs(x) <= 5.12, x, 5.12 / np.where(x == 0, 1, x))
        meas = np.stack([clipped[:, : n // 2].sum(1), clipped[:, n // 2:].sum(1)], axis=1)
        sched.tell(obj, meas)
        hist.append(len(a))
    assert hist[-1] > hist[5] > hist[0] >= 1


def test_bandit_scheduler_over_device_emitters():
    """The reference's `me_map_elites` configuration in small (examples/sphere.py): a UCB1 bandit over a pool of
    ES and variation emitters, all inserting into one device archive."""
    from pyribs_b200.archives import GridArchive
    from pyribs_b200.emitters import EvolutionStrategyEmitter, GaussianEmitter, IsoLineEmitter
    from pyribs_b200.schedulers import BanditScheduler

    n = 12
    a = GridArchive(solution_dim=n, dims=(30, 30), ranges=[(-30.72, 30.72)] * 2, seed=4)
    pool = ([EvolutionStrategyEmitter(a, x0=np.zeros(n), sigma0=0.5, ranker="2imp", batch_size=16, seed=s)
             for s in range(3)] +
            [GaussianEmitter(a, sigma=0.3, x0=np.zeros(n), batch_size=16, seed=10 + s) for s in range(2)] +
            [IsoLineEmitter(a, x0=np.zeros(n), batch_size=16, seed=20 + s) for s in range(2)])
    sched = BanditScheduler(a, pool, num_active=3, reselect="terminated")
    sizes = []
    for it in range(30):
        x = sched.ask()
        assert x.shape == (48, n) and sched.active.sum() == 3
        obj = -np.sum((x - 2.0) ** 2, axis=1)
        clipped = np.where(np.abs(x) <= 5.12, x, 5.12 / np.where(x == 0, 1, x))
        meas = np.stack([clipped[:, : n // 2].sum(1), clipped[:, n // 2:].sum(1)], axis=1)
        sched.tell(obj, meas)
        sizes.append(len(a))
    assert sizes[-1] > sizes[3] > 0
    assert sched._selection.sum() == 30 * 48 and 0 < sched._success.sum() <= sched._selection.sum()
    assert np.count_nonzero(sched._selection) >= 4  # the variation emitters (no restart counter) rotate
