"""EvolutionStrategyEmitter (reference ribs/emitters/_evolution_strategy_emitter.py:20-265,
bounds handling ribs/emitters/_emitter_base.py:47-123).

STAND-IN, not a built component. The restart / ranking policy of this emitter is host logic with exactly one
legal behaviour, so this file is necessarily a restatement of the reference's; nothing here is on the GPU. It
exists so that examples, tests and benchmarks run on boxes where `ribs` is not installed. Where `ribs` IS
installed, the reference's own emitter drives the device strategies of this repo directly -- pass
`es=pyribs_b200.emitters.opt.<class>` (a callable, ribs/emitters/opt/__init__.py:149-155) to
`ribs.emitters.EvolutionStrategyEmitter`; tests/test_gpu_ref_scheduler.py::test_reference_emitter_over_device_es
does exactly that. The product of this package is the strategies (emitters/opt, csrc/es*.cu) and the archives."""
from __future__ import annotations

import numbers

import numpy as np

from pyribs_b200._utils import check_shape, validate_batch
from pyribs_b200.emitters.opt import _get_es
from pyribs_b200.emitters.rankers import _get_ranker


def _process_bounds(bounds, solution_dim, dtype):
    lower = np.full(solution_dim, -np.inf, dtype=dtype)
    upper = np.full(solution_dim, np.inf, dtype=dtype)
    if bounds is None:
        return lower, upper
    if len(bounds) != solution_dim:
        raise ValueError("If it is an array-like, bounds must have length solution_dim")
    for i, bnd in enumerate(bounds):
        if bnd is None:
            continue
        if len(bnd) != 2:
            raise ValueError("All entries of bounds must be length 2")
        lower[i] = -np.inf if bnd[0] is None else bnd[0]
        upper[i] = np.inf if bnd[1] is None else bnd[1]
    return lower, upper


class EvolutionStrategyEmitter:
    """Adapts an evolution strategy to a ranking of the solutions' archive outcomes.

    Same constructor, `ask` / `tell` and properties as the reference. The ES state and the
    sampled solutions stay on the GPU between `ask` and `tell`; `ask()` hands the caller a
    NumPy copy (drop-in), `ask(return_device=True)` the CUDA tensor.
    """

    def __init__(self, archive, *, x0, sigma0, ranker="2imp", es="cma_es", es_kwargs=None, selection_rule="filter",
                 restart_rule="no_improvement", bounds=None, lower_bounds=None, upper_bounds=None, batch_size=None,
                 seed=None):
        self._archive = archive
        self._solution_dim = archive.solution_dim
        sdt = archive.dtypes["solution"]
        if bounds is not None and (lower_bounds is not None or upper_bounds is not None):
            raise ValueError(
                "Cannot specify both bounds and lower_bounds/upper_bounds; either specify bounds or specify "
                "lower_bounds/upper_bounds.")
        if bounds is not None:
            self._lower_bounds, self._upper_bounds = _process_bounds(bounds, self._solution_dim, sdt)
        else:
            self._lower_bounds = (np.full(self._solution_dim, -np.inf, dtype=sdt) if lower_bounds is None else
                                  np.asarray(lower_bounds, dtype=sdt))
            self._upper_bounds = (np.full(self._solution_dim, np.inf, dtype=sdt) if upper_bounds is None else
                                  np.asarray(upper_bounds, dtype=sdt))
            check_shape(self._lower_bounds, "lower_bounds", self._solution_dim, "solution_dim")
            check_shape(self._upper_bounds, "upper_bounds", self._solution_dim, "solution_dim")
        seed_sequence = seed if isinstance(seed, np.random.SeedSequence) else np.random.SeedSequence(seed)
        opt_seed, ranker_seed = seed_sequence.spawn(2)  # :108-113
        self._x0 = np.asarray(x0, dtype=sdt).copy()
        check_shape(self._x0, "x0", archive.solution_dim, "archive.solution_dim")
        self._sigma0 = sigma0
        if selection_rule not in ["mu", "filter"]:
            raise ValueError(f"Invalid selection_rule {selection_rule}")
        self._selection_rule = selection_rule
        self._restart_rule = restart_rule
        self._restarts = 0
        self._itrs = 0
        _ = self._check_restart(0)  # validates restart_rule
        self._opt = _get_es(es, sigma0=sigma0, batch_size=batch_size, solution_dim=self._solution_dim, seed=opt_seed,
                            dtype=sdt, lower_bounds=self._lower_bounds, upper_bounds=self._upper_bounds,
                            **(es_kwargs if es_kwargs is not None else {}))
        self._opt.reset(self._x0)
        self._ranker = _get_ranker(ranker, ranker_seed)
        self._ranker.reset(self, archive)
        self._batch_size = self._opt.batch_size

    @property
    def archive(self):
        return self._archive

    @property
    def solution_dim(self):
        return self._solution_dim

    @property
    def lower_bounds(self):
        return self._lower_bounds

    @property
    def upper_bounds(self):
        return self._upper_bounds

    @property
    def x0(self):
        return self._x0

    @property
    def batch_size(self):
        return self._batch_size

    @property
    def restarts(self):
        return self._restarts

    @property
    def itrs(self):
        return self._itrs

    def ask(self, **kwargs):
        return self._opt.ask(**kwargs)

    def _check_restart(self, num_parents):
        if isinstance(self._restart_rule, numbers.Integral):
            return self._itrs % self._restart_rule == 0
        if self._restart_rule == "no_improvement":
            return num_parents == 0
        if self._restart_rule == "basic":
            return False
        raise ValueError(f"Invalid restart_rule {self._restart_rule}")

    def tell(self, solution, objective, measures, add_info, **fields):
        """:203-265: rank, pick the parents, update the ES, restart from a random elite when the
        ES has converged or the restart rule fires."""
        data, add_info = validate_batch(
            self._archive, {"solution": solution, "objective": objective, "measures": measures, **fields}, add_info)
        self._itrs += 1
        status = np.asarray(add_info["status"].cpu().numpy() if hasattr(add_info["status"], "cpu") else
                            add_info["status"])
        host_info = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in add_info.items()}
        host_data = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in data.items() if k != "solution"}
        new_sols = status.astype(bool).sum()
        indices, ranking_values = self._ranker.rank(self, self._archive, host_data, host_info)
        num_parents = new_sols if self._selection_rule == "filter" else self._batch_size // 2
        self._opt.tell(indices, ranking_values, num_parents)
        if self._opt.check_stop(ranking_values[indices]) or self._check_restart(new_sols):
            new_x0 = self._archive.sample_elites(1)["solution"][0]
            self._opt.reset(new_x0)
            self._ranker.reset(self, self._archive)
            self._restarts += 1
        if hasattr(self._opt, "prefetch_ask"):
            self._opt.prefetch_ask()  # sampling for the next round overlaps with the other emitters' tells
