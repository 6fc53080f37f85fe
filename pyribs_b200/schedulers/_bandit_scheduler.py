"""BanditScheduler: a UCB1 bandit picks which emitters of a pool are active.

Call contract and selection rule of ribs/schedulers/_bandit_scheduler.py:82-427 (constructor checks :82-170,
selection in `ask` :233-318, credit assignment in `tell` :415-427). The policy is host arithmetic over a handful
of emitters; the batches themselves go through the same device insertion as `Scheduler` (one index_of shared with
`result_archive` when the geometries match).

STAND-IN, not a built component: the reference's `ribs.schedulers.BanditScheduler` works unchanged on top of
these archives and emitters (tests/test_gpu_ref_scheduler.py::test_reference_bandit_scheduler_drives_these_classes
runs it and compares trajectories with this file's). This restatement of its host policy exists only for boxes
where `ribs` is not installed.
"""
from __future__ import annotations

import numpy as np

from pyribs_b200._utils import arr_readonly
from pyribs_b200.schedulers._scheduler import Scheduler


class BanditScheduler:
    def __init__(self, archive, emitter_pool, result_archive=None, *, num_active, reselect="terminated", zeta=0.05,
                 add_mode="batch"):
        if num_active < 1:
            raise ValueError("num_active cannot be less than 1.")
        try:
            pool_size = len(emitter_pool)
        except TypeError as exc:
            raise TypeError("`emitter_pool` must be a list of emitter objects.") from exc
        if pool_size < num_active:
            raise ValueError(f"The emitter pool must contain at least num_active emitters, but only {pool_size} "
                             "are given.")
        if len({id(e) for e in emitter_pool}) != pool_size:
            raise ValueError("Not all emitters passed in were unique (i.e. some emitters had the same id). If "
                             "emitters were created with something like [EmitterClass(...)] * n, instead use "
                             "[EmitterClass(...) for _ in range(n)] so that all emitters are unique instances.")
        self._solution_dim = emitter_pool[0].solution_dim
        for k, emitter in enumerate(emitter_pool[1:]):
            if emitter.solution_dim != self._solution_dim:
                raise ValueError(f"All emitters must have the same solution dim, but Emitter {k} has dimension "
                                 f"{emitter.solution_dim}, while Emitter 0 has dimension {self._solution_dim}")
        if reselect not in ("terminated", "all"):
            raise ValueError(f"add_mode must either be 'terminated' or 'all', but it was '{reselect}'")
        if add_mode not in ("single", "batch"):
            raise ValueError(f"add_mode must either be 'batch' or 'single', but it was '{add_mode}'")
        if archive is result_archive:
            raise ValueError("`archive` has same id as `result_archive` -- Note that `BanditScheduler.result_archive` "
                             "already defaults to be the same as `archive` if you pass `result_archive=None`")
        pool = np.empty(pool_size, dtype=object)
        for k, emitter in enumerate(emitter_pool):
            pool[k] = emitter
        self._pool = pool
        self._num_active = int(num_active)
        self._reselect = reselect
        self._zeta = zeta
        self._active = np.zeros(pool_size, dtype=bool)
        self._success = np.zeros(pool_size, dtype=float)    # solutions of the emitter that entered the archive
        self._selection = np.zeros(pool_size, dtype=float)  # solutions the emitter was asked for
        self._restarts = np.zeros(pool_size, dtype=int)
        self._num_emitted = np.full(pool_size, None, dtype=object)
        self._cur_solutions = []
        self._last_called = None
        # insertion (batch / single, result_archive with a shared index) is the plain scheduler's
        self._inserter = Scheduler(archive, list(emitter_pool), result_archive, add_mode=add_mode, pool_emitters=False)

    # ---- properties ---------------------------------------------------------------------------------
    @property
    def archive(self):
        return self._inserter.archive

    @property
    def emitter_pool(self):
        return self._pool

    @property
    def active(self):
        return arr_readonly(self._active, view=True)

    @property
    def result_archive(self):
        return self._inserter.result_archive

    def ask_dqd(self):
        raise NotImplementedError("ask_dqd() is not supported by BanditScheduler.")

    def tell_dqd(self, objective, measures, jacobian, **fields):
        raise NotImplementedError("tell_dqd() is not supported by BanditScheduler.")

    # ---- selection ------------------------------------------------------------------------------------
    def _choose_active(self):
        """_bandit_scheduler.py:248-302: which emitters give up their slot, then UCB1 fills the slots."""
        if self._reselect == "terminated":
            # an emitter has terminated when its restart counter moved; one without a counter restarts every time
            counters = np.array([e.restarts if hasattr(e, "restarts") else -1 for e in self._pool])
            release = counters > self._restarts
            release[counters < 0] = True
            self._restarts = counters
        else:
            release = self._active.copy()
        # top up to num_active with the first inactive emitters of the pool (the first iterations)
        missing = self._num_active - self._active.sum()
        k = 0
        while missing > 0:
            release[k] = False
            if not self._active[k]:
                self._active[k] = True
                missing -= 1
            k += 1
        self._active[release] = False
        if not release.any():
            return
        # UCB1 score; never-selected emitters score +inf. (Object array, as the reference's full_like of its
        # emitter array is: the tie order of argsort then matches it.)
        score = np.full(len(self._pool), np.inf, dtype=object)
        seen = self._selection != 0
        if seen.any():
            score[seen] = (self._success[seen] / self._selection[seen] +
                           self._zeta * np.sqrt(np.log(self._success.sum()) / self._selection[seen]))
        n_on = self._active.sum()
        for k in np.argsort(score)[::-1]:
            if n_on >= self._num_active:
                break
            if not self._active[k]:
                self._active[k] = True
                n_on += 1

    def ask(self):
        if self._last_called == "ask":
            raise RuntimeError("ask cannot be called immediately after " + self._last_called)
        self._last_called = "ask"
        self._choose_active()
        batches = []
        for k in np.flatnonzero(self._active):
            sols = self._pool[k].ask()
            batches.append(sols)
            self._num_emitted[k] = len(sols)
        self._cur_solutions = (np.concatenate(batches, axis=0) if batches
                               else np.empty((0, self._solution_dim)))
        return self._cur_solutions

    def tell(self, objective, measures, **fields):
        if self._last_called != "ask":
            raise RuntimeError("tell() was called without calling ask().")
        self._last_called = "tell"
        data = {"solution": self._cur_solutions, "objective": np.asarray(objective), "measures": np.asarray(measures),
                **{name: np.asarray(v) for name, v in fields.items()}}
        for name, arr in data.items():
            if len(arr) != len(self._cur_solutions):
                raise ValueError(f"{name} should have length {len(self._cur_solutions)} (this is the number of "
                                 f"solutions output by ask()) but has length {len(arr)}")
        add_info = self._inserter._add_to_archives(data)
        add_info = {name: (v.cpu().numpy() if hasattr(v, "cpu") else np.asarray(v)) for name, v in add_info.items()}
        begin = 0
        for k in np.flatnonzero(self._active):
            end = begin + self._num_emitted[k]
            self._selection[k] += self._num_emitted[k]
            self._success[k] += np.count_nonzero(add_info["status"][begin:end])  # :420-421
            self._pool[k].tell(**{name: arr[begin:end] for name, arr in data.items()},
                               add_info={name: arr[begin:end] for name, arr in add_info.items()})
            begin = end
