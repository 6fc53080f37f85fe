"""Deterministic stand-ins for an archive and emitters, shared by oracle/gen_golden.py (which drives the
reference's BanditScheduler with them) and the CPU test of pyribs_b200's BanditScheduler. No GPU involved: the
scheduler under test is host policy (UCB1 over emitter statistics)."""
import numpy as np


class StubArchive:
    """`add` reports a seeded pseudo-random status per row (what the bandit's credit assignment consumes)."""

    def __init__(self, seed):
        self._rng = np.random.default_rng(seed)
        self._n = 0

    @property
    def empty(self):
        return self._n == 0

    def add(self, solution, objective, measures, **fields):
        # emitters that mark their solutions with a large first coordinate succeed more often
        p = np.clip(0.15 + 0.1 * np.asarray(solution)[:, 0], 0.0, 0.95)
        status = (self._rng.random(len(solution)) < p).astype(np.int32) * self._rng.integers(1, 3, len(solution))
        self._n += int(np.count_nonzero(status))
        return {"status": status.astype(np.int32), "value": np.asarray(objective, dtype=np.float64) - 1.0}


class StubEmitter:
    def __init__(self, tag, dim, batch, restart_every=None):
        self.tag, self.solution_dim, self.batch = tag, dim, batch
        self._restart_every = restart_every
        self.asked = 0
        self.told = []
        if restart_every is not None:
            self.restarts = 0

    def ask(self):
        self.asked += 1
        return np.full((self.batch, self.solution_dim), float(self.tag))

    def tell(self, solution, objective, measures, add_info, **fields):
        assert len(solution) == len(objective) == len(measures) == len(add_info["status"]) == self.batch
        assert np.all(solution == float(self.tag))
        self.told.append(int(np.count_nonzero(add_info["status"])))
        if self._restart_every is not None and len(self.told) % self._restart_every == 0:
            self.restarts += 1


def make_pool(case):
    if case == "restarting":  # every emitter terminates now and then; batch sizes differ
        return [StubEmitter(t, 3, 4 + t % 3, restart_every=2 + t % 4) for t in range(7)]
    if case == "no_counter":  # emitters without a `restarts` attribute are reselected every iteration
        return [StubEmitter(t, 3, 5, restart_every=(None if t % 2 else 3)) for t in range(6)]
    raise ValueError(case)


def run_bandit(scheduler_cls, case, reselect, num_active, iters=40, zeta=0.05, seed=123):
    """Returns (active masks per iteration, final success / selection counts) of a run."""
    pool = make_pool(case)
    sched = scheduler_cls(StubArchive(seed), pool, num_active=num_active, reselect=reselect, zeta=zeta)
    masks = []
    for _ in range(iters):
        sols = sched.ask()
        masks.append(np.array(sched.active, dtype=bool).copy())
        sched.tell(np.arange(len(sols), dtype=np.float64), np.zeros((len(sols), 2)))
    return np.array(masks), np.array(sched._success, dtype=np.float64), np.array(sched._selection, dtype=np.float64)


CASES = [("restarting", "terminated", 3), ("restarting", "all", 2), ("no_counter", "terminated", 2),
         ("no_counter", "all", 4)]
