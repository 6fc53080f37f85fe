"""Pooled execution of many EvolutionStrategyEmitters.

The reference's Scheduler loops over its emitters in Python: `emitter.ask()` for each
(ribs/schedulers/_scheduler.py:202-205), then `emitter.tell(...)` for each (:403-412). With 15-100
small CMA-ES emitters (BASELINE.json configs[0] and configs[2]) that loop is launch latency and host
overhead. Here the strategies of all emitters are stacked in HBM (`CMAPool`) and one batched launch
sequence serves all of them (csrc/es.cu `qd_cma_pool_*`), while `EmitterPool` applies the emitter
policy of `EvolutionStrategyEmitter.tell` (ribs/emitters/_evolution_strategy_emitter.py:203-265:
ranking, parent selection, restart rule) to all emitters at once with vectorised NumPy.

Nothing about the results changes: every strategy keeps its own Philox stream (seed, counter), the
batched kernels repeat the single-strategy arithmetic operation for operation, ranking uses the same
NumPy sorts row by row, and restarts draw from the archive in emitter order -- a pooled run and an
emitter-by-emitter run produce the same archive (tests/test_gpu_es.py::test_pooled_scheduler_*).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from pyribs_b200 import _lib
from pyribs_b200._utils import arr_readonly
from pyribs_b200.emitters.opt._cma_es import CMAEvolutionStrategy
from pyribs_b200.emitters.opt._es_base import es_weights, stream_ptr
from pyribs_b200.emitters.rankers import (ImprovementRanker, ObjectiveRanker, TwoStageImprovementRanker,
                                          TwoStageObjectiveRanker)

_SC_DTYPE = np.dtype([(k, np.float64) for k in ("cc", "cs", "c1", "cmu", "mueff", "damps", "wsum", "hsig_exp")])


class CMAPool:
    """Stacked device state of E CMA-ES strategies with identical (solution_dim, batch_size).

    Adopting a strategy rebinds its tensors to views of the stacks, so its own methods (`reset`,
    `set_eigensystem`, the state properties) keep working on the pooled memory, and moves it onto the
    pool's stream. `ask` / `tell` of an adopted strategy route through the pool."""

    def __init__(self, strategies):
        es0 = strategies[0]
        self.strategies = list(strategies)
        self.E, self.n, self.lam = len(strategies), es0.solution_dim, es0.batch_size
        self._device = es0._device
        self._L = _lib.lib()
        E, n, lam = self.E, self.n, self.lam
        f64 = dict(dtype=torch.float64, device=self._device)
        with torch.cuda.device(self._device):
            self._stream = torch.cuda.Stream(device=self._device)
            self._event = torch.cuda.Event()
            self.ws_stride = (n + 8 + 7) // 8 * 8
            self.mean, self.pc, self.ps = (torch.zeros((E, n), **f64) for _ in range(3))
            self.C, self.B, self.T, self.invsqrt = (torch.zeros((E, n, n), **f64) for _ in range(4))
            self.eig = torch.ones((E, n), **f64)
            self.status = torch.zeros((E, 8), **f64)
            self.new_mean = torch.zeros((E, n), **f64)
            self.ws = torch.zeros((E, self.ws_stride), **f64)
            self.X = torch.zeros((E, lam, n), **f64)
            self.z = torch.zeros((E, lam, n), **f64)
            self._host = torch.empty((E * lam, n), dtype=torch.float64).pin_memory()
            self._seeds = torch.from_numpy(np.array([es._seed for es in strategies], dtype=np.uint64).view(np.int64)).to(
                self._device)
            # host -> device staging of one tell: ids | mu | scalars | weights | ranks
            self._off_mu = _align(4 * E)
            self._off_sc = self._off_mu + 8 * E
            self._off_w = self._off_sc + _SC_DTYPE.itemsize * E
            self._off_rank = self._off_w + 8 * E * lam
            nbytes = self._off_rank + 8 * E * lam
            self._stage_h = torch.zeros(nbytes, dtype=torch.uint8).pin_memory()
            self._stage_d = torch.zeros(nbytes, dtype=torch.uint8, device=self._device)
            self._stage_np = self._stage_h.numpy()
            self._off_h = torch.zeros(E, dtype=torch.int64).pin_memory()
            self._off_d = torch.zeros(E, dtype=torch.int64, device=self._device)
            for e, es in enumerate(strategies):
                self._adopt(e, es)
            self._stream.wait_stream(torch.cuda.current_stream())
        self._abi = _lib.CmaPool(E=E, n=n, lam=lam, ws_stride=self.ws_stride,
                                 **{k: getattr(self, k).data_ptr() for k in ("mean", "pc", "ps", "C", "B", "eig", "T",
                                                                             "invsqrt", "status", "new_mean", "ws",
                                                                             "X", "z")})
        self._fresh = set()  # slots whose samples of the current round were not handed out yet
        self._host_valid = False
        self._params = {}

    def _adopt(self, e, es):
        es._cancel_prefetch()
        if es._stream is not None:
            es._stream.synchronize()
        for name in ("mean", "pc", "ps", "C", "B", "eig", "T", "invsqrt", "status"):
            stack = getattr(self, name)
            stack[e].copy_(getattr(es, "_" + name))
            setattr(es, "_" + name, stack[e])
        es._new_mean = self.new_mean[e]
        es._ws = self.ws[e]
        es._solutions = self.X[e]
        es._z = self.z[e]
        es._stream, es._ask_event = self._stream, self._event
        es._pool, es._slot = self, e

    # ---- ask ------------------------------------------------------------------------------------
    def ask_all(self, z=None):
        """Samples lambda solutions for every strategy: one refresh sequence for the strategies whose
        eigensystem is stale (_cma_es.py:51-82), one Philox fill, one batched GEMM, one D2H copy."""
        S = self.strategies
        # X is about to be overwritten: get behind whatever the caller's stream still does with the last round's
        self._stream.wait_stream(torch.cuda.current_stream(self._device))
        with torch.cuda.device(self._device), torch.cuda.stream(self._stream):
            stale = [e for e, es in enumerate(S) if es.current_eval > es.updated_eval + es.lazy_gap_evals]
            if stale:
                ids = torch.tensor(stale, dtype=torch.int32).to(self._device, non_blocking=False)
                _lib.check(self._L.qd_cma_pool_refresh(C.byref(self._abi), ids.data_ptr(), len(stale), stream_ptr()))
                for e in stale:
                    S[e].updated_eval = S[e].current_eval
            if z is None:
                self._off_h.copy_(torch.from_numpy(np.array([es._rng_offset for es in S], dtype=np.int64)))
                self._off_d.copy_(self._off_h, non_blocking=True)
                _lib.check(self._L.qd_cma_pool_ask(C.byref(self._abi), self._seeds.data_ptr(), self._off_d.data_ptr(),
                                                   stream_ptr()))
                adv = (self.lam * self.n + 1) // 2
                for es in S:
                    es._rng_offset += adv
            else:
                self.z.copy_(torch.as_tensor(np.asarray(z, dtype=np.float64)).reshape(self.E, self.lam, self.n))
                _lib.check(self._L.qd_cma_pool_ask(C.byref(self._abi), None, None, stream_ptr()))
            self._event.record(self._stream)
        self._host_valid = False
        self._fresh = set(range(self.E))

    def _sync_host(self):
        if not self._host_valid:
            with torch.cuda.device(self._device), torch.cuda.stream(self._stream):
                self._host.copy_(self.X.view(self.E * self.lam, self.n), non_blocking=True)
                self._event.record(self._stream)
            self._host_valid = True
        self._event.synchronize()

    def solutions_host(self):
        self._sync_host()
        return arr_readonly(self._host.numpy(), view=True)

    def solutions_device(self):
        torch.cuda.current_stream().wait_event(self._event)
        return self.X.view(self.E * self.lam, self.n)

    def ask_one(self, e, return_device=False):
        """`strategy.ask()` of an adopted strategy (a scheduler that asks emitter by emitter): the first
        call of a round samples for everyone."""
        if e not in self._fresh:
            self.ask_all()
        self._fresh.discard(e)
        if return_device:
            torch.cuda.current_stream().wait_event(self._event)
            return self.X[e]
        self._sync_host()
        out = self._host.numpy()[e * self.lam:(e + 1) * self.lam].copy()
        es = self.strategies[e]
        if np.dtype(es.dtype) != np.float64:
            out = out.astype(es.dtype)
        return arr_readonly(out)

    # ---- tell -----------------------------------------------------------------------------------
    def _strat_params(self, mu):
        p = self._params.get(mu)
        if p is None:
            w, mueff, cc, cs, c1, cmu = self.strategies[0]._calc_strat_params(mu)
            damps = 1 + 2 * max(0, np.sqrt((mueff - 1) / (self.n + 1)) - 1) + cs
            p = self._params[mu] = (w, mueff, cc, cs, c1, cmu, damps, np.sum(w))
        return p

    def tell_many(self, slots, ranking, num_parents):
        """_cma_es.py:274-328 for the strategies `slots`; `ranking[i]` / `num_parents[i]` belong to slots[i]."""
        E, lam = self.E, self.lam
        st = self._stage_np
        ids = st[: 4 * E].view(np.int32)
        mu = st[self._off_mu: self._off_mu + 8 * E].view(np.int64)
        sc = st[self._off_sc: self._off_sc + _SC_DTYPE.itemsize * E].view(_SC_DTYPE)
        w = st[self._off_w: self._off_w + 8 * E * lam].view(np.float64).reshape(E, lam)
        rank = st[self._off_rank: self._off_rank + 8 * E * lam].view(np.int64).reshape(E, lam)
        n_act = 0
        for i, e in enumerate(slots):
            es = self.strategies[e]
            r = np.asarray(ranking[i])
            es.current_eval += len(r)
            m = int(num_parents[i])
            if m == 0:
                continue
            wv, mueff, cc, cs, c1, cmu, damps, wsum = self._strat_params(m)
            ids[n_act] = e
            mu[e] = m
            sc[e] = (cc, cs, c1, cmu, mueff, damps, wsum, 2 * es.current_eval / es.batch_size)
            w[e, :m] = wv
            rank[e, :m] = r[:m]
            n_act += 1
        if n_act == 0:
            return
        with torch.cuda.device(self._device), torch.cuda.stream(self._stream):
            self._stage_d.copy_(self._stage_h, non_blocking=True)
            base = self._stage_d.data_ptr()
            _lib.check(self._L.qd_cma_pool_tell(C.byref(self._abi), base, n_act, base + self._off_rank, base + self._off_w,
                                                base + self._off_mu, base + self._off_sc, stream_ptr()))
            self._event.record(self._stream)
        # the staging buffer is rewritten by the next tell: wait until this copy has been consumed
        self._event.synchronize()

    def status_host(self):
        with torch.cuda.stream(self._stream):
            return self.status.cpu().numpy()


def _align(x, a=64):
    return (x + a - 1) // a * a


_POOLABLE_RANKERS = (ImprovementRanker, TwoStageImprovementRanker, ObjectiveRanker, TwoStageObjectiveRanker)


class EmitterPool:
    """The emitter policy of EvolutionStrategyEmitter.tell for all emitters at once."""

    def __init__(self, emitters, archive):
        self.emitters = list(emitters)
        self.archive = archive
        self.cma = CMAPool([em._opt for em in emitters])
        em0 = emitters[0]
        self.E, self.lam = len(emitters), em0.batch_size
        self._two_stage = isinstance(em0._ranker, (TwoStageImprovementRanker, TwoStageObjectiveRanker))
        self._by_objective = isinstance(em0._ranker, (ObjectiveRanker, TwoStageObjectiveRanker))
        self._filter = em0._selection_rule == "filter"

    @staticmethod
    def try_build(emitters, archive):
        """Returns a pool when every emitter is a plain, unbounded CMA-ES EvolutionStrategyEmitter of the same
        shape and policy on `archive`; None otherwise (the scheduler then works emitter by emitter)."""
        from pyribs_b200.emitters._evolution_strategy_emitter import EvolutionStrategyEmitter

        if len(emitters) < 2:
            return None
        em0 = emitters[0]
        for em in emitters:
            if type(em) is not EvolutionStrategyEmitter or type(em._opt) is not CMAEvolutionStrategy:
                return None
            if em.archive is not archive or em._opt._bounded or em._opt._pool is not None:
                return None
            if (em.batch_size, em.solution_dim) != (em0.batch_size, em0.solution_dim):
                return None
            if type(em._ranker) is not type(em0._ranker) or type(em._ranker) not in _POOLABLE_RANKERS:
                return None
            if em._selection_rule != em0._selection_rule or em._restart_rule != em0._restart_rule:
                return None
        if not isinstance(em0.solution_dim, (int, np.integer)) or em0.solution_dim > 150:
            return None  # the batched Jacobi eigensolver holds one matrix per CTA
        if np.dtype(archive.dtypes["solution"]) != np.float64:
            return None
        return EmitterPool(emitters, archive)

    def ask(self, return_device=False):
        self.cma.ask_all()
        self.cma._fresh.clear()
        return self.cma.solutions_device() if return_device else self.cma.solutions_host()

    def solutions_device(self):
        return self.cma.solutions_device()

    def tell(self, objective, add_info):
        """objective (E*lam,), add_info {"status", "value"} (E*lam,) host arrays."""
        E, lam = self.E, self.lam
        status = np.asarray(add_info["status"]).reshape(E, lam)
        if self._by_objective:
            key = (objective.cpu().numpy() if hasattr(objective, "cpu") else np.asarray(objective)).reshape(E, lam)
        else:
            key = np.asarray(add_info["value"]).reshape(E, lam)
        if self._two_stage:  # np.flip(np.lexsort((values, status))) row by row (rankers.py:168-203)
            idx = np.flip(np.lexsort((key, status), axis=-1), axis=1)
        else:
            idx = np.flip(np.argsort(key, axis=1), axis=1)
        new_sols = status.astype(bool).sum(axis=1)
        num_parents = new_sols if self._filter else np.full(E, lam // 2)
        self.cma.tell_many(range(E), idx, num_parents)
        # check_stop of every strategy from one status read (_cma_es.py:161-179)
        st = self.cma.status_host()
        first, last = idx[:, 0], idx[:, -1]
        rows = np.arange(E)
        dv = key[rows, first] - key[rows, last]
        if self._two_stage:
            flat = np.hypot(status[rows, first].astype(np.float64) - status[rows, last], dv) < 1e-12
        else:
            flat = np.abs(dv) < 1e-12
        if lam < 2:
            flat[:] = False
        with np.errstate(divide="ignore", invalid="ignore"):
            cond = st[:, 2] / st[:, 3]
        stop = (cond > 1e14) | (st[:, 0] * np.sqrt(st[:, 2]) < 1e-11) | flat
        for e, em in enumerate(self.emitters):
            em._itrs += 1
            em._opt.condition_number = float(cond[e])
            if stop[e] or em._check_restart(new_sols[e]):
                new_x0 = self.archive.sample_elites(1)["solution"][0]
                em._opt.reset(new_x0)
                em._ranker.reset(em, self.archive)
                em._restarts += 1
