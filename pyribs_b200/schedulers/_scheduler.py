"""Minimal ask/tell Scheduler with the reference's call contract
(ribs/schedulers/_scheduler.py:180-213 ask, :250-296 _add_to_archives, :361-412 tell).

The reference Scheduler works unchanged on top of these archives and emitters (they keep its
API); this one exists so that examples, tests and benchmarks of this repo run where the
reference package is not installed.
"""
from __future__ import annotations

import warnings

import numpy as np


class Scheduler:
    def __init__(self, archive, emitters, result_archive=None, *, add_mode="batch", pool_emitters=True):
        if len(emitters) == 0:
            raise ValueError("Pass in at least one emitter to the scheduler.")
        if add_mode not in ("single", "batch"):
            raise ValueError("add_mode must either be 'batch' or 'single', but it was '{add_mode}'")
        if archive is result_archive:
            raise ValueError("`archive` has same id as `result_archive` -- note that `Scheduler.result_archive` "
                             "already defaults to be the same as `archive` if you pass `result_archive=None`")
        self._archive = archive
        self._emitters = list(emitters)
        self._result_archive = result_archive
        self._add_mode = add_mode
        self._last_called = None
        self._cur_solutions = None
        self._num_emitted = [None] * len(self._emitters)
        # fast path: homogeneous CMA-ES emitters run as one batched pool --
        # same results, one launch sequence per ask / tell instead of one per emitter
        self._pool = None
        if pool_emitters and add_mode == "batch":
            from pyribs_b200.emitters._emitter_pool import EmitterPool

            self._pool = EmitterPool.try_build(self._emitters, archive)

    @property
    def archive(self):
        return self._archive

    @property
    def emitters(self):
        return self._emitters

    @property
    def result_archive(self):
        return self._archive if self._result_archive is None else self._result_archive

    def ask(self, return_device=False):
        """`return_device=True` (pooled emitters only) hands out the CUDA tensor of the samples, for
        evaluators that run on the GPU; `tell` then also accepts CUDA tensors."""
        if self._last_called == "ask":
            raise RuntimeError("ask cannot be called immediately after ask")
        self._last_called = "ask"
        if self._pool is not None:
            self._cur_solutions = self._pool.ask(return_device)
            self._num_emitted = [self._pool.lam] * len(self._emitters)
            return self._cur_solutions
        if return_device:
            raise ValueError("return_device=True needs pooled emitters (homogeneous CMA-ES emitters, add_mode='batch')")
        sols = []
        for i, emitter in enumerate(self._emitters):
            s = emitter.ask()
            sols.append(s)
            self._num_emitted[i] = len(s)
        self._cur_solutions = np.concatenate(sols, axis=0)
        return self._cur_solutions

    def _add_to_archives(self, data):
        if self._add_mode == "batch":
            add_info = self._archive.add(**data)
            if self._result_archive is not None:
                # both archives index the same measures: when their geometry is identical the cell indices
                # computed for `archive` are handed to `result_archive` (one index_of for the double insert)
                last = getattr(self._archive, "_last_idx", None)
                if last is not None and hasattr(self._result_archive, "same_geometry") and \
                        self._result_archive.same_geometry(self._archive):
                    self._result_archive._shared_idx = last
                self._result_archive.add(**data)
        else:
            infos = []
            for i in range(len(data["solution"])):
                row = {k: v[i] for k, v in data.items()}
                infos.append(self._archive.add_single(**row))
                if self._result_archive is not None:
                    self._result_archive.add_single(**row)
            add_info = {k: np.asarray([d[k] for d in infos]) for k in infos[0]} if infos else {}
        if self._archive.empty:
            warnings.warn("`archive` is empty after adding solutions, and there is no way to recover from this.",
                          stacklevel=3)
        return add_info

    def tell(self, objective, measures, **fields):
        if self._last_called != "ask":
            raise RuntimeError("tell() was called without calling ask().")
        self._last_called = "tell"
        def host_or_device(v):
            return v if hasattr(v, "is_cuda") else np.asarray(v)

        data = {"solution": self._cur_solutions, "objective": host_or_device(objective),
                "measures": host_or_device(measures), **{k: host_or_device(v) for k, v in fields.items()}}
        for name, arr in data.items():
            if len(arr) != len(self._cur_solutions):
                raise ValueError(f"{name} should have length {len(self._cur_solutions)} (this is the number of "
                                 f"solutions output by ask()) but has length {len(arr)}")
        if self._pool is not None:
            # the samples are still on the device: insert from there, rank on the host, update every strategy at once
            data["solution"] = self._pool.solutions_device()
            add_info = self._add_to_archives(data)
            host_info = {k: (v.cpu().numpy() if hasattr(v, "cpu") else np.asarray(v)) for k, v in add_info.items()}
            self._pool.tell(data["objective"], host_info)
            return
        add_info = self._add_to_archives(data)
        pos = 0
        for emitter, n in zip(self._emitters, self._num_emitted):
            end = pos + n
            emitter.tell(**{k: v[pos:end] for k, v in data.items()},
                         add_info={k: v[pos:end] for k, v in add_info.items()})
            pos = end
