"""Rankers used by EvolutionStrategyEmitter (reference ribs/emitters/rankers.py:136-398).

Ranking a batch of a few thousand scalars is a host-side sort in the reference; it stays one
here (status / value come back from `add` as NumPy arrays anyway), which also makes the tie
order identical to the reference by construction: one-key rankers are
`np.flip(np.argsort(v))`, two-stage rankers `np.flip(np.lexsort((v, status)))`.
"""
from __future__ import annotations

import numpy as np


class RankerBase:
    """`rank(emitter, archive, data, add_info) -> (indices, ranking_values)`; `reset` is optional."""

    def __init__(self, seed=None):
        self._rng = np.random.default_rng(seed)

    def rank(self, emitter, archive, data, add_info):
        raise NotImplementedError

    def reset(self, emitter, archive):
        pass


def _descending(values):
    return np.flip(np.argsort(values))


def _descending_two_stage(status, values):
    keys = np.stack((status, values), axis=-1)
    # lexsort sorts by the LAST key first: (values, status) -> status is the primary key
    return np.flip(np.lexsort((keys[:, 1], keys[:, 0]))), keys


class ImprovementRanker(RankerBase):
    """rankers.py:136-159: by improvement value."""

    def rank(self, emitter, archive, data, add_info):
        return _descending(add_info["value"]), add_info["value"]


class TwoStageImprovementRanker(RankerBase):
    """rankers.py:168-203: by status (new > improved > not added), then improvement value."""

    def rank(self, emitter, archive, data, add_info):
        return _descending_two_stage(add_info["status"], add_info["value"])


class _RandomDirectionMixin:
    _target_measure_dir = None

    @property
    def target_measure_dir(self):
        return self._target_measure_dir

    @target_measure_dir.setter
    def target_measure_dir(self, value):
        self._target_measure_dir = value

    def reset(self, emitter, archive):
        """rankers.py:258-262: Gaussian direction scaled to the archive's measure ranges."""
        ranges = archive.upper_bounds - archive.lower_bounds
        self._target_measure_dir = self._rng.standard_normal(len(ranges)) * ranges

    def _projections(self, data):
        if self._target_measure_dir is None:
            raise RuntimeError("target measure direction not set")
        return np.dot(data["measures"], self._target_measure_dir)


class RandomDirectionRanker(_RandomDirectionMixin, RankerBase):
    """rankers.py:212-262."""

    def rank(self, emitter, archive, data, add_info):
        proj = self._projections(data)
        return _descending(proj), proj


class TwoStageRandomDirectionRanker(_RandomDirectionMixin, RankerBase):
    """rankers.py:271-338."""

    def rank(self, emitter, archive, data, add_info):
        return _descending_two_stage(add_info["status"], self._projections(data))


class ObjectiveRanker(RankerBase):
    """rankers.py:341-361."""

    def rank(self, emitter, archive, data, add_info):
        return _descending(data["objective"]), data["objective"]


class TwoStageObjectiveRanker(RankerBase):
    """rankers.py:364-398."""

    def rank(self, emitter, archive, data, add_info):
        return _descending_two_stage(add_info["status"], data["objective"])


_NAME_TO_RANKER_MAP = {
    "ImprovementRanker": ImprovementRanker,
    "TwoStageImprovementRanker": TwoStageImprovementRanker,
    "RandomDirectionRanker": RandomDirectionRanker,
    "TwoStageRandomDirectionRanker": TwoStageRandomDirectionRanker,
    "ObjectiveRanker": ObjectiveRanker,
    "TwoStageObjectiveRanker": TwoStageObjectiveRanker,
    "imp": ImprovementRanker,
    "2imp": TwoStageImprovementRanker,
    "rd": RandomDirectionRanker,
    "2rd": TwoStageRandomDirectionRanker,
    "obj": ObjectiveRanker,
    "2obj": TwoStageObjectiveRanker,
}


def _get_ranker(klass, seed):
    """rankers.py:755-785."""
    if isinstance(klass, str):
        if klass in _NAME_TO_RANKER_MAP:
            return _NAME_TO_RANKER_MAP[klass](seed)
        raise ValueError(f"`{klass}` is not the full or abbreviated name of a valid ranker")
    if callable(klass):
        ranker = klass(seed)
        if isinstance(ranker, RankerBase) or hasattr(ranker, "rank"):
            return ranker
        raise ValueError(f"Callable `{klass}` did not return an instance of RankerBase.")
    raise ValueError(f"`{klass}` is neither a callable nor a string")
