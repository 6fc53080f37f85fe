"""Emitters of the hot path (drop-in for `ribs.emitters`): the CMA-family EvolutionStrategyEmitter and the
MAP-Elites variation emitters."""
from pyribs_b200.emitters._evolution_strategy_emitter import EvolutionStrategyEmitter
from pyribs_b200.emitters._variation_emitters import GaussianEmitter, IsoLineEmitter

__all__ = ["EvolutionStrategyEmitter", "GaussianEmitter", "IsoLineEmitter"]
