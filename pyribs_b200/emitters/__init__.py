"""Emitters of the hot path (drop-in for `ribs.emitters.EvolutionStrategyEmitter`)."""
from pyribs_b200.emitters._evolution_strategy_emitter import EvolutionStrategyEmitter

__all__ = ["EvolutionStrategyEmitter"]
