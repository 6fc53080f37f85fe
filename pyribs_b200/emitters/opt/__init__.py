"""Evolution strategies of the hot path + the name registry (`ribs/emitters/opt/__init__.py:115-156`)."""
from pyribs_b200.emitters.opt._cma_es import CMAEvolutionStrategy
from pyribs_b200.emitters.opt._es_base import DeviceES
from pyribs_b200.emitters.opt._lm_ma_es import LMMAEvolutionStrategy
from pyribs_b200.emitters.opt._openai_es import OpenAIEvolutionStrategy
from pyribs_b200.emitters.opt._sep_cma_es import SeparableCMAEvolutionStrategy

_NAME_TO_ES_MAP = {
    "CMAEvolutionStrategy": CMAEvolutionStrategy,
    "SeparableCMAEvolutionStrategy": SeparableCMAEvolutionStrategy,
    "LMMAEvolutionStrategy": LMMAEvolutionStrategy,
    "OpenAIEvolutionStrategy": OpenAIEvolutionStrategy,
    "cma_es": CMAEvolutionStrategy,
    "sep_cma_es": SeparableCMAEvolutionStrategy,
    "lm_ma_es": LMMAEvolutionStrategy,
    "openai_es": OpenAIEvolutionStrategy,
}


def _get_es(klass, **es_kwargs):
    """Resolves `es=` exactly like the reference: a registered name or any callable that
    returns an object with the EvolutionStrategyBase interface."""
    if isinstance(klass, str):
        if klass in _NAME_TO_ES_MAP:
            return _NAME_TO_ES_MAP[klass](**es_kwargs)
        raise ValueError(f"`{klass}` is not the full or abbreviated name of a valid evolution strategy")
    if callable(klass):
        es = klass(**es_kwargs)
        for attr in ("reset", "ask", "tell", "check_stop"):
            if not hasattr(es, attr):
                raise ValueError(f"Callable `{klass}` did not return an evolution strategy.")
        return es
    raise ValueError(f"`{klass}` is neither a callable nor a string")


def register_with_reference():
    """Declares the device strategies virtual subclasses of the reference's abstract `EvolutionStrategyBase`
    (ribs/emitters/opt/_evolution_strategy_base.py:25), so that `ribs.emitters.EvolutionStrategyEmitter(...,
    es=pyribs_b200.emitters.opt.SeparableCMAEvolutionStrategy)` passes the isinstance check of the reference's
    registry (ribs/emitters/opt/__init__.py:149-155). This is the binding a maintainer adds; it
    needs `ribs` importable and is never called by this package itself."""
    from ribs.emitters.opt import EvolutionStrategyBase

    EvolutionStrategyBase.register(DeviceES)


__all__ = ["CMAEvolutionStrategy", "DeviceES", "LMMAEvolutionStrategy", "OpenAIEvolutionStrategy",
           "SeparableCMAEvolutionStrategy", "_get_es", "register_with_reference"]
