from pyribs_b200.schedulers._scheduler import Scheduler

__all__ = ["Scheduler"]
