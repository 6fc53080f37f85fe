from pyribs_b200.schedulers._bandit_scheduler import BanditScheduler
from pyribs_b200.schedulers._scheduler import Scheduler

__all__ = ["BanditScheduler", "Scheduler"]
