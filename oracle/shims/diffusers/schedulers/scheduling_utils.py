from enum import Enum


class KarrasDiffusionSchedulers(Enum):
    DDIMScheduler = 1
    DDPMScheduler = 2


class SchedulerMixin:
    pass
