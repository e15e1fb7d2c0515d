import logging as _pylogging
from collections import OrderedDict
from dataclasses import fields

import torch


class BaseOutput(OrderedDict):
    """dataclass-style output that also indexes like a tuple/dict (diffusers.utils.BaseOutput)"""

    def __post_init__(self):
        for f in fields(self):
            v = getattr(self, f.name)
            if v is not None:
                self[f.name] = v

    def __getitem__(self, k):
        if isinstance(k, str):
            return dict(self.items())[k]
        return self.to_tuple()[k]

    def to_tuple(self):
        return tuple(self[k] for k in self.keys())


class _Logging:
    @staticmethod
    def get_logger(name):
        return _pylogging.getLogger(name)

    @staticmethod
    def set_verbosity_error():
        pass


logging = _Logging()


def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    """diffusers.utils.randn_tensor: draw on the generator's device, then move"""
    device = device or torch.device("cpu")
    gen_device = generator.device if generator is not None else torch.device(device)
    return torch.randn(shape, generator=generator, device=gen_device, dtype=dtype).to(device)


def apply_forward_hook(method):
    return method


def deprecate(*args, **kwargs):
    pass


def is_accelerate_available():
    return False


def is_accelerate_version(*args, **kwargs):
    return False
