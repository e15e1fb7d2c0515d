import functools
import inspect
import json
from types import SimpleNamespace


class FrozenDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class ConfigMixin:
    config_name = None

    def register_to_config(self, **kwargs):
        cfg = dict(getattr(self, "_internal_dict", {}))
        cfg.update(kwargs)
        self._internal_dict = FrozenDict(cfg)

    @property
    def config(self):
        return self._internal_dict

    @classmethod
    def from_config(cls, config, **kwargs):
        if isinstance(config, str):
            with open(config) as f:
                config = json.load(f)
        sig = inspect.signature(cls.__init__).parameters
        init = {k: v for k, v in dict(config).items() if k in sig and not k.startswith("_")}
        init.update({k: v for k, v in kwargs.items() if k in sig})
        return cls(**init)

    @classmethod
    def load_config(cls, path, **kwargs):
        with open(path) as f:
            return json.load(f)


def register_to_config(init):
    @functools.wraps(init)
    def inner(self, *args, **kwargs):
        sig = inspect.signature(init)
        params = [p for p in sig.parameters.values()][1:]
        cfg = {p.name: p.default for p in params if p.default is not inspect.Parameter.empty}
        for p, a in zip(params, args):
            cfg[p.name] = a
        cfg.update(kwargs)
        ConfigMixin.register_to_config(self, **cfg)  # diffusers registers before running __init__
        init(self, *args, **kwargs)
    return inner
