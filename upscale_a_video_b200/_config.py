"""Minimal config plumbing mirroring what callers use from diffusers' ConfigMixin
(`from_config(json-dict)`, `.config.<key>`, `register_to_config`): inference_upscale_a_video.py:104-121."""
from __future__ import annotations

import inspect
import json
import os


class FrozenConfig(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class ConfigMixin:
    def _init_config(self, locals_: dict):
        sig = inspect.signature(type(self).__init__).parameters
        cfg = {k: locals_[k] for k in sig if k not in ("self", "args", "kwargs") and k in locals_}
        object.__setattr__(self, "_internal_dict", FrozenConfig(cfg))

    def register_to_config(self, **kwargs):
        cfg = dict(getattr(self, "_internal_dict", {}))
        cfg.update(kwargs)
        object.__setattr__(self, "_internal_dict", FrozenConfig(cfg))

    @property
    def config(self):
        return self._internal_dict

    @classmethod
    def load_config(cls, path):
        if os.path.isdir(path):
            path = os.path.join(path, "config.json")
        with open(path) as f:
            return json.load(f)

    @classmethod
    def from_config(cls, config, **kwargs):
        if isinstance(config, (str, os.PathLike)):
            config = cls.load_config(config)
        sig = inspect.signature(cls.__init__).parameters
        init = {k: v for k, v in dict(config).items() if k in sig and not k.startswith("_")}
        init.update({k: v for k, v in kwargs.items() if k in sig})
        return cls(**init)
