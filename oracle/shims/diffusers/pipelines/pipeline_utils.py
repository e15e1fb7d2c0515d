import torch
from ..configuration_utils import ConfigMixin


class DiffusionPipeline(ConfigMixin):
    def register_modules(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def device(self):
        for m in (getattr(self, "unet", None), getattr(self, "vae", None)):
            if isinstance(m, torch.nn.Module):
                return next(m.parameters()).device
        return torch.device("cpu")

    @property
    def _execution_device(self):
        return self.device

    def to(self, device):
        return self
