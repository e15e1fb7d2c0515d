"""DDIMScheduler — drop-in for /root/reference/models_video/scheduling_ddim.py (same config keys, `set_timesteps`,
`scale_model_input`, `step_v0`, `step_vt`, `step`, `add_noise`, `init_noise_sigma`, `timesteps`).

The schedule tables live on the host exactly as in the reference (`alphas_cumprod` is a CPU fp32 tensor,
scheduling_ddim.py:163-169); every tensor update is ONE fused CUDA kernel (csrc/sampler.cu) that replays the
reference's per-op rounding, instead of 3-8 ATen pointwise kernels plus a device->host sync per step
(`self.alphas_cumprod[timestep]` with a CUDA timestep, scheduling_ddim.py:404,459)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import ops
from ._config import ConfigMixin

_PRED = {"epsilon": 0, "sample": 1, "v_prediction": 2}


@dataclass
class DDIMSchedulerOutput:
    prev_sample: Optional[torch.Tensor] = None
    pred_original_sample: Optional[torch.Tensor] = None


class DDIMScheduler(ConfigMixin):
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", trained_betas=None, clip_sample: bool = True,
                 set_alpha_to_one: bool = True, steps_offset: int = 0, prediction_type: str = "epsilon",
                 thresholding: bool = False, dynamic_thresholding_ratio: float = 0.995,
                 clip_sample_range: float = 1.0, sample_max_value: float = 1.0):
        self._init_config(locals())
        # scheduling_ddim.py:147-161
        if trained_betas is not None:
            self.betas = torch.tensor(trained_betas, dtype=torch.float32)
        elif beta_schedule == "linear":
            self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        elif beta_schedule == "squaredcos_cap_v2":
            import math
            f = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2  # noqa: E731
            self.betas = torch.tensor([min(1 - f((i + 1) / num_train_timesteps) / f(i / num_train_timesteps), 0.999)
                                       for i in range(num_train_timesteps)], dtype=torch.float32)
        else:
            raise NotImplementedError(f"{beta_schedule} does is not implemented for {self.__class__}")
        if thresholding:
            raise NotImplementedError("dynamic thresholding is not on the Upscale-A-Video sampling path")
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_timesteps(self, num_inference_steps: int, device=None):
        """scheduling_ddim.py:237-259"""
        if num_inference_steps > self.config.num_train_timesteps:
            raise ValueError(
                f"`num_inference_steps`: {num_inference_steps} cannot be larger than `self.config.train_timesteps`:"
                f" {self.config.num_train_timesteps} as the unet model trained with this scheduler can only handle"
                f" maximal {self.config.num_train_timesteps} timesteps.")
        self.num_inference_steps = num_inference_steps
        step_ratio = self.config.num_train_timesteps // self.num_inference_steps
        ts = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
        ts = ts + self.config.steps_offset
        self.timesteps_host = [int(t) for t in ts]  # no device->host sync inside the sampling loop
        self.timesteps = torch.from_numpy(ts).to(device)

    # -- host-side scalars, computed with the same fp32 CPU tensor ops as the reference --------------------
    def _coefs(self, timestep, eta: float = 0.0):
        t = int(timestep)
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        a = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b = 1 - a
        variance = ((1 - a_prev) / (1 - a)) * (1 - a / a_prev)
        std = eta * variance ** 0.5
        return dict(sa=float(a ** 0.5), sb=float(b ** 0.5), sa_prev=float(a_prev ** 0.5),
                    dir=float((1 - a_prev - std ** 2) ** 0.5), std=float(std))

    def _check(self):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        if self.config.prediction_type not in _PRED:
            raise ValueError(f"prediction_type given as {self.config.prediction_type} must be one of `epsilon`, `sample`, or"
                             " `v_prediction`")

    def step_v0(self, model_output, timestep, sample, eta: float = 0.0, use_clipped_model_output: bool = False,
                generator=None, variance_noise=None, return_dict: bool = True):
        """scheduling_ddim.py:383-433 (pred_original_sample only)"""
        self._check()
        c = self._coefs(timestep)
        x0 = ops.ddim_step_v0(model_output.contiguous(), sample.contiguous(), _PRED[self.config.prediction_type],
                              c["sa"], c["sb"], self.config.clip_sample, self.config.clip_sample_range)
        if not return_dict:
            return (x0,)
        return DDIMSchedulerOutput(pred_original_sample=x0)

    def step_vt(self, v0, model_output, timestep, sample, eta: float = 0.0, use_clipped_model_output: bool = False,
                generator=None, variance_noise=None, return_dict: bool = True):
        """scheduling_ddim.py:436-520 (x_{t-1} from a possibly edited x0)"""
        self._check()
        if use_clipped_model_output:
            raise NotImplementedError("use_clipped_model_output is not used by the Upscale-A-Video pipeline")
        c = self._coefs(timestep, eta)
        noise = None
        if eta > 0:
            if variance_noise is not None and generator is not None:
                raise ValueError("Cannot pass both generator and variance_noise. Please make sure that either `generator` or"
                                 " `variance_noise` stays `None`.")
            noise = variance_noise
            if noise is None:
                gdev = generator.device if generator is not None else model_output.device
                noise = torch.randn(model_output.shape, generator=generator, device=gdev,
                                    dtype=model_output.dtype).to(model_output.device)
            noise = noise.contiguous()
        prev = ops.ddim_step_vt(v0.contiguous(), model_output.contiguous(), sample.contiguous(),
                                _PRED[self.config.prediction_type], c["sa"], c["sb"], c["sa_prev"], c["dir"],
                                self.config.clip_sample, self.config.clip_sample_range, c["std"], noise)
        if not return_dict:
            return (prev,)
        return DDIMSchedulerOutput(prev_sample=prev)

    def step(self, model_output, timestep, sample, eta: float = 0.0, use_clipped_model_output: bool = False,
             generator=None, variance_noise=None, return_dict: bool = True):
        """scheduling_ddim.py:262-380 == step_v0 followed by step_vt"""
        x0 = self.step_v0(model_output, timestep, sample).pred_original_sample
        prev = self.step_vt(x0, model_output, timestep, sample, eta, use_clipped_model_output, generator,
                            variance_noise).prev_sample
        if not return_dict:
            return (prev,)
        return DDIMSchedulerOutput(prev_sample=prev, pred_original_sample=x0)

    def add_noise(self, original_samples, noise, timesteps):
        """scheduling_ddim.py:524-545 (also serves as the `low_res_scheduler`'s DDPMScheduler.add_noise):
        alphas_cumprod is cast to the sample dtype BEFORE ** 0.5."""
        dt = original_samples.dtype
        ac = self.alphas_cumprod.to(dtype=dt)
        ts = [int(t) for t in torch.as_tensor(timesteps).flatten().tolist()]
        x = original_samples.contiguous()
        nz = noise.contiguous()
        if len(ts) == 1:
            groups = [(x, nz, ts[0])]
            out = None
        else:
            assert len(ts) == x.shape[0], "one timestep per batch item"
            out = torch.empty_like(x)
            groups = [(x[i], nz[i], ts[i]) for i in range(len(ts))]
        res = []
        for xi, ni, t in groups:
            a = float((ac[t] ** 0.5))
            s = float(((1 - ac[t]) ** 0.5))
            res.append(ops.add_noise(xi, ni, a, s))
        if out is None:
            return res[0]
        for i, r in enumerate(res):
            out[i].copy_(r)
        return out

    def __len__(self):
        return self.config.num_train_timesteps


class DDPMScheduler(DDIMScheduler):
    """Only `add_noise` is used from the pipeline's `low_res_scheduler` (pipeline_upscale_a_video.py:548)."""

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", **kwargs):
        super().__init__(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                         beta_schedule=beta_schedule)
