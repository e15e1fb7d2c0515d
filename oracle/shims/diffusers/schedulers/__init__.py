"""DDPMScheduler.add_noise (used as `low_res_scheduler`, pipeline_upscale_a_video.py:548): same arithmetic as the
in-tree copy at models_video/scheduling_ddim.py:524-545."""
import torch


class DDPMScheduler:
    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear"):
        if beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(beta_schedule)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)

    def add_noise(self, original_samples, noise, timesteps):
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        timesteps = timesteps.to(original_samples.device)
        a = (ac[timesteps] ** 0.5).flatten()
        while len(a.shape) < len(original_samples.shape):
            a = a.unsqueeze(-1)
        s = ((1 - ac[timesteps]) ** 0.5).flatten()
        while len(s.shape) < len(original_samples.shape):
            s = s.unsqueeze(-1)
        return a * original_samples + s * noise
