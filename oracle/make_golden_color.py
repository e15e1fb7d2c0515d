"""make_golden_color.py — mint the colour-fix fixtures by running the UNMODIFIED reference functions on CPU.

TEST INFRASTRUCTURE; build container only (needs /root/reference):

    python oracle/make_golden_color.py        -> tests/golden/color.pt

Inputs are seeded synthetic frames: a smooth "decoded" clip with a colour cast and noise, and its low-resolution source.
Outputs are what `models_video/color_correction.py` and the CLI's post-processing lines
(`inference_upscale_a_video.py:323-333, 354-356`, executed here op for op) produce."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = ["/root/reference"]
import importlib.util  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_color_correction", "/root/reference/models_video/color_correction.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

OUT = os.path.join(ROOT, "tests", "golden", "color.pt")


def synth(T, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, h), torch.linspace(0, 1, w), indexing="ij")
    base = torch.stack([torch.sin(6 * xx + 2 * yy), torch.cos(5 * yy - xx), xx * yy * 2 - 1])  # (3, h, w)
    lr = torch.stack([base * (0.8 - 0.1 * t) + 0.05 * t for t in range(T)])  # (T, 3, h, w)
    lr = (lr + 0.05 * torch.randn(T, 3, h, w, generator=g)).clamp(-1, 1)
    hr = F.interpolate(lr, scale_factor=4, mode="bilinear")
    cast = torch.tensor([1.15, 0.9, 1.05]).view(1, 3, 1, 1)
    hr = hr * cast + torch.tensor([0.08, -0.05, 0.02]).view(1, 3, 1, 1) + 0.1 * torch.randn(T, 3, 4 * h, 4 * w, generator=g)
    return lr.contiguous(), hr.contiguous()  # hr deliberately leaves [-1, 1] in places (clamp path of the packing)


def main():
    cases = {}
    for name, (T, h, w, seed) in {"t2_16x24": (2, 16, 24, 3), "t1_13x19": (1, 13, 19, 5)}.items():
        lr, hr = synth(T, h, w, seed)
        # inference_upscale_a_video.py:325-331 (tensors are already "t c h w" here)
        up = F.interpolate(lr, scale_factor=4, mode="bicubic")
        adain = ref.adaptive_instance_normalization(hr, up)
        wave = ref.wavelet_reconstruction(hr, up)
        c_mean, c_std = ref.calc_mean_std(hr)
        high, low = ref.wavelet_decomposition(hr)
        blur4 = ref.wavelet_blur(hr, 4)

        def pack(x):  # :354-356
            v = (x / 2 + 0.5).clamp(0, 1) * 255
            v = v.permute(0, 2, 3, 1).contiguous()
            return torch.from_numpy(v.cpu().numpy().astype(np.uint8))

        cases[name] = {"lr": lr, "hr": hr, "bicubic": up, "adain": adain, "wavelet": wave, "mean": c_mean, "std": c_std,
                       "pack_hr": pack(hr), "pack_adain": pack(adain)}
        if T == 1:  # intermediate tensors only for the small case (fixture size)
            cases[name].update({"high": high, "low": low, "blur4": blur4})
    torch.save(cases, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
