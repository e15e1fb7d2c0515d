"""color_oracle.py — CPU restatement of the post-decode colour fix and output packing
(/root/reference/models_video/color_correction.py and /root/reference/inference_upscale_a_video.py:323-357).

TEST INFRASTRUCTURE.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU legs may import this file; the
product package (`upscale_a_video_b200/`) never does.

Parity status: the reference has no tests for this path; the restatement is pinned against the UNMODIFIED reference
functions run in the build container by `oracle/make_golden_color.py` (fixtures: `tests/golden/color.pt`, checked on CPU by
`tests/test_oracle_golden.py`).  Everything is written with explicit arithmetic (no F.conv2d / F.interpolate /
Tensor.var) so that it documents the exact op order the CUDA kernels replay.
"""
from __future__ import annotations

import torch


def calc_mean_std(feat: torch.Tensor, eps: float = 1e-5):
    """color_correction.py:45-58: var is UNBIASED (Tensor.var default), eps added to the variance"""
    assert feat.dim() == 4, "The input feature should be 4D tensor."
    b, c = feat.shape[:2]
    x = feat.reshape(b, c, -1).double()
    n = x.shape[-1]
    mean = x.sum(-1) / n
    var = ((x * x).sum(-1) - x.sum(-1) * mean) / (n - 1)
    var = var.clamp_min(0).float() + eps
    return mean.float().reshape(b, c, 1, 1), var.sqrt().reshape(b, c, 1, 1)


def adaptive_instance_normalization(content: torch.Tensor, style: torch.Tensor) -> torch.Tensor:
    """color_correction.py:60-73"""
    s_mean, s_std = calc_mean_std(style)
    c_mean, c_std = calc_mean_std(content)
    normalized = (content - c_mean) / c_std
    return normalized * s_std + s_mean


_K = ((0.0625, 0.125, 0.0625), (0.125, 0.25, 0.125), (0.0625, 0.125, 0.0625))


def wavelet_blur(image: torch.Tensor, radius: int) -> torch.Tensor:
    """color_correction.py:75-93: replicate pad by `radius`, depthwise 3x3 with dilation `radius`.
    Accumulated row-major from zero (the weights are powers of two: every product is exact)."""
    H, W = image.shape[-2:]
    ys = torch.arange(H)
    xs = torch.arange(W)
    acc = torch.zeros_like(image)
    for ky in range(3):
        yy = (ys + (ky - 1) * radius).clamp(0, H - 1)
        for kx in range(3):
            xx = (xs + (kx - 1) * radius).clamp(0, W - 1)
            acc = acc + _K[ky][kx] * image[..., yy, :][..., :, xx]
    return acc


def wavelet_decomposition(image: torch.Tensor, levels: int = 5):
    """color_correction.py:95-103"""
    high = torch.zeros_like(image)
    low = image
    for i in range(levels):
        low = wavelet_blur(image, 2 ** i)
        high = high + (image - low)
        image = low
    return high, low


def wavelet_reconstruction(content: torch.Tensor, style: torch.Tensor) -> torch.Tensor:
    """color_correction.py:105-118"""
    content_high, _ = wavelet_decomposition(content)
    _, style_low = wavelet_decomposition(style)
    return content_high + style_low


def _cubic1(x, A):
    return ((A + 2) * x - (A + 3)) * x * x + 1


def _cubic2(x, A):
    return ((A * x - 5 * A) * x + 8 * A) * x - 4 * A


def _cubic_axis(n_in: int, scale: int):
    """indices (n_out, 4) and weights (n_out, 4) of ATen's upsample_bicubic2d along one axis: align_corners=False,
    source index = (dst + 0.5) / scale - 0.5 (NOT clamped at 0 for cubic), A = -0.75, border-clamped taps"""
    dst = torch.arange(n_in * scale, dtype=torch.float32)
    real = (1.0 / scale) * (dst + 0.5) - 0.5
    fl = torch.floor(real)
    t = real - fl
    A = -0.75
    w = torch.stack([_cubic2(t + 1.0, A), _cubic1(t, A), _cubic1(1.0 - t, A), _cubic2((1.0 - t) + 1.0, A)], dim=1)
    idx = (fl.long()[:, None] + torch.arange(-1, 3)[None, :]).clamp(0, n_in - 1)
    return idx, w


def bicubic_upsample(x: torch.Tensor, scale: int = 4) -> torch.Tensor:
    """F.interpolate(x, scale_factor=scale, mode='bicubic') (inference_upscale_a_video.py:327): x interpolation of the
    four neighbouring rows first, then y, as ATen's kernel does"""
    h, w = x.shape[-2:]
    iy, wy = _cubic_axis(h, scale)
    ix, wx = _cubic_axis(w, scale)
    rows = None
    for k in range(4):  # x pass
        term = x[..., :, ix[:, k]] * wx[:, k]
        rows = term if rows is None else rows + term
    out = None
    for k in range(4):  # y pass
        term = rows[..., iy[:, k], :] * wy[:, k][:, None]
        out = term if out is None else out + term
    return out


def color_fix_frames(output: torch.Tensor, vframes: torch.Tensor, color_fix: str) -> torch.Tensor:
    """inference_upscale_a_video.py:323-333: output / vframes are (1, c, t, H, W) / (1, c, t, h, w); returns (t, c, H, W)"""
    out = output.squeeze(0).permute(1, 0, 2, 3).contiguous()
    if color_fix in ("AdaIn", "Wavelet"):
        lr = bicubic_upsample(vframes.squeeze(0).permute(1, 0, 2, 3).contiguous(), 4)
        out = adaptive_instance_normalization(out, lr) if color_fix == "AdaIn" else wavelet_reconstruction(out, lr)
    return out


def pack_video_uint8(frames: torch.Tensor) -> torch.Tensor:
    """inference_upscale_a_video.py:354-356: (t c h w) -> (t h w c) uint8; numpy astype truncates toward zero"""
    v = (frames / 2 + 0.5).clamp(0, 1) * 255
    return v.permute(0, 2, 3, 1).contiguous().to(torch.int32).to(torch.uint8)
