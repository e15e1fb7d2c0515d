"""GPU parity of the post-decode colour fix + packing kernels (csrc/postprocess.cu) through the
C ABI: against the committed reference fixtures (tests/golden/color.pt, minted by oracle/make_golden_color.py from the
unmodified reference functions), against the CPU oracle on other seeded shapes incl. the edge cases (frames smaller
than the largest dilation, odd sizes, a single frame), and through size-independent properties at the full 4x frame
size.  Tolerances: fp32 with a different association order of reductions / 9-tap sums -> 5e-6 absolute on [-1.3, 1.3]
data; the uint8 packing is integer work and must be bit-exact."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(autouse=True)
def _setup(uav_lib):
    torch.manual_seed(0)


def _maxerr(a, b):
    return (a.float().cpu() - b.float().cpu()).abs().max().item()


def test_color_fix_against_reference_fixtures():
    from upscale_a_video_b200 import color_correction as cc
    cases = torch.load(os.path.join(G, "color.pt"), map_location="cpu", weights_only=False)
    for name, c in cases.items():
        lr, hr = c["lr"].cuda(), c["hr"].cuda()
        up = cc.upsample_lr_frames(lr, 4)
        assert _maxerr(up, c["bicubic"]) < 2e-6, name
        mean, std = cc.calc_mean_std(hr)
        assert _maxerr(mean, c["mean"]) < 1e-6 and _maxerr(std, c["std"]) < 1e-6, name
        ref_up = c["bicubic"].cuda()
        assert _maxerr(cc.adaptive_instance_normalization(hr, ref_up), c["adain"]) < 5e-6, name
        assert _maxerr(cc.wavelet_reconstruction(hr, ref_up), c["wavelet"]) < 2e-6, name
        if "high" in c:
            high, low = cc.wavelet_decomposition(hr)
            assert _maxerr(high, c["high"]) < 2e-6 and _maxerr(low, c["low"]) < 1e-6, name
            assert _maxerr(cc.wavelet_blur(hr, 4), c["blur4"]) < 1e-6, name
        assert torch.equal(cc.pack_video_uint8(hr).cpu(), c["pack_hr"]), name
        assert torch.equal(cc.pack_video_uint8(c["adain"].cuda()).cpu(), c["pack_adain"]), name
        # the whole CLI block (inference_upscale_a_video.py:323-333) from the pipeline's (1, c, t, H, W) output
        for mode, key in (("Wavelet", "wavelet"), ("AdaIn", "adain")):
            out = cc.color_fix_frames(hr.permute(1, 0, 2, 3)[None], lr.permute(1, 0, 2, 3)[None], mode)
            assert _maxerr(out, c[key]) < 1e-5, (name, mode)


@pytest.mark.parametrize("T,C,h,w", [(1, 3, 3, 5), (2, 3, 9, 7), (3, 3, 31, 45), (1, 1, 2, 2), (2, 4, 12, 20)])
def test_color_fix_against_oracle(T, C, h, w):
    from oracle import color_oracle as co
    from upscale_a_video_b200 import color_correction as cc
    lr = (torch.rand(T, C, h, w) * 2 - 1)
    hr = (torch.randn(T, C, 4 * h, 4 * w) * 0.6).clamp(-1.3, 1.3)
    up_o = co.bicubic_upsample(lr, 4)
    up = cc.upsample_lr_frames(lr.cuda(), 4)
    assert _maxerr(up, up_o) < 2e-6
    assert _maxerr(cc.adaptive_instance_normalization(hr.cuda(), up_o.cuda()), co.adaptive_instance_normalization(hr, up_o)) < 1e-5
    assert _maxerr(cc.wavelet_reconstruction(hr.cuda(), up_o.cuda()), co.wavelet_reconstruction(hr, up_o)) < 2e-6
    h_o, l_o = co.wavelet_decomposition(hr)
    h_g, l_g = cc.wavelet_decomposition(hr.cuda())
    assert _maxerr(h_g, h_o) < 2e-6 and _maxerr(l_g, l_o) < 1e-6
    assert torch.equal(cc.pack_video_uint8(hr.cuda()).cpu(), co.pack_video_uint8(hr))


def test_pack_uint8_known_values_and_rounding():
    from upscale_a_video_b200 import color_correction as cc
    # (x / 2 + 0.5) * 255 truncated: boundaries, clamps, and a value where an FMA-contracted x * 0.5 + 0.5 would differ
    vals = torch.tensor([-1.0, 1.0, -2.0, 3.0, 0.0, 1.0 - 2.0 ** -23, -1.0 + 2.0 ** -24, 0.00392157, 0.9999, -0.9999,
                         2.0 / 255 - 1.0, 4.0 / 255 - 1.0 - 1e-7], dtype=torch.float32)
    x = vals.view(1, 1, 1, -1).repeat(1, 3, 2, 1).cuda()
    ref = ((x.cpu() / 2 + 0.5).clamp(0, 1) * 255).permute(0, 2, 3, 1).to(torch.int32).to(torch.uint8)
    assert torch.equal(cc.pack_video_uint8(x).cpu(), ref)
    # exhaustive over a dense grid of fp32 inputs in [-1.01, 1.01]
    dense = torch.linspace(-1.01, 1.01, 3 * 1001 * 997).view(1, 3, 1001, 997).cuda()
    ref = ((dense / 2 + 0.5).clamp(0, 1) * 255).permute(0, 2, 3, 1).to(torch.int32).to(torch.uint8)  # torch op sequence on GPU
    assert torch.equal(cc.pack_video_uint8(dense), ref)


def test_full_size_properties():
    """BASELINE config-2 output size (8 frames 1280x2304): properties that need no oracle run"""
    from upscale_a_video_b200 import color_correction as cc
    T, h, w = 8, 320, 576
    lr = (torch.rand(T, 3, h, w, device="cuda") * 2 - 1)
    hr = torch.nn.functional.interpolate(lr, scale_factor=4, mode="nearest") * 0.8 + 0.1 + 0.05 * torch.randn(T, 3, 4 * h, 4 * w, device="cuda")
    up = cc.upsample_lr_frames(lr, 4)
    assert up.shape == hr.shape
    # bicubic reproduces constants and is bounded by the overshoot of the A=-0.75 kernel (sum |w| <= 1.28 per axis)
    const = cc.upsample_lr_frames(torch.full((1, 3, h, w), 0.37, device="cuda"), 4)
    assert (const - 0.37).abs().max().item() < 1e-6
    assert up.abs().max().item() < 1.28 * 1.28 + 1e-3
    # AdaIN: the result carries the style's per-plane statistics; deterministic
    a1 = cc.adaptive_instance_normalization(hr, up)
    a2 = cc.adaptive_instance_normalization(hr, up)
    assert torch.equal(a1, a2)
    m_a, s_a = cc.calc_mean_std(a1)
    m_s, s_s = cc.calc_mean_std(up)
    assert (m_a - m_s).abs().max().item() < 1e-5 and (s_a / s_s - 1).abs().max().item() < 1e-4
    # wavelet: high + low telescopes back to the image; fixing an image with itself is the identity
    high, low = cc.wavelet_decomposition(hr)
    assert (high + low - hr).abs().max().item() < 1e-5
    assert (cc.wavelet_reconstruction(hr, hr) - hr).abs().max().item() < 1e-5
    # linearity of the decomposition
    h2, l2 = cc.wavelet_decomposition(hr * 0.5)
    assert (h2 - high * 0.5).abs().max().item() < 1e-6 and (l2 - low * 0.5).abs().max().item() < 1e-6
    # packing: shape / layout and agreement with the torch op sequence on the same GPU
    pk = cc.pack_video_uint8(a1)
    assert pk.shape == (T, 4 * h, 4 * w, 3) and pk.dtype == torch.uint8
    ref = ((a1 / 2 + 0.5).clamp(0, 1) * 255).permute(0, 2, 3, 1).to(torch.int32).to(torch.uint8)
    assert torch.equal(pk, ref)


def test_errors_are_loud():
    from upscale_a_video_b200 import color_correction as cc, _lib
    with pytest.raises(RuntimeError):
        cc.calc_mean_std(torch.zeros(1, 3, 4, 4))  # CPU tensor: no fallback
    with pytest.raises(AssertionError):
        cc.calc_mean_std(torch.zeros(3, 4, 4, device="cuda"))
    lib = _lib.load()
    assert lib.uav_wavelet_level(None, 1, 4, 4, 1, None, None, 0, None, None) != 0
    assert b"uav_wavelet_level" in lib.uav_last_error_string()
    x = torch.zeros(1, 1, 4, 4, device="cuda")
    assert lib.uav_wavelet_level(x.data_ptr(), 1, 4, 4, 1, x.data_ptr(), None, 0, None, None) != 0  # aliasing refused
