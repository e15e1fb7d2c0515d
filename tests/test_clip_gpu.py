"""GPU parity of the CLIP text encoder (clip_text.py; uav_attention_causal, GELU epilogues) against the fixtures minted from
transformers' CLIPTextModel: relative L2 error of the fp16 path vs the fp32 reference <= 3e-3 (the emulated-kernel host test
measures 1.0e-3); the causal attention kernel is also checked alone against torch."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def test_attention_causal_kernel(uav_lib):
    from upscale_a_video_b200 import ops
    torch.manual_seed(0)
    for b, heads, d, n in [(2, 16, 64, 77), (1, 4, 16, 77), (3, 8, 128, 128), (1, 2, 64, 1), (2, 12, 64, 33)]:
        c = heads * d
        qkv = torch.randn(b, n, 3 * c, device="cuda").half()
        o = ops.attention_causal(qkv[..., :c], qkv[..., c:2 * c], qkv[..., 2 * c:], heads)
        q, k, v = (t.float().reshape(b, n, heads, d).transpose(1, 2) for t in (qkv[..., :c], qkv[..., c:2 * c], qkv[..., 2 * c:]))
        s = q @ k.transpose(-1, -2) * d ** -0.5 + torch.full((n, n), float("-inf"), device="cuda").triu(1)
        ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(b, n, c)
        assert (o.float() - ref).abs().max().item() < 3e-3, (b, heads, d, n)


def test_clip_text_model_vs_transformers_fixtures(uav_lib):
    from oracle.weights import make_state_dict
    from upscale_a_video_b200.clip_text import CLIPTextConfig, CLIPTextModel
    g = torch.load(os.path.join(G, "clip.pt"), weights_only=False)
    for name, c in g["cases"].items():
        m = CLIPTextModel(CLIPTextConfig(**c["config"]))
        m.load_state_dict(make_state_dict(c["shapes"], g["seed"]), strict=True)
        m = m.half().eval().cuda()
        out = m(c["input_ids"].cuda())[0]
        ref = c["last_hidden_state"]
        err = ((out.float().cpu()[..., ::c["col_stride"]] - ref).norm() / ref.norm()).item()
        print(f"[clip {name}] rel L2 err {err:.3e}")
        assert err < 3e-3
        assert torch.equal(out, m(c["input_ids"].cuda())[0])
