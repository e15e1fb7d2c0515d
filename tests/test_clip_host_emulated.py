"""CPU check of the CLIP text encoder's host logic (fused q|k|v packing, residual / activation placement, embedding lookup)
with emulated kernels against the fixtures minted from transformers' CLIPTextModel (tests/golden/clip.pt)."""
import os

import pytest
import torch

import emu_ops

G = os.path.join(os.path.dirname(__file__), "golden")


class Emu:
    ACT_NONE, ACT_GELU, ACT_QUICK_GELU = 0, 6, 7
    layer_norm = staticmethod(emu_ops.layer_norm)

    @staticmethod
    def linear(a, w, bias=None, *, residual=None, act=0, out=None, out_dtype=torch.float16, **_):
        v = a.float().reshape(-1, a.shape[-1]) @ w.float().t()
        if bias is not None:
            v = v + bias.float()
        if act == 6:
            v = torch.nn.functional.gelu(v)
        elif act == 7:
            v = v * torch.sigmoid(1.702 * v)
        if residual is not None:
            v = v + residual.float().reshape(-1, residual.shape[-1])
        return v.reshape(*a.shape[:-1], -1).to(out_dtype)

    @staticmethod
    def attention_causal(q, k, v, heads, *, scale=None, out=None):
        b, n, c = q.shape
        d = c // heads
        qq, kk, vv = (t.float().reshape(b, n, heads, d).transpose(1, 2) for t in (q, k, v))
        s = qq @ kk.transpose(-1, -2) * (d ** -0.5 if scale is None else scale) + torch.full((n, n), float("-inf")).triu(1)
        return (torch.softmax(s, -1) @ vv).transpose(1, 2).reshape(b, n, c).half()


def test_clip_text_host_logic_vs_transformers_fixtures(monkeypatch):
    from oracle.weights import make_state_dict
    from upscale_a_video_b200 import _lib, clip_text
    monkeypatch.setattr(clip_text, "ops", Emu)
    monkeypatch.setattr(_lib, "require_cuda", lambda t, who: None)
    g = torch.load(os.path.join(G, "clip.pt"), weights_only=False)
    for name, c in g["cases"].items():
        m = clip_text.CLIPTextModel(clip_text.CLIPTextConfig(**c["config"]))
        assert {k: list(v.shape) for k, v in m.state_dict().items()} == c["shapes"]  # transformers' keys and shapes
        m.load_state_dict(make_state_dict(c["shapes"], g["seed"]), strict=True)
        out = m.half().eval()(c["input_ids"])[0]
        ref = c["last_hidden_state"]
        err = ((out.float()[..., ::c["col_stride"]] - ref).norm() / ref.norm()).item()
        print(f"[clip host-emulated {name}] rel L2 err {err:.3e}")
        assert out.dtype == torch.float16 and err < 3e-3
