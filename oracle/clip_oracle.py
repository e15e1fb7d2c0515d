"""clip_oracle.py — CPU restatement of the CLIP text encoder that produces `prompt_embeds`
(`self.text_encoder(input_ids, attention_mask=None)[0]`, /root/reference/models_video/pipeline_upscale_a_video.py:239-245;
the encoder itself is the pip dependency `transformers.CLIPTextModel`, not vendored in /root/reference).

TEST INFRASTRUCTURE (see oracle/uav_oracle.py for the import rules).  Parity status: pinned against the installed
`transformers` CLIPTextModel (5.5.0 here; the arithmetic of `CLIPTextTransformer` — pre-LayerNorm blocks, causal mask,
final LayerNorm — has not changed since the 4.2x the reference targets) by `oracle/make_golden_clip.py` ->
`tests/golden/clip.pt`.  Functional over a flat state dict with transformers' key names (`text_model.*`).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def _act(x, name: str):
    if name == "gelu":
        return F.gelu(x)
    if name == "quick_gelu":
        return x * torch.sigmoid(1.702 * x)
    raise ValueError(name)


def clip_text_forward(sd: SD, cfg: dict, input_ids: torch.Tensor) -> torch.Tensor:
    """CLIPTextTransformer.forward without padding mask: last_hidden_state (b, n, hidden)"""
    p = "text_model."
    b, n = input_ids.shape
    heads, eps = cfg["num_attention_heads"], cfg.get("layer_norm_eps", 1e-5)
    x = sd[p + "embeddings.token_embedding.weight"][input_ids] + sd[p + "embeddings.position_embedding.weight"][:n][None]
    hidden = x.shape[-1]
    d = hidden // heads
    causal = torch.full((n, n), float("-inf")).triu(1)
    for i in range(cfg["num_hidden_layers"]):
        lp = f"{p}encoder.layers.{i}."
        h = F.layer_norm(x, (hidden,), sd[lp + "layer_norm1.weight"], sd[lp + "layer_norm1.bias"], eps)
        q = F.linear(h, sd[lp + "self_attn.q_proj.weight"], sd[lp + "self_attn.q_proj.bias"]) * d ** -0.5
        k = F.linear(h, sd[lp + "self_attn.k_proj.weight"], sd[lp + "self_attn.k_proj.bias"])
        v = F.linear(h, sd[lp + "self_attn.v_proj.weight"], sd[lp + "self_attn.v_proj.bias"])
        q, k, v = (t.view(b, n, heads, d).transpose(1, 2) for t in (q, k, v))
        a = torch.softmax(q @ k.transpose(-1, -2) + causal, dim=-1) @ v
        a = a.transpose(1, 2).reshape(b, n, hidden)
        x = x + F.linear(a, sd[lp + "self_attn.out_proj.weight"], sd[lp + "self_attn.out_proj.bias"])
        h = F.layer_norm(x, (hidden,), sd[lp + "layer_norm2.weight"], sd[lp + "layer_norm2.bias"], eps)
        h = _act(F.linear(h, sd[lp + "mlp.fc1.weight"], sd[lp + "mlp.fc1.bias"]), cfg["hidden_act"])
        x = x + F.linear(h, sd[lp + "mlp.fc2.weight"], sd[lp + "mlp.fc2.bias"])
    return F.layer_norm(x, (hidden,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], eps)
