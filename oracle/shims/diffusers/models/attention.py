"""FeedForward / GEGLU / AdaLayerNorm / AttentionBlock of diffusers 0.16, restated from their behaviour
(the reference vendors a 0.11 copy at models_video/diffusers_attention.py:249-381,735-858
which this must agree with).  AttentionBlock: GroupNorm -> q,k,v Linear -> softmax(q k^T / sqrt(C/heads)) in fp32
-> proj_attn -> (+ residual) / rescale."""
import math
import torch
import torch.nn.functional as F
from torch import nn


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, dropout=0.0, activation_fn="geglu", final_dropout=False):
        super().__init__()
        assert activation_fn == "geglu"
        inner = int(dim * mult)
        self.net = nn.ModuleList([GEGLU(dim, inner), nn.Dropout(dropout), nn.Linear(inner, dim_out or dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class AdaLayerNorm(nn.Module):
    def __init__(self, embedding_dim, num_embeddings):
        super().__init__()
        self.emb = nn.Embedding(num_embeddings, embedding_dim)
        self.silu = nn.SiLU()
        self.linear = nn.Linear(embedding_dim, embedding_dim * 2)
        self.norm = nn.LayerNorm(embedding_dim, elementwise_affine=False)

    def forward(self, x, timestep):
        scale, shift = torch.chunk(self.linear(self.silu(self.emb(timestep))), 2)
        return self.norm(x) * (1 + scale) + shift


class AttentionBlock(nn.Module):
    def __init__(self, channels, num_head_channels=None, norm_num_groups=32, rescale_output_factor=1.0, eps=1e-5):
        super().__init__()
        self.channels = channels
        self.num_heads = channels // num_head_channels if num_head_channels is not None else 1
        self.group_norm = nn.GroupNorm(num_channels=channels, num_groups=norm_num_groups, eps=eps, affine=True)
        self.query = nn.Linear(channels, channels)
        self.key = nn.Linear(channels, channels)
        self.value = nn.Linear(channels, channels)
        self.rescale_output_factor = rescale_output_factor
        self.proj_attn = nn.Linear(channels, channels, bias=True)
        self._use_memory_efficient_attention_xformers = False

    def forward(self, hidden_states):
        residual = hidden_states
        b, c, h, w = hidden_states.shape
        x = self.group_norm(hidden_states).view(b, c, h * w).transpose(1, 2)
        q, k, v = self.query(x), self.key(x), self.value(x)
        hd = c // self.num_heads

        def split(t):
            return t.reshape(b, -1, self.num_heads, hd).permute(0, 2, 1, 3).reshape(b * self.num_heads, -1, hd)

        q, k, v = split(q), split(k), split(v)
        scores = torch.baddbmm(torch.empty(q.shape[0], q.shape[1], k.shape[1], dtype=q.dtype, device=q.device),
                               q, k.transpose(-1, -2), beta=0, alpha=1 / math.sqrt(hd))
        probs = torch.softmax(scores.float(), dim=-1).type(scores.dtype)
        o = torch.bmm(probs, v)
        o = o.reshape(b, self.num_heads, -1, hd).permute(0, 2, 1, 3).reshape(b, -1, c)
        o = self.proj_attn(o).transpose(-1, -2).reshape(b, c, h, w)
        return (o + residual) / self.rescale_output_factor
