"""Restatement of rotary-embedding-torch 0.2.3 `RotaryEmbedding` as used at
models_video/unet_video.py:203 and attention.py:709-711:
freqs_j = theta^(-2j/dim) kept as an nn.Parameter named `freqs`; positions 0..n-1 along seq_dim=-2;
angles repeated pairwise; x*cos + rotate_half(x)*sin on the first `dim` features, interleaved pairs."""
import torch
from torch import nn


def _rotate_half(x):
    x = x.reshape(*x.shape[:-1], -1, 2)
    x1, x2 = x.unbind(dim=-1)
    return torch.stack((-x2, x1), dim=-1).reshape(*x.shape[:-2], -1)


class RotaryEmbedding(nn.Module):
    def __init__(self, dim, theta=10000):
        super().__init__()
        self.freqs = nn.Parameter(1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim)), requires_grad=False)

    def rotate_queries_or_keys(self, t, seq_dim=-2):
        n = t.shape[seq_dim]
        pos = torch.arange(n, device=t.device).type_as(self.freqs)
        ang = torch.einsum("i,j->ij", pos, self.freqs)        # (n, dim/2)
        ang = torch.repeat_interleave(ang, 2, dim=-1)          # (n, dim)  "(n r), r=2"
        rot = ang.shape[-1]
        tl, tr = t[..., :rot], t[..., rot:]
        tl = tl * ang.cos().to(t.dtype) + _rotate_half(tl) * ang.sin().to(t.dtype)
        return torch.cat((tl, tr), dim=-1)
