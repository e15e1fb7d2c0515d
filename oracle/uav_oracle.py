"""uav_oracle.py — CPU restatement (plain PyTorch, fp32/fp64 capable) of the Upscale-A-Video sampling path.

TEST INFRASTRUCTURE.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this file; the product package (`upscale_a_video_b200/`) never does.

Parity status: the reference ships no tests or golden vectors, so this restatement is pinned
against outputs of the UNMODIFIED reference modules themselves, run in the build container through
`oracle/shims` by `oracle/make_golden.py`; the resulting vectors are committed under `tests/golden/` and
`tests/test_oracle_golden.py` checks this file against them on CPU.

Everything is functional over a flat state dict with the reference's own key names, in the reference's
"b c t h w" layout.  Each function cites the reference code it restates (paths relative to
/root/reference/models_video/).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# ------------------------------------------------------------------------------------------------
# primitives
# ------------------------------------------------------------------------------------------------
def _has(sd: SD, key: str) -> bool:
    return key in sd


def linear(sd: SD, p: str, x):
    """nn.Linear"""
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def group_norm(sd: SD, p: str, x, groups: int, eps: float):
    """nn.GroupNorm on (b c t h w) or (n c h w): statistics over all trailing dims (resnet.py:231,267)."""
    return F.group_norm(x, groups, sd[p + ".weight"], sd[p + ".bias"], eps)


def inflated_conv(sd: SD, p: str, x, stride=1, padding=None):
    """InflatedConv3d: 2-D conv applied per frame (resnet.py:94-101)."""
    w = sd[p + ".weight"]
    if padding is None:
        padding = w.shape[-1] // 2
    b, c, t, h, wd = x.shape
    y = F.conv2d(x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, wd), w, sd.get(p + ".bias"), stride=stride, padding=padding)
    return y.reshape(b, t, *y.shape[1:]).permute(0, 2, 1, 3, 4)


def conv3d(sd: SD, p: str, x):
    """nn.Conv3d with 'same' zero padding, stride 1 (resnet.py:332,348,361,461)."""
    w = sd[p + ".weight"]
    pad = tuple((k - 1) // 2 for k in w.shape[2:])
    return F.conv3d(x, w, sd.get(p + ".bias"), padding=pad)


def resnet_block3d(sd: SD, p: str, x, temb, eps: float, groups: int = 32, groups_out: Optional[int] = None):
    """ResnetBlock3D.forward (resnet.py:264-294), time_embedding_norm='default', output_scale_factor=1."""
    groups_out = groups if groups_out is None else groups_out
    h = F.silu(group_norm(sd, p + ".norm1", x, groups, eps))
    h = inflated_conv(sd, p + ".conv1", h)
    if temb is not None and _has(sd, p + ".time_emb_proj.weight"):
        h = h + linear(sd, p + ".time_emb_proj", F.silu(temb))[:, :, None, None, None]
    h = F.silu(group_norm(sd, p + ".norm2", h, groups_out, eps))
    h = inflated_conv(sd, p + ".conv2", h)
    if _has(sd, p + ".conv_shortcut.weight"):
        x = inflated_conv(sd, p + ".conv_shortcut", x)
    return x + h


def resnet_block3d_cnn(sd: SD, p: str, x, temb, eps: float = 1e-6, groups: int = 32):
    """ResnetBlock3DCNN.forward (resnet.py:363-393): same block with (k,1,1) temporal nn.Conv3d."""
    h = F.silu(group_norm(sd, p + ".norm1", x, groups, eps))
    h = conv3d(sd, p + ".conv1", h)
    if temb is not None and _has(sd, p + ".time_emb_proj.weight"):
        h = h + linear(sd, p + ".time_emb_proj", F.silu(temb))[:, :, None, None, None]
    h = F.silu(group_norm(sd, p + ".norm2", h, groups, eps))
    h = conv3d(sd, p + ".conv2", h)
    if _has(sd, p + ".conv_shortcut.weight"):
        x = conv3d(sd, p + ".conv_shortcut", x)
    return x + h


def resnet_block3d_plus(sd: SD, p: str, x, eps: float = 1e-6, groups: int = 32, groups_out: Optional[int] = None):
    """ResnetBlock3D_plus.forward (resnet.py:464-500): ResnetBlock3D + GN -> SiLU -> Conv3d 3x3x3 residual."""
    groups_out = groups if groups_out is None else groups_out
    out = resnet_block3d(sd, p, x, None, eps, groups, groups_out)
    h = F.silu(group_norm(sd, p + ".norm_3d", out, groups_out, eps))
    return out + conv3d(sd, p + ".conv_3d", h)


def upsample3d(sd: SD, p: str, x, output_size=None):
    """Upsample3D.forward (resnet.py:126-158): nearest x(1,2,2) (or explicit size) then conv."""
    if output_size is None:
        x = F.interpolate(x, scale_factor=[1.0, 2.0, 2.0], mode="nearest")
    else:
        x = F.interpolate(x, size=output_size, mode="nearest")
    return inflated_conv(sd, p + ".conv", x)


def downsample3d(sd: SD, p: str, x, padding: int):
    """Downsample3D.forward (resnet.py:185-197): stride-2 3x3 conv; padding=0 pads (0,1,0,1) first."""
    if padding == 0:
        x = F.pad(x, (0, 1, 0, 1))
    return inflated_conv(sd, p + ".conv", x, stride=2, padding=padding)


# ------------------------------------------------------------------------------------------------
# attention (attention.py)
# ------------------------------------------------------------------------------------------------
def _heads_to_batch(t, heads):
    b, n, c = t.shape
    return t.reshape(b, n, heads, c // heads).permute(0, 2, 1, 3).reshape(b * heads, n, c // heads)


def _batch_to_heads(t, heads):
    bh, n, d = t.shape
    return t.reshape(bh // heads, heads, n, d).permute(0, 2, 1, 3).reshape(bh // heads, n, d * heads)


def cross_attention(sd: SD, p: str, x, ctx, heads: int):
    """CrossAttention.forward/_attention (attention.py:148-238): softmax(q k^T * d^-0.5) v, then to_out[0]."""
    q = linear(sd, p + ".to_q", x)
    src = x if ctx is None else ctx
    k = linear(sd, p + ".to_k", src)
    v = linear(sd, p + ".to_v", src)
    d = q.shape[-1] // heads
    q, k, v = _heads_to_batch(q, heads), _heads_to_batch(k, heads), _heads_to_batch(v, heads)
    scores = torch.baddbmm(torch.empty(q.shape[0], q.shape[1], k.shape[1], dtype=q.dtype, device=q.device), q, k.transpose(-1, -2),
                           beta=0, alpha=d ** -0.5)
    probs = scores.softmax(dim=-1)
    o = _batch_to_heads(torch.bmm(probs, v), heads)
    return linear(sd, p + ".to_out.0", o)


def rel_pos_bias(sd: SD, p: str, n: int, num_buckets: int = 32, max_distance: int = 32):
    """RelativePositionBias.forward (attention.py:735-773) -> (heads, n, n)."""
    q_pos = torch.arange(n)
    rel = q_pos[None, :] - q_pos[:, None]  # k - q
    nb = num_buckets // 2
    neg = -rel
    ret = (neg < 0).long() * nb
    a = neg.abs()
    max_exact = nb // 2
    is_small = a < max_exact
    large = max_exact + (torch.log(a.float().clamp(min=1) / max_exact) / math.log(max_distance / max_exact)
                         * (nb - max_exact)).long()
    large = torch.min(large, torch.full_like(large, nb - 1))
    bucket = ret + torch.where(is_small, a, large)
    emb = sd[p + ".relative_attention_bias.weight"]  # (num_buckets, heads)
    return emb[bucket.to(emb.device)].permute(2, 0, 1)


def rotary(freqs, t):
    """rotary-embedding-torch 0.2.3 rotate_queries_or_keys, seq_dim=-2."""
    n = t.shape[-2]
    ang = torch.arange(n, dtype=freqs.dtype, device=freqs.device)[:, None] * freqs[None, :]
    ang = torch.repeat_interleave(ang, 2, dim=-1)
    rot = ang.shape[-1]
    tl, tr = t[..., :rot], t[..., rot:]
    x = tl.reshape(*tl.shape[:-1], -1, 2)
    rh = torch.stack((-x[..., 1], x[..., 0]), dim=-1).reshape(tl.shape)
    tl = tl * ang.cos().to(t.dtype) + rh * ang.sin().to(t.dtype)
    return torch.cat((tl, tr), dim=-1)


def temporal_attention(sd: SD, p: str, x, heads: int, rotary_freqs):
    """TemporalAttention.forward/_attention (attention.py:644-733): x is ((b hw), f, c)."""
    f = x.shape[1]
    bias = rel_pos_bias(sd, p + ".time_rel_pos_bias", f)
    q = linear(sd, p + ".to_q", x)
    k = linear(sd, p + ".to_k", x)
    v = linear(sd, p + ".to_v", x)
    d = q.shape[-1] // heads

    def split(t):
        return t.reshape(t.shape[0], f, heads, d).permute(0, 2, 1, 3)

    q = (d ** -0.5) * split(q)
    k, v = split(k), split(v)
    q = rotary(rotary_freqs, q)
    k = rotary(rotary_freqs, k)
    scores = torch.einsum("bhid,bhjd->bhij", q, k) + bias.to(q.dtype)
    scores = scores - scores.amax(dim=-1, keepdim=True)
    probs = scores.softmax(dim=-1)
    o = torch.einsum("bhij,bhjd->bhid", probs, v).permute(0, 2, 1, 3).reshape(x.shape[0], f, heads * d)
    return linear(sd, p + ".to_out.0", o)


def feed_forward(sd: SD, p: str, x):
    """diffusers FeedForward with GEGLU (in-tree copy diffusers_attention.py:735-823)."""
    h, gate = linear(sd, p + ".net.0.proj", x).chunk(2, dim=-1)
    return linear(sd, p + ".net.2", h * F.gelu(gate))


def layer_norm(sd: SD, p: str, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def basic_transformer_block(sd: SD, p: str, x, ctx, heads: int, video_length: int, only_cross: bool, rotary_freqs):
    """BasicTransformerBlock.forward (attention.py:523-564); x: ((b f), hw, c)."""
    n = layer_norm(sd, p + ".norm1", x)
    x = cross_attention(sd, p + ".attn1", n, ctx if only_cross else None, heads) + x
    if _has(sd, p + ".attn2.to_q.weight"):
        n = layer_norm(sd, p + ".norm2", x)
        x = cross_attention(sd, p + ".attn2", n, ctx, heads) + x
    bf, hw, c = x.shape
    b = bf // video_length
    xt = x.reshape(b, video_length, hw, c).permute(0, 2, 1, 3).reshape(b * hw, video_length, c)
    n = layer_norm(sd, p + ".norm_temporal", xt)
    xt = temporal_attention(sd, p + ".attn_temporal", n, heads, rotary_freqs) + xt
    x = xt.reshape(b, hw, video_length, c).permute(0, 2, 1, 3).reshape(bf, hw, c)
    return feed_forward(sd, p + ".ff", layer_norm(sd, p + ".norm3", x)) + x


def transformer3d(sd: SD, p: str, x, ctx, heads: int, only_cross: bool, rotary_freqs, groups: int = 32):
    """Transformer3DModel.forward (attention.py:359-411), use_linear_projection=True."""
    b, c, f, h, w = x.shape
    ctx_rep = ctx.repeat_interleave(f, dim=0)  # 'b n c -> (b f) n c'
    x = resnet_block3d_cnn(sd, p + ".resblock_temporal", x, None)
    xf = x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
    residual = xf
    hs = F.group_norm(xf, groups, sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-6)
    hs = hs.permute(0, 2, 3, 1).reshape(b * f, h * w, c)
    hs = linear(sd, p + ".proj_in", hs)
    i = 0
    while _has(sd, f"{p}.transformer_blocks.{i}.norm1.weight"):
        hs = basic_transformer_block(sd, f"{p}.transformer_blocks.{i}", hs, ctx_rep, heads, f, only_cross, rotary_freqs)
        i += 1
    hs = linear(sd, p + ".proj_out", hs)
    hs = hs.reshape(b * f, h, w, c).permute(0, 3, 1, 2)
    out = hs + residual
    return out.reshape(b, f, c, h, w).permute(0, 2, 1, 3, 4)


def temporal_module3d(sd: SD, p: str, x, temb):
    """TemporalModule3D.forward (temporal_module.py:175-194), attention_block_types=("","")."""
    h = resnet_block3d_cnn(sd, p + ".resblocks_3d_temporal", x, temb)
    h = resnet_block3d(sd, p + ".resblocks_3d_spatial", h, temb, 1e-6)
    h = inflated_conv(sd, p + ".shift_conv", h)
    return x + h


# ------------------------------------------------------------------------------------------------
# UNetVideoModel.forward (unet_video.py:404-574)
# ------------------------------------------------------------------------------------------------
def timestep_embedding(t, dim: int, flip_sin_to_cos: bool, freq_shift: float):
    """diffusers get_timestep_embedding."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device) / (half - freq_shift)
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([emb.sin(), emb.cos()], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


def unet_forward(sd: SD, cfg: dict, sample, timestep, low_res, encoder_hidden_states, class_labels, taps=None):
    """`taps` (optional dict) receives the output of every top-level stage, for stage-wise parity debugging"""
    def _tap(name, v):
        if taps is not None:
            taps[name] = v
    boc = cfg["block_out_channels"]
    heads = cfg["attention_head_dim"]
    eps = cfg.get("norm_eps", 1e-5)
    groups = cfg.get("norm_num_groups", 32)
    n_blocks = len(boc)
    oca = cfg.get("only_cross_attention", False)
    oca = [oca] * n_blocks if isinstance(oca, bool) else list(oca)
    heads_l = [heads] * n_blocks if isinstance(heads, int) else list(heads)
    dtype = sd["conv_in.weight"].dtype
    freqs = sd["temporal_rotary_emb.freqs"]

    sample = torch.cat([sample, low_res], dim=1)
    n_up = sum(1 for i in range(n_blocks) if i != n_blocks - 1)
    forward_size = any(s % (2 ** n_up) != 0 for s in sample.shape[-2:])

    t = timestep if torch.is_tensor(timestep) else torch.tensor([timestep])
    t = t.to(sample.device).reshape(-1).expand(sample.shape[0])
    t_emb = timestep_embedding(t, boc[0], cfg.get("flip_sin_to_cos", True), cfg.get("freq_shift", 0)).to(dtype)
    emb = linear(sd, "time_embedding.linear_2", F.silu(linear(sd, "time_embedding.linear_1", t_emb)))
    if _has(sd, "class_embedding.weight"):
        emb = emb + sd["class_embedding.weight"][class_labels.to(sample.device)].to(dtype)

    x = inflated_conv(sd, "conv_in", sample)
    _tap("conv_in", x)
    skips = [x]
    for i, btype in enumerate(cfg["down_block_types"]):
        p = f"down_blocks.{i}"
        j = 0
        while _has(sd, f"{p}.resnets.{j}.norm1.weight"):
            x = resnet_block3d(sd, f"{p}.resnets.{j}", x, emb, eps, groups)
            if btype == "CrossAttnDownBlock3D":
                x = transformer3d(sd, f"{p}.attentions.{j}", x, encoder_hidden_states, heads_l[i], oca[i], freqs, groups)
            skips.append(x)
            j += 1
        if _has(sd, f"{p}.downsamplers.0.conv.weight"):
            x = downsample3d(sd, f"{p}.downsamplers.0", x, cfg.get("downsample_padding", 1))
            skips.append(x)
        _tap(f"down{i}", x)
        if i in cfg["down_temporal_idx"]:
            x = temporal_module3d(sd, f"down_temp_blocks.{i}", x, emb)
        _tap(f"down_temp{i}", x)

    x = resnet_block3d(sd, "mid_block.resnets.0", x, emb, eps, groups)
    x = transformer3d(sd, "mid_block.attentions.0", x, encoder_hidden_states, heads_l[-1], False, freqs, groups)
    x = resnet_block3d(sd, "mid_block.resnets.1", x, emb, eps, groups)
    _tap("mid", x)
    if cfg["mid_temporal"]:
        x = temporal_module3d(sd, "mid_temp_block", x, emb)
    _tap("mid_temp", x)

    oca_r = list(reversed(oca))
    heads_r = list(reversed(heads_l))
    for i, btype in enumerate(cfg["up_block_types"]):
        p = f"up_blocks.{i}"
        nres = 0
        while _has(sd, f"{p}.resnets.{nres}.norm1.weight"):
            nres += 1
        res, skips = skips[-nres:], skips[:-nres]
        is_final = i == n_blocks - 1
        up_size = skips[-1].shape[2:] if (not is_final and forward_size) else None
        for j in range(nres):
            x = torch.cat([x, res[-1]], dim=1)
            res = res[:-1]
            x = resnet_block3d(sd, f"{p}.resnets.{j}", x, emb, eps, groups)
            if btype == "CrossAttnUpBlock3D":
                x = transformer3d(sd, f"{p}.attentions.{j}", x, encoder_hidden_states, heads_r[i], oca_r[i], freqs, groups)
        if _has(sd, f"{p}.upsamplers.0.conv.weight"):
            x = upsample3d(sd, f"{p}.upsamplers.0", x, up_size)
        _tap(f"up{i}", x)
        if i in cfg["up_temporal_idx"]:
            x = temporal_module3d(sd, f"up_temp_blocks.{i}", x, emb)
        _tap(f"up_temp{i}", x)

    x = F.silu(group_norm(sd, "conv_norm_out", x, groups, eps))
    return inflated_conv(sd, "conv_out", x)


# ------------------------------------------------------------------------------------------------
# AutoencoderKLVideo (autoencoder_kl_cond_video.py, vae_video.py)
# ------------------------------------------------------------------------------------------------
# attention over N = h*w tokens materialises an N x N score matrix per frame in the reference; at the BASELINE frame size
# (N = 184 320) that is 136 GB per frame, so above ATTN_DENSE_LIMIT score elements the oracle evaluates the SAME softmax
# row-block by row-block ("exact": identical arithmetic per query row, fp32) or through torch's fused SDPA ("sdpa": what a
# user of the reference has to enable to run this size at all; used only by bench.py's reference-GPU timing leg).
ATTN_DENSE_LIMIT = 1 << 28
ATTN_LARGE_IMPL = "exact"
ATTN_Q_BLOCK = 4096


def _single_head_attention(q, k, v, denom: float):
    n, nq, _ = q.shape
    if n * nq * k.shape[1] <= ATTN_DENSE_LIMIT:
        return torch.softmax((q @ k.transpose(-1, -2)) / denom, dim=-1) @ v
    if ATTN_LARGE_IMPL == "sdpa":
        return F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None], scale=1.0 / denom)[:, 0]
    out = torch.empty_like(q)
    kt = k.transpose(-1, -2)
    for i in range(n):
        for s in range(0, nq, ATTN_Q_BLOCK):
            out[i, s:s + ATTN_Q_BLOCK] = torch.softmax((q[i, s:s + ATTN_Q_BLOCK] @ kt[i]) / denom, dim=-1) @ v[i]
    return out


def attention_block(sd: SD, p: str, x, groups: int, eps: float):
    """diffusers AttentionBlock, 1 head (in-tree copy diffusers_attention.py:330-381); x: (n c h w)."""
    n, c, h, w = x.shape
    hs = F.group_norm(x, groups, sd[p + ".group_norm.weight"], sd[p + ".group_norm.bias"], eps)
    hs = hs.reshape(n, c, h * w).transpose(1, 2)
    q, k, v = linear(sd, p + ".query", hs), linear(sd, p + ".key", hs), linear(sd, p + ".value", hs)
    o = linear(sd, p + ".proj_attn", _single_head_attention(q, k, v, math.sqrt(c)))
    return o.transpose(-1, -2).reshape(n, c, h, w) + x


def _vae_mid(sd: SD, p: str, x, groups: int, plus: bool):
    """UNetMidBlock3D(_plus).forward (unet_blocks.py:735-745, 905-915)."""
    blk = (lambda q, y: resnet_block3d_plus(sd, q, y, 1e-6, groups)) if plus else \
          (lambda q, y: resnet_block3d(sd, q, y, None, 1e-6, groups))
    x = blk(p + ".resnets.0", x)
    b, c, t, h, w = x.shape
    xf = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    xf = attention_block(sd, p + ".attentions.0", xf, groups, 1e-6)
    x = xf.reshape(b, t, c, h, w).permute(0, 2, 1, 3, 4)
    return blk(p + ".resnets.1", x)


def vae_decode(sd: SD, cfg: dict, z, img=None, w_lr: float = 1.0):
    """AutoencoderKLVideo.decode -> Decoder.forward (autoencoder_kl_cond_video.py:199-226, vae_video.py:365-405)."""
    groups = cfg.get("norm_num_groups", 32)
    plus = cfg["up_block_types"][0] == "UpDecoderBlock3D_plus"
    x = inflated_conv(sd, "post_quant_conv", z)
    x = inflated_conv(sd, "decoder.conv_in", x)
    if cfg.get("condition_img", False):
        cond = resnet_block3d_plus(sd, "decoder.condition_in.0", img, 1e-6, 3, 32)
        cond = resnet_block3d_plus(sd, "decoder.condition_in.1", cond, 1e-6, 32)
        # Fuse_sft_block.forward (resnet.py:73-79)
        e = torch.cat([cond, x], dim=1)
        e = resnet_block3d(sd, "decoder.condition_fuse.shared.0", e, None, 1e-6)
        e = resnet_block3d(sd, "decoder.condition_fuse.shared.1", e, None, 1e-6)
        scale = inflated_conv(sd, "decoder.condition_fuse.scale", e)
        shift = inflated_conv(sd, "decoder.condition_fuse.shift", e)
        x = x + w_lr * (x * scale + shift)
    x = _vae_mid(sd, "decoder.mid_block", x, groups, plus)
    for i in range(len(cfg["up_block_types"])):
        p = f"decoder.up_blocks.{i}"
        j = 0
        while _has(sd, f"{p}.resnets.{j}.norm1.weight"):
            if plus:
                x = resnet_block3d_plus(sd, f"{p}.resnets.{j}", x, 1e-6, groups)
            else:
                x = resnet_block3d(sd, f"{p}.resnets.{j}", x, None, 1e-6, groups)
            j += 1
        if _has(sd, f"{p}.upsamplers.0.conv.weight"):
            x = upsample3d(sd, f"{p}.upsamplers.0", x)
    x = F.silu(group_norm(sd, "decoder.conv_norm_out", x, groups, 1e-6))
    return inflated_conv(sd, "decoder.conv_out", x)


def vae_encode_moments(sd: SD, cfg: dict, x):
    """AutoencoderKLVideo.encode -> Encoder.forward -> quant_conv (autoencoder_kl_cond_video.py:174-185,
    vae_video.py:117-156); returns the (mean, logvar) moments tensor."""
    groups = cfg.get("norm_num_groups", 32)
    x = inflated_conv(sd, "encoder.conv_in", x)
    for i in range(len(cfg["down_block_types"])):
        p = f"encoder.down_blocks.{i}"
        j = 0
        while _has(sd, f"{p}.resnets.{j}.norm1.weight"):
            x = resnet_block3d(sd, f"{p}.resnets.{j}", x, None, 1e-6, groups)
            j += 1
        if _has(sd, f"{p}.downsamplers.0.conv.weight"):
            x = downsample3d(sd, f"{p}.downsamplers.0", x, 0)
    x = _vae_mid(sd, "encoder.mid_block", x, groups, False)
    x = F.silu(group_norm(sd, "encoder.conv_norm_out", x, groups, 1e-6))
    x = inflated_conv(sd, "encoder.conv_out", x)
    return inflated_conv(sd, "quant_conv", x)


# ------------------------------------------------------------------------------------------------
# DDIMScheduler (scheduling_ddim.py)
# ------------------------------------------------------------------------------------------------
class DDIM:
    """set_timesteps / step_v0 / step_vt / add_noise (scheduling_ddim.py:129-184, 237-259, 383-545).
    Scalars stay 0-dim CPU fp32 tensors exactly like the reference, so tensor dtype governs rounding."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 clip_sample=True, set_alpha_to_one=True, steps_offset=0, prediction_type="epsilon",
                 clip_sample_range=1.0, **_):
        if beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(beta_schedule)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.n_train = num_train_timesteps
        self.clip_sample, self.clip_range = clip_sample, clip_sample_range
        self.steps_offset, self.prediction_type = steps_offset, prediction_type
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None

    def set_timesteps(self, n):
        self.num_inference_steps = n
        ratio = self.n_train // n
        self.timesteps = torch.arange(0, n).mul(ratio).flip(0).long() + self.steps_offset

    def step_v0(self, model_output, timestep, sample):
        a = self.alphas_cumprod[timestep]
        b = 1 - a
        if self.prediction_type == "epsilon":
            x0 = (sample - b ** 0.5 * model_output) / a ** 0.5
        elif self.prediction_type == "sample":
            x0 = model_output
        else:
            x0 = (a ** 0.5) * sample - (b ** 0.5) * model_output
        if self.clip_sample:
            x0 = x0.clamp(-self.clip_range, self.clip_range)
        return x0

    def step_vt(self, x0, model_output, timestep, sample, eta=0.0):
        prev_t = timestep - self.n_train // self.num_inference_steps
        a = self.alphas_cumprod[timestep]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b = 1 - a
        if self.prediction_type == "epsilon":
            eps = model_output
        elif self.prediction_type == "sample":
            eps = (sample - a ** 0.5 * x0) / b ** 0.5
        else:
            eps = (a ** 0.5) * model_output + (b ** 0.5) * sample
        if self.clip_sample:
            x0 = x0.clamp(-self.clip_range, self.clip_range)
        var = ((1 - a_prev) / (1 - a)) * (1 - a / a_prev)
        std = eta * var ** 0.5
        direction = (1 - a_prev - std ** 2) ** 0.5 * eps
        return a_prev ** 0.5 * x0 + direction

    def add_noise(self, x, noise, timesteps):
        ac = self.alphas_cumprod.to(device=x.device, dtype=x.dtype)
        timesteps = timesteps.to(x.device)
        a = (ac[timesteps] ** 0.5).flatten()
        s = ((1 - ac[timesteps]) ** 0.5).flatten()
        while a.dim() < x.dim():
            a, s = a.unsqueeze(-1), s.unsqueeze(-1)
        return a * x + s * noise


# ------------------------------------------------------------------------------------------------
# Propagation (propagation_module.py), learnable=False
# ------------------------------------------------------------------------------------------------
def flow_warp(x, flow, mode="bilinear"):
    """flow_warp (propagation_module.py:104-135): x (n c h w), flow (n h w 2), zeros padding, align_corners=True."""
    _, _, h, w = x.shape
    gy, gx = torch.meshgrid(torch.arange(0, h).type_as(x), torch.arange(0, w).type_as(x), indexing="ij")
    grid = torch.stack((gx, gy), 2)
    v = grid + flow
    vx = 2.0 * v[..., 0] / max(w - 1, 1) - 1.0
    vy = 2.0 * v[..., 1] / max(h - 1, 1) - 1.0
    return F.grid_sample(x, torch.stack((vx, vy), dim=3).to(x), mode=mode, padding_mode="zeros", align_corners=True)


def fb_consistency(flow_fw, flow_bw, alpha1, alpha2):
    """fbConsistencyCheck (propagation_module.py:140-149)."""
    bw_warped = flow_warp(flow_bw, flow_fw.permute(0, 2, 3, 1))
    diff = flow_fw + bw_warped
    mag = (flow_fw ** 2).sum(1, keepdim=True) + (bw_warped ** 2).sum(1, keepdim=True)
    return ((diff ** 2).sum(1, keepdim=True) < alpha1 * mag + alpha2).to(flow_fw)


def propagation(x, flows_forward, flows_backward, interpolation="bilinear", mode="fuse", fuse_scale=0.5,
                alpha1=0.01, alpha2=0.5):
    """Propagation.forward, learnable=False (propagation_module.py:194-281)."""
    b, c, t, h, w = x.shape
    s = 1.0 * w / flows_forward.shape[-1]
    ff = F.interpolate(flows_forward, (t - 1, h, w), mode="area") * s
    fb = F.interpolate(flows_backward, (t - 1, h, w), mode="area") * s
    prev = [x[:, :, i] for i in range(t)]
    for name in ("backward", "forward"):
        if name == "backward":
            frame_idx = list(range(t))[::-1]
            flow_idx = frame_idx
            f_prop, f_check = ff, fb
        else:
            frame_idx = list(range(t))
            flow_idx = list(range(-1, t - 1))
            f_prop, f_check = fb, ff
        outs = []
        feat_prop = None
        for i, idx in enumerate(frame_idx):
            cur = prev[idx]
            if i == 0:
                feat_prop = cur
            else:
                fp = f_prop[:, :, flow_idx[i]]
                fc = f_check[:, :, flow_idx[i]]
                mask = fb_consistency(fp, fc, alpha1, alpha2)
                warped = flow_warp(feat_prop, fp.permute(0, 2, 3, 1), interpolation)
                if mode == "fuse":
                    warped = warped * fuse_scale + cur * (1 - fuse_scale)
                feat_prop = mask * warped + (1 - mask) * cur
            outs.append(feat_prop)
        if name == "backward":
            outs = outs[::-1]
        prev = outs
    return torch.stack(prev, dim=2)


# ------------------------------------------------------------------------------------------------
# VideoUpscalePipeline.__call__ (pipeline_upscale_a_video.py:436-717), prompt_embeds given
# ------------------------------------------------------------------------------------------------
def unet_windows(T: int, short_seq: int = 8, overlap: int = 2) -> List[tuple]:
    """window grid of pipeline_upscale_a_video.py:601-625 (incl. the re-anchored / duplicated last window)."""
    out = []
    for s in range(0, T, short_seq - overlap):
        e = min(T, s + short_seq)
        if e - s < short_seq:
            s = e - short_seq
        out.append((s, e))
    return out


def pipeline_call(unet_sd: SD, unet_cfg: dict, vae_sd: SD, vae_cfg: dict, sched: DDIM, low_res_sched: DDIM, *,
                  image, prompt_embeds, noise, latents, flows_bi=None, num_inference_steps=30, guidance_scale=6.0,
                  noise_level=120, propagation_steps: Sequence[int] = (), w_lr=1.0, return_latents=False):
    """`prompt_embeds` = cat[negative, positive] (pipeline...:319); `noise`/`latents` are the two randn draws
    of pipeline...:547 and :424 (passed in so that the generator is out of the comparison)."""
    dtype = prompt_embeds.dtype
    image_dec = image.float()
    image = image.to(dtype)
    nl = torch.tensor([noise_level], dtype=torch.long)
    image = low_res_sched.add_noise(image, noise, nl)
    cfg_on = guidance_scale > 1.0
    image = torch.cat([image] * (2 if cfg_on else 1))
    denoise_level = torch.cat([nl] * image.shape[0])
    sched.set_timesteps(num_inference_steps)
    latents = latents * sched.init_noise_sigma
    T = image.shape[2]
    for i, t in enumerate(sched.timesteps):
        lat_in = torch.cat([latents] * 2) if cfg_on else latents
        if T > 8:
            preds: List[Optional[torch.Tensor]] = [None] * T
            for (s, e) in unet_windows(T):
                out = unet_forward(unet_sd, unet_cfg, lat_in[:, :, s:e], t, image[:, :, s:e], prompt_embeds, denoise_level)
                for k, idx in enumerate(range(s, e)):
                    preds[idx] = out[:, :, k:k + 1] if preds[idx] is None else preds[idx] * 0.5 + out[:, :, k:k + 1] * 0.5
            noise_pred = torch.cat(preds, dim=2)
        else:
            noise_pred = unet_forward(unet_sd, unet_cfg, lat_in, t, image, prompt_embeds, nl)
        if cfg_on:
            u, c = noise_pred.chunk(2)
            noise_pred = u + guidance_scale * (c - u)
        x0 = sched.step_v0(noise_pred, t, latents)
        if flows_bi is not None and i in propagation_steps:
            x0 = propagation(x0, flows_bi[0].to(latents), flows_bi[1].to(latents), "nearest", "fuse", 0.5, 0.001, 0.05)
        latents = sched.step_vt(x0, noise_pred, t, latents)
    latents = latents.float()
    frames = []
    for s in range(0, T, 3):
        e = min(T, s + 3)
        z = (1 / vae_cfg["scaling_factor"]) * latents[:, :, s:e]
        frames.append(vae_decode(vae_sd, vae_cfg, z, image_dec[:, :, s:e], w_lr).clamp(-1, 1).float())
    out = torch.cat(frames, dim=2)
    return (out, latents) if return_latents else out
