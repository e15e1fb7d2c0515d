"""raft_oracle.py — CPU restatement of the RAFT bidirectional optical flow that feeds `flows_bi`
(/root/reference/models_video/RAFT/{raft_bi,raft,corr,update,extractor}.py, utils/utils.py).

TEST INFRASTRUCTURE.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU legs may import this file; the
product package (`upscale_a_video_b200/`) never does.  The product side of this row is NOT built yet; the
oracle and its fixtures are in place so that the CUDA path can be brought up against them.

Parity status: pinned against the UNMODIFIED reference modules run in the build container by
`oracle/make_golden_raft.py` (fixtures `tests/golden/raft.pt`, checked on CPU by `tests/test_oracle_golden.py`).
Functional over a flat state dict with the reference's key names (`fnet.*`, `cnet.*`, `update_block.*`: the keys of
`raft-things.pth` minus the `module.` prefix that `raft_bi.py:27-29` strips).  Only the configuration the pipeline uses is
restated: `small=False`, `mixed_precision=False`, `alternate_corr=False`, `test_mode=True`, eval-mode BatchNorm.
Reference behaviour kept: frames below 128 px on a side make the last correlation-pyramid level 1x1 and
`bilinear_sampler` divide by zero (NaN flows) — the fixtures therefore use >= 124 px frames.
"""
from __future__ import annotations

from math import ceil
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
CORR_LEVELS, CORR_RADIUS = 4, 4  # raft.py:38-41
HDIM = CDIM = 128


def _conv(sd: SD, p: str, x, stride=1, padding=0):
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=padding)


def _norm(sd: SD, p: str, x, kind: str):
    """extractor.py:13-37,124-134: 'instance' = nn.InstanceNorm2d defaults (no affine, no running stats, eps 1e-5);
    'batch' = eval-mode BatchNorm2d (running statistics, eps 1e-5)"""
    if kind == "instance":
        return F.instance_norm(x, eps=1e-5)
    if kind == "batch":
        return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                            training=False, eps=1e-5)
    raise ValueError(kind)


def residual_block(sd: SD, p: str, x, kind: str, stride: int):
    """extractor.py:6-58"""
    y = F.relu(_norm(sd, p + ".norm1", _conv(sd, p + ".conv1", x, stride=stride, padding=1), kind))
    y = F.relu(_norm(sd, p + ".norm2", _conv(sd, p + ".conv2", y, padding=1), kind))
    if stride != 1:
        # nn.Sequential(conv1x1 stride, norm3): norm3 is registered both as `.norm3` and as `.downsample.1`
        x = _norm(sd, p + ".downsample.1", _conv(sd, p + ".downsample.0", x, stride=stride), kind)
    return F.relu(x + y)


def basic_encoder(sd: SD, p: str, x, kind: str):
    """extractor.py:118-190 (BasicEncoder.forward; dropout inactive in eval)"""
    x = F.relu(_norm(sd, p + ".norm1", _conv(sd, p + ".conv1", x, stride=2, padding=3), kind))
    for name, stride in (("layer1", 1), ("layer2", 2), ("layer3", 2)):
        x = residual_block(sd, f"{p}.{name}.0", x, kind, stride)
        x = residual_block(sd, f"{p}.{name}.1", x, kind, 1)
    return _conv(sd, p + ".conv2", x)


def coords_grid(batch: int, ht: int, wd: int):
    """utils.py:73-76: channel 0 = x, channel 1 = y"""
    ys, xs = torch.meshgrid(torch.arange(ht), torch.arange(wd), indexing="ij")
    return torch.stack([xs, ys], dim=0).float()[None].repeat(batch, 1, 1, 1)


def bilinear_sampler(img, coords):
    """utils.py:57-70: pixel coordinates -> grid_sample(align_corners=True), zero padding"""
    H, W = img.shape[-2:]
    xgrid, ygrid = coords.split([1, 1], dim=-1)
    xgrid = 2 * xgrid / (W - 1) - 1
    ygrid = 2 * ygrid / (H - 1) - 1
    return F.grid_sample(img, torch.cat([xgrid, ygrid], dim=-1), align_corners=True)


def corr_pyramid(fmap1, fmap2):
    """corr.py:12-28,52-60: all-pairs correlation / sqrt(dim), then 3 x avg_pool2d(2)"""
    b, dim, ht, wd = fmap1.shape
    corr = torch.matmul(fmap1.view(b, dim, ht * wd).transpose(1, 2), fmap2.view(b, dim, ht * wd))
    corr = corr / torch.sqrt(torch.tensor(dim).float())
    corr = corr.reshape(b * ht * wd, 1, ht, wd)
    pyr = [corr]
    for _ in range(CORR_LEVELS - 1):
        corr = F.avg_pool2d(corr, 2, stride=2)
        pyr.append(corr)
    return pyr


def corr_lookup(pyr, coords):
    """corr.py:30-50.  Note the upstream quirk kept by the reference: `delta = stack(meshgrid(dy, dx))` puts the dy offsets
    in the component that is added to x (and dx to y), i.e. the (2r+1)^2 window is indexed [x-offset][y-offset]."""
    r = CORR_RADIUS
    coords = coords.permute(0, 2, 3, 1)
    b, h1, w1, _ = coords.shape
    d = torch.linspace(-r, r, 2 * r + 1)
    delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), dim=-1)
    out = []
    for i, corr in enumerate(pyr):
        centroid = coords.reshape(b * h1 * w1, 1, 1, 2) / 2 ** i
        out.append(bilinear_sampler(corr, centroid + delta.view(1, 2 * r + 1, 2 * r + 1, 2)).view(b, h1, w1, -1))
    return torch.cat(out, dim=-1).permute(0, 3, 1, 2).contiguous().float()


def motion_encoder(sd: SD, p: str, flow, corr):
    """update.py:79-98 (BasicMotionEncoder)"""
    cor = F.relu(_conv(sd, p + ".convc1", corr))
    cor = F.relu(_conv(sd, p + ".convc2", cor, padding=1))
    flo = F.relu(_conv(sd, p + ".convf1", flow, padding=3))
    flo = F.relu(_conv(sd, p + ".convf2", flo, padding=1))
    out = F.relu(_conv(sd, p + ".conv", torch.cat([cor, flo], dim=1), padding=1))
    return torch.cat([out, flow], dim=1)


def sep_conv_gru(sd: SD, p: str, h, x):
    """update.py:33-60 (SepConvGRU): a horizontal (1x5) then a vertical (5x1) GRU step"""
    for s, pad in (("1", (0, 2)), ("2", (2, 0))):
        hx = torch.cat([h, x], dim=1)
        z = torch.sigmoid(_conv(sd, f"{p}.convz{s}", hx, padding=pad))
        r = torch.sigmoid(_conv(sd, f"{p}.convr{s}", hx, padding=pad))
        q = torch.tanh(_conv(sd, f"{p}.convq{s}", torch.cat([r * h, x], dim=1), padding=pad))
        h = (1 - z) * h + z * q
    return h


def update_block(sd: SD, p: str, net, inp, corr, flow):
    """update.py:115-139 (BasicUpdateBlock): returns (net, 0.25 * mask, delta_flow)"""
    motion = motion_encoder(sd, p + ".encoder", flow, corr)
    net = sep_conv_gru(sd, p + ".gru", net, torch.cat([inp, motion], dim=1))
    delta = _conv(sd, p + ".flow_head.conv2", F.relu(_conv(sd, p + ".flow_head.conv1", net, padding=1)), padding=1)
    mask = _conv(sd, p + ".mask.2", F.relu(_conv(sd, p + ".mask.0", net, padding=1)))
    return net, 0.25 * mask, delta


def upsample_flow(flow, mask):
    """raft.py:73-84: convex combination of the 3x3 neighbourhood of 8 * flow, weights = softmax over the 9 taps"""
    N, _, H, W = flow.shape
    mask = torch.softmax(mask.view(N, 1, 9, 8, 8, H, W), dim=2)
    up = F.unfold(8 * flow, [3, 3], padding=1).view(N, 2, 9, 1, 1, H, W)
    up = torch.sum(mask * up, dim=2).permute(0, 1, 4, 2, 5, 3)
    return up.reshape(N, 2, 8 * H, 8 * W)


def raft_forward(sd: SD, image1, image2, iters: int = 12) -> Tuple[torch.Tensor, torch.Tensor]:
    """raft.py:87-143 with test_mode=True: returns (flow at 1/8 resolution, convex-upsampled flow)"""
    n = image1.shape[0]
    fmaps = basic_encoder(sd, "fnet", torch.cat([image1, image2], dim=0), "instance").float()
    fmap1, fmap2 = fmaps[:n], fmaps[n:]
    pyr = corr_pyramid(fmap1, fmap2)
    cnet = basic_encoder(sd, "cnet", image1, "batch")
    net, inp = torch.tanh(cnet[:, :HDIM]), torch.relu(cnet[:, HDIM:])
    H, W = image1.shape[-2:]
    coords0 = coords_grid(n, H // 8, W // 8)
    coords1 = coords0.clone()
    flow_up = None
    for _ in range(iters):
        corr = corr_lookup(pyr, coords1)
        net, mask, delta = update_block(sd, "update_block", net, inp, corr, coords1 - coords0)
        coords1 = coords1 + delta
        flow_up = upsample_flow(coords1 - coords0, mask)
    return coords1 - coords0, flow_up


def resize_flow_pytorch(flow, newh: int, neww: int):
    """raft_bi.py:11-16.  Quirk kept: the scale factors are applied to ROWS 0 and 1 of every channel (`flow[:, :, 0]`,
    `flow[:, :, 1]`), not to the x / y channels; when newh == oldh and neww == oldw both factors are 1 and it is a no-op."""
    oldh, oldw = flow.shape[-2:]
    flow = F.interpolate(flow, (newh, neww), mode="bilinear")
    flow[:, :, 0] *= newh / oldh
    flow[:, :, 1] *= neww / oldw
    return flow


def raft_bi_forward(sd: SD, frames, iters: int = 20):
    """raft_bi.py:47-68: frames (b, c, t, h, w) in [-1, 1]; returns forward / backward flows (b, 2, t-1, h, w)"""
    B, C, T, H, W = frames.shape
    H_, W_ = int(ceil(H / 8) * 8), int(ceil(W / 8) * 8)
    frames = F.interpolate(frames, (T, H_, W_), mode="trilinear")
    f1 = frames[:, :, :-1].permute(0, 2, 1, 3, 4).reshape(B * (T - 1), C, H_, W_).contiguous()
    f2 = frames[:, :, 1:].permute(0, 2, 1, 3, 4).reshape(B * (T - 1), C, H_, W_).contiguous()
    _, fwd = raft_forward(sd, f1, f2, iters)
    _, bwd = raft_forward(sd, f2, f1, iters)
    fwd, bwd = resize_flow_pytorch(fwd, H, W), resize_flow_pytorch(bwd, H, W)
    back = lambda x: x.reshape(B, T - 1, 2, H, W).permute(0, 2, 1, 3, 4).contiguous()  # noqa: E731
    return back(fwd), back(bwd)


def short_clip_len(width: int) -> int:
    """raft_bi.py:73-80"""
    if width <= 640:
        return 12
    if width <= 720:
        return 8
    if width <= 1280:
        return 4
    return 2


def raft_bi_forward_slicing(sd: SD, frames, iters: int = 20):
    """raft_bi.py:71-104: clips of `short_clip_len` frames, each (but the first) re-reading the previous frame"""
    n, clip = frames.shape[2], short_clip_len(frames.shape[-1])
    if n <= clip:
        return raft_bi_forward(sd, frames, iters)
    fs, bs = [], []
    for f in range(0, n, clip):
        end = min(n, f + clip)
        a, b = raft_bi_forward(sd, frames[:, :, (f if f == 0 else f - 1):end], iters)
        fs.append(a)
        bs.append(b)
    return torch.cat(fs, dim=2), torch.cat(bs, dim=2)


def synth_clip(T: int, H: int, W: int, seed: int):
    """seeded test clip (1, 3, T, H, W) in [-1, 1]: a textured pattern translating by ~1.5 px / frame plus noise.
    Shared by the fixture generator and the tests, so the fixtures only store (T, H, W, seed)."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
    frames = []
    for t in range(T):
        dx, dy = 1.5 * t / W, 0.8 * t / H
        base = torch.stack([torch.sin(20 * (xx - dx) + 9 * (yy - dy)), torch.cos(17 * (yy - dy) - 5 * (xx - dx)),
                            torch.sin(31 * (xx - dx) * (yy - dy) + 1.0)])
        frames.append(0.7 * base + 0.1 * torch.randn(3, H, W, generator=g))
    return torch.stack(frames, dim=1)[None].clamp(-1, 1)
