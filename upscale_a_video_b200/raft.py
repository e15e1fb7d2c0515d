"""RAFT bidirectional optical flow on the B200 kernels — the mirror of the reference's `models_video/RAFT`
(`raft.py`, `raft_bi.py`, `extractor.py`, `update.py`, `corr.py`): same class names, constructor
arguments, state-dict keys (the keys of `raft-things.pth` minus the `module.` prefix) and call signatures, so
`RAFT_bi(...).forward_slicing(vframes)` drops in under `inference_upscale_a_video.py:125,190-194`.

torch.nn modules are parameter holders only.  The arithmetic runs in `csrc/`: every convolution and the all-pairs
correlation on the tcgen05 implicit-GEMM kernel (`uav_conv2d_taps` / `uav_linear`), the rest in `csrc/raft.cu`.  The
reference runs RAFT in fp32; here activations are fp16 with fp32 accumulation, the correlation volume, coordinates and
flows stay fp32 (upstream RAFT's own `mixed_precision` mode makes the same split).  Exact work removal: eval-mode
BatchNorm is folded into the context encoder's convolutions, the z and r gate convolutions share one GEMM, the 1/sqrt(256)
of the correlation is folded (as 1/4 each) into the feature encoder's output conv, and the upsampling mask head runs in
the last iteration only (the reference computes and discards it in the others).  There is no CPU path.
"""
from __future__ import annotations

from math import ceil
from typing import List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .layers import PackedModule

__all__ = ["RAFT", "RAFT_bi", "initialize_RAFT", "resize_flow_pytorch"]

HDIM = CDIM = 128
CORR_LEVELS, CORR_RADIUS = 4, 4


# ------------------------------------------------------------------------------------------------
# parameter holders (names / registration order of the reference modules)
# ------------------------------------------------------------------------------------------------
def _norm(kind: str, planes: int) -> nn.Module:
    if kind == "batch":
        return nn.BatchNorm2d(planes)
    if kind == "instance":
        return nn.InstanceNorm2d(planes)
    raise NotImplementedError(f"norm_fn={kind!r}: the pipeline uses 'instance' (fnet) and 'batch' (cnet) only")


class ResidualBlock(nn.Module):
    """extractor.py:6-58"""

    def __init__(self, in_planes: int, planes: int, norm_fn: str = "group", stride: int = 1):
        super().__init__()
        self.conv1 = nn.Conv2d(in_planes, planes, kernel_size=3, padding=1, stride=stride)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, padding=1)
        self.norm1, self.norm2 = _norm(norm_fn, planes), _norm(norm_fn, planes)
        self.stride = stride
        if stride == 1:
            self.downsample = None
        else:
            self.norm3 = _norm(norm_fn, planes)
            self.downsample = nn.Sequential(nn.Conv2d(in_planes, planes, kernel_size=1, stride=stride), self.norm3)


class BasicEncoder(nn.Module):
    """extractor.py:118-190"""

    def __init__(self, output_dim: int = 128, norm_fn: str = "batch", dropout: float = 0.0):
        super().__init__()
        self.norm_fn = norm_fn
        self.norm1 = _norm(norm_fn, 64)
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3)
        self.layer1 = nn.Sequential(ResidualBlock(64, 64, norm_fn, 1), ResidualBlock(64, 64, norm_fn, 1))
        self.layer2 = nn.Sequential(ResidualBlock(64, 96, norm_fn, 2), ResidualBlock(96, 96, norm_fn, 1))
        self.layer3 = nn.Sequential(ResidualBlock(96, 128, norm_fn, 2), ResidualBlock(128, 128, norm_fn, 1))
        self.conv2 = nn.Conv2d(128, output_dim, kernel_size=1)


class FlowHead(nn.Module):
    def __init__(self, input_dim=128, hidden_dim=256):
        super().__init__()
        self.conv1 = nn.Conv2d(input_dim, hidden_dim, 3, padding=1)
        self.conv2 = nn.Conv2d(hidden_dim, 2, 3, padding=1)


class SepConvGRU(nn.Module):
    def __init__(self, hidden_dim=128, input_dim=192 + 128):
        super().__init__()
        for s, k, p in (("1", (1, 5), (0, 2)), ("2", (5, 1), (2, 0))):
            for g in "zrq":
                setattr(self, f"conv{g}{s}", nn.Conv2d(hidden_dim + input_dim, hidden_dim, k, padding=p))


class BasicMotionEncoder(nn.Module):
    def __init__(self):
        super().__init__()
        cor_planes = CORR_LEVELS * (2 * CORR_RADIUS + 1) ** 2
        self.convc1 = nn.Conv2d(cor_planes, 256, 1, padding=0)
        self.convc2 = nn.Conv2d(256, 192, 3, padding=1)
        self.convf1 = nn.Conv2d(2, 128, 7, padding=3)
        self.convf2 = nn.Conv2d(128, 64, 3, padding=1)
        self.conv = nn.Conv2d(64 + 192, 128 - 2, 3, padding=1)


class BasicUpdateBlock(nn.Module):
    def __init__(self, hidden_dim=128):
        super().__init__()
        self.encoder = BasicMotionEncoder()
        self.gru = SepConvGRU(hidden_dim=hidden_dim, input_dim=128 + hidden_dim)
        self.flow_head = FlowHead(hidden_dim, hidden_dim=256)
        self.mask = nn.Sequential(nn.Conv2d(128, 256, 3, padding=1), nn.ReLU(inplace=True), nn.Conv2d(256, 64 * 9, 1, padding=0))


# ------------------------------------------------------------------------------------------------
# weight packing
# ------------------------------------------------------------------------------------------------
def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def _fold_bn(w: torch.Tensor, b: torch.Tensor, bn: Optional[nn.Module]):
    """eval-mode BatchNorm2d after a conv == per-output-channel affine of the conv"""
    if bn is None or not isinstance(bn, nn.BatchNorm2d):
        return w, b
    s = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    return w * s.view(-1, 1, 1, 1), (b - bn.running_mean.detach().float()) * s + bn.bias.detach().float()


def _pack_conv(conv: nn.Conv2d, bn: Optional[nn.Module] = None, scale: float = 1.0, cout_pad: int = 0):
    """stride-1 conv -> (w [Cout][kh][kw][Cin_pad8] fp16, bias fp32, kh, kw, pad_top, pad_left)"""
    w, b = conv.weight.detach().float(), conv.bias.detach().float()
    w, b = _fold_bn(w, b, bn)
    w, b = w * scale, b * scale
    cout, cin, kh, kw = w.shape
    wp = F.pad(w.permute(0, 2, 3, 1), (0, _pad8(cin) - cin))
    if cout_pad > cout:
        wp = F.pad(wp, (0, 0, 0, 0, 0, 0, 0, cout_pad - cout))
        b = F.pad(b, (0, cout_pad - cout))
    ph, pw = conv.padding
    return wp.to(torch.float16).contiguous(), b.contiguous(), kh, kw, ph, pw


def _pack_conv_s2(conv: nn.Conv2d, bn: Optional[nn.Module] = None):
    """k x k stride-2 conv with padding p as a ceil-sized stride-1 conv over the space-to-depth input
    ((N, H/2, W/2, 4*Cpad), channel = (row phase * 2 + column phase) * Cpad + c): input row 2i + ky - p = 2(i + dy) + a."""
    w, b = conv.weight.detach().float(), conv.bias.detach().float()
    w, b = _fold_bn(w, b, bn)
    cout, cin, k, _ = w.shape
    p = conv.padding[0]
    cp = _pad8(cin)
    dys = [(ky - p) // 2 for ky in range(k)]  # floor division
    lo, hi = min(dys), max(dys)
    k2 = hi - lo + 1
    w2 = torch.zeros(cout, k2, k2, 4 * cp, dtype=torch.float32, device=w.device)
    for ky in range(k):
        dy, a = (ky - p) // 2, (ky - p) % 2
        for kx in range(k):
            dx, bb = (kx - p) // 2, (kx - p) % 2
            w2[:, dy - lo, dx - lo, (a * 2 + bb) * cp:(a * 2 + bb) * cp + cin] = w[:, :, ky, kx]
    return w2.to(torch.float16).contiguous(), b.contiguous(), k2, k2, -lo, -lo


def _s2d(x: torch.Tensor) -> torch.Tensor:
    """(N, H, W, C) -> (N, H/2, W/2, 4C), channel = (row phase * 2 + column phase) * C + c"""
    n, h, w, c = x.shape
    return x.view(n, h // 2, 2, w // 2, 2, c).permute(0, 1, 3, 2, 4, 5).reshape(n, h // 2, w // 2, 4 * c).contiguous()


def _conv(x, pk, act=ops.ACT_NONE, out=None, out_dtype=torch.float16):
    w, b, kh, kw, pt, pl = pk
    return ops.conv2d_taps(x, w, b, pad_top=pt, pad_left=pl, act=act, out=out, out_dtype=out_dtype)


# ------------------------------------------------------------------------------------------------
# RAFT
# ------------------------------------------------------------------------------------------------
class RAFT(PackedModule):
    """raft.py:24-143 (`small=False`, `alternate_corr=False`; `args.mixed_precision` is irrelevant: see module docstring)"""

    def __init__(self, args=None):
        super().__init__()
        if args is not None and getattr(args, "small", False):
            raise NotImplementedError("RAFT-small is not on the Upscale-A-Video path (raft_bi.py:24)")
        if args is not None and getattr(args, "alternate_corr", False):
            raise NotImplementedError("alternate_corr needs upstream's alt_cuda_corr; the reference sets it False (raft_bi.py:26)")
        self.args = args
        self.hidden_dim, self.context_dim = HDIM, CDIM
        self.fnet = BasicEncoder(output_dim=256, norm_fn="instance")
        self.cnet = BasicEncoder(output_dim=HDIM + CDIM, norm_fn="batch")
        self.update_block = BasicUpdateBlock(hidden_dim=HDIM)

    # ---- packing ----
    def _pack_encoder(self, enc: BasicEncoder, prefix: str, out_scale: float):
        pk = self._packed()
        bn = enc.norm_fn == "batch"

        def get(key, fn):
            return pk.tensor(prefix + key, fn)

        P = {"conv1": get("conv1", lambda: _pack_conv_s2(enc.conv1, enc.norm1 if bn else None)),
             "conv2": get("conv2", lambda: _pack_conv(enc.conv2, None, out_scale))}
        for ln in ("layer1", "layer2", "layer3"):
            for bi, blk in enumerate(getattr(enc, ln)):
                k = f"{ln}.{bi}."
                if blk.stride == 1:
                    P[k + "conv1"] = get(k + "conv1", lambda blk=blk: _pack_conv(blk.conv1, blk.norm1 if bn else None))
                else:
                    P[k + "conv1"] = get(k + "conv1", lambda blk=blk: _pack_conv_s2(blk.conv1, blk.norm1 if bn else None))
                    P[k + "down"] = get(k + "down", lambda blk=blk: _pack_conv_s2(blk.downsample[0], blk.norm3 if bn else None))
                P[k + "conv2"] = get(k + "conv2", lambda blk=blk: _pack_conv(blk.conv2, blk.norm2 if bn else None))
        return P

    def _pack_update(self):
        pk, ub = self._packed(), self.update_block

        def get(key, fn):
            return pk.tensor("ub." + key, fn)

        def cat_out(*convs):
            parts = [_pack_conv(c) for c in convs]
            return (torch.cat([p[0] for p in parts], 0).contiguous(), torch.cat([p[1] for p in parts]).contiguous(), *parts[0][2:])

        e, g = ub.encoder, ub.gru

        def convc1():
            w, b, *_ = _pack_conv(e.convc1)  # [256][1][1][328]
            return w.view(256, -1).contiguous(), b

        P = {"convc1": get("convc1", convc1), "convc2": get("convc2", lambda: _pack_conv(e.convc2)),
             "convf1": get("convf1", lambda: _pack_conv(e.convf1)), "convf2": get("convf2", lambda: _pack_conv(e.convf2)),
             "conv": get("conv", lambda: _pack_conv(e.conv, cout_pad=128)),
             "zr1": get("zr1", lambda: cat_out(g.convz1, g.convr1)), "q1": get("q1", lambda: _pack_conv(g.convq1)),
             "zr2": get("zr2", lambda: cat_out(g.convz2, g.convr2)), "q2": get("q2", lambda: _pack_conv(g.convq2)),
             "fh1": get("fh1", lambda: _pack_conv(ub.flow_head.conv1)), "fh2": get("fh2", lambda: _pack_conv(ub.flow_head.conv2)),
             "m0": get("m0", lambda: _pack_conv(ub.mask[0]))}

        def m2():  # `mask = .25 * self.mask(net)` (update.py:137) folded into the 1x1 conv
            w, b, *_ = _pack_conv(ub.mask[2], None, 0.25)
            return w.view(576, -1).contiguous(), b

        P["m2"] = get("m2", m2)
        return P

    # ---- encoders ----
    def _encode(self, enc: BasicEncoder, P, img: torch.Tensor) -> torch.Tensor:
        """img (N, H, W, 8) fp16 -> (N, H/8, W/8, out) fp16"""
        inst = enc.norm_fn == "instance"
        act = ops.ACT_NONE if inst else ops.ACT_RELU  # BatchNorm is folded: ReLU rides the conv epilogue

        def norm_relu(t, relu=True):
            return ops.instnorm_relu(t, relu) if inst else t

        x = norm_relu(_conv(_s2d(img), P["conv1"], act))
        for ln in ("layer1", "layer2", "layer3"):
            for bi, blk in enumerate(getattr(enc, ln)):
                k = f"{ln}.{bi}."
                if blk.stride == 1:
                    src, skip = x, x
                else:
                    src = _s2d(x)
                    skip = norm_relu(_conv(src, P[k + "down"], ops.ACT_NONE), relu=False)
                y = norm_relu(_conv(src, P[k + "conv1"], act))
                y = norm_relu(_conv(y, P[k + "conv2"], act))
                x = ops.add_relu(skip, y)
        return _conv(x, P["conv2"])

    # ---- forward ----
    @torch.no_grad()
    def forward(self, image1: torch.Tensor, image2: torch.Tensor, iters: int = 12, flow_init=None, test_mode: bool = True):
        """raft.py:87-143: images (N, 3, H, W) in [-1, 1] on the GPU, H and W multiples of 8; returns
        (flow at 1/8 resolution (N, 2, H/8, W/8), convex-upsampled flow (N, 2, H, W)), both fp32"""
        if not test_mode:
            raise NotImplementedError("training mode (list of per-iteration predictions) is not on the sampling path")
        if not image1.is_cuda:
            raise RuntimeError("RAFT: expected CUDA tensors (no CPU fallback)")
        return self._forward_impl(image1, image2, iters, flow_init)

    def _forward_impl(self, image1, image2, iters, flow_init):
        n, _, H, W = image1.shape
        assert H % 8 == 0 and W % 8 == 0, "RAFT.forward: H and W must be multiples of 8 (RAFT_bi resizes first)"
        h8, w8 = H // 8, W // 8
        hw8, rows = h8 * w8, n * h8 * w8
        dev = image1.device
        Pf = self._pack_encoder(self.fnet, "fnet.", 0.25)  # (f1 / 4) . (f2 / 4) = f1 . f2 / sqrt(256)   (corr.py:60)
        Pc = self._pack_encoder(self.cnet, "cnet.", 1.0)
        Pu = self._pack_update()

        img = torch.zeros(2 * n, H, W, 8, dtype=torch.float16, device=dev)
        img[..., :3] = torch.cat([image1, image2], 0).permute(0, 2, 3, 1)
        fmap = self._encode(self.fnet, Pf, img)  # (2n, h8, w8, 256)
        # all-pairs correlation, one fp16 GEMM with fp32 output per image pair (corr.py:52-60)
        corr0 = torch.empty(rows, h8, w8, dtype=torch.float32, device=dev)
        for i in range(n):
            ops.linear(fmap[i].view(hw8, 256), fmap[n + i].view(hw8, 256), None,
                       out=corr0[i * hw8:(i + 1) * hw8].view(hw8, hw8))
        pyr = [corr0]
        for _ in range(CORR_LEVELS - 1):
            pyr.append(ops.avgpool2x2_f32(pyr[-1]))

        # GRU state / input buffers: HX = [h | inp | motion], RHX = [r*h | inp | motion]
        HX = torch.empty(rows, 384, dtype=torch.float16, device=dev)
        RHX = torch.empty(rows, 384, dtype=torch.float16, device=dev)
        cnet = self._encode(self.cnet, Pc, img[:n])  # (n, h8, w8, 256)
        ops.raft_split_tanh_relu(cnet.view(rows, 256), HX[:, :128], HX[:, 128:256], RHX[:, 128:256])

        ys, xs = torch.meshgrid(torch.arange(h8, device=dev), torch.arange(w8, device=dev), indexing="ij")
        coords0 = torch.stack([xs, ys], dim=-1).float().view(1, hw8, 2).repeat(n, 1, 1).view(rows, 2).contiguous()
        coords1 = coords0.clone()
        if flow_init is not None:
            coords1 += flow_init.permute(0, 2, 3, 1).reshape(rows, 2).float()
        flow16 = torch.zeros(rows, 8, dtype=torch.float16, device=dev)
        ops.raft_flow_update(coords1, None, h8, w8, flow16=flow16)
        corrfeat = torch.empty(rows, 328, dtype=torch.float16, device=dev)
        CF = torch.empty(n, h8, w8, 256, dtype=torch.float16, device=dev)
        HX4, RHX4 = HX.view(n, h8, w8, 384), RHX.view(n, h8, w8, 384)
        delta = torch.empty(n, h8, w8, 2, dtype=torch.float32, device=dev)
        R = ops.ACT_RELU
        for _ in range(iters):
            ops.raft_corr_lookup(pyr, coords1, corrfeat)
            # BasicMotionEncoder (update.py:88-98)
            cor = ops.linear(corrfeat, *Pu["convc1"], act=R).view(n, h8, w8, 256)
            _conv(cor, Pu["convc2"], R, out=CF[..., :192])
            flo = _conv(flow16.view(n, h8, w8, 8), Pu["convf1"], R)
            _conv(flo, Pu["convf2"], R, out=CF[..., 192:])
            _conv(CF, Pu["conv"], R, out=HX4[..., 256:])           # 126 channels + 2 zero columns ...
            ops.raft_flow_update(coords1, None, h8, w8, dst_a=HX[:, 382:])  # ... which hold the flow (update.py:98)
            ops.copy_channels(HX4[..., 256:], RHX4[..., 256:])
            # SepConvGRU (update.py:43-60): horizontal then vertical
            for s in ("1", "2"):
                zr = _conv(HX4, Pu["zr" + s], ops.ACT_SIGMOID).view(rows, 256)
                ops.raft_gru_rh(zr, HX[:, :128], RHX[:, :128])
                q = _conv(RHX4, Pu["q" + s], ops.ACT_TANH).view(rows, 128)
                ops.raft_gru_update(zr, q, HX[:, :128])
            fh = _conv(HX4[..., :128], Pu["fh1"], R)
            _conv(fh, Pu["fh2"], out=delta, out_dtype=torch.float32)
            ops.raft_flow_update(coords1, delta.view(rows, 2), h8, w8, flow16=flow16)  # F(t+1) = F(t) + delta
        m = _conv(HX4[..., :128], Pu["m0"], R)
        mask = ops.linear(m.view(rows, 256), *Pu["m2"])
        flow_up = ops.raft_convex_upsample(coords1, mask, n, h8, w8)
        flow_lo = (coords1 - coords0).view(n, h8, w8, 2).permute(0, 3, 1, 2).contiguous()
        return flow_lo, flow_up


def resize_flow_pytorch(flow, newh, neww):
    """raft_bi.py:11-16, including its quirk: the factors scale ROWS 0 and 1 of every channel, not the x / y channels;
    identity when the size does not change (every multiple-of-8 input, e.g. 320x576)"""
    oldh, oldw = flow.shape[-2:]
    if (oldh, oldw) == (newh, neww):
        return flow
    flow = F.interpolate(flow, (newh, neww), mode="bilinear")
    flow[:, :, 0] *= newh / oldh
    flow[:, :, 1] *= neww / oldw
    return flow


def initialize_RAFT(model_path: Optional[str] = "pretrained_models/raft-things.pth", device="cuda") -> RAFT:
    """raft_bi.py:19-33: loads `raft-things.pth` (keys carry the DataParallel `module.` prefix)"""
    model = RAFT()
    if model_path is not None:
        sd = torch.load(model_path, map_location="cpu")
        model.load_state_dict({k[len("module."):] if k.startswith("module.") else k: v for k, v in sd.items()})
    return model.to(device).eval()


class RAFT_bi(nn.Module):
    """raft_bi.py:35-104"""

    def __init__(self, model_path: Optional[str] = "weights/raft-things.pth", device="cuda"):
        super().__init__()
        self.fix_raft = initialize_RAFT(model_path, device=device)
        for p in self.fix_raft.parameters():
            p.requires_grad = False
        self.eval()

    @torch.no_grad()
    def forward(self, gt_local_frames: torch.Tensor, iters: int = 20) -> Tuple[torch.Tensor, torch.Tensor]:
        B, C, T, H, W = gt_local_frames.shape
        H_, W_ = int(ceil(H / 8) * 8), int(ceil(W / 8) * 8)
        frames = gt_local_frames.float()
        if (H_, W_) != (H, W):  # raft_bi.py:53 (trilinear with an unchanged T = per-frame bilinear); glue, not a hot kernel
            frames = F.interpolate(frames, (T, H_, W_), mode="trilinear")
        f1 = frames[:, :, :-1].permute(0, 2, 1, 3, 4).reshape(B * (T - 1), C, H_, W_)
        f2 = frames[:, :, 1:].permute(0, 2, 1, 3, 4).reshape(B * (T - 1), C, H_, W_)
        # both directions in ONE batched call: pairs (f1 -> f2) and (f2 -> f1)
        _, up = self.fix_raft(torch.cat([f1, f2], 0), torch.cat([f2, f1], 0), iters=iters, test_mode=True)
        fwd, bwd = up[:B * (T - 1)], up[B * (T - 1):]
        fwd, bwd = resize_flow_pytorch(fwd, H, W), resize_flow_pytorch(bwd, H, W)

        def back(x):
            return x.reshape(B, T - 1, 2, H, W).permute(0, 2, 1, 3, 4).contiguous()

        return back(fwd), back(bwd)

    @torch.no_grad()
    def forward_slicing(self, gt_local_frames: torch.Tensor, iters: int = 20):
        """raft_bi.py:71-104"""
        width = gt_local_frames.size(-1)
        clip = 12 if width <= 640 else 8 if width <= 720 else 4 if width <= 1280 else 2
        n = gt_local_frames.size(2)
        if n <= clip:
            return self.forward(gt_local_frames, iters=iters)
        fs: List[torch.Tensor] = []
        bs: List[torch.Tensor] = []
        for f in range(0, n, clip):
            end = min(n, f + clip)
            a, b = self.forward(gt_local_frames[:, :, (f if f == 0 else f - 1):end], iters=iters)
            fs.append(a)
            bs.append(b)
        return torch.cat(fs, dim=2), torch.cat(bs, dim=2)
