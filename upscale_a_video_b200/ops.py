"""Thin torch-tensor wrappers over the C ABI (include/uav_b200.h).

All activations are channels-last fp16 CUDA tensors: a reference "b c t h w" tensor is held as
(b, t, h, w, c).  A tensor may be a channel slice `buf[..., c0:c1]` of a wider buffer (its pixel
stride `ld` is then the buffer's channel count).  PyTorch is used for memory and streams only.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import Epilogue

ACT_NONE, ACT_SILU, ACT_GEGLU = 0, 1, 2
F16, F32 = 0, 1


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _pixel_ld(x: torch.Tensor) -> int:
    """element distance between consecutive pixels; validates the channels-last layout"""
    assert x.is_cuda, "uav_b200 ops need CUDA tensors (no CPU fallback)"
    assert x.stride(-1) == 1 or x.shape[-1] == 1, "channel dim must be contiguous"
    if x.dim() == 1:
        return x.shape[0]
    ld = x.stride(-2)
    exp = ld
    for d in range(x.dim() - 2, -1, -1):
        if x.shape[d] != 1:
            assert x.stride(d) == exp, f"tensor is not a dense channels-last view: {x.shape} {x.stride()}"
        exp *= x.shape[d]
    return ld


def _epi(out: torch.Tensor, bias=None, rowvec=None, rows_per_vec=0, residual=None, act=ACT_NONE) -> Epilogue:
    e = Epilogue()
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous()
        e.bias = bias.data_ptr()
    if rowvec is not None:
        assert rowvec.dtype == torch.float16 and rowvec.stride(-1) == 1
        e.rowvec = rowvec.data_ptr()
        e.rows_per_vec = rows_per_vec
        e.ld_rowvec = rowvec.stride(0) if rowvec.dim() > 1 else rowvec.shape[0]
    if residual is not None:
        assert residual.dtype == torch.float16
        e.residual = residual.data_ptr()
        e.ld_res = _pixel_ld(residual)
    e.act = act
    e.out_dtype = F16 if out.dtype == torch.float16 else F32
    assert out.dtype in (torch.float16, torch.float32)
    e.ld_out = _pixel_ld(out)
    return e


def linear(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *, out=None,
           residual=None, rowvec=None, rows_per_vec=0, act=ACT_NONE, out_dtype=torch.float16):
    """out[..., N] = epilogue(a[..., K] @ w[N, K]^T); a fp16 (rows may be a channel-slice view)."""
    assert a.dtype == torch.float16 and w.dtype == torch.float16 and w.is_contiguous()
    K = a.shape[-1]
    N = w.shape[0]
    assert w.shape[1] == K
    M = a.numel() // K
    n_out = N // 2 if act == ACT_GEGLU else N
    if out is None:
        out = torch.empty(*a.shape[:-1], n_out, dtype=out_dtype, device=a.device)
    assert out.shape[-1] == n_out and out.numel() // n_out == M
    e = _epi(out, bias, rowvec, rows_per_vec, residual, act)
    lib = _lib.load()
    _lib.check(lib.uav_linear(a.data_ptr(), M, K, _pixel_ld(a) if a.dim() > 1 else K, w.data_ptr(), N,
                              out.data_ptr(), C.byref(e), _stream()), "uav_linear")
    return out


def conv2d(x: torch.Tensor, w: torch.Tensor, bias=None, *, stride=1, pad_mode=0, out=None, residual=None,
           rowvec=None, rows_per_vec=0, act=ACT_NONE, out_dtype=torch.float16):
    """x: (..., H, W, Cin) channels-last fp16 (leading dims = images); w: (Cout, k, k, Cin) fp16."""
    assert x.dtype == torch.float16 and w.dtype == torch.float16 and w.is_contiguous()
    *lead, H, W, Cin = x.shape
    Cout, k, k2, Cin_w = w.shape
    assert k == k2 and Cin_w == Cin
    NB = 1
    for d in lead:
        NB *= d
    Ho, Wo = (H, W) if stride == 1 else (H // 2, W // 2)
    if out is None:
        out = torch.empty(*lead, Ho, Wo, Cout, dtype=out_dtype, device=x.device)
    assert tuple(out.shape) == (*lead, Ho, Wo, Cout), (out.shape, (*lead, Ho, Wo, Cout))
    e = _epi(out, bias, rowvec, rows_per_vec, residual, act)
    lib = _lib.load()
    _lib.check(lib.uav_conv2d(x.data_ptr(), NB, H, W, Cin, _pixel_ld(x), w.data_ptr(), Cout, k, stride,
                              pad_mode, out.data_ptr(), C.byref(e), _stream()), "uav_conv2d")
    return out


def conv_temporal(x: torch.Tensor, w: torch.Tensor, bias=None, *, out=None, residual=None, rowvec=None,
                  rows_per_vec=0, act=ACT_NONE, out_dtype=torch.float16):
    """x: (B, T, H, W, Cin); w: (Cout, k, Cin) — nn.Conv3d (k,1,1), zero padding (k-1)/2 in t."""
    assert x.dtype == torch.float16 and w.dtype == torch.float16 and w.is_contiguous()
    B, T, H, W, Cin = x.shape
    Cout, k, Cin_w = w.shape
    assert Cin_w == Cin
    if out is None:
        out = torch.empty(B, T, H, W, Cout, dtype=out_dtype, device=x.device)
    e = _epi(out, bias, rowvec, rows_per_vec, residual, act)
    lib = _lib.load()
    _lib.check(lib.uav_conv_temporal(x.data_ptr(), B, T, H * W, Cin, _pixel_ld(x), w.data_ptr(), Cout, k,
                                     out.data_ptr(), C.byref(e), _stream()), "uav_conv_temporal")
    return out


def conv3d(x: torch.Tensor, w: torch.Tensor, bias=None, *, out=None, residual=None, act=ACT_NONE,
           out_dtype=torch.float16):
    """x: (B, T, H, W, Cin); w: (Cout, 3, 3, 3, Cin) — nn.Conv3d 3x3x3, padding 1."""
    assert x.dtype == torch.float16 and w.dtype == torch.float16 and w.is_contiguous()
    B, T, H, W, Cin = x.shape
    Cout = w.shape[0]
    assert tuple(w.shape) == (Cout, 3, 3, 3, Cin)
    if out is None:
        out = torch.empty(B, T, H, W, Cout, dtype=out_dtype, device=x.device)
    e = _epi(out, bias, None, 0, residual, act)
    lib = _lib.load()
    _lib.check(lib.uav_conv3d(x.data_ptr(), B, T, H, W, Cin, _pixel_ld(x), w.data_ptr(), Cout,
                              out.data_ptr(), C.byref(e), _stream()), "uav_conv3d")
    return out
