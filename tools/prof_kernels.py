"""One kernel shape per invocation, for `ncu --set full` captures (tools/run_ncu_round2.sh).  usage: prof_kernels.py <which>"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from upscale_a_video_b200 import ops

which = sys.argv[1]
dev = "cuda"


def rnd(*s, scale=1.0):
    return (torch.randn(*s, device=dev) * scale).half()


if which == "geglu":       # FeedForward net.0 at the H/2 level (quarter of the rows: same per-tile behaviour)
    a, w, b = rnd(2, 4, 23040, 512), rnd(4096, 512, scale=0.02), torch.zeros(4096, device=dev)
    f = lambda: ops.linear(a, w, b, act=ops.ACT_GEGLU)
elif which == "linear_res":  # attention to_out / proj_out: Linear 512->512 + residual, M = 737 280
    a, w, b, r = rnd(2, 8, 46080, 512), rnd(512, 512, scale=0.02), torch.zeros(512, device=dev), rnd(2, 8, 46080, 512)
    out = torch.empty_like(r)
    f = lambda: ops.linear(a, w, b, residual=r, out=out)
elif which == "linear_plain":  # attention to_q / text-attention to_q: Linear 512->512, no residual, M = 737 280
    a, w, b = rnd(2, 8, 46080, 512), rnd(512, 512, scale=0.02), torch.zeros(512, device=dev)
    out = torch.empty(2, 8, 46080, 512, device=dev, dtype=torch.float16)
    f = lambda: ops.linear(a, w, b, out=out)
elif which == "linear_res_gn":
    a, w, b, r = rnd(2, 8, 46080, 512), rnd(512, 512, scale=0.02), torch.zeros(512, device=dev), rnd(2, 8, 46080, 512)
    out = torch.empty_like(r)
    f = lambda: ops.linear(a, w, b, residual=r, out=out, gn_stats=True)
elif which == "conv_gn":     # conv3x3 512->512 on 16 x 160x288 with the GroupNorm statistics epilogue
    x, w, b = rnd(2, 8, 160, 288, 512), rnd(512, 3, 3, 512, scale=0.02), torch.zeros(512, device=dev)
    out = torch.empty(2, 8, 160, 288, 512, device=dev, dtype=torch.float16)
    f = lambda: ops.conv2d(x, w, b, out=out, gn_stats=True)
elif which == "conv":        # the representative launch of bench.py's roofline block
    x, w, b = rnd(16, 160, 288, 512), rnd(512, 3, 3, 512, scale=0.02), torch.zeros(512, device=dev)
    out = torch.empty(16, 160, 288, 512, device=dev, dtype=torch.float16)
    f = lambda: ops.conv2d(x, w, b, out=out)
elif which == "gn_fused":    # GroupNorm from producer statistics (reduce + apply) on the 2 949 120-pixel, 256-channel level
    x = ops.conv2d(rnd(2, 8, 320, 576, 64), rnd(256, 3, 3, 64, scale=0.05), None, gn_stats=True)
    g, bt = torch.ones(256, device=dev), torch.zeros(256, device=dev)
    f = lambda: ops.group_norm(x, g, bt, 32, 1e-5, silu=True, n_outer=2, stats=x.uav_gn, batch=2)
elif which == "gn_plain":
    x = rnd(2, 8, 320, 576, 256)
    g, bt = torch.ones(256, device=dev), torch.zeros(256, device=dev)
    f = lambda: ops.group_norm(x, g, bt, 32, 1e-5, silu=True, n_outer=2)
elif which == "fa512":       # VAE mid-block attention at a quarter of the config-2 token count
    n, C = 46080, 512
    qkv = rnd(1, n, 3 * C)
    out = torch.empty(1, n, C, device=dev, dtype=torch.float16)
    f = lambda: ops.attention(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], 1, out=out)
elif which == "fa128":       # UNet spatial self-attention at the H/8 level
    qkv = rnd(16, 2880, 3 * 1024)
    out = torch.empty(16, 2880, 1024, device=dev, dtype=torch.float16)
    f = lambda: ops.attention(qkv[..., :1024], qkv[..., 1024:2048], qkv[..., 2048:], 8, out=out)
elif which == "cross":       # text cross-attention at the H/2 level
    q, kv = rnd(16, 46080, 512), rnd(2, 77, 1024)
    out = torch.empty_like(q)
    f = lambda: ops.attention(q, kv[..., :512], kv[..., 512:], 8, kv_batch_div=8, out=out)
else:
    raise SystemExit(f"unknown kernel {which}")
for _ in range(3):
    f()
torch.cuda.synchronize()
