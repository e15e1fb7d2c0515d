import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from upscale_a_video_b200 import ops
def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup): fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters
for (B, T, H, W, C) in [(2, 8, 320, 576, 256), (2, 8, 320, 576, 768), (2, 8, 160, 288, 512), (2, 8, 40, 72, 1024)]:
    x = torch.randn(B, T, H, W, C, device="cuda").half()
    g = torch.ones(C, device="cuda"); b = torch.zeros(C, device="cuda")
    out = torch.empty_like(x)
    ms = timeit(lambda: ops.group_norm(x, g, b, 32, 1e-5, silu=True, n_outer=B, out=out))
    print(json.dumps({"shape": [B, T, H, W, C], "ms": ms, "GBps_3pass": 6.0 * x.numel() / ms / 1e6}))
