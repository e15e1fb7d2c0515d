"""Linear K=512 -> N=512 at M = 737 280 (the most frequent launch of a UNet forward: 39 of 319) alone, with and without
the residual operand / statistics epilogue, inputs rotated past L2.  Prints ms, TFLOP/s and the HBM rate of the
algorithmic bytes.  Env knobs of igemm.cu (UAV_IGEMM_RES_MODE, ...) apply: run once per setting."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from upscale_a_video_b200 import ops

dev = torch.device("cuda")
print({k: v for k, v in os.environ.items() if k.startswith("UAV_")})


def run(M, K, N, residual, gn_stats, act=0, nbuf=3, iters=12):
    a = [torch.randn(16, M // 16, K, device=dev).half() for _ in range(nbuf)]
    w = (torch.randn(N, K, device=dev) * 0.02).half()
    b = torch.zeros(N, device=dev)
    n_out = N // 2 if act == 2 else N
    outs = [torch.empty(16, M // 16, n_out, device=dev, dtype=torch.float16) for _ in range(nbuf)]
    res = [torch.randn(16, M // 16, n_out, device=dev).half() for _ in range(nbuf)] if residual else None

    def one(i):
        ops.linear(a[i % nbuf], w, b, out=outs[i % nbuf], residual=res[i % nbuf] if residual else None, act=act,
                   gn_stats=gn_stats)
    for i in range(3):
        one(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        one(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    gb = 2.0 * (M * K + M * n_out * (2 if residual else 1) + N * K) / 1e9
    print(f"linear M{M} K{K} N{N} act{act} residual={int(residual)} gn_stats={int(gn_stats)}: {ms * 1000:7.1f} us  "
          f"{2.0 * M * K * N / ms / 1e9:6.0f} TF/s  {gb / ms * 1000:6.0f} GB/s")


for residual, gn in ((False, False), (True, False), (True, True), (False, True)):
    run(737280, 512, 512, residual, gn)
run(737280, 512, 1536, False, False)
run(737280, 2048, 512, True, False)
run(737280, 512, 4096, False, False, act=2)
run(184320, 512, 512, True, False, nbuf=8, iters=32)
run(46080, 1024, 1024, True, False, nbuf=16, iters=64)
