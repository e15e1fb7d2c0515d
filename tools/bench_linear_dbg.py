import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from upscale_a_video_b200 import ops
M, K, N = 737280, 512, 512
a = torch.randn(M, K, device="cuda").half()
w = (torch.randn(N, K, device="cuda") * 0.02).half()
b = torch.zeros(N, device="cuda")
out = torch.empty(M, N, device="cuda", dtype=torch.float16)
for _ in range(3): ops.linear(a, w, b, out=out)
torch.cuda.synchronize()
s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(20): ops.linear(a, w, b, out=out)
e.record(); torch.cuda.synchronize()
print(os.environ.get("UAV_IGEMM_DBG", "0"), "linear 512x512 ms", s.elapsed_time(e) / 20)
