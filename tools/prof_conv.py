import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from upscale_a_video_b200 import ops
NB, H, W, Cin, Cout = 16, 160, 288, 512, 512
x = torch.randn(NB, H, W, Cin, device="cuda").half()
w = (torch.randn(Cout, 3, 3, Cin, device="cuda") * 0.02).half()
b = torch.zeros(Cout, device="cuda")
out = torch.empty(NB, H, W, Cout, device="cuda", dtype=torch.float16)
for _ in range(3):
    ops.conv2d(x, w, b, out=out)
torch.cuda.synchronize()
