import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from upscale_a_video_b200 import ops
# VAE mid-block attention at a quarter of the config-2 token count (keeps the ncu replay short) + temporal attention
n, C = 46080, 512
qkv = torch.randn(1, n, 3 * C, device="cuda").half()
out = torch.empty(1, n, C, device="cuda", dtype=torch.float16)
for _ in range(2):
    ops.attention(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], 1, out=out)
B, Fr, HW, heads, d = 2, 8, 160 * 288, 8, 64
x = torch.randn(B, Fr, HW, 3 * C, device="cuda").half()
freqs = 1.0 / (10000 ** (torch.arange(0, 32, 2).float() / 32))
ang = torch.arange(Fr).float()[:, None] * freqs[None, :]
rot = torch.stack([ang.cos(), ang.sin()], dim=-1).contiguous().cuda()
bias = (torch.randn(heads, Fr, Fr) * 0.3).cuda()
o2 = torch.empty(B, Fr, HW, C, device="cuda", dtype=torch.float16)
for _ in range(2):
    ops.temporal_attention(x[..., :C], x[..., C:2 * C], x[..., 2 * C:], heads, rot, bias, out=o2)
torch.cuda.synchronize()
