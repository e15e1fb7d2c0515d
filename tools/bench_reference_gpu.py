"""The reference's own GPU execution of one UNet forward at config 2 (informative; diffusers/xformers are not installed,
so this runs the oracle = the same torch op sequence as the reference modules, fp16, cuDNN/cuBLAS, on this GPU)."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import uav_oracle as O
from oracle.weights import make_state_dict
cfg = json.load(open("upscale_a_video_b200/configs/unet_video_config.json"))
shapes = json.load(open("tests/golden/shapes_unet.json"))
sd = {k: v.cuda().half() for k, v in make_state_dict(shapes, 1234).items()}
B, T, H, W = 2, 8, 320, 576
lat = torch.randn(B, 4, T, H, W, device="cuda", dtype=torch.float16)
low = torch.randn(B, 3, T, H, W, device="cuda", dtype=torch.float16)
ctx = (torch.randn(B, 77, 1024, device="cuda") * 0.3).half()
torch.backends.cudnn.benchmark = True
with torch.no_grad():
    for _ in range(2):
        O.unet_forward(sd, cfg, lat, torch.tensor(500), low, ctx, torch.tensor([120], device="cuda"))
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(3):
        O.unet_forward(sd, cfg, lat, torch.tensor(500), low, ctx, torch.tensor([120], device="cuda"))
    e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 3
print(json.dumps({"what": "reference torch fp16 UNet forward B=2,T=8,320x576 (cuDNN/cuBLAS, materialised attention)",
                  "ms": ms, "tflops_alg": 319.96 / ms * 1e3, "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9}))
