"""per-shape timing of one UNet forward at config 2 (B=2,T=8,320x576): where the 400 ms go"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from upscale_a_video_b200 import ops
dev = torch.device("cuda")
from upscale_a_video_b200 import UNetVideoModel
cfg = json.load(open("upscale_a_video_b200/configs/unet_video_config.json"))
unet = UNetVideoModel.from_config(cfg)
unet.load_state_dict(bench.seeded_state_dict(unet, 1234)); unet = unet.half().eval().to(dev)
lat = torch.randn(2, 4, 8, 320, 576, device=dev, dtype=torch.float16)
low = torch.randn(2, 3, 8, 320, 576, device=dev, dtype=torch.float16)
ctx = (torch.randn(2, 77, 1024, device=dev) * 0.3).half()
for _ in range(2):
    unet(lat, 500, low, encoder_hidden_states=ctx, class_labels=torch.tensor([120]))
with ops.Profile() as prof:
    unet(lat, 500, low, encoder_hidden_states=ctx, class_labels=torch.tensor([120]))
bt = prof.by_tag()
tot = sum(d["ms"] for d in bt.values())
print(f"total {tot:.1f} ms")
for (kind, tag), d in sorted(bt.items(), key=lambda kv: -kv[1]["ms"])[:45]:
    tf = d["flops"] / d["ms"] / 1e9 if d["flops"] else 0
    gb = d["bytes"] / d["ms"] / 1e6
    print(f"{d['ms']:8.2f} ms {100*d['ms']/tot:5.1f}%  x{d['launches']:3d}  {tf:7.0f} TF/s {gb:7.0f} GB/s  {kind:10s} {tag}")
