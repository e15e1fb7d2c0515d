"""One UNet forward at config 2 (B=2, T=8, 320x576, cfg_shared_input) inside a cudaProfilerStart/Stop window, for
`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv` (the launch list of the step the
roofline in bench.py is quoted on).  Numbers printed under ncu are not bench values."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from upscale_a_video_b200 import UNetVideoModel
from upscale_a_video_b200.synthetic import seeded_state_dict

dev = torch.device("cuda")
cfg = json.load(open(os.path.join(os.path.dirname(__file__), "..", "upscale_a_video_b200", "configs", "unet_video_config.json")))
unet = UNetVideoModel.from_config(cfg)
unet.load_state_dict(seeded_state_dict(unet, 1234))
unet = unet.half().eval().to(dev)
lat = torch.randn(1, 4, 8, 320, 576, device=dev, dtype=torch.float16).repeat(2, 1, 1, 1, 1)
low = torch.randn(1, 3, 8, 320, 576, device=dev, dtype=torch.float16).repeat(2, 1, 1, 1, 1)
ctx = (torch.randn(2, 77, 1024, device=dev) * 0.3).half()
kw = dict(encoder_hidden_states=ctx, class_labels=torch.tensor([120]), cfg_shared_input=True)
unet(lat, 500, low, **kw)
torch.cuda.synchronize()
torch.cuda.profiler.start()
unet(lat, 500, low, **kw)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
