"""stage-wise parity: product UNet (fp16 CUDA kernels) vs the oracle run with fp16 torch ops on the same GPU"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import uav_oracle as O
from oracle.weights import make_state_dict
from upscale_a_video_b200 import UNetVideoModel
G = "tests/golden"
case = sys.argv[1] if len(sys.argv) > 1 else "t2_20x28_upsize"
cfg = json.load(open("upscale_a_video_b200/configs/unet_video_config.json"))
shapes = json.load(open(f"{G}/shapes_unet.json"))
sd = make_state_dict(shapes, 1234)
m = UNetVideoModel.from_config(cfg); m.load_state_dict(sd); m = m.half().eval().cuda()
c = torch.load(f"{G}/unet.pt", weights_only=False)[case]
sample, low, ctx = c["sample"].cuda().half(), c["low_res"].cuda().half(), c["ctx"].cuda().half()
taps = {}
m.__dict__["_debug_taps"] = taps
out = m(sample, torch.tensor(c["timestep"]), low, encoder_hidden_states=ctx, class_labels=c["class_labels"].cuda()).sample
rtaps = {}
sd32 = {k: v.cuda() for k, v in sd.items()}
ref = O.unet_forward(sd32, cfg, sample.float(), torch.tensor(c["timestep"]), low.float(), ctx.float(), c["class_labels"], taps=rtaps)
for k, v in rtaps.items():
    mine = taps[k].permute(0, 4, 1, 2, 3).float()
    mine = mine[:, : v.shape[1]]
    print(f"{k:12s} shape {tuple(v.shape)} rel err {((mine - v).norm() / v.norm()).item():.3e}")
print("final", ((out.float() - ref).norm() / ref.norm()).item())
