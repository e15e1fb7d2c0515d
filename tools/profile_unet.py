"""per-shape timing of one UNet forward at config 2 (B=2 = the two CFG halves of one clip, T=8, 320x576), as the pipeline
calls it (cfg_shared_input=True), plus A/B totals of the round-2 switches toggled in-process.  usage: profile_unet.py [--ab]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from upscale_a_video_b200 import UNetVideoModel, layers, ops
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


def run(tag, detail=False):
    for _ in range(2):
        unet(lat, 500, low, **kw)
    with ops.Profile() as prof:
        unet(lat, 500, low, **kw)
    bt = prof.by_tag()
    tot = sum(d["ms"] for d in bt.values())
    kinds = {}
    for (kind, _), d in bt.items():
        kinds[kind] = kinds.get(kind, 0.0) + d["ms"]
    print(f"[{tag}] total {tot:.1f} ms  " + "  ".join(f"{k} {v:.1f}" for k, v in sorted(kinds.items(), key=lambda kv: -kv[1])))
    if detail:
        for (kind, t), d in sorted(bt.items(), key=lambda kv: -kv[1]["ms"])[:60]:
            tf = d["flops"] / d["ms"] / 1e9 if d["flops"] else 0
            gb = d["bytes"] / d["ms"] / 1e6
            print(f"{d['ms']:8.2f} ms {100 * d['ms'] / tot:5.1f}%  x{d['launches']:3d}  {tf:7.0f} TF/s {gb:7.0f} GB/s  {kind:10s} {t}")
    return tot


run("default", detail=True)
if "--ab" in sys.argv:
    from upscale_a_video_b200 import unet_video
    layers.GN_STATS_LINEAR = False
    run("no statistics from Linear / 1x1 producers (UAV_GN_STATS_LINEAR=0)")
    layers.GN_STATS_LINEAR = True
    layers.VIRTUAL_CONCAT = False
    run("concat buffers, main branch in place (UAV_VIRTUAL_CONCAT=0)")
    layers.GN_STATS_LINEAR = False
    run("UAV_VIRTUAL_CONCAT=0 + UAV_GN_STATS_LINEAR=0")
    layers.GN_STATS_LINEAR = True
    layers.VIRTUAL_CONCAT = True
    unet_video.FUSED_CONV_OUT = False
    run("separate conv_norm_out / conv_out kernels (UAV_FUSED_CONV_OUT=0)")
    unet_video.FUSED_CONV_OUT = True
    ops.GN_FUSED_STATS = False
    run("GN statistics by their own pass (UAV_GN_FUSED_STATS=0)")
    ops.GN_FUSED_STATS = True
    run("default again")
