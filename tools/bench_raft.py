"""RAFT_bi.forward_slicing at the BASELINE clip (8 frames 320x576, 20 iterations, both directions = 14 image pairs):
CUDA-event timing + per-kind breakdown (ops.Profile).  Random-init weights of the reference architecture."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.weights import make_state_dict
from oracle.raft_oracle import synth_clip
from upscale_a_video_b200 import build, ops
from upscale_a_video_b200.raft import RAFT, RAFT_bi

build.build()
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
m = RAFT()
m.load_state_dict(make_state_dict(json.load(open(os.path.join(G, "shapes_raft.json"))), 777), strict=True)
bi = RAFT_bi(model_path=None)
bi.fix_raft = m.cuda().eval()
frames = synth_clip(8, 320, 576, 4).cuda()
for _ in range(2):
    bi.forward_slicing(frames, iters=20)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(3):
    f, b = bi.forward_slicing(frames, iters=20)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / 3
with ops.Profile() as prof:
    bi.forward_slicing(frames, iters=20)
summ = prof.summary()
print(json.dumps({"name": "RAFT_bi.forward_slicing 8x320x576, 20 iters, 14 pairs", "ms": ms, "flows_per_s": 14 / (ms / 1e3),
                  "igemm_ms": summ.get("igemm", {}).get("ms"), "igemm_tflops": (summ["igemm"]["flops"] / summ["igemm"]["ms"] / 1e9)
                  if "igemm" in summ else None, "kinds": {k: round(v["ms"], 2) for k, v in summ.items()}}))
