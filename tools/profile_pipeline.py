"""Whole-clip profile at BASELINE config 2 (8 frames 320x576 -> 4x, 30 DDIM steps): CUDA-event time per kernel kind
and per shape, split into the DDIM loop (UNet) and the decode (VAE).  `python tools/profile_pipeline.py`"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from upscale_a_video_b200 import ops

dev = torch.device("cuda")
pipe = bench.build_pipeline(dev)
T = 8
C2 = bench.CONFIGS["c2"]
image, fw, bw, pe = bench.synth_inputs(T, C2["h"], C2["w"], dev)
neg, pos = pe.half().to(dev).chunk(2)
kw = dict(num_inference_steps=C2["steps"], guidance_scale=bench.GUIDANCE, noise_level=bench.NOISE_LEVEL,
          propagation_steps=list(C2["prop"]), prompt_embeds=pos, negative_prompt_embeds=neg)


def run():
    gen = torch.Generator(device=dev).manual_seed(10)
    return pipe(None, image=image.to(dev), flows_bi=[fw.to(dev), bw.to(dev)], generator=gen, **kw).images


run()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
run()
e.record()
torch.cuda.synchronize()
print(f"clip wall (device events): {s.elapsed_time(e):.1f} ms")

# decode separately
orig_decode = pipe.decode_latents_vsr if hasattr(pipe, "decode_latents_vsr") else None
with ops.Profile() as prof:
    run()
bt = prof.by_tag()
tot = sum(d["ms"] for d in bt.values())
print(f"sum of profiled kernels {tot:.1f} ms over {sum(d['launches'] for d in bt.values())} launches")
kinds = {}
for (kind, tag), d in bt.items():
    kinds[kind] = kinds.get(kind, 0.0) + d["ms"]
for k, ms in sorted(kinds.items(), key=lambda kv: -kv[1]):
    print(f"  {ms:9.1f} ms {100 * ms / tot:5.1f}%  {k}")
print("top shapes:")
for (kind, tag), d in sorted(bt.items(), key=lambda kv: -kv[1]["ms"])[:40]:
    tf = d["flops"] / d["ms"] / 1e9 if d["flops"] else 0
    gb = d["bytes"] / d["ms"] / 1e6
    print(f"{d['ms']:9.1f} ms {100 * d['ms'] / tot:5.1f}%  x{d['launches']:5d}  {tf:7.0f} TF/s {gb:7.0f} GB/s  {kind:10s} {tag}")
