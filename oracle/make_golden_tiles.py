"""Mint golden tile plans by EXECUTING the reference's own tile loop (inference_upscale_a_video.py:207-304) — read from
the reference checkout at run time, never copied into this repo — with a recording stub in place of the pipeline.

    python oracle/make_golden_tiles.py        (build container only; writes tests/golden/tiles.json)

For each (h, w, tile_size) the stub pipeline returns the nearest-x4 upsampling of the padded input tile, so the pasted
output must equal the nearest-x4 upsampling of the whole frame iff the paste geometry is right; every input slice and
paste box is recorded."""
import json
import math
import os
import textwrap
import types

import torch

REF = "/root/reference/inference_upscale_a_video.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "tiles.json")


def reference_tile_block():
    lines = open(REF).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.strip() == "if args.perform_tile:")
    end = next(i for i in range(start, len(lines)) if lines[i].strip() == "else:" and lines[i].startswith("        else:"))
    body = "\n".join(lines[start + 1:end])
    return textwrap.dedent(body)


def run_case(h, w, tile_size, t=2):
    code = reference_tile_block()
    vframes = torch.arange(h * w, dtype=torch.float32).reshape(1, 1, 1, h, w).repeat(1, 3, t, 1, 1)
    rec = []

    class Out:
        def __init__(self, images):
            self.images = images

    def pipeline(prompt, image=None, flows_bi=None, **kw):
        rec.append({"in_shape": list(image.shape[-2:]), "first": float(image[0, 0, 0, 0, 0]), "last": float(image[0, 0, 0, -1, -1])})
        return Out(image.repeat_interleave(4, dim=-2).repeat_interleave(4, dim=-1))

    args = types.SimpleNamespace(tile_size=tile_size, inference_steps=1, guidance_scale=1.0, noise_level=0, n_prompt="",
                                 propagation_steps=[])
    env = dict(args=args, vframes=vframes, b=1, c=3, t=t, h=h, w=w, math=math, torch=torch, pipeline=pipeline,
               flows_bi=None, prompt="", generator=None, index_str="", print=lambda *a, **k: None)
    exec(code, env)
    output = env["output"]
    expect = vframes.repeat_interleave(4, dim=-2).repeat_interleave(4, dim=-1)
    # recover each tile's input box from the recorded corner values
    tiles = []
    for r in rec:
        y0, x0 = divmod(int(r["first"]), w)
        y1, x1 = divmod(int(r["last"]), w)
        tiles.append([y0, y1 + 1, x0, x1 + 1])
    return {"h": h, "w": w, "tile_size": tile_size, "tiles_in": tiles, "paste_exact": bool(torch.equal(output, expect)),
            "tiles_x": env["tiles_x"], "tiles_y": env["tiles_y"]}


if __name__ == "__main__":
    cases = []
    for (h, w) in [(320, 576), (540, 960), (384, 384), (400, 400), (256, 320), (320, 320), (321, 577), (180, 320), (720, 1280),
                   (300, 500), (512, 512), (513, 770), (64, 64)]:
        for ts in (256, 320):
            cases.append(run_case(h, w, ts))
    json.dump(cases, open(OUT, "w"))
    print(len(cases), "cases;", sum(c["paste_exact"] for c in cases), "paste-exact")
