"""make_golden_clip.py — mint CLIP-text-encoder fixtures from the installed `transformers` CLIPTextModel (CPU).

TEST INFRASTRUCTURE; build container only:   python oracle/make_golden_clip.py   -> tests/golden/clip.pt
Weights are regenerated from (seed, key) by oracle/weights.py; two configurations: a tiny one and one with the width /
head layout of the SD-x4-upscaler text encoder (hidden 1024, 16 heads of 64, "gelu") cut to 2 layers."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from transformers import CLIPTextConfig, CLIPTextModel  # noqa: E402

from oracle.weights import make_state_dict  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SEED = 4242
CONFIGS = {
    "tiny_quick_gelu": dict(vocab_size=1000, hidden_size=64, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4,
                            max_position_embeddings=77, hidden_act="quick_gelu"),
    "wide_gelu": dict(vocab_size=2048, hidden_size=1024, intermediate_size=4096, num_hidden_layers=2, num_attention_heads=16,
                      max_position_embeddings=77, hidden_act="gelu"),
}


def main():
    cases = {}
    for name, kw in CONFIGS.items():
        m = CLIPTextModel(CLIPTextConfig(**kw)).eval()
        shapes = {k: list(v.shape) for k, v in m.state_dict().items()}
        m.load_state_dict(make_state_dict(shapes, SEED), strict=True)
        g = torch.Generator().manual_seed(len(name))
        ids = torch.randint(0, kw["vocab_size"], (2, 77), generator=g)
        ids[1, 40:] = kw["vocab_size"] - 1  # a padded prompt (eos repeated), as the tokenizer produces
        with torch.no_grad():
            out = m(ids)[0]
        cs = 4 if kw["hidden_size"] > 256 else 1  # the wide case stores every 4th column
        cases[name] = {"config": dict(kw, layer_norm_eps=1e-5), "shapes": shapes, "input_ids": ids, "col_stride": cs,
                       "last_hidden_state": out[..., ::cs].clone()}
    torch.save({"seed": SEED, "cases": cases}, os.path.join(OUT, "clip.pt"))
    print("wrote clip.pt", os.path.getsize(os.path.join(OUT, "clip.pt")), "bytes")


if __name__ == "__main__":
    main()
