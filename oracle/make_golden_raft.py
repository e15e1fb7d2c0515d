"""make_golden_raft.py — mint RAFT fixtures by running the UNMODIFIED reference RAFT (models_video/RAFT) on CPU.

TEST INFRASTRUCTURE; build container only (needs /root/reference + oracle/shims):

    python oracle/make_golden_raft.py        -> tests/golden/raft.pt, tests/golden/shapes_raft.json

Neither weights nor input clips are stored: the clips come from `oracle.raft_oracle.synth_clip(T, H, W, seed)` and every tensor of the state dict is regenerated from (seed, key) by oracle/weights.py.  The
reference's `RAFT_bi.__init__` loads `raft-things.pth` (absent offline), so the module is assembled the way
`initialize_RAFT` does (raft_bi.py:19-33) minus the checkpoint load, and `RAFT_bi.forward` / `forward_slicing` are called
unbound on a stand-in object holding `fix_raft`.  Frames must be >= 128 px on each side: the 4-level correlation pyramid of
smaller inputs ends in a 1x1 level and `bilinear_sampler` then divides by (W - 1) = 0 (reference behaviour: NaN flows).
The larger fixtures store every `stride`-th pixel of the flows."""
import argparse
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle", "shims"), "/root/reference", ROOT]

from models_video.RAFT.raft import RAFT  # noqa: E402
from models_video.RAFT import raft_bi as ref_bi  # noqa: E402

from oracle.weights import make_state_dict  # noqa: E402
from oracle.raft_oracle import synth_clip  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SEED = 777


def build_reference():
    args = argparse.ArgumentParser()  # the reference uses the parser object itself as its namespace (raft_bi.py:22-26)
    args.small = False
    args.mixed_precision = False
    args.alternate_corr = False
    model = RAFT(args).eval()
    shapes = {k: list(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict(make_state_dict(shapes, SEED), strict=True)
    return model, shapes


def main():
    model, shapes = build_reference()
    json.dump(shapes, open(os.path.join(OUT, "shapes_raft.json"), "w"), indent=0, sort_keys=True)
    holder = types.SimpleNamespace(fix_raft=model)
    holder.forward = lambda frames, iters=20: ref_bi.RAFT_bi.forward(holder, frames, iters=iters)
    cases = {}
    with torch.no_grad():
        # (a) one RAFT call (raft.py:87-143): low-resolution and upsampled flow after 3 iterations
        clip = synth_clip(2, 128, 136, 1)
        lo, up = model(clip[0, :, 0][None], clip[0, :, 1][None], iters=3, test_mode=True)
        cases["raft_128x136_it3"] = {"clip": [2, 128, 136, 1], "iters": 3, "flow_lo": lo, "flow_up": up}
        # (b) RAFT_bi.forward on a size that is NOT a multiple of 8 (trilinear resize + the resize_flow_pytorch quirk)
        clip = synth_clip(3, 124, 132, 2)
        f, b = ref_bi.RAFT_bi.forward(holder, clip, iters=2)
        cases["bi_124x132_it2"] = {"clip": [3, 124, 132, 2], "iters": 2, "stride": 2, "fwd": f[..., ::2, ::2].clone(), "bwd": b[..., ::2, ::2].clone()}
        # (c) forward_slicing with more frames than one short clip (width <= 640 -> 12 frames per clip)
        clip = synth_clip(13, 128, 128, 3)
        f, b = ref_bi.RAFT_bi.forward_slicing(holder, clip, iters=1)
        cases["slicing_13f_128x128_it1"] = {"clip": [13, 128, 128, 3], "iters": 1, "stride": 4, "fwd": f[..., ::4, ::4].clone(), "bwd": b[..., ::4, ::4].clone()}
        # (d) one pair at the BASELINE.json frame size (320x576 -> 40x72 grid, 2880^2 correlation volume), every 8th pixel stored
        clip = synth_clip(2, 320, 576, 4)
        lo, up = model(clip[0, :, 0][None], clip[0, :, 1][None], iters=4, test_mode=True)
        cases["raft_320x576_it4"] = {"clip": [2, 320, 576, 4], "iters": 4, "stride": 8, "flow_lo": lo,
                                     "flow_up": up[..., ::8, ::8].clone()}
    torch.save({"seed": SEED, "cases": cases}, os.path.join(OUT, "raft.pt"))
    print("wrote raft.pt", os.path.getsize(os.path.join(OUT, "raft.pt")), "bytes;", len(shapes), "tensors in the state dict")


if __name__ == "__main__":
    main()
