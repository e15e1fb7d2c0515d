"""GPU parity of UNetVideoModel.forward (full 691 M-parameter config) against the golden vectors minted from the
unmodified reference (fp32 CPU).  The product computes in fp16 (fp32 accumulate); the acceptance band is tied to
what the reference's own fp16 execution (same torch ops, cuDNN/cuBLAS, run here through the oracle on this GPU)
deviates from its fp32 result: the B200 path must be no worse than 1.5x that drift (+ a small floor), and both
numbers are printed.  A tight elementwise rtol=1e-3 after ~130 stacked fp16 layers is not meaningful even for the
reference against itself."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


@pytest.fixture(scope="module")
def unet(uav_lib):
    from oracle.weights import make_state_dict
    from upscale_a_video_b200.unet_video import UNetVideoModel
    meta = json.load(open(os.path.join(G, "meta.json")))
    cfg = json.load(open(os.path.join(os.path.dirname(__file__), "..", "upscale_a_video_b200", "configs",
                                      "unet_video_config.json")))
    shapes = json.load(open(os.path.join(G, "shapes_unet.json")))
    sd = make_state_dict(shapes, meta["seed_unet"])
    m = UNetVideoModel.from_config(cfg)
    assert {k: list(v.shape) for k, v in m.state_dict().items()} == shapes  # drop-in: identical keys and shapes
    m.load_state_dict(sd, strict=True)
    return m.half().eval().cuda(), sd, cfg


@pytest.mark.parametrize("case", ["t3_16x24", "t2_20x28_upsize", "t8_8x8"])
def test_unet_forward_vs_golden(unet, case):
    from oracle import uav_oracle as O
    m, sd, cfg = unet
    c = torch.load(os.path.join(G, "unet.pt"), weights_only=False)[case]
    sample, low, ctx = c["sample"].cuda().half(), c["low_res"].cuda().half(), c["ctx"].cuda().half()
    out = m(sample, torch.tensor(c["timestep"]), low, encoder_hidden_states=ctx, class_labels=c["class_labels"].cuda()).sample
    assert out.shape == c["out"].shape and out.dtype == torch.float16
    torch.cuda.synchronize()
    err = _rel(out.cpu(), c["out"])
    # the reference's own fp16 drift on this GPU (oracle = same torch ops as the reference modules)
    sd16 = {k: v.cuda().half() for k, v in sd.items()}
    ref16 = O.unet_forward(sd16, cfg, sample, torch.tensor(c["timestep"]), low, ctx, c["class_labels"])
    err_ref = _rel(ref16.cpu(), c["out"])
    print(f"\n[unet {case}] rel L2 err vs fp32 golden: uav_b200 {err:.3e} | reference-fp16 (torch) {err_ref:.3e}")
    assert err <= max(1.5 * err_ref, 5e-3), (err, err_ref)
    # second call (cached prompt K/V, packed weights) must be bit-identical
    out2 = m(sample, torch.tensor(c["timestep"]), low, encoder_hidden_states=ctx, class_labels=c["class_labels"].cuda()).sample
    assert torch.equal(out, out2)


def test_unet_shared_cfg_prefix(unet):
    """cfg_shared_input=True (prefix computed once for both CFG halves) == the plain batch-2 forward up to fp16
    reduction-order noise (GroupNorm partial sums are grouped differently for batch 1 and 2)"""
    m, sd, cfg = unet
    c = torch.load(os.path.join(G, "unet.pt"), weights_only=False)["t3_16x24"]
    sample = c["sample"][:1].repeat(2, 1, 1, 1, 1).cuda().half()
    low = c["low_res"][:1].repeat(2, 1, 1, 1, 1).cuda().half()
    ctx = c["ctx"].cuda().half()
    a = m(sample, 601, low, encoder_hidden_states=ctx, class_labels=torch.tensor([120])).sample
    b = m(sample, 601, low, encoder_hidden_states=ctx, class_labels=torch.tensor([120]), cfg_shared_input=True).sample
    err = _rel(b, a)
    print(f"\n[unet shared-prefix] rel L2 diff vs unshared {err:.3e}")
    # two fp16 executions whose GroupNorm statistics are summed in a different order (batch-1 prefix vs batch 2) decorrelate
    # their rounding errors: the difference of two runs that are each ~2.5e-3 from the fp32 result is ~sqrt(2) x that
    assert err < 5e-3
    assert not torch.equal(a[0], a[1])  # the two halves differ (different text rows)
