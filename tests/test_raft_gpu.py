"""GPU parity of the RAFT bidirectional flow (upscale_a_video_b200/raft.py, csrc/raft.cu, uav_conv2d_taps) against the
reference fixtures (tests/golden/raft.pt, minted from the unmodified reference RAFT / RAFT_bi by oracle/make_golden_raft.py)
and against the fp32 CPU oracle on other shapes.  The reference computes in fp32; the product uses fp16 activations with fp32
accumulation and fp32 correlation / coordinates (upstream RAFT's own mixed-precision split), so flows are compared by
end-point error: mean EPE <= 0.05 px and max EPE <= 0.5 px on flows of 1-10 px magnitude; the non-GEMM kernels are also
checked one by one against their torch formulas at tight tolerances."""
import json
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(autouse=True)
def _setup(uav_lib):
    torch.manual_seed(0)


def _model():
    from oracle.weights import make_state_dict
    from upscale_a_video_b200.raft import RAFT
    g = torch.load(os.path.join(G, "raft.pt"), map_location="cpu", weights_only=False)
    shapes = json.load(open(os.path.join(G, "shapes_raft.json")))
    sd = make_state_dict(shapes, g["seed"])
    m = RAFT()
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval(), sd, g["cases"]


def _epe(a, b):
    d = (a.float().cpu() - b.float().cpu())
    e = torch.sqrt((d ** 2).sum(dim=1 if a.dim() == 4 else 1))
    return e.mean().item(), e.max().item()


def test_raft_forward_against_reference_fixture():
    from oracle import raft_oracle as R
    m, sd, cases = _model()
    c = cases["raft_128x136_it3"]
    clip = R.synth_clip(*c["clip"])
    lo, up = m(clip[0, :, 0][None].cuda(), clip[0, :, 1][None].cuda(), iters=c["iters"], test_mode=True)
    mean_lo, max_lo = _epe(lo, c["flow_lo"])
    mean_up, max_up = _epe(up, c["flow_up"])
    print(f"[raft] 128x136 it3: EPE lo mean {mean_lo:.4f} max {max_lo:.4f} (1/8 px units); up mean {mean_up:.4f} max {max_up:.4f} px")
    assert mean_lo < 0.01 and max_lo < 0.1
    assert mean_up < 0.05 and max_up < 0.5


def test_raft_forward_baseline_size():
    """one pair at the BASELINE.json frame size: 40x72 grid, 2880 x 2880 correlation volume"""
    from oracle import raft_oracle as R
    m, sd, cases = _model()
    c = cases["raft_320x576_it4"]
    clip = R.synth_clip(*c["clip"])
    lo, up = m(clip[0, :, 0][None].cuda(), clip[0, :, 1][None].cuda(), iters=c["iters"], test_mode=True)
    st = c["stride"]
    mean_lo, max_lo = _epe(lo, c["flow_lo"])
    mean_up, max_up = _epe(up[..., ::st, ::st], c["flow_up"])
    print(f"[raft] 320x576 it4: EPE lo mean {mean_lo:.4f} max {max_lo:.4f}; up mean {mean_up:.4f} max {max_up:.4f} px "
          f"(flow magnitude mean {c['flow_up'].abs().mean().item():.1f} px)")
    assert mean_up < 0.05 and max_up < 0.5


def test_raft_bi_against_reference_fixtures():
    from oracle import raft_oracle as R
    from upscale_a_video_b200.raft import RAFT_bi
    m, sd, cases = _model()
    bi = RAFT_bi(model_path=None)
    bi.fix_raft = m
    for name in ("bi_124x132_it2", "slicing_13f_128x128_it1"):
        c = cases[name]
        frames = R.synth_clip(*c["clip"]).cuda()
        f, b = bi.forward_slicing(frames, iters=c["iters"])
        st = c["stride"]
        for got, ref, tag in ((f, c["fwd"], "fwd"), (b, c["bwd"], "bwd")):
            g = got[..., ::st, ::st].permute(0, 2, 1, 3, 4).reshape(-1, 2, *got[..., ::st, ::st].shape[-2:])
            r = ref.permute(0, 2, 1, 3, 4).reshape(-1, 2, *ref.shape[-2:])
            mean_e, max_e = _epe(g, r)
            print(f"[raft_bi] {name} {tag}: EPE mean {mean_e:.4f} max {max_e:.4f} px")
            assert mean_e < 0.05 and max_e < 0.5, (name, tag)


def test_raft_kernels_against_torch():
    from upscale_a_video_b200 import ops
    dev = "cuda"
    # InstanceNorm + ReLU (C = 64 / 96 / 128, odd pixel counts)
    for n, h, w, c in [(3, 17, 23, 64), (2, 20, 36, 96), (1, 8, 9, 128)]:
        x = (torch.randn(n, h, w, c, device=dev) * 1.7 + 0.3).half()
        ref = F.relu(F.instance_norm(x.float().permute(0, 3, 1, 2), eps=1e-5)).permute(0, 2, 3, 1)
        assert (ops.instnorm_relu(x, True).float() - ref).abs().max().item() < 4e-3
        ref2 = F.instance_norm(x.float().permute(0, 3, 1, 2), eps=1e-5).permute(0, 2, 3, 1)
        assert (ops.instnorm_relu(x, False).float() - ref2).abs().max().item() < 4e-3
    a, b = torch.randn(5, 7, 9, 64, device=dev).half(), torch.randn(5, 7, 9, 64, device=dev).half()
    assert torch.equal(ops.add_relu(a, b), F.relu(a.float() + b.float()).half())
    # pyramid + lookup vs the oracle's formulation on the GPU
    from oracle import raft_oracle as R
    nimg, h8, w8 = 2, 16, 18
    corr = torch.randn(nimg * h8 * w8, h8, w8, device=dev)
    pyr = [corr]
    for _ in range(3):
        pyr.append(ops.avgpool2x2_f32(pyr[-1]))
    ref_pyr = [corr[:, None]]
    for _ in range(3):
        ref_pyr.append(F.avg_pool2d(ref_pyr[-1], 2, stride=2))
    for got, ref in zip(pyr, ref_pyr):
        assert (got - ref[:, 0]).abs().max().item() < 1e-6
    coords = R.coords_grid(nimg, h8, w8).to(dev) + torch.randn(nimg, 2, h8, w8, device=dev) * 3.0  # incl. out of range
    ref_feat = R.corr_lookup([p.cpu() for p in ref_pyr], coords.cpu())  # (n, 324, h8, w8)
    out = torch.full((nimg * h8 * w8, 328), 7.0, dtype=torch.float16, device=dev)
    ops.raft_corr_lookup(pyr, coords.permute(0, 2, 3, 1).reshape(-1, 2).contiguous(), out)
    got = out[:, :324].float().cpu().view(nimg, h8, w8, 324).permute(0, 3, 1, 2)
    assert (got - ref_feat).abs().max().item() < 5e-3  # fp16 output of O(1) values
    assert (out[:, 324:] == 0).all()
    # convex upsampling
    coords1 = (R.coords_grid(nimg, h8, w8) + torch.randn(nimg, 2, h8, w8)).to(dev)
    mask = torch.randn(nimg, 576, h8, w8, device=dev).half()
    ref_up = R.upsample_flow((coords1.cpu() - R.coords_grid(nimg, h8, w8)), mask.float().cpu())
    got_up = ops.raft_convex_upsample(coords1.permute(0, 2, 3, 1).reshape(-1, 2).contiguous(),
                                      mask.permute(0, 2, 3, 1).reshape(-1, 576).contiguous(), nimg, h8, w8)
    assert (got_up.cpu() - ref_up).abs().max().item() < 1e-4


@pytest.mark.parametrize("kh,kw,pt,pl,cin,cout,act", [(7, 7, 3, 3, 8, 128, 3), (1, 5, 0, 2, 384, 256, 4), (5, 1, 2, 0, 384, 128, 5),
                                                      (4, 4, 2, 2, 32, 64, 0), (2, 2, 1, 1, 256, 96, 3), (3, 3, 1, 1, 256, 2, 0)])
def test_conv2d_taps(kh, kw, pt, pl, cin, cout, act):
    from upscale_a_video_b200 import ops
    x = torch.randn(3, 20, 28, cin, device="cuda").half()
    w = (torch.randn(cout, kh, kw, cin, device="cuda") / (kh * kw * cin) ** 0.5).half()
    b = torch.randn(cout, device="cuda") * 0.1
    out_dtype = torch.float32 if cout == 2 else torch.float16
    y = ops.conv2d_taps(x, w, b, pad_top=pt, pad_left=pl, act=act, out_dtype=out_dtype)
    xp = F.pad(x.float().permute(0, 3, 1, 2), (pl, kw - 1 - pl, pt, kh - 1 - pt))
    ref = F.conv2d(xp, w.float().permute(0, 3, 1, 2), b)
    ref = {0: ref, 3: F.relu(ref), 4: torch.sigmoid(ref), 5: torch.tanh(ref)}[act].permute(0, 2, 3, 1)
    assert (y.float() - ref).abs().max().item() < 6e-3
