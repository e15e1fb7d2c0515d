"""CPU check of the RAFT host logic (weight packing: BatchNorm folding, space-to-depth stride-2 filters, fused z|r gates,
padded channels; buffer plumbing of the GRU loop; batched bidirectional call) with every `ops.*` kernel wrapper replaced by a
plain-torch stand-in that mimics the kernels' contracts (fp32 math, fp16 outputs).  The kernels themselves are checked on the
GPU (tests/test_raft_gpu.py); this test pins everything around them against the reference fixtures without a GPU."""
import json
import os

import pytest
import torch
import torch.nn.functional as F

G = os.path.join(os.path.dirname(__file__), "golden")


def _act(x, act):
    return {0: x, 3: F.relu(x), 4: torch.sigmoid(x), 5: torch.tanh(x)}[act]


class EmuOps:
    ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH = 0, 3, 4, 5

    @staticmethod
    def conv2d_taps(x, w, bias=None, *, pad_top, pad_left, out=None, residual=None, act=0, out_dtype=torch.float16):
        cout, kh, kw, cin = w.shape
        xp = F.pad(x.float().permute(0, 3, 1, 2), (pad_left, kw - 1 - pad_left, pad_top, kh - 1 - pad_top))
        y = _act(F.conv2d(xp, w.float().permute(0, 3, 1, 2), bias), act).permute(0, 2, 3, 1)
        if out is None:
            return y.to(out_dtype).contiguous()
        out.copy_(y)
        return out

    @staticmethod
    def linear(a, w, bias=None, *, out=None, act=0, out_dtype=torch.float16, **_):
        y = _act(F.linear(a.float(), w.float(), bias), act)
        if out is None:
            return y.to(out_dtype)
        out.copy_(y)
        return out

    @staticmethod
    def instnorm_relu(x, relu=True, eps=1e-5):
        y = F.instance_norm(x.float().permute(0, 3, 1, 2), eps=eps).permute(0, 2, 3, 1)
        return (F.relu(y) if relu else y).half().contiguous()

    @staticmethod
    def add_relu(a, b):
        return F.relu(a.float() + b.float()).half()

    @staticmethod
    def raft_split_tanh_relu(cnet, net, inp_a, inp_b):
        c = cnet.shape[1] // 2
        net.copy_(torch.tanh(cnet[:, :c].float()))
        inp_a.copy_(F.relu(cnet[:, c:].float()))
        if inp_b is not None:
            inp_b.copy_(F.relu(cnet[:, c:].float()))

    @staticmethod
    def avgpool2x2_f32(x):
        return F.avg_pool2d(x[:, None], 2, stride=2)[:, 0].contiguous()

    @staticmethod
    def raft_corr_lookup(levels, coords, out):
        from oracle import raft_oracle as R
        pix = coords.shape[0]
        feat = R.corr_lookup([lv[:, None] for lv in levels], coords.view(pix, 1, 1, 2).permute(0, 3, 1, 2))  # (pix, 324, 1, 1)
        out[:, :324] = feat.view(pix, 324)
        out[:, 324:] = 0
        return out

    @staticmethod
    def raft_gru_rh(zr, h, out):
        c = h.shape[1]
        out.copy_(zr[:, c:2 * c].float() * h.float())

    @staticmethod
    def raft_gru_update(zr, q, h):
        c = h.shape[1]
        z = zr[:, :c].float()
        h.copy_((1 - z) * h.float() + z * q.float())

    @staticmethod
    def raft_flow_update(coords1, delta, h8, w8, flow16=None, dst_a=None, dst_b=None):
        if delta is not None:
            coords1 += delta[:, :2]
        rows = coords1.shape[0]
        r = torch.arange(rows)
        c0 = torch.stack([(r % w8).float(), ((r // w8) % h8).float()], dim=1)
        fl = (coords1 - c0).half()
        for d in (flow16, dst_a, dst_b):
            if d is not None:
                d[:, :2] = fl

    @staticmethod
    def raft_convex_upsample(coords1, mask, nimg, h8, w8):
        from oracle import raft_oracle as R
        flow = (coords1.view(nimg, h8, w8, 2).permute(0, 3, 1, 2) - R.coords_grid(nimg, h8, w8))
        return R.upsample_flow(flow, mask.float().view(nimg, h8, w8, 576).permute(0, 3, 1, 2))

    @staticmethod
    def copy_channels(src, dst):
        dst.copy_(src)


@pytest.fixture()
def emu(monkeypatch):
    from upscale_a_video_b200 import raft as raft_mod
    monkeypatch.setattr(raft_mod, "ops", EmuOps)
    return raft_mod


def _epe(a, b):
    e = torch.sqrt(((a.float() - b.float()) ** 2).sum(dim=1))
    return e.mean().item(), e.max().item()


def test_raft_host_logic_against_reference_fixtures(emu):
    from oracle import raft_oracle as R
    from oracle.weights import make_state_dict
    g = torch.load(os.path.join(G, "raft.pt"), map_location="cpu", weights_only=False)
    shapes = json.load(open(os.path.join(G, "shapes_raft.json")))
    m = emu.RAFT()
    m.load_state_dict(make_state_dict(shapes, g["seed"]), strict=True)
    m.eval()
    c = g["cases"]["raft_128x136_it3"]
    clip = R.synth_clip(*c["clip"])
    with torch.no_grad():
        lo, up = m._forward_impl(clip[0, :, 0][None], clip[0, :, 1][None], c["iters"], None)
    mean_lo, max_lo = _epe(lo, c["flow_lo"])
    mean_up, max_up = _epe(up, c["flow_up"])
    print(f"[raft host, fp16-emulated kernels] EPE lo mean {mean_lo:.4f} max {max_lo:.4f}; up mean {mean_up:.4f} max {max_up:.4f} px")
    assert mean_up < 0.05 and max_up < 0.5 and mean_lo < 0.01 and max_lo < 0.1

    c = g["cases"]["raft_320x576_it4"]  # BASELINE frame size
    clip = R.synth_clip(*c["clip"])
    with torch.no_grad():
        lo, up = m._forward_impl(clip[0, :, 0][None], clip[0, :, 1][None], c["iters"], None)
    st = c["stride"]
    mean_up, max_up = _epe(up[..., ::st, ::st], c["flow_up"])
    print(f"[raft host 320x576] EPE up mean {mean_up:.4f} max {max_up:.4f} px")
    assert mean_up < 0.05 and max_up < 0.5

    # RAFT_bi: both directions in one batched call, non-multiple-of-8 size (resize in, row-quirk resize out)
    bi = emu.RAFT_bi(model_path=None, device="cpu")
    bi.fix_raft = m
    m.forward = lambda a, b, iters=12, flow_init=None, test_mode=True: m._forward_impl(a, b, iters, flow_init)
    c = g["cases"]["bi_124x132_it2"]
    with torch.no_grad():
        f, b = bi.forward(R.synth_clip(*c["clip"]), iters=c["iters"])
    st = c["stride"]
    for got, ref in ((f, c["fwd"]), (b, c["bwd"])):
        gg = got[..., ::st, ::st].permute(0, 2, 1, 3, 4).reshape(-1, 2, *ref.shape[-2:])
        rr = ref.permute(0, 2, 1, 3, 4).reshape(-1, 2, *ref.shape[-2:])
        mean_e, max_e = _epe(gg, rr)
        assert mean_e < 0.05 and max_e < 0.5, (mean_e, max_e)
