"""CPU: pin the oracle (oracle/uav_oracle.py) to the golden vectors minted from the UNMODIFIED reference
(oracle/make_golden.py).  fp32 both sides, so the tolerance is accumulation-order noise only."""
import json
import os

import pytest
import torch

from oracle import uav_oracle as O
from oracle.weights import make_state_dict

G = os.path.join(os.path.dirname(__file__), "golden")
META = json.load(open(os.path.join(G, "meta.json")))
CFG = os.path.join(os.path.dirname(__file__), "..", "upscale_a_video_b200", "configs")


def _load(name):
    return torch.load(os.path.join(G, name), map_location="cpu", weights_only=False)


def _cfg(name):
    return json.load(open(os.path.join(CFG, name)))


def _sd(kind, dtype=torch.float32):
    shapes = json.load(open(os.path.join(G, f"shapes_{kind}.json")))
    seed = META["seed_unet"] if kind == "unet" else META["seed_vae"]
    return make_state_dict(shapes, seed, dtype)


@pytest.fixture(scope="module")
def unet_sd():
    return _sd("unet")


@pytest.mark.parametrize("case", ["t3_16x24", "t2_20x28_upsize", "t8_8x8"])
def test_unet_forward(unet_sd, case):
    c = _load("unet.pt")[case]
    with torch.no_grad():
        out = O.unet_forward(unet_sd, _cfg("unet_video_config.json"), c["sample"], torch.tensor(c["timestep"]),
                             c["low_res"], c["ctx"], c["class_labels"])
    torch.testing.assert_close(out, c["out"], rtol=1e-4, atol=1e-4)


def test_vae():
    v = _load("vae.pt")
    sd3, cfg3 = _sd("vae_3d"), _cfg("vae_3d_config.json")
    with torch.no_grad():
        c = v["vae3d_decode"]
        torch.testing.assert_close(O.vae_decode(sd3, cfg3, c["z"], c["img"], c["w_lr"]), c["out"], rtol=1e-4, atol=1e-4)
        c = v["vae3d_encode"]
        torch.testing.assert_close(O.vae_encode_moments(sd3, cfg3, c["x"]), c["moments"], rtol=1e-4, atol=1e-4)
        sdv, cfgv = _sd("vae_video"), _cfg("vae_video_config.json")
        c = v["vaevideo_decode"]
        torch.testing.assert_close(O.vae_decode(sdv, cfgv, c["z"], c["img"], c["w_lr"]), c["out"], rtol=1e-4, atol=1e-4)


def test_scheduler():
    s = _load("scheduler.pt")
    x, mo = s["inputs"]["x"], s["inputs"]["model_output"]
    for name, kw in META["sched_cfgs"].items():
        for dt in (torch.float32, torch.float16):
            d = O.DDIM(**kw)
            for steps in (30, 2):
                rec = s[f"{name}/{str(dt)[6:]}/{steps}"]
                d.set_timesteps(steps)
                assert torch.equal(d.timesteps, rec["timesteps"])
                for r in rec["steps"]:
                    t = d.timesteps[r["i"]]
                    x0 = d.step_v0(mo.to(dt), t, x.to(dt))
                    assert torch.equal(x0, r["x0"]), (name, dt, steps, r["i"])
                    assert torch.equal(d.step_vt(x0, mo.to(dt), t, x.to(dt)), r["prev"])
            assert torch.equal(d.add_noise(x.to(dt), mo.to(dt), torch.tensor([120])), s[f"{name}/{str(dt)[6:]}/add_noise"])


def test_propagation():
    p = _load("propagation.pt")
    i = p["inputs"]
    for dt in (torch.float32, torch.float16):
        for interp, mode, a1, a2 in (("nearest", "fuse", 0.001, 0.05), ("bilinear", "copy", 0.01, 0.5)):
            out = O.propagation(i["x"].to(dt), i["flows_forward"].to(dt), i["flows_backward"].to(dt), interp, mode, 0.5, a1, a2)
            assert torch.equal(out, p[f"{str(dt)[6:]}/{interp}_{mode}"]), (dt, interp)


def test_unet_windows():
    # window grid incl. the re-anchored / duplicated last window (pipeline_upscale_a_video.py:621-625)
    assert O.unet_windows(11) == [(0, 8), (3, 11)]
    assert O.unet_windows(14) == [(0, 8), (6, 14), (6, 14)]
    assert O.unet_windows(16) == [(0, 8), (6, 14), (8, 16)]
    assert O.unet_windows(32) == [(0, 8), (6, 14), (12, 20), (18, 26), (24, 32), (24, 32)]


@pytest.mark.parametrize("case", ["c1_t1_64x64", "t11_16x16_prop"])
def test_pipeline(unet_sd, case):
    c = _load("pipeline.pt")[case]
    vae_kind = c["vae"]
    vsd, vcfg = _sd(vae_kind), _cfg(f"{vae_kind}_config.json")
    sched = O.DDIM(**META["sched_cfgs"]["v_scaled_offset"])
    low = O.DDIM(beta_schedule="scaled_linear")
    with torch.no_grad():
        out, lat = O.pipeline_call(unet_sd, _cfg("unet_video_config.json"), vsd, vcfg, sched, low, image=c["image"],
                                   prompt_embeds=c["prompt_embeds"], noise=c["noise"], latents=c["latents"],
                                   flows_bi=c["flows"], num_inference_steps=c["steps"], guidance_scale=c["guidance_scale"],
                                   noise_level=c["noise_level"], propagation_steps=c["propagation_steps"], w_lr=c["w_lr"],
                                   return_latents=True)
    torch.testing.assert_close(lat, c["latents_out"], rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(out, c["out"], rtol=1e-3, atol=2e-3)


# ------------------------------------------------------------------------------------------------
# colour fix + packing: oracle/color_oracle.py vs the reference's own functions
# ------------------------------------------------------------------------------------------------
def test_color_oracle_against_reference_fixtures():
    from oracle import color_oracle as co

    cases = _load("color.pt")
    assert set(cases) == {"t2_16x24", "t1_13x19"}
    for name, c in cases.items():
        lr, hr = c["lr"], c["hr"]
        up = co.bicubic_upsample(lr, 4)
        assert (up - c["bicubic"]).abs().max().item() < 2e-6, name
        mean, std = co.calc_mean_std(hr)
        assert (mean - c["mean"]).abs().max().item() < 1e-6 and (std - c["std"]).abs().max().item() < 1e-6, name
        # tolerance in the test: fp32 reductions / 9-tap sums in a different association order
        assert (co.adaptive_instance_normalization(hr, c["bicubic"]) - c["adain"]).abs().max().item() < 5e-6, name
        assert (co.wavelet_reconstruction(hr, c["bicubic"]) - c["wavelet"]).abs().max().item() < 2e-6, name
        if "high" in c:
            high, low = co.wavelet_decomposition(hr)
            assert (high - c["high"]).abs().max().item() < 2e-6 and (low - c["low"]).abs().max().item() < 1e-6, name
            assert (co.wavelet_blur(hr, 4) - c["blur4"]).abs().max().item() < 1e-6, name
        # integer work: bit-exact
        assert torch.equal(co.pack_video_uint8(hr), c["pack_hr"]), name
        assert torch.equal(co.pack_video_uint8(c["adain"]), c["pack_adain"]), name
        # the CLI block end to end (inference_upscale_a_video.py:323-333)
        out = co.color_fix_frames(hr.permute(1, 0, 2, 3)[None], lr.permute(1, 0, 2, 3)[None], "Wavelet")
        assert (out - c["wavelet"]).abs().max().item() < 5e-6, name


# ------------------------------------------------------------------------------------------------
# RAFT bidirectional flow: oracle/raft_oracle.py vs the reference's own RAFT / RAFT_bi
# ------------------------------------------------------------------------------------------------
def test_raft_oracle_against_reference_fixtures():
    from oracle import raft_oracle as R

    g = _load("raft.pt")
    shapes = json.load(open(os.path.join(G, "shapes_raft.json")))
    assert len(shapes) == 179
    sd = make_state_dict(shapes, g["seed"])
    cases = g["cases"]
    with torch.no_grad():
        c = cases["raft_128x136_it3"]
        clip = R.synth_clip(*c["clip"])
        lo, up = R.raft_forward(sd, clip[0, :, 0][None], clip[0, :, 1][None], c["iters"])
        # fp32 with the same ATen ops in the same order: only thread-count dependent reduction order differs
        assert (lo - c["flow_lo"]).abs().max().item() < 1e-4 and (up - c["flow_up"]).abs().max().item() < 1e-3
        assert c["flow_up"].abs().max().item() > 0.5  # the fixture is not a trivial zero flow
        c = cases["bi_124x132_it2"]  # H, W not multiples of 8: trilinear resize in, resize_flow_pytorch (row quirk) out
        f, b = R.raft_bi_forward(sd, R.synth_clip(*c["clip"]), c["iters"])
        assert f.shape == (1, 2, 2, 124, 132)
        st = c["stride"]
        assert (f[..., ::st, ::st] - c["fwd"]).abs().max().item() < 1e-3 and (b[..., ::st, ::st] - c["bwd"]).abs().max().item() < 1e-3
        c = cases["slicing_13f_128x128_it1"]  # 13 frames > one 12-frame short clip
        f, b = R.raft_bi_forward_slicing(sd, R.synth_clip(*c["clip"]), c["iters"])
        assert f.shape == (1, 2, 12, 128, 128)
        st = c["stride"]
        assert (f[..., ::st, ::st] - c["fwd"]).abs().max().item() < 1e-3 and (b[..., ::st, ::st] - c["bwd"]).abs().max().item() < 1e-3
        c = cases["raft_320x576_it4"]  # BASELINE frame size
        clip = R.synth_clip(*c["clip"])
        lo, up = R.raft_forward(sd, clip[0, :, 0][None], clip[0, :, 1][None], c["iters"])
        st = c["stride"]
        assert (lo - c["flow_lo"]).abs().max().item() < 1e-3 and (up[..., ::st, ::st] - c["flow_up"]).abs().max().item() < 1e-2
    assert [R.short_clip_len(w) for w in (576, 640, 641, 720, 960, 1280, 1281)] == [12, 12, 8, 8, 4, 4, 2]


# ------------------------------------------------------------------------------------------------
# CLIP text encoder: oracle/clip_oracle.py vs transformers' CLIPTextModel
# ------------------------------------------------------------------------------------------------
def test_clip_oracle_against_transformers_fixtures():
    from oracle import clip_oracle as Co

    g = _load("clip.pt")
    for name, c in g["cases"].items():
        sd = make_state_dict(c["shapes"], g["seed"])
        with torch.no_grad():
            out = Co.clip_text_forward(sd, c["config"], c["input_ids"])
        assert (out[..., ::c["col_stride"]] - c["last_hidden_state"]).abs().max().item() < 2e-5, name
