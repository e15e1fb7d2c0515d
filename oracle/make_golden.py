"""make_golden.py — mint golden vectors by running the UNMODIFIED reference (/root/reference) on CPU.

TEST INFRASTRUCTURE; runs only in the build container (needs /root/reference + oracle/shims):

    PYTHONPATH=oracle/shims:/root/reference:. python oracle/make_golden.py

Writes tests/golden/*.pt (inputs + reference outputs, fp32, small) and tests/golden/shapes_*.json (state-dict
key -> shape tables).  Weights are NOT stored: they are regenerated from (seed, key) by oracle/weights.py.
The reference publishes no golden vectors of its own, so these files are what pins the oracle
(`oracle/uav_oracle.py`, checked by tests/test_oracle_golden.py) and, through it, the CUDA path.
"""
import json
import os
import sys
import zlib

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle", "shims"), "/root/reference", ROOT]

import diffusers.utils as du  # shim  # noqa: E402
from diffusers.schedulers import DDPMScheduler  # shim  # noqa: E402
from models_video import AutoencoderKLVideo, Propagation, UNetVideoModel  # noqa: E402
from models_video.pipeline_upscale_a_video import VideoUpscalePipeline  # noqa: E402
import models_video.pipeline_upscale_a_video as ref_pipe_mod  # noqa: E402
from models_video.scheduling_ddim import DDIMScheduler  # noqa: E402

from oracle.weights import make_state_dict  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SEED_UNET, SEED_VAE = 1234, 4321
SCHED_CFGS = {
    "eps_linear_clip": dict(),  # defaults of scheduling_ddim.py:131-146
    "v_scaled_offset": dict(beta_schedule="scaled_linear", clip_sample=False, steps_offset=1,
                            prediction_type="v_prediction", set_alpha_to_one=False),
    "sample_linear": dict(prediction_type="sample", clip_sample=False),
}


def synth_flows(T, H, W, seed=1):
    """smooth 3-px field + noise; backward ~ -forward so the consistency mask is mixed 0/1."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")
    fw = torch.stack([3.0 * torch.sin(yy / 7.0 + 0.3) + 0 * xx, 3.0 * torch.cos(xx / 5.0) + 0 * yy])
    fw = fw[None, :, None].repeat(1, 1, T - 1, 1, 1) + 0.5 * torch.randn(1, 2, T - 1, H, W, generator=g)
    bw = -fw + 0.4 * torch.randn(1, 2, T - 1, H, W, generator=g)
    fw[..., :2, :] += 40.0  # a band that leaves the image (out-of-bounds sampling)
    return fw, bw


def build(cls, cfg_path, seed):
    cfg = json.load(open(cfg_path))
    m = cls.from_config(cfg)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(make_state_dict(shapes, seed), strict=True)
    return m.eval(), cfg, shapes


class FakeTokenizer:
    """deterministic stand-in for CLIPTokenizer (weights are not in the repo, README.md:78-101)"""
    model_max_length = 77

    def __call__(self, prompt, padding=None, max_length=None, truncation=None, return_tensors=None):
        prompts = [prompt] if isinstance(prompt, str) else list(prompt)
        L = max_length or self.model_max_length
        ids = torch.zeros(len(prompts), L, dtype=torch.long)
        for i, p in enumerate(prompts):
            g = torch.Generator().manual_seed(zlib.crc32(p.encode()))
            ids[i] = torch.randint(0, 1000, (L,), generator=g)
        return type("Enc", (), {"input_ids": ids})()

    def batch_decode(self, ids):
        return [""] * len(ids)


class FakeTextEncoder(torch.nn.Module):
    def __init__(self, seed=2):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.table = torch.nn.Parameter(torch.randn(1000, 1024, generator=g) * 0.3, requires_grad=False)
        self.config = type("Cfg", (), {})()

    @property
    def dtype(self):
        return self.table.dtype

    def forward(self, ids, attention_mask=None):
        return (self.table[ids],)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_grad_enabled(False)
    torch.set_num_threads(os.cpu_count())
    cfgdir = "/root/reference/configs"

    # ------------------------------------------------------------------ UNet
    unet, ucfg, ushapes = build(UNetVideoModel, f"{cfgdir}/unet_video_config.json", SEED_UNET)
    json.dump({k: list(v) for k, v in ushapes.items()}, open(f"{OUT}/shapes_unet.json", "w"))
    cases = {}
    for name, (B, T, H, W, t, cl) in {
        "t3_16x24": (2, 3, 16, 24, 601, [120]),
        "t2_20x28_upsize": (2, 2, 20, 28, 34, [80, 80]),
        "t8_8x8": (2, 8, 8, 8, 958, [150, 150]),
    }.items():
        g = torch.Generator().manual_seed(zlib.crc32(name.encode()))
        sample = torch.randn(B, 4, T, H, W, generator=g)
        low = torch.randn(B, 3, T, H, W, generator=g)
        ctx = torch.randn(B, 77, 1024, generator=g) * 0.3
        out = unet(sample, torch.tensor(t), low, encoder_hidden_states=ctx, class_labels=torch.tensor(cl)).sample
        cases[name] = dict(sample=sample, low_res=low, ctx=ctx, timestep=t, class_labels=torch.tensor(cl), out=out)
        print("unet", name, out.abs().mean().item())
    torch.save(cases, f"{OUT}/unet.pt")

    # ------------------------------------------------------------------ VAE
    vcases = {}
    vae3d, v3cfg, v3shapes = build(AutoencoderKLVideo, f"{cfgdir}/vae_3d_config.json", SEED_VAE)
    json.dump({k: list(v) for k, v in v3shapes.items()}, open(f"{OUT}/shapes_vae_3d.json", "w"))
    g = torch.Generator().manual_seed(7)
    z = torch.randn(1, 4, 3, 16, 24, generator=g)
    img = torch.rand(1, 3, 3, 16, 24, generator=g) * 2 - 1
    vcases["vae3d_decode"] = dict(z=z, img=img, w_lr=1.0, out=vae3d.decode(z, img, 1.0).sample)
    x = torch.rand(1, 3, 2, 32, 48, generator=g) * 2 - 1
    vcases["vae3d_encode"] = dict(x=x, moments=vae3d.encode(x).latent_dist.parameters)
    vaev, vvcfg, vvshapes = build(AutoencoderKLVideo, f"{cfgdir}/vae_video_config.json", SEED_VAE)
    json.dump({k: list(v) for k, v in vvshapes.items()}, open(f"{OUT}/shapes_vae_video.json", "w"))
    z = torch.randn(1, 4, 2, 12, 16, generator=g)
    img = torch.rand(1, 3, 2, 12, 16, generator=g) * 2 - 1
    vcases["vaevideo_decode"] = dict(z=z, img=img, w_lr=0.7, out=vaev.decode(z, img, 0.7).sample)
    torch.save(vcases, f"{OUT}/vae.pt")
    print("vae done")

    # ------------------------------------------------------------------ scheduler
    scases = {}
    g = torch.Generator().manual_seed(11)
    x = torch.randn(1, 4, 3, 8, 8, generator=g)
    mo = torch.randn(1, 4, 3, 8, 8, generator=g)
    for name, kw in SCHED_CFGS.items():
        for dt in (torch.float32, torch.float16):
            s = DDIMScheduler(**kw)
            for steps in (30, 2):
                s.set_timesteps(steps)
                rec = []
                for i in (0, len(s.timesteps) // 2, len(s.timesteps) - 1):
                    t = s.timesteps[i]
                    x0 = s.step_v0(mo.to(dt), t, x.to(dt)).pred_original_sample
                    xp = s.step_vt(x0, mo.to(dt), t, x.to(dt)).prev_sample
                    rec.append(dict(i=i, t=int(t), x0=x0, prev=xp))
                scases[f"{name}/{str(dt)[6:]}/{steps}"] = dict(timesteps=s.timesteps.clone(), steps=rec)
            nz = s.add_noise(x.to(dt), mo.to(dt), torch.tensor([120]))
            scases[f"{name}/{str(dt)[6:]}/add_noise"] = nz
    scases["inputs"] = dict(x=x, model_output=mo)
    torch.save(scases, f"{OUT}/scheduler.pt")
    print("scheduler done")

    # ------------------------------------------------------------------ propagation
    pcases = {}
    prop = Propagation(4, learnable=False)
    g = torch.Generator().manual_seed(13)
    T, H, W = 5, 24, 32
    x = torch.randn(1, 4, T, H, W, generator=g)
    fw, bw = synth_flows(T, H, W)
    for dt in (torch.float32, torch.float16):
        for (interp, mode, a1, a2) in (("nearest", "fuse", 0.001, 0.05), ("bilinear", "copy", 0.01, 0.5)):
            try:
                out = prop(x.to(dt), fw.to(dt), bw.to(dt), interpolation=interp, mode=mode, fuse_scale=0.5,
                           alpha1=a1, alpha2=a2)
            except RuntimeError as e:  # CPU half grid_sample may be unsupported
                print("propagation", dt, interp, "skipped:", e)
                continue
            pcases[f"{str(dt)[6:]}/{interp}_{mode}"] = out
    pcases["inputs"] = dict(x=x, flows_forward=fw, flows_backward=bw)
    torch.save(pcases, f"{OUT}/propagation.pt")
    print("propagation done", list(pcases))

    # ------------------------------------------------------------------ pipeline
    draws = []
    orig_randn = ref_pipe_mod.randn_tensor

    def rec_randn(shape, generator=None, device=None, dtype=None, layout=None):
        t = orig_randn(shape, generator=generator, device=device, dtype=dtype)
        draws.append(t.clone())
        return t

    ref_pipe_mod.randn_tensor = rec_randn
    plcases = {}
    for name, (vae, T, H, W, steps, psteps, w_lr) in {
        "c1_t1_64x64": (vae3d, 1, 64, 64, 2, [], 1.0),
        "t11_16x16_prop": (vaev, 11, 16, 16, 3, [1], 0.8),
    }.items():
        sched = DDIMScheduler(**SCHED_CFGS["v_scaled_offset"])
        low = DDPMScheduler(beta_schedule="scaled_linear")
        te, tok = FakeTextEncoder(), FakeTokenizer()
        pipe = VideoUpscalePipeline(text_encoder=te, tokenizer=tok, low_res_scheduler=low, scheduler=sched, vae=vae,
                                    unet=unet, propagator=Propagation(4, learnable=False))
        g = torch.Generator().manual_seed(zlib.crc32(name.encode()))
        image = torch.rand(1, 3, T, H, W, generator=g) * 2 - 1
        flows = list(synth_flows(T, H, W)) if T > 1 else None
        draws.clear()
        gen = torch.Generator().manual_seed(10)  # inference_upscale_a_video.py:197
        out, lat = pipe("a cat", image=image, flows_bi=flows, num_inference_steps=steps, guidance_scale=6.0,
                        noise_level=120, negative_prompt="blur", generator=gen, propagation_steps=psteps,
                        w_lr=w_lr, return_dict=False)
        pe = torch.cat([te(tok("blur", max_length=77).input_ids)[0], te(tok("a cat").input_ids)[0]])
        plcases[name] = dict(image=image, flows=flows, steps=steps, propagation_steps=psteps, w_lr=w_lr,
                             guidance_scale=6.0, noise_level=120, prompt_embeds=pe, noise=draws[0], latents=draws[1],
                             out=out, latents_out=lat, vae="vae_3d" if vae is vae3d else "vae_video")
        print("pipeline", name, out.shape, out.abs().mean().item())
    torch.save(plcases, f"{OUT}/pipeline.pt")
    json.dump(dict(seed_unet=SEED_UNET, seed_vae=SEED_VAE, sched_cfgs=SCHED_CFGS,
                   reference_commit="10ca1d75", torch=torch.__version__), open(f"{OUT}/meta.json", "w"), indent=1)


if __name__ == "__main__":
    main()
