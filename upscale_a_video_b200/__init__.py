"""uav_b200 — B200-native (sm_100a) implementation of the Upscale-A-Video diffusion sampling path.

Public surface mirrors the reference (`/root/reference/models_video/`): `VideoUpscalePipeline`,
`UNetVideoModel`, `AutoencoderKLVideo`, `DDIMScheduler`, `Propagation`.  The arithmetic runs in
hand-written CUDA kernels behind a C ABI (`include/uav_b200.h`, `csrc/`); there is no CPU path.
"""
__version__ = "0.1.0"
