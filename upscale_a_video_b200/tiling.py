"""Spatial tile driver — the tile loop of /root/reference/inference_upscale_a_video.py:200-304 as a
plan + executor, so that large frames can be dealt to GPUs tile by tile.

`plan_tiles` reproduces the reference geometry exactly (tiles of `tile_size` plus `overlap` LR pixels of context on every
side, the last row / column merged into its neighbour when the remainder is <= overlap, hard paste of the central region,
no blending); it is pinned by `tests/golden/tiles.json`, which is produced by EXECUTING the reference's own loop
(`oracle/make_golden_tiles.py`).  `upscale_tiled` runs the pipeline per tile.  With torch.distributed initialised, tiles are
dealt round-robin to ranks (each tile is an independent pipeline run: no per-step collective at all) and the pasted outputs
are combined with one all_reduce at the end.  The reference consumes ONE generator sequentially over the tiles
(inference...:197,268); to keep that stream every rank draws the noise of every tile in loop order and uses its own.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch

from . import sharding
from .pipeline_upscale_a_video import randn_tensor


@dataclass(frozen=True)
class Tile:
    in_box: Tuple[int, int, int, int]    # (y0, y1, x0, x1) on the LR frame, including the overlap context
    out_box: Tuple[int, int, int, int]   # (y0, y1, x0, x1) on the 4x output frame
    src_box: Tuple[int, int, int, int]   # (y0, y1, x0, x1) inside the tile's 4x output


def needs_tiling(h: int, w: int) -> bool:
    """inference_upscale_a_video.py:201-202"""
    return h * w >= 384 * 384


def plan_tiles(h: int, w: int, tile_size: int = 256, overlap: int = 64, scale: int = 4) -> List[Tile]:
    tiles_x, tiles_y = math.ceil(w / tile_size), math.ceil(h / tile_size)
    # a trailing tile whose fresh area would not exceed the overlap is merged into its neighbour (inference...:220-227)
    merge_w = (tiles_x - 1) * tile_size + overlap >= w
    merge_h = (tiles_y - 1) * tile_size + overlap >= h
    if merge_w:
        tiles_x -= 1
    if merge_h:
        tiles_y -= 1
    out_h, out_w = h * scale, w * scale
    plan = []
    for y in range(tiles_y):
        for x in range(tiles_x):
            x0, y0 = x * tile_size, y * tile_size
            x1, y1 = min(x0 + tile_size, w), min(y0 + tile_size, h)
            px0, px1 = max(x0 - overlap, 0), min(x1 + overlap, w)
            py0, py1 = max(y0 - overlap, 0), min(y1 + overlap, h)
            last_x, last_y = (x == tiles_x - 1 and merge_w), (y == tiles_y - 1 and merge_h)
            ox0, oy0 = x0 * scale, y0 * scale
            ox1 = out_w if last_x else x1 * scale
            oy1 = out_h if last_y else y1 * scale
            sx0, sy0 = (x0 - px0) * scale, (y0 - py0) * scale
            plan.append(Tile((py0, py1, px0, px1), (oy0, oy1, ox0, ox1), (sy0, sy0 + oy1 - oy0, sx0, sx0 + ox1 - ox0)))
    return plan


_SOLO_GROUPS = {}


def _solo_group(rank: int, world: int):
    """single-rank process groups, created once per world size (new_group is a collective and a communicator that is
    never destroyed would leak one NCCL communicator per rank per clip)"""
    if world not in _SOLO_GROUPS:
        import torch.distributed as dist
        _SOLO_GROUPS[world] = [dist.new_group([r]) for r in range(world)]  # every rank creates every group
    return _SOLO_GROUPS[world][rank]


@torch.no_grad()
def upscale_tiled(pipeline, image: torch.Tensor, flows_bi: Optional[list] = None, generator=None, tile_size: int = 256,
                  overlap: int = 64, process_group=None, **pipe_kwargs) -> torch.Tensor:
    """image: (1, 3, T, H, W) LR clip in [-1, 1] on the GPU.  Returns the (1, 3, T, 4H, 4W) output like the reference's
    tile branch.  `pipe_kwargs` go to `VideoUpscalePipeline.__call__` (prompt / prompt_embeds, steps, guidance, ...)."""
    b, c, t, h, w = image.shape
    plan = plan_tiles(h, w, tile_size, overlap)
    rank, world = sharding.world_info(process_group)
    out = image.new_zeros((b, c, t, 4 * h, 4 * w), dtype=torch.float32)
    dtype = None
    for key in ("prompt_embeds", "negative_prompt_embeds"):
        if pipe_kwargs.get(key) is not None:
            dtype = pipe_kwargs[key].dtype
    if dtype is None:
        dtype = getattr(pipeline.text_encoder, "dtype", torch.float16)
    c_lat = pipeline.vae.config.latent_channels
    # tiles are independent pipeline runs: inside a tile the pipeline must not shard windows over the same ranks
    saved_group = pipeline.process_group
    solo = None
    if world > 1:
        solo = _solo_group(rank, world)
    try:
        pipeline.process_group = solo if world > 1 else saved_group
        for i, tl in enumerate(plan):
            py0, py1, px0, px1 = tl.in_box
            tile = image[:, :, :, py0:py1, px0:px1]
            # the generator stream of the reference: per tile, first the LR noise, then the initial latents
            noise = randn_tensor(tile.shape, generator=generator, device=image.device, dtype=dtype)
            latents = randn_tensor((b, c_lat, t, py1 - py0, px1 - px0), generator=generator, device=image.device, dtype=dtype)
            if i % world != rank:
                continue
            flows = None
            if flows_bi is not None:
                flows = [f[:, :, :, py0:py1, px0:px1] for f in flows_bi]
            res = pipeline(image=tile, flows_bi=flows, noise=noise, latents=latents, **pipe_kwargs).images
            oy0, oy1, ox0, ox1 = tl.out_box
            sy0, sy1, sx0, sx1 = tl.src_box
            out[:, :, :, oy0:oy1, ox0:ox1] = res[:, :, :, sy0:sy1, sx0:sx1]
    finally:
        pipeline.process_group = saved_group
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(out, group=process_group)  # paste regions are disjoint: sum == union
    return out
