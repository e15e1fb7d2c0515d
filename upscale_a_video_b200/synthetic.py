"""Seeded random-init weights for benchmarking without checkpoints (the shipped reference weights are a Google-Drive
download, README.md:78-101 of the reference; there is no network on the benchmark boxes).

Every tensor is drawn from its own CPU generator seeded by (seed, crc32(key)), so a state dict is reproducible on any
machine from the key -> shape table alone.  Scales keep activations O(1) through 100+ layers; tensors the reference
zero-initialises (attention.py:490, temporal_module.py:172, resnet.py:461) get non-zero values like everything else so
that those branches do real work.  The parity tests draw the oracle's weights with the same rule
(`tests/test_host_logic.py::test_synthetic_weights_match_the_oracle_rule`)."""
from __future__ import annotations

import math
import zlib
from typing import Dict, Sequence

import torch


def seeded_tensor(key: str, shape: Sequence[int], seed: int) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 63))
    shape = tuple(shape)
    if key.endswith(".freqs"):  # rotary frequencies are a deterministic table, not a learned draw
        dim = shape[0] * 2
        return 1.0 / (10000 ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
    x = torch.randn(shape, generator=g, dtype=torch.float32)
    parts = key.split(".")
    leaf, parent = parts[-1], (parts[-2] if len(parts) > 1 else "")
    if leaf == "running_var":
        return 0.5 + 0.5 * x.abs()
    if leaf == "running_mean":
        return 0.1 * x
    if leaf == "num_batches_tracked":
        return torch.zeros(shape, dtype=torch.long)
    if leaf == "bias":
        return x * 0.05
    if "norm" in parent:
        return 1.0 + 0.1 * x
    if "class_embedding" in key:
        return x * 0.1
    if "relative_attention_bias" in key:
        return x * 0.5
    if len(shape) >= 2:
        return x * (1.0 / math.sqrt(math.prod(shape[1:])))
    return x * 0.05


def seeded_state_dict(module_or_shapes, seed: int, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """`module_or_shapes`: an nn.Module (its own state-dict keys / shapes are used) or a {key: shape} table"""
    if hasattr(module_or_shapes, "state_dict"):
        shapes = {k: tuple(v.shape) for k, v in module_or_shapes.state_dict().items()}
    else:
        shapes = module_or_shapes
    return {k: seeded_tensor(k, s, seed).to(dtype) for k, s in shapes.items()}
