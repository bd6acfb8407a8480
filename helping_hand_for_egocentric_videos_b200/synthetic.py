"""Synthetic weights / clips for benchmarks and smoke tests (there are no checkpoints or datasets offline).

`randomize_` re-draws every parameter of a module in place from a seeded generator so that no part of the path is
degenerate: the reference constructs the temporal attention with zero qkv weights (model/LaviLa.py:236-242), which
would leave the temporal kernel with nothing to compute.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn


@torch.no_grad()
def randomize_(module: nn.Module, seed: int = 0) -> nn.Module:
    g = torch.Generator().manual_seed(seed)
    for name, p in sorted(module.named_parameters()):
        leaf = name.split(".")[-1]
        r = torch.randn(p.shape, generator=g, dtype=torch.float32)
        is_norm = ("norm" in name or "ln_" in name) and p.dim() == 1
        if is_norm and leaf == "weight":
            t = 1.0 + 0.1 * r
        elif p.dim() <= 1 and name != "logit_scale":
            t = 0.05 * r
        elif name == "logit_scale":
            t = torch.tensor(math.log(1 / 0.07))
        elif name in ("query_embed.weight", "frame_index.weight"):
            t = 0.5 * r
        elif leaf in ("pos_embed", "temporal_embed", "cls_token", "positional_embedding") or name == "token_embedding.weight":
            t = 0.1 * r
        else:
            fan_in = 1
            for d in p.shape[1:]:
                fan_in *= d
            t = r * (0.8 / math.sqrt(fan_in))
        p.copy_(t.to(p.dtype))
    return module


@torch.no_grad()
def randomize_on_device_(module: nn.Module, seed: int = 0) -> nn.Module:
    """randomize_ with the numbers drawn on the parameters' own (CUDA) device: the same scaling rules, a different
    stream -- for benchmark modules built directly on the GPU (seconds saved per 400 M-parameter tower)."""
    first = next(module.parameters())
    g = torch.Generator(device=first.device).manual_seed(seed)
    for name, p in sorted(module.named_parameters()):
        leaf = name.split(".")[-1]
        if name == "logit_scale":
            p.fill_(math.log(1 / 0.07))
            continue
        r = torch.randn(p.shape, generator=g, dtype=torch.float32, device=p.device)
        is_norm = ("norm" in name or "ln_" in name) and p.dim() == 1
        if is_norm and leaf == "weight":
            t = 1.0 + 0.1 * r
        elif p.dim() <= 1:
            t = 0.05 * r
        elif name in ("query_embed.weight", "frame_index.weight"):
            t = 0.5 * r
        elif leaf in ("pos_embed", "temporal_embed", "cls_token", "positional_embedding") or name == "token_embedding.weight":
            t = 0.1 * r
        else:
            fan_in = 1
            for d in p.shape[1:]:
                fan_in *= d
            t = r * (0.8 / math.sqrt(fan_in))
        p.copy_(t.to(p.dtype))
    return module


def synthetic_clips(batch: int, frames: int, size: int = 224, seed: int = 1234, device="cuda", pinned=False):
    """N(0,1) pixels: clips are mean/std normalised in the real pipeline (run/test_EgoMCQ.py:230-233)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, frames, 3, size, size, generator=g)
    if pinned:
        return x.pin_memory()
    return x.to(device)
