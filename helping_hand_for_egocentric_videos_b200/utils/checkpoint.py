"""Checkpoint adapters the reference scripts apply before ``load_state_dict`` (host-side dictionary work, no kernels):

* ``strip_module_prefix``      -- LaViLa checkpoints are saved from DistributedDataParallel (run/test_EgoMCQ.py:221-226)
* ``remap_keys``               -- OpenAI-CLIP visual keys -> TimeSformer keys (model/LaviLa.py:19-53)
* ``inflate_positional_embeds``-- temporal embedding of a T'-frame checkpoint -> T frames (run/test_egtea.py:46-96)
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn.functional as F


def strip_module_prefix(state_dict, prefix: str = "module."):
    """Drop a leading ``module.`` from every key (keys without it are kept as they are)."""
    return OrderedDict((k[len(prefix):] if k.startswith(prefix) else k, v) for k, v in state_dict.items())


_CLIP_TO_TIMESFORMER = {
    "class_embedding": "cls_token", "positional_embedding": "pos_embed", "conv1.weight": "patch_embed.proj.weight",
    "ln_pre.weight": "ln_pre.weight", "ln_pre.bias": "ln_pre.bias", "ln_post.weight": "norm.weight",
    "ln_post.bias": "norm.bias",
}
_CLIP_BLOCK = {
    "attn.in_proj_weight": "attn.qkv.weight", "attn.in_proj_bias": "attn.qkv.bias", "attn.out_proj.weight": "attn.proj.weight",
    "attn.out_proj.bias": "attn.proj.bias", "ln_1.weight": "norm1.weight", "ln_1.bias": "norm1.bias",
    "mlp.c_fc.weight": "mlp.fc1.weight", "mlp.c_fc.bias": "mlp.fc1.bias", "mlp.c_proj.weight": "mlp.fc2.weight",
    "mlp.c_proj.bias": "mlp.fc2.bias", "ln_2.weight": "norm2.weight", "ln_2.bias": "norm2.bias",
}


def remap_keys(clip_state_dict, transformer_layers=12):
    """CLIP visual-tower state_dict -> SpaceTimeTransformer keys.  ``proj`` is skipped (loaded separately, possible dim
    mismatch); the class / positional embeddings gain their leading singleton dims; a key outside the table (e.g. a
    block index >= transformer_layers) raises KeyError, as in the reference."""
    out = OrderedDict()
    for key, value in clip_state_dict.items():
        if key == "proj":
            continue
        if key in _CLIP_TO_TIMESFORMER:
            new = _CLIP_TO_TIMESFORMER[key]
            if key == "class_embedding":
                value = value.unsqueeze(0).unsqueeze(0)
            elif key == "positional_embedding":
                value = value.unsqueeze(0)
        else:
            parts = key.split(".")
            tail = ".".join(parts[3:])
            if parts[:2] != ["transformer", "resblocks"] or not parts[2].isdigit() \
                    or int(parts[2]) >= transformer_layers or tail not in _CLIP_BLOCK:
                raise KeyError(key)
            new = "blocks.%s.%s" % (parts[2], _CLIP_BLOCK[tail])
        out[new] = value
    return out


def inflate_positional_embeds(current_model_state_dict, new_state_dict, num_frames=4, load_temporal_fix='bilinear',
                              name='visual.temporal_embed', dim=1):
    """Resize ``new_state_dict[name]`` ([1, T', D]) to ``num_frames``: truncate when the checkpoint has more frames,
    otherwise zero-pad ('zeros') or interpolate ('interp' = nearest, 'bilinear').  Modifies and returns new_state_dict."""
    if name not in new_state_dict or name not in current_model_state_dict:
        return new_state_dict
    loaded = new_state_dict[name]
    have, want, width = loaded.shape[dim], num_frames, loaded.shape[-1]
    if have > want:
        new_state_dict[name] = loaded[:, :want, :]
    elif have < want:
        if load_temporal_fix == 'zeros':
            grown = torch.zeros([loaded.shape[0], want, width])
            grown[:, :have] = loaded
        elif load_temporal_fix in ('interp', 'bilinear'):
            mode = 'bilinear' if load_temporal_fix == 'bilinear' else 'nearest'
            grown = F.interpolate(loaded.unsqueeze(0), (want, width), mode=mode).squeeze(0)
        else:
            raise NotImplementedError
        new_state_dict[name] = grown
    if new_state_dict[name].shape[dim] != current_model_state_dict[name].shape[dim]:
        raise NotImplementedError(
            'Loading models with different spatial resolution / patch number not yet implemented, sorry.')
    return new_state_dict
