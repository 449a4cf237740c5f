"""Deterministic, non-degenerate synthetic weights and batches (TEST INFRASTRUCTURE ONLY).

The reference zero-initialises every adaLN weight, the prompt-block output projections and the final
linear (dit.py:615-627), so a parity check at default init is vacuous (output exactly 0).  These
helpers fill a state_dict with fan-in-scaled Gaussians as a pure function of (key, shape, seed), so the
reference (dev container) and the CUDA path (GPU box) can be given identical weights without shipping
them.
"""
from __future__ import annotations

import zlib

import torch


def synth_tensor(key: str, shape, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31 - 1))
    shape = tuple(shape)
    if len(shape) == 1:
        if key.endswith("bias"):
            return 0.1 * torch.randn(shape, generator=g)
        return 1.0 + 0.1 * torch.randn(shape, generator=g)  # LayerNorm gains
    if len(shape) == 2:
        fan_in = shape[1]
    elif len(shape) == 3:  # expert banks (E, in, out)
        fan_in = shape[1]
    else:  # conv (out, in, kh, kw)
        fan_in = shape[1] * shape[2] * shape[3]
    gain = 1.0
    if "adaLN_modulation" in key:
        gain = 0.5  # keep (1+scale) and the gates O(1) but not wild
    return gain * torch.randn(shape, generator=g) / fan_in ** 0.5


def synth_state_dict(template: dict, seed: int, keep=("pos_embed", "mask_token")) -> dict:
    """template: key -> tensor (only shapes are used, except for buffers in `keep`)."""
    out = {}
    for k in sorted(template):
        out[k] = template[k].detach().clone().float() if k in keep else synth_tensor(k, template[k].shape, seed)
    return out


def synth_batch(b: int, c: int, res: int, seed: int, cap_len: int = 77, cap_dim: int = 1024, drop_prob: float = 0.1):
    """The batch contract of latents_loader.py:43-70: fp16 latents (already VAE-scaled), fp16 caption
    embeddings (B,1,77,1024), float64 keep-mask for caption dropout."""
    g = torch.Generator().manual_seed(seed)
    lat = (0.8 * torch.randn(b, c, res, res, generator=g)).half()
    cap = torch.randn(b, 1, cap_len, cap_dim, generator=g).half()
    drop = (torch.rand(b, generator=g) >= drop_prob).double()
    return {"image_latents": lat, "caption_latents": cap, "drop_caption_mask": drop}


def replay_draws(seed: int, x_shape, tokens: int, mask_ratio: float):
    """The three draws the reference makes from the global CPU generator, in order
    (model.py:182 randn, model.py:188 randn_like, utils.py:390 rand)."""
    g = torch.Generator().manual_seed(seed)
    rnd = torch.randn([x_shape[0], 1, 1, 1], generator=g)
    eps = torch.randn(x_shape, generator=g)
    noise = torch.rand(x_shape[0], tokens, generator=g) if mask_ratio > 0 else None
    return rnd, eps, noise
