"""Host-side helpers that keep the reference's `micro_diffusion.models.utils` names importable
(train.py:9 imports `text_encoder_embedding_format` from here)."""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

DATA_TYPES = {"float16": torch.float16, "bfloat16": torch.bfloat16, "float32": torch.float32}


def text_encoder_embedding_format(enc: str) -> Tuple[int, int]:
    """(sequence length, embedding width) of the supported text encoders (reference utils.py:501-513)."""
    if enc in ("stabilityai/stable-diffusion-2-base", "runwayml/stable-diffusion-v1-5", "CompVis/stable-diffusion-v1-4"):
        return 77, 1024
    if enc == "openclip:hf-hub:apple/DFN5B-CLIP-ViT-H-14-378":
        return 77, 1024
    if enc == "DeepFloyd/t5-v1_1-xxl":
        return 120, 4096
    raise ValueError(f"Please specifcy the sequence and embedding size of {enc} encoder")


def get_2d_sincos_pos_embed(embed_dim: int, grid_size: int, pos_interp_scale: float = 1.0, base_size: int = 16) -> np.ndarray:
    """Fixed 2-D sin-cos table, [grid*grid, embed_dim] (reference utils.py:330-379): fp32 coordinates divided by
    (grid/base)/scale, fp64 frequencies 10000^(-i/(dim/4)), column-coordinate half first, [sin | cos] per axis."""
    assert embed_dim % 4 == 0
    coords = np.arange(grid_size, dtype=np.float32) / (grid_size / base_size) / pos_interp_scale
    col, row = np.meshgrid(coords, coords)
    omega = 1.0 / 10000 ** (np.arange(embed_dim // 4, dtype=np.float64) / (embed_dim / 4.0))
    halves = []
    for pos in (col, row):
        ang = pos.reshape(-1)[:, None] * omega[None, :]
        halves += [np.sin(ang), np.cos(ang)]
    return np.concatenate(halves, axis=1)


class DistLoss:
    """Sum-reduced running loss (reference utils.py:598-613 is a torchmetrics.Metric; torchmetrics is used when
    installed so Composer's metric plumbing sees the type it expects)."""

    def __new__(cls, **kwargs):
        try:
            from torchmetrics import Metric
        except ImportError:
            return super().__new__(cls)

        class _DistLoss(Metric):
            def __init__(self, **kw):
                super().__init__(**kw)
                self.add_state("loss", default=torch.tensor(0.0), dist_reduce_fx="sum")
                self.add_state("batches", default=torch.tensor(0), dist_reduce_fx="sum")

            def update(self, value):
                self.loss += value
                self.batches += 1

            def compute(self):
                return self.loss.float() / self.batches

        return _DistLoss(**kwargs)

    def __init__(self, **kwargs):
        self.loss = torch.tensor(0.0)
        self.batches = torch.tensor(0)

    def update(self, value):
        self.loss = self.loss.to(value.device) + value.detach()
        self.batches = self.batches + 1

    def compute(self):
        return self.loss.float() / self.batches

    def reset(self):
        self.loss = torch.tensor(0.0)
        self.batches = torch.tensor(0)
