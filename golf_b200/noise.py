"""Noise sources (models/noise.py:20-35).  Only the generator the GOLF configs use."""
from __future__ import annotations

import torch

from .ctrl import Controllable

__all__ = ["NoiseInterface", "StandardNormalNoise"]


class NoiseInterface(Controllable):
    def forward(self, ref, *args, **kwargs):
        raise NotImplementedError


class StandardNormalNoise(NoiseInterface):
    """White N(0,1) noise shaped like `ref` (models/noise.py:30-35).  The draw comes from
    torch's generator so seeding behaves exactly as with the reference."""

    def forward(self, ref, *args, **kwargs):
        return torch.randn_like(ref)
