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
    torch's generator so seeding behaves exactly as with the reference.

    fused=True (opt-in, inference): SourceFilterSynth then lets the noise-FIR kernel draw the samples itself
    (Philox4x32-10 + Box-Muller keyed by `rng_state`, golf_noise_fir_design_fwd): same distribution, a different
    stream than torch.randn, and no [B,T] noise tensor in memory.  Calling the module directly always uses torch."""

    def __init__(self, fused: bool = False):
        super().__init__()
        self.fused = fused
        self._rng = {}

    def rng_state(self, device) -> torch.Tensor:
        """{seed, offset} (int64[2]) of the in-kernel generator on `device`; the seed is drawn from torch's generator
        at first use, the offset advances with every fused decoder pass"""
        from . import functional as G

        key = torch.device(device)
        if key not in self._rng:
            self._rng[key] = G.new_rng_state(key)
        return self._rng[key]

    def forward(self, ref, *args, **kwargs):
        return torch.randn_like(ref)
