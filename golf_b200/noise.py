"""Noise sources (models/noise.py:20-35).  Only the generator the GOLF configs use."""
from __future__ import annotations

import torch

from .ctrl import Controllable

__all__ = ["NoiseInterface", "StandardNormalNoise"]

RNG_SLOT = [0]  # which generator state fused noise uses: 0 = eager calls; GraphedSynth gives every capture its own


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
        self._rng, self._seed = {}, None

    def rng_state(self, device) -> torch.Tensor:
        """{seed, offset} (int64[2]) of the in-kernel generator on `device`; the seed is drawn from torch's generator
        at first use, the offset advances with every fused decoder pass.  Each CUDA-graph capture of the decoder
        (golf_b200.graphs.GraphedSynth sets RNG_SLOT) gets a state of its own, so passes replayed concurrently on
        different streams neither share a draw nor race on the offset."""
        from . import functional as G

        key = (torch.device(device), RNG_SLOT[0])
        if key not in self._rng:
            if self._seed is None:
                self._seed = int(torch.randint(0, 2**61, (1,), dtype=torch.int64).item())
            self._rng[key] = G.new_rng_state(key[0], (self._seed + 0x9E3779B97F4A7C15 * RNG_SLOT[0]) % (2**63))
        return self._rng[key]

    def forward(self, ref, *args, **kwargs):
        return torch.randn_like(ref)
