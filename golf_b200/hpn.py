"""Harmonic-plus-noise wiring of GOLF-v1 / the ISMIR-23 checkpoints (models/hpn.py:11-57):
harmonic branch through its own LTV filter, noise branch through the zero-phase FIR,
summed, then a static end filter."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .audiotensor import AudioTensor
from .ctrl import Synth


class HarmonicPlusNoiseSynth(Synth):
    def __init__(self, harm_oscillator, noise_generator, harm_filter, noise_filter, end_filter):
        super().__init__()
        self.harm_oscillator = harm_oscillator
        self.noise_generator = noise_generator
        self.harm_filter = harm_filter
        self.noise_filter = noise_filter
        self.end_filter = end_filter

    def forward(self, phase, harm_oscillator_params: Tuple, noise_generator_params: Tuple, harm_filter_params: Tuple,
                noise_filter_params: Tuple, voicing: Optional[AudioTensor] = None, **other_params):
        if voicing is not None:
            assert torch.all(voicing >= 0) and torch.all(voicing <= 1)
            phase = phase * voicing
        harm = self.harm_oscillator(phase, *harm_oscillator_params)
        noise = self.noise_generator(harm, *noise_generator_params)
        out = self.harm_filter(harm, *harm_filter_params) + self.noise_filter(noise, *noise_filter_params)
        return self.end_filter(out)
