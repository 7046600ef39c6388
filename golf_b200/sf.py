"""Source-filter decoder wiring (models/sf.py:13-64): oscillator (+ filtered noise) ->
end filter -> room filter.  Child names and order are part of the plugin API: they fix the
layout of the encoder's output (harm_oscillator | noise_generator | noise_filter |
end_filter | room_filter)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn.functional as F

from .audiotensor import AudioTensor
from .ctrl import PassThrough, Synth
from .filters import LTVZeroPhaseFIRFilter


class SourceFilterSynth(Synth):
    def __init__(self, harm_oscillator, noise_generator, noise_filter, end_filter, room_filter=None,
                 subtract_harmonics: bool = True):
        super().__init__()
        self.subtract_harmonics = subtract_harmonics
        self.harm_oscillator = harm_oscillator
        self.noise_generator = noise_generator
        self.noise_filter = noise_filter
        self.end_filter = end_filter
        self.room_filter = room_filter if room_filter is not None else PassThrough()

    def forward(self, phase, harm_oscillator_params: Tuple, noise_generator_params: Tuple, noise_filter_params: Tuple,
                end_filter_params: Tuple, voicing: Optional[AudioTensor] = None, target: Optional[AudioTensor] = None,
                **other_params):
        harm = self.harm_oscillator(phase, *harm_oscillator_params)
        if voicing is not None:
            assert torch.all(voicing >= 0) and torch.all(voicing <= 1)
            harm = harm * F.threshold(voicing, 0.5, 0)
        noise = self.noise_generator(harm, *noise_generator_params)
        if isinstance(self.noise_filter, LTVZeroPhaseFIRFilter):
            src = self.noise_filter(noise, *noise_filter_params, add=harm)  # harm + filtered noise, one pass
        else:
            src = harm + self.noise_filter(noise, *noise_filter_params)
        if self.subtract_harmonics:
            src = src - self.noise_filter(harm, *noise_filter_params)
        if target is not None:
            return self.end_filter.reverse(src, target, *end_filter_params)
        return self.room_filter(self.end_filter(src, *end_filter_params))
