"""Source-filter decoder wiring (models/sf.py:13-64): oscillator (+ filtered noise) ->
end filter -> room filter.  Child names and order are part of the plugin API: they fix the
layout of the encoder's output (harm_oscillator | noise_generator | noise_filter |
end_filter | room_filter)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn.functional as F

from .audiotensor import AudioTensor, hop_of, like, plain
from .ctrl import PassThrough, Synth
from .filters import LTVMinimumPhaseFilterPrecise, LTVZeroPhaseFIRFilter
from .noise import StandardNormalNoise
from .synth import IndexedGlottalFlowTable

# Inference runs the three independent branches of the decoder on separate CUDA streams
# (forked from and joined back into the caller's stream; captured as parallel branches by a CUDA
# graph): the oscillator, the noise draw + FIR design, and the end filter's chunk transition
# matrices (which need only the coefficients).  "off" keeps everything on one stream.
CONCURRENT = "auto"
_SIDE_STREAMS = {}


def _side_streams(dev: torch.device, n: int):
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _SIDE_STREAMS or len(_SIDE_STREAMS[key]) < n:
        _SIDE_STREAMS[key] = [torch.cuda.Stream(device=dev) for _ in range(n)]
    return _SIDE_STREAMS[key]


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
        if voicing is None and target is None and self._can_run_concurrent(phase, noise_filter_params, end_filter_params):
            return self._forward_concurrent(phase, harm_oscillator_params, noise_filter_params, end_filter_params)
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

    # ------------------------------------------------------------- concurrent inference path
    def _can_run_concurrent(self, phase, noise_filter_params, end_filter_params) -> bool:
        if CONCURRENT == "off" or torch.is_grad_enabled() or self.subtract_harmonics:
            return False
        return (isinstance(self.harm_oscillator, IndexedGlottalFlowTable) and type(self.noise_generator) is StandardNormalNoise
                and type(self.noise_filter) is LTVZeroPhaseFIRFilter and type(self.end_filter) is LTVMinimumPhaseFilterPrecise
                and len(noise_filter_params) == 1 and len(end_filter_params) == 2 and plain(phase).is_cuda
                and hop_of(noise_filter_params[0]) == hop_of(end_filter_params[0]))

    def _forward_concurrent(self, phase, harm_oscillator_params, noise_filter_params, end_filter_params):
        """Same arithmetic as the sequential path (bit-identical output for the same noise); only
        the launch order / stream assignment differs:

            main : oscillator ------------------------------+-> noise FIR (+harm) -+-> z, stitch, solve -> room
            s_fir: randn, exp, irfft (FIR design) ----------+                      |
            s_phi: end-filter chunk transition matrices  ---------------------------+
        """
        log_mag = noise_filter_params[0]
        gain, a = end_filter_params
        dev = plain(phase).device
        main = torch.cuda.current_stream(dev)
        s_phi, s_fir = _side_streams(dev, 2)
        t_osc = self.harm_oscillator.out_length(phase)
        n_taps = 2 * (plain(log_mag).shape[-1] - 1)
        hop = hop_of(log_mag)
        t_src = self.noise_filter.out_length(t_osc, plain(log_mag).shape[1], n_taps, hop)
        s_phi.wait_stream(main)
        s_fir.wait_stream(main)
        with torch.cuda.stream(s_phi):
            ws = self.end_filter.responses(t_src, gain, a)
            ws.record_stream(main)
        with torch.cuda.stream(s_fir):
            noise = torch.randn(plain(phase).shape[0], t_osc, dtype=torch.float32, device=dev)
            raw = self.noise_filter.raw_kernels(log_mag)
            noise.record_stream(main)
            raw.record_stream(main)
        harm = self.harm_oscillator(phase, *harm_oscillator_params)
        main.wait_stream(s_fir)
        src = self.noise_filter.apply_raw(like(harm, noise, 1), raw, hop, add=harm)
        main.wait_stream(s_phi)
        return self.room_filter(self.end_filter.finish(src, gain, a, ws))
