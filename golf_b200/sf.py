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
from . import functional as G
from .filters import LTIAcousticFilter, LTVMinimumPhaseFilter, LTVMinimumPhaseFilterPrecise, LTVZeroPhaseFIRFilter
from . import filters as filters_mod
from . import synth as synth_mod
from .noise import StandardNormalNoise
from .synth import IndexedGlottalFlowTable

# Inference scheduling.  At the reference's batch sizes most kernels of the decoder are latency
# bound (serial recurrences, short grids), so one stream leaves the GPU mostly idle.  Utterances are
# independent, so the batch is cut into SPLIT groups that run the whole decoder on their own CUDA
# streams (forked from and joined back into the caller's stream; a CUDA graph captures them as
# parallel branches): one group's serial stitch/solve overlaps another group's throughput kernels.
# Within a group the noise draw and the FIR design run beside the oscillator, each on its own stream.  "off" = one stream.
CONCURRENT = "auto"
SPLIT = 1
FUSE_ROOM = True  # GOLF-ss end filter + room FIR through golf_lpc_ss_room_fwd (False: two module calls)
FUSED = True      # inference on the shipped GOLF-ss configuration through golf_synth_fused_fwd (False: module by module)
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
            if FUSED and self._can_run_fused(phase, harm_oscillator_params, noise_filter_params, end_filter_params):
                return self._forward_fused(phase, harm_oscillator_params, noise_filter_params, end_filter_params)
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
        return self._end_and_room(src, end_filter_params)

    def _end_and_room(self, src, end_filter_params):
        """room_filter(end_filter(src, gain, a)) (models/sf.py:64).  For the sample-wise filter followed by the learned
        room FIR the two share one launch chain (golf_lpc_ss_room_fwd): the FIR runs as the last phase of the filter's
        per-sequence cluster kernel, on y while it is still in L2."""
        if (FUSE_ROOM and type(self.end_filter) is LTVMinimumPhaseFilterPrecise and type(self.room_filter) is LTIAcousticFilter
                and len(end_filter_params) == 2 and plain(src).is_cuda and self.room_filter.kernel.numel() <= 252):
            gain, a = end_filter_params
            hop, ex_hop = hop_of(gain), hop_of(src)
            if hop % ex_hop == 0 and hop_of(a, hop) == hop and plain(a).shape[1] == plain(gain).shape[1]:
                y = G.lpc_ss_room(plain(src), plain(gain), plain(a), self.room_filter.kernel, hop // ex_hop)
                return like(src, y, ex_hop)
        return self.room_filter(self.end_filter(src, *end_filter_params))

    # ------------------------------------------------------------- fused inference path
    def _can_run_fused(self, phase, harm_oscillator_params, noise_filter_params, end_filter_params) -> bool:
        """the shipped GOLF-ss inference configuration (cfg/ae/decoder/golf-precise.yaml), no phase offset"""
        lm, (gain, a) = noise_filter_params[0], end_filter_params
        hop = hop_of(lm)
        return (len(harm_oscillator_params) == 1 and hop_of(phase) >= 1 and hop_of(a, hop) == hop
                and plain(a).shape[1] == plain(gain).shape[1] == plain(lm).shape[1]
                and G.noise_fir_design_supported(plain(lm).shape[-1], hop)
                and (type(self.room_filter) is PassThrough or (type(self.room_filter) is LTIAcousticFilter and self.room_filter.kernel.numel() <= 252))
                and plain(phase).dtype == torch.float32)

    def _forward_fused(self, phase, harm_oscillator_params, noise_filter_params, end_filter_params):
        """One C call for the whole pass (golf_synth_fused_fwd): five launches, no library kernel.  The noise draw is
        torch.randn (one ATen launch, exact torch stream) unless the generator module opts into the in-kernel one."""
        osc, (w,), lm, (gain, a) = self.harm_oscillator, harm_oscillator_params, noise_filter_params[0], end_filter_params
        ph = plain(phase)
        dev = ph.device
        if synth_mod.CHECK_INPUTS == "sync":
            assert bool(((ph >= 0) & (ph <= 0.5)).all()), "phase (cycles/sample) must lie in [0, 0.5]"
            assert bool(((plain(w) >= 0) & (plain(w) <= 1)).all()), "table_select_weight must lie in [0, 1]"
        n_mag = plain(lm).shape[-1]
        noise = rng = None
        if getattr(self.noise_generator, "fused", False):
            rng = self.noise_generator.rng_state(dev)
        else:
            noise = torch.randn(ph.shape[0], osc.out_length(phase), dtype=torch.float32, device=dev)
        room_k = self.room_filter.kernel if type(self.room_filter) is LTIAcousticFilter else None
        dk = osc.decimater.kernel if osc.oversampling > 1 else None
        y = G.synth_fused(ph, hop_of(phase), plain(w), hop_of(w), osc.table, dk, osc.oversampling, osc.equal_energy,
                          osc.phase_accumulation, plain(lm), self.noise_filter._window(2 * (n_mag - 1), plain(lm), scaled=False),
                          plain(gain), plain(a), hop_of(lm), room_k, noise, rng)
        return like(phase, y, 1)

    # ------------------------------------------------------------- concurrent inference path
    def _can_run_concurrent(self, phase, noise_filter_params, end_filter_params) -> bool:
        if CONCURRENT == "off" or torch.is_grad_enabled() or self.subtract_harmonics:
            return False
        return (isinstance(self.harm_oscillator, IndexedGlottalFlowTable) and type(self.noise_generator) is StandardNormalNoise
                and type(self.noise_filter) is LTVZeroPhaseFIRFilter and type(self.end_filter) is LTVMinimumPhaseFilterPrecise
                and len(noise_filter_params) == 1 and len(end_filter_params) == 2 and plain(phase).is_cuda
                and hop_of(noise_filter_params[0]) == hop_of(end_filter_params[0]))

    def _forward_concurrent(self, phase, harm_oscillator_params, noise_filter_params, end_filter_params):
        """Same arithmetic per utterance as the sequential path; the noise generator is called once
        per group, so the draw differs from a single full-batch randn (same distribution)."""
        dev = plain(phase).device
        B = plain(phase).shape[0]
        n = max(1, min(int(SPLIT), B))
        main = torch.cuda.current_stream(dev)
        streams = _side_streams(dev, 3 * n)
        bounds = [(B * i) // n for i in range(n + 1)]
        cut = lambda x, lo, hi: like(x, plain(x)[lo:hi], hop_of(x))
        outs = []
        for i in range(n):
            lo, hi = bounds[i], bounds[i + 1]
            s_run, s_fir, s_rng = streams[3 * i], streams[3 * i + 1], streams[3 * i + 2]
            s_run.wait_stream(main)
            with torch.cuda.stream(s_run):
                y = self._forward_group(s_run, s_fir, s_rng, cut(phase, lo, hi), tuple(cut(x, lo, hi) for x in harm_oscillator_params),
                                        cut(noise_filter_params[0], lo, hi), tuple(cut(x, lo, hi) for x in end_filter_params))
                plain(y).record_stream(main)
            outs.append(y)
        for i in range(n):
            main.wait_stream(streams[3 * i])
        if n == 1:
            return outs[0]
        return like(outs[0], torch.cat([plain(o) for o in outs], 0), hop_of(outs[0]))

    def _forward_group(self, s_run, s_fir, s_rng, phase, harm_oscillator_params, log_mag, end_filter_params):
        """one group:  s_run: oscillator ------------------------+-> noise FIR (+harm) -> end filter -> room
                       s_fir: exp+pack, irfft (FIR design) -------+
                       s_rng: randn ------------------------------+
        The two side branches are independent of each other and each shorter than the oscillator."""
        dev = plain(phase).device
        hop = hop_of(log_mag)
        t_osc = self.harm_oscillator.out_length(phase)
        in_kernel_design = filters_mod.FUSED_DESIGN and G.noise_fir_design_supported(plain(log_mag).shape[-1], hop)
        s_rng.wait_stream(s_run)
        with torch.cuda.stream(s_rng):
            noise = torch.randn(plain(phase).shape[0], t_osc, dtype=torch.float32, device=dev)
            noise.record_stream(s_run)
        if not in_kernel_design:
            s_fir.wait_stream(s_run)
            with torch.cuda.stream(s_fir):
                raw = self.noise_filter.raw_kernels(log_mag)
                raw.record_stream(s_run)
        harm = self.harm_oscillator(phase, *harm_oscillator_params)
        s_run.wait_stream(s_rng)
        if in_kernel_design:
            src = self.noise_filter(like(harm, noise, 1), log_mag, add=harm)
        else:
            s_run.wait_stream(s_fir)
            src = self.noise_filter.apply_raw(like(harm, noise, 1), raw, hop, add=harm)
        return self._end_and_room(src, end_filter_params)
