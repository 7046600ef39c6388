"""Glottal-flow wavetable oscillators (models/synth.py:26-340) on the sm_100a kernels.

  GlottalFlowTable                      table construction + `generate` (bilinear read)
  IndexedGlottalFlowTable               one selection weight per control frame
  DownsampledIndexedGlottalFlowTable    + the frame-rate MLP that predicts the weight

Buffer / parameter names match the reference checkpoints: `table`, `R_d_values`
(persistent), `decimater.kernel` (non-persistent), `model.{1,3}.{weight,bias}`.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from . import functional as G
from .audiotensor import AudioTensor, hop_of, like, plain
from .ctrl import Controllable, wrap_ctrl_fn
from .utils import lf_period_v1, lf_period_v2

__all__ = [
    "OscillatorInterface",
    "GlottalFlowTable",
    "IndexedGlottalFlowTable",
    "DownsampledIndexedGlottalFlowTable",
]

# "sync": assert input ranges on the host like the reference does (one device sync per
# call, models/synth.py:26-29,219-222); "off": skip them.
CHECK_INPUTS = "sync"


def _check_phase(m, args):
    p = plain(args[0])
    assert p.ndim == 2, p.shape
    if CHECK_INPUTS == "sync":
        assert bool(((p >= 0) & (p <= 0.5)).all()), "phase (cycles/sample) must lie in [0, 0.5]"


def _check_out(m, args, out):
    assert plain(out).ndim == 2
    assert hop_of(out, -1) == 1


class OscillatorInterface(Controllable):
    def __init__(self) -> None:
        super().__init__()
        self._input_handle = self.register_forward_pre_hook(_check_phase)
        self._output_handle = self.register_forward_hook(_check_out)

    def forward(self, phase, *args, **kwargs):
        raise NotImplementedError


class Decimate(nn.Module):
    """Anti-aliasing FIR + stride (kazane.Decimate restated -- third-party, absent, parity
    unpinned): Hann-windowed sinc, `zeros` crossings per side at the output
    rate.  Only holds the taps; the convolution is fused into the oscillator kernel."""

    def __init__(self, q: int, zeros: int = 16):
        super().__init__()
        self.q, self.zeros = q, zeros
        n = torch.arange(-zeros * q, zeros * q + 1, dtype=torch.float64)
        h = torch.sinc(n / q) / q * torch.hann_window(2 * zeros * q + 1, periodic=False, dtype=torch.float64)
        self.register_buffer("kernel", h.float())


class GlottalFlowTable(OscillatorInterface):
    def __init__(self, table_size: int = 100, table_type: str = "derivative", normalize_method: str = "constant_power",
                 align_peak: bool = True, trainable: bool = False, min_R_d: float = 0.3, max_R_d: float = 2.7,
                 lf_v2: bool = False, **kwargs):
        super().__init__()
        self.register_buffer("R_d_values", torch.exp(torch.linspace(math.log(min_R_d), math.log(max_R_d), table_size)))
        if lf_v2:
            table = lf_period_v2(self.R_d_values, **kwargs)
        else:
            # the 0-dim float32 tensor goes through as it does in the reference (models/synth.py:83-84): its Newton
            # iterations then run in float32 with the math.* calls evaluated on the float32-rounded arguments
            table = torch.stack([lf_period_v1(R_d=r, **kwargs) for r in self.R_d_values])
        if table_type == "flow":
            table = table.cumsum(dim=1)
        elif table_type != "derivative":
            raise ValueError(f"unknown table_type: {table_type}")
        if align_peak:  # rotate every period so the main excitation lines up
            pos = table.argmin(1) if table_type == "derivative" else table.argmax(1)
            target = int(pos.max())
            table = torch.stack([torch.roll(row, target - int(p)) for row, p in zip(table, pos)])
        if normalize_method == "constant_power":
            table = table / table.norm(dim=1, keepdim=True) * math.sqrt(table.shape[1])
        elif normalize_method == "peak":
            if table_type == "flow":
                table = table / table.max(dim=1, keepdim=True).values
        elif normalize_method is not None:
            raise ValueError(f"unknown normalize_method: {normalize_method}")
        if trainable:  # the table adjoint is golf_glottal_osc_bwd (csrc/osc.cu)
            self.register_parameter("table", nn.Parameter(table))
        else:
            self.register_buffer("table", table)

    @staticmethod
    def generate(wrapped_phase, tables):
        """wrapped_phase [B,N] (hop 1) in [0,1), tables [B,R,P] at hop h: bilinear read over
        (time/h, phase*P) with the period wrapping around (models/synth.py:124-177)."""
        assert hop_of(wrapped_phase) == 1
        out = G.wavetable_read(plain(wrapped_phase), plain(tables), hop_of(tables))
        return like(wrapped_phase, out, 1)


class IndexedGlottalFlowTable(GlottalFlowTable):
    # "exact": running phase in 64-bit fixed point (default); "aten_cpu": the reference CPU arithmetic
    # (float64 accumulate, float32 round, then mod 1) for parity checks
    phase_accumulation = "exact"

    def __init__(self, *args, oversampling: int = 1, equal_energy: bool = False, **kwargs):
        super().__init__(*args, **kwargs)
        self.ctrl = wrap_ctrl_fn(split_size=(1,), trsfm_fn=lambda x: (torch.sigmoid(x),))
        self.equal_energy = equal_energy
        self.oversampling = oversampling
        if oversampling > 1:
            self.decimater = Decimate(oversampling)
            self.decimater.register_buffer("kernel", self.decimater.kernel, persistent=False)

    @staticmethod
    def out_length(phase) -> int:
        """samples produced for a phase track [B,Np] at hop h: (Np-1)*h + 1 (synth.py:242-262)"""
        return (plain(phase).shape[1] - 1) * hop_of(phase) + 1

    def forward(self, phase, table_select_weight, phase_offset=None):
        w = plain(table_select_weight)
        assert w.dim() == 2
        p0 = None
        if phase_offset is not None:  # models/synth.py:251-252; a constant per utterance rides on the phase prefix
            p0 = plain(phase_offset).reshape(w.shape[0], -1)
            if p0.shape[1] != 1:
                raise NotImplementedError("phase_offset: only one value per utterance ([B] or [B,1]) is fused")
            if self.phase_accumulation != "exact":
                raise NotImplementedError("phase_offset needs phase_accumulation = 'exact'")
        if CHECK_INPUTS == "sync":
            assert bool(((w >= 0) & (w <= 1)).all()), "table_select_weight must lie in [0, 1]"
        dk = self.decimater.kernel if self.oversampling > 1 else None
        y = G.glottal_osc(plain(phase), hop_of(phase), w, hop_of(table_select_weight), self.table, dk,
                          self.oversampling, self.equal_energy, self.phase_accumulation, p0)
        return like(phase, y, 1)


def get_downsampler(hop_rate: int, in_channels: int, output_channels: int) -> nn.Sequential:
    """frame-rate head: average-pool by hop_rate, 1x1 conv, GLU, 1x1 conv (models/synth.py:297-315);
    module indices 1 and 3 carry the checkpoint's weights"""
    return nn.Sequential(
        nn.AvgPool1d(kernel_size=hop_rate, stride=hop_rate, padding=hop_rate // 2),
        nn.Conv1d(in_channels, in_channels * 2, kernel_size=1),
        nn.GLU(dim=1),
        nn.Conv1d(in_channels, output_channels, kernel_size=1),
    )


class DownsampledIndexedGlottalFlowTable(IndexedGlottalFlowTable):
    def __init__(self, hop_rate: int, in_channels: int, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.hop_rate = hop_rate
        self.model = get_downsampler(hop_rate, in_channels, 1)

        def trsfm(h):
            x = plain(h).transpose(1, 2)
            if x.is_cuda:  # float32 products like the reference's CPU path (cuDNN would run these 1x1 convolutions in TF32)
                with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                    w = self.model(x)
            else:
                w = self.model(x)
            return (like(h, w.squeeze(1).sigmoid(), hop_of(h) * self.hop_rate),)

        self.ctrl = wrap_ctrl_fn(split_size=(in_channels,), trsfm_fn=trsfm)
