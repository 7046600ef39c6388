"""Drop-ins for the reference's frame-wise LPC synthesis modules (models/lpc.py), backed by the sm_100a kernels:

  lpc_synthesis(source, gains, a)   models/lpc.py:11-16    lfilter-shaped per-channel all-pole (golf_lfilter_allpole_*)
  LPCSynth                          models/lpc.py:19-59    one utterance: ex [T], lpc [F, 1+M] (gain first)
  BatchLPCSynth                     models/lpc.py:62-91    ex [B,T], gain [B,F], a [B,F,M]        (golf_lpc_frames_*)
  BatchSecondOrderLPCSynth          models/lpc.py:94-131   ex [B,T], gain [B,F], biquads [B,F,K,3] (golf_biquad_cascade_*)

Same constructor arguments and the same `_kernel` buffer (diag(window) as a conv_transpose1d weight, [win,1,win]) so
state dicts are interchangeable; the kernels read the window from its diagonal.  Frames are zero-padded by
(window_size - hop_length) // 2 and carry one gain per frame (no interpolation) -- unlike LTVMinimumPhaseFilter.
All differentiable (ex, gain, coefficients).  CUDA tensors only (GolfError otherwise; no CPU path).
"""
from __future__ import annotations

import torch
from torch import Tensor, nn

from . import functional as G
from .functional import lpc_synthesis  # noqa: F401  (re-export under the reference's name)
from .utils import get_window_fn

__all__ = ["lpc_synthesis", "LPCSynth", "BatchLPCSynth", "BatchSecondOrderLPCSynth"]


class LPCSynth(nn.Module):
    def __init__(self, hop_length: int, window_size: int = None, window: str = "hann"):
        super().__init__()
        window_fn = get_window_fn(window)
        self.hop_length = hop_length
        self.window_size = hop_length * 4 if window_size is None else window_size
        self.padding = (self.window_size - self.hop_length) // 2
        self.register_buffer("_kernel", torch.diag(window_fn(self.window_size).float()).unsqueeze(1))

    def _window(self) -> Tensor:
        return torch.diagonal(self._kernel[:, 0, :]).contiguous()

    def forward(self, ex: Tensor, lpc: Tensor):
        assert ex.ndim == 1
        assert lpc.ndim == 2
        n_frames = (ex.shape[0] + 2 * self.padding - self.window_size) // self.hop_length + 1
        assert n_frames == lpc.shape[0], f"{n_frames} != {lpc.shape}"
        gain, a = lpc[..., 0], lpc[..., 1:]
        return G.lpc_frames(ex[None], gain[None], a[None], self._window(), self.hop_length)[0]


class BatchLPCSynth(LPCSynth):
    def forward(self, ex: Tensor, gain: Tensor, a: Tensor):
        assert ex.ndim == 2
        assert gain.ndim == 2
        assert a.ndim == 3
        assert a.shape[1] == gain.shape[1]
        return G.lpc_frames(ex, gain, a, self._window(), self.hop_length)


class BatchSecondOrderLPCSynth(LPCSynth):
    def forward(self, ex: Tensor, gain: Tensor, biquads: Tensor):
        assert ex.ndim == 2
        assert gain.ndim == 2
        assert biquads.ndim == 4 and biquads.shape[-1] == 3
        return G.biquad_ff(ex, gain, biquads, self._window(), self.hop_length)
