"""Import the UNMODIFIED reference from /root/reference -- build-container only.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference's L1/L2 modules
(models.audiotensor, ctrl, utils, lpc, filters, synth, noise, sf, hpn) import and
run under torch 2.11 once five absent third-party modules are stubbed in
``sys.modules``.  The stubs below are *restatements*, labelled
as such wherever their output is used:

  pyworld, diffsptk(.functional)  -- imported by the reference but never called on
                                     the GOLF path (models/utils.py:4, filters.py:9-16)
  torch_fftconv(.functional)      -- fft_conv1d == F.conv1d mathematically
                                     (models/filters.py:6,439,447)
  torchlpc.sample_wise_lpc        -- oracle.golf_oracle.sample_wise_lpc + the adjoint
                                     recurrence for autograd (models/filters.py:17,112)
  kazane.Decimate                 -- oracle.golf_oracle.decimate (models/synth.py:6,208)

``/root/reference`` does not exist on the GPU box; ``available()`` says whether
the tree is present, and everything that needs it must skip when it is not.
"""
from __future__ import annotations

import os
import sys
import types
import typing

import torch
import torch.nn.functional as F

from . import golf_oracle as O

REF_ROOT = os.environ.get("GOLF_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "models", "audiotensor"))


class _SampleWiseLPC(torch.autograd.Function):
    """Restated torchlpc.sample_wise_lpc with its adjoint:
    u_t = g_t - sum_i A[t+i+1, i] u_{t+i+1};  dx = u;  dA[t,i] = -u_t y_{t-1-i}."""

    @staticmethod
    def forward(ctx, x, A, zi):
        y = O.sample_wise_lpc(x, A, zi).to(x.dtype)
        ctx.save_for_backward(A, y, zi if zi is not None else torch.empty(0))
        ctx.has_zi = zi is not None
        return y

    @staticmethod
    def backward(ctx, g):
        A, y, zi = ctx.saved_tensors
        B, T, M = A.shape
        # adjoint recurrence = same filter on reversed time with A shifted per tap
        Ash = torch.zeros_like(A)
        for i in range(M):
            if T - i - 1 > 0:
                Ash[:, : T - i - 1, i] = A[:, i + 1 :, i]
        u = O.sample_wise_lpc(g.flip(1), Ash.flip(1)).flip(1).to(g.dtype)
        zi_pad = zi.flip(1) if ctx.has_zi else y.new_zeros(B, M)
        ypad = torch.cat([zi_pad, y], 1)  # ypad[:, M + t] = y[t]
        hist = torch.stack([ypad[:, M - 1 - i : M - 1 - i + T] for i in range(M)], -1)
        dA = -u.unsqueeze(-1) * hist
        dzi = None
        if ctx.has_zi:
            # y[-1-j] enters sample t (t <= j) through tap i = t + j: dzi_j = -sum_t A[t,t+j] u_t
            dzi = torch.zeros_like(zi)
            for j in range(M):
                for t in range(min(T, M - j)):
                    dzi[:, j] -= A[:, t, t + j] * u[:, t]
        return u, dA, dzi


def _sample_wise_lpc(x, a, zi=None):
    return _SampleWiseLPC.apply(x, a, zi)


class _Decimate(torch.nn.Module):
    def __init__(self, q: int = 2, zeros: int = 16, **_):
        super().__init__()
        self.q, self.zeros = q, zeros
        self.register_buffer("kernel", O.decimate_kernel(q, zeros))

    def forward(self, x):
        shape = x.shape
        y = F.conv1d(x.reshape(-1, 1, shape[-1]), self.kernel[None, None].to(x.dtype), stride=self.q, padding=self.zeros * self.q)
        return y.view(*shape[:-1], -1)


def _fft_conv1d(x, w, *args, **kwargs):
    return F.conv1d(x, w, *args, **kwargs)


def install_stubs() -> None:
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    def _absent(*a, **k):
        raise RuntimeError("stubbed third-party function called on a path that should not need it")

    if "pyworld" not in sys.modules:
        mod("pyworld", dio=_absent, stonemask=_absent, harvest=_absent)
    if "diffsptk" not in sys.modules:
        names = ["MLSA", "MelCepstralAnalysis", "MelGeneralizedCepstrumToSpectrum", "PQMF", "IPQMF", "STFT"]
        d = mod("diffsptk", **{n: type(n, (torch.nn.Module,), {}) for n in names})
        d.functional = mod("diffsptk.functional", lsp2lpc=_absent)
    if "torch_fftconv" not in sys.modules:
        t = mod("torch_fftconv", fft_conv1d=_fft_conv1d)
        t.functional = mod("torch_fftconv.functional", fft_conv1d=_fft_conv1d)
    if "torchlpc" not in sys.modules:
        mod("torchlpc", sample_wise_lpc=_sample_wise_lpc)
    if "kazane" not in sys.modules:
        mod("kazane", Decimate=_Decimate)
    if not hasattr(torch, "Any"):  # models/lru/recurrence.py:28 under torch 2.11
        torch.Any = typing.Any


def import_harness():
    """The reference's Lightning-side harness, unmodified: returns (ltng.ae, test_rtf) with `lightning` provided by
    oracle/lightning_standin.py when the real package is absent."""
    import importlib

    import_reference()
    from . import lightning_standin

    try:
        importlib.import_module("lightning.pytorch")
    except ImportError:
        lightning_standin.install()
    for name, attrs in (("frechet_audio_distance", {"FrechetAudioDistance": object}), ("soundfile", {})):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except ImportError:
                m = types.ModuleType(name)
                m.__dict__.update(attrs)
                sys.modules[name] = m
    return importlib.import_module("ltng.ae"), importlib.import_module("test_rtf")


def import_reference():
    """Returns the reference's `models` package (imported from REF_ROOT, unmodified)."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.append(REF_ROOT)
    import importlib

    models = importlib.import_module("models")
    for sub in ("audiotensor", "ctrl", "utils", "lpc", "filters", "synth", "noise", "sf", "hpn"):
        importlib.import_module(f"models.{sub}")
    return models
