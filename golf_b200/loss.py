"""Multi-scale spectral loss of the GOLF training step on the B200 tensor cores -- drop-in for `loss.spec.SSSLoss` /
`loss.spec.MSSLoss` (loss/spec.py:11-67; criterion of cfg/ae/vctk.yaml:58-67: n_ffts 509 / 1021 / 2053, alpha 1, overlap 0.75,
window "hann").

    loss = ratio * sum_scales [ mean |S_pred - S_true| + alpha * mean |log2(S_true + eps) - log2(S_pred + eps)| ],
    S = torchaudio.transforms.Spectrogram(n_fft, hop_length=int(n_fft - n_fft * overlap), power=1, window_fn=hann)

The reference computes six STFTs with cuFFT; the shipped sizes are primes, so each is a Bluestein transform.  Here the
STFT is a DFT-as-GEMM on tcgen05 (csrc/mss.cu): framing + window, a TMA-fed `tcgen05.mma.kind::tf32` GEMM with
error-compensated products whose epilogue turns (re, im) into magnitudes and loss terms, and -- in the same forward call --
the gradient with respect to the prediction (adjoint GEMM + overlap-add), so `backward()` is one multiply.  CUDA float32
tensors only; raises GolfError otherwise (no CPU path).
"""
from __future__ import annotations

import ctypes
from typing import Sequence

import torch
from torch import nn

from . import _lib
from ._lib import GolfError, check
from .audiotensor import plain
from .functional import _cuda_f32_view, _on, _ptr, _stream, _workspace

__all__ = ["SSSLoss", "MSSLoss", "mss_loss"]

_TABLES = {}  # (device index, n_fft) -> DFT bases (never evicted: captured graphs may hold the pointers)


def _tables(n_fft: int, device: torch.device) -> torch.Tensor:
    key = (device.index if device.index is not None else torch.cuda.current_device(), int(n_fft))
    t = _TABLES.get(key)
    if t is None:
        lib = _lib.lib()
        nbytes = lib.golf_mss_tables_bytes(int(n_fft))
        if nbytes == 0:
            raise GolfError(f"mss_loss: unsupported n_fft {n_fft}")
        t = torch.empty(nbytes // 4, dtype=torch.float32, device=device)
        with _on(device):
            check(lib.golf_mss_build_tables(int(n_fft), _ptr(t), _stream()), "golf_mss_build_tables")
        torch.cuda.current_stream(device).synchronize()  # other streams may use the table next
        _TABLES[key] = t
    return t


def _rows2(t, name):
    t = _cuda_f32_view(plain(t) if isinstance(t, torch.Tensor) else t, name)
    if t.ndim != 2:
        raise GolfError(f"mss_loss: {name} must be [B, L], got {tuple(t.shape)}")
    return t if t.stride(1) == 1 else t.contiguous()


class _MSS(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, n_ffts, hops, alpha, ratio, eps, prec3):
        if ctx.needs_input_grad[1]:
            raise GolfError("mss_loss: the target is data (no gradient is computed for it); detach it")
        p, t = _rows2(pred, "pred"), _rows2(target, "target")
        if p.shape != t.shape:
            raise GolfError(f"mss_loss: pred {tuple(p.shape)} vs target {tuple(t.shape)}")
        B, L = p.shape
        dev = p.device
        lib = _lib.lib()
        n = len(n_ffts)
        c_ffts = (ctypes.c_int * n)(*[int(v) for v in n_ffts])
        c_hops = (ctypes.c_int * n)(*[int(v) for v in hops])
        tabs = [_tables(v, dev) for v in n_ffts]
        c_tabs = (ctypes.c_void_p * n)(*[tb.data_ptr() for tb in tabs])
        nbytes = lib.golf_mss_workspace_bytes(B, L, c_ffts, c_hops, n)
        if nbytes == 0:
            raise GolfError(f"mss_loss: unsupported configuration B={B} L={L} n_ffts={list(n_ffts)} hops={list(hops)}")
        ws = _workspace(nbytes, dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        need_grad = ctx.needs_input_grad[0]
        d_pred = torch.empty(B, L, dtype=torch.float32, device=dev) if need_grad else None
        with _on(dev):
            rc = lib.golf_mss_loss(_ptr(p), p.stride(0), _ptr(t), t.stride(0), B, L, c_ffts, c_hops, n, c_tabs, float(alpha), float(ratio),
                                   float(eps), _ptr(loss), _ptr(d_pred), L, int(prec3), _ptr(ws), ws.numel(), _stream())
        check(rc, "golf_mss_loss")
        ctx.d_pred = d_pred
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        d = ctx.d_pred
        return (None if d is None else d * g), None, None, None, None, None, None, None


def mss_loss(pred, target, n_ffts: Sequence[int], alpha: float = 1.0, ratio: float = 1.0, overlap: float = 0.75, eps: float = 1e-8,
             precision: int = 3) -> torch.Tensor:
    """the functional form; differentiable in `pred` (the target is data).  precision: 3 = error-compensated (3 x TF32,
    float32-grade) products in the forward and the adjoint GEMMs (default); 1 = forward only (the adjoint runs single TF32
    products: gradients good to ~1e-3 on noise-like spectra, worse in deep spectral valleys; 0.2 ms faster at B = 32 x 2 s)"""
    hops = [int(n - n * overlap) for n in n_ffts]
    return _MSS.apply(pred, target, tuple(int(n) for n in n_ffts), tuple(hops), float(alpha), float(ratio), float(eps), int(precision))


_SPECTROGRAM_DEFAULTS = {"win_length": None, "pad": 0, "normalized": False, "wkwargs": None, "center": True, "pad_mode": "reflect",
                         "onesided": True, "return_complex": None, "power": 1}


def _check_window(window: str, kwargs: dict):
    """the kernels implement torchaudio.transforms.Spectrogram at its defaults (the only way the reference's configs call it:
    cfg/ae/vctk.yaml:58-67, ckpts/ismir23/*/config.yaml `criterion`); explicit arguments are accepted when they say the same"""
    if window not in ("hann", "hanning"):
        raise ValueError(f"golf_b200.loss: only the Hann window is built into the kernels (got {window!r})")
    for k, v in kwargs.items():
        if k in ("n_fft", "hop_length"):
            continue
        if k not in _SPECTROGRAM_DEFAULTS:
            raise ValueError(f"golf_b200.loss: unsupported Spectrogram argument {k!r}")
        ok = v == _SPECTROGRAM_DEFAULTS[k] or (k == "onesided" and v is None) or (k == "win_length" and v == kwargs.get("n_fft"))
        if not ok:
            raise ValueError(f"golf_b200.loss: Spectrogram argument {k}={v!r} differs from the default the kernels implement")


class SSSLoss(nn.Module):
    """single-scale spectral loss (loss/spec.py:11-31)"""

    eps = 1e-8

    def __init__(self, alpha: float = 1.0, window: str = "hann", **kwargs):
        super().__init__()
        _check_window(window, kwargs)
        self.alpha = alpha
        self.n_fft = int(kwargs.get("n_fft", 400))
        self.hop_length = int(kwargs.get("hop_length", self.n_fft // 2))

    def forward(self, pred, target):
        return _MSS.apply(pred, target, (self.n_fft,), (self.hop_length,), float(self.alpha), 1.0, float(self.eps), 3)


class MSSLoss(nn.Module):
    """multi-scale spectral loss (loss/spec.py:34-67); all scales in one call"""

    def __init__(self, n_ffts: list, alpha=1.0, ratio=1.0, overlap=0.75, window: str = "hann", precision: int = 3, **kwargs):
        super().__init__()
        _check_window(window, kwargs)
        self.precision = int(precision)
        self.n_ffts = [int(n) for n in n_ffts]
        self.hops = [int(n - n * overlap) for n in self.n_ffts]
        self.alpha, self.ratio = float(alpha), float(ratio)

    def forward(self, x_pred, x_true):
        return _MSS.apply(x_pred, x_true, tuple(self.n_ffts), tuple(self.hops), self.alpha, self.ratio, 1e-8, self.precision)
