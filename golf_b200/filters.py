"""Drop-in filter modules of the GOLF decoder, backed by the sm_100a kernels.

Same constructor arguments, `.ctrl` protocol, buffers/parameters and forward signatures as
the reference classes of the same name (models/filters.py), so they are selected by
`class_path: golf_b200.filters.<Name>` in the reference's YAML and load its checkpoints:

  LTVMinimumPhaseFilterPrecise   models/filters.py:64-113   (GOLF-ss end filter)
  LTVMinimumPhaseFilter          models/filters.py:116-195  (GOLF-ff end filter)
  LTVZeroPhaseFIRFilter          models/filters.py:340-384  (noise filter)
  LTIAcousticFilter              models/filters.py:426-456  (room filter)

Inputs must live on a CUDA device (GolfError otherwise; no CPU fallback).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.nn.functional as F
from torch import nn

from . import functional as G
from ._lib import GolfError
from .audiotensor import AudioTensor, hop_of, like, plain
from .ctrl import Controllable, wrap_ctrl_fn
from .utils import biquads2lpc, get_logits2biquads, get_window_fn, rc2lpc

FUSED_DESIGN = True  # inference: FIR taps designed inside the FIR kernel where supported (False: cuFFT route)

__all__ = [
    "FilterInterface",
    "LTVFilterInterface",
    "LTVMinimumPhaseFilterPrecise",
    "LTVMinimumPhaseFilter",
    "LTVZeroPhaseFIRFilter",
    "LTIAcousticFilter",
    "convert2samplewise",
]


class FilterInterface(Controllable):
    def forward(self, ex, *args, **kwargs):
        raise NotImplementedError


class LTVFilterInterface(FilterInterface):
    def reverse(self, ex, *args, **kwargs):
        raise NotImplementedError


def _logits2lpc(kind: str, max_abs: float):
    if kind in ("coef", "conj", "real"):
        to_bq = get_logits2biquads(kind, max_abs)

        def fn_bq(lg):  # one CUDA launch each way (golf_biquad_params_fwd / _bwd) where it applies, else the torch product
            sections = lg.view(lg.shape[0], lg.shape[1], -1, 2)
            if lg.is_cuda and lg.dtype == torch.float32 and sections.shape[2] <= 16:
                return G.logits2lpc(sections, kind, max_abs)
            return biquads2lpc(to_bq(sections))

        return fn_bq, 0
    if kind == "rc2lpc":

        def fn(lg):  # one CUDA launch (golf_rc2lpc_fwd / _bwd) where it applies, else the torch recursion
            if lg.is_cuda and lg.dtype == torch.float32 and lg.shape[-1] <= 40:
                return G.rc2lpc(lg, max_abs)
            return rc2lpc(lg.tanh() * max_abs)

        return fn, 0
    if kind == "lsp2lpc":

        def fn(lg):
            from diffsptk.functional import lsp2lpc  # optional third-party, as in the reference

            return lsp2lpc(lg.softmax(-1).cumsum(-1).roll(1, -1) * torch.pi)[..., 1:]

        return fn, 1
    raise ValueError(f"Unknown lpc_parameterisation: {kind}")


class LTVMinimumPhaseFilterPrecise(LTVFilterInterface):
    """Sample-wise time-varying all-pole filter: y[t] = ex[t]*gain(t) - sum_i a_i(t) y[t-1-i]
    with gain and a linearly interpolated from frame rate.  One fused CUDA path
    (golf_lpc_ss_fwd/_bwd); the [B,T,M] coefficient tensor is never materialised."""

    def __init__(self, lpc_order: int = None, lpc_parameterisation: str = "rc2lpc", max_abs_value: float = 1.0):
        super().__init__()
        to_lpc, extra = _logits2lpc(lpc_parameterisation, max_abs_value)
        if lpc_order is not None:
            self.ctrl = wrap_ctrl_fn(
                split_size=(1, lpc_order + extra),
                trsfm_fn=lambda log_gain, logits: (torch.exp(log_gain), logits.new_tensor(to_lpc(plain(logits)))),
            )

    def forward(self, ex, gain, a):
        x, g, c = plain(ex), plain(gain), plain(a)
        assert x.ndim == 2 and g.ndim == 2 and c.ndim == 3
        assert c.shape[1] == g.shape[1]
        hop, ex_hop = hop_of(gain), hop_of(ex)
        assert hop % ex_hop == 0 and hop_of(a, hop) == hop, (ex_hop, hop, hop_of(a))
        return like(ex, G.lpc_ss(x, g, c, hop // ex_hop), ex_hop)

    def reverse(self, ex, y, gain, a) -> Tuple[AudioTensor, AudioTensor]:
        """inverse-filter the target (models/filters.py:186-195): returns (ex*gain, A(z) y)"""
        return _reverse(ex, y, gain, a)

    # two-call form used by SourceFilterSynth's concurrent inference path: the chunk transition
    # matrices need only `a`, so they are computed on a side stream while the source is synthesised
    def responses(self, t_ex: int, gain, a, ex_hop: int = 1):
        g, c = plain(gain), plain(a)
        hop = hop_of(gain) // ex_hop
        return G.lpc_ss_responses(c, G.lpc_ss_length(t_ex, c.shape[1], hop), hop)

    def finish(self, ex, gain, a, ws):
        ex_hop = hop_of(ex)
        return like(ex, G.lpc_ss_finish(plain(ex), plain(gain), plain(a), hop_of(gain) // ex_hop, ws), ex_hop)


def _reverse(ex, y, gain, a):
    resid = G.lpc_inverse(plain(y), plain(a), hop_of(a) // hop_of(y))
    return ex * gain, like(y, resid, hop_of(y))


def _check_ff_geometry(hop: int, W: int, M: int, needs_grad: bool) -> None:
    """What the GOLF-ff kernels are built for (golf_lpc_ff_fwd / _bwd), checked where the module is called so that an
    unsupported configuration fails in forward() with its reason -- not as UNSUPPORTED from the first backward()."""
    why = None
    if M > 40:
        why = f"LPC order {M} > 40"
    elif hop < 40:
        why = f"hop {hop} < 40"
    elif W % hop != 0 or not 2 <= W // hop <= 8:
        why = f"window_length {W} must be 2..8 times the hop {hop}"
    elif needs_grad:
        if not any(m >= M and W % m == 0 and hop % m == 0 for m in (4, 8, 12, 16, 20, 24, 32, 40)):
            why = (f"training needs a kernel order in 4, 8, 12, 16, 20, 24, 32, 40 that is >= {M} and divides both "
                   f"window_length {W} and hop {hop}; inference (torch.no_grad) works")
    if why is not None:
        raise GolfError(f"LTVMinimumPhaseFilter: {why}; use LTVMinimumPhaseFilterPrecise (any hop, order <= 40) or the reference module")


class LTVMinimumPhaseFilter(LTVMinimumPhaseFilterPrecise):
    """Frame-wise variant: windows of `window_length` every hop, per-frame LTI all-pole from
    zero state, windowed overlap-add, normalised (golf_lpc_ff_fwd).  The reference keeps a
    dense diag(window) buffer `_kernel` (non-persistent); only the window itself is kept
    here, under the same non-persistent status so state dicts stay interchangeable."""

    def __init__(self, window: str, window_length: int, centred: bool = True, **kwargs):
        super().__init__(**kwargs)
        self.register_buffer("_window", get_window_fn(window)(window_length).float(), persistent=False)
        self.centred = centred

    def forward(self, ex, gain, a):
        x, g, c = plain(ex), plain(gain), plain(a)
        assert c.shape[1] == g.shape[1]
        hop = hop_of(gain) // hop_of(ex)
        W = self._window.shape[0]
        assert W >= hop * 2, f"{W} < {hop * 2}"
        _check_ff_geometry(hop, W, c.shape[-1], torch.is_grad_enabled() and any(t.requires_grad for t in (x, g, c)))
        if not self.centred:
            x = x[..., hop // 2 :]
        y = G.lpc_ff(x, g, c, self._window, hop)
        if not self.centred:
            y = F.pad(y[:, None], (hop // 2, 0), "reflect")[:, 0]
        return like(ex, y, hop_of(ex))


class LTVZeroPhaseFIRFilter(LTVFilterInterface):
    """Time-varying zero-phase FIR from log-magnitudes: kernel_k = window * fftshift(irfft(exp(
    log_mag_k))) (frame rate, cuFFT) applied block-wise by golf_noise_fir_fwd.  `add` fuses
    the following `harm + filtered_noise` (models/sf.py:53-56)."""

    def __init__(self, window: str, conv_method: str = "direct", n_mag: int = None):
        super().__init__()
        if conv_method not in ("direct", "fft"):
            raise ValueError(f"Unknown conv_method: {conv_method}")
        self.window_fn = get_window_fn(window)
        if n_mag is not None:
            self.ctrl = wrap_ctrl_fn(split_size=(n_mag,), trsfm_fn=lambda x: (x,))

    @staticmethod
    def get_zero_phase_fir(log_mag: torch.Tensor) -> torch.Tensor:
        fir = torch.fft.irfft(torch.exp(log_mag) + 0j, dim=-1)
        return torch.fft.fftshift(fir, dim=-1)

    def windowing(self, kernel: torch.Tensor) -> torch.Tensor:
        return kernel * self.window_fn(kernel.shape[-1], device=kernel.device, dtype=kernel.dtype)

    def _window(self, K: int, like_t: torch.Tensor, scaled: bool = True) -> torch.Tensor:
        """the window on like_t's device; scaled: window / K (raw_kernels() leaves the inverse FFT unnormalised, the 1/K
        rides on the window).  Entries are never evicted (a captured CUDA graph may hold their pointers) and a new
        entry is complete before any stream can see it."""
        cache = self.__dict__.setdefault("_win_cache", {})
        key = (K, like_t.device, scaled)
        if key not in cache:
            if like_t.is_cuda and torch.cuda.is_current_stream_capturing():
                raise GolfError("LTVZeroPhaseFIRFilter: first use inside a CUDA-graph capture; run the module once eagerly first")
            w = self.window_fn(K, device=like_t.device, dtype=torch.float32)
            cache[key] = w / K if scaled else w
            if like_t.is_cuda:
                torch.cuda.current_stream(like_t.device).synchronize()
        return cache[key]

    def raw_kernels(self, log_mag) -> torch.Tensor:
        """frame-rate half of the inference path: K * irfft(exp(log_mag)) -- one kernel for exp + complex
        packing, cuFFT without its scaling pass; 1/K, fftshift and the window are applied by the FIR kernel
        while it stages the taps (apply_raw)"""
        return torch.fft.irfft(G.exp_complex(plain(log_mag)), dim=-1, norm="forward")

    def apply_raw(self, ex, raw, hop: int, add=None):
        y = G.ltv_fir_blocks(plain(ex), raw, hop // hop_of(ex), None if add is None else plain(add),
                             window=self._window(raw.shape[-1], raw))
        return like(ex, y, hop_of(ex))

    @staticmethod
    def out_length(t_ex: int, frames: int, n_taps: int, hop: int) -> int:
        return G._fir_blocks_count(t_ex, frames, n_taps, hop) * hop

    def forward(self, ex, log_mag, add=None):
        hop = hop_of(log_mag) // hop_of(ex)
        lm, x = plain(log_mag), plain(ex)
        add = None if add is None else plain(add)
        need_grad = torch.is_grad_enabled() and (lm.requires_grad or x.requires_grad or (add is not None and add.requires_grad))
        if need_grad:  # differentiable path: final taps built by torch (frame rate), FIR + adjoint in CUDA
            kernel = self.windowing(self.get_zero_phase_fir(lm))
            y = G.ltv_fir_blocks(x, kernel, hop, add)
        elif lm.is_cuda and FUSED_DESIGN and G.noise_fir_design_supported(lm.shape[-1], hop):
            # inference, shipped geometry: the taps are designed inside the FIR kernel (no FFT library, no [B,F,K] tensor)
            y = G.noise_fir_design(x, lm, self._window(2 * (lm.shape[-1] - 1), lm, scaled=False), hop, add)
        else:  # inference: cuFFT gives the raw impulse responses, shift + window ride along in the FIR kernel
            raw = self.raw_kernels(lm)
            y = G.ltv_fir_blocks(x, raw, hop, add, window=self._window(raw.shape[-1], raw))
        return like(ex, y, hop_of(ex))


class LTVZeroPhaseFIRFilterPrecise(LTVZeroPhaseFIRFilter):
    """Sample-wise twin of LTVZeroPhaseFIRFilter (models/filters.py:286-337): the K-tap kernel applied at
    sample t is the linear interpolation (AudioTensor.reduce_hop_length) of the two frame kernels around t,

        y[t] = sum_j xpad[t + j] * (l0(t) k_f[j] + l1(t) k_{f+1}[j]),   f = t // hop,

    output length min(T, (F-1)*hop + 1).  The filter is linear in the kernel, so this is
    l0(t) * (block FIR with k_f)(t) + l1(t) * (block FIR with k_{f+1})(t): two launches of the block-FIR
    kernel (golf_noise_fir_fwd) over the same input, one with the kernel tensor shifted by a frame, and an
    element-wise blend -- the [B, T, K] interpolated-kernel tensor of the reference (K floats per SAMPLE)
    never exists.  Differentiable through the block FIR's adjoints."""

    def __init__(self, window: str, n_mag: int = None):
        super().__init__(window, "direct", n_mag)

    def forward(self, ex, log_mag):
        hop = hop_of(log_mag) // hop_of(ex)
        lm, x = plain(log_mag), plain(ex)
        assert x.ndim == 2 and lm.ndim == 3
        B, T = x.shape
        Fr = lm.shape[1]
        kernel = self.windowing(self.get_zero_phase_fir(lm))  # [B, F, K] frame rate (torch / cuFFT)
        K = kernel.shape[-1]
        n_out = min(T, (Fr - 1) * hop + 1)
        n_blk = -(-n_out // hop)
        # the block FIR pads (K-1)//2 on both sides and emits whole blocks: extend the input with zeros (the
        # reference's right padding) so that blocks 0 .. n_blk-1 exist
        p = (K - 1) // 2
        need = (n_blk - 1) * hop + (K + hop - 1) - 2 * p
        xe = F.pad(x, (0, max(need - T, 0)))
        k0 = kernel[:, :n_blk]
        k1 = kernel[:, torch.arange(1, n_blk + 1, device=kernel.device).clamp(max=Fr - 1)]
        y0 = G.ltv_fir_blocks(xe, k0.contiguous(), hop)[:, :n_out]
        y1 = G.ltv_fir_blocks(xe, k1.contiguous(), hop)[:, :n_out]
        # ATen's upsample weights at sample t (align_corners=True): src = scale * t, l1 = src - floor(src)
        t = torch.arange(n_out, device=x.device, dtype=torch.float32)
        scale = torch.tensor((Fr - 1) / max((Fr - 1) * hop, 1), dtype=torch.float32, device=x.device)
        blk = torch.div(torch.arange(n_out, device=x.device), hop, rounding_mode="floor").to(torch.float32)
        l1 = (scale * t - blk).clamp(0.0, 1.0)
        return like(ex, (1.0 - l1) * y0 + l1 * y1, hop_of(ex))


class LTIAcousticFilter(FilterInterface):
    """Learned room response: out = ex + conv(ex delayed, kernel) with `length-1` free taps
    (parameter name `kernel`, as in the checkpoints)."""

    def __init__(self, length: int, conv_method: str = "direct"):
        super().__init__()
        if conv_method not in ("direct", "fft"):
            raise ValueError(f"Unknown conv_method: {conv_method}")
        self.kernel = nn.Parameter(torch.zeros(length - 1))
        self._padding = length - 1

    def forward(self, ex):
        return like(ex, G.room_fir(plain(ex), self.kernel), hop_of(ex))

    @property
    def impulse_response(self):
        return torch.cat([self.kernel, torch.ones(1, device=self.kernel.device)]).flip(0)


def convert2samplewise(config: dict) -> dict:
    """Rewrite a decoder config so frame-wise filters become their sample-wise twins
    (models/filters.py:793-809): GOLF-ff weights evaluated as GOLF-ss ("GOLF-fs")."""
    for key, value in list(config.items()):
        if key == "class_path" and ".LTVMinimumPhaseFilter" in value and not value.endswith("Precise"):
            config["class_path"] = value.rsplit(".", 1)[0] + ".LTVMinimumPhaseFilterPrecise"
            for k in ("window", "window_length", "centred"):
                config.get("init_args", {}).pop(k, None)
            return config
        if key == "class_path" and value.endswith(".LTVZeroPhaseFIRFilter"):
            config["class_path"] = value + "Precise"
            config.get("init_args", {}).pop("conv_method", None)
            return config
        if isinstance(value, dict):
            config[key] = convert2samplewise(value)
    return config
