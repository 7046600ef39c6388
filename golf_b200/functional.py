"""Tensor-level entry points over the C ABI (include/golf_b200.h).

Functional twins of what the reference's hot path calls:
  sample_wise_lpc(x, a, zi)       torchlpc.sample_wise_lpc        (models/filters.py:112)
  lpc_ss(ex, gain, a, hop)        LTVMinimumPhaseFilterPrecise    (models/filters.py:99-113)
  lpc_ff(ex, gain, a, hop, win)   LTVMinimumPhaseFilter           (models/filters.py:131-184)
  biquad_ff(...)                  BatchSecondOrderLPCSynth        (models/lpc.py:94-131)
  lpc_frames(...)                 LPCSynth / BatchLPCSynth        (models/lpc.py:19-91)
  lpc_synthesis(x, gains, a)      lpc_synthesis -> lfilter        (models/lpc.py:11-16)
  logits2biquads / logits2lpc     get_logits2biquads, biquads2lpc (models/utils.py:444-525)
  lpc_inverse(y, a, hop)          reverse()/fir_filt              (models/filters.py:186-195)
  ltv_fir_blocks(ex, kernel, hop) LTVZeroPhaseFIRFilter           (models/filters.py:360-384)
  room_fir(x, k)                  LTIAcousticFilter               (models/filters.py:443-450)
  glottal_osc(...)                IndexedGlottalFlowTable         (models/synth.py:213-263)
  wavetable_read(...)             GlottalFlowTable.generate       (models/synth.py:124-177)

Every function requires CUDA float32 tensors and raises otherwise -- there is no
CPU path in this package (the CPU restatement lives in oracle/, test-only).
Kernels are enqueued on torch's current stream.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from ._lib import GolfError, check


def _cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise GolfError(f"{name}: expected a CUDA tensor (golf_b200 has no CPU path), got {getattr(t, 'device', type(t))}")
    if t.dtype != torch.float32:
        raise GolfError(f"{name}: expected float32, got {t.dtype}")
    t = t.detach()
    if type(t) is not torch.Tensor:  # AudioTensor and friends: plain view of the same storage
        t = t.as_subclass(torch.Tensor)
    return t.contiguous()


def _rows(t: torch.Tensor, name: str) -> torch.Tensor:
    """[B,T] with unit inner stride; the row stride may exceed T (views are fine)."""
    t = _cuda_f32_view(t, name)
    return t if t.stride(1) == 1 and t.stride(0) >= t.shape[1] else t.contiguous()


def _cuda_f32_view(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise GolfError(f"{name}: expected a CUDA tensor (golf_b200 has no CPU path), got {getattr(t, 'device', type(t))}")
    if t.dtype != torch.float32:
        raise GolfError(f"{name}: expected float32, got {t.dtype}")
    t = t.detach()
    return t.as_subclass(torch.Tensor) if type(t) is not torch.Tensor else t


def _ptr(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class _on:
    """`with _on(device):` -- make `device` current for the launch; free when it already is"""

    __slots__ = ("dev", "prev")

    def __init__(self, device):
        self.dev = device.index if device.index is not None else torch.cuda.current_device()

    def __enter__(self):
        self.prev = torch.cuda.current_device()
        if self.prev != self.dev:
            torch.cuda.set_device(self.dev)

    def __exit__(self, *exc):
        if self.prev != self.dev:
            torch.cuda.set_device(self.prev)


def _workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------- GOLF-ss
def lpc_ss_length(t_ex: int, frames: int, hop: int) -> int:
    return min(int(t_ex), (int(frames) - 1) * int(hop) + 1)


def _lpc_ss_fwd(ex, gain, a, zi, hop: int, chunk: int = 0, passes: int = 15, ws=None):
    ex = _rows(ex, "ex")
    a = _cuda_f32(a, "a")
    gain = None if gain is None else _cuda_f32(gain, "gain")
    zi = None if zi is None else _cuda_f32(zi, "zi")
    B, Tex = ex.shape
    Fr, M = a.shape[1], a.shape[2]
    if a.shape[0] != B or (gain is not None and tuple(gain.shape) != (B, Fr)) or (zi is not None and tuple(zi.shape) != (B, M)):
        raise GolfError(f"lpc_ss: inconsistent shapes ex{tuple(ex.shape)} gain{None if gain is None else tuple(gain.shape)} a{tuple(a.shape)}")
    L = lpc_ss_length(Tex, Fr, hop)
    y = torch.empty(B, L, dtype=torch.float32, device=ex.device)
    lib = _lib.lib()
    nbytes = lib.golf_lpc_ss_workspace_bytes(B, L, M, hop, chunk)
    if nbytes == 0:
        raise GolfError(f"lpc_ss: unsupported configuration B={B} L={L} M={M} hop={hop} chunk={chunk}")
    if ws is None:
        ws = _workspace(nbytes, ex.device)
    elif ws.numel() < nbytes or ws.device != ex.device:
        raise GolfError("lpc_ss: the workspace handed over by lpc_ss_responses does not fit this call")
    with _on(ex.device):
        rc = lib.golf_lpc_ss_fwd_passes(_ptr(ex), ex.stride(0), _ptr(gain), _ptr(a), _ptr(zi), _ptr(y), B, L, Fr, M, hop, chunk,
                                        _ptr(ws), ws.numel(), passes, _stream())
    check(rc, "golf_lpc_ss_fwd")
    return y


def lpc_ss_responses(a, L: int, hop: int, chunk: int = 0) -> torch.Tensor:
    """First half of the two-call form (include/golf_b200.h): the chunk transition matrices of the
    filter `a` [B,F,M] over L samples, enqueued on the CURRENT stream -- they do not depend on the
    excitation, so a caller can run this on a side stream while the excitation is produced.
    Returns the workspace to hand to lpc_ss_finish (after joining the streams)."""
    a = _cuda_f32(a, "a")
    B, Fr, M = a.shape
    lib = _lib.lib()
    nbytes = lib.golf_lpc_ss_workspace_bytes(B, L, M, hop, chunk)
    if nbytes == 0:
        raise GolfError(f"lpc_ss: unsupported configuration B={B} L={L} M={M} hop={hop} chunk={chunk}")
    ws = _workspace(nbytes, a.device)
    with _on(a.device):
        rc = lib.golf_lpc_ss_fwd_passes(0, 0, 0, _ptr(a), 0, 0, B, L, Fr, M, hop, chunk, _ptr(ws), ws.numel(), 1, _stream())
    check(rc, "golf_lpc_ss_fwd(responses)")
    return ws


def lpc_ss_finish(ex, gain, a, hop: int, ws: torch.Tensor, zi=None, chunk: int = 0, refine: bool = True) -> torch.Tensor:
    """Second half: zero-state responses (a solve from rest), stitch, solve, refinement.  Same
    result, bit for bit, as lpc_ss(ex, gain, a, hop).  Inference path (no autograd)."""
    return _lpc_ss_fwd(ex, gain, a, zi, hop, chunk, 16 | 2 | 4 | (8 if refine else 0), ws=ws)


def _lpc_ss_bwd(gy, y, ex, gain, a, zi, hop: int, need, chunk: int = 0, refine: bool = True):
    """need = (ex, gain, a, zi) booleans -> gradients (None where not needed)."""
    gy = _cuda_f32(gy, "gy")
    y = _cuda_f32(y, "y")
    ex = _rows(ex, "ex")
    a = _cuda_f32(a, "a")
    gain = None if gain is None else _cuda_f32(gain, "gain")
    zi = None if zi is None else _cuda_f32(zi, "zi")
    B, L = y.shape
    Fr, M = a.shape[1], a.shape[2]
    dev = y.device
    d_ex = torch.empty(B, L, dtype=torch.float32, device=dev) if need[0] else None
    d_gain = torch.empty(B, Fr, dtype=torch.float32, device=dev) if (need[1] and gain is not None) else None
    d_a = torch.empty(B, Fr, M, dtype=torch.float32, device=dev) if need[2] else None
    d_zi = torch.empty(B, M, dtype=torch.float32, device=dev) if (need[3] and zi is not None) else None
    lib = _lib.lib()
    ws = _workspace(lib.golf_lpc_ss_bwd_workspace_bytes(B, L, M, hop, chunk), dev)
    with _on(dev):
        rc = lib.golf_lpc_ss_bwd(_ptr(gy), _ptr(y), _ptr(ex), ex.stride(0), _ptr(gain), _ptr(a), _ptr(zi), _ptr(d_ex),
                                 _ptr(d_gain), _ptr(d_a), _ptr(d_zi), B, L, Fr, M, hop, chunk, 1 if refine else 0, _ptr(ws), ws.numel(), _stream())
    check(rc, "golf_lpc_ss_bwd")
    return d_ex, d_gain, d_a, d_zi


class _LpcSS(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ex, gain, a, zi, hop, chunk, refine):
        y = _lpc_ss_fwd(ex, gain, a, zi, hop, chunk, 15 if refine else 7)
        ctx.save_for_backward(ex, gain, a, zi, y)
        ctx.hop, ctx.chunk, ctx.refine = hop, chunk, refine
        return y

    @staticmethod
    def backward(ctx, gy):
        ex, gain, a, zi, y = ctx.saved_tensors
        need = ctx.needs_input_grad[:4]
        d_ex, d_gain, d_a, d_zi = _lpc_ss_bwd(gy, y, ex, gain, a, zi, ctx.hop, need, ctx.chunk, ctx.refine)
        if d_ex is not None and d_ex.shape[1] < ex.shape[1]:  # ex was longer than the filtered span
            d_ex = torch.nn.functional.pad(d_ex, (0, ex.shape[1] - d_ex.shape[1]))
        return d_ex, d_gain, d_a, d_zi, None, None, None


def lpc_ss(ex, gain, a, hop: int, zi=None, chunk: int = 0, refine: bool = True) -> torch.Tensor:
    """y[t] = ex[t]*up(gain)[t] - sum_i up(a)[t,i] y[t-1-i]; ex [B,T], gain [B,F], a [B,F,M] at
    `hop`; returns [B, min(T,(F-1)*hop+1)].  Differentiable in ex, gain, a, zi.
    refine=False skips the chunk-boundary refinement round (faster; up to ~10x less accurate
    on high-gain filters)."""
    return _LpcSS.apply(ex, gain, a, zi, int(hop), int(chunk), bool(refine))


def _lpc_ss_room_fwd(ex, gain, a, zi, room_k, hop: int, chunk: int = 0, refine: bool = True, keep_y: bool = False):
    ex = _rows(ex, "ex")
    a = _cuda_f32(a, "a")
    gain = None if gain is None else _cuda_f32(gain, "gain")
    zi = None if zi is None else _cuda_f32(zi, "zi")
    room_k = _cuda_f32(room_k, "room kernel")
    B, Tex = ex.shape
    Fr, M = a.shape[1], a.shape[2]
    if a.shape[0] != B or (gain is not None and tuple(gain.shape) != (B, Fr)) or (zi is not None and tuple(zi.shape) != (B, M)):
        raise GolfError(f"lpc_ss_room: inconsistent shapes ex{tuple(ex.shape)} gain{None if gain is None else tuple(gain.shape)} a{tuple(a.shape)}")
    L = lpc_ss_length(Tex, Fr, hop)
    out = torch.empty(B, L, dtype=torch.float32, device=ex.device)
    y = torch.empty(B, L, dtype=torch.float32, device=ex.device) if keep_y else None
    lib = _lib.lib()
    nbytes = lib.golf_lpc_ss_room_workspace_bytes(B, L, M, hop, chunk)
    if nbytes == 0:
        raise GolfError(f"lpc_ss_room: unsupported configuration B={B} L={L} M={M} hop={hop} chunk={chunk}")
    ws = _workspace(nbytes, ex.device)
    with _on(ex.device):
        rc = lib.golf_lpc_ss_room_fwd(_ptr(ex), ex.stride(0), _ptr(gain), _ptr(a), _ptr(zi), _ptr(room_k), room_k.numel(), _ptr(y),
                                      _ptr(out), B, L, Fr, M, hop, chunk, 1 if refine else 0, _ptr(ws), ws.numel(), _stream())
    check(rc, "golf_lpc_ss_room_fwd")
    return out, y


class _LpcSSRoom(torch.autograd.Function):
    """room_fir(lpc_ss(ex, gain, a), k) in one pass; the adjoint chains the two stages' adjoints on the saved y"""

    @staticmethod
    def forward(ctx, ex, gain, a, room_k, hop, refine):
        need = any(ctx.needs_input_grad[:4])
        out, y = _lpc_ss_room_fwd(ex, gain, a, None, room_k, hop, 0, refine, keep_y=need)
        if need:
            ctx.save_for_backward(ex, gain, a, room_k, y)
        ctx.hop, ctx.refine = hop, refine
        return out

    @staticmethod
    def backward(ctx, gout):
        ex, gain, a, room_k, y = ctx.saved_tensors
        gout = _cuda_f32(gout, "gout")
        kc = _cuda_f32(room_k, "room kernel")
        B, L = y.shape
        gy = torch.empty_like(y)
        d_k = torch.empty_like(kc) if ctx.needs_input_grad[3] else None
        with _on(gout.device):
            rc = _lib.lib().golf_room_fir_bwd(_ptr(gout), _ptr(y), _ptr(kc), _ptr(gy), _ptr(d_k), B, L, kc.numel(), _stream())
        check(rc, "golf_room_fir_bwd")
        d_ex = d_gain = d_a = None
        if any(ctx.needs_input_grad[:3]):
            d_ex, d_gain, d_a, _ = _lpc_ss_bwd(gy, y, ex, gain, a, None, ctx.hop, tuple(ctx.needs_input_grad[:3]) + (False,), 0, ctx.refine)
            if d_ex is not None and d_ex.shape[1] < ex.shape[1]:
                d_ex = torch.nn.functional.pad(d_ex, (0, ex.shape[1] - d_ex.shape[1]))
        return d_ex, d_gain, d_a, d_k, None, None


def lpc_ss_room(ex, gain, a, room_k, hop: int, refine: bool = True) -> torch.Tensor:
    """GOLF-ss end filter followed by the learned room FIR (models/sf.py:64) -- the filter's stitch / solve and the
    128-tap FIR share one launch (golf_lpc_ss_room_fwd).  Differentiable in ex, gain, a and the room taps."""
    return _LpcSSRoom.apply(ex, gain, a, room_k, int(hop), bool(refine))


def sample_wise_lpc(x, a, zi=None) -> torch.Tensor:
    """torchlpc.sample_wise_lpc: x [B,T], a [B,T,M] sample-rate coefficients, zi [B,M]."""
    if x.ndim != 2 or a.ndim != 3 or a.shape[:2] != x.shape:
        raise GolfError(f"sample_wise_lpc: x{tuple(x.shape)} a{tuple(a.shape)}")
    return _LpcSS.apply(x, None, a, zi, 1, 0, True)


# ------------------------------------------------------------------------- GOLF-ff
def _lpc_ff_fwd(ex, gain, a, window, hop: int) -> torch.Tensor:
    ex = _rows(ex, "ex")
    gain, a, window = _cuda_f32(gain, "gain"), _cuda_f32(a, "a"), _cuda_f32(window, "window")
    B, Tex = ex.shape
    Fr, M = a.shape[1], a.shape[2]
    win = window.numel()
    Le = lpc_ss_length(Tex, Fr, hop)
    n_frames = (Le + 2 * (win // 2) - win) // hop + 1
    if n_frames > Fr:
        raise AssertionError(f"{n_frames} frames but only {Fr} control frames")  # filters.py:157
    out_len = (n_frames - 1) * hop + win - 2 * (win // 2)
    y = torch.empty(B, out_len, dtype=torch.float32, device=ex.device)
    with _on(ex.device):
        rc = _lib.lib().golf_lpc_ff_fwd(_ptr(ex), ex.stride(0), _ptr(gain), _ptr(a), _ptr(window), _ptr(y), B, Tex, Fr, M,
                                        hop, win, _stream())
    check(rc, "golf_lpc_ff_fwd")
    return y


class _LpcFF(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ex, gain, a, window, hop):
        y = _lpc_ff_fwd(ex, gain, a, window, hop)
        ctx.save_for_backward(ex, gain, a, window)
        ctx.hop = hop
        return y

    @staticmethod
    def backward(ctx, gy):
        ex, gain, a, window = ctx.saved_tensors
        gy = _cuda_f32(gy, "gy")
        exr = _rows(ex, "ex")
        gain_c, a_c, win_c = _cuda_f32(gain, "gain"), _cuda_f32(a, "a"), _cuda_f32(window, "window")
        B, Tex = exr.shape
        Fr, M = a_c.shape[1], a_c.shape[2]
        win, hop = win_c.numel(), ctx.hop
        need = ctx.needs_input_grad
        dev = gy.device
        d_ex = torch.empty(B, Tex, dtype=torch.float32, device=dev) if need[0] else None
        d_gain = torch.empty(B, Fr, dtype=torch.float32, device=dev) if need[1] else None
        d_a = torch.empty(B, Fr, M, dtype=torch.float32, device=dev) if need[2] else None
        lib = _lib.lib()
        ws = _workspace(lib.golf_lpc_ff_bwd_workspace_bytes(B, Tex, Fr, hop, win), dev)
        with _on(dev):
            rc = lib.golf_lpc_ff_bwd(_ptr(gy), _ptr(exr), exr.stride(0), _ptr(gain_c), _ptr(a_c), _ptr(win_c), _ptr(d_ex),
                                     Tex, _ptr(d_gain), _ptr(d_a), B, Tex, Fr, M, hop, win, _ptr(ws), ws.numel(), _stream())
        check(rc, "golf_lpc_ff_bwd")
        return d_ex, d_gain, d_a, None, None


def lpc_ff(ex, gain, a, window, hop: int) -> torch.Tensor:
    """Frame-wise filter + windowed OLA; differentiable in ex, gain, a."""
    return _LpcFF.apply(ex, gain, a, window, int(hop))


def _biquad_ff_fwd(ex, gain, biquads, window, hop: int) -> torch.Tensor:
    ex = _rows(ex, "ex")
    gain, biquads, window = _cuda_f32(gain, "gain"), _cuda_f32(biquads, "biquads"), _cuda_f32(window, "window")
    B, Tex = ex.shape
    Fr, K = biquads.shape[1], biquads.shape[2]
    win = window.numel()
    pad = (win - hop) // 2
    n_frames = (Tex + 2 * pad - win) // hop + 1
    if n_frames > Fr:
        raise AssertionError(f"{n_frames} frames but only {Fr} control frames")  # lpc.py:102-104
    y = torch.empty(B, (n_frames - 1) * hop + win - 2 * pad, dtype=torch.float32, device=ex.device)
    with _on(ex.device):
        rc = _lib.lib().golf_biquad_cascade_fwd(_ptr(ex), ex.stride(0), _ptr(gain), _ptr(biquads), _ptr(window), _ptr(y), B, Tex,
                                                Fr, K, hop, win, _stream())
    check(rc, "golf_biquad_cascade_fwd")
    return y


class _BiquadFF(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ex, gain, biquads, window, hop):
        y = _biquad_ff_fwd(ex, gain, biquads, window, hop)
        ctx.save_for_backward(ex, gain, biquads, window)
        ctx.hop = hop
        return y

    @staticmethod
    def backward(ctx, gy):
        ex, gain, biquads, window = ctx.saved_tensors
        gy = _cuda_f32(gy, "gy")
        exr = _rows(ex, "ex")
        gain_c, bq_c, win_c = _cuda_f32(gain, "gain"), _cuda_f32(biquads, "biquads"), _cuda_f32(window, "window")
        B, Tex = exr.shape
        Fr, K = bq_c.shape[1], bq_c.shape[2]
        win, hop = win_c.numel(), ctx.hop
        need, dev = ctx.needs_input_grad, gy.device
        d_ex = torch.empty(B, Tex, dtype=torch.float32, device=dev)
        d_gain = torch.empty(B, Fr, dtype=torch.float32, device=dev) if need[1] else None
        d_bq = torch.empty(B, Fr, K, 3, dtype=torch.float32, device=dev) if need[2] else None
        lib = _lib.lib()
        ws = _workspace(lib.golf_biquad_cascade_bwd_workspace_bytes(B, Tex, Fr, K, hop, win), dev)
        with _on(dev):
            rc = lib.golf_biquad_cascade_bwd(_ptr(gy), _ptr(exr), exr.stride(0), _ptr(gain_c), _ptr(bq_c), _ptr(win_c), _ptr(d_ex),
                                             _ptr(d_gain), _ptr(d_bq), B, Tex, Fr, K, hop, win, _ptr(ws), ws.numel(), _stream())
        check(rc, "golf_biquad_cascade_bwd")
        return (d_ex if need[0] else None), d_gain, d_bq, None, None


def biquad_ff(ex, gain, biquads, window, hop: int) -> torch.Tensor:
    """BatchSecondOrderLPCSynth.forward (models/lpc.py:94-131): ex [B,T], gain [B,F], biquads [B,F,K,3], window [win].
    Differentiable in ex, gain and the sections (golf_biquad_cascade_fwd / _bwd)."""
    return _BiquadFF.apply(ex, gain, biquads, window, int(hop))


# ------------------------------------------------- LPCSynth / BatchLPCSynth (models/lpc.py:19-91)
def _lpc_frames_fwd(ex, gain, a, window, hop: int) -> torch.Tensor:
    ex = _rows(ex, "ex")
    gain, a, window = _cuda_f32(gain, "gain"), _cuda_f32(a, "a"), _cuda_f32(window, "window")
    B, Tex = ex.shape
    Fr, M = a.shape[1], a.shape[2]
    win = window.numel()
    pad = (win - hop) // 2
    n_frames = (Tex + 2 * pad - win) // hop + 1
    if n_frames > Fr:
        raise AssertionError(f"{n_frames} frames but only {Fr} control frames")  # lpc.py:71
    lib = _lib.lib()
    out_len = lib.golf_lpc_frames_out_length(Tex, Fr, hop, win)
    if out_len <= 0:
        raise GolfError(f"lpc_frames: unsupported geometry T={Tex} F={Fr} hop={hop} win={win}")
    y = torch.empty(B, out_len, dtype=torch.float32, device=ex.device)
    with _on(ex.device):
        rc = lib.golf_lpc_frames_fwd(_ptr(ex), ex.stride(0), _ptr(gain), _ptr(a), _ptr(window), _ptr(y), B, Tex, Fr, M, hop, win,
                                     _stream())
    check(rc, "golf_lpc_frames_fwd")
    return y


class _LpcFrames(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ex, gain, a, window, hop):
        y = _lpc_frames_fwd(ex, gain, a, window, hop)
        ctx.save_for_backward(ex, gain, a, window)
        ctx.hop = hop
        return y

    @staticmethod
    def backward(ctx, gy):
        ex, gain, a, window = ctx.saved_tensors
        gy = _cuda_f32(gy, "gy")
        exr = _rows(ex, "ex")
        gain_c, a_c, win_c = _cuda_f32(gain, "gain"), _cuda_f32(a, "a"), _cuda_f32(window, "window")
        B, Tex = exr.shape
        Fr, M = a_c.shape[1], a_c.shape[2]
        win, hop = win_c.numel(), ctx.hop
        need, dev = ctx.needs_input_grad, gy.device
        d_ex = torch.empty(B, Tex, dtype=torch.float32, device=dev)
        d_gain = torch.empty(B, Fr, dtype=torch.float32, device=dev) if need[1] else None
        d_a = torch.empty(B, Fr, M, dtype=torch.float32, device=dev) if need[2] else None
        lib = _lib.lib()
        ws = _workspace(lib.golf_lpc_frames_bwd_workspace_bytes(B, Tex, Fr, hop, win), dev)
        with _on(dev):
            rc = lib.golf_lpc_frames_bwd(_ptr(gy), _ptr(exr), exr.stride(0), _ptr(gain_c), _ptr(a_c), _ptr(win_c), _ptr(d_ex),
                                         _ptr(d_gain), _ptr(d_a), B, Tex, Fr, M, hop, win, _ptr(ws), ws.numel(), _stream())
        check(rc, "golf_lpc_frames_bwd")
        return (d_ex if need[0] else None), d_gain, d_a, None, None


def lpc_frames(ex, gain, a, window, hop: int) -> torch.Tensor:
    """BatchLPCSynth.forward (models/lpc.py:62-91): per-frame gain and LTI all-pole, zero-pad (win-hop)/2, window OLA."""
    return _LpcFrames.apply(ex, gain, a, window, int(hop))


# ------------------------------------------------- lpc_synthesis (models/lpc.py:11-16), lfilter-shaped
class _LfilterAllpole(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gains, a):
        xr = _rows(x, "source")
        a_c = _cuda_f32(a, "a")
        g_c = _cuda_f32(gains, "gains") if gains is not None else None
        C, N = xr.shape
        if a_c.ndim != 2 or a_c.shape[0] != C or (g_c is not None and g_c.numel() != C):
            raise GolfError(f"lpc_synthesis: source{tuple(xr.shape)} gains{None if g_c is None else tuple(g_c.shape)} a{tuple(a_c.shape)}")
        y = torch.empty(C, N, dtype=torch.float32, device=xr.device)
        with _on(xr.device):
            rc = _lib.lib().golf_lfilter_allpole_fwd(_ptr(xr), xr.stride(0), _ptr(g_c), _ptr(a_c), _ptr(y), C, N, a_c.shape[1], _stream())
        check(rc, "golf_lfilter_allpole_fwd")
        ctx.save_for_backward(x, gains, a, y)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, gains, a, y = ctx.saved_tensors
        gy = _cuda_f32(gy, "gy")
        xr = _rows(x, "source")
        a_c = _cuda_f32(a, "a")
        g_c = _cuda_f32(gains, "gains") if gains is not None else None
        C, N = xr.shape
        M = a_c.shape[1]
        need, dev = ctx.needs_input_grad, gy.device
        d_x = torch.empty(C, N, dtype=torch.float32, device=dev) if need[0] else None
        d_g = torch.empty(C, dtype=torch.float32, device=dev) if (need[1] and gains is not None) else None
        d_a = torch.empty(C, M, dtype=torch.float32, device=dev) if need[2] else None
        lib = _lib.lib()
        ws = _workspace(lib.golf_lfilter_allpole_bwd_workspace_bytes(C, N), dev)
        with _on(dev):
            rc = lib.golf_lfilter_allpole_bwd(_ptr(gy), _ptr(xr), xr.stride(0), _ptr(y), _ptr(g_c), _ptr(a_c), _ptr(d_x), _ptr(d_g),
                                              _ptr(d_a), C, N, M, _ptr(ws), ws.numel(), _stream())
        check(rc, "golf_lfilter_allpole_bwd")
        if d_g is not None:
            d_g = d_g.view(gains.shape)
        return d_x, d_g, d_a


def lpc_synthesis(source, gains, a) -> torch.Tensor:
    """models/lpc.py:11-16: `lfilter(source, [1, a], [gains, 0, ...], clamp=False)` for source [C,N], gains [C], a [C,M]:
    y[c,n] = gains[c] source[c,n] - sum_i a[c,i] y[c,n-1-i] from zero state.  Differentiable in all three."""
    if source.ndim != 2:
        raise GolfError(f"lpc_synthesis: source must be [C,N], got {tuple(source.shape)}")
    return _LfilterAllpole.apply(source, gains, a)


# ------------------------------------------------- biquad parameterisations (models/utils.py:444-525)
_BIQUAD_REPS = {"coef": 0, "conj": 1, "real": 2}


class _BiquadParams(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, rep, rho, want_bq, want_a):
        lg = _cuda_f32(logits, "logits")
        K = lg.shape[-2]
        N = lg.numel() // (2 * K)
        lead = tuple(lg.shape[:-2])
        bq = torch.empty(*lead, K, 3, dtype=torch.float32, device=lg.device) if want_bq else None
        a = torch.empty(*lead, 2 * K, dtype=torch.float32, device=lg.device) if want_a else None
        with _on(lg.device):
            rc = _lib.lib().golf_biquad_params_fwd(_ptr(lg), _ptr(bq), _ptr(a), N, K, rep, rho, _stream())
        check(rc, "golf_biquad_params_fwd")
        ctx.save_for_backward(logits)
        ctx.cfg = (rep, rho, N, K)
        ctx.set_materialize_grads(False)
        outs = tuple(t for t in (bq, a) if t is not None)
        ctx.slots = (want_bq, want_a)
        return outs if len(outs) > 1 else outs[0]

    @staticmethod
    def backward(ctx, *grads):
        (logits,) = ctx.saved_tensors
        rep, rho, N, K = ctx.cfg
        grads = list(grads)
        d_bq = grads.pop(0) if ctx.slots[0] else None
        d_a = grads.pop(0) if ctx.slots[1] else None
        if d_bq is None and d_a is None:
            return None, None, None, None, None
        lg = _cuda_f32(logits, "logits")
        d_bq = _cuda_f32(d_bq, "d_biquads") if d_bq is not None else None
        d_a = _cuda_f32(d_a, "d_a") if d_a is not None else None
        d_lg = torch.empty_like(lg)
        with _on(lg.device):
            rc = _lib.lib().golf_biquad_params_bwd(_ptr(lg), _ptr(d_bq), _ptr(d_a), _ptr(d_lg), N, K, rep, rho, _stream())
        check(rc, "golf_biquad_params_bwd")
        return d_lg, None, None, None, None


def logits2biquads(logits, rep_type: str = "coef", max_abs_pole: float = 0.99) -> torch.Tensor:
    """get_logits2biquads(rep_type, max_abs_pole)(logits) (models/utils.py:487-525): [...,K,2] -> sections [...,K,3]."""
    if rep_type not in _BIQUAD_REPS:
        raise ValueError(f"Unknown rep_type: {rep_type}, expected coef, conj or real")
    return _BiquadParams.apply(logits, _BIQUAD_REPS[rep_type], float(max_abs_pole), True, False)


def logits2lpc(logits, rep_type: str = "coef", max_abs_pole: float = 0.99) -> torch.Tensor:
    """biquads2lpc(get_logits2biquads(...)(logits)) as composed at models/filters.py:73-78: [...,K,2] -> a [...,2K]
    (one launch each way; the sections never reach HBM)."""
    if rep_type not in _BIQUAD_REPS:
        raise ValueError(f"Unknown rep_type: {rep_type}, expected coef, conj or real")
    return _BiquadParams.apply(logits, _BIQUAD_REPS[rep_type], float(max_abs_pole), False, True)


def _lpc_inverse_fwd(y, a, hop: int) -> torch.Tensor:
    y = _rows(y, "y")
    a = _cuda_f32(a, "a")
    B, T = y.shape
    Fr, M = a.shape[1], a.shape[2]
    L = lpc_ss_length(T, Fr, hop)
    r = torch.empty(B, L, dtype=torch.float32, device=y.device)
    with _on(y.device):
        rc = _lib.lib().golf_lpc_inverse_fwd(_ptr(y), y.stride(0), _ptr(a), _ptr(r), B, L, Fr, M, hop, _stream())
    check(rc, "golf_lpc_inverse_fwd")
    return r


class _LpcInverse(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, a, hop):
        r = _lpc_inverse_fwd(y, a, hop)
        ctx.save_for_backward(y, a)
        ctx.hop = hop
        return r

    @staticmethod
    def backward(ctx, g):
        y, a = ctx.saved_tensors
        g = _cuda_f32(g, "g")
        yr, ac = _rows(y, "y"), _cuda_f32(a, "a")
        B, T = yr.shape
        Fr, M = ac.shape[1], ac.shape[2]
        L = g.shape[1]
        d_yl = torch.empty(B, L, dtype=torch.float32, device=g.device) if ctx.needs_input_grad[0] else None
        d_a = torch.empty_like(ac) if ctx.needs_input_grad[1] else None
        with _on(g.device):
            rc = _lib.lib().golf_lpc_inverse_bwd(_ptr(g), _ptr(yr), yr.stride(0), _ptr(ac), _ptr(d_yl), _ptr(d_a), B, L, Fr, M,
                                                 ctx.hop, _stream())
        check(rc, "golf_lpc_inverse_bwd")
        d_y = None
        if d_yl is not None:  # samples of y beyond the output length never reach it
            d_y = torch.nn.functional.pad(d_yl, (0, T - L)) if T > L else d_yl
        return d_y, d_a, None


def lpc_inverse(y, a, hop: int) -> torch.Tensor:
    """Inverse (analysis) filter r[t] = y[t] + sum_i a_up[t,i] y[t-1-i] (models/filters.py:186-195 +
    fir_filt, models/utils.py:433-441) on frame-rate coefficients; differentiable in y and a."""
    return _LpcInverse.apply(y, a, int(hop))


# ---------------------------------------------------------------------- FIR stages
def _fir_blocks_count(T: int, Fr: int, K: int, hop: int) -> int:
    p = (K - 1) // 2
    return min((T + 2 * p - (K + hop - 1)) // hop + 1, Fr)


def _ltv_fir_fwd(ex, kernel, hop, add=None, window=None):
    ex = _rows(ex, "ex")
    kernel = _cuda_f32(kernel, "kernel")
    window = None if window is None else _cuda_f32(window, "window")
    B, T = ex.shape
    Fr, K = kernel.shape[1], kernel.shape[2]
    n_blocks = _fir_blocks_count(T, Fr, K, hop)
    if add is not None:
        add = _rows(add, "add")
        if add.shape[1] < n_blocks * hop:
            raise GolfError("ltv_fir_blocks: `add` shorter than the output")
    y = torch.empty(B, n_blocks * hop, dtype=torch.float32, device=ex.device)
    with _on(ex.device):
        rc = _lib.lib().golf_noise_fir_fwd(_ptr(ex), ex.stride(0), _ptr(kernel), _ptr(window), _ptr(add),
                                           0 if add is None else add.stride(0), _ptr(y), B, T, Fr, K, hop, _stream())
    check(rc, "golf_noise_fir_fwd")
    return y


class _LtvFir(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ex, kernel, add, hop):
        y = _ltv_fir_fwd(ex, kernel, hop, add)
        ctx.save_for_backward(ex, kernel)
        ctx.hop, ctx.has_add = hop, add is not None
        ctx.add_len = 0 if add is None else add.shape[1]
        return y

    @staticmethod
    def backward(ctx, gy):
        ex, kernel = ctx.saved_tensors
        gy = _cuda_f32(gy, "gy")
        exr = _rows(ex, "ex")
        kern = _cuda_f32(kernel, "kernel")
        B, T = exr.shape
        Fr, K = kern.shape[1], kern.shape[2]
        d_ex = torch.empty(B, T, dtype=torch.float32, device=gy.device) if ctx.needs_input_grad[0] else None
        d_k = torch.empty_like(kern) if ctx.needs_input_grad[1] else None
        with _on(gy.device):
            rc = _lib.lib().golf_noise_fir_bwd(_ptr(gy), _ptr(exr), exr.stride(0), _ptr(kern), _ptr(d_ex), _ptr(d_k), B, T, Fr, K,
                                               ctx.hop, _stream())
        check(rc, "golf_noise_fir_bwd")
        d_add = None
        if ctx.has_add and ctx.needs_input_grad[2]:
            d_add = torch.nn.functional.pad(gy, (0, ctx.add_len - gy.shape[1])) if ctx.add_len > gy.shape[1] else gy
        return d_ex, d_k, d_add, None


def ltv_fir_blocks(ex, kernel, hop: int, add=None, window=None) -> torch.Tensor:
    """Block FIR with per-frame kernels [B,F,K]; optional fused `add + result`.  With `window`
    ([K]) the kernel argument is the raw irfft output and fftshift + windowing are fused into the
    tap staging (inference path, no autograd)."""
    if window is not None:
        return _ltv_fir_fwd(ex, kernel, int(hop), add, window)
    return _LtvFir.apply(ex, kernel, add, int(hop))


def noise_fir_design_supported(n_mag: int, hop: int) -> bool:
    return bool(_lib.lib().golf_noise_fir_design_supported(int(n_mag), int(hop)))


def new_rng_state(device, seed: Optional[int] = None) -> torch.Tensor:
    """{seed, offset} for the in-kernel noise generator: an int64[2] device tensor.  seed=None draws it from torch's
    (CPU) generator, so torch.manual_seed governs it."""
    if seed is None:
        seed = int(torch.randint(0, 2**62, (1,), dtype=torch.int64).item())
    return torch.tensor([seed, 0], dtype=torch.int64, device=device)


def philox_normal(B: int, T: int, rng_state: torch.Tensor) -> torch.Tensor:
    """the N(0,1) draw the fused noise kernels make for (rng_state), written to memory (tests, debugging)"""
    out = torch.empty(B, T, dtype=torch.float32, device=rng_state.device)
    with _on(out.device):
        rc = _lib.lib().golf_philox_normal(_ptr(out), B, T, _ptr(rng_state), _stream())
    check(rc, "golf_philox_normal")
    return out


def rng_advance(rng_state: torch.Tensor) -> None:
    with _on(rng_state.device):
        rc = _lib.lib().golf_rng_advance(_ptr(rng_state), _stream())
    check(rc, "golf_rng_advance")


def noise_fir_design(ex, log_mag, window, hop: int, add=None, rng_state=None) -> torch.Tensor:
    """LTVZeroPhaseFIRFilter.forward with the taps designed inside the FIR kernel (golf_noise_fir_design_fwd): ex [B,T]
    white noise (or None + rng_state: drawn in the kernel; give T through `add` or log_mag), log_mag [B,F,256],
    window [510] (unscaled), optional fused `add + result`.  Inference path (no autograd)."""
    log_mag = _cuda_f32(log_mag, "log_mag")
    window = _cuda_f32(window, "window")
    B, Fr, n_mag = log_mag.shape
    K = 2 * (n_mag - 1)
    if window.numel() != K:
        raise GolfError(f"noise_fir_design: window has {window.numel()} taps, expected {K}")
    if add is not None:
        add = _rows(add, "add")
    if ex is not None:
        ex = _rows(ex, "ex")
        T = ex.shape[1]
    elif rng_state is not None and add is not None:
        T = add.shape[1]
    else:
        raise GolfError("noise_fir_design: pass the noise (ex) or rng_state together with `add` (which fixes the length)")
    n_blocks = _fir_blocks_count(T, Fr, K, hop)
    if add is not None and add.shape[1] < n_blocks * hop:
        raise GolfError("noise_fir_design: `add` shorter than the output")
    y = torch.empty(B, n_blocks * hop, dtype=torch.float32, device=log_mag.device)
    with _on(log_mag.device):
        rc = _lib.lib().golf_noise_fir_design_fwd(_ptr(ex), 0 if ex is None else ex.stride(0), 0 if ex is not None else _ptr(rng_state),
                                                  _ptr(log_mag), _ptr(window), _ptr(add), 0 if add is None else add.stride(0), _ptr(y),
                                                  B, T, Fr, n_mag, hop, _stream())
    check(rc, "golf_noise_fir_design_fwd")
    return y


def synth_fused(phase, phase_hop: int, w, w_hop: int, table, dec_kernel, oversampling: int, equal_energy: bool, accumulate: str,
                log_mag, fir_window, gain, a, hop: int, room_k=None, noise=None, rng_state=None, refine: bool = True,
                workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The whole GOLF-ss decoder pass through golf_synth_fused_fwd (five launches, no library kernel): oscillator ->
    + FIR-filtered noise (taps designed in-kernel) -> sample-wise LPC filter -> room FIR.  noise [B,>=T_osc] is the
    white-noise draw; noise=None uses the in-kernel generator with rng_state (advanced by the call).  Inference only."""
    phase, w, table = _cuda_f32(phase, "phase"), _cuda_f32(w, "w"), _cuda_f32(table, "table")
    log_mag, fir_window = _cuda_f32(log_mag, "log_mag"), _cuda_f32(fir_window, "window")
    gain, a = _cuda_f32(gain, "gain"), _cuda_f32(a, "a")
    dk = None if dec_kernel is None else _cuda_f32(dec_kernel, "dec_kernel")
    rk = None if room_k is None else _cuda_f32(room_k, "room kernel")
    B, Np = phase.shape
    Fw = w.shape[1]
    n_tab, P = table.shape
    Fr, M = a.shape[1], a.shape[2]
    n_mag = log_mag.shape[2]
    if log_mag.shape[:2] != (B, Fr) or tuple(gain.shape) != (B, Fr) or a.shape[0] != B or w.shape[0] != B:
        raise GolfError("synth_fused: inconsistent control shapes")
    if noise is None and rng_state is None:
        raise GolfError("synth_fused: pass the white-noise draw or an rng_state")
    if noise is not None:
        noise = _rows(noise, "noise")
    zeros = 0 if dk is None else (dk.numel() - 1) // (2 * oversampling)
    lib = _lib.lib()
    L = lib.golf_synth_fused_out_length(Np, phase_hop, oversampling, Fr, hop, n_mag)
    nbytes = lib.golf_synth_fused_workspace_bytes(B, Np, phase_hop, Fw, P, oversampling, Fr, M, hop, n_mag)
    if L <= 0 or nbytes == 0:
        raise GolfError("synth_fused: unsupported configuration")
    ws = workspace if workspace is not None else _workspace(nbytes, phase.device)
    if ws.numel() < nbytes or ws.data_ptr() % 256:
        raise GolfError("synth_fused: workspace too small or not 256-byte aligned")
    out = torch.empty(B, L, dtype=torch.float32, device=phase.device)
    with _on(phase.device):
        rc = lib.golf_synth_fused_fwd(_ptr(phase), _ptr(w), _ptr(table), _ptr(dk), _ptr(noise), 0 if noise is None else noise.stride(0),
                                      0 if noise is not None else _ptr(rng_state), _ptr(log_mag), _ptr(fir_window), _ptr(gain), _ptr(a),
                                      _ptr(rk), 0 if rk is None else rk.numel(), _ptr(out), B, Np, phase_hop, Fw, w_hop, n_tab, P,
                                      oversampling, zeros, _OSC_MODES[accumulate], 1 if equal_energy else 0, Fr, M, hop, n_mag,
                                      1 if refine else 0, _ptr(ws), ws.numel(), _stream())
    check(rc, "golf_synth_fused_fwd")
    return out


class _RoomFir(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k):
        xc, kc = _cuda_f32(x, "x"), _cuda_f32(k, "k")
        B, T = xc.shape
        out = torch.empty_like(xc)
        with _on(xc.device):
            rc = _lib.lib().golf_room_fir_fwd(_ptr(xc), _ptr(kc), _ptr(out), B, T, kc.numel(), _stream())
        check(rc, "golf_room_fir_fwd")
        ctx.save_for_backward(xc, kc)
        return out

    @staticmethod
    def backward(ctx, gy):
        xc, kc = ctx.saved_tensors
        gy = _cuda_f32(gy, "gy")
        B, T = xc.shape
        d_x = torch.empty_like(xc) if ctx.needs_input_grad[0] else None
        d_k = torch.empty_like(kc) if ctx.needs_input_grad[1] else None
        with _on(gy.device):
            rc = _lib.lib().golf_room_fir_bwd(_ptr(gy), _ptr(xc), _ptr(kc), _ptr(d_x), _ptr(d_k), B, T, kc.numel(), _stream())
        check(rc, "golf_room_fir_bwd")
        return d_x, d_k


def room_fir(x, k) -> torch.Tensor:
    """out[t] = x[t] + sum_j k[j] x[t-len(k)+j]; differentiable in x and k."""
    return _RoomFir.apply(x, k)


# ---------------------------------------------------------------------- oscillator
_OSC_MODES = {"exact": 0, "fp64": 0, "aten_cpu": 1}


class _GlottalOsc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, phase, w, table, dec_kernel, phase_hop, w_hop, os_, equal_energy, mode, phase0=None):
        phase_c, w_c, table_c = _cuda_f32(phase, "phase"), _cuda_f32(w, "w"), _cuda_f32(table, "table")
        p0 = None
        if phase0 is not None:
            if not phase0.is_cuda:
                raise GolfError("glottal_osc: phase0 must be a CUDA tensor")
            p0 = phase0.detach().to(torch.float64).reshape(-1).contiguous()
        if p0 is not None and p0.numel() != phase_c.shape[0]:
            raise GolfError("glottal_osc: phase0 must hold one value per utterance")
        dk = None if dec_kernel is None else _cuda_f32(dec_kernel, "dec_kernel")
        B, Np = phase_c.shape
        Fw = w_c.shape[1]
        n_tab, P = table_c.shape
        zeros = 0 if dk is None else (dk.numel() - 1) // (2 * os_)
        N = (Np - 1) * phase_hop * os_ + 1
        out = torch.empty(B, (N - 1) // os_ + 1, dtype=torch.float32, device=phase_c.device)
        lib = _lib.lib()
        ws = _workspace(lib.golf_glottal_osc_workspace_bytes(B, Np, phase_hop, Fw, P, os_), phase_c.device)
        with _on(phase_c.device):
            if p0 is None:
                rc = lib.golf_glottal_osc_fwd(_ptr(phase_c), _ptr(w_c), _ptr(table_c), _ptr(dk), _ptr(out), B, Np, phase_hop, Fw,
                                              w_hop, n_tab, P, os_, zeros, mode, 1 if equal_energy else 0, _ptr(ws), ws.numel(), _stream())
            else:
                rc = lib.golf_glottal_osc_fwd_from(_ptr(phase_c), _ptr(w_c), _ptr(table_c), _ptr(dk), _ptr(out), _ptr(p0), B, Np,
                                                   phase_hop, Fw, w_hop, n_tab, P, os_, zeros, mode, 1 if equal_energy else 0,
                                                   _ptr(ws), ws.numel(), _stream())
        check(rc, "golf_glottal_osc_fwd")
        if p0 is not None and any(ctx.needs_input_grad[:3]):
            raise GolfError("glottal_osc: the adjoint does not take an initial phase (streaming is an inference path)")
        ctx.save_for_backward(phase_c, w_c, table_c, dk)
        ctx.geom = (phase_hop, w_hop, os_, zeros, equal_energy, mode)
        return out

    @staticmethod
    def backward(ctx, gout):
        if ctx.needs_input_grad[0]:
            raise GolfError("glottal_osc: the gradient w.r.t. phase is not implemented (detach f0, as the shipped configs do)")
        need_w, need_t = ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        if not (need_w or need_t):
            return (None,) * 10
        phase_c, w_c, table_c, dk = ctx.saved_tensors
        phase_hop, w_hop, os_, zeros, equal_energy, mode = ctx.geom
        gout = _cuda_f32(gout, "gout")
        B, Np = phase_c.shape
        Fw = w_c.shape[1]
        n_tab, P = table_c.shape
        d_w = torch.empty_like(w_c) if need_w else None
        d_table = torch.empty_like(table_c) if need_t else None
        lib = _lib.lib()
        ws = _workspace(lib.golf_glottal_osc_workspace_bytes(B, Np, phase_hop, Fw, P, os_), gout.device)
        with _on(gout.device):
            rc = lib.golf_glottal_osc_bwd(_ptr(gout), _ptr(phase_c), _ptr(w_c), _ptr(table_c), _ptr(dk), _ptr(d_w), _ptr(d_table),
                                          B, Np, phase_hop, Fw, w_hop, n_tab, P, os_, zeros, mode, 1 if equal_energy else 0,
                                          _ptr(ws), ws.numel(), _stream())
        check(rc, "golf_glottal_osc_bwd" + (" (table gradient: exact-phase mode, oversampling 1/2/4, power-of-two table length)" if need_t else ""))
        return None, d_w, d_table, None, None, None, None, None, None, None


def glottal_osc(phase, phase_hop: int, w, w_hop: int, table, dec_kernel=None, oversampling: int = 1,
                equal_energy: bool = False, accumulate: str = "exact", phase0=None) -> torch.Tensor:
    """IndexedGlottalFlowTable.forward; differentiable in the table-selection weight `w` and in the table (trainable tables).  phase0 [B]: running
    phase (cycles) each utterance starts from (a per-utterance constant `phase_offset`, exact-phase mode)."""
    return _GlottalOsc.apply(phase, w, table, dec_kernel, int(phase_hop), int(w_hop), int(oversampling), bool(equal_energy),
                             _OSC_MODES[accumulate], phase0)


def wavetable_read(wrapped, tables, hop_tab: int) -> torch.Tensor:
    wrapped, tables = _cuda_f32(wrapped, "wrapped_phase"), _cuda_f32(tables, "tables")
    B, N = wrapped.shape
    R, P = tables.shape[1], tables.shape[2]
    out = torch.empty_like(wrapped)
    with _on(wrapped.device):
        rc = _lib.lib().golf_wavetable_read_fwd(_ptr(wrapped), _ptr(tables), _ptr(out), B, N, R, P, hop_tab, _stream())
    check(rc, "golf_wavetable_read_fwd")
    return out


# ------------------------------------------------------------------------- helpers
def linear_upsample(x, hop: int) -> torch.Tensor:
    """F.interpolate(linear, align_corners=True) along the last dim (ATen arithmetic)."""
    x = _cuda_f32(x, "x")
    n = x.shape[-1]
    rows = x.numel() // n
    out = torch.empty(x.shape[:-1] + ((n - 1) * hop + 1,), dtype=torch.float32, device=x.device)
    with _on(x.device):
        rc = _lib.lib().golf_linear_upsample(_ptr(x), _ptr(out), rows, n, hop, _stream())
    check(rc, "golf_linear_upsample")
    return out


class _Rc2Lpc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, max_abs):
        lg = _cuda_f32(logits, "logits")
        M = lg.shape[-1]
        a = torch.empty_like(lg)
        with _on(lg.device):
            rc = _lib.lib().golf_rc2lpc_fwd(_ptr(lg), _ptr(a), lg.numel() // M, M, float(max_abs), _stream())
        check(rc, "golf_rc2lpc_fwd")
        ctx.save_for_backward(lg)
        ctx.max_abs = float(max_abs)
        return a

    @staticmethod
    def backward(ctx, g):
        (lg,) = ctx.saved_tensors
        g = _cuda_f32(g, "d_a")
        M = lg.shape[-1]
        d = torch.empty_like(lg)
        with _on(lg.device):
            rc = _lib.lib().golf_rc2lpc_bwd(_ptr(lg), _ptr(g), _ptr(d), lg.numel() // M, M, ctx.max_abs, _stream())
        check(rc, "golf_rc2lpc_bwd")
        return d, None


def rc2lpc(logits, max_abs: float = 1.0) -> torch.Tensor:
    """a = step_up(tanh(logits)*max_abs), last dim = order (models/utils.py:581-593 with the tanh * max_abs of
    models/filters.py:80): one launch instead of ~3M tiny torch ops; differentiable for M <= 40."""
    return _Rc2Lpc.apply(logits, float(max_abs))


def exp_complex(x) -> torch.Tensor:
    """exp(x) + 0j as complex64 in one pass (first step of the zero-phase FIR design, inference path)."""
    x = _cuda_f32(x, "x")
    out = torch.empty(x.shape, dtype=torch.complex64, device=x.device)
    with _on(x.device):
        rc = _lib.lib().golf_exp_to_complex(_ptr(x), out.data_ptr(), x.numel(), _stream())
    check(rc, "golf_exp_to_complex")
    return out
