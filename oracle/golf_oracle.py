"""CPU restatement of GOLF's synthesis hot path -- TEST INFRASTRUCTURE ONLY.

See ``oracle/__init__.py`` for who may import this.  Every function cites the
reference file:line (under /root/reference) whose behaviour it restates.  The
sample recurrences run in C (``golf_oracle.c``, OpenMP over the batch like the
reference's numba ``prange`` / ATen ``parallel_for``); frame-rate transforms and
FIR stages are torch-CPU, because that *is* the reference's arithmetic (ATen).

Pinning status (also in DESIGN.md):
  * ff filter, noise FIR, room FIR, wavetable read, control transforms, linear
    upsample: pinned against the unmodified reference modules run in the build
    container (tests/golden/*.npz, made by tests/golden/make_golden.py) and, for
    the LTI recurrence and the upsample, bit-exact against torchaudio / ATen.
  * ss recurrence (torchlpc.sample_wise_lpc) and the 4x decimator
    (kazane.Decimate): third-party, unpinned, absent -> PARITY UNPINNED; the
    restatement follows the reference call sites and is cross-checked against
    lfilter for constant coefficients and against fp64.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "golf_oracle.c")
_SO = os.path.join(_HERE, "_build", "libgolf_oracle.so")
_LIB = None


# --------------------------------------------------------------------------- build
def build(force: bool = False) -> str:
    """Compile golf_oracle.c (gcc, OpenMP when the runtime is there)."""
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(_SRC):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    base = ["-O2", "-fPIC", "-ffp-contract=off", "-fvisibility=hidden", "-shared", "-o", _SO, _SRC, "-lm"]
    last = None
    for cc in ("/usr/bin/gcc", "gcc", "cc"):
        for omp in (["-fopenmp"], []):
            try:
                subprocess.run([cc] + omp + base, check=True, capture_output=True)
                return _SO
            except (OSError, subprocess.CalledProcessError) as e:  # try next recipe
                last = e
    raise RuntimeError(f"could not build the oracle: {last}")


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.oracle_num_threads.restype = ctypes.c_int
    return _LIB


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def set_num_threads(n: int) -> None:
    lib().oracle_set_num_threads(ctypes.c_int(int(n)))


def _p(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _i(v) -> ctypes.c_int64:
    return ctypes.c_int64(int(v))


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to("cpu", torch.float32).contiguous()


# ------------------------------------------------------------------ recurrences (C)
def sample_wise_lpc(x: torch.Tensor, A: torch.Tensor, zi: Optional[torch.Tensor] = None) -> torch.Tensor:
    """torchlpc.sample_wise_lpc contract (call site models/filters.py:112): x [B,T],
    A [B,T,M], zi [B,M] (zi[:,0] = y[-1]).  float32 or float64."""
    dt = torch.float64 if x.dtype == torch.float64 else torch.float32
    x = x.detach().to("cpu", dt).contiguous()
    A = A.detach().to("cpu", dt).contiguous()
    zi = None if zi is None else zi.detach().to("cpu", dt).contiguous()
    B, T = x.shape
    M = A.shape[2]
    assert A.shape[:2] == (B, T)
    y = torch.empty_like(x)
    fn = lib().oracle_sample_wise_lpc_f64 if dt == torch.float64 else lib().oracle_sample_wise_lpc_f32
    fn(_p(x), _p(A), _p(zi), _p(y), _i(B), _i(T), _i(M))
    return y


def allpole_lti(x: torch.Tensor, a: torch.Tensor) -> torch.Tensor:
    """IIR core of torchaudio lfilter with a0 == 1, b == [1,0,..] (models/lpc.py:11-16).
    x [C,N], a [C,M] -> [C,N].  Bit-identical to lfilter on CPU in float32."""
    dt = torch.float64 if x.dtype == torch.float64 else torch.float32
    x = x.detach().to("cpu", dt).contiguous()
    a = a.detach().to("cpu", dt).contiguous()
    C, N = x.shape
    M = a.shape[1]
    y = torch.empty_like(x)
    fn = lib().oracle_allpole_lti_f64 if dt == torch.float64 else lib().oracle_allpole_lti_f32
    fn(_p(x), _p(a), _p(y), _i(C), _i(N), _i(M))
    return y


def biquad_cascade(x: torch.Tensor, biquads: torch.Tensor) -> torch.Tensor:
    """models/lpc.py:115-118: x [C,N], biquads [C,K,3] -> [C,N]."""
    x, biquads = _f32(x), _f32(biquads)
    C, N = x.shape
    K = biquads.shape[1]
    y = torch.empty_like(x)
    lib().oracle_biquad_cascade_f32(_p(x), _p(biquads), _p(y), _i(C), _i(N), _i(K))
    return y


def linear_upsample_c(x: torch.Tensor, hop: int) -> torch.Tensor:
    """C twin of F.interpolate(linear, align_corners=True) along the last dim."""
    x = _f32(x)
    n = x.shape[-1]
    rows = x.numel() // n
    out = torch.empty(x.shape[:-1] + ((n - 1) * hop + 1,), dtype=torch.float32)
    lib().oracle_linear_upsample_f32(_p(x), _p(out), _i(rows), _i(n), _i(hop))
    return out


# -------------------------------------------------------- rate handling (audiotensor)
def upsample_time(x: torch.Tensor, hop: int) -> torch.Tensor:
    """models/audiotensor/audiotensor.py:11-17,77-99: linear upsample of the time axis
    (dim 1) to length (n-1)*hop+1 with ATen's own kernel."""
    if hop == 1:
        return x
    xt = x.transpose(1, -1) if x.ndim > 2 else x
    n = xt.shape[-1]
    up = F.interpolate(xt.reshape(-1, 1, n), (n - 1) * hop + 1, mode="linear", align_corners=True)
    up = up.view(*xt.shape[:-1], -1)
    return up.transpose(1, -1) if x.ndim > 2 else up


def mixed_rate_mul(ex: torch.Tensor, gain: torch.Tensor, hop: int) -> torch.Tensor:
    """audiotensor.py:134-152 for `ex * gain`: upsample the coarser operand, truncate
    both to the shorter."""
    g = upsample_time(gain, hop)
    n = min(ex.shape[1], g.shape[1])
    return ex[:, :n] * g[:, :n]


# ------------------------------------------------------------- GOLF-ss filter (a10)
def lpc_ss(ex: torch.Tensor, gain: torch.Tensor, a: torch.Tensor, hop: int) -> torch.Tensor:
    """LTVMinimumPhaseFilterPrecise.forward, models/filters.py:99-113, literally:
    materialise gain and a at sample rate, multiply, recurrence."""
    ex, gain, a = _f32(ex), _f32(gain), _f32(a)
    e = mixed_rate_mul(ex, gain, hop)
    a_up = upsample_time(a, hop)[:, : e.shape[1]]
    e = e[:, : a_up.shape[1]]
    return sample_wise_lpc(e, a_up)


def lpc_ss_fused(ex: torch.Tensor, gain: torch.Tensor, a: torch.Tensor, hop: int, double: bool = False) -> torch.Tensor:
    """Same result as lpc_ss without the [B,T,M] temporary (coefficients interpolated
    inside the C loop with the pinned ATen arithmetic); `double` = fp64 truth."""
    ex, gain, a = _f32(ex), _f32(gain), _f32(a)
    B, Tex = ex.shape
    Fr, M = a.shape[1], a.shape[2]
    L = min(Tex, (Fr - 1) * hop + 1)
    y = torch.empty(B, L, dtype=torch.float64 if double else torch.float32)
    fn = lib().oracle_lpc_ss_f64 if double else lib().oracle_lpc_ss_f32
    fn(_p(ex), _p(gain), _p(a), _p(y), _i(B), _i(Tex), _i(Fr), _i(M), _i(hop))
    return y


# ------------------------------------------------------------- GOLF-ff filter (a11)
def get_window(name: str, n: int) -> torch.Tensor:
    """models/utils.py:417-430 (torch windows are periodic by default)."""
    fns = {
        "hanning": torch.hann_window,
        "hamming": torch.hamming_window,
        "blackman": torch.blackman_window,
        "bartlett": torch.bartlett_window,
    }
    return fns[name](n)


def _frames_and_ola(e: torch.Tensor, hop: int, W: int, n_ctrl: int):
    p = W // 2
    padded = F.pad(e, (p, p))
    n_frames = (padded.shape[1] - W) // hop + 1
    assert n_frames <= n_ctrl, f"{n_frames} frames but only {n_ctrl} control frames"
    idx = torch.arange(n_frames)[:, None] * hop + torch.arange(W)[None, :]
    return padded[:, idx], n_frames, p


def _overlap_add(frames: torch.Tensor, win: torch.Tensor, hop: int, p: int) -> torch.Tensor:
    """conv_transpose1d(diag(window), stride=hop, padding=p) of [frames; ones] followed
    by the division (models/filters.py:169-180), written as an index-add."""
    B, n_frames, W = frames.shape
    full = (n_frames - 1) * hop + W
    pos = (torch.arange(n_frames)[:, None] * hop + torch.arange(W)[None, :]).reshape(-1)
    win = win.to(frames.dtype)
    acc = torch.zeros(B, full, dtype=frames.dtype)
    acc.index_add_(1, pos, (frames * win).reshape(B, -1))
    norm = torch.zeros(full, dtype=frames.dtype)
    norm.index_add_(0, pos, win.repeat(n_frames))
    out_len = full - 2 * p
    return acc[:, p : p + out_len] / norm[p : p + out_len]


def lpc_ff(ex, gain, a, hop: int, window_length: int, centred: bool = True, window: str = "hanning", double: bool = False) -> torch.Tensor:
    """LTVMinimumPhaseFilter.forward, models/filters.py:131-184.  double=True: the same computation in float64 (truth for the
    float32 floor of resonant frames; returns float64)."""
    if double:
        ex, gain, a = (t.detach().to("cpu", torch.float64) for t in (ex, gain, a))
    else:
        ex, gain, a = _f32(ex), _f32(gain), _f32(a)
    assert window_length >= 2 * hop
    if not centred:
        ex = ex[:, hop // 2 :]
    e = mixed_rate_mul(ex, gain, hop)
    frames, n_frames, p = _frames_and_ola(e, hop, window_length, a.shape[1])
    B = e.shape[0]
    coef = a[:, :n_frames].reshape(B * n_frames, -1)
    filt = allpole_lti(frames.reshape(B * n_frames, window_length), coef).view(B, n_frames, window_length)
    y = _overlap_add(filt, get_window(window, window_length), hop, p)
    if not centred:
        y = F.pad(y[:, None], (hop // 2, 0), "reflect")[:, 0]
    return y


def biquad_ff(ex, gain, biquads, hop: int, window_length: Optional[int] = None, window: str = "hann") -> torch.Tensor:
    """BatchSecondOrderLPCSynth.forward, models/lpc.py:94-131 (padding (W-hop)//2,
    gain applied per frame, cascade of K sections, Hann OLA + normalise)."""
    ex, gain, biquads = _f32(ex), _f32(gain), _f32(biquads)
    W = hop * 4 if window_length is None else window_length
    p = (W - hop) // 2
    padded = F.pad(ex, (p, p))
    n_frames = (padded.shape[1] - W) // hop + 1
    assert n_frames <= biquads.shape[1]
    idx = torch.arange(n_frames)[:, None] * hop + torch.arange(W)[None, :]
    frames = padded[:, idx] * gain[:, :n_frames, None]
    B = ex.shape[0]
    K = biquads.shape[2]
    filt = biquad_cascade(frames.reshape(B * n_frames, W), biquads[:, :n_frames].reshape(B * n_frames, K, 3))
    win = torch.hann_window(W) if window in ("hann", "hanning") else get_window(window, W)
    return _overlap_add(filt.view(B, n_frames, W), win, hop, p)


# ------------------------------------------------------ inverse / analysis filter (a17)
def lpc_inverse(y: torch.Tensor, a: torch.Tensor, hop: int) -> torch.Tensor:
    """LTVMinimumPhaseFilter.reverse + fir_filt, models/filters.py:186-195,
    models/utils.py:433-441: r[t] = y[t] + sum_i a_up[t,i] y[t-1-i].  Differentiable when the inputs
    require grad (torch ops only)."""
    if not (y.requires_grad or a.requires_grad):
        y, a = _f32(y), _f32(a)
    a_up = upsample_time(a, hop)
    n = min(y.shape[1], a_up.shape[1])
    y, a_up = y[:, :n], a_up[:, :n]
    M = a.shape[2]
    yp = F.pad(y, (M, 0))
    hist = torch.stack([yp[:, M - 1 - i : M - 1 - i + n] for i in range(M)], dim=-1)
    return y + (hist * a_up).sum(-1)


# ----------------------------------------------------------------- noise FIR (a8)
def zero_phase_fir(log_mag: torch.Tensor, window: str = "hanning") -> torch.Tensor:
    """models/filters.py:294-306: exp -> irfft -> fftshift -> window."""
    fir = torch.fft.irfft(torch.exp(log_mag) + 0j, dim=-1)
    fir = torch.fft.fftshift(fir, dim=-1)
    return fir * get_window(window, fir.shape[-1])


def ltv_fir_blocks(ex: torch.Tensor, kernel: torch.Tensor, hop: int) -> torch.Tensor:
    """models/filters.py:360-384: block k (hop samples) is the valid cross-correlation
    of padded ex[k*hop : k*hop+K+hop-1] with kernel k."""
    if not (ex.requires_grad or kernel.requires_grad):
        ex, kernel = _f32(ex), _f32(kernel)
    B, T = ex.shape
    K = kernel.shape[-1]
    p = (K - 1) // 2
    padded = F.pad(ex, (p, p))
    n_blocks = (padded.shape[1] - (K + hop - 1)) // hop + 1
    n_blocks = min(n_blocks, kernel.shape[1])
    # one group per (utterance, block): the reference's own formulation (grouped conv1d over the
    # unfolded input), which is also what makes this leg a fair CPU baseline
    segs = padded.unfold(1, K + hop - 1, hop)[:, :n_blocks]
    out = F.conv1d(segs.reshape(1, B * n_blocks, K + hop - 1), kernel[:, :n_blocks].reshape(B * n_blocks, 1, K),
                   groups=B * n_blocks)
    return out.view(B, -1)


def noise_fir(ex, log_mag, hop: int, window: str = "hanning") -> torch.Tensor:
    """LTVZeroPhaseFIRFilter.forward, models/filters.py:350-384."""
    return ltv_fir_blocks(ex, zero_phase_fir(_f32(log_mag), window), hop)


def noise_fir_precise(ex: torch.Tensor, log_mag: torch.Tensor, hop: int, window: str = "hanning") -> torch.Tensor:
    """LTVZeroPhaseFIRFilterPrecise.forward, models/filters.py:308-337: the frame kernels are upsampled
    to sample rate (reduce_hop_length -> F.interpolate linear, align_corners=True, audiotensor.py:11-17),
    the input is padded (K-1)//2 left / K-1-(K-1)//2 right and unfolded into K-windows, and every sample is
    the dot product of its window with its own kernel; the mixed-rate matmul truncates to the shorter
    operand: min(T, (F-1)*hop + 1) samples."""
    kernel = zero_phase_fir(log_mag if log_mag.requires_grad else _f32(log_mag), window)  # [B, F, K]
    B, Fr, K = kernel.shape
    n_up = (Fr - 1) * hop + 1
    up = F.interpolate(kernel.transpose(1, 2), n_up, mode="linear", align_corners=True).transpose(1, 2)  # [B, n_up, K]
    pl = (K - 1) // 2
    win = F.pad(ex, (pl, K - 1 - pl)).unfold(1, K, 1)  # [B, T, K]
    n = min(win.shape[1], n_up)
    return (win[:, :n] * up[:, :n]).sum(-1)


# ------------------------------------------------------------------ room FIR (a18)
def room_fir(x: torch.Tensor, k: torch.Tensor) -> torch.Tensor:
    """LTIAcousticFilter.forward, models/filters.py:443-450:
    out[t] = x[t] + sum_{j<len(k)} k[j] x[t-len(k)+j]."""
    if not (x.requires_grad or k.requires_grad):
        x, k = _f32(x), _f32(k)
    n = k.numel()
    xp = F.pad(x[:, None, :-1], (n, 0))
    return x + F.conv1d(xp, k[None, None, :])[:, 0]


# ------------------------------------------------------------- control transforms
def rc2lpc(rc: torch.Tensor) -> torch.Tensor:
    """Step-up recursion, models/utils.py:581-593 (rc already tanh'ed/scaled)."""
    M = rc.shape[-1]
    poly = torch.ones_like(rc[..., :1])
    for n in range(M):
        ext = F.pad(poly, (0, 1))
        poly = ext + rc[..., n : n + 1] * ext.flip(-1)
    return poly[..., 1:] if M > 1 else rc


def logits2biquads(logits: torch.Tensor, rep: str = "coef", rho: float = 0.99) -> torch.Tensor:
    """models/utils.py:487-525; logits [...,K,2] -> [...,K,3]."""
    l0, l1 = logits[..., 0], logits[..., 1]
    if rep == "coef":
        a1 = torch.tanh(l0) * rho * 2
        a2 = 0.5 * ((2 - a1.abs()) * torch.tanh(l1) * rho + a1.abs())
    elif rep == "conj":
        mag = torch.sigmoid(l0) * rho
        a1 = -2 * mag * torch.tanh(l1)
        a2 = mag.square()
    elif rep == "real":
        z1, z2 = torch.tanh(l0) * rho, torch.tanh(l1) * rho
        a1, a2 = -z1 - z2, z1 * z2
    else:
        raise ValueError(rep)
    return torch.stack([torch.ones_like(a1), a1, a2], -1)


def biquads2lpc(biquads: torch.Tensor) -> torch.Tensor:
    """Polynomial product of the sections (models/utils.py:444-484) -> [..., 2K]."""
    K = biquads.shape[-2]
    poly = biquads[..., 0, :]
    for k in range(1, K):
        nxt = biquads[..., k, :]
        out = torch.zeros(poly.shape[:-1] + (poly.shape[-1] + 2,), dtype=poly.dtype)
        for j in range(3):
            out[..., j : j + poly.shape[-1]] += poly * nxt[..., j : j + 1]
        poly = out
    return poly[..., 1:]


# ------------------------------------------------------- glottal wavetable (a4-a6)
def lf_table_v2(Rd: torch.Tensor, points: int) -> torch.Tensor:
    """LF glottal-flow derivative, one period per R_d (models/utils.py:363-400)."""
    Rd = Rd.reshape(-1, 1)
    Ra = 0.048 * Rd - 0.01
    Rk = 0.118 * Rd + 0.224
    Rg = (Rk / 4) * (0.5 + 1.2 * Rk) / (0.11 * Rd - Ra * (0.5 + 1.2 * Rk))
    Ta, Tp = Ra, 1 / (2 * Rg)
    Te = Tp + Tp * Rk
    eps = 1 / Ta
    shift = torch.exp(-eps * (1 - Te))
    delta = 1 - shift
    rhs = ((1 / eps) * (shift - 1) + (1 - Te) * shift) / delta
    upper = -(-(Te - Tp) / 2 + rhs)
    omega = torch.pi / Tp
    s = torch.sin(omega * Te)
    alpha = torch.log(-torch.pi * s * upper / (Tp * 2)) / (Tp / 2 - Te)
    E0 = -1 / (s * torch.exp(alpha * Te))
    t = torch.linspace(0, 1, points + 1)[None, :-1]
    rise = E0 * torch.exp(alpha * t) * torch.sin(omega * t)
    ret = (shift - torch.exp(-eps * (t - Te))) / delta
    return torch.where(t < Te, rise, ret)


def glottal_table(table_size=100, points=2048, min_R_d=0.3, max_R_d=2.7) -> Tuple[torch.Tensor, torch.Tensor]:
    """GlottalFlowTable.__init__ for the shipped options (derivative, align_peak,
    constant_power, lf_v2), models/synth.py:59-120."""
    Rd = torch.exp(torch.linspace(math.log(min_R_d), math.log(max_R_d), table_size))
    tab = lf_table_v2(Rd, points)
    peak = tab.argmin(1)
    tgt = int(peak.max())
    tab = torch.stack([torch.roll(tab[i], tgt - int(peak[i])) for i in range(table_size)])
    tab = tab / tab.norm(dim=1, keepdim=True) * math.sqrt(points)
    return tab, Rd


def select_tables(table: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """Row interpolation by the selection weight, models/synth.py:223-237. w [B,Fw]."""
    n = table.shape[0]
    raw = w * (n - 1)
    lo = raw.long().clamp_(0, n - 2)
    p = (raw - lo).unsqueeze(-1)
    return table[lo] * (1 - p) + table[lo + 1] * p


def wavetable_read(wrapped: torch.Tensor, tables: torch.Tensor, hop_tab: int) -> torch.Tensor:
    """GlottalFlowTable.generate, models/synth.py:124-177, with F.grid_sample's bilinear
    align_corners=True arithmetic written out (unnormalise, floor, 4 weighted taps)."""
    B, N = wrapped.shape
    blocks = (N + hop_tab - 1) // hop_tab
    if tables.shape[1] < blocks + 1:
        tables = torch.cat([tables, tables[:, -1:].expand(-1, blocks + 1 - tables.shape[1], -1)], 1)
    else:
        tables = tables[:, : blocks + 1]
    P = tables.shape[2]
    img = torch.cat([tables, tables[:, :, :1]], 2)  # column P wraps to column 0
    gx = wrapped * 2 - 1
    gy = torch.arange(N, dtype=wrapped.dtype)[None, :] / (hop_tab * blocks) * 2 - 1
    ix = ((gx + 1) / 2) * P  # (W-1) with W = P+1
    iy = ((gy + 1) / 2) * blocks  # (H-1) with H = blocks+1
    x0, y0 = ix.floor(), iy.floor()
    fx, fy = ix - x0, iy - y0
    x0, y0 = x0.long(), y0.long()

    def tap(yy, xx):
        ok = (xx >= 0) & (xx <= P) & (yy >= 0) & (yy <= blocks)
        v = img[torch.arange(B)[:, None], yy.clamp(0, blocks), xx.clamp(0, P)]
        return torch.where(ok, v, torch.zeros_like(v))

    return (
        tap(y0, x0) * ((1 - fx) * (1 - fy))
        + tap(y0, x0 + 1) * (fx * (1 - fy))
        + tap(y0 + 1, x0) * ((1 - fx) * fy)
        + tap(y0 + 1, x0 + 1) * (fx * fy)
    )


def decimate_kernel(q: int, zeros: int = 16) -> torch.Tensor:
    """Restated kazane.Decimate low-pass (models/synth.py:208; third-party, absent,
    PARITY UNPINNED): Hann-windowed sinc, `zeros` zero-crossings a side at the
    decimated rate, cutoff at the new Nyquist, unit DC gain before windowing."""
    half = zeros * q
    n = torch.arange(-half, half + 1, dtype=torch.float64)
    h = torch.sinc(n / q) / q * torch.hann_window(2 * half + 1, periodic=False, dtype=torch.float64)
    return h.float()


def decimate(x: torch.Tensor, q: int, zeros: int = 16) -> torch.Tensor:
    """y[m] = sum_n h[n] xpad[m q + n], xpad = x zero-padded by zeros*q each side;
    len_out = (len_in - 1)//q + 1."""
    h = decimate_kernel(q, zeros)
    return F.conv1d(x[:, None], h[None, None], stride=q, padding=zeros * q)[:, 0]


def phase_accumulate(upsampled_phase: torch.Tensor, mode: str = "fp32") -> torch.Tensor:
    """models/synth.py:250-255: running sum then mod 1.  "fp32" is the reference's
    arithmetic (ATen CPU cumsum is a sequential float accumulate); "fp64" is truth."""
    if mode == "fp32":
        return torch.cumsum(upsampled_phase.float(), 1) % 1
    return (torch.cumsum(upsampled_phase.double(), 1) % 1).float()


def glottal_osc(
    phase: torch.Tensor,
    phase_hop: int,
    w: torch.Tensor,
    w_hop: int,
    table: torch.Tensor,
    oversampling: int = 4,
    equal_energy: bool = True,
    accumulate: str = "fp32",
    wrapped_override: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """IndexedGlottalFlowTable.forward, models/synth.py:213-263.  phase [B,Np] at
    `phase_hop` (cycles/sample), w [B,Fw] at `w_hop`."""
    phase = _f32(phase)
    if not table.requires_grad:  # a trainable table (models/synth.py:117-118) keeps its autograd graph
        table = _f32(table)
    if not w.requires_grad:
        w = _f32(w)
    tables = select_tables(table, w)
    up = upsample_time(phase / oversampling, phase_hop * oversampling)
    wrapped = phase_accumulate(up, accumulate) if wrapped_override is None else wrapped_override
    y = wavetable_read(wrapped, tables, w_hop * oversampling)
    if equal_energy:
        y = y * torch.rsqrt(up)
    if oversampling > 1:
        y = decimate(y, oversampling)
    return y


# ------------------------------------------------------------ whole decoder (a1)
def source_filter_synth(
    phase, phase_hop, w, w_hop, log_mag, gain, a, hop, noise, table, room_kernel,
    variant: str = "ss", window_length: int = 960, oversampling: int = 4,
    accumulate: str = "fp32", stages: Optional[dict] = None,
) -> torch.Tensor:
    """SourceFilterSynth.forward with subtract_harmonics=False, voicing=None
    (models/sf.py:47-64; cfg/ae/decoder/golf{,-precise}.yaml).  `noise` is the
    randn_like draw, injected so both sides see the same samples."""
    harm = glottal_osc(phase, phase_hop, w, w_hop, table, oversampling, True, accumulate)
    nz = noise_fir(noise[:, : harm.shape[1]], log_mag, hop)
    n = min(harm.shape[1], nz.shape[1])
    src = harm[:, :n] + nz[:, :n]
    if variant == "ss":
        y = lpc_ss(src, gain, a, hop)
    else:
        y = lpc_ff(src, gain, a, hop, window_length)
    out = room_fir(y, room_kernel) if room_kernel is not None else y
    if stages is not None:
        stages.update(harm=harm, noise_filtered=nz, src=src, lpc=y, out=out)
    return out
