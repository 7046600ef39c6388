"""Frame-rate control transforms and init-time tables (differentiable torch code; these
run at 100 Hz, not at audio rate).  Reference: models/utils.py:308-525,581-593."""
from __future__ import annotations

import math
from typing import Callable

import torch
import torch.nn.functional as F


def get_window_fn(window: str = "hann") -> Callable[..., torch.Tensor]:
    """name -> window constructor; torch windows are periodic (models/utils.py:417-430)."""
    table = {
        "hanning": torch.hann_window,
        "hamming": torch.hamming_window,
        "blackman": torch.blackman_window,
        "bartlett": torch.bartlett_window,
    }
    if window in table:
        return table[window]
    try:
        from scipy.signal import get_window

        get_window(window, 8)
    except Exception as e:  # noqa: BLE001
        raise ValueError(f"Unknown window function {window}") from e
    return lambda n, **kw: torch.tensor(get_window(window, n), **kw)


def rc2lpc(rc: torch.Tensor) -> torch.Tensor:
    """Reflection coefficients [B,F,M] -> direct-form a_1..a_M (step-up recursion),
    models/utils.py:581-593."""
    assert rc.ndim == 3
    M = rc.shape[-1]
    if M == 1:
        return rc
    poly = torch.cat([torch.ones_like(rc[..., :1]), rc[..., :1]], -1)
    for n in range(1, M):
        ext = F.pad(poly, (0, 1))
        poly = ext + rc[..., n : n + 1] * ext.flip(-1)
    return poly[..., 1:]


def get_logits2biquads(rep_type: str, max_abs_pole: float = 0.99) -> Callable[[torch.Tensor], torch.Tensor]:
    """logits [...,2] -> stable section [1, a1, a2] (models/utils.py:487-525)."""
    rho = max_abs_pole
    if rep_type == "coef":

        def fn(lg):
            assert lg.shape[-1] == 2
            a1 = torch.tanh(lg[..., 0]) * rho * 2
            mag = a1.abs()
            a2 = 0.5 * ((2 - mag) * torch.tanh(lg[..., 1]) * rho + mag)
            return torch.stack([torch.ones_like(a1), a1, a2], -1)

    elif rep_type == "conj":

        def fn(lg):
            assert lg.shape[-1] == 2
            r = torch.sigmoid(lg[..., 0]) * rho
            a1 = -2 * r * torch.tanh(lg[..., 1])
            return torch.stack([torch.ones_like(a1), a1, r.square()], -1)

    elif rep_type == "real":

        def fn(lg):
            assert lg.shape[-1] == 2
            z1, z2 = torch.tanh(lg[..., 0]) * rho, torch.tanh(lg[..., 1]) * rho
            return torch.stack([torch.ones_like(z1), -z1 - z2, z1 * z2], -1)

    else:
        raise ValueError(f"Unknown rep_type: {rep_type}, expected coef, conj or real")
    return fn


def coeff_product(polys) -> torch.Tensor:
    """Product of polynomials given as a sequence of [N, d_i+1] coefficient rows
    (models/utils.py:444-460), pairwise tree like the reference so the float rounding
    order matches."""
    n = len(polys)
    if n == 1:
        return polys[0]
    hi, lo = coeff_product(polys[n // 2 :]), coeff_product(polys[: n // 2])
    if hi.shape[1] > lo.shape[1]:
        hi, lo = lo, hi
    w = hi.unsqueeze(1).flip(2)
    return F.conv1d(lo.unsqueeze(0), w, padding=w.shape[2] - 1, groups=lo.shape[0]).squeeze(0)


def biquads2lpc(biquads: torch.Tensor) -> torch.Tensor:
    """[...,K,3] sections -> [...,2K] polynomial taps (models/utils.py:480-484)."""
    assert biquads.shape[-1] == 3
    flat = biquads.reshape(-1, *biquads.shape[-2:]).transpose(0, 1)
    return coeff_product(flat).reshape(*biquads.shape[:-2], -1)[..., 1:]


# ---- LF glottal-flow-derivative wavetables (init-time) ------------------------------
def lf_period_v2(Rd: torch.Tensor, points: int = 1024) -> torch.Tensor:
    """One period per R_d value, closed form (models/utils.py:363-400)."""
    Rd = torch.as_tensor(Rd).reshape(-1, 1)
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
    return torch.where(t < Te, rise, ret).squeeze()


def lf_period_v1(R_d=0.3, T_0: float = 5.0, n_iter_eps: int = 5, n_iter_a: int = 100, points: int = 1000) -> torch.Tensor:
    """Iterative LF fit used by the ISMIR-23 checkpoints (models/utils.py:308-360).  R_d may be a Python float or a
    0-dim tensor; with a float32 tensor (what GlottalFlowTable passes, like the reference) every derived quantity
    and both Newton loops are float32 tensors."""
    R_ap = 0.048 * R_d - 0.01
    R_kp = 0.118 * R_d + 0.224
    R_gp = 0.25 * R_kp * (0.5 + 1.2 * R_kp) / (0.11 * R_d - R_ap * (0.5 + 1.2 * R_kp))
    T_a = R_ap * T_0
    T_p = 0.5 * T_0 / R_gp
    T_e = T_p * (R_kp + 1)
    T_b = T_0 - T_e
    w_g = math.pi / T_p
    E_e, a, eps = 1.0, 1.0, 1.0
    for _ in range(n_iter_eps):  # Newton on eps*T_a = 1 - exp(-eps*T_b)
        eps = abs(eps - (eps * T_a + math.expm1(-eps * T_b)) / (T_a - T_b * math.exp(-eps * T_b)))
    for _ in range(n_iter_a):  # Newton on the zero-net-flow condition
        E_0 = -E_e * math.exp(-a * T_e) / math.sin(w_g * T_e)
        A_o = E_0 * math.exp(a * T_e) / math.sqrt(w_g**2 + a**2) * math.sin(w_g * T_e - math.atan(w_g / a)) + E_0 * w_g / (w_g**2 + a**2)
        A_r = -E_e / (eps**2 * T_a) * (1 - math.exp(-eps * T_b) * (1 + eps * T_b))
        grad = (1 - 2 * a * A_r / E_e) * math.sin(w_g * T_e) - w_g * T_e * math.exp(-a * T_e)
        a = a - (A_o + A_r) / grad
    t = torch.linspace(0, T_0, points + 1)[:-1]
    t_open, t_ret = t[t < T_e], t[t >= T_e]
    opening = E_0 * torch.exp(a * t_open) * torch.sin(w_g * t_open)
    closing = -E_e / eps / T_a * (torch.exp(-eps * (t_ret - T_e)) - math.exp(-eps * T_b))
    return torch.cat([opening, closing])
