"""Golden vectors for the frame-wise LPC synthesis modules and the biquad parameterisations, produced by the
UNMODIFIED reference (models/lpc.py, models/utils.py) on the CPU with torchaudio's lfilter and torch autograd.

    python tests/golden/make_golden_lpc_modules.py        (build container only: needs /root/reference)

Writes tests/golden/lpc_modules.npz:
  lpc_synthesis (models/lpc.py:11-16)            x [C,N], gains [C], a [C,M] -> y, and d_x, d_gains, d_a for a fixed upstream g
  LPCSynth / BatchLPCSynth (lpc.py:19-91)        ex, gain, a -> y and gradients
  BatchSecondOrderLPCSynth (lpc.py:94-131)       ex, gain, biquads (coef parameterisation, K=4 and K=11) -> y and gradients
  get_logits2biquads + biquads2lpc (utils.py:444-525)  coef | conj | real, K=11 (ISMIR-23 configuration): biquads, a, and
                                                 d_logits for fixed upstream gradients on a and on the sections
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import refimport  # noqa: E402
from make_golden import np32, smooth  # noqa: E402


def main():
    refimport.import_reference()
    from models.lpc import BatchLPCSynth, BatchSecondOrderLPCSynth, LPCSynth, lpc_synthesis
    from models.utils import biquads2lpc, get_logits2biquads, rc2lpc

    torch.manual_seed(2434)
    out = {}
    H = 240

    # ---- lpc_synthesis: 70 channels (more than two warps of channels), N not a multiple of the tile
    C, N, M = 70, 515, 22
    x = torch.randn(C, N, requires_grad=True)
    gains = torch.exp(torch.randn(C) - 1).requires_grad_()
    a = rc2lpc(torch.tanh(0.3 * torch.randn(1, C, M)))[0].detach().requires_grad_()
    y = lpc_synthesis(x, gains, a)
    g = torch.randn_like(y)
    dx, dg, da = torch.autograd.grad(y, (x, gains, a), g)
    out.update(ls_x=np32(x), ls_gains=np32(gains), ls_a=np32(a), ls_y=np32(y), ls_g=np32(g), ls_dx=np32(dx), ls_dgains=np32(dg), ls_da=np32(da))

    # ---- BatchLPCSynth (and LPCSynth on row 0): T chosen so the frame count equals F
    for M in (8, 22):
        T = 4800
        Fr = T // H
        ex = torch.randn(2, T, requires_grad=True)
        gain = torch.exp(smooth(torch.randn(2, Fr)) - 2).requires_grad_()
        a = rc2lpc(torch.tanh(0.15 * smooth(torch.randn(2, Fr, M)))).detach().requires_grad_()
        mod = BatchLPCSynth(hop_length=H, window="hanning")
        y = mod(ex, gain, a)
        g = torch.randn_like(y)
        dex, dgain, da = torch.autograd.grad(y, (ex, gain, a), g)
        p = f"bl{M}_"
        out.update({p + "ex": np32(ex), p + "gain": np32(gain), p + "a": np32(a), p + "y": np32(y), p + "g": np32(g),
                    p + "dex": np32(dex), p + "dgain": np32(dgain), p + "da": np32(da)})
        if M == 8:
            single = LPCSynth(hop_length=H, window="hanning")
            y1 = single(ex[0].detach(), torch.cat([gain[0, :, None], a[0]], -1).detach())
            out["lp8_y"] = np32(y1)

    # ---- BatchSecondOrderLPCSynth
    for K in (4, 11):
        T = 4800
        Fr = T // H + 2  # more control frames than signal frames: the tail gets zero gradient
        ex = torch.randn(2, T, requires_grad=True)
        gain = torch.exp(smooth(torch.randn(2, Fr)) - 2).requires_grad_()
        # moderate resonances: with scale 0.6 / rho 0.99 the cascade's output reaches 1e8 and float32 implementations
        # (the reference's included) only agree to 1e-3
        bq = get_logits2biquads("coef", 0.9)((0.35 if K == 4 else 0.25) * smooth(torch.randn(2, Fr, K, 2))).detach().requires_grad_()
        mod = BatchSecondOrderLPCSynth(hop_length=H, window="hanning")
        y = mod(ex, gain, bq)
        print("cascade K", K, "max|y|", float(y.abs().max()), "rms", float(y.square().mean().sqrt()))
        g = torch.randn_like(y)
        dex, dgain, dbq = torch.autograd.grad(y, (ex, gain, bq), g)
        p = f"bq{K}_"
        out.update({p + "ex": np32(ex), p + "gain": np32(gain), p + "biquads": np32(bq), p + "y": np32(y), p + "g": np32(g),
                    p + "dex": np32(dex), p + "dgain": np32(dgain), p + "dbiquads": np32(dbq)})

    # ---- parameterisations, K = 11, rho = 0.99 (ckpts/ismir23/glottal_d_f1/config.yaml:103-110)
    K = 11
    logits = (0.8 * smooth(torch.randn(2, 40, K, 2))).detach()
    out["pm_logits"] = np32(logits)
    g_a = torch.randn(2, 40, 2 * K)
    g_bq = torch.randn(2, 40, K, 3)
    out["pm_g_a"], out["pm_g_bq"] = np32(g_a), np32(g_bq)
    for rep in ("coef", "conj", "real"):
        lg = logits.clone().requires_grad_()
        bq = get_logits2biquads(rep, 0.99)(lg)
        a = biquads2lpc(bq)
        (d_from_a,) = torch.autograd.grad(a, lg, g_a, retain_graph=True)
        (d_from_bq,) = torch.autograd.grad(bq, lg, g_bq)
        out.update({f"pm_{rep}_biquads": np32(bq), f"pm_{rep}_a": np32(a), f"pm_{rep}_dlogits_a": np32(d_from_a),
                    f"pm_{rep}_dlogits_bq": np32(d_from_bq)})
    out["hop"] = H
    np.savez(os.path.join(HERE, "lpc_modules.npz"), **out)
    print({k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
