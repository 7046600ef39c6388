#!/usr/bin/env python
"""Golden vectors for LTVZeroPhaseFIRFilterPrecise (models/filters.py:286-337), produced by the
UNMODIFIED reference module imported from /root/reference (build container only; see
oracle/refimport.py for the stubs, none of which is on this path).

    python tests/golden/make_golden_precise_fir.py   ->  tests/golden/fir_precise.npz

Cases: T longer than the control range (output truncated to (F-1)*hop+1), T shorter, and the
autograd of the reference module for a fixed upstream gradient.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refimport  # noqa: E402


def main():
    refimport.import_reference()
    from models.audiotensor import AudioTensor
    from models.filters import LTVZeroPhaseFIRFilterPrecise

    torch.manual_seed(2434)
    H, Fr, n_mag = 240, 11, 64
    f = LTVZeroPhaseFIRFilterPrecise("hanning", n_mag=n_mag)
    lm = (0.5 * torch.randn(2, Fr, n_mag) - 2.0)
    out = {"hop": H, "log_mag": lm.numpy()}
    for tag, Tn in (("long", 2641), ("short", 2000)):
        ex = torch.randn(2, Tn)
        y = f(AudioTensor(ex), AudioTensor(lm, hop_length=H)).as_tensor()
        out[f"ex_{tag}"], out[f"y_{tag}"] = ex.numpy(), y.numpy()
    with torch.enable_grad():
        ex = torch.tensor(out["ex_long"]).requires_grad_()
        lmg = lm.clone().requires_grad_()
        y = f(AudioTensor(ex), AudioTensor(lmg, hop_length=H)).as_tensor()
        torch.manual_seed(7)
        up = torch.randn_like(y)
        d_ex, d_lm = torch.autograd.grad(y, (ex, lmg), up)
        out.update(up=up.numpy(), d_ex=d_ex.numpy(), d_log_mag=d_lm.numpy())
    path = os.path.join(HERE, "fir_precise.npz")
    np.savez(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
