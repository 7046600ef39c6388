"""Calibrate the adaptive-refinement trigger: error vs float64 and time for several tolerances."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import synthetic_controls, rel_rms, golden, T
from golf_b200 import functional as G, _lib
from oracle import golf_oracle as O
dev = "cuda:0"
L = _lib.lib()
cases = {}
g = golden("controls_gt"); H = int(g["hop"])
gain, a = T(g["gain"]), T(g["a"])
cases["encoder-derived"] = (torch.randn(gain.shape[0], (gain.shape[1] - 1) * H, generator=torch.Generator().manual_seed(0)), gain, a, H)
f = golden("filters_rand")
for M in (8, 22):
    cases[f"filters_rand M{M}"] = (T(f[f"ex_{M}"]), T(f[f"gain_{M}"]), T(f[f"a_{M}"]), int(f["hop"]))
gain, a = synthetic_controls(2, 39, 22, seed=22 + 256)
cases["synthetic gain x100 (hop 256)"] = (torch.randn(2, 9600, generator=torch.Generator().manual_seed(1)), gain, a, 256)
gain, a = synthetic_controls(32, 200, 22, seed=9)
cases["synthetic B32 2s"] = (torch.randn(32, 48000, generator=torch.Generator().manual_seed(5)), gain, a, 240)
for name, (ex, gain, a, H) in cases.items():
    ref64 = O.lpc_ss_fused(ex, gain, a, H, double=True); floor = rel_rms(O.lpc_ss_fused(ex, gain, a, H), ref64)
    exd, gd, ad = ex.to(dev), gain.to(dev), a.to(dev)
    print(f"{name}: float32 floor {floor:.2e}")
    for tol in (0.0, 5e-5, 1e-4, 2e-4, 5e-4, 1e-3, 3e-3, 1e-2, 1e9):
        L.golf_lpc_ss_set_refine_tolerance(tol)
        y = G.lpc_ss(exd, gd, ad, H)
        for _ in range(3): G.lpc_ss(exd, gd, ad, H)
        torch.cuda.synchronize(); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10): G.lpc_ss(exd, gd, ad, H)
        e.record(); torch.cuda.synchronize()
        print(f"   tol {tol:7.0e}: err vs f64 {rel_rms(y, ref64):.2e}   {s.elapsed_time(e)*100:.0f} us")
