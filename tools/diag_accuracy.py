"""Row-wise accuracy of the GOLF-ss filter on the RTF-grid synthetic controls (B = 128): GPU vs float64 truth next to
the float32 floor (oracle32 vs oracle64), for refinement off / adaptive / forced and the tail on / off.
usage: python tools/diag_accuracy.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import synthetic_controls
from golf_b200 import functional as G, _lib
from oracle import golf_oracle as O
O.build()
L = _lib.lib()
dev = "cuda:0"
def rows(x, y):
    x, y = x.double().cpu(), y.double().cpu()
    return (((x - y) ** 2).mean(-1) / (y ** 2).mean(-1)).sqrt()
for M, hop in ((22, 240), (32, 120), (32, 240)):
    B, Tn = 128, 48000
    gain, a = synthetic_controls(B, Tn // hop + 1, M, seed=100 + M + hop)
    ex = torch.randn(B, Tn, generator=torch.Generator().manual_seed(M))
    r32, r64 = O.lpc_ss_fused(ex, gain, a, hop), O.lpc_ss_fused(ex, gain, a, hop, double=True)
    floor = rows(r32, r64)
    exd, gd, ad = ex.to(dev), gain.to(dev), a.to(dev)
    print(f"M={M} hop={hop}: float32 floor max {floor.max():.2e} (row {int(floor.argmax())}) median {floor.median():.2e}")
    for tail in (1, 0):
        L.golf_lpc_ss_set_tail(tail)
        for name, tol, refine in (("no refine", 1e-4, False), ("adaptive 1e-4", 1e-4, True), ("adaptive 1e-5", 1e-5, True), ("adaptive 1e-6", 1e-6, True), ("forced", 0.0, True)):
            L.golf_lpc_ss_set_refine_tolerance(tol)
            y = G.lpc_ss(exd, gd, ad, hop, refine=refine)
            e = rows(y, r64)
            worst = int(e.argmax())
            print(f"   tail={tail} {name:14s}: max {e.max():.2e} (row {worst}, floor there {floor[worst]:.2e})  rows > 1e-4: {int((e > 1e-4).sum())}  rows > 10x floor: {int((e > 10 * floor).sum())}")
L.golf_lpc_ss_set_tail(0); L.golf_lpc_ss_set_refine_tolerance(1e-4)
