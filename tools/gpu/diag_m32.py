"""GOLF-ss filter at the slow corner of the RTF grid (order 32): time and accuracy."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import synthetic_controls
from golf_b200 import functional as G
from oracle import golf_oracle as O
O.build(); dev = "cuda:0"
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for M, hop, B in ((32, 240, 32), (32, 120, 32), (32, 240, 1), (40, 240, 32), (20, 240, 32), (20, 120, 32)):
    gain, a = synthetic_controls(B, 48000 // hop + 1, M, seed=100 + M + hop)
    ex = torch.randn(B, 48000, generator=torch.Generator().manual_seed(M))
    r64 = O.lpc_ss_fused(ex[:4], gain[:4], a[:4], hop, double=True)
    exd, gd, ad = ex.to(dev), gain.to(dev), a.to(dev)
    y = G.lpc_ss(exd, gd, ad, hop)
    e = (((y[:4].cpu().double() - r64) ** 2).mean(1) / (r64 ** 2).mean(1)).sqrt().max()
    print(f"M={M} hop={hop} B={B}: filter {timeit(lambda: G.lpc_ss(exd, gd, ad, hop)):8.1f} us   rel err vs f64 {float(e):.2e}", flush=True)
