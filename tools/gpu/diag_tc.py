"""A/B of the GOLF-ss response pass: FP32 kernel (mode 0) vs tensor cores 3xTF32 (1) vs 1xTF32 (2).
Accuracy per utterance against the float64 oracle on the bench inputs (encoder-derived controls, B = 32 x 2 s) and on the
RTF-grid synthetic controls; device time of the response pass alone and of the whole filter (CUDA events, L2-warm)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from conftest import synthetic_controls
from golf_b200 import functional as G, _lib
from oracle import golf_oracle as O
O.build()
L = _lib.lib()
dev = "cuda:0"
def rows(x, y):
    x, y = x.double().cpu(), y.double().cpu()
    return (((x - y) ** 2).mean(-1) / (y ** 2).mean(-1)).sqrt()
def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
cases = []
s = bench.make_inputs(1, 32)[0]
ex = torch.randn(32, bench.T, generator=torch.Generator().manual_seed(7)) * 0.1
cases.append(("encoder-derived controls B=32 M=22 hop=240", ex, s["gain"], s["a"], 240))
for M, hop, B in ((22, 240, 128), (20, 240, 32), (22, 120, 32)):
    gain, a = synthetic_controls(B, 48000 // hop + 1, M, seed=100 + M + hop)
    cases.append((f"synthetic B={B} M={M} hop={hop}", torch.randn(B, 48000, generator=torch.Generator().manual_seed(M)), gain, a, hop))
for name, ex, gain, a, hop in cases:
    r32, r64 = O.lpc_ss_fused(ex, gain, a, hop), O.lpc_ss_fused(ex, gain, a, hop, double=True)
    floor = rows(r32, r64)
    print(f"{name}: float32 floor max {floor.max():.2e} median {floor.median():.2e}")
    exd, gd, ad = ex.to(dev), gain.to(dev), a.to(dev)
    for mode in (0, 1):
        L.golf_lpc_ss_set_response(mode)
        for rname, tol, refine in (("no refine", 1e-4, False), ("adaptive 1e-4", 1e-4, True), ("forced", 0.0, True)):
            L.golf_lpc_ss_set_refine_tolerance(tol)
            y = G.lpc_ss(exd, gd, ad, hop, refine=refine)
            e = rows(y, r64)
            t_all = timeit(lambda: G.lpc_ss(exd, gd, ad, hop, refine=refine))
            print(f"   mode {mode} {rname:14s}: max {e.max():.2e} median {e.median():.2e} rows > 1e-4: {int((e > 1e-4).sum())}  > 10x floor: {int((e > 10 * floor).sum())}   filter {t_all:7.1f} us")
        B, Tn = exd.shape
        Fr, M = ad.shape[1], ad.shape[2]
        Lf = G.lpc_ss_length(Tn, Fr, hop)
        ws = torch.empty(L.golf_lpc_ss_workspace_bytes(B, Lf, M, hop, 0), dtype=torch.uint8, device=dev)
        yb = torch.empty(B, Lf, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        t_r = timeit(lambda: L.golf_lpc_ss_fwd_passes(exd.data_ptr(), exd.stride(0), gd.data_ptr(), ad.data_ptr(), 0, yb.data_ptr(), B, Lf, Fr, M, hop, 0, ws.data_ptr(), ws.numel(), 1, st))
        print(f"   mode {mode} response pass alone: {t_r:7.1f} us")
L.golf_lpc_ss_set_response(0); L.golf_lpc_ss_set_refine_tolerance(1e-4)
