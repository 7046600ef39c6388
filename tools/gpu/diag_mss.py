"""MSS loss on tcgen05 vs the torch.stft restatement of loss/spec.py: value and gradient against float64, timing against cuFFT."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from golf_b200 import loss as GL
dev = "cuda:0"
NF = (509, 1021, 2053)
def ref_loss(pred, true, dtype):
    p, t = pred.to(dtype), true.to(dtype)
    tot = 0
    for n in NF:
        w = torch.hann_window(n, dtype=dtype, device=p.device)
        sp, st = (torch.stft(v, n, hop_length=int(n - n * 0.75), window=w, return_complex=True).abs() for v in (p, t))
        tot = tot + (sp - st).abs().mean() + ((st + 1e-8).log2() - (sp + 1e-8).log2()).abs().mean()
    return tot
def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())
g = torch.Generator().manual_seed(0)
for B, L, kind in ((2, 12000, "noise"), (4, 47760, "harmonic"), (32, 47760, "harmonic")):
    t_ax = torch.arange(L) / 24000.0
    if kind == "noise":
        true = 0.05 * torch.randn(B, L, generator=g); pred = true + 0.02 * torch.randn(B, L, generator=g)
    else:  # harmonic signals with 60 dB of spectral dynamic range + a little noise
        f0 = 100 + 150 * torch.rand(B, 1, generator=g)
        true = sum((0.5 ** k) * torch.sin(2 * torch.pi * k * f0 * t_ax) for k in range(1, 12)) * 0.1 + 1e-4 * torch.randn(B, L, generator=g)
        pred = sum((0.55 ** k) * torch.sin(2 * torch.pi * k * (f0 * 1.003) * t_ax + 0.3) for k in range(1, 12)) * 0.1 + 1e-4 * torch.randn(B, L, generator=g)
    pd = pred.to(dev).requires_grad_(); td = true.to(dev)
    ours = GL.mss_loss(pd, td, NF)
    (g_ours,) = torch.autograd.grad(ours, pd)
    p32 = pred.to(dev).requires_grad_()
    l32 = ref_loss(p32, td, torch.float32)
    (g32,) = torch.autograd.grad(l32, p32)
    if B <= 4:
        p64 = pred.clone().double().requires_grad_()
        l64 = ref_loss(p64, true, torch.float64)
        (g64,) = torch.autograd.grad(l64, p64)
        print(f"B={B} L={L} {kind}: loss ours {float(ours):.7f} torch32 {float(l32):.7f} f64 {float(l64):.7f} | rel err ours {abs(float(ours)-float(l64))/float(l64):.2e} torch32 {abs(float(l32)-float(l64))/float(l64):.2e}"
              f" | grad rel ours {rel(g_ours.cpu(), g64):.2e} torch32 {rel(g32.cpu(), g64):.2e}", flush=True)
    else:
        print(f"B={B} L={L} {kind}: loss ours {float(ours):.7f} torch32 {float(l32):.7f} rel {abs(float(ours)-float(l32))/float(l32):.2e} | grad rel vs torch32 {rel(g_ours, g32):.2e}", flush=True)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def ours_fb():
    pd.grad = None
    GL.mss_loss(pd, td, NF).backward()
def torch_fb():
    p32.grad = None
    ref_loss(p32, td, torch.float32).backward()
with torch.no_grad():
    t_of = timeit(lambda: GL.mss_loss(pd.detach(), td, NF)); t_tf = timeit(lambda: ref_loss(p32.detach(), td, torch.float32))
print(f"B=32 forward only: ours {t_of:.3f} ms, torch/cuFFT {t_tf:.3f} ms;  forward+backward: ours {timeit(ours_fb):.3f} ms, torch/cuFFT {timeit(torch_fb):.3f} ms")
