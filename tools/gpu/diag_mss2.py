import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from golf_b200 import loss as GL
from test_gpu_mss import ref_loss, signals, rel
dev = "cuda:0"
for (B, L, kind, seed) in ((5, 4097, "noise", 5 + 4097), (3, 24001, "harmonic", 3 + 24001), (2, 12000, "noise", 12002)):
    pred, true = signals(B, L, seed, kind)
    for nf in ((509,), (1021,), (2053,)):
        p64 = pred.clone().double().requires_grad_()
        l64 = ref_loss(p64, true, nf); (g64,) = torch.autograd.grad(l64, p64)
        p32 = pred.clone().requires_grad_()
        l32 = ref_loss(p32, true, nf, dtype=torch.float32); (g32,) = torch.autograd.grad(l32, p32)
        pd = pred.to(dev).requires_grad_()
        ours = GL.mss_loss(pd, true.to(dev), nf); (go,) = torch.autograd.grad(ours, pd)
        pd1 = pred.to(dev).requires_grad_(); (go1,) = torch.autograd.grad(GL.mss_loss(pd1, true.to(dev), nf, precision=1), pd1)
        e = (go.cpu().double() - g64)
        pos = e.abs().argmax().item()
        pad = nf[0] // 2
        # error energy by region: left edge [0, pad], interior, right edge
        def reg(a, b): return float(e[:, a:b].norm() / g64[:, a:b].norm())
        print(f"B={B} L={L} {kind} n_fft={nf[0]}: loss rel {abs(float(ours)-float(l64))/float(l64):.1e} grad rel ours {rel(go, g64):.2e} (1x bwd {rel(go1, g64):.2e}) torch32 {rel(g32, g64):.2e} | left {reg(0, pad+1):.2e} mid {reg(pad+1, L-pad-1):.2e} right {reg(L-pad-1, L):.2e} | worst at b={pos // L} t={pos % L}")
