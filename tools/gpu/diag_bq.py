import sys; sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import torch
from conftest import T, golden, rel_rms
from golf_b200.lpc import BatchSecondOrderLPCSynth
g = golden("lpc_modules"); DEV = "cuda:0"
for K in (4, 11):
    p = f"bq{K}_"; H = int(g["hop"])
    mod = BatchSecondOrderLPCSynth(hop_length=H, window="hanning").to(DEV)
    ex, gain, bq = [T(g[p + n]).to(DEV).requires_grad_() for n in ("ex", "gain", "biquads")]
    y = mod(ex, gain, bq)
    print(K, "y", rel_rms(y, T(g[p + "y"])), "max|y|", float(y.abs().max()))
    dex, dgain, dbq = torch.autograd.grad(y, (ex, gain, bq), T(g[p + "g"]).to(DEV))
    print("  dex", rel_rms(dex, T(g[p + "dex"])), "dgain", rel_rms(dgain, T(g[p + "dgain"])), "dbq", rel_rms(dbq.reshape(1, -1), T(g[p + "dbiquads"]).reshape(1, -1)))
    r = T(g[p + "dbiquads"]); d = dbq.cpu()
    for c in range(3): print("   col", c, rel_rms(d[..., c].reshape(1, -1), r[..., c].reshape(1, -1)))
    print("  tail", float(dbq[:, y.shape[1] // H:].abs().max()), float(dgain[:, y.shape[1] // H:].abs().max()))
