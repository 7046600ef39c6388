"""Small driver for ncu: runs each filter kernel a few times at the bench shape (B=32, 2 s)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from golf_b200 import functional as G
dev = torch.device("cuda:0")
B, T, H, M = 32, 48000, 240, 22
Fr = T // H
torch.manual_seed(0)
k = torch.tanh(0.15 * torch.randn(B, Fr, M))
a = k.clone()
# cheap stable coefficients: small random reflection-like taps
a = (0.3 * torch.randn(B, Fr, M) / (1 + torch.arange(M))).to(dev)
gain = torch.exp(torch.randn(B, Fr) - 6).to(dev)
ex = torch.randn(B, T, device=dev)
win = torch.hann_window(4 * H, device=dev)
kern = torch.randn(B, Fr, 510, device=dev) * 0.01
rk = torch.randn(127, device=dev) * 0.01
which = sys.argv[1] if len(sys.argv) > 1 else "all"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
for _ in range(n):
    if which in ("all", "ss"): G.lpc_ss(ex, gain, a, H)
    if which in ("all", "ff"): G.lpc_ff(ex, gain, a, win, H)
    if which in ("all", "fir"): G.ltv_fir_blocks(ex, kern, H); G.room_fir(ex, rk)
torch.cuda.synchronize()
print("done")
