"""Stage-by-stage check of golf_mss_loss for one scale: frames, target magnitudes, spectral gradient G, frame gradients, d_pred."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from golf_b200 import _lib, loss as GL
from test_gpu_mss import signals
L_ = _lib.lib(); dev = "cuda:0"
B, L, n = 2, 12000, 509
pred, true = signals(B, L, 12002, "noise")
hop = int(n - n * 0.75); nb = n // 2 + 1; N = 2 * nb; Kp = (n + 31) // 32 * 32; Kb = (N + 31) // 32 * 32; nfr = 1 + L // hop; rows = B * nfr; nbp = (nb + 3) // 4 * 4
c_ff = (ctypes.c_int * 1)(n); c_h = (ctypes.c_int * 1)(hop)
tab = GL._tables(n, torch.device(dev)); c_t = (ctypes.c_void_p * 1)(tab.data_ptr())
nbytes = L_.golf_mss_workspace_bytes(B, L, c_ff, c_h, 1)
ws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
pd, td = pred.to(dev), true.to(dev)
loss = torch.zeros(1, device=dev); dp = torch.zeros(B, L, device=dev)
rc = L_.golf_mss_loss(pd.data_ptr(), L, td.data_ptr(), L, B, L, c_ff, c_h, 1, c_t, 1.0, 1.0, 1e-8, loss.data_ptr(), dp.data_ptr(), L, 3, ws.data_ptr(), nbytes, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize(); assert rc == 0
al = lambda x: (x + 255) // 256 * 256
fbytes = al(rows * Kp * 4); sbytes = al(rows * nbp * 4)
wf = ws.view(torch.float32)
fr_p = wf[: rows * Kp].view(rows, Kp)[:, :n].double().cpu()
dfr = wf[fbytes // 4 : fbytes // 4 + rows * Kp].view(rows, Kp)[:, :n].double().cpu()
s_true = wf[2 * fbytes // 4 : 2 * fbytes // 4 + rows * nbp].view(rows, nbp)[:, :nb].double().cpu()
G = wf[(2 * fbytes + sbytes) // 4 : (2 * fbytes + sbytes) // 4 + rows * Kb].view(rows, Kb)[:, :N].double().cpu()
# float64 reference of every stage
w = torch.hann_window(n, dtype=torch.float64)
def frames(x):
    xp = torch.nn.functional.pad(x.double()[:, None], (n // 2, n // 2), mode="reflect")[:, 0]
    return xp.unfold(1, n, hop)[:, :nfr].reshape(rows, n) * w
Fp = frames(pred).requires_grad_(); Ft = frames(true)
k = torch.arange(nb, dtype=torch.float64); t = torch.arange(n, dtype=torch.float64)
ang = 2 * torch.pi * torch.outer(t, k) / n
C, S = torch.cos(ang), -torch.sin(ang)
def mag(F): return ((F @ C) ** 2 + (F @ S) ** 2).sqrt()
Sp, St = mag(Fp), mag(Ft)
lossr = ((Sp - St).abs().mean() + ((St + 1e-8).log2() - (Sp + 1e-8).log2()).abs().mean())
(dF,) = torch.autograd.grad(lossr, Fp)
rel = lambda a, b: float((a - b).norm() / b.norm())
print("loss", float(loss), float(lossr))
print("frames rel", rel(fr_p, Fp.detach()), " s_true rel", rel(s_true, St))
cnt = rows * nb
print("dFr rel", rel(dfr / cnt, dF), " worst rows:", torch.topk(((dfr / cnt - dF) ** 2).sum(1), 5))
# G reference: d loss / d (re, im)
re, im = (Fp.detach() @ C), (Fp.detach() @ S)
gs = torch.sign(Sp - St) + torch.sign((Sp + 1e-8).log2() - (St + 1e-8).log2()) / ((Sp + 1e-8) * torch.log(torch.tensor(2.0, dtype=torch.float64)))
Gr = torch.stack([gs * re / Sp, gs * im / Sp], -1).reshape(rows, N).detach()
print("G rel", rel(G, Gr), " worst rows:", torch.topk(((G - Gr) ** 2).sum(1), 5).indices.tolist(), "nfr", nfr)
bad = ((G - Gr).abs() > 1e-3 * Gr.abs().max()).nonzero()
print("bad G entries", bad[:10].tolist(), len(bad))
r, c = 95, 385
b_ = c // 2
print("G ours", G[r, c - 1 : c + 1].tolist(), "ref", Gr[r, c - 1 : c + 1].tolist())
print("Sp", float(Sp[r, b_]), "St", float(St[r, b_]), "s_true ours", float(s_true[r, b_]), "re im", float(re[r, b_]), float(im[r, b_]), "gs", float(gs[r, b_]))
print("neighbours ours", G[r, c - 5 : c + 5].tolist()); print("neighbours ref ", Gr[r, c - 5 : c + 5].tolist())
