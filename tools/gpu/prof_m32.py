import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import synthetic_controls
from golf_b200 import functional as G
B, M, hop = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
gain, a = synthetic_controls(B, 48000 // hop + 1, M, seed=100 + M + hop)
ex = torch.randn(B, 48000, generator=torch.Generator().manual_seed(M))
exd, gd, ad = ex.cuda(), gain.cuda(), a.cuda()
for _ in range(3): G.lpc_ss(exd, gd, ad, hop)
torch.cuda.synchronize()
