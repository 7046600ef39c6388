import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from golf_b200 import loss as GL
dev = "cuda:0"
g = torch.Generator().manual_seed(0)
pred = (0.05 * torch.randn(32, 47760, generator=g)).to(dev).requires_grad_()
true = (0.05 * torch.randn(32, 47760, generator=g)).to(dev)
for _ in range(3):
    pred.grad = None
    GL.mss_loss(pred, true, (509, 1021, 2053)).backward()
torch.cuda.synchronize()
