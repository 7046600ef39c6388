"""Time the oscillator adjoints (d_w alone, d_w + d_table) at the bench geometry: 32 x 2 s, sample-rate f0, os 4."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from golf_b200 import functional as G
dev = "cuda:0"
dec = bench.build_decoder(torch.device(dev), "ss")
osc = dec.harm_oscillator
s = bench.make_inputs(1, bench.BATCH, seed=1)[0]
ph = s["phase"].to(dev)
w = torch.rand(bench.BATCH, bench.T // 2400 + 1, device=dev)
dk = osc.decimater.kernel
out = {}
for name, tab_grad in (("d_w", False), ("d_w+d_table", True)):
    wg = w.clone().requires_grad_()
    tg = osc.table.detach().clone().requires_grad_(tab_grad)
    y = G.glottal_osc(ph, 1, wg, 2400, tg, dk, 4, True, "exact")
    up = torch.randn_like(y)
    ins = (wg, tg) if tab_grad else (wg,)
    for _ in range(3):
        torch.autograd.grad(y, ins, up, retain_graph=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        torch.autograd.grad(y, ins, up, retain_graph=True)
    e1.record()
    torch.cuda.synchronize()
    out[name + "_ms"] = e0.elapsed_time(e1) / 20
print(json.dumps(out))
