"""Driver for ncu: the GOLF-ss filter (+ fused room FIR) alone at the bench shape.  usage: python tools/prof_tail.py [n]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from golf_b200 import functional as G
dev = torch.device("cuda:0")
s = {k: v.to(dev) for k, v in bench.make_inputs(1, bench.BATCH)[0].items()}
src = torch.randn(bench.BATCH, bench.T - bench.HOP, device=dev) * 0.1
k = bench.room_kernel().to(dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
with torch.no_grad():
    for i in range(n):
        G._lpc_ss_room_fwd(src, s["gain"], s["a"], None, k, bench.HOP)
torch.cuda.synchronize()
print("done")
