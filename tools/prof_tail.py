"""Driver for ncu / the instrumented build: the GOLF-ss filter (+ fused room FIR) alone at the bench shape.
usage: python tools/prof_tail.py [n] [batch]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from golf_b200 import functional as G
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else bench.BATCH
s = {k: v[:B].to(dev) for k, v in bench.make_inputs(1, max(B, 1))[0].items()}
src = torch.randn(B, bench.T - bench.HOP, device=dev) * 0.1
k = bench.room_kernel().to(dev)
with torch.no_grad():
    for i in range(n):
        G._lpc_ss_room_fwd(src, s["gain"], s["a"], None, k, bench.HOP)
        torch.cuda.synchronize()
print("done")
