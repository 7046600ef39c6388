"""time the GOLF-ss passes with whatever library GOLF_B200_SO points to (tools/resp_nc_sweep.sh)"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from golf_b200 import functional as G
dev = torch.device("cuda:0")
s = {k: v.to(dev) for k, v in bench.make_inputs(1, bench.BATCH)[0].items()}
src = torch.randn(bench.BATCH, bench.T - bench.HOP, device=dev)
def t(f, n=30):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
ref = None
print(os.environ.get("GOLF_B200_SO", "default"),
      f"responses+z {t(lambda: G._lpc_ss_fwd(src, s['gain'], s['a'], None, bench.HOP, 0, passes=1)):.1f} us",
      f"all {t(lambda: G._lpc_ss_fwd(src, s['gain'], s['a'], None, bench.HOP, 0, passes=15)):.1f} us")
y = G.lpc_ss(src, s["gain"], s["a"], bench.HOP)
print("  checksum", float(y.double().pow(2).mean().sqrt()))
