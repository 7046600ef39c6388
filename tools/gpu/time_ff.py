"""Time the GOLF-ff end filter alone (golf_lpc_ff_fwd) at the bench geometry for the library GOLF_B200_SO selects."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from golf_b200 import _lib, functional as G

dev = "cuda:0"
out = {"so": os.path.basename(_lib.SO_PATH)}
for B in (1, 32):
    for hop, M in ((120, 22), (240, 12), (240, 22)):
        T = 47760
        F = T // hop + 2
        gen = torch.Generator().manual_seed(0)
        ex = torch.randn(B, T, generator=gen).to(dev)
        gain = torch.rand(B, F, generator=gen).to(dev) + 0.5
        a = G.rc2lpc((0.5 * torch.randn(B, F, M, generator=gen)).to(dev))
        win = torch.hann_window(4 * hop, periodic=False, device=dev)
        fn = lambda: G.lpc_ff(ex, gain, a, win, hop)
        for _ in range(5):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out[f"B{B}_hop{hop}_M{M}_us"] = round(e0.elapsed_time(e1) / 50 * 1e3, 1)
print(json.dumps(out))
# variant builds with -DGOLF_FF_TIMING: phase clocks of CTA 0 of the last launch
import ctypes
try:
    fn = _lib.lib()._dll.golf_debug_ff_clocks if hasattr(_lib.lib(), "_dll") else None
except Exception:  # noqa: BLE001
    fn = None
if fn is None:
    try:
        fn = ctypes.CDLL(_lib.SO_PATH).golf_debug_ff_clocks
    except AttributeError:
        fn = None
if fn is not None:
    buf = (ctypes.c_longlong * 8)()
    fn(buf)
    c = list(buf)
    print("phase cycles (CTA 0, last launch): stage copies %d, gain pass %d, coefficient load %d, serial %d, write-out %d" %
          (c[1] - c[0], c[2] - c[1], c[3] - c[2], c[4] - c[3], c[5] - c[4]))
