"""CUDA-event timing of the GOLF-ss decoder step and of its stages (bench shape, device-resident inputs).
usage: python tools/step_events.py [n]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from golf_b200 import functional as G, sf, synth as gsynth
from golf_b200.audiotensor import AudioTensor
from golf_b200.graphs import GraphedSynth
dev = torch.device("cuda:0")
dec = bench.build_decoder(dev)
gsynth.CHECK_INPUTS = "off"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
sets = [{k: v.to(dev) for k, v in s.items()} for s in bench.make_inputs(bench.N_SETS, bench.BATCH)]
A = lambda s, k, h: AudioTensor(s[k], hop_length=h)
P = lambda s: dict(phase=A(s, "phase", 1), harm_oscillator_params=(A(s, "w", 2400),), noise_generator_params=(),
                   noise_filter_params=(A(s, "log_mag", 240),), end_filter_params=(A(s, "gain", 240), A(s, "a", 240)))

def t(fn, n=n):
    with torch.no_grad():
        for i in range(5): fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n): fn(i)
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

for mode, split in (("off", 1), ("auto", 1), ("auto", 2), ("auto", 4), ("auto", 8), ("auto", 16)):
    sf.CONCURRENT, sf.SPLIT = mode, split
    with torch.no_grad():
        gs = [GraphedSynth(dec, P(s)) for s in sets]
    print(f"graph replay, CONCURRENT={mode} SPLIT={split}: {t(lambda i: gs[i % len(gs)].replay()):8.1f} us  ({gs[0].kernels_captured} kernels)")
    del gs
sf.CONCURRENT, sf.SPLIT = "auto", 1
# raw PCIe: pinned host <-> device copies of the bench's per-step payloads
hp = torch.empty(13289088 // 4).pin_memory(); dp = torch.empty_like(hp, device=dev)
ho = torch.empty(6113280 // 4).pin_memory(); do = torch.empty_like(ho, device=dev)
print(f"H2D 13.3 MB pinned: {t(lambda i: dp.copy_(hp, non_blocking=True)):8.1f} us   D2H 6.1 MB pinned: {t(lambda i: ho.copy_(do, non_blocking=True)):8.1f} us")
s = sets[0]
with torch.no_grad():
    harm = dec.harm_oscillator(A(s, "phase", 1), A(s, "w", 2400))
    noise = dec.noise_generator(harm)
    src = dec.noise_filter(noise, A(s, "log_mag", 240), add=harm)
    y = dec.end_filter(src, A(s, "gain", 240), A(s, "a", 240))
    raw = dec.noise_filter.raw_kernels(A(s, "log_mag", 240))

def graphed(fn):
    with torch.no_grad():
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn(); fn()
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
    return lambda i: g.replay()

srct, L = src.as_tensor(), y.shape[1]
win510 = torch.hann_window(510, device=dev)
rng = G.new_rng_state(dev, 1)
stages = [
    ("oscillator", lambda: dec.harm_oscillator(A(s, "phase", 1), A(s, "w", 2400))),
    ("randn", lambda: dec.noise_generator(harm)),
    ("exp+irfft", lambda: dec.noise_filter.raw_kernels(A(s, "log_mag", 240))),
    ("noise FIR (+harm)", lambda: dec.noise_filter.apply_raw(noise, raw, 240, add=harm)),
    ("noise FIR + in-kernel design", lambda: dec.noise_filter(noise, A(s, "log_mag", 240), add=harm)),
    ("... + in-kernel noise", lambda: G.noise_fir_design(None, s["log_mag"], win510, 240, add=harm.as_tensor(), rng_state=rng)),
    ("lpc_ss all (15)", lambda: G._lpc_ss_fwd(srct, s["gain"], s["a"], None, 240, 0, 15)),
    ("lpc_ss no refine (7)", lambda: G._lpc_ss_fwd(srct, s["gain"], s["a"], None, 240, 0, 7)),
    ("lpc_ss responses+z (1)", lambda: G._lpc_ss_fwd(srct, s["gain"], s["a"], None, 240, 0, 1)),
    ("lpc_ss Phi only", lambda: G.lpc_ss_responses(s["a"], L, 240)),
    ("room FIR", lambda: dec.room_filter(y)),
]
ws = G.lpc_ss_responses(s["a"], L, 240)
stages += [
    ("lpc_ss z-solve (16)", lambda: G._lpc_ss_fwd(srct, s["gain"], s["a"], None, 240, 0, 16, ws=ws)),
    ("lpc_ss stitch (2)", lambda: G._lpc_ss_fwd(srct, s["gain"], s["a"], None, 240, 0, 2, ws=ws)),
    ("lpc_ss solve (4)", lambda: G._lpc_ss_fwd(srct, s["gain"], s["a"], None, 240, 0, 4, ws=ws)),
    ("lpc_ss finish (30)", lambda: G._lpc_ss_fwd(srct, s["gain"], s["a"], None, 240, 0, 30, ws=ws)),
    ("lpc_ss tail only (14)", lambda: G._lpc_ss_fwd(srct, s["gain"], s["a"], None, 240, 0, 14, ws=ws)),
    ("lpc_ss tail no refine (6)", lambda: G._lpc_ss_fwd(srct, s["gain"], s["a"], None, 240, 0, 6, ws=ws)),
    ("lpc_ss + room fused", lambda: G._lpc_ss_room_fwd(srct, s["gain"], s["a"], None, dec.room_filter.kernel, 240)),
]
for name, fn in stages:
    print(f"{name:28s} {t(graphed(fn)):8.1f} us")
