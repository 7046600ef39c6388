"""Do parallel branches of a captured CUDA graph really overlap on this box?  Pairs of decoder kernels
captured (a) back to back on one stream, (b) on two forked streams; replay time of each."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from golf_b200 import functional as G, synth as gsynth
from golf_b200.audiotensor import AudioTensor
dev = torch.device("cuda:0")
dec = bench.build_decoder(dev)
gsynth.CHECK_INPUTS = "off"
s = {k: v.to(dev) for k, v in bench.make_inputs(1, bench.BATCH)[0].items()}
A = lambda k, h: AudioTensor(s[k], hop_length=h)
with torch.no_grad():
    harm = dec.harm_oscillator(A("phase", 1), A("w", 2400))
    noise = dec.noise_generator(harm)
    raw = dec.noise_filter.raw_kernels(A("log_mag", 240))
    src = dec.noise_filter.apply_raw(noise, raw, 240, add=harm)
srct = src.as_tensor()
L = G.lpc_ss_length(srct.shape[1], 200, 240)
ws1 = G.lpc_ss_responses(s["a"], L, 240)
ws2 = G.lpc_ss_responses(s["a"], L, 240)
G._lpc_ss_fwd(srct, s["gain"], s["a"], None, 240, 0, 15, ws=ws1)
G._lpc_ss_fwd(srct, s["gain"], s["a"], None, 240, 0, 15, ws=ws2)
K = {
    "stitch": lambda ws: G._lpc_ss_fwd(srct, s["gain"], s["a"], None, 240, 0, 2, ws=ws),
    "solve": lambda ws: G._lpc_ss_fwd(srct, s["gain"], s["a"], None, 240, 0, 4, ws=ws),
    "phi": lambda ws: G._lpc_ss_fwd(srct, s["gain"], s["a"], None, 240, 0, 1, ws=ws),
    "osc": lambda ws: dec.harm_oscillator(A("phase", 1), A("w", 2400)),
    "fir": lambda ws: dec.noise_filter.apply_raw(noise, raw, 240, add=harm),
    "sleep": lambda ws: torch.cuda._sleep(100000),
}

def capture(fa, fb, parallel):
    side, aux = torch.cuda.Stream(), torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.no_grad():
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fa(ws1); fb(ws2)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            cur = torch.cuda.current_stream()
            if parallel:
                aux.wait_stream(cur)
                with torch.cuda.stream(aux):
                    fb(ws2)
                fa(ws1)
                cur.wait_stream(aux)
            else:
                fa(ws1); fb(ws2)
    return g

def t(g, n=30):
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

for a, b in (("sleep", "sleep"), ("stitch", "stitch"), ("solve", "solve"), ("stitch", "phi"), ("solve", "phi"), ("osc", "phi"), ("osc", "fir"), ("osc", "solve"), ("fir", "stitch")):
    print(f"{a:7s} + {b:7s}: serial {t(capture(K[a], K[b], False)):7.1f} us   forked {t(capture(K[a], K[b], True)):7.1f} us")
