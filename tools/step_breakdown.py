"""Host/device breakdown of one GOLF-ss decoder step (bench shape)."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from golf_b200 import synth as gsynth, functional as G
from golf_b200.audiotensor import AudioTensor
dev = torch.device("cuda:0")
dec = bench.build_decoder(dev)
gsynth.CHECK_INPUTS = "off"
s = {k: v.to(dev) for k, v in bench.make_inputs(1, bench.BATCH)[0].items()}
A = lambda k, h: AudioTensor(s[k], hop_length=h)
def full():
    return dec(phase=A("phase", 1), harm_oscillator_params=(A("w", 2400),), noise_generator_params=(),
               noise_filter_params=(A("log_mag", 240),), end_filter_params=(A("gain", 240), A("a", 240)))
def t(fn, n=30):
    with torch.no_grad():
        for _ in range(5): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(n): fn()
        t_enq = (time.perf_counter() - t0) / n
        torch.cuda.synchronize(); t_tot = (time.perf_counter() - t0) / n
    return t_enq * 1e6, t_tot * 1e6
print("full step: enqueue %.0f us, total %.0f us" % t(full))
with torch.no_grad():
    harm = dec.harm_oscillator(A("phase", 1), A("w", 2400))
    noise = dec.noise_generator(harm)
    src = dec.noise_filter(noise, A("log_mag", 240), add=harm)
    y = dec.end_filter(src, A("gain", 240), A("a", 240))
print("osc       : enqueue %.0f us, total %.0f us" % t(lambda: dec.harm_oscillator(A("phase", 1), A("w", 2400))))
print("noise gen : enqueue %.0f us, total %.0f us" % t(lambda: dec.noise_generator(harm)))
print("noise fir : enqueue %.0f us, total %.0f us" % t(lambda: dec.noise_filter(noise, A("log_mag", 240), add=harm)))
print("lpc ss    : enqueue %.0f us, total %.0f us" % t(lambda: dec.end_filter(src, A("gain", 240), A("a", 240))))
print("room      : enqueue %.0f us, total %.0f us" % t(lambda: dec.room_filter(y)))
print("G.lpc_ss  : enqueue %.0f us, total %.0f us" % t(lambda: G.lpc_ss(src.as_tensor(), s["gain"], s["a"], 240)))
print("G.lpc_ss norefine: enqueue %.0f us, total %.0f us" % t(lambda: G.lpc_ss(src.as_tensor(), s["gain"], s["a"], 240, refine=False)))
