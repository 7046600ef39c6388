"""Config 5 of BASELINE.json: real-time-factor sweep of the decoder (test_rtf.py:240-248 semantics:
decoder only, 10 runs, drop min and max, mean -- but timed with CUDA events and explicit syncs),
B in {1,8,32,128}, hop in {120,240} (window = 4*hop), LPC order in {12,20,32}; f0 = 150 Hz constant
at hop 120 as in test_rtf.py:219-223; 2 s of audio.  Also filter-only timings and the CPU oracle
port for the same shapes (bounded: B <= 32).

    python tools/rtf_sweep.py [--cpu] > profiles/r1_rtf_sweep.json
"""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import synthetic_controls, smooth
from golf_b200 import filters, noise, sf, synth, functional as G
from golf_b200 import synth as gsynth
from golf_b200.audiotensor import AudioTensor
from golf_b200.graphs import GraphedSynth

dev = torch.device("cuda:0")
SR, SEC = 24000, 2.0
T = int(SR * SEC)
with_cpu = "--cpu" in sys.argv
quick = "--quick" in sys.argv  # hop 240, order 22 only
if "GOLF_TAIL" in os.environ:  # A/B of the GOLF-ss schedules: 0 light stitch / solve launches, 1 cluster tail, 2 automatic
    from golf_b200 import _lib
    _lib.lib().golf_lpc_ss_set_tail(int(os.environ["GOLF_TAIL"]))


def trimmed_mean(xs):
    xs = sorted(xs)[1:-1]
    return sum(xs) / len(xs)


def time_gpu(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(n):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        out.append(s.elapsed_time(e) * 1e-3)
    return trimmed_mean(out)


def decoder(variant, hop, M):
    end = (filters.LTVMinimumPhaseFilterPrecise(lpc_order=M) if variant == "ss"
           else filters.LTVMinimumPhaseFilter(window="hanning", window_length=4 * hop, lpc_order=M))
    return sf.SourceFilterSynth(
        synth.DownsampledIndexedGlottalFlowTable(hop_rate=10, in_channels=64, oversampling=4, equal_energy=True, lf_v2=True, points=2048),
        noise.StandardNormalNoise(), filters.LTVZeroPhaseFIRFilter("hanning", n_mag=256), end,
        filters.LTIAcousticFilter(128, "fft"), subtract_harmonics=False).to(dev).eval()


rows = []
gsynth.CHECK_INPUTS = "off"
for hop in ((240,) if quick else (120, 240)):
    for M in ((22,) if quick else (12, 20, 32)):
        for B in (1, 8, 32, 128):
            Fr = T // hop + 1
            gain, a = synthetic_controls(B, Fr, M, seed=hop + M)
            gen = torch.Generator().manual_seed(B)
            log_mag = smooth(torch.randn(B, Fr, 256, generator=gen)) - 4
            w = torch.rand(B, T // (hop * 10) + 1, generator=gen)
            phase = torch.full((B, T // 120 + 1), 150.0 / SR)
            ex = torch.randn(B, T, generator=gen)
            row = {"hop": hop, "lpc_order": M, "batch": B}
            for variant in ("ss", "ff"):
                dec = decoder(variant, hop, M)
                params = dict(phase=AudioTensor(phase.to(dev), hop_length=120), harm_oscillator_params=(AudioTensor(w.to(dev), hop_length=hop * 10),),
                              noise_generator_params=(), noise_filter_params=(AudioTensor(log_mag.to(dev), hop_length=hop),),
                              end_filter_params=(AudioTensor(gain.to(dev), hop_length=hop), AudioTensor(a.to(dev), hop_length=hop)))
                with torch.no_grad():
                    t_eager = time_gpu(lambda: dec(**params))
                    gs = GraphedSynth(dec, params)
                    t_graph = time_gpu(lambda: gs(**params))
                row[f"{variant}_decoder_eager_ms"] = t_eager * 1e3
                row[f"{variant}_decoder_graph_ms"] = t_graph * 1e3
                row[f"{variant}_rtf"] = t_graph / (B * SEC)
                row[f"{variant}_samples_per_s"] = B * T / t_graph
            exd, gd, ad = ex.to(dev), gain.to(dev), a.to(dev)
            win = torch.hann_window(4 * hop, device=dev)
            with torch.no_grad():
                row["ss_filter_ms"] = time_gpu(lambda: G.lpc_ss(exd, gd, ad, hop)) * 1e3
                row["ff_filter_ms"] = time_gpu(lambda: G.lpc_ff(exd, gd, ad, win, hop)) * 1e3
            if with_cpu and B <= 32:
                from oracle import golf_oracle as O
                O.set_num_threads(os.cpu_count()); torch.set_num_threads(os.cpu_count())
                for name, fn in (("ss", lambda: O.lpc_ss(ex, gain, a, hop)), ("ff", lambda: O.lpc_ff(ex, gain, a, hop, 4 * hop))):
                    fn(); ts = []
                    for _ in range(5):
                        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
                    row[f"cpu_{name}_filter_ms"] = min(ts) * 1e3
                row["cpu_cores"] = os.cpu_count()
            rows.append(row)
            print(json.dumps(row), file=sys.stderr)
print(json.dumps({"seconds": SEC, "sample_rate": SR, "rows": rows}, indent=1))
