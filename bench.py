#!/usr/bin/env python
"""bench.py -- GOLF-ss synthesis throughput (audio samples/s) on N B200s, one process per GPU.

Workload (BASELINE.json metric / configs[2] at N=1): the full GOLF-ss decoder forward
(cfg/ae/decoder/golf-precise.yaml: DownsampledIndexedGlottalFlowTable 4x oversampled ->
StandardNormalNoise -> LTVZeroPhaseFIRFilter(n_mag 256) -> LTVMinimumPhaseFilterPrecise(M=22)
-> LTIAcousticFilter(128)), batch 32 x 2 s @ 24 kHz PER GPU (weak scaling; utterances are
independent, so ranks shard the batch and there is no data-path collective), training-mode
inputs (sample-rate f0), controls = the real encoder's output on the reference's sample
wavs (tests/golden/controls_gt.npz) tiled to the batch.

A step = one decoder pass over one batch.  Sample-count convention: B x 48 000
nominal samples per pass on both arms.

  value     whole-job samples/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e       same through the public nn.Module API with HOST (pinned) inputs: H2D of the
            step's controls and D2H of the waveform inside the timed region
  roofline  dominant kernel (GOLF-ss chunk-response pass), timed alone with CUDA events
  cpu_baseline / --impl reference
            the oracle port of the reference decoder (torch-CPU ops + OpenMP C recurrence,
            the shape of the reference's own CPU path) on this box's host cores

usage: python bench.py [--gpus N] [--steps K] [--warmup W] [--impl golf|reference]
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR, SECONDS, BATCH, HOP, ORDER, N_MAG, OS = 24000, 2.0, 32, 240, 22, 256, 4
T = int(SR * SECONDS)
FRAMES = T // HOP  # 200: training-mode control frames (models/unet.py:160-162)
N_SETS = 8         # rotating input sets: 8 x ~26 MB > 126 MB L2


# ----------------------------------------------------------------------------- inputs
def make_inputs(n_sets: int, batch: int, seed: int = 2434):
    """host float32 tensors for n_sets batches: phase [B,T] (cycles/sample, hop 1), w [B,21]@2400,
    log_mag [B,F,256], gain [B,F], a [B,F,M] @240."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "controls_gt.npz"))
    gen = torch.Generator().manual_seed(seed)
    sets = []
    n_utt = g["gain"].shape[0]
    for s in range(n_sets):
        idx = (torch.arange(batch) + s) % n_utt
        # f0: random walk in [80, 400] Hz at sample rate; 20 % of 0.1 s segments "unvoiced" ->
        # replaced by one U(50, 500) draw per item (ltng/ae.py:96-101)
        walk = torch.cumsum(torch.randn(batch, T // 240, generator=gen) * 4, 1) + torch.empty(batch, 1).uniform_(120, 300, generator=gen)
        f0 = torch.nn.functional.interpolate(walk.clamp(80, 400)[:, None], T, mode="linear", align_corners=True)[:, 0]
        unv = (torch.rand(batch, T // 2400, generator=gen) < 0.2).repeat_interleave(2400, 1)
        f0 = torch.where(unv, torch.empty(batch, 1).uniform_(50, 500, generator=gen).expand(-1, T), f0)
        sets.append(dict(
            phase=(f0 / SR).contiguous(),
            w=torch.tensor(g["w"])[idx].contiguous(),
            log_mag=torch.tensor(g["log_mag"])[idx, :FRAMES].contiguous(),
            gain=torch.tensor(g["gain"])[idx, :FRAMES].contiguous(),
            a=torch.tensor(g["a"])[idx, :FRAMES].contiguous(),
        ))
    return sets


def room_kernel():
    return torch.tensor(np.load(os.path.join(ROOT, "tests", "golden", "table.npz"))["room_kernel_ss"])


# ------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arm
def cpu_decoder_pass(O, s, table, rk, noise):
    return O.source_filter_synth(s["phase"], 1, s["w"], 2400, s["log_mag"], s["gain"], s["a"], HOP, noise, table, rk,
                                 variant="ss", oversampling=OS)


def run_cpu(steps: int, warmup: int, batch: int):
    """the reference's CPU path, restated (oracle port): returns (samples/s, cores, seconds/pass)"""
    from oracle import golf_oracle as O

    O.build()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    O.set_num_threads(cores)
    table, _ = O.glottal_table()
    rk = room_kernel()
    sets = make_inputs(2, batch)
    noise = torch.randn(batch, T)
    for i in range(warmup):
        cpu_decoder_pass(O, sets[i % 2], table, rk, noise)
    t0 = time.perf_counter()
    for i in range(steps):
        noise = torch.randn(batch, T)  # the reference draws noise inside the decoder
        cpu_decoder_pass(O, sets[i % 2], table, rk, noise)
    dt = (time.perf_counter() - t0) / steps
    return batch * T / dt, cores, dt


def cpu_line(args):
    val, cores, dt = run_cpu(args.steps, max(args.warmup, 1), BATCH)
    sample = f"{args.steps} full decoder passes of {BATCH} x {SECONDS:g} s ({dt:.3f} s each) after {max(args.warmup, 1)} warm-up"
    return {
        "impl": "reference", "metric": "audio samples/sec, GOLF-ss synthesis 24 kHz batch 32 x 2 s", "value": val,
        "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference = Python + torchlpc/kazane (absent, not installable): timed arm is the oracle port of its CPU path",
    }


def workload_config(n):
    return {"workload": "GOLF-ss decoder forward (SourceFilterSynth, cfg/ae/decoder/golf-precise.yaml), training-mode inputs",
            "batch_per_gpu": BATCH, "global_batch": BATCH * n, "seconds": SECONDS, "sample_rate": SR, "hop": HOP,
            "lpc_order": ORDER, "n_mag": N_MAG, "oversampling": OS, "room_taps": 128, "parallelism": f"batch-shard x{n}, no collective",
            "l2": f"{N_SETS} rotating input sets (~{N_SETS * 26} MB) > 126 MB L2",
            "launch": "CUDA-graph replay of decoder(**params) (golf_b200.graphs.GraphedSynth), several passes in flight on alternating streams (ReplayRing, see in_flight); host-side input-range asserts are not part of a replay"}


# ----------------------------------------------------------------------------- GPU arm
def build_decoder(dev, variant: str = "ss"):
    from golf_b200 import filters, noise, sf, synth

    end = (filters.LTVMinimumPhaseFilterPrecise(lpc_order=ORDER, lpc_parameterisation="rc2lpc") if variant == "ss"
           else filters.LTVMinimumPhaseFilter(window="hanning", window_length=4 * HOP, lpc_order=ORDER, lpc_parameterisation="rc2lpc"))

    dec = sf.SourceFilterSynth(
        synth.DownsampledIndexedGlottalFlowTable(hop_rate=10, in_channels=64, oversampling=OS, equal_energy=True,
                                                 table_type="derivative", normalize_method="constant_power", align_peak=True,
                                                 trainable=False, min_R_d=0.3, max_R_d=2.7, lf_v2=True, points=2048),
        noise.StandardNormalNoise(), filters.LTVZeroPhaseFIRFilter("hanning", conv_method="direct", n_mag=N_MAG),
        end, filters.LTIAcousticFilter(128, "fft"), subtract_harmonics=False)
    dec.room_filter.kernel.data = room_kernel()
    return dec.to(dev).eval()


def _bind_to_gpu_numa_node(index: int) -> None:
    """Pin this rank to the CPU cores next to its GPU before any pinned host buffer is allocated (first touch puts
    the pages on that NUMA node): with one rank per GPU the H2D / D2H copies of the e2e measurement then stay on
    the GPU's own socket instead of crossing the inter-socket link.  Best effort."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:  # noqa: BLE001  (no NVML, no permission, single-node box: nothing to do)
        pass


def run_gpu(args):
    import torch.distributed as dist

    import golf_b200
    from golf_b200 import functional as G
    from golf_b200.audiotensor import AudioTensor
    from golf_b200.sharding import max_over_ranks

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- golf_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:  # N = 1 keeps every host core for the cpu_baseline leg
        _bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(2434 + rank)

    dec = build_decoder(dev)
    host_sets = [{k: v.pin_memory() for k, v in s.items()} for s in make_inputs(N_SETS, BATCH, seed=2434 + rank)]
    dev_sets = [{k: v.to(dev) for k, v in s.items()} for s in host_sets]

    def params_of(s):
        return dict(phase=AudioTensor(s["phase"], hop_length=1), harm_oscillator_params=(AudioTensor(s["w"], hop_length=2400),),
                    noise_generator_params=(), noise_filter_params=(AudioTensor(s["log_mag"], hop_length=HOP),),
                    end_filter_params=(AudioTensor(s["gain"], hop_length=HOP), AudioTensor(s["a"], hop_length=HOP)))

    # one captured CUDA graph per resident input set (golf_b200.graphs.GraphedSynth, public API):
    # a replay is one launch, so the timed region measures the GPU, not Python's enqueue rate
    from golf_b200.graphs import GraphedSynth, PipelinedSynth

    # decoder passes in flight (1, 2, 4 or 8; measured on B200: 0.315 / 0.256 / 0.227 ms per step for 1 / 2 / 4);
    # the host-to-host pipeline is bound by the H2D copy from two passes on
    IN_FLIGHT = int(os.environ.get("GOLF_BENCH_IN_FLIGHT", "4"))
    DEPTH, E2E_STREAMS = 4, 2
    with torch.no_grad():
        graphed = [GraphedSynth(dec, params_of(s)) for s in dev_sets]
        pipe = PipelinedSynth(dec, params_of(dev_sets[0]), depth=DEPTH, compute_streams=E2E_STREAMS)
    from golf_b200.graphs import ReplayRing

    ring = ReplayRing(graphed, streams=IN_FLIGHT)
    out_host = [torch.empty(BATCH, pipe.out_len, dtype=torch.float32).pin_memory() for _ in range(DEPTH)]

    # the resident input sets already sit in their graphs' static inputs: a step is one graph replay; consecutive
    # steps alternate between IN_FLIGHT streams (golf_b200.graphs.ReplayRing, public API) so the serial tail of one
    # pass overlaps the next pass -- K passes are still K complete, independent decoder calls
    def step_dev(i):
        return ring.submit(i)

    # host (pinned) controls -> H2D into a slot's graph inputs -> replay -> D2H of the waveform into pinned
    # host memory, EVERY step; consecutive steps overlap on three streams (PipelinedSynth, public API)
    def step_e2e(i):
        pipe.submit(out_host[i % DEPTH], **params_of(host_sets[i % N_SETS]))
        return pipe.slots[0]._out

    def step_e2e_serial(i):  # the same without overlap: latency of one host-to-host call
        y = graphed[0](**params_of(host_sets[i % N_SETS])).as_tensor()
        out_host[0].copy_(y, non_blocking=True)
        return y

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, fork=None, join=None):
        with torch.no_grad():
            for i in range(warmup):
                fn(i)
            barrier()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            l0 = golf_b200.launch_count()
            ev[0].record()
            if fork:
                fork()  # the pipeline's streams start after ev[0] ...
            for i in range(steps):
                out = fn(warmup + i)
            if join:
                join()  # ... and ev[1] waits for all of them
            ev[1].record()
            barrier()
            ms = ev[0].elapsed_time(ev[1])
            launches = golf_b200.launch_count() - l0
        ms = max_over_ranks(ms, dev)  # the job is as slow as its slowest rank
        return ms, launches, out

    with ClockSampler(local) as clocks:
        ms, launches, out = timed(step_dev, args.steps, args.warmup, ring.fork_from, ring.join_into)
        ms_seq, _, _ = timed(lambda i: graphed[i % N_SETS].replay(), args.steps, args.warmup)  # one pass at a time
        ms_e2e, _, _ = timed(step_e2e, args.steps, args.warmup, pipe.fork_from, pipe.join_into)
        ms_ser, _, _ = timed(step_e2e_serial, args.steps, args.warmup)
    launches = args.steps * graphed[0].kernels_captured  # golf_b200 kernels replayed inside the timed region
    n_out = out.shape[1]
    total = world * BATCH * T
    value = total * args.steps / (ms * 1e-3)
    e2e = total * args.steps / (ms_e2e * 1e-3)
    h2d = sum(v.numel() * 4 for v in host_sets[0].values())
    d2h = BATCH * n_out * 4

    # ---- dominant kernel alone: GOLF-ss chunk-response pass
    roof = None
    if rank == 0:
        s = dev_sets[0]
        src = torch.randn(BATCH, T - HOP, device=dev)
        L = G.lpc_ss_length(src.shape[1], FRAMES, HOP)
        with torch.no_grad():
            for _ in range(3):
                G._lpc_ss_fwd(src, s["gain"], s["a"], None, HOP, 0, passes=1)
            torch.cuda.synchronize()
            n = 20
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(n):
                sd = dev_sets[i % N_SETS]
                G._lpc_ss_fwd(src, sd["gain"], sd["a"], None, HOP, 0, passes=1)
            e1.record()
            torch.cuda.synchronize()
        k_ms = e0.elapsed_time(e1) / n
        alg_bytes = (8 + 4 * (ORDER + 1) / HOP) * BATCH * L
        peak, src_peak = 6650.0, "fallback"
        try:
            peak, src_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
        except Exception:
            pass
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))["ss_response_kernel"]["dram_bytes_per_launch"]
        except Exception:
            pass
        ach = alg_bytes / (k_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "ss_response_kernel<24,0> (GOLF-ss pass 1, timed alone, includes the per-call workspace alloc)",
                "achieved": ach, "peak": peak, "peak_source": src_peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k_ms,
                "note": "FP32-issue/latency bound by design (M(M+1) FMA per sample for the time-parallel split), see DESIGN.md"}

    # ---- BASELINE.json configs[1]: the GOLF-ff decoder (frame-wise filter) on the same controls, device-resident,
    # reported beside the headline (never instead of it); a failure here must not cost the main line
    ff = None
    try:
        if world > 1:
            raise RuntimeError("N = 1 only (no collective inside an optional measurement)")
        with torch.no_grad():
            dec_ff = build_decoder(dev, "ff")
            graphed_ff = [GraphedSynth(dec_ff, params_of(s)) for s in dev_sets]
        ms_ff, _, out_ff = timed(lambda i: graphed_ff[i % N_SETS](**params_of(dev_sets[i % N_SETS])), args.steps, args.warmup)
        ff = {"workload": "GOLF-ff decoder forward (cfg/ae/decoder/golf.yaml: LTVMinimumPhaseFilter, hanning 960), same controls",
              "value": total * args.steps / (ms_ff * 1e-3), "unit": "samples/s", "ms_per_step": ms_ff / args.steps,
              "output_samples_per_utterance": int(out_ff.shape[1])}
    except Exception as e:  # noqa: BLE001
        ff = None if world > 1 else {"error": f"{type(e).__name__}: {e}"[:200]}

    if rank != 0:
        return None
    line = {
        "metric": "audio samples/sec, GOLF-ss synthesis 24 kHz batch 32 x 2 s", "value": value, "unit": "samples/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(world), "clocks": clocks.summary(),
        "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps, "pipeline": f"{DEPTH} slots, H2D / graph replay ({E2E_STREAMS} streams) / D2H",
                "serial_ms_per_step": ms_ser / args.steps},
        "gpu_launches": int(launches), "roofline": roof, "output_samples_per_utterance": int(n_out),
        "rtf": (ms / args.steps * 1e-3) / (BATCH * SECONDS), "golf_ff": ff,
        "in_flight": IN_FLIGHT, "ms_per_step_one_at_a_time": ms_seq / args.steps,
    }
    if world == 1 and not args.no_cpu:
        val, cores, dt = run_cpu(args.cpu_steps, 1, BATCH)
        line["cpu_baseline"] = {"value": val, "unit": "samples/s", "cores": cores, "kind": "port",
                                "sample": f"{args.cpu_steps} full decoder passes of {BATCH} x {SECONDS:g} s on the host ({dt:.3f} s each), oracle port of the reference CPU path"}
    if world > 1:
        dist.destroy_process_group()
    return line


def main():
    # stdout carries exactly ONE JSON line.  Libraries write there too (NCCL prints its version banner with
    # printf when NCCL_DEBUG is WARN/VERSION/INFO): keep a private handle on the real stdout for the result and
    # point file descriptor 1 at stderr for everything else.
    sys.stdout.flush()
    result_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="golf", choices=["golf", "reference"])
    ap.add_argument("--cpu-steps", type=int, default=10)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        if int(os.environ.get("RANK", 0)) != 0:
            return
        args.steps = min(args.steps, 20)  # ~0.7 s per pass on 8 cores: keep the arm within minutes
        print(json.dumps(cpu_line(args)), file=result_out, flush=True)
        return
    line = run_gpu(args)
    if line is not None:
        print(json.dumps(line), file=result_out, flush=True)


if __name__ == "__main__":
    main()
