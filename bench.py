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
            step's controls (one packed block) and D2H of the waveform inside the timed region
  roofline  the GOLF-ss filter (chunk responses + cluster tail) on its algorithmic bytes, plus one
            entry per kernel of the pass, each timed alone with CUDA events on pre-allocated workspaces
  parity    one pass with an injected noise draw checked against the CPU oracle (N = 1)
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


def run_cpu(steps: int, warmup: int, batch: int, keep: dict = None):
    """the reference's CPU path, restated (oracle port): returns (samples/s, cores, seconds/pass).  keep: filled with the
    inputs and the output of one pass (bench's parity check of the GPU arm)"""
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
        out = cpu_decoder_pass(O, sets[i % 2], table, rk, noise)
    dt = (time.perf_counter() - t0) / steps
    if keep is not None:
        keep.update(inputs=sets[(steps - 1) % 2], noise=noise, out=out)
    return batch * T / dt, cores, dt


def cpu_line(args):
    val, cores, dt = run_cpu(args.steps, max(args.warmup, 1), BATCH)
    sample = f"{args.steps} full decoder passes of {BATCH} x {SECONDS:g} s ({dt:.3f} s each) after {max(args.warmup, 1)} warm-up"
    cfg = workload_config(1)
    cfg.update(parallelism=f"{cores} host threads (OpenMP over utterances + torch intra-op), one process (rank 0 of {args.gpus})",
               launch="CPU: oracle port of the reference decoder, noise drawn per pass with torch.randn",
               l2="n/a (host)", noise="torch.randn per pass")
    return {
        "impl": "reference", "metric": "audio samples/sec, GOLF-ss synthesis 24 kHz batch 32 x 2 s", "value": val,
        "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference = Python + torchlpc/kazane (absent, not installable): timed arm is the oracle port of its CPU path; "
                f"it runs {BATCH} utterances per step on rank 0 whatever --gpus says",
    }


NOISE_MODE = os.environ.get("GOLF_BENCH_NOISE", "fused")  # "fused": in-kernel Philox generator; "torch": torch.randn tensor


def workload_config(n):
    return {"workload": "GOLF-ss decoder forward (SourceFilterSynth, cfg/ae/decoder/golf-precise.yaml), training-mode inputs",
            "batch_per_gpu": BATCH, "global_batch": BATCH * n, "seconds": SECONDS, "sample_rate": SR, "hop": HOP,
            "lpc_order": ORDER, "n_mag": N_MAG, "oversampling": OS, "room_taps": 128, "parallelism": f"batch-shard x{n}, no collective",
            "l2": f"{N_SETS} rotating input sets (~{N_SETS * 26} MB) > 126 MB L2",
            "noise": "white noise drawn inside the noise-FIR kernel (Philox4x32-10 + Box-Muller, StandardNormalNoise(fused=True))"
                     if NOISE_MODE == "fused" else "torch.randn tensor per pass (StandardNormalNoise())",
            "launch": "CUDA-graph replay of decoder(**params) (golf_b200.graphs.GraphedSynth -> golf_synth_fused_fwd), several passes in flight on alternating streams (ReplayRing, see in_flight); host-side input-range asserts are not part of a replay"}


# ----------------------------------------------------------------------------- GPU arm
def build_decoder(dev, variant: str = "ss", fused_noise: bool = False):
    from golf_b200 import filters, noise, sf, synth

    end = (filters.LTVMinimumPhaseFilterPrecise(lpc_order=ORDER, lpc_parameterisation="rc2lpc") if variant == "ss"
           else filters.LTVMinimumPhaseFilter(window="hanning", window_length=4 * HOP, lpc_order=ORDER, lpc_parameterisation="rc2lpc"))

    dec = sf.SourceFilterSynth(
        synth.DownsampledIndexedGlottalFlowTable(hop_rate=10, in_channels=64, oversampling=OS, equal_energy=True,
                                                 table_type="derivative", normalize_method="constant_power", align_peak=True,
                                                 trainable=False, min_R_d=0.3, max_R_d=2.7, lf_v2=True, points=2048),
        noise.StandardNormalNoise(fused=fused_noise), filters.LTVZeroPhaseFIRFilter("hanning", conv_method="direct", n_mag=N_MAG),
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


def torchaudio_ff_filter_cuda(ex, gain, a, hop, diag_kernel):
    """The reference's own GPU path for the GOLF-ff end filter, op for op (models/filters.py:141-180 + models/lpc.py:11-16):
    linearly up-sampled gain, zero padding, unfold, torchaudio.functional.lfilter (libtorchaudio's iir_cu_kernel on CUDA,
    the only Blackwell kernel the reference reaches on this path), overlap-add by conv_transpose1d with a dense diag(window)
    kernel, normalisation.  Library code only -- the comparator of the `golf_ff.vs_torchaudio_cuda` key."""
    import torch.nn.functional as Fn
    from torchaudio.functional import lfilter

    B, Tn = ex.shape
    Fr, win = gain.shape[1], diag_kernel.shape[0]
    up = Fn.interpolate(gain[:, None], (Fr - 1) * hop + 1, mode="linear", align_corners=True)[:, 0]
    L = min(Tn, up.shape[1])
    e = Fn.pad(ex[:, :L] * up[:, :L], (win // 2, win // 2))
    fr = e.unfold(1, win, hop)
    n = fr.shape[1]
    A = torch.cat([torch.ones_like(a[:, :n, :1]), a[:, :n]], -1).reshape(B * n, -1)
    Bc = torch.zeros_like(A)
    Bc[:, 0] = 1
    filt = lfilter(fr.reshape(B * n, win), A, Bc, False).view(B, n, win).transpose(1, 2)
    tmp = Fn.conv_transpose1d(torch.cat([filt, filt.new_ones(1, win, n)], 0), diag_kernel, stride=hop, padding=win // 2).squeeze(1)
    return tmp[:-1] / tmp[-1]


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

    fused_noise = NOISE_MODE == "fused"
    if "GOLF_BENCH_TAIL" in os.environ:  # A/B: 0 = separate stitch / solve launches, 1 = cluster tail, 2 = by batch size (default)
        golf_b200_lib().golf_lpc_ss_set_tail(int(os.environ["GOLF_BENCH_TAIL"]))
    dec = build_decoder(dev, fused_noise=fused_noise)
    raw_sets = make_inputs(N_SETS, BATCH, seed=2434 + rank)
    dev_sets = [{k: v.to(dev) for k, v in s.items()} for s in raw_sets]

    def params_of(s):
        return dict(phase=AudioTensor(s["phase"], hop_length=1), harm_oscillator_params=(AudioTensor(s["w"], hop_length=2400),),
                    noise_generator_params=(), noise_filter_params=(AudioTensor(s["log_mag"], hop_length=HOP),),
                    end_filter_params=(AudioTensor(s["gain"], hop_length=HOP), AudioTensor(s["a"], hop_length=HOP)))

    # one captured CUDA graph per resident input set (golf_b200.graphs.GraphedSynth, public API):
    # a replay is one launch, so the timed region measures the GPU, not Python's enqueue rate
    from golf_b200.graphs import GraphedSynth, PipelinedSynth, ReplayRing

    # decoder passes in flight (1, 2, 4 or 8); the host-to-host pipeline is bound by the H2D copy from two passes on
    IN_FLIGHT = int(os.environ.get("GOLF_BENCH_IN_FLIGHT", "8"))  # measured on B200: 2 -> 6.23e9, 4 -> 7.3e9, 8 -> 7.46e9 samples/s
    DEPTH, E2E_STREAMS = int(os.environ.get("GOLF_BENCH_DEPTH", "8")), int(os.environ.get("GOLF_BENCH_E2E_STREAMS", "4"))  # e2e pipeline slots / replay streams; measured on B200 (slots, streams): (4, 2) e2e 5.52e9 / frame-rate-f0 e2e 6.07e9, (8, 4) 5.54e9 / 7.12e9
    with torch.no_grad():
        graphed = [GraphedSynth(dec, params_of(s)) for s in dev_sets]
        pipe = PipelinedSynth(dec, params_of(dev_sets[0]), depth=DEPTH, compute_streams=E2E_STREAMS, packed=True)
    ring = ReplayRing(graphed, streams=IN_FLIGHT)
    out_host = [torch.empty(BATCH, pipe.out_len, dtype=torch.float32).pin_memory() for _ in range(DEPTH)]
    # host side of the e2e measurement: every input set lives in ONE pinned block with the layout of the pipeline's
    # packed static inputs (PipelinedSynth.host_staging), so a step's controls cross PCIe as a single copy
    host_flats = []
    for s in raw_sets:
        flat, views = pipe.host_staging()
        views["phase"].copy_(s["phase"])
        views["harm_oscillator_params"][0].copy_(s["w"])
        views["noise_filter_params"][0].copy_(s["log_mag"])
        views["end_filter_params"][0].copy_(s["gain"])
        views["end_filter_params"][1].copy_(s["a"])
        host_flats.append(flat)
    host_sets = [{k: v.pin_memory() for k, v in s.items()} for s in raw_sets[:2]]

    # the resident input sets already sit in their graphs' static inputs: a step is one graph replay; consecutive
    # steps alternate between IN_FLIGHT streams (golf_b200.graphs.ReplayRing, public API) so the serial tail of one
    # pass overlaps the next pass -- K passes are still K complete, independent decoder calls
    def step_dev(i):
        return ring.submit(i)

    # host (pinned) controls -> ONE H2D into a slot's graph inputs -> replay -> D2H of the waveform into pinned
    # host memory, EVERY step; consecutive steps overlap on three streams (PipelinedSynth, public API)
    def step_e2e(i):
        pipe.submit_flat(out_host[i % DEPTH], host_flats[i % N_SETS])
        return pipe.slots[0]._out

    def step_e2e_serial(i):  # the same without overlap: latency of one host-to-host call
        y = graphed[0](**params_of(host_sets[i % 2])).as_tensor()
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
    h2d = int(host_flats[0].numel() * 4)
    d2h = BATCH * n_out * 4

    # ---- copy-only ceiling of the e2e pipeline on this rank: the same H2D + D2H traffic with no kernels in between
    def copy_only(i):
        with torch.cuda.stream(pipe.s_in):
            pipe.slots[i % DEPTH]._flat.copy_(host_flats[i % N_SETS], non_blocking=True)
        with torch.cuda.stream(pipe.s_out):
            out_host[i % DEPTH].copy_(plain_out, non_blocking=True)

    plain_out = pipe.slots[0]._out.as_tensor()
    ms_copy, _, _ = timed(copy_only, args.steps, args.warmup, pipe.fork_from, pipe.join_into)

    # ---- second e2e line: frame-rate f0 (test_rtf.py:219-223: f0 at hop 120 instead of one value per sample)
    e2e_fr = None
    try:
        with torch.no_grad():
            def params_fr(sd):
                p_ = params_of(sd)
                p_["phase"] = AudioTensor(sd["phase"][:, ::120].contiguous(), hop_length=120)
                return p_

            pipe_fr = PipelinedSynth(dec, params_fr(dev_sets[0]), depth=DEPTH, compute_streams=E2E_STREAMS, packed=True)
            flats_fr = []
            for sd in raw_sets:
                flat, views = pipe_fr.host_staging()
                views["phase"].copy_(sd["phase"][:, ::120])
                views["harm_oscillator_params"][0].copy_(sd["w"])
                views["noise_filter_params"][0].copy_(sd["log_mag"])
                views["end_filter_params"][0].copy_(sd["gain"])
                views["end_filter_params"][1].copy_(sd["a"])
                flats_fr.append(flat)
            out_fr = [torch.empty(BATCH, pipe_fr.out_len, dtype=torch.float32).pin_memory() for _ in range(DEPTH)]
        ms_fr, _, _ = timed(lambda i: pipe_fr.submit_flat(out_fr[i % DEPTH], flats_fr[i % N_SETS]), args.steps, args.warmup,
                            pipe_fr.fork_from, pipe_fr.join_into)
        e2e_fr = {"value": total * args.steps / (ms_fr * 1e-3), "unit": "samples/s", "ms_per_step": ms_fr / args.steps,
                  "h2d_bytes_per_step": int(flats_fr[0].numel() * 4), "d2h_bytes_per_step": BATCH * pipe_fr.out_len * 4,
                  "what": "same pipeline with f0 at frame rate (hop 120, test_rtf.py:219-223) instead of one value per sample"}
        del pipe_fr
    except Exception as e:  # noqa: BLE001
        e2e_fr = {"error": f"{type(e).__name__}: {e}"[:200]}

    # ---- roofline: the GOLF-ss filter and every kernel of the pass, each timed alone on pre-allocated workspaces
    roof = None
    if rank == 0:
        roof = kernel_rooflines(dev, dec, dev_sets, G, AudioTensor)

    # ---- BASELINE.json configs[1]: the GOLF-ff decoder (frame-wise filter) on the same controls, device-resident,
    # reported beside the headline (never instead of it); a failure here must not cost the main line
    ff = None
    try:
        with torch.no_grad():
            dec_ff = build_decoder(dev, "ff")
            graphed_ff = [GraphedSynth(dec_ff, params_of(s)) for s in dev_sets]
            ring_ff = ReplayRing(graphed_ff, streams=IN_FLIGHT)
        ms_ff1, _, out_ff = timed(lambda i: graphed_ff[i % N_SETS].replay(), args.steps, args.warmup)
        ms_ff, _, _ = timed(lambda i: ring_ff.submit(i), args.steps, args.warmup, ring_ff.fork_from, ring_ff.join_into)
        ff = {"workload": "GOLF-ff decoder forward (cfg/ae/decoder/golf.yaml: LTVMinimumPhaseFilter, hanning 960), same controls",
              "value": total * args.steps / (ms_ff * 1e-3), "unit": "samples/s", "ms_per_step": ms_ff / args.steps,
              "ms_per_step_one_at_a_time": ms_ff1 / args.steps, "in_flight": IN_FLIGHT,
              "output_samples_per_utterance": int(out_ff.shape[1])}
        if rank == 0:
            ff["vs_torchaudio_cuda"] = ff_vs_torchaudio(dev, dev_sets, G)
        del graphed_ff, ring_ff
    except Exception as e:  # noqa: BLE001
        ff = {"error": f"{type(e).__name__}: {e}"[:200]}

    if rank != 0:
        return None
    line = {
        "metric": "audio samples/sec, GOLF-ss synthesis 24 kHz batch 32 x 2 s", "value": value, "unit": "samples/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(world), "clocks": clocks.summary(),
        "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps,
                "pipeline": f"{DEPTH} slots, one packed H2D / graph replay ({E2E_STREAMS} streams) / D2H",
                "serial_ms_per_step": ms_ser / args.steps,
                "copy_only_ms_per_step": ms_copy / args.steps,
                "copy_only_note": "the same H2D + D2H copies with no kernel between them (max over ranks): the ceiling PCIe / host memory set for this pipeline"},
        "e2e_frame_rate_f0": e2e_fr,
        "gpu_launches": int(launches), "kernels_per_pass": int(graphed[0].kernels_captured), "roofline": roof,
        "output_samples_per_utterance": int(n_out),
        "rtf": (ms / args.steps * 1e-3) / (BATCH * SECONDS), "golf_ff": ff,
        "in_flight": IN_FLIGHT, "ms_per_step_one_at_a_time": ms_seq / args.steps,
    }
    if world == 1 and not args.no_cpu:
        line["fit_step"] = fit_step_extra()
        keep = {}
        val, cores, dt = run_cpu(args.cpu_steps, 1, BATCH, keep)
        line["cpu_baseline"] = {"value": val, "unit": "samples/s", "cores": cores, "kind": "port",
                                "sample": f"{args.cpu_steps} full decoder passes of {BATCH} x {SECONDS:g} s on the host ({dt:.3f} s each), oracle port of the reference CPU path"}
        line["parity"] = parity_check(dev, keep, AudioTensor)
    if world > 1:
        dist.destroy_process_group()
    return line


def fit_step_extra():
    """BASELINE.json configs[3] (the autoencode.py fit step) as an extra key, N = 1: tools/fit_step.py in a child process,
    once with the tcgen05 multi-scale spectral loss (golf_b200.loss) and once with the torch.stft / cuFFT restatement of the
    reference's loss.  The multi-GPU runs of the same script are under profiles/.  Never costs the main line."""
    out = {"what": "VoiceAutoEncoder.training_step semantics (ltng/ae.py:86-143) with a 6.09 M-parameter stand-in encoder: encoder -> "
                   ".ctrl -> golf_b200 decoder -> MSS loss -> CUDA adjoints -> clip 0.5 -> Adam, 32 x 2 s, eager (tools/fit_step.py)"}
    for key, impl in (("golf_mss_loss", "tcgen05"), ("torch_stft_loss", "torch")):
        try:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fit_step.py"), "10", "ss", impl], capture_output=True,
                               text=True, timeout=300, env={**os.environ, "RANK": "0", "WORLD_SIZE": "1", "LOCAL_RANK": "0"})
            rec = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
            out[key] = {"ms_per_step": rec["ms_per_step"], "samples_per_s": rec["samples_per_s"], "split_ms": rec["split_ms"]}
        except Exception as e:  # noqa: BLE001
            out[key] = {"error": f"{type(e).__name__}: {e}"[:200]}
    return out


def parity_check(dev, keep, AudioTensor):
    """the benched decoder on the inputs and the noise draw of the CPU arm's last pass, compared sample by sample with
    that pass's output (oracle as the checker): max over utterances of rms(gpu - cpu) / rms(cpu)"""
    from golf_b200 import noise as gnoise

    s, noise, ref = keep["inputs"], keep["noise"].to(dev), keep["out"]

    class Injected(gnoise.NoiseInterface):
        def forward(self, ref_, *a):
            return AudioTensor(noise[:, : ref_.shape[1]])

    dec = build_decoder(dev)
    dec.noise_generator = Injected()
    dec.harm_oscillator.phase_accumulation = "aten_cpu"  # the reference's float32 running phase (oracle arithmetic)
    A = lambda t, hop: AudioTensor(t.to(dev), hop_length=hop)
    with torch.no_grad():
        out = dec(phase=A(s["phase"], 1), harm_oscillator_params=(A(s["w"], 2400),), noise_generator_params=(),
                  noise_filter_params=(A(s["log_mag"], HOP),), end_filter_params=(A(s["gain"], HOP), A(s["a"], HOP))).as_tensor().cpu()
    if out.shape != ref.shape:
        return {"error": f"shape {tuple(out.shape)} vs oracle {tuple(ref.shape)}"}
    d = ((out.double() - ref.double()) ** 2).mean(1).sqrt() / (ref.double() ** 2).mean(1).sqrt()
    return {"rel_rms_max": float(d.max()), "rel_rms_median": float(d.median()), "tolerance": 1e-4, "pass": bool(d.max() < 1e-4),
            "checked": f"{out.shape[0]} x {out.shape[1]} samples of one decoder pass (modules of the benched decoder, injected noise draw, "
                       "reference phase arithmetic) against the CPU oracle's output for the same inputs",
            "unpinned_third_party": ["torchlpc.sample_wise_lpc (restated 3x independently, tests/test_oracle.py)", "kazane.Decimate (restated)"]}


def _time_loop(fn, n=20, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def kernel_rooflines(dev, dec, dev_sets, G, AudioTensor):
    """Per-kernel and filter-level roofline entries.  Each kernel runs alone (CUDA events around 20 launches, rotating
    input sets, workspaces allocated once outside the loop); `achieved` = the kernel's OWN algorithmic bytes / time.
    DRAM traffic per launch comes from the committed ncu capture (profiles/ncu_summary.json), null when absent."""
    peak, src_peak = 6650.0, "fallback"
    try:
        peak, src_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        pass
    try:
        ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))
    except Exception:
        ncu = {}
    lib = golf_b200_lib()
    Ls = T - HOP  # samples the filter emits per utterance
    ctl = 4 * (ORDER + 1) / HOP  # bytes per sample of frame-rate (gain, a)
    src = [torch.randn(BATCH, Ls, device=dev) * 0.05 for _ in range(N_SETS)]
    harm = torch.randn(BATCH, T, device=dev) * 0.05
    noise = torch.randn(BATCH, T, device=dev)
    win = torch.hann_window(2 * (N_MAG - 1), device=dev)
    rk = dec.room_filter.kernel.detach()
    osc = dec.harm_oscillator
    ws = G._workspace(lib.golf_lpc_ss_room_workspace_bytes(BATCH, Ls, ORDER, HOP, 0), dev)
    y = torch.empty(BATCH, Ls, device=dev)
    out = torch.empty(BATCH, Ls, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def lpc(i, passes, room):
        sd = dev_sets[i % N_SETS]
        if room:
            rc = lib.golf_lpc_ss_room_fwd(src[i % N_SETS].data_ptr(), Ls, sd["gain"].data_ptr(), sd["a"].data_ptr(), 0, rk.data_ptr(), rk.numel(),
                                          0, out.data_ptr(), BATCH, Ls, FRAMES, ORDER, HOP, 0, 1, ws.data_ptr(), ws.numel(), st)
        else:
            rc = lib.golf_lpc_ss_fwd_passes(src[i % N_SETS].data_ptr(), Ls, sd["gain"].data_ptr(), sd["a"].data_ptr(), 0, y.data_ptr(), BATCH, Ls,
                                            FRAMES, ORDER, HOP, 0, ws.data_ptr(), ws.numel(), passes, st)
        assert rc == 0, rc

    rng = G.new_rng_state(dev, 1)
    entries = []

    def add(name, kernels, fn, alg_bytes, flops=None, note=None):
        with torch.no_grad():
            ms = _time_loop(fn)
        ach = alg_bytes / (ms * 1e-3) / 1e9
        traffic = sum(ncu[k]["dram_bytes_per_launch"] for k in kernels if k in ncu) if all(k in ncu for k in kernels) else None
        e = {"name": name, "kernels": kernels, "ms": ms, "algorithmic_bytes": alg_bytes, "achieved_GBps": ach, "frac_hbm": ach / peak,
             "traffic": traffic}
        if flops:
            e["fp32_tflops"] = flops / (ms * 1e-3) / 1e12
            e["frac_fp32"] = e["fp32_tflops"] / FP32_PEAK_TFLOPS
        if note:
            e["note"] = note
        entries.append(e)
        return e

    n = BATCH * Ls
    add("GOLF-ss chunk responses (pass 1)", ["ss_response_kernel"], lambda i: lpc(i, 1, False), (4 + ctl) * n,
        flops=2.0 * ORDER * (ORDER + 1) * n, note="reads ex + controls; its output (chunk transition blocks, 15 MB) is an L2-resident intermediate")
    lpc(0, 1, False)
    tail_kernels = ["ss_tail_kernel"] if golf_b200_lib().golf_lpc_ss_get_tail() else ["ss_stitch_kernel", "ss_solve_sys_kernel"]
    add("GOLF-ss tail: stitch + solve + adaptive refinement (passes 2-4" + (", one cluster launch)" if len(tail_kernels) == 1 else ", four light launches)"),
        tail_kernels, lambda i: lpc(i, 14, False),
        (8 + ctl) * n, flops=2.0 * ORDER * n, note="reads ex + controls, writes y; timed on the chunk blocks of one response pass")
    filt = add("GOLF-ss filter + room FIR (golf_lpc_ss_room_fwd: responses + tail + room)", ["ss_response_kernel"] + tail_kernels + ([] if len(tail_kernels) == 1 else ["room_fir_kernel"]),
               lambda i: lpc(i, 15, True), (8 + ctl) * n, flops=2.0 * (ORDER * (ORDER + 1) + ORDER + 128) * n)
    add("oscillator (knot prefix + flow / 4x decimation)", ["osc_knot_prefix_q64_kernel", "osc_flow_v2_kernel"],
        lambda i: G.glottal_osc(dev_sets[i % N_SETS]["phase"], 1, dev_sets[i % N_SETS]["w"], 2400, osc.table, osc.decimater.kernel,
                                osc.oversampling, osc.equal_energy, osc.phase_accumulation),
        8.0 * BATCH * T, note="reads phase (sample rate), writes harm")
    add("noise branch: FIR design + 510-tap block FIR + harm (noise tensor given)", ["noise_fir_design_kernel"],
        lambda i: G.noise_fir_design(noise, dev_sets[i % N_SETS]["log_mag"], win, HOP, add=harm), (12 + 4 * N_MAG / HOP) * n,
        flops=2.0 * (2 * (N_MAG - 1) + N_MAG * N_MAG / 2 / HOP) * n, note="reads noise + harm + log_mag, writes src")
    add("noise branch with the in-kernel generator", ["noise_fir_design_kernel"],
        lambda i: G.noise_fir_design(None, dev_sets[i % N_SETS]["log_mag"], win, HOP, add=harm, rng_state=rng), (8 + 4 * N_MAG / HOP) * n,
        flops=2.0 * (2 * (N_MAG - 1) + N_MAG * N_MAG / 2 / HOP) * n, note="reads harm + log_mag, writes src")
    roof = {"bound": "hbm", "kernel": filt["name"], "achieved": filt["achieved_GBps"], "peak": peak, "peak_source": src_peak, "unit": "GB/s",
            "frac": filt["frac_hbm"], "traffic": filt["traffic"], "algorithmic_bytes_per_launch": filt["algorithmic_bytes"],
            "kernel_ms": filt["ms"],
            "fp32": {"tflops": filt["fp32_tflops"], "peak_tflops": FP32_PEAK_TFLOPS, "frac": filt["frac_fp32"],
                     "peak_source": "148 SMs x 128 FMA/clk x 1.965 GHz x 2"},
            "note": "FP32-issue / latency bound by design (M(M+1) FMA per sample buys the time-parallel split; see DESIGN.md 3.1): "
                    "frac is against the HBM peak on SURVEY 8(d)'s 8.383 B/sample; traffic sums the kernels' DRAM bytes from the committed "
                    "`ncu --set full` capture, which flushes the caches before every launch -- in a pass's natural cache state the "
                    "15 MB of chunk blocks between the kernels stay in L2 and the same kernels move about 17 MB (profiles/r2_dram_per_pass.txt)",
            "kernels": entries}
    return roof


FP32_PEAK_TFLOPS = 148 * 128 * 1.965e9 * 2 / 1e12


def golf_b200_lib():
    from golf_b200 import _lib

    return _lib.lib()


def ff_vs_torchaudio(dev, dev_sets, G):
    """GOLF-ff end filter alone: golf_lpc_ff_fwd against the reference's own CUDA route (torchaudio lfilter + dense OLA)"""
    win = torch.hann_window(4 * HOP, device=dev)
    diag = torch.diag(win).unsqueeze(1)
    ex = torch.randn(BATCH, T - HOP, device=dev) * 0.05
    with torch.no_grad():
        ms_g = _time_loop(lambda i: G.lpc_ff(ex, dev_sets[i % N_SETS]["gain"], dev_sets[i % N_SETS]["a"], win, HOP), n=20)
        ms_t = _time_loop(lambda i: torchaudio_ff_filter_cuda(ex, dev_sets[i % N_SETS]["gain"], dev_sets[i % N_SETS]["a"], HOP, diag), n=5, warm=2)
        a_, b_ = G.lpc_ff(ex, dev_sets[0]["gain"], dev_sets[0]["a"], win, HOP), torchaudio_ff_filter_cuda(ex, dev_sets[0]["gain"], dev_sets[0]["a"], HOP, diag)
    d = ((a_.double() - b_.double()) ** 2).mean(1).sqrt() / (b_.double() ** 2).mean(1).sqrt()
    return {"golf_ms": ms_g, "torchaudio_cuda_ms": ms_t, "ratio": ms_t / ms_g, "rel_rms_max": float(d.max()),
            "what": f"GOLF-ff end filter, {BATCH} x {T - HOP} samples ({BATCH * FRAMES} frames of {4 * HOP}), M = {ORDER}: golf_lpc_ff_fwd vs "
                    "torchaudio.functional.lfilter on CUDA + unfold / conv_transpose1d exactly as models/filters.py:141-180"}


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
