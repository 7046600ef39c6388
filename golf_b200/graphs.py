"""CUDA-graph replay of a decoder call.

A GOLF decoder step is ~15 short kernels; launched eagerly from Python the host needs longer
to enqueue them than a B200 needs to run them.  `GraphedSynth` captures
`decoder(**params)` once (same shapes, static buffers) and replays it: new control tensors
are copied into the static inputs (device or pinned-host sources), one graph launch runs the
whole step, the result is the static output tensor (valid until the next call).

    gs = GraphedSynth(decoder, example_params)     # params as ltng/ae.py builds them
    y = gs(**params)                               # AudioTensor [B, T'] at hop 1

Inference only (torch.no_grad); the oscillator's host-side range asserts are skipped
(nothing may synchronise inside a capture) -- call the decoder eagerly once if you want them.
"""
from __future__ import annotations

from typing import Any, Dict

import torch
from torch.utils._pytree import tree_flatten, tree_unflatten

import itertools

from . import noise as _noise
from . import synth as _synth

_RNG_SLOTS = itertools.count(1)
from .audiotensor import hop_of, like, plain


class GraphedSynth:
    def __init__(self, decoder: torch.nn.Module, example_params: Dict[str, Any], warmup: int = 3, packed: bool = False):
        """packed: the static inputs are views into ONE flat float32 device buffer (each tensor starts on a 256-byte
        boundary), so a step's controls arrive with a single host-to-device copy from a pinned staging buffer of the
        same layout (`host_staging()` / `load_flat()`) instead of one copy per control tensor."""
        leaves, self._spec = tree_flatten(example_params)
        self._is_tensor = [isinstance(v, torch.Tensor) for v in leaves]
        dev = next(plain(v).device for v, t in zip(leaves, self._is_tensor) if t and plain(v).is_cuda)
        self._flat, self._offsets = None, None
        if packed:
            off, self._offsets = 0, []
            for v, t in zip(leaves, self._is_tensor):
                if t:
                    if plain(v).dtype != torch.float32:
                        raise ValueError("GraphedSynth(packed=True): every tensor input must be float32")
                    self._offsets.append(off)
                    off += (plain(v).numel() + 63) // 64 * 64
                else:
                    self._offsets.append(None)
            self._flat = torch.zeros(max(off, 64), dtype=torch.float32, device=dev)
            self._static = []
            for v, t, o in zip(leaves, self._is_tensor, self._offsets):
                if t:
                    view = self._flat[o : o + plain(v).numel()].view(plain(v).shape)
                    view.copy_(plain(v).detach())
                    self._static.append(view)
                else:
                    self._static.append(v)
        else:
            self._static = [plain(v).detach().to(dev, copy=True) if t else v for v, t in zip(leaves, self._is_tensor)]
        self._hops = [hop_of(v, None) if t else None for v, t in zip(leaves, self._is_tensor)]
        self._refs = [v if t else None for v, t in zip(leaves, self._is_tensor)]
        self.decoder = decoder
        self.graph = torch.cuda.CUDAGraph()
        checks, _synth.CHECK_INPUTS = _synth.CHECK_INPUTS, "off"
        slot, _noise.RNG_SLOT[0] = _noise.RNG_SLOT[0], next(_RNG_SLOTS)  # this capture's own in-kernel generator state
        try:
            with torch.no_grad():
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    for _ in range(max(warmup, 1)):  # first-use work (attribute sets, lazy buffers) happens here
                        decoder(**self._wrapped())
                torch.cuda.current_stream(dev).wait_stream(side)
                from ._lib import launch_count

                n0 = launch_count()
                with torch.cuda.graph(self.graph):
                    self._out = decoder(**self._wrapped())
                self.kernels_captured = launch_count() - n0  # golf_b200 kernels replayed per call
        finally:
            _synth.CHECK_INPUTS = checks
            _noise.RNG_SLOT[0] = slot

    def _wrapped(self):
        leaves = [like(r, s, h) if (t and h is not None) else s for s, t, h, r in zip(self._static, self._is_tensor, self._hops, self._refs)]
        return tree_unflatten(leaves, self._spec)

    def load(self, **params) -> int:
        """copy new control tensors (device or pinned host) into the static inputs, on the current
        stream; returns the bytes copied"""
        leaves, _ = tree_flatten(params)
        n = 0
        for dst, src, t in zip(self._static, leaves, self._is_tensor):
            if t:
                s = plain(src)
                if s.data_ptr() != dst.data_ptr():
                    dst.copy_(s, non_blocking=True)
                    n += dst.numel() * dst.element_size()
        return n

    def host_staging(self):
        """(flat, params): a pinned host buffer with the packed layout and a params pytree of views into it (same
        structure and hop lengths as the example) -- fill the views, then `load_flat(flat)` is one H2D copy"""
        if self._flat is None:
            raise ValueError("host_staging() needs GraphedSynth(packed=True)")
        flat = torch.zeros(self._flat.numel(), dtype=torch.float32).pin_memory()
        leaves = []
        for s_, t, h, r, o in zip(self._static, self._is_tensor, self._hops, self._refs, self._offsets):
            if t:
                view = flat[o : o + s_.numel()].view(s_.shape)
                leaves.append(like(r, view, h) if h is not None else view)
            else:
                leaves.append(s_)
        return flat, tree_unflatten(leaves, self._spec)

    def load_flat(self, host_flat: torch.Tensor) -> int:
        """one copy of the whole packed control block (pinned host or device) into the static inputs"""
        self._flat.copy_(host_flat, non_blocking=True)
        return self._flat.numel() * 4

    def replay(self):
        self.graph.replay()
        return self._out

    def __call__(self, **params):
        self.load(**params)
        return self.replay()


class PipelinedSynth:
    """Host-to-host synthesis as a software pipeline over `depth` GraphedSynth slots:

        copy-in stream : H2D(i+1)            pinned host controls -> slot's static inputs
        compute stream :          graph(i)   one CUDA-graph replay of the decoder
        copy-out stream:                D2H(i-1)   waveform -> pinned host buffer

    PCIe is full duplex and independent of the SMs, so in steady state a step costs
    max(H2D, compute, D2H) instead of their sum.  Slots are recycled in order; events keep a
    slot's inputs from being overwritten before its replay has run and its output from being
    overwritten before it has been copied out.

        pipe = PipelinedSynth(decoder, example_params, depth=3)
        t = pipe.submit(out_host, **host_params)     # enqueue only, returns a ticket
        pipe.wait(t)                                 # out_host ([B, pipe.out_len], pinned, contiguous) is valid
    """

    def __init__(self, decoder: torch.nn.Module, example_params: Dict[str, Any], depth: int = 3, compute_streams: int = 1,
                 packed: bool = False):
        """compute_streams > 1: consecutive replays alternate between that many streams, so the latency-bound
        tail of one decoder pass (serial stitch / solve, a few SMs busy) overlaps the throughput kernels of the
        next; depth must be a multiple of it (a slot always replays on the same stream).
        packed: slots take their controls as ONE packed block (GraphedSynth(packed=True)); `submit_flat`."""
        if depth % compute_streams:
            raise ValueError("PipelinedSynth: depth must be a multiple of compute_streams")
        self.slots = [GraphedSynth(decoder, example_params, packed=packed) for _ in range(depth)]
        out = plain(self.slots[0]._out)
        self.device = out.device
        self.out_len = out.shape[1]
        self.kernels_captured = self.slots[0].kernels_captured
        self.s_in, self.s_out = (torch.cuda.Stream(device=self.device) for _ in range(2))
        self.s_runs = [torch.cuda.Stream(device=self.device) for _ in range(compute_streams)]
        self.s_run = self.s_runs[0]
        self._ran = [None] * depth    # replay of the slot's previous use finished
        self._read = [None] * depth   # copy-out of the slot's previous use finished
        self._n = 0
        self.h2d_bytes = 0

    def fork_from(self, stream=None):
        """order the pipeline after everything already enqueued on `stream` (default: current)"""
        stream = stream or torch.cuda.current_stream(self.device)
        for s in (self.s_in, self.s_out, *self.s_runs):
            s.wait_stream(stream)

    def join_into(self, stream=None):
        stream = stream or torch.cuda.current_stream(self.device)
        for s in (self.s_in, self.s_out, *self.s_runs):
            stream.wait_stream(s)

    def host_staging(self):
        """a pinned staging buffer + views for one step's controls (packed mode), see GraphedSynth.host_staging"""
        return self.slots[0].host_staging()

    def submit_flat(self, out_host: torch.Tensor, host_flat: torch.Tensor) -> int:
        """packed mode: the step's controls are one pinned block -> one H2D copy"""
        return self._submit(out_host, lambda slot: slot.load_flat(host_flat))

    def submit(self, out_host: torch.Tensor, **host_params) -> int:
        return self._submit(out_host, lambda slot: slot.load(**host_params))

    def _submit(self, out_host: torch.Tensor, load) -> int:
        k = self._n % len(self.slots)
        slot = self.slots[k]
        with torch.cuda.stream(self.s_in):
            if self._ran[k] is not None:
                self.s_in.wait_event(self._ran[k])
            self.h2d_bytes = load(slot)
            loaded = torch.cuda.Event()
            loaded.record(self.s_in)
        s_run = self.s_runs[k % len(self.s_runs)]
        with torch.cuda.stream(s_run):
            s_run.wait_event(loaded)
            if self._read[k] is not None:
                s_run.wait_event(self._read[k])
            y = plain(slot.replay())
            self._ran[k] = torch.cuda.Event()
            self._ran[k].record(s_run)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self._ran[k])
            # out_host must be a CONTIGUOUS pinned [B, out_len] tensor: a strided host destination makes
            # torch stage the copy through pageable memory and block
            out_host.copy_(y, non_blocking=True)
            self._read[k] = torch.cuda.Event()
            self._read[k].record(self.s_out)
        self._n += 1
        return k

    def wait(self, ticket: int) -> None:
        self._read[ticket].synchronize()


class ReplayRing:
    """Device-resident throughput mode: a ring of GraphedSynth instances (one per resident input set) replayed
    round-robin on `streams` CUDA streams, so that `streams` decoder passes are in flight at a time -- the serial,
    latency-bound tail of one pass (stitch / solve: a few SMs) runs under the throughput kernels of the next.
    A graph always replays on the same stream (len(graphs) must be a multiple of `streams`), so a graph never
    overlaps itself and its static buffers are never in use twice.

        ring = ReplayRing(graphs, streams=2)
        ring.fork_from(); [ring.submit(i) for i in range(k)]; ring.join_into()
    """

    def __init__(self, graphs, streams: int = 2):
        if len(graphs) % streams:
            raise ValueError("ReplayRing: the number of graphs must be a multiple of the number of streams")
        self.graphs = list(graphs)
        self.device = plain(self.graphs[0]._out).device
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(streams)]

    def fork_from(self, stream=None):
        stream = stream or torch.cuda.current_stream(self.device)
        for s in self.streams:
            s.wait_stream(stream)

    def join_into(self, stream=None):
        stream = stream or torch.cuda.current_stream(self.device)
        for s in self.streams:
            stream.wait_stream(s)

    def submit(self, i: int):
        g = self.graphs[i % len(self.graphs)]
        with torch.cuda.stream(self.streams[i % len(self.streams)]):
            return g.replay()
