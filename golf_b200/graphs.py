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

from . import synth as _synth
from .audiotensor import hop_of, like, plain


class GraphedSynth:
    def __init__(self, decoder: torch.nn.Module, example_params: Dict[str, Any], warmup: int = 3):
        leaves, self._spec = tree_flatten(example_params)
        self._is_tensor = [isinstance(v, torch.Tensor) for v in leaves]
        dev = next(plain(v).device for v, t in zip(leaves, self._is_tensor) if t and plain(v).is_cuda)
        self._static = [plain(v).detach().to(dev, copy=True) if t else v for v, t in zip(leaves, self._is_tensor)]
        self._hops = [hop_of(v, None) if t else None for v, t in zip(leaves, self._is_tensor)]
        self._refs = [v if t else None for v, t in zip(leaves, self._is_tensor)]
        self.decoder = decoder
        self.graph = torch.cuda.CUDAGraph()
        checks, _synth.CHECK_INPUTS = _synth.CHECK_INPUTS, "off"
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

    def _wrapped(self):
        leaves = [like(r, s, h) if (t and h is not None) else s for s, t, h, r in zip(self._static, self._is_tensor, self._hops, self._refs)]
        return tree_unflatten(leaves, self._spec)

    def __call__(self, **params):
        leaves, _ = tree_flatten(params)
        for dst, src, t in zip(self._static, leaves, self._is_tensor):
            if t:
                s = plain(src)
                if s.data_ptr() != dst.data_ptr():
                    dst.copy_(s, non_blocking=True)
        self.graph.replay()
        return self._out
