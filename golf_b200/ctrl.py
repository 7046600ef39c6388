"""Control plumbing: how a decoder tells the encoder to split and transform its logits.

Protocol (identical to the reference's, models/ctrl.py:17-69, so modules from either
side can sit in the same Synth): every controllable module has an attribute
`ctrl(next_fn) -> fn(split_sizes, trsfm_fns)`, which appends the module's own
(split_size, trsfm_fn) and forwards to `next_fn`.  `Synth.split_sizes_and_trsfms` chains
the children in REGISTRATION order and returns (split_sizes, trsfm_fns, arg_names) with
arg_names = "<child>_params".
"""
from __future__ import annotations

from typing import Callable, Tuple

import torch

from . import _interop

if _interop.INTEROP:
    Controllable = _interop.ref_ctrl.Controllable
    PassThrough = _interop.ref_ctrl.PassThrough
    Synth = _interop.ref_ctrl.Synth
    wrap_ctrl_fn = _interop.ref_ctrl.wrap_ctrl_fn
else:

    def _identity(split_sizes, trsfm_fns):
        return split_sizes, trsfm_fns

    def wrap_ctrl_fn(split_size: Tuple[int, ...] = (), trsfm_fn: Callable = lambda *x: ()):
        def ctrl(next_fn):
            def split_and_trsfm(split_sizes, trsfm_fns):
                return next_fn(split_sizes + (split_size,), trsfm_fns + (trsfm_fn,))

            return split_and_trsfm

        return ctrl

    class Controllable(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.ctrl = wrap_ctrl_fn()

    class PassThrough(Controllable):
        def forward(self, x, *args, **kwargs):
            return x

    class Synth(torch.nn.Module):
        @property
        def split_sizes_and_trsfms(self):
            kids = [(n, m) for n, m in self.named_children() if isinstance(m, Controllable)]
            fn = _identity
            for _, m in reversed(kids):
                fn = m.ctrl(fn)
            sizes, trsfms = fn((), ())
            return sizes, trsfms, tuple(n + "_params" for n, _ in kids)
