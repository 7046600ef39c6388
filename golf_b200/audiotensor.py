"""Rate-aware tensor: a torch.Tensor that knows how many audio samples one step of its
time axis (dim 1) spans.  Same contract as the reference's L1 type
(models/audiotensor/audiotensor.py:30-195, pinned by tests/test_time_tensor.py:18-28):

  * AudioTensor(data, hop_length=1); data.ndim >= 2, time on dim 1;
  * any torch function called with two or more AudioTensors first brings them to the
    gcd of their hops by LINEAR upsampling (F.interpolate, align_corners=True, length
    (n-1)*factor+1), right-pads missing trailing dims, and truncates all of them to the
    shortest; the result carries the (common) hop; 1-D results carry hop -1;
  * reduce_hop_length / increase_hop_length / set_hop_length / truncate / unfold /
    new_tensor / as_tensor / steps.

Inside the reference's process this module simply re-exports the reference's class (see
_interop.py) so isinstance checks on either side agree.  The hot-path kernels never go
through this arithmetic: they take frame-rate controls and interpolate in-kernel.
"""
from __future__ import annotations

from math import gcd
from typing import Sequence

import torch
import torch.nn.functional as F
from torch.utils._pytree import tree_flatten, tree_unflatten

from . import _interop


def linear_upsample(x: torch.Tensor, factor: int) -> torch.Tensor:
    """Last-dim linear interpolation onto (n-1)*factor+1 points (ends coincide)."""
    n = x.shape[-1]
    y = F.interpolate(x.reshape(-1, 1, n), size=(n - 1) * factor + 1, mode="linear", align_corners=True)
    return y.reshape(*x.shape[:-1], -1)


class _AudioTensor(torch.Tensor):
    hop_length: int

    @staticmethod
    def __new__(cls, data, hop_length: int = 1, requires_grad=None):
        t = data if isinstance(data, torch.Tensor) else torch.as_tensor(data)
        if requires_grad is None:
            return t.as_subclass(cls)
        return torch.Tensor._make_subclass(cls, t, requires_grad)

    def __init__(self, data, hop_length: int = 1, requires_grad=None):
        if self.ndim < 2:
            raise AssertionError("AudioTensor must have at least 2 dimensions")
        self.hop_length = int(hop_length)

    def __repr__(self):
        return f"AudioTensor(hop_length={getattr(self, 'hop_length', '?')}, {torch.Tensor.__repr__(self.as_tensor())})"

    # ---- plain views -------------------------------------------------------------
    def as_tensor(self) -> torch.Tensor:
        return self.as_subclass(torch.Tensor)

    def new_tensor(self, data: torch.Tensor) -> "_AudioTensor":
        return type(self)(data, hop_length=self.hop_length)

    def _need_rate(self, what: str):
        if self.hop_length < 0:
            raise ValueError(f"Cannot call {what} on an AudioTensor with hop_length < 0")

    @property
    def steps(self) -> int:
        self._need_rate("steps")
        return self.size(1) if self.ndim >= 2 else torch.iinfo(torch.int32).max

    def truncate(self, steps: int):
        self._need_rate("truncate")
        return self if (self.ndim < 2 or steps >= self.size(1)) else self.narrow(1, 0, steps)

    # ---- rate changes --------------------------------------------------------------
    def reduce_hop_length(self, factor: int = None):
        self._need_rate("reduce_hop_length")
        factor = self.hop_length if factor is None else factor
        assert factor >= 1 and self.hop_length % factor == 0, (self.hop_length, factor)
        if factor == 1:
            return self
        x = self.as_tensor()
        x = linear_upsample(x.transpose(1, -1), factor).transpose(1, -1) if x.ndim > 2 else linear_upsample(x, factor)
        return type(self)(x, hop_length=self.hop_length // factor)

    def increase_hop_length(self, factor: int):
        self._need_rate("increase_hop_length")
        assert factor > 0, "factor must be positive"
        if factor == 1:
            return self
        return type(self)(self.as_tensor()[:, ::factor].clone(), hop_length=self.hop_length * factor)

    def set_hop_length(self, hop_length: int):
        self._need_rate("set_hop_length")
        if hop_length > self.hop_length:
            assert hop_length % self.hop_length == 0
            return self.increase_hop_length(hop_length // self.hop_length)
        if hop_length < self.hop_length:
            assert self.hop_length % hop_length == 0
            return self.reduce_hop_length(self.hop_length // hop_length)
        return self

    def unfold(self, size: int, step: int = 1):
        self._need_rate("unfold")
        assert self.ndim == 2
        return type(self)(self.as_tensor().unfold(1, size, step), hop_length=self.hop_length * step)

    # ---- mixed-rate dispatch -------------------------------------------------------
    @classmethod
    def align(cls, tensors: Sequence["_AudioTensor"]):
        """common hop (gcd) -> equal rank -> equal number of steps"""
        g = gcd(*(t.hop_length for t in tensors))
        out = [t.reduce_hop_length(t.hop_length // g) for t in tensors]
        rank = max(t.ndim for t in out)
        out = [t.as_tensor().reshape(t.shape + (1,) * (rank - t.ndim)).as_subclass(cls) if t.ndim < rank else t for t in out]
        for o in out:
            o.hop_length = g
        n = min(t.size(1) for t in out)
        return [t.truncate(n) for t in out]

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if not kwargs and len(args) == 1:  # attribute getters and unary ops: nothing to align
            ret = super().__torch_function__(func, types, args, kwargs)
            if isinstance(ret, _AudioTensor):
                ret.hop_length = args[0].__dict__.get("hop_length", -1) if ret.ndim > 1 else -1
            return ret
        leaves, spec = tree_flatten((args, kwargs))
        rated = [i for i, v in enumerate(leaves) if isinstance(v, _AudioTensor) and getattr(v, "hop_length", -1) > 0]
        if len(rated) > 1:
            for i, v in zip(rated, cls.align([leaves[i] for i in rated])):
                leaves[i] = v
        hop = next((v.hop_length for v in leaves if isinstance(v, _AudioTensor) and getattr(v, "hop_length", -1) > 0), -1)
        a, k = tree_unflatten(leaves, spec)
        ret = super().__torch_function__(func, types, a, k)
        out_leaves, out_spec = tree_flatten(ret)
        for v in out_leaves:
            if isinstance(v, _AudioTensor):
                v.hop_length = hop if v.ndim > 1 else -1
        return tree_unflatten(out_leaves, out_spec)


AudioTensor = _interop.ref_audiotensor.AudioTensor if _interop.INTEROP else _AudioTensor


def hop_of(x, default: int = 1) -> int:
    return int(x.__dict__.get("hop_length", default)) if hasattr(x, "__dict__") else default


_NoTF = torch._C.DisableTorchFunctionSubclass


def plain(x: torch.Tensor) -> torch.Tensor:
    """the underlying torch.Tensor (keeps autograd history); no __torch_function__ round trip"""
    if type(x) is torch.Tensor:
        return x
    with _NoTF():
        return x.as_subclass(torch.Tensor)


def like(ref, data: torch.Tensor, hop_length: int = 1):
    """wrap `data` in the same AudioTensor class the caller handed us"""
    cls = type(ref) if (type(ref) is not torch.Tensor and isinstance(ref, torch.Tensor) and "hop_length" in getattr(ref, "__dict__", {})) else AudioTensor
    with _NoTF():
        return cls(data, hop_length=hop_length)
