"""A minimal stand-in for the `lightning` package -- TEST INFRASTRUCTURE (see oracle/__init__.py), build container only.

The reference's harness (`ltng/ae.py`, `ltng/vocoder.py`, `ltng/cli.py`, `test_rtf.py`) imports `lightning`, which is
not installed in this image and cannot be (no network).  This module registers just enough of its surface in
``sys.modules`` for those files to be imported UNMODIFIED from /root/reference and for the inference-side calls
`test_rtf.py` makes to work:

  lightning.pytorch.LightningModule      nn.Module + log/log_dict (recorded, not reduced), `logger`, `device`,
                                         `load_from_checkpoint(path, map_location, **init_kwargs)` (strict state-dict load
                                         of ckpt["state_dict"], what Lightning does when init args are passed explicitly)
  lightning.pytorch.cli.LightningCLI / LightningArgumentParser, lightning.pytorch.callbacks.{Callback,BasePredictionWriter},
  lightning.{LightningModule,Trainer}, lightning.fabric.utilities.cloud_io.get_filesystem      import-only placeholders

No training loop: `autoencode.py fit` needs the real package.  tools/fit_step.py is this repo's stand-in for that step.
"""
from __future__ import annotations

import sys
import types

import torch


class LightningModule(torch.nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        self.logged = {}
        self.logger = None
        self.trainer = None

    def log(self, name, value, *args, **kwargs):
        self.logged[name] = value

    def log_dict(self, values, *args, **kwargs):
        self.logged.update(values)

    def save_hyperparameters(self, *args, **kwargs):
        pass

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, strict=True, **kwargs):
        ckpt = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=True)
        model = cls(**kwargs)
        model.load_result = model.load_state_dict(ckpt["state_dict"], strict=strict)
        model.on_load_checkpoint(ckpt) if hasattr(model, "on_load_checkpoint") else None
        return model.to(map_location) if map_location is not None else model


class _Placeholder:
    def __init__(self, *args, **kwargs):
        raise RuntimeError("lightning stand-in: only LightningModule is functional (oracle/lightning_standin.py)")


def install() -> None:
    if "lightning" in sys.modules and not getattr(sys.modules["lightning"], "__golf_standin__", False):
        return  # the real package is present: use it

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__golf_standin__ = True
        sys.modules[name] = m
        return m

    ph = lambda n: type(n, (_Placeholder,), {})  # noqa: E731
    Trainer, Callback, Writer = ph("Trainer"), type("Callback", (), {}), type("BasePredictionWriter", (), {"__init__": lambda self, *a, **k: None})
    root = mod("lightning", LightningModule=LightningModule, Trainer=Trainer)
    root.pytorch = mod("lightning.pytorch", LightningModule=LightningModule, Trainer=Trainer)
    root.pytorch.cli = mod("lightning.pytorch.cli", LightningCLI=ph("LightningCLI"), LightningArgumentParser=ph("LightningArgumentParser"))
    root.pytorch.callbacks = mod("lightning.pytorch.callbacks", Callback=Callback, BasePredictionWriter=Writer)
    root.fabric = mod("lightning.fabric")
    root.fabric.utilities = mod("lightning.fabric.utilities")
    root.fabric.utilities.cloud_io = mod("lightning.fabric.utilities.cloud_io", get_filesystem=lambda *a, **k: None)
