"""Reference interop: when golf_b200 modules are selected by class_path inside the
reference's own process (ltng/ae.py, test_rtf.py, autoencode.py), they must BE the
reference's boundary types -- `models.sf.Synth.split_sizes_and_trsfms` counts children
with `isinstance(m, models.ctrl.Controllable)` (models/ctrl.py:59-69) and the oscillator
hooks assert `isinstance(out, models.audiotensor.AudioTensor)` (models/synth.py:32-35).

So: if the reference's `models.ctrl` / `models.audiotensor` are importable (and look like
GOLF's), golf_b200 re-uses those classes as its bases; otherwise it uses its own
implementations (golf_b200/audiotensor.py, golf_b200/ctrl.py), which follow the same
protocol.  GOLF_B200_INTEROP=0 forces the standalone types, =1 requires the reference.
"""
from __future__ import annotations

import importlib
import os

_mode = os.environ.get("GOLF_B200_INTEROP", "auto")

ref_ctrl = None
ref_audiotensor = None

if _mode != "0":
    try:
        _c = importlib.import_module("models.ctrl")
        _a = importlib.import_module("models.audiotensor")
        if all(hasattr(_c, n) for n in ("Controllable", "wrap_ctrl_fn", "PassThrough", "Synth")) and hasattr(_a, "AudioTensor"):
            ref_ctrl, ref_audiotensor = _c, _a
    except Exception:  # not inside the reference tree
        if _mode == "1":
            raise

INTEROP = ref_ctrl is not None
