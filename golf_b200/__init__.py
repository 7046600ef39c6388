"""golf_b200 -- B200 (sm_100a) implementation of GOLF's sample-recurrent synthesis hot path.

Drop-in modules (same class names / init args / `.ctrl` / state-dict keys as the reference,
iamycy/golf `models/`): select them with `class_path: golf_b200.filters.LTVMinimumPhaseFilter`
etc., or import `golf_b200.functional` for the tensor-level ops.  The compute lives in
`_lib/libgolf_b200.so` (C ABI, include/golf_b200.h); there is no CPU fallback.
"""
from ._lib import ABI_VERSION, GolfError, SO_PATH, launch_count  # noqa: F401
from .audiotensor import AudioTensor  # noqa: F401

__version__ = "0.1.0"
