"""ctypes binding of libgolf_b200.so (include/golf_b200.h).  No fallback: if the
library is missing or a call fails, this raises -- the product path never routes
around the CUDA kernels."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_float, c_int, c_int64, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("GOLF_B200_SO") or os.path.join(_HERE, "_lib", "libgolf_b200.so")  # env: A/B builds (tools/)

ABI_VERSION = 8
_lib = None

P = c_void_p
_SIGS = {
    "golf_abi_version": (c_int, []),
    "golf_strerror": (ctypes.c_char_p, [c_int]),
    "golf_last_cuda_error": (c_int, []),
    "golf_launch_count": (c_uint64, []),
    "golf_set_pdl": (None, [c_int]),
    "golf_lpc_ss_workspace_bytes": (c_size_t, [c_int] * 5),
    "golf_lpc_ss_set_refine_tolerance": (None, [c_float]),
    "golf_lpc_ss_get_refine_tolerance": (c_float, []),
    "golf_lpc_ss_set_solver": (None, [c_int]),
    "golf_lpc_ss_fwd": (c_int, [P, c_int64, P, P, P, P] + [c_int] * 6 + [P, c_size_t, P]),
    "golf_lpc_ss_fwd_passes": (c_int, [P, c_int64, P, P, P, P] + [c_int] * 6 + [P, c_size_t, c_int, P]),
    "golf_lpc_ss_set_tail": (None, [c_int]),
    "golf_lpc_ss_get_tail": (c_int, []),
    "golf_lpc_ss_set_response": (None, [c_int]),
    "golf_lpc_ss_get_response": (c_int, []),
    "golf_lpc_ss_room_workspace_bytes": (c_size_t, [c_int] * 5),
    "golf_lpc_ss_room_fwd": (c_int, [P, c_int64, P, P, P, P, c_int, P, P] + [c_int] * 7 + [P, c_size_t, P]),
    "golf_lpc_ss_bwd_workspace_bytes": (c_size_t, [c_int] * 5),
    "golf_lpc_ss_bwd": (c_int, [P, P, P, c_int64, P, P, P, P, P, P, P] + [c_int] * 7 + [P, c_size_t, P]),
    "golf_lpc_ff_set_exact_order": (None, [c_int]),
    "golf_lpc_ff_fwd": (c_int, [P, c_int64, P, P, P, P] + [c_int] * 6 + [P]),
    "golf_lpc_ff_bwd_workspace_bytes": (c_size_t, [c_int] * 5),
    "golf_lpc_ff_bwd": (c_int, [P, P, c_int64, P, P, P, P, c_int64, P, P] + [c_int] * 6 + [P, c_size_t, P]),
    "golf_biquad_ff_fwd": (c_int, [P, c_int64, P, P, P, P] + [c_int] * 6 + [P]),
    "golf_biquad_cascade_fwd": (c_int, [P, c_int64, P, P, P, P] + [c_int] * 6 + [P]),
    "golf_biquad_cascade_bwd_workspace_bytes": (c_size_t, [c_int] * 6),
    "golf_biquad_cascade_bwd": (c_int, [P, P, c_int64, P, P, P, P, P, P] + [c_int] * 6 + [P, c_size_t, P]),
    "golf_lpc_frames_out_length": (c_int, [c_int] * 4),
    "golf_lpc_frames_fwd": (c_int, [P, c_int64, P, P, P, P] + [c_int] * 6 + [P]),
    "golf_lpc_frames_bwd_workspace_bytes": (c_size_t, [c_int] * 5),
    "golf_lpc_frames_bwd": (c_int, [P, P, c_int64, P, P, P, P, P, P] + [c_int] * 6 + [P, c_size_t, P]),
    "golf_lfilter_allpole_fwd": (c_int, [P, c_int64, P, P, P, c_int, c_int, c_int, P]),
    "golf_lfilter_allpole_bwd_workspace_bytes": (c_size_t, [c_int, c_int]),
    "golf_lfilter_allpole_bwd": (c_int, [P, P, c_int64, P, P, P, P, P, P, c_int, c_int, c_int, P, c_size_t, P]),
    "golf_biquad_params_fwd": (c_int, [P, P, P, c_int, c_int, c_int, c_float, P]),
    "golf_biquad_params_bwd": (c_int, [P, P, P, P, c_int, c_int, c_int, c_float, P]),
    "golf_mss_tables_bytes": (c_size_t, [c_int]),
    "golf_mss_build_tables": (c_int, [c_int, P, P]),
    "golf_mss_workspace_bytes": (c_size_t, [c_int, c_int, P, P, c_int]),
    "golf_mss_loss": (c_int, [P, c_int64, P, c_int64, c_int, c_int, P, P, c_int, P, c_float, c_float, c_float, P, P, c_int64, c_int, P, c_size_t, P]),
    "golf_mss_gemm": (c_int, [P, c_int64, P, c_int64, P, c_int64, c_int, c_int, c_int, c_int, c_int, P]),
    "golf_lpc_inverse_fwd": (c_int, [P, c_int64, P, P] + [c_int] * 5 + [P]),
    "golf_lpc_inverse_bwd": (c_int, [P, P, c_int64, P, P, P] + [c_int] * 5 + [P]),
    "golf_noise_fir_fwd": (c_int, [P, c_int64, P, P, P, c_int64, P] + [c_int] * 5 + [P]),
    "golf_fir_set_variant": (None, [c_int]),
    "golf_noise_fir_bwd": (c_int, [P, P, c_int64, P, P, P] + [c_int] * 5 + [P]),
    "golf_noise_fir_design_supported": (c_int, [c_int, c_int]),
    "golf_noise_fir_design_fwd": (c_int, [P, c_int64, P, P, P, P, c_int64, P] + [c_int] * 5 + [P]),
    "golf_philox_normal": (c_int, [P, c_int, c_int, P, P]),
    "golf_rng_advance": (c_int, [P, P]),
    "golf_synth_fused_workspace_bytes": (c_size_t, [c_int] * 10),
    "golf_synth_fused_out_length": (c_int, [c_int] * 6),
    "golf_synth_fused_fwd": (c_int, [P, P, P, P, P, c_int64, P, P, P, P, P, P, c_int, P] + [c_int] * 16 + [P, c_size_t, P]),
    "golf_room_fir_fwd": (c_int, [P, P, P, c_int, c_int, c_int, P]),
    "golf_room_fir_bwd": (c_int, [P, P, P, P, P, c_int, c_int, c_int, P]),
    "golf_glottal_osc_workspace_bytes": (c_size_t, [c_int] * 6),
    "golf_glottal_osc_set_variant": (None, [c_int]),
    "golf_glottal_osc_fwd": (c_int, [P, P, P, P, P] + [c_int] * 11 + [P, c_size_t, P]),
    "golf_glottal_osc_fwd_from": (c_int, [P, P, P, P, P, P] + [c_int] * 11 + [P, c_size_t, P]),
    "golf_glottal_osc_bwd_w": (c_int, [P, P, P, P, P, P] + [c_int] * 11 + [P, c_size_t, P]),
    "golf_glottal_osc_bwd": (c_int, [P, P, P, P, P, P, P] + [c_int] * 11 + [P, c_size_t, P]),
    "golf_wavetable_read_fwd": (c_int, [P, P, P] + [c_int] * 5 + [P]),
    "golf_linear_upsample": (c_int, [P, P, c_int, c_int, c_int, P]),
    "golf_rc2lpc_fwd": (c_int, [P, P, c_int, c_int, c_float, P]),
    "golf_rc2lpc_bwd": (c_int, [P, P, P, c_int, c_int, c_float, P]),
    "golf_exp_to_complex": (c_int, [P, P, c_int64, P]),
}
EXPORTS = tuple(_SIGS)


class GolfError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise GolfError(
                f"{SO_PATH} is missing: build it with `python -m golf_b200.build` "
                "(or __graft_entry__.build()).  golf_b200 has no CPU/PyTorch fallback."
            )
        L = ctypes.CDLL(SO_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        if L.golf_abi_version() != ABI_VERSION:
            raise GolfError(f"ABI mismatch: library {L.golf_abi_version()} vs binding {ABI_VERSION}; rebuild")
        _lib = L
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        L = lib()
        msg = L.golf_strerror(rc).decode()
        extra = f" (cudaError {L.golf_last_cuda_error()})" if rc == -4 else ""
        raise GolfError(f"{what}: {msg}{extra}")


def launch_count() -> int:
    return int(lib().golf_launch_count())
