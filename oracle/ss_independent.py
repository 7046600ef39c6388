"""Two further restatements of torchlpc.sample_wise_lpc -- TEST INFRASTRUCTURE ONLY.

`oracle/golf_oracle.c` holds ONE restatement of the GOLF-ss recurrence (third-party
`torchlpc`, requirements.txt:19, unpinned and absent: PARITY UNPINNED at the source).  It is
used as oracle, as the golden-generation stub and as the CPU baseline, so a mistake in it
would go unnoticed.  This file adds two implementations that share no code and no structure
with it, both derived only from the call-site contract (models/filters.py:107-112,
models/lpc.py:11-16, models/lru/lru.py:9-15):

  y[b,t] = x[b,t] - sum_{i<M} A[b,t,i] * y[b,t-1-i],     y[b,-1-j] = zi[b,j]

* `sample_wise_lpc_banded`  -- no recurrence at all: the definition is the unit lower
  triangular banded linear system (I + L) y = x' with L[t, t-1-i] = A[t,i]; LAPACK's banded
  solver (scipy.linalg.solve_banded, float64) solves it.  The initial state moves to the
  right-hand side.
* `sample_wise_lpc_padded`  -- the *shape* of torchlpc's numba CPU kernel as published
  (recalled, not verifiable here): one padded buffer [zi reversed | x] per row, updated in
  place, `padded[t+M] -= A[t,i] * padded[t+M-i-1]` for i ascending; numba `prange` over the
  batch when numba is importable, plain NumPy loops otherwise.

tests/test_oracle.py checks the three against each other (float64: ~1e-12; float32: the
rounding floor) and the `zi` ordering through the split-continuation identity.
"""
from __future__ import annotations

from typing import Optional

import numpy as np


def sample_wise_lpc_banded(x: np.ndarray, A: np.ndarray, zi: Optional[np.ndarray] = None) -> np.ndarray:
    """float64 LAPACK solve of the defining linear system; x [B,T], A [B,T,M], zi [B,M]."""
    from scipy.linalg import solve_banded

    x = np.asarray(x, dtype=np.float64)
    A = np.asarray(A, dtype=np.float64)
    B, T = x.shape
    M = A.shape[2]
    y = np.empty_like(x)
    for b in range(B):
        rhs = x[b].copy()
        if zi is not None:
            z = np.asarray(zi[b], dtype=np.float64)
            for t in range(min(M, T)):
                # taps that reach before t = 0: i >= t  ->  y[t-1-i] = zi[i-t]
                rhs[t] -= np.dot(A[b, t, t:], z[: M - t])
        # banded storage: ab[l, j] = matrix[j + l, j]  (l = 0 diagonal .. M sub-diagonals)
        ab = np.zeros((M + 1, T))
        ab[0] = 1.0
        for i in range(M):
            # entry (row t, column t-1-i) = A[t, i]  ->  ab[i+1, t-1-i]
            n = T - 1 - i
            if n > 0:
                ab[i + 1, :n] = A[b, i + 1 :, i]
        y[b] = solve_banded((M, 0), ab, rhs, check_finite=False)
    return y


def _padded_rows(x, A, zi, out):
    B, T = x.shape
    M = A.shape[2]
    for b in range(B):
        buf = np.empty(T + M, dtype=x.dtype)
        for j in range(M):
            buf[j] = zi[b, M - 1 - j]
        buf[M:] = x[b]
        for t in range(T):
            for i in range(M):
                buf[t + M] -= A[b, t, i] * buf[t + M - i - 1]
        out[b] = buf[M:]


try:  # the reference's CPU path is numba; use it when present so the shape (and speed) match
    import numba as _nb

    @_nb.njit(parallel=True, fastmath=False, cache=False)
    def _padded_rows_nb(x, A, zi, out):  # pragma: no cover  (compiled)
        B, T = x.shape
        M = A.shape[2]
        for b in _nb.prange(B):
            buf = np.empty(T + M, dtype=x.dtype)
            for j in range(M):
                buf[j] = zi[b, M - 1 - j]
            for t in range(T):
                buf[M + t] = x[b, t]
            for t in range(T):
                for i in range(M):
                    buf[t + M] -= A[b, t, i] * buf[t + M - i - 1]
            for t in range(T):
                out[b, t] = buf[M + t]
except Exception:  # noqa: BLE001
    _padded_rows_nb = None


def sample_wise_lpc_padded(x: np.ndarray, A: np.ndarray, zi: Optional[np.ndarray] = None, use_numba: bool = True) -> np.ndarray:
    """same dtype in, same dtype out (float32 or float64)"""
    x = np.ascontiguousarray(x)
    A = np.ascontiguousarray(A, dtype=x.dtype)
    B, T = x.shape
    M = A.shape[2]
    z = np.zeros((B, M), dtype=x.dtype) if zi is None else np.ascontiguousarray(zi, dtype=x.dtype)
    out = np.empty_like(x)
    if use_numba and _padded_rows_nb is not None:
        _padded_rows_nb(x, A, z, out)
    else:
        _padded_rows(x, A, z, out)
    return out
