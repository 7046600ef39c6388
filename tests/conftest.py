import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REL_TOL = 1e-4  # north_star: float32 match within 1e-4 relative RMS on identical inputs


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the reference tree at /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    has_gpu = torch.cuda.is_available()
    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))


def rel_rms(x, y):
    """max over rows of rms(x - y) / rms(y), in float64"""
    x = torch.as_tensor(x).detach().double().cpu()
    y = torch.as_tensor(y).detach().double().cpu()
    assert x.shape == y.shape, (tuple(x.shape), tuple(y.shape))
    x, y = x.reshape(x.shape[0], -1), y.reshape(y.shape[0], -1)
    return float((((x - y) ** 2).mean(-1) / (y**2).mean(-1).clamp_min(1e-300)).sqrt().max())


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def T(arr):
    return torch.tensor(np.asarray(arr))


def smooth(x, n=16):
    """variance-preserving moving average over n frames along dim 1 (SURVEY 8d recipe)"""
    xt = x.transpose(1, -1) if x.ndim > 2 else x
    flat = xt.reshape(-1, 1, xt.shape[-1])
    y = torch.nn.functional.conv1d(torch.nn.functional.pad(flat, (n - 1, 0), mode="replicate"), torch.ones(1, 1, n) / n)
    y = (y * n**0.5).view(xt.shape)
    return y.transpose(1, -1) if x.ndim > 2 else y


def synthetic_controls(B, frames, M, seed=0, scale=0.15):
    """stable smooth reflection-coefficient trajectories -> (gain [B,F], a [B,F,M])"""
    from oracle import golf_oracle as O

    g = torch.Generator().manual_seed(seed)
    a = O.rc2lpc(torch.tanh(scale * smooth(torch.randn(B, frames, M, generator=g))))
    gain = torch.exp(smooth(torch.randn(B, frames, generator=g)) - 6)
    return gain, a


@pytest.fixture(scope="session")
def oracle():
    from oracle import golf_oracle as O

    O.build()
    return O


@pytest.fixture(scope="session")
def reference():
    from oracle import refimport

    if not refimport.available():
        pytest.skip("reference tree not present (GPU box)")
    return refimport.import_reference()
