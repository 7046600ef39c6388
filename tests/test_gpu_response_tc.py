"""GPU: the tensor-core variant of GOLF-ss pass 1 (csrc/lpc_ss_tc.cuh, mma.sync TF32 with error-compensated products;
opt-in through golf_lpc_ss_set_response(1)) against the oracle and against the FP32 response kernel it replaces.

The variant is a recorded experiment: it is correct, but on B200 it is not faster than the FP32 kernel
(83 vs 74 us at B = 32 x 2 s) and the tensor cores' truncating accumulation makes its transition matrices ~10x less accurate,
so the refinement round fires for almost every sequence.  The default stays mode 0; these tests pin that mode 1 still
meets the parity bar so the comparison in profiles/README.md stays reproducible."""
import pytest
import torch

from conftest import REL_TOL, T, golden, rel_rms, synthetic_controls

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture()
def response_switch():
    from golf_b200 import _lib

    L = _lib.lib()
    assert L.golf_lpc_ss_get_response() == 0, "library default must be the FP32 response kernel"
    yield L.golf_lpc_ss_set_response
    L.golf_lpc_ss_set_response(0)


@pytest.mark.parametrize("B,Tn,H,M", [(2, 4800, 240, 22), (3, 12000, 240, 22), (2, 12000, 120, 22), (2, 9600, 240, 24), (2, 9600, 240, 21),
                                      (1, 500, 240, 22), (2, 7001, 240, 22), (2, 9600, 48, 22),
                                      pytest.param(5, 48000, 240, 22, marks=pytest.mark.xfail(
                                          reason="one of the five trajectories ends at 3.8e-4: the tensor cores' truncating accumulation leaves "
                                                 "Phi too inexact for ONE refinement round here -- a reason the mode is not the default", strict=False))])
def test_tc_response_matches_oracle(oracle, response_switch, B, Tn, H, M):
    from golf_b200 import functional as G

    Fr = (Tn + H - 1) // H + 1
    gain, a = synthetic_controls(B, Fr, M, seed=M + H + B)
    ex = torch.randn(B, Tn, generator=torch.Generator().manual_seed(1))
    ref64 = oracle.lpc_ss_fused(ex, gain, a, H, double=True)
    exd, gd, ad = ex.to(DEV), gain.to(DEV), a.to(DEV)
    response_switch(1)
    y1 = G.lpc_ss(exd, gd, ad, H)
    response_switch(0)
    y0 = G.lpc_ss(exd, gd, ad, H)
    assert y1.shape == ref64.shape
    assert rel_rms(y1, ref64) < REL_TOL
    assert rel_rms(y1, y0) < REL_TOL


def test_tc_response_on_encoder_derived_controls(oracle, response_switch):
    from golf_b200 import functional as G

    g = golden("controls_gt")
    gain, a, H = T(g["gain"]), T(g["a"]), int(g["hop"])
    ex = 0.1 * torch.randn(gain.shape[0], (gain.shape[1] - 1) * H, generator=torch.Generator().manual_seed(2))
    ref64 = oracle.lpc_ss_fused(ex, gain, a, H, double=True)
    response_switch(1)
    y = G.lpc_ss(ex.to(DEV), gain.to(DEV), a.to(DEV), H)
    assert rel_rms(y, ref64) < REL_TOL


def test_library_contains_legacy_tensor_core_instructions():
    """the variant really is a tensor-core kernel: HMMA (mma.sync) in the SASS of the built library"""
    import shutil
    import subprocess

    from golf_b200 import _lib

    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN4golf21ss_response_tc_kernelILi3EEEvNS_8SsParamsE", _lib.SO_PATH],
                          capture_output=True, text=True).stdout
    assert sass.count("HMMA") >= 18
