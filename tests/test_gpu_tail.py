"""GPU: the one-launch tail of GOLF-ss (csrc/lpc_ss_tail.cuh: two-level stitch + solve + refinement + room FIR by a
thread-block cluster per sequence) against the oracle, against the separate stitch / solve launches it replaces, and the
fused room FIR against the stand-alone one."""
import pytest
import torch

from conftest import REL_TOL, T, golden, rel_rms, synthetic_controls

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def G():
    from golf_b200 import functional

    return functional


@pytest.fixture()
def tail_switch():
    from golf_b200 import _lib

    L = _lib.lib()
    yield L.golf_lpc_ss_set_tail
    L.golf_lpc_ss_set_tail(0)  # library default: the light stitch / solve launches
    L.golf_lpc_ss_set_refine_tolerance(1e-4)


def cu(*ts):
    return [t.to(DEV) for t in ts]


@pytest.mark.parametrize("B,Tn,H,M", [
    (2, 4800, 240, 22), (3, 12000, 240, 22), (2, 12000, 120, 22), (2, 9600, 240, 16), (2, 9600, 240, 20), (2, 9600, 240, 32),
    (1, 500, 240, 22), (1, 23, 240, 22), (2, 481, 240, 22), (1, 70001, 240, 22), (2, 62000, 120, 30), (5, 48000, 240, 22), (2, 9984, 256, 30), (2, 9984, 128, 30)])
def test_tail_matches_oracle_and_separate_launches(G, oracle, tail_switch, B, Tn, H, M):
    Fr = (Tn + H - 1) // H + 1
    gain, a = synthetic_controls(B, Fr, M, seed=M + H + B)
    ex = torch.randn(B, Tn, generator=torch.Generator().manual_seed(1))
    ref = oracle.lpc_ss_fused(ex, gain, a, H)
    ref64 = oracle.lpc_ss_fused(ex, gain, a, H, double=True)
    exd, gd, ad = cu(ex, gain, a)
    tail_switch(1)
    y1 = G.lpc_ss(exd, gd, ad, H)
    tail_switch(0)
    y0 = G.lpc_ss(exd, gd, ad, H)
    assert y1.shape == ref.shape
    assert rel_rms(y1, ref) < REL_TOL and rel_rms(y1, ref64) < REL_TOL
    floor = rel_rms(ref, ref64)  # what float32 itself costs on this input
    assert rel_rms(y1, y0) < max(2e-5, 4 * floor)
    assert rel_rms(y1, ref64) < 2 * rel_rms(y0, ref64) + 1e-6  # no less accurate than the path it replaces


def test_tail_initial_state_and_dense_coefficients(G, oracle, tail_switch):
    """hop == 1 (torchlpc.sample_wise_lpc surface) with zi goes through the same planner"""
    from conftest import smooth

    B, Tn, M = 2, 6000, 22
    g = torch.Generator().manual_seed(3)
    A = oracle.rc2lpc(torch.tanh(0.15 * smooth(torch.randn(B, Tn, M, generator=g), 64)))
    x, zi = torch.randn(B, Tn, generator=g), torch.randn(B, M, generator=g)
    ref = oracle.sample_wise_lpc(x, A, zi)
    for mode in (1, 0):
        tail_switch(mode)
        assert rel_rms(G.sample_wise_lpc(*cu(x, A, zi)), ref) < 2e-5


def test_tail_refinement_round(G, oracle, tail_switch):
    """forced refinement (tolerance 0) on the encoder-derived controls and on the ill-conditioned 4-coincident-pole
    case: the two-level mismatch propagation must reach the same accuracy as the sequential one"""
    from golf_b200 import _lib

    L = _lib.lib()
    g = golden("controls_gt")
    gain, a, H = T(g["gain"]), T(g["a"]), int(g["hop"])
    ex = torch.randn(gain.shape[0], (gain.shape[1] - 1) * H, generator=torch.Generator().manual_seed(0))
    ref32 = oracle.lpc_ss_fused(ex, gain, a, H)
    ref64 = oracle.lpc_ss_fused(ex, gain, a, H, double=True)
    floor = rel_rms(ref32, ref64)
    for tol in (0.0, 1e-4):
        L.golf_lpc_ss_set_refine_tolerance(tol)
        tail_switch(1)
        y = G.lpc_ss(*cu(ex, gain, a), H)
        assert rel_rms(y, ref64) < 3 * floor, tol
    f = golden("filters_rand")
    ex, gain, a, H = T(f["ex_8"]), T(f["gain_8"]), T(f["a_8"]), int(f["hop"])
    ref64 = oracle.lpc_ss_fused(ex, gain, a, H, double=True)
    floor = rel_rms(oracle.lpc_ss_fused(ex, gain, a, H), ref64)
    L.golf_lpc_ss_set_refine_tolerance(1e-4)
    errs = {}
    for mode in (1, 0):
        tail_switch(mode)
        y = G.lpc_ss(*cu(ex, gain, a), H)
        assert torch.isfinite(y).all()
        errs[mode] = rel_rms(y, ref64)
    assert errs[1] < max(REL_TOL, 100 * floor), errs
    assert errs[1] < 3 * errs[0] + 1e-6, errs


def test_tail_propagates_nonfinite(G, tail_switch):
    tail_switch(1)
    gain, a = synthetic_controls(1, 41, 22)
    ex = torch.randn(1, 9600)
    ex[0, 5000] = float("inf")
    y = G.lpc_ss(*cu(ex, gain, a), 240)
    assert torch.isfinite(y[0, :5000]).all() and not torch.isfinite(y[0, 5000:]).any()


@pytest.mark.parametrize("B,Tn,H,M,n", [(2, 9600, 240, 22, 127), (3, 12001, 120, 22, 127), (1, 300, 240, 22, 127), (2, 9600, 240, 12, 127),
                                        (2, 9600, 240, 32, 63), (32, 47760, 240, 22, 127)])
def test_fused_room_equals_separate_room(G, oracle, tail_switch, B, Tn, H, M, n):
    Fr = (Tn + H - 1) // H + 1
    gain, a = synthetic_controls(B, Fr, M, seed=7 + M)
    g = torch.Generator().manual_seed(4)
    ex = torch.randn(B, Tn, generator=g)
    k = 0.05 * torch.randn(n, generator=g)
    exd, gd, ad, kd = cu(ex, gain, a, k)
    tail_switch(1)
    fused = G.lpc_ss_room(exd, gd, ad, kd, H)
    y = G.lpc_ss(exd, gd, ad, H)
    sep = G.room_fir(y, kd)
    assert fused.shape == sep.shape
    assert torch.equal(fused, sep)  # same y, same tap order
    ref = oracle.room_fir(oracle.lpc_ss_fused(ex, gain, a, H), k)
    assert rel_rms(fused, ref) < REL_TOL
    tail_switch(0)  # fallback inside the library: filter launches + stand-alone FIR
    assert rel_rms(G.lpc_ss_room(exd, gd, ad, kd, H), ref) < REL_TOL


def test_fused_room_gradients(G, tail_switch):
    tail_switch(1)
    B, Tn, H, M = 2, 4800, 240, 22
    gain, a = synthetic_controls(B, Tn // H + 1, M, seed=11)
    g = torch.Generator().manual_seed(5)
    ex, k, up = torch.randn(B, Tn, generator=g), 0.05 * torch.randn(127, generator=g), torch.randn(B, Tn, generator=g)
    leaves = [t.to(DEV).requires_grad_() for t in (ex, gain, a, k)]
    out = G.lpc_ss_room(*leaves, H)
    g1 = torch.autograd.grad(out, leaves, up.to(DEV))
    leaves2 = [t.to(DEV).requires_grad_() for t in (ex, gain, a, k)]
    out2 = G.room_fir(G.lpc_ss(leaves2[0], leaves2[1], leaves2[2], H), leaves2[3])
    g2 = torch.autograd.grad(out2, leaves2, up.to(DEV))
    assert torch.equal(out, out2)
    for x, y in zip(g1, g2):
        assert rel_rms(x.reshape(B if x.ndim > 1 else 1, -1), y.reshape(B if y.ndim > 1 else 1, -1)) < 1e-5


def test_decoder_uses_fused_room(G, tail_switch):
    """SourceFilterSynth routes end filter + room filter through the fused entry point and gets the same samples"""
    import bench
    from golf_b200 import noise as gnoise, sf
    from golf_b200.audiotensor import AudioTensor

    tail_switch(1)
    s = {k: v[:3] for k, v in bench.make_inputs(1, 4)[0].items()}
    noise = torch.randn(3, bench.T, generator=torch.Generator().manual_seed(2)).to(DEV)

    class Fixed(gnoise.NoiseInterface):
        def forward(self, ref_, *args):
            return AudioTensor(noise[:, : ref_.shape[1]])

    dec = bench.build_decoder(torch.device(DEV), "ss")
    dec.noise_generator = Fixed()
    A = lambda t, hop: AudioTensor(t.to(DEV), hop_length=hop)
    P = dict(phase=A(s["phase"], 1), harm_oscillator_params=(A(s["w"], 2400),), noise_generator_params=(),
             noise_filter_params=(A(s["log_mag"], bench.HOP),), end_filter_params=(A(s["gain"], bench.HOP), A(s["a"], bench.HOP)))
    outs = {}
    try:
        for fuse in (True, False):
            sf.FUSE_ROOM = fuse
            with torch.no_grad():
                outs[fuse] = dec(**P).as_tensor()
    finally:
        sf.FUSE_ROOM = True
    assert torch.equal(outs[True], outs[False])
