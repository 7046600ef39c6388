"""GPU: edge shapes of every entry point against the oracle -- minimum and maximum orders, single samples,
inputs shorter than a hop / a kernel / the room response, odd batch sizes, hops that defeat every fast
path, dense (hop 1) coefficients with an initial state -- and the error behaviour on invalid input."""
import pytest
import torch

from conftest import REL_TOL, rel_rms, smooth, synthetic_controls

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def G():
    from golf_b200 import functional

    return functional


def cu(*ts):
    return [t.to(DEV) for t in ts]


def gen(seed):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize("B,Tn,H,M", [
    (1, 1, 240, 22),      # a single sample
    (1, 2, 1, 1),         # dense coefficients, order 1, two samples
    (5, 239, 240, 22),    # shorter than one hop
    (1, 241, 240, 1),     # order 1, one sample past a frame boundary
    (2, 3000, 240, 2),    # smallest padded-order bucket
    (2, 3000, 240, 40),   # largest supported order
    (3, 2500, 250, 39),   # hop divisible by no padded order -> generic solve
    (1, 4001, 7, 3),      # tiny prime hop
    (65, 960, 240, 22),   # odd batch larger than a warp's worth of chunks
    (2, 20000, 5000, 22), # very long frames (chunks subdivide a hop)
    (1, 720000, 240, 22), # 30 s utterance: 3000 chunks on the serial stitch
    (300, 2400, 240, 22), # many short utterances
])
def test_lpc_ss_edge_shapes(G, oracle, B, Tn, H, M):
    Fr = (Tn + H - 1) // H + 1
    gain, a = synthetic_controls(B, Fr, M, seed=Tn + M, scale=0.1)
    ex = torch.randn(B, Tn, generator=gen(Tn))
    ref = oracle.lpc_ss_fused(ex, gain, a, H)
    y = G.lpc_ss(*cu(ex, gain, a), H)
    assert y.shape == ref.shape
    assert torch.isfinite(y).all() and rel_rms(y, ref) < REL_TOL
    # and with control frames to spare (truncation to the shorter operand, audiotensor.py:134-152)
    gain2, a2 = synthetic_controls(B, Fr + 3, M, seed=Tn + M, scale=0.1)
    assert G.lpc_ss(*cu(ex, gain2, a2), H).shape == oracle.lpc_ss_fused(ex, gain2, a2, H).shape


@pytest.mark.parametrize("M", [1, 2, 23, 40])
def test_sample_wise_lpc_orders_and_initial_state(G, oracle, M):
    B, Tn = 3, 777
    g = gen(M)
    A = oracle.rc2lpc(torch.tanh((0.2 if M < 30 else 0.08) * smooth(torch.randn(B, Tn, M, generator=g), 32)))
    x, zi = torch.randn(B, Tn, generator=g), torch.randn(B, M, generator=g)
    assert rel_rms(G.sample_wise_lpc(*cu(x, A, zi)), oracle.sample_wise_lpc(x, A, zi)) < REL_TOL
    # gradient through the same call (autograd of the definition on the CPU is too slow for M = 40: check
    # a directional derivative against a central difference in float64 of the oracle instead)
    xg, Ag, zg = (t.to(DEV).requires_grad_() for t in (x, A, zi))
    up = torch.randn(B, Tn, generator=g)
    d_x, d_A, d_z = torch.autograd.grad(G.sample_wise_lpc(xg, Ag, zg), (xg, Ag, zg), up.to(DEV))
    vx, vA, vz = torch.randn(B, Tn, generator=g), 1e-2 * torch.randn(B, Tn, M, generator=g), torch.randn(B, M, generator=g)
    lhs = float((d_x.cpu() * vx).sum() + (d_A.cpu() * vA).sum() + (d_z.cpu() * vz).sum())
    eps = 1e-3
    f = lambda s: (oracle.sample_wise_lpc((x + s * vx).double(), (A + s * vA).double(), (zi + s * vz).double()).double() * up.double()).sum()
    rhs = float((f(eps) - f(-eps)) / (2 * eps))
    assert abs(lhs - rhs) <= 2e-3 * max(abs(rhs), 1.0), (lhs, rhs)


@pytest.mark.parametrize("B,Tn,H,M,W", [(1, 480, 240, 22, 480), (2, 700, 120, 1, 480), (1, 5000, 240, 40, 960), (3, 1203, 200, 22, 800)])
def test_lpc_ff_edge_shapes(G, oracle, B, Tn, H, M, W):
    Fr = Tn // H + 2
    gain, a = synthetic_controls(B, Fr, M, seed=Tn, scale=0.1)
    ex = torch.randn(B, Tn, generator=gen(Tn))
    ref = oracle.lpc_ff(ex, gain, a, H, W)
    y = G.lpc_ff(*cu(ex, gain, a, torch.hann_window(W)), H)
    assert y.shape == ref.shape and rel_rms(y, ref) < REL_TOL


@pytest.mark.parametrize("Tn,H,K,Fr", [
    (241, 240, 510, 1),   # exactly one block (the shortest input the reference accepts)
    (300, 240, 510, 5),   # more frames than blocks
    (64, 4, 2, 40),       # two taps, tiny hop
    (1000, 8, 6, 200),    # many tiny blocks
    (5000, 1024, 510, 6), # hop longer than 512: scalar kernel
    (3000, 250, 510, 14), # hop not a multiple of 4: scalar kernel
    (2400, 240, 1022, 10),# n_mag 512
])
def test_noise_fir_edge_shapes(G, oracle, Tn, H, K, Fr):
    ex, kern = torch.randn(2, Tn, generator=gen(K)), 0.05 * torch.randn(2, Fr, K, generator=gen(H))
    ref = oracle.ltv_fir_blocks(ex, kern, H)
    y = G.ltv_fir_blocks(*cu(ex, kern), H)
    assert y.shape == ref.shape and rel_rms(y, ref) < 1e-5


@pytest.mark.parametrize("Tn,n", [(1, 127), (50, 127), (127, 127), (128, 127), (1025, 127), (3000, 1), (3000, 255)])
def test_room_fir_edge_shapes(G, oracle, Tn, n):
    x, k = torch.randn(3, Tn, generator=gen(Tn)), 0.05 * torch.randn(n, generator=gen(n))
    ref = oracle.room_fir(x, k)
    y = G.room_fir(*cu(x, k))
    assert y.shape == ref.shape and rel_rms(y, ref) < 1e-5
    xg, kg = x.to(DEV).requires_grad_(), k.to(DEV).requires_grad_()
    xr, kr = x.clone().requires_grad_(), k.clone().requires_grad_()
    up = torch.randn(ref.shape, generator=gen(7))
    g_x, g_k = torch.autograd.grad(oracle.room_fir(xr, kr), (xr, kr), up)
    d_x, d_k = torch.autograd.grad(G.room_fir(xg, kg), (xg, kg), up.to(DEV))
    assert rel_rms(d_x, g_x) < 1e-5 and rel_rms(d_k[None], g_k[None]) < 1e-5


@pytest.mark.parametrize("B,Np,phase_hop,os_", [(1, 2, 120, 4), (2, 3, 1, 4), (1, 41, 120, 1), (3, 2401, 1, 2), (2, 9, 600, 4)])
def test_oscillator_edge_shapes(G, oracle, B, Np, phase_hop, os_):
    g = gen(Np)
    n_out = (Np - 1) * phase_hop + 1
    w_hop = 2400
    Fw = (n_out + w_hop - 1) // w_hop + 1
    ph = (80 + 300 * torch.rand(B, Np, generator=g)) / 24000
    w = torch.rand(B, Fw, generator=g)
    table, _ = oracle.glottal_table()
    dk = oracle.decimate_kernel(os_) if os_ > 1 else None
    y = G.glottal_osc(ph.to(DEV), phase_hop, w.to(DEV), w_hop, table.to(DEV), None if dk is None else dk.to(DEV), os_, True, "fp64")
    ref = oracle.glottal_osc(ph, phase_hop, w, w_hop, table, os_, True, "fp64")
    # exact-phase mode reads the table at exact fixed-point coordinates; the reference's grid_sample rounds the
    # column coordinate to 24 bits on its [-1, 1] detour (up to 3e-5 of a column, visible where the table is
    # steep).  Without oversampling no decimation filter averages that rounding noise out (nor the 2-ulp rsqrt and
    # the 1-ulp row coordinate of the faithful mode below): 5e-5 instead of 1e-5, still inside the 1e-4 bar
    tol = 1e-5 if os_ > 1 else 5e-5
    assert y.shape == ref.shape and rel_rms(y, ref) < tol, rel_rms(y, ref)
    y32 = G.glottal_osc(ph.to(DEV), phase_hop, w.to(DEV), w_hop, table.to(DEV), None if dk is None else dk.to(DEV), os_, True, "aten_cpu")
    e32 = rel_rms(y32, oracle.glottal_osc(ph, phase_hop, w, w_hop, table, os_, True, "fp32"))
    assert e32 < tol, e32


def test_invalid_inputs_raise(G):
    from golf_b200._lib import GolfError

    gain, a = synthetic_controls(2, 11, 22)
    ex = torch.randn(2, 2400)
    with pytest.raises((GolfError, AssertionError)):
        G.lpc_ss(ex, gain, a, 240)  # CPU tensors: no fallback
    with pytest.raises((GolfError, AssertionError, RuntimeError)):
        G.lpc_ss(*cu(ex, gain[:1], a), 240)  # batch mismatch
    with pytest.raises((GolfError, AssertionError)):
        gain41, a41 = synthetic_controls(2, 11, 41)
        G.lpc_ss(*cu(ex, gain41, a41), 240)  # order beyond the compiled buckets
    with pytest.raises((GolfError, AssertionError, RuntimeError)):
        G.lpc_ss(*cu(ex[:, :0], gain, a), 240)  # empty signal
    with pytest.raises((GolfError, AssertionError)):
        G.room_fir(torch.randn(2, 100).to(DEV).double(), torch.randn(127).to(DEV))  # not float32
