"""GPU: the CUDA path (through the C ABI) against the CPU oracle and the reference's golden
vectors.  Bar (north_star): relative RMS <= 1e-4 per utterance in float32 on identical inputs;
most kernels sit two orders of magnitude inside it, and the bounds below say so."""
import numpy as np
import pytest
import torch

from conftest import REL_TOL, T, golden, rel_rms, smooth, synthetic_controls

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def G():
    from golf_b200 import functional

    return functional


def cu(*ts):
    return [t.to(DEV) for t in ts]


# ------------------------------------------------------------------ GOLF-ss forward
@pytest.mark.parametrize("B,Tn,H,M", [
    (2, 4800, 240, 22), (3, 12000, 240, 22), (2, 12000, 120, 22), (2, 9600, 240, 12), (2, 9600, 240, 20),
    (2, 9600, 240, 32), (2, 9600, 240, 8), (2, 7001, 240, 22), (1, 500, 240, 22), (1, 23, 240, 22),
    (2, 9600, 256, 22), (2, 9600, 100, 22), (2, 5000, 2400, 22)])
def test_lpc_ss_matches_oracle(G, oracle, B, Tn, H, M):
    Fr = (Tn + H - 1) // H + 1
    gain, a = synthetic_controls(B, Fr, M, seed=M + H)
    ex = torch.randn(B, Tn, generator=torch.Generator().manual_seed(1))
    ref = oracle.lpc_ss_fused(ex, gain, a, H)
    y = G.lpc_ss(*cu(ex, gain, a), H)
    assert y.shape == ref.shape
    assert rel_rms(y, ref) < REL_TOL
    assert rel_rms(y, oracle.lpc_ss_fused(ex, gain, a, H, double=True)) < REL_TOL


def test_lpc_ss_encoder_derived_controls(G, oracle):
    """the demanding set: controls from the real encoder on gt_*.wav (pole radius up to 0.996)"""
    g = golden("controls_gt")
    gain, a, H = T(g["gain"]), T(g["a"]), int(g["hop"])
    ex = torch.randn(gain.shape[0], (gain.shape[1] - 1) * H, generator=torch.Generator().manual_seed(0))
    ref32 = oracle.lpc_ss_fused(ex, gain, a, H)
    ref64 = oracle.lpc_ss_fused(ex, gain, a, H, double=True)
    y = G.lpc_ss(*cu(ex, gain, a), H)
    floor = rel_rms(ref32, ref64)  # what float32 itself costs on this input (~1.3e-5)
    assert rel_rms(y, ref32) < REL_TOL and rel_rms(y, ref64) < REL_TOL
    assert rel_rms(y, ref64) < 3 * floor
    for chunk in (480, 960):
        assert rel_rms(G.lpc_ss(*cu(ex, gain, a), H, chunk=chunk), ref32) < REL_TOL


@pytest.mark.parametrize("M", [20, 22])
def test_lpc_ss_reference_golden(G, M):
    g = golden("filters_rand")
    y = G.lpc_ss(*cu(T(g[f"ex_{M}"]), T(g[f"gain_{M}"]), T(g[f"a_{M}"])), int(g["hop"]))
    assert rel_rms(y, T(g[f"ss_{M}"])) < 1e-5


def test_lpc_ss_ill_conditioned_biquad_poles(G, oracle):
    """4 coincident pole pairs at radius ~0.998 (gain x700): float32 itself is only good to
    ~8e-4 here (oracle32 vs float64), so the bar is float64 truth within a small multiple of
    that floor rather than 1e-4."""
    g = golden("filters_rand")
    ex, gain, a, H = T(g["ex_8"]), T(g["gain_8"]), T(g["a_8"]), int(g["hop"])
    ref64 = oracle.lpc_ss_fused(ex, gain, a, H, double=True)
    floor = rel_rms(oracle.lpc_ss_fused(ex, gain, a, H), ref64)
    y = G.lpc_ss(*cu(ex, gain, a), H)
    assert torch.isfinite(y).all()
    assert rel_rms(y, ref64) < max(REL_TOL, 100 * floor)


def test_sample_wise_lpc_dense_coefficients(G, oracle):
    """torchlpc.sample_wise_lpc surface: sample-rate A, initial state zi, order 1 (lru.py:9-15)"""
    B, Tn, M = 2, 3000, 6
    g = torch.Generator().manual_seed(3)
    A = oracle.rc2lpc(torch.tanh(0.3 * smooth(torch.randn(B, Tn, M, generator=g), 64)))
    x, zi = torch.randn(B, Tn, generator=g), torch.randn(B, M, generator=g)
    assert rel_rms(G.sample_wise_lpc(*cu(x, A, zi)), oracle.sample_wise_lpc(x, A, zi)) < 1e-5
    assert rel_rms(G.sample_wise_lpc(*cu(x, A)), oracle.sample_wise_lpc(x, A)) < 1e-5
    lam = torch.rand(B, Tn, 1, generator=g) * 0.9
    assert rel_rms(G.sample_wise_lpc(*cu(x, -lam, zi[:, :1])), oracle.sample_wise_lpc(x, -lam, zi[:, :1])) < 1e-5


def test_lpc_ss_propagates_nonfinite(G):
    gain, a = synthetic_controls(1, 21, 22)
    ex = torch.randn(1, 4800)
    ex[0, 1000] = float("inf")
    y = G.lpc_ss(*cu(ex, gain, a), 240)
    assert torch.isfinite(y[0, :1000]).all() and not torch.isfinite(y[0, 1000:]).any()


def test_lpc_ss_full_size_linearity(G):
    """BASELINE size (B=32 x 2 s, M=22): size-independent properties -- superposition and
    scaling of the excitation -- since the oracle is too slow to be the per-element check"""
    B, Tn, H, M = 32, 48000, 240, 22
    gain, a = synthetic_controls(B, Tn // H, M, seed=9)
    g = torch.Generator().manual_seed(5)
    x1, x2 = torch.randn(B, Tn, generator=g), torch.randn(B, Tn, generator=g)
    gain, a, x1, x2 = cu(gain, a, x1, x2)
    y1, y2, y12 = G.lpc_ss(x1, gain, a, H), G.lpc_ss(x2, gain, a, H), G.lpc_ss(x1 + 0.5 * x2, gain, a, H)
    assert y1.shape == (B, (Tn // H - 1) * H + 1)
    assert rel_rms(y12, y1 + 0.5 * y2) < 2e-5
    # inverse filter undoes the synthesis filter (encode -> decode round trip)
    e = G.lpc_inverse(y1, a, H)
    up = torch.nn.functional.interpolate(gain[:, None], (gain.shape[1] - 1) * H + 1, mode="linear", align_corners=True)[:, 0]
    assert rel_rms(e, (x1[:, : y1.shape[1]] * up[:, : y1.shape[1]])) < 1e-3


# ----------------------------------------------------------------- GOLF-ss backward
def test_lpc_ss_gradients_reference_golden(G):
    g = golden("grads_ss")
    ex, gain, a = (T(g[k]).to(DEV).requires_grad_() for k in ("ex", "gain", "a"))
    y = G.lpc_ss(ex, gain, a, int(g["hop"]))
    assert rel_rms(y, T(g["ss_y"])) < 1e-5
    dex, dgain, da = torch.autograd.grad(y, (ex, gain, a), T(g["ss_up"]).to(DEV))
    assert rel_rms(dex, T(g["ss_dex"])) < REL_TOL
    assert rel_rms(dgain, T(g["ss_dgain"])) < REL_TOL
    assert rel_rms(da.flatten(1), T(g["ss_da"]).flatten(1)) < REL_TOL


def test_sample_wise_lpc_gradients_match_autograd_of_definition(G):
    """small case, float64 autograd through the literal recurrence as the truth"""
    B, Tn, M = 2, 64, 3
    g = torch.Generator().manual_seed(8)
    x = torch.randn(B, Tn, generator=g)
    A = 0.2 * torch.randn(B, Tn, M, generator=g)
    zi = torch.randn(B, M, generator=g)
    up = torch.randn(B, Tn, generator=g)
    xd, Ad, zd = (t.double().requires_grad_() for t in (x, A, zi))
    hist = [zd[:, j] for j in range(M)]  # hist[j] = y[t-1-j]
    ys = []
    for t in range(Tn):
        yt = xd[:, t] - sum(Ad[:, t, i] * hist[i] for i in range(M))
        ys.append(yt)
        hist = [yt] + hist[:-1]
    yref = torch.stack(ys, 1)
    gref = torch.autograd.grad(yref, (xd, Ad, zd), up.double())
    xg, Ag, zg = (t.to(DEV).requires_grad_() for t in (x, A, zi))
    y = G.sample_wise_lpc(xg, Ag, zg)
    assert rel_rms(y, yref) < 1e-5
    got = torch.autograd.grad(y, (xg, Ag, zg), up.to(DEV))
    for a_, b_ in zip(got, gref):
        assert rel_rms(a_.flatten(1), b_.flatten(1)) < 1e-5


# -------------------------------------------------------------------------- GOLF-ff
@pytest.mark.parametrize("M", [20, 22])
def test_lpc_ff_reference_golden(G, M):
    g = golden("filters_rand")
    H = int(g["hop"])
    y = G.lpc_ff(*cu(T(g[f"ex_{M}"]), T(g[f"gain_{M}"]), T(g[f"a_{M}"]), torch.hann_window(4 * H)), H)
    assert y.shape == g[f"ff_{M}"].shape
    assert rel_rms(y, T(g[f"ff_{M}"])) < 1e-5


@pytest.mark.parametrize("B,Tn,H,M", [(2, 48000, 240, 22), (2, 24000, 120, 22), (1, 9600, 240, 32), (1, 9600, 240, 12), (3, 9600, 240, 8), (1, 1000, 240, 22)])
def test_lpc_ff_matches_oracle(G, oracle, B, Tn, H, M):
    Fr = Tn // H + 1
    gain, a = synthetic_controls(B, Fr, M, seed=5)
    ex = torch.randn(B, Tn, generator=torch.Generator().manual_seed(2))
    ref = oracle.lpc_ff(ex, gain, a, H, 4 * H)
    y = G.lpc_ff(*cu(ex, gain, a, torch.hann_window(4 * H)), H)
    assert y.shape == ref.shape and rel_rms(y, ref) < REL_TOL


def test_lpc_ff_gradients_reference_golden(G):
    """adjoint of the frame-wise filter vs autograd of the reference module (torchaudio's
    DifferentiableIIR + conv_transpose1d + interpolate backward)"""
    g = golden("grads_ss")
    H = int(g["hop"])
    ex, gain, a = (T(g[k]).to(DEV).requires_grad_() for k in ("ex", "gain", "a"))
    y = G.lpc_ff(ex, gain, a, torch.hann_window(4 * H).to(DEV), H)
    assert rel_rms(y, T(g["ff_y"])) < 1e-5
    dex, dgain, da = torch.autograd.grad(y, (ex, gain, a), T(g["ff_up"]).to(DEV))
    assert rel_rms(dex, T(g["ff_dex"])) < REL_TOL
    assert rel_rms(dgain, T(g["ff_dgain"])) < REL_TOL
    assert rel_rms(da.flatten(1), T(g["ff_da"]).flatten(1)) < REL_TOL


def test_lpc_ff_gradients_ragged(G, oracle):
    """lengths that are not a multiple of the hop, hop 120, a different order: float64 autograd
    through a literal per-frame implementation as the truth"""
    # the last case: window 1024 / hop 256 / order 22 -- the padded order 24 divides neither, the adjoint runs at order 32
    for (Tn, H, M) in [(2500, 120, 12), (3001, 240, 20), (2100, 256, 22)]:
        B, W = 2, 4 * H
        Fr = Tn // H + 1
        gain, a = synthetic_controls(B, Fr, M, seed=3)
        gen = torch.Generator().manual_seed(9)
        ex = torch.randn(B, Tn, generator=gen)
        exd, gd, ad = (t.double().requires_grad_() for t in (ex, gain, a))
        up_g = torch.nn.functional.interpolate(gd[:, None], (Fr - 1) * H + 1, mode="linear", align_corners=True)[:, 0]
        n = min(Tn, up_g.shape[1])
        e = exd[:, :n] * up_g[:, :n]
        pad = torch.nn.functional.pad(e, (W // 2, W // 2))
        nf = (pad.shape[1] - W) // H + 1
        win = torch.hann_window(W, dtype=torch.float64)
        full = torch.zeros(B, (nf - 1) * H + W, dtype=torch.float64)
        norm = torch.zeros((nf - 1) * H + W, dtype=torch.float64)
        for k in range(nf):
            x = pad[:, k * H : k * H + W]
            hist = [torch.zeros(B, dtype=torch.float64)] * M
            outs = []
            for i in range(W):
                yv = x[:, i] - sum(ad[:, k, j] * hist[j] for j in range(M))
                outs.append(yv)
                hist = [yv] + hist[:-1]
            full[:, k * H : k * H + W] = full[:, k * H : k * H + W] + torch.stack(outs, 1) * win
            norm[k * H : k * H + W] += win
        ref = full[:, W // 2 : W // 2 + (nf - 1) * H] / norm[W // 2 : W // 2 + (nf - 1) * H]
        up = torch.randn(ref.shape, generator=gen)
        gref = torch.autograd.grad(ref, (exd, gd, ad), up.double())
        exg, gg, ag = (t.to(DEV).requires_grad_() for t in (ex, gain, a))
        y = G.lpc_ff(exg, gg, ag, torch.hann_window(W).to(DEV), H)
        assert y.shape == ref.shape and rel_rms(y, ref) < 1e-5
        got = torch.autograd.grad(y, (exg, gg, ag), up.to(DEV))
        for a_, b_ in zip(got, gref):
            assert rel_rms(a_.flatten(1), b_.flatten(1)) < REL_TOL


def test_biquad_cascade_reference_golden(G, oracle):
    g = golden("filters_rand")
    H = int(g["hop"])
    ex, gain, bq = T(g["ex_8"]), T(g["gain_8"]), T(g["biquads_8"])
    y = G.biquad_ff(*cu(ex, gain, bq, torch.hann_window(4 * H)), H)
    assert y.shape == g["bq_cascade_8"].shape
    assert rel_rms(y, T(g["bq_cascade_8"])) < REL_TOL
    assert rel_rms(y, oracle.biquad_ff(ex, gain, bq, H)) < REL_TOL


@pytest.mark.parametrize("M", [8, 20, 22])
def test_inverse_filter_reference_golden(G, M):
    g = golden("filters_rand")
    r = G.lpc_inverse(*cu(T(g[f"target_{M}"]), T(g[f"a_{M}"])), int(g["hop"]))
    assert rel_rms(r, T(g[f"inverse_{M}"])) < 1e-5


def test_inverse_filter_gradients_match_autograd_of_the_oracle(G, oracle):
    """adjoint of the analysis filter (inverse-target training, ltng/vocoder.py:192-198): d_y, d_a vs torch
    autograd through the CPU restatement; ragged length (y longer than the control range)"""
    g = golden("filters_rand")
    H, M = int(g["hop"]), 22
    a = T(g[f"a_{M}"])
    y = torch.cat([T(g[f"target_{M}"]), torch.randn(a.shape[0], 37, generator=torch.Generator().manual_seed(5))], 1)[:, : (a.shape[1] - 1) * H + 30]
    yr, ar = y.clone().requires_grad_(), a.clone().requires_grad_()
    ref = oracle.lpc_inverse(yr, ar, H)
    up = torch.randn(ref.shape, generator=torch.Generator().manual_seed(6))
    gy, ga = torch.autograd.grad(ref, (yr, ar), up)
    yg, ag = y.to(DEV).requires_grad_(), a.to(DEV).requires_grad_()
    r = G.lpc_inverse(yg, ag, H)
    assert r.shape == ref.shape and rel_rms(r, ref.detach()) < 1e-5
    d_y, d_a = torch.autograd.grad(r, (yg, ag), up.to(DEV))
    assert d_y.shape == y.shape and rel_rms(d_y, gy) < 1e-5
    assert rel_rms(d_a.flatten(1), ga.flatten(1)) < 1e-5


# ----------------------------------------------------------------------- FIR stages
@pytest.mark.parametrize("variant", ["ss", "ff"])
def test_fir_stages_reference_golden(G, oracle, variant):
    g = golden(f"stages_{variant}")
    H = int(g["hop"])
    kern = oracle.zero_phase_fir(T(g["log_mag"]))
    noise = T(g["noise"])[:, : g["harm"].shape[1]]
    y = G.ltv_fir_blocks(*cu(noise, kern), H)
    assert y.shape == g["noise_filtered"].shape and rel_rms(y, T(g["noise_filtered"])) < 1e-5
    y2 = G.ltv_fir_blocks(*cu(noise, kern), H, add=T(g["harm"]).to(DEV))
    assert rel_rms(y2, T(g["harm"])[:, : y.shape[1]] + T(g["noise_filtered"])) < 1e-5
    r = G.room_fir(*cu(T(g["lpc"]), T(g["room_kernel"])))
    assert rel_rms(r, T(g["out"])) < 1e-5


@pytest.mark.parametrize("T_,n", [(1, 127), (5, 3), (1023, 127), (1024, 127), (1025, 127), (4099, 1), (47760, 127), (9001, 300)])
def test_room_fir_ragged_sizes(G, oracle, T_, n):
    """LTIAcousticFilter on lengths around the kernel's 1024-output tile, odd lengths and tap counts other than the shipped 127"""
    gen = torch.Generator().manual_seed(T_ + n)
    x, k = torch.randn(3, T_, generator=gen), 0.1 * torch.randn(n, generator=gen)
    ref = oracle.room_fir(x, k)
    y = G.room_fir(*cu(x, k))
    assert y.shape == ref.shape and rel_rms(y, ref) < 1e-5


def test_noise_fir_ragged_sizes(G, oracle):
    g = torch.Generator().manual_seed(0)
    for (Tn, H, K, Fr) in [(4801, 240, 510, 25), (2000, 120, 254, 10), (700, 100, 62, 9)]:
        ex, kern = torch.randn(2, Tn, generator=g), 0.05 * torch.randn(2, Fr, K, generator=g)
        ref = oracle.ltv_fir_blocks(ex, kern, H)
        y = G.ltv_fir_blocks(*cu(ex, kern), H)
        assert y.shape == ref.shape and rel_rms(y, ref) < 1e-5


def test_noise_fir_packed_fp32_variant_is_bit_identical(G, oracle):
    """fma.rn.f32x2 kernel (16 outputs per lane, duplicated strip) vs the scalar register tile: every
    output sums its taps in the same order, so the results must agree bit for bit -- across hops
    (2, 4 and 1 blocks per warp; a hop that is not a multiple of 16), odd lengths, an `add` operand
    whose rows are not 16-byte aligned, and the fused fftshift + window staging."""
    from golf_b200 import _lib

    g = torch.Generator().manual_seed(3)
    cases = [(4801, 240, 510, 25, 0), (2000, 120, 254, 10, 1), (3000, 200, 510, 20, 0), (5000, 480, 510, 9, 3), (700, 100, 62, 9, 0)]
    try:
        for (Tn, H, K, Fr, off) in cases:
            ex, kern = torch.randn(3, Tn, generator=g), 0.05 * torch.randn(3, Fr, K, generator=g)
            add = torch.randn(3, Tn + 8, generator=g).to(DEV)[:, off : off + Tn]
            outs = []
            for mode in (0, 1):
                _lib.lib().golf_fir_set_variant(mode)
                outs.append((G.ltv_fir_blocks(*cu(ex, kern), H), G.ltv_fir_blocks(*cu(ex, kern), H, add=add),
                             G.ltv_fir_blocks(*cu(ex, kern), H, window=torch.hann_window(K).to(DEV))))
            for y0, y1 in zip(*outs):
                assert torch.equal(y0, y1), (Tn, H, K, Fr)
            assert rel_rms(outs[1][0], oracle.ltv_fir_blocks(ex, kern, H)) < 1e-5
    finally:
        _lib.lib().golf_fir_set_variant(1)


def test_fir_gradients_match_autograd_of_the_oracle(G, oracle):
    """adjoints of the block FIR and the room FIR vs torch autograd through the CPU restatement"""
    g = torch.Generator().manual_seed(12)
    B, Tn, H, K, Fr = 2, 2400, 240, 510, 11
    ex, kern = torch.randn(B, Tn, generator=g), 0.05 * torch.randn(B, Fr, K, generator=g)
    add = torch.randn(B, Tn, generator=g)
    exr, kr = ex.clone().requires_grad_(), kern.clone().requires_grad_()
    ref = oracle.ltv_fir_blocks(exr, kr, H)
    up = torch.randn(ref.shape, generator=g)
    g_ex, g_k = torch.autograd.grad(ref, (exr, kr), up)
    exg, kg, addg = ex.to(DEV).requires_grad_(), kern.to(DEV).requires_grad_(), add.to(DEV).requires_grad_()
    y = G.ltv_fir_blocks(exg, kg, H, add=addg)
    assert rel_rms(y, ref.detach() + add[:, : ref.shape[1]]) < 1e-5
    d_ex, d_k, d_add = torch.autograd.grad(y, (exg, kg, addg), up.to(DEV))
    assert rel_rms(d_ex, g_ex) < 1e-5 and rel_rms(d_k.flatten(1), g_k.flatten(1)) < 1e-5
    assert torch.equal(d_add[:, : up.shape[1]].cpu(), up) and not d_add[:, up.shape[1]:].any()
    # room
    x, k = torch.randn(B, 5000, generator=g), 0.05 * torch.randn(127, generator=g)
    xr, krr = x.clone().requires_grad_(), k.clone().requires_grad_()
    ref = oracle.room_fir(xr, krr)
    up = torch.randn(ref.shape, generator=g)
    g_x, g_kk = torch.autograd.grad(ref, (xr, krr), up)
    xg, kgg = x.to(DEV).requires_grad_(), k.to(DEV).requires_grad_()
    d_x, d_kk = torch.autograd.grad(G.room_fir(xg, kgg), (xg, kgg), up.to(DEV))
    assert rel_rms(d_x, g_x) < 1e-5 and rel_rms(d_kk[None], g_kk[None]) < 1e-5


def test_noise_fir_fused_shift_and_window(G, oracle):
    g = golden("stages_ss")
    H = int(g["hop"])
    raw = torch.fft.irfft(torch.exp(T(g["log_mag"])) + 0j, dim=-1)
    noise = T(g["noise"])[:, : g["harm"].shape[1]]
    y = G.ltv_fir_blocks(*cu(noise, raw), H, window=torch.hann_window(510).to(DEV))
    assert rel_rms(y, T(g["noise_filtered"])) < 1e-5


# ----------------------------------------------------------------------- oscillator
def test_oscillator_reference_golden(G, oracle):
    g = golden("stages_ss")
    table, _ = oracle.glottal_table()
    dk = oracle.decimate_kernel(4)
    ph, w = T(g["phase"]), T(g["w"])
    args = (int(g["phase_hop"]), w.to(DEV), int(g["w_hop"]), table.to(DEV), dk.to(DEV), 4, True)
    y = G.glottal_osc(ph.to(DEV), *args, "aten_cpu")  # the reference's CPU phase arithmetic
    assert y.shape == g["harm"].shape and rel_rms(y, T(g["harm"])) < 1e-5
    y64 = G.glottal_osc(ph.to(DEV), *args, "fp64")
    truth = oracle.glottal_osc(ph, int(g["phase_hop"]), w, int(g["w_hop"]), table, 4, True, "fp64")
    assert rel_rms(y64, truth) < 1e-5
    # default mode is closer to exact phase than the reference is, and within the bar of it
    assert rel_rms(y64, truth) < rel_rms(T(g["harm"]), truth)
    assert rel_rms(y64, T(g["harm"])) < REL_TOL


def test_oscillator_sample_rate_f0(G, oracle):
    """training-mode input: per-sample f0 (ltng/ae.py:96-101), 2 utterances x 1 s"""
    B, Tn = 2, 24000
    gen = torch.Generator().manual_seed(4)
    f0 = (200 + 0.8 * torch.cumsum(torch.randn(B, Tn, generator=gen), 1)).clamp(80, 400)
    ph, w = f0 / 24000, torch.rand(B, Tn // 2400 + 1, generator=gen)
    table, _ = oracle.glottal_table()
    dk = oracle.decimate_kernel(4)
    for mode, omode in (("aten_cpu", "fp32"), ("fp64", "fp64")):
        y = G.glottal_osc(ph.to(DEV), 1, w.to(DEV), 2400, table.to(DEV), dk.to(DEV), 4, True, mode)
        ref = oracle.glottal_osc(ph, 1, w, 2400, table, 4, True, omode)
        assert y.shape == ref.shape == (B, Tn) and rel_rms(y, ref) < 2e-5


def test_oscillator_weight_gradient(G, oracle):
    """d/dw of the oscillator vs torch autograd through the CPU restatement (float64 phase)"""
    B, Tn = 2, 9600
    gen = torch.Generator().manual_seed(14)
    f0 = (180 + 0.5 * torch.cumsum(torch.randn(B, Tn, generator=gen), 1)).clamp(80, 400)
    ph = f0 / 24000
    w = (0.2 + 0.6 * torch.rand(B, Tn // 2400 + 1, generator=gen))
    table, _ = oracle.glottal_table()
    dk = oracle.decimate_kernel(4)
    wr = w.clone().requires_grad_()
    ref = oracle.glottal_osc(ph, 1, wr, 2400, table, 4, True, "fp64")
    up = torch.randn(ref.shape, generator=gen)
    (g_ref,) = torch.autograd.grad(ref, wr, up)
    wg = w.to(DEV).requires_grad_()
    y = G.glottal_osc(ph.to(DEV), 1, wg, 2400, table.to(DEV), dk.to(DEV), 4, True, "exact")
    assert rel_rms(y, ref.detach()) < 2e-5
    (g_w,) = torch.autograd.grad(y, wg, up.to(DEV))
    assert rel_rms(g_w, g_ref) < 1e-3  # sums of ~1e4 terms in a different order, float atomics


@pytest.mark.parametrize("os_,phase_hop,equal_energy", [(4, 1, True), (4, 240, False), (2, 1, True), (1, 1, False)])
def test_oscillator_table_gradient(G, oracle, os_, phase_hop, equal_energy):
    """d/dtable of the oscillator (GlottalFlowTable(trainable=True), models/synth.py:117-118) together with d/dw, vs torch
    autograd through the CPU restatement (float64 phase); both gradients from one backward"""
    B, Tn = 3, 9600
    gen = torch.Generator().manual_seed(15 + os_)
    f0 = (180 + 0.5 * torch.cumsum(torch.randn(B, Tn // phase_hop + (phase_hop > 1), generator=gen), 1)).clamp(80, 400)
    ph = f0 / 24000
    w = torch.rand(B, Tn // 2400 + 1, generator=gen)
    w[0, 0], w[1, -1] = 0.0, 1.0  # both ends of the table
    table, _ = oracle.glottal_table()
    dk = oracle.decimate_kernel(os_) if os_ > 1 else None
    wr, tr = w.clone().requires_grad_(), table.clone().requires_grad_()
    ref = oracle.glottal_osc(ph, phase_hop, wr, 2400, tr, os_, equal_energy, "fp64")
    up = torch.randn(ref.shape, generator=gen)
    gw_ref, gt_ref = torch.autograd.grad(ref, (wr, tr), up)
    wg, tg = w.to(DEV).requires_grad_(), table.to(DEV).requires_grad_()
    y = G.glottal_osc(ph.to(DEV), phase_hop, wg, 2400, tg, None if dk is None else dk.to(DEV), os_, equal_energy, "exact")
    assert rel_rms(y, ref.detach()) < 2e-5
    g_w, g_t = torch.autograd.grad(y, (wg, tg), up.to(DEV))
    assert g_t.shape == table.shape and rel_rms(g_w, gw_ref) < 1e-3
    assert rel_rms(g_t, gt_ref) < 1e-3  # bilinear scatter; the column split of a sample at a cell edge moves with the phase's last bit
    untouched = gt_ref == 0
    assert torch.equal(g_t.cpu()[untouched], gt_ref[untouched])  # rows no frame selected stay exactly zero
    # table gradient alone (w detached): same numbers
    y2 = G.glottal_osc(ph.to(DEV), phase_hop, w.to(DEV), 2400, tg, None if dk is None else dk.to(DEV), os_, equal_energy, "exact")
    (g_t2,) = torch.autograd.grad(y2, tg, up.to(DEV))
    assert rel_rms(g_t2, g_t) < 1e-5


def test_oscillator_table_gradient_unsupported_modes_fail_loudly(G, oracle):
    from golf_b200 import GolfError

    table, _ = oracle.glottal_table()
    tg = table.to(DEV).requires_grad_()
    ph, w = torch.full((1, 4800), 200 / 24000), torch.rand(1, 3)
    y = G.glottal_osc(ph.to(DEV), 1, w.to(DEV), 2400, tg, oracle.decimate_kernel(4).to(DEV), 4, True, "aten_cpu")
    with pytest.raises(GolfError, match="table gradient"):
        y.sum().backward()


def test_wavetable_read_matches_generate(G, oracle):
    gen = torch.Generator().manual_seed(6)
    table, _ = oracle.glottal_table()
    w = torch.rand(2, 4, generator=gen)
    tabs = oracle.select_tables(table, w)
    wr = torch.rand(2, 20000, generator=gen)
    assert rel_rms(G.wavetable_read(*cu(wr, tabs), 9600), oracle.wavetable_read(wr, tabs, 9600)) < 1e-5


# -------------------------------------------------------------------------- helpers
def test_linear_upsample_bit_exact(G, oracle):
    x = torch.randn(5, 201, generator=torch.Generator().manual_seed(0))
    for hop in (240, 120, 7):
        assert torch.equal(G.linear_upsample(x.to(DEV), hop).cpu(), oracle.upsample_time(x, hop))


def test_rc2lpc_kernel(G, oracle):
    lg = 0.5 * torch.randn(300, 22, generator=torch.Generator().manual_seed(0))
    a = G.rc2lpc(lg.to(DEV)).cpu()
    ref = oracle.rc2lpc(torch.tanh(lg))
    assert (a - ref).abs().max() < 5e-6 * ref.abs().max()


@pytest.mark.parametrize("M", [1, 2, 22, 40])
def test_rc2lpc_gradient_matches_torch_autograd(G, oracle, M):
    """adjoint of the step-up recursion vs autograd through the torch restatement (models/utils.py:581-593)"""
    from golf_b200.utils import rc2lpc as rc2lpc_torch

    g = torch.Generator().manual_seed(M)
    lg = (0.5 * torch.randn(3, 50, M, generator=g)).to(DEV)
    up = torch.randn(3, 50, M, generator=g).to(DEV)
    l1, l2 = lg.clone().requires_grad_(), lg.clone().requires_grad_()
    a1 = G.rc2lpc(l1, 0.99)
    a2 = rc2lpc_torch(torch.tanh(l2) * 0.99)
    assert rel_rms(a1.flatten(1), a2.flatten(1)) < 1e-6
    (d1,), (d2,) = torch.autograd.grad(a1, l1, up), torch.autograd.grad(a2, l2, up)
    assert rel_rms(d1.flatten(1), d2.flatten(1)) < 1e-5
