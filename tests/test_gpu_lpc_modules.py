"""GPU: the frame-wise LPC synthesis row (models/lpc.py) and the biquad parameterisations (models/utils.py:444-525)
through the C ABI against the REFERENCE's own outputs and autograd gradients (tests/golden/lpc_modules.npz, written
by tests/golden/make_golden_lpc_modules.py from the unmodified reference on the CPU).

Bar: relative RMS <= 1e-4 (north_star); forward results sit near 1e-6, gradients near 1e-5."""
import pytest
import torch

from conftest import REL_TOL, T, golden, rel_rms

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def G():
    from golf_b200 import functional

    return functional


@pytest.fixture(scope="module")
def g():
    return golden("lpc_modules")


def cu(*ts, grad=False):
    return [t.to(DEV).requires_grad_(grad) for t in ts]


def flat(t):
    return t.reshape(1, -1)


# ---------------------------------------------------------------- a12: lpc_synthesis (lfilter-shaped twin)
def test_lpc_synthesis_forward_and_gradients(G, g):
    x, gains, a = cu(T(g["ls_x"]), T(g["ls_gains"]), T(g["ls_a"]), grad=True)
    y = G.lpc_synthesis(x, gains, a)
    assert rel_rms(y, T(g["ls_y"])) < 1e-5
    dx, dg, da = torch.autograd.grad(y, (x, gains, a), T(g["ls_g"]).to(DEV))
    assert rel_rms(dx, T(g["ls_dx"])) < REL_TOL
    assert rel_rms(flat(dg), flat(T(g["ls_dgains"]))) < REL_TOL
    assert rel_rms(da, T(g["ls_da"])) < REL_TOL


def test_lpc_synthesis_views_and_orders(G, oracle):
    """strided rows (an `unfold` view, as models/lpc.py:36 passes), N shorter than a tile, M in {1, 4, 40}"""
    gen = torch.Generator().manual_seed(3)
    for C, N, M in ((5, 3, 1), (33, 100, 4), (64, 960, 40), (3, 17, 22)):
        sig = torch.randn(N + 7 * (C - 1), generator=gen)
        x = sig.unfold(0, N, 7)  # [C, N], row stride 7 < N -> the wrapper must make it contiguous
        a = oracle.rc2lpc(torch.tanh(0.3 * torch.randn(1, C, M, generator=gen)))[0]
        gains = torch.rand(C, generator=gen) + 0.5
        y = G.lpc_synthesis(x.to(DEV), gains.to(DEV), a.to(DEV))
        ref = oracle.allpole_lti(x.contiguous() * gains[:, None], a)
        assert rel_rms(y, ref) < 1e-5, (C, N, M)


# ---------------------------------------------------------------- BatchLPCSynth / LPCSynth
@pytest.mark.parametrize("M", [8, 22])
def test_batch_lpc_synth_reference_golden(g, M):
    from golf_b200.lpc import BatchLPCSynth

    p = f"bl{M}_"
    H = int(g["hop"])
    mod = BatchLPCSynth(hop_length=H, window="hanning").to(DEV)
    ex, gain, a = cu(T(g[p + "ex"]), T(g[p + "gain"]), T(g[p + "a"]), grad=True)
    y = mod(ex, gain, a)
    assert y.shape == g[p + "y"].shape
    assert rel_rms(y, T(g[p + "y"])) < 1e-5
    dex, dgain, da = torch.autograd.grad(y, (ex, gain, a), T(g[p + "g"]).to(DEV))
    assert rel_rms(dex, T(g[p + "dex"])) < REL_TOL
    assert rel_rms(dgain, T(g[p + "dgain"])) < REL_TOL
    assert rel_rms(flat(da), flat(T(g[p + "da"]))) < REL_TOL


def test_lpc_synth_single_utterance(g):
    from golf_b200.lpc import LPCSynth

    H = int(g["hop"])
    mod = LPCSynth(hop_length=H, window="hanning").to(DEV)
    lpc = torch.cat([T(g["bl8_gain"])[0, :, None], T(g["bl8_a"])[0]], -1)
    y = mod(T(g["bl8_ex"])[0].to(DEV), lpc.to(DEV))
    assert y.shape == g["lp8_y"].shape
    assert rel_rms(flat(y), flat(T(g["lp8_y"]))) < 1e-5


def test_lpc_modules_state_dict_keys():
    """`_kernel` [win,1,win] is the reference's only buffer (models/lpc.py:30-32)"""
    from golf_b200.lpc import BatchSecondOrderLPCSynth

    sd = BatchSecondOrderLPCSynth(hop_length=120).state_dict()
    assert list(sd) == ["_kernel"] and tuple(sd["_kernel"].shape) == (480, 1, 480)


# ---------------------------------------------------------------- a14: the cascade and its adjoint
@pytest.mark.parametrize("K", [4, 11])
def test_biquad_cascade_forward_and_gradients(g, K):
    from golf_b200.lpc import BatchSecondOrderLPCSynth

    p = f"bq{K}_"
    H = int(g["hop"])
    mod = BatchSecondOrderLPCSynth(hop_length=H, window="hanning").to(DEV)
    ex, gain, bq = cu(T(g[p + "ex"]), T(g[p + "gain"]), T(g[p + "biquads"]), grad=True)
    y = mod(ex, gain, bq)
    assert rel_rms(y, T(g[p + "y"])) < 5e-5  # K cascaded resonators: float32 rounding-order differences reach 1e-5
    dex, dgain, dbq = torch.autograd.grad(y, (ex, gain, bq), T(g[p + "g"]).to(DEV))
    assert rel_rms(dex, T(g[p + "dex"])) < REL_TOL
    assert rel_rms(dgain, T(g[p + "dgain"])) < REL_TOL
    assert rel_rms(flat(dbq), flat(T(g[p + "dbiquads"]))) < REL_TOL
    # control frames beyond the last signal frame receive exactly zero
    n_frames = y.shape[1] // H
    assert float(dbq[:, n_frames:].abs().max()) == 0.0 and float(dgain[:, n_frames:].abs().max()) == 0.0


def test_biquad_cascade_unnormalised_a0(G, oracle):
    """lfilter divides by a0 (models/lpc.py:118 passes the sections as given): a0 != 1 and its gradient"""
    gen = torch.Generator().manual_seed(5)
    H, Tn, K = 64, 1024, 3
    Fr = Tn // H
    ex = torch.randn(2, Tn, generator=gen)
    gain = torch.rand(2, Fr, generator=gen) + 0.5
    bq = oracle.logits2biquads(0.5 * torch.randn(2, Fr, K, 2, generator=gen), "coef", 0.95)
    a0 = 0.5 + torch.rand(2, Fr, K, 1, generator=gen)
    bq = bq * a0
    win = torch.hann_window(4 * H)
    up = torch.randn(2, Tn, generator=gen)
    exr, gainr, bqr = [t.clone().double().requires_grad_() for t in (ex, gain, bq)]

    def torch_ref(e, gn, q):  # the definition in float64, autograd through an explicit loop
        W, pad = 4 * H, (4 * H - H) // 2
        fr = torch.nn.functional.pad(e, (pad, pad)).unfold(1, W, H) * gn[:, :, None]
        y = fr
        for j in range(K):
            c0, c1, c2 = q[:, :, j, 0], q[:, :, j, 1], q[:, :, j, 2]
            prev1 = torch.zeros_like(c0)
            prev2 = torch.zeros_like(c0)
            cols = []
            for n in range(W):
                v = (y[:, :, n] - c1 * prev1 - c2 * prev2) / c0
                cols.append(v)
                prev2, prev1 = prev1, v
            y = torch.stack(cols, -1)
        w = win.double()
        yw = (y * w).transpose(1, 2)
        ola = torch.nn.functional.fold(yw, (1, (Fr - 1) * H + W), (1, W), stride=(1, H))[:, 0, 0]
        nrm = torch.nn.functional.fold(w[None, :, None].expand(1, W, Fr), (1, (Fr - 1) * H + W), (1, W), stride=(1, H))[:, 0, 0]
        return ola[:, pad:-pad] / nrm[:, pad:-pad]  # crop first: the window is 0 at the very first sample

    ref = torch_ref(exr, gainr, bqr)
    rdex, rdgain, rdbq = torch.autograd.grad(ref, (exr, gainr, bqr), up.double())
    e, gn, q = cu(ex, gain, bq, grad=True)
    y = G.biquad_ff(e, gn, q, win.to(DEV), H)
    assert rel_rms(y, ref) < 1e-5
    dex, dgain, dbq = torch.autograd.grad(y, (e, gn, q), up.to(DEV))
    assert rel_rms(dex, rdex) < REL_TOL
    assert rel_rms(dgain, rdgain) < REL_TOL
    assert rel_rms(flat(dbq), flat(rdbq)) < REL_TOL


# ---------------------------------------------------------------- a16: parameterisations + polynomial product
@pytest.mark.parametrize("rep", ["coef", "conj", "real"])
def test_biquad_params_kernel_reference_golden(G, g, rep):
    lg = T(g["pm_logits"]).to(DEV).requires_grad_()
    bq = G.logits2biquads(lg, rep, 0.99)
    a = G.logits2lpc(lg, rep, 0.99)
    assert torch.allclose(bq.cpu(), T(g[f"pm_{rep}_biquads"]), atol=2e-6)
    ref_a = T(g[f"pm_{rep}_a"])
    assert float((a.detach().cpu() - ref_a).abs().max()) <= 2e-6 * float(ref_a.abs().max())
    (d_a,) = torch.autograd.grad(a, lg, T(g["pm_g_a"]).to(DEV))
    (d_bq,) = torch.autograd.grad(bq, lg, T(g["pm_g_bq"]).to(DEV))
    assert rel_rms(flat(d_a), flat(T(g[f"pm_{rep}_dlogits_a"]))) < 2e-5
    assert rel_rms(flat(d_bq), flat(T(g[f"pm_{rep}_dlogits_bq"]))) < 2e-5


def test_filter_ctrl_uses_the_kernel_for_biquad_parameterisations(g):
    """LTVMinimumPhaseFilter(lpc_parameterisation="coef") -- the ISMIR-23 end filter (models/filters.py:73-78) --
    transforms its logits with one launch and matches the reference's a"""
    from golf_b200 import _lib
    from golf_b200.filters import LTVMinimumPhaseFilter

    f = LTVMinimumPhaseFilter(window="hanning", window_length=480, centred=False, lpc_order=22, lpc_parameterisation="coef",
                              max_abs_value=0.99)
    from golf_b200.audiotensor import AudioTensor

    lg = AudioTensor(T(g["pm_logits"]).reshape(2, 40, 22).to(DEV), hop_length=120)
    log_gain = AudioTensor(torch.zeros(2, 40, device=DEV), hop_length=120)
    sizes, trsfms = f.ctrl(lambda s, t: (s, t))((), ())
    assert sizes == ((1, 22),)
    n0 = _lib.launch_count()
    gain, a = trsfms[0](log_gain, lg)
    assert _lib.launch_count() - n0 == 1
    ref_a = T(g["pm_coef_a"])
    assert float((torch.as_tensor(a).cpu() - ref_a).abs().max()) <= 2e-6 * float(ref_a.abs().max())
