"""CPU: pin the oracle (oracle/golf_oracle.py + .c) against the reference's outputs.

* committed golden vectors produced by the UNMODIFIED reference (tests/golden/make_golden.py)
* bit-exactness against the installed third-party ops the reference calls (torchaudio lfilter,
  ATen linear interpolation)
* live comparison with the reference modules when /root/reference is present
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import REL_TOL, T, golden, rel_rms, smooth, synthetic_controls


def test_lti_matches_torchaudio_bitwise(oracle):
    from torchaudio.functional import lfilter

    g = torch.Generator().manual_seed(0)
    _, a = synthetic_controls(1, 7, 22, seed=3)
    a = a[0]
    x = torch.randn(7, 960, generator=g)
    A = torch.cat([torch.ones(7, 1), a], 1)
    Bc = torch.zeros(7, 23)
    Bc[:, 0] = 1
    ref = lfilter(x, A, Bc, False)
    assert torch.equal(oracle.allpole_lti(x, a), ref)


def test_upsample_matches_aten_bitwise(oracle):
    x = torch.randn(3, 201, generator=torch.Generator().manual_seed(1))
    for hop in (240, 120, 4, 480):
        ref = F.interpolate(x[:, None], (201 - 1) * hop + 1, mode="linear", align_corners=True)[:, 0]
        assert torch.equal(oracle.linear_upsample_c(x, hop), ref)


def test_ss_fused_equals_materialised(oracle):
    gain, a = synthetic_controls(2, 41, 22, seed=5)
    ex = torch.randn(2, 9700, generator=torch.Generator().manual_seed(2))
    assert torch.equal(oracle.lpc_ss(ex, gain, a, 240), oracle.lpc_ss_fused(ex, gain, a, 240))


def test_ss_constant_coefficients_equal_lfilter(oracle):
    """cross-check (i) of SURVEY 8c: time-invariant coefficients must reduce to lfilter"""
    from torchaudio.functional import lfilter

    _, a = synthetic_controls(2, 1, 22, seed=7)
    x = torch.randn(2, 5000, generator=torch.Generator().manual_seed(3))
    A = a.expand(2, 5000, 22).contiguous()
    y = oracle.sample_wise_lpc(x, A)
    ref = lfilter(x, torch.cat([torch.ones(2, 1), a[:, 0]], 1), F.pad(torch.ones(2, 1), (0, 22)), False)
    assert rel_rms(y, ref) < 2e-5  # same maths, opposite tap order: float32 rounding only


def test_order1_is_leaky_integrator(oracle):
    """cross-check (ii): sample_wise_lpc(u, -lambda, zi) == h_t = lambda_t h_{t-1} + u_t (lru.py:9-15)"""
    g = torch.Generator().manual_seed(4)
    u = torch.randn(2, 300, generator=g, dtype=torch.float64)
    lam = torch.rand(2, 300, 1, generator=g, dtype=torch.float64) * 0.9
    h0 = torch.randn(2, 1, generator=g, dtype=torch.float64)
    y = oracle.sample_wise_lpc(u, -lam, h0)
    h = h0[:, 0].clone()
    for t in range(300):
        h = lam[:, t, 0] * h + u[:, t]
        assert torch.allclose(y[:, t], h, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("M", [8, 20, 22])
def test_filters_against_reference_golden(oracle, M):
    g = golden("filters_rand")
    H = int(g["hop"])
    ex, gain, a = T(g[f"ex_{M}"]), T(g[f"gain_{M}"]), T(g[f"a_{M}"])
    assert torch.equal(oracle.lpc_ss(ex, gain, a, H), T(g[f"ss_{M}"]))  # same C loop as the golden run's stub
    assert rel_rms(oracle.lpc_ff(ex, gain, a, H, 4 * H), T(g[f"ff_{M}"])) < 1e-6
    assert rel_rms(oracle.lpc_ff(ex, gain, a, H, 4 * H, centred=False), T(g[f"ffnc_{M}"])) < 1e-6
    assert rel_rms(oracle.lpc_inverse(T(g[f"target_{M}"]), a, H), T(g[f"inverse_{M}"])) < 1e-6


def test_biquad_cascade_against_reference_golden(oracle):
    g = golden("filters_rand")
    y = oracle.biquad_ff(T(g["ex_8"]), T(g["gain_8"]), T(g["biquads_8"]), int(g["hop"]))
    # 4 near-coincident pole pairs at radius 0.99: float32 itself is only good to ~1e-3 here
    assert rel_rms(y, T(g["bq_cascade_8"])) < 5e-3
    assert y.shape == g["bq_cascade_8"].shape


@pytest.mark.parametrize("variant", ["ss", "ff"])
def test_decoder_stages_against_reference_golden(oracle, variant):
    g = golden(f"stages_{variant}")
    table, Rd = oracle.glottal_table()
    tb = golden("table")
    assert np.abs(table[::9].numpy() - tb["table_rows"]).max() < 2e-6
    assert np.abs(Rd.numpy() - tb["R_d_values"]).max() == 0
    st = {}
    out = oracle.source_filter_synth(
        T(g["phase"]), int(g["phase_hop"]), T(g["w"]), int(g["w_hop"]), T(g["log_mag"]), T(g["gain"]), T(g["a"]),
        int(g["hop"]), T(g["noise"]), table, T(g["room_kernel"]), variant=variant, stages=st)
    assert rel_rms(st["harm"], T(g["harm"])) < 1e-6
    assert rel_rms(st["noise_filtered"], T(g["noise_filtered"])) < 2e-6
    # identical-input stage checks (the stage's own golden input)
    src = T(g["harm"])[:, : g["noise_filtered"].shape[1]] + T(g["noise_filtered"])
    lpc = oracle.lpc_ss(src, T(g["gain"]), T(g["a"]), int(g["hop"])) if variant == "ss" else \
        oracle.lpc_ff(src, T(g["gain"]), T(g["a"]), int(g["hop"]), int(g["window_length"]))
    assert rel_rms(lpc, T(g["lpc"])) < 1e-6
    assert rel_rms(oracle.room_fir(T(g["lpc"]), T(g["room_kernel"])), T(g["out"])) < 1e-6
    # end to end (errors of the early stages pass through a resonant filter)
    assert rel_rms(out, T(g["out"])) < REL_TOL
    assert out.shape == g["out"].shape


def test_control_transforms(oracle):
    from golf_b200 import utils as U

    lg = torch.randn(2, 9, 22, generator=torch.Generator().manual_seed(0))
    assert torch.equal(oracle.rc2lpc(torch.tanh(lg)), U.rc2lpc(torch.tanh(lg)))
    bq_l = torch.randn(2, 9, 4, 2, generator=torch.Generator().manual_seed(1))
    for rep in ("coef", "conj", "real"):
        bq = oracle.logits2biquads(bq_l, rep, 0.99)
        assert torch.allclose(bq, U.get_logits2biquads(rep, 0.99)(bq_l), atol=1e-7)
        assert torch.allclose(oracle.biquads2lpc(bq), U.biquads2lpc(bq), atol=2e-6)


@pytest.mark.reference
def test_oracle_against_live_reference(oracle, reference):
    """build container only: the restatement vs the reference modules on fresh inputs"""
    from models.audiotensor import AudioTensor
    from models.filters import LTIAcousticFilter, LTVMinimumPhaseFilter, LTVMinimumPhaseFilterPrecise, LTVZeroPhaseFIRFilter
    from models.utils import rc2lpc

    H, M, B, Tn = 120, 12, 2, 6000
    Fr = Tn // H + 1
    g = torch.Generator().manual_seed(11)
    a = rc2lpc(torch.tanh(0.15 * smooth(torch.randn(B, Fr, M, generator=g))))
    gain = torch.exp(smooth(torch.randn(B, Fr, generator=g)) - 6)
    ex = torch.randn(B, Tn, generator=g)
    A = (AudioTensor(ex), AudioTensor(gain, hop_length=H), AudioTensor(a, hop_length=H))
    with torch.no_grad():
        assert torch.equal(LTVMinimumPhaseFilterPrecise(lpc_order=M)(*A).as_tensor(), oracle.lpc_ss(ex, gain, a, H))
        ff = LTVMinimumPhaseFilter(window="hanning", window_length=4 * H, lpc_order=M)(*A).as_tensor()
        assert rel_rms(oracle.lpc_ff(ex, gain, a, H, 4 * H), ff) < 1e-6
        lm = smooth(torch.randn(B, Fr, 65, generator=g)) - 4
        nf = LTVZeroPhaseFIRFilter(window="hanning", n_mag=65)(AudioTensor(ex), AudioTensor(lm, hop_length=H)).as_tensor()
        assert rel_rms(oracle.noise_fir(ex, lm, H), nf) < 2e-6
        room = LTIAcousticFilter(128, "fft")
        room.kernel.data = torch.randn(127, generator=g) * 0.05
        assert rel_rms(oracle.room_fir(ex, room.kernel.data), room(AudioTensor(ex)).as_tensor()) < 1e-6


def test_precise_fir_oracle_matches_reference_golden(oracle):
    """LTVZeroPhaseFIRFilterPrecise (models/filters.py:286-337): restatement vs the reference module's own
    output (tests/golden/make_golden_precise_fir.py), long and short inputs"""
    g = golden("fir_precise")
    H = int(g["hop"])
    for tag in ("long", "short"):
        y = oracle.noise_fir_precise(T(g[f"ex_{tag}"]), T(g["log_mag"]), H)
        assert y.shape == g[f"y_{tag}"].shape
        assert rel_rms(y, T(g[f"y_{tag}"])) < 2e-6


# ---------------------------------------------------------------- independent ss restatements
def test_ss_three_independent_restatements_agree(oracle):
    """The GOLF-ss recurrence is third-party code that cannot be installed here (PARITY UNPINNED at the
    source).  Three implementations that share no code -- the C loop, a LAPACK banded solve of the
    defining linear system, and the padded-buffer in-place shape of torchlpc's numba kernel -- must agree:
    float64 to ~1e-12, float32 to the rounding floor, with and without an initial state."""
    from oracle import ss_independent as S

    g = golden("controls_gt")
    H = int(g["hop"])
    a_f, gain_f = T(g["a"])[:2, :21], T(g["gain"])[:2, :21]  # encoder-derived (pole radius up to 0.996)
    Tn = 20 * H
    A = oracle.upsample_time(a_f, H)[:, :Tn].contiguous()
    x = torch.randn(2, Tn, generator=torch.Generator().manual_seed(0)) * oracle.upsample_time(gain_f, H)[:, :Tn]
    zi = 1e-3 * torch.randn(2, A.shape[2], generator=torch.Generator().manual_seed(1))
    for z in (None, zi):
        zd = None if z is None else z.double()
        c64 = oracle.sample_wise_lpc(x.double(), A.double(), zd)
        banded = torch.from_numpy(S.sample_wise_lpc_banded(x.numpy(), A.numpy(), None if z is None else z.numpy()))
        padded64 = torch.from_numpy(S.sample_wise_lpc_padded(x.double().numpy(), A.double().numpy(), None if zd is None else zd.numpy()))
        assert rel_rms(c64, banded) < 1e-9  # the fp64 difference is the conditioning of the banded solve
        assert rel_rms(padded64, banded) < 1e-9
        c32 = oracle.sample_wise_lpc(x, A, z)
        padded32 = torch.from_numpy(S.sample_wise_lpc_padded(x.numpy(), A.numpy(), None if z is None else z.numpy()))
        floor = rel_rms(c32, banded)
        assert floor < REL_TOL and rel_rms(padded32, banded) < REL_TOL
        assert rel_rms(padded32, c32) < 4 * floor + 1e-7
    # the numba and the plain-NumPy versions of the padded shape are the same arithmetic
    small = slice(0, 600)
    p_np = S.sample_wise_lpc_padded(x[:, small].numpy(), A[:, small].numpy(), zi.numpy(), use_numba=False)
    p_nb = S.sample_wise_lpc_padded(x[:, small].numpy(), A[:, small].numpy(), zi.numpy(), use_numba=True)
    assert np.abs(p_np - p_nb).max() <= 1e-6 * np.abs(p_np).max()


def test_ss_zi_ordering_split_continuation(oracle):
    """`zi[:, j] = y[-1-j]` (recalled from the package, unpinned): whatever the convention, filtering a
    signal in two pieces -- the second started from the last M outputs of the first, newest first -- must
    reproduce the one-shot result exactly; any other ordering of zi breaks this identity."""
    from oracle import ss_independent as S

    g = torch.Generator().manual_seed(5)
    B, Tn, M, cut = 2, 700, 6, 333
    A = oracle.rc2lpc(torch.tanh(0.3 * smooth(torch.randn(B, Tn, M, generator=g), 32))).double()
    x = torch.randn(B, Tn, generator=g, dtype=torch.float64)
    full = oracle.sample_wise_lpc(x, A)
    zi = full[:, cut - M : cut].flip(1)  # newest first
    tail = oracle.sample_wise_lpc(x[:, cut:].contiguous(), A[:, cut:].contiguous(), zi.contiguous())
    assert torch.equal(tail, full[:, cut:])
    wrong = oracle.sample_wise_lpc(x[:, cut:].contiguous(), A[:, cut:].contiguous(), zi.flip(1).contiguous())
    assert not torch.allclose(wrong, full[:, cut:])
    for fn in (S.sample_wise_lpc_banded, S.sample_wise_lpc_padded):
        t2 = torch.from_numpy(fn(x[:, cut:].numpy(), A[:, cut:].numpy(), zi.numpy()))
        assert rel_rms(t2, full[:, cut:]) < 1e-10
