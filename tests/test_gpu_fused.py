"""GPU: the fused noise branch (csrc/fir_fused.cu: in-kernel FIR design, optional in-kernel white noise) and the
one-call decoder (golf_synth_fused_fwd) against the oracle and against the module-by-module path they replace."""
import pytest
import torch

from conftest import REL_TOL, T, golden, rel_rms, smooth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def G():
    from golf_b200 import functional

    return functional


def hann(K):
    return torch.hann_window(K, device=DEV)


@pytest.mark.parametrize("B,Tn,Fr,hop", [(2, 9600, 41, 240), (3, 12001, 60, 240), (1, 4800, 12, 240), (2, 48000, 200, 240),
                                         (2, 7000, 40, 200), (1, 2000, 40, 256), (32, 48000, 200, 240)])
def test_fir_design_matches_oracle_and_fft_route(G, oracle, B, Tn, Fr, hop):
    """a8 + f1: taps from log_mag inside the FIR kernel == irfft route (oracle: torch.fft on the CPU, the reference's
    arithmetic) and == the library's own cuFFT route"""
    g = torch.Generator().manual_seed(Fr + hop)
    lm = smooth(torch.randn(B, Fr, 256, generator=g)) - 4
    ex = torch.randn(B, Tn, generator=g)
    add = torch.randn(B, Tn, generator=g)
    ref = oracle.noise_fir(ex, lm, hop)
    y = G.noise_fir_design(ex.to(DEV), lm.to(DEV), hann(510), hop)
    assert y.shape == ref.shape
    assert rel_rms(y, ref) < 2e-6
    raw = torch.fft.irfft(G.exp_complex(lm.to(DEV)), dim=-1, norm="forward")
    y_fft = G.ltv_fir_blocks(ex.to(DEV), raw, hop, None, window=hann(510) / 510)
    assert rel_rms(y, y_fft) < 2e-6
    ya = G.noise_fir_design(ex.to(DEV), lm.to(DEV), hann(510), hop, add=add.to(DEV))
    assert torch.equal(ya, add.to(DEV)[:, : y.shape[1]] + y)


def test_fir_design_reference_golden(G):
    g = golden("stages_ss")
    y = G.noise_fir_design(T(g["noise"]).to(DEV)[:, : g["harm"].shape[1]], T(g["log_mag"]).to(DEV), hann(510), int(g["hop"]))
    assert y.shape == g["noise_filtered"].shape
    assert rel_rms(y, T(g["noise_filtered"])) < 2e-6


def test_fir_design_unsupported_geometry_is_reported(G):
    from golf_b200 import GolfError

    assert not G.noise_fir_design_supported(65, 240) and not G.noise_fir_design_supported(256, 120)
    with pytest.raises(GolfError):
        G.noise_fir_design(torch.randn(1, 4800, device=DEV), torch.zeros(1, 21, 65, device=DEV), hann(128), 240)


def test_philox_normal_statistics_and_determinism(G):
    st = G.new_rng_state(DEV, seed=1234)
    x = G.philox_normal(8, 1 << 18, st)
    assert torch.equal(x, G.philox_normal(8, 1 << 18, st))  # counter-based: same state, same draw
    v = x.double()
    assert abs(float(v.mean())) < 3e-3 and abs(float(v.var()) - 1) < 5e-3
    assert abs(float((v**3).mean())) < 1e-2 and abs(float((v**4).mean()) - 3) < 3e-2
    for lag in (1, 2, 3, 4, 5, 240):
        assert abs(float((v[:, lag:] * v[:, :-lag]).mean())) < 3e-3, lag
    assert abs(float((v[0] * v[1]).mean())) < 5e-3  # utterances are independent streams
    assert float(v.abs().max()) > 4.2 and torch.isfinite(x).all()
    G.rng_advance(st)
    y = G.philox_normal(8, 1 << 18, st)
    assert int(st[1]) == 1 and abs(float((y.double() * v).mean())) < 3e-3  # a new, uncorrelated draw
    other = G.philox_normal(8, 4096, G.new_rng_state(DEV, seed=1235))
    assert not torch.equal(other, x[:, :4096])
    assert torch.equal(G.philox_normal(8, 4099, G.new_rng_state(DEV, seed=1234))[:, :4096], x[:, :4096])  # ragged tail


@pytest.mark.parametrize("Tn", [9600, 9601, 7003])
def test_in_kernel_noise_equals_filtering_the_same_draw(G, Tn):
    B, Fr, hop = 3, 41, 240
    g = torch.Generator().manual_seed(3)
    lm = (smooth(torch.randn(B, Fr, 256, generator=g)) - 4).to(DEV)
    add = torch.randn(B, Tn, generator=g).to(DEV)
    st = G.new_rng_state(DEV, seed=99)
    G.rng_advance(st)
    fused = G.noise_fir_design(None, lm, hann(510), hop, add=add, rng_state=st)
    draw = G.philox_normal(B, Tn, st)
    explicit = G.noise_fir_design(draw, lm, hann(510), hop, add=add)
    assert torch.equal(fused, explicit)


def _decoder_inputs(bench, B):
    s = {k: v[:B] for k, v in bench.make_inputs(1, max(B, 1))[0].items()}
    return s


@pytest.mark.parametrize("B", [1, 3])
def test_synth_fused_equals_module_path_and_oracle(G, oracle, B):
    """+N1: golf_synth_fused_fwd == oscillator -> FIR(+harm) -> lpc_ss -> room modules, and both match the oracle"""
    import bench
    from golf_b200 import noise as gnoise, sf
    from golf_b200.audiotensor import AudioTensor

    s = _decoder_inputs(bench, B)
    noise = torch.randn(B, bench.T, generator=torch.Generator().manual_seed(2))
    dec = bench.build_decoder(torch.device(DEV), "ss")
    dec.harm_oscillator.phase_accumulation = "aten_cpu"
    osc = dec.harm_oscillator
    d = {k: v.to(DEV) for k, v in s.items()}
    out = G.synth_fused(d["phase"], 1, d["w"], 2400, osc.table, osc.decimater.kernel, osc.oversampling, osc.equal_energy, "aten_cpu",
                        d["log_mag"], hann(510), d["gain"], d["a"], bench.HOP, dec.room_filter.kernel, noise.to(DEV))

    class Fixed(gnoise.NoiseInterface):
        def forward(self, ref_, *args):
            return AudioTensor(noise.to(DEV)[:, : ref_.shape[1]])

    dec.noise_generator = Fixed()
    A = lambda t, hop: AudioTensor(t.to(DEV), hop_length=hop)
    with torch.no_grad():
        mod = dec(phase=A(s["phase"], 1), harm_oscillator_params=(A(s["w"], 2400),), noise_generator_params=(),
                  noise_filter_params=(A(s["log_mag"], bench.HOP),), end_filter_params=(A(s["gain"], bench.HOP), A(s["a"], bench.HOP)))
    assert torch.equal(out, mod.as_tensor())
    ref = oracle.source_filter_synth(s["phase"], 1, s["w"], 2400, s["log_mag"], s["gain"], s["a"], bench.HOP, noise,
                                     osc.table.cpu(), bench.room_kernel(), variant="ss", oversampling=bench.OS)
    assert out.shape == ref.shape
    assert rel_rms(out, ref) < REL_TOL


def test_decoder_module_routes_through_fused_call(G):
    """SourceFilterSynth with the stock modules: FUSED on/off give the same samples for the same torch seed; the
    opt-in in-kernel generator equals a pass fed with the very draw it makes"""
    import bench
    from golf_b200 import sf
    from golf_b200.audiotensor import AudioTensor

    B = 2
    s = _decoder_inputs(bench, B)
    dec = bench.build_decoder(torch.device(DEV), "ss")
    A = lambda t, hop: AudioTensor(t.to(DEV), hop_length=hop)
    P = dict(phase=A(s["phase"], 1), harm_oscillator_params=(A(s["w"], 2400),), noise_generator_params=(),
             noise_filter_params=(A(s["log_mag"], bench.HOP),), end_filter_params=(A(s["gain"], bench.HOP), A(s["a"], bench.HOP)))
    outs = {}
    try:
        for fused in (True, False):
            sf.FUSED = fused
            torch.manual_seed(7)
            with torch.no_grad():
                outs[fused] = dec(**P).as_tensor()
    finally:
        sf.FUSED = True
    assert outs[True].shape == (B, bench.T - bench.HOP)
    assert torch.equal(outs[True], outs[False])
    # in-kernel generator
    dec.noise_generator.fused = True
    st = dec.noise_generator.rng_state(torch.device(DEV))
    before = st.clone()
    with torch.no_grad():
        y1 = dec(**P).as_tensor()
        y2 = dec(**P).as_tensor()
    assert int(st[1]) == int(before[1]) + 2 and not torch.equal(y1, y2) and torch.isfinite(y1).all()
    osc = dec.harm_oscillator
    d = {k: v.to(DEV) for k, v in s.items()}
    draw = G.philox_normal(B, bench.T, before)
    y_ref = G.synth_fused(d["phase"], 1, d["w"], 2400, osc.table, osc.decimater.kernel, osc.oversampling, osc.equal_energy,
                          osc.phase_accumulation, d["log_mag"], hann(510), d["gain"], d["a"], bench.HOP, dec.room_filter.kernel, draw)
    assert torch.equal(y1, y_ref)


def test_fused_decoder_under_cuda_graph_draws_fresh_noise(G):
    """the generator state lives in device memory and is advanced by the pass itself, so a replayed graph does not
    repeat its noise"""
    import bench
    from golf_b200.audiotensor import AudioTensor
    from golf_b200.graphs import GraphedSynth

    s = _decoder_inputs(bench, 2)
    dec = bench.build_decoder(torch.device(DEV), "ss")
    dec.noise_generator.fused = True
    A = lambda t, hop: AudioTensor(t.to(DEV), hop_length=hop)
    P = dict(phase=A(s["phase"], 1), harm_oscillator_params=(A(s["w"], 2400),), noise_generator_params=(),
             noise_filter_params=(A(s["log_mag"], bench.HOP),), end_filter_params=(A(s["gain"], bench.HOP), A(s["a"], bench.HOP)))
    from golf_b200 import _lib

    _lib.lib().golf_lpc_ss_set_tail(1)  # the five-launch schedule (cluster tail); the default keeps the light launches
    try:
        with torch.no_grad():
            gs = GraphedSynth(dec, P)
    finally:
        _lib.lib().golf_lpc_ss_set_tail(0)
    assert gs.kernels_captured <= 6
    a = gs.replay().as_tensor().clone()
    b = gs.replay().as_tensor().clone()
    assert not torch.equal(a, b) and torch.isfinite(a).all() and torch.isfinite(b).all()
