"""GPU: the drop-in nn.Modules, end to end, against the reference's stage-by-stage golden."""
import pytest
import torch

from conftest import REL_TOL, T, golden, rel_rms

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def build_decoder(variant, g):
    from golf_b200 import filters, noise, sf, synth

    end = (filters.LTVMinimumPhaseFilterPrecise(lpc_order=22) if variant == "ss"
           else filters.LTVMinimumPhaseFilter(window="hanning", window_length=int(g["window_length"]), lpc_order=22))
    dec = sf.SourceFilterSynth(
        synth.DownsampledIndexedGlottalFlowTable(hop_rate=10, in_channels=64, oversampling=4, equal_energy=True,
                                                 table_type="derivative", normalize_method="constant_power", align_peak=True,
                                                 trainable=False, min_R_d=0.3, max_R_d=2.7, lf_v2=True, points=2048),
        noise.StandardNormalNoise(), filters.LTVZeroPhaseFIRFilter("hanning", n_mag=256), end,
        filters.LTIAcousticFilter(128, "fft"), subtract_harmonics=False)
    dec.room_filter.kernel.data = T(g["room_kernel"])
    return dec.to(DEV).eval()


def fixed_noise(noise):
    """stands in for StandardNormalNoise so both sides see the same draw"""
    from golf_b200.ctrl import Controllable

    class FixedNoise(Controllable):
        def forward(self, ref, *a):
            return type(ref)(noise[:, : ref.shape[1]], hop_length=1)

    return FixedNoise()


@pytest.mark.parametrize("variant", ["ss", "ff"])
def test_source_filter_synth_reference_golden(variant):
    from golf_b200.audiotensor import AudioTensor

    g = golden(f"stages_{variant}")
    dec = build_decoder(variant, g)
    dec.harm_oscillator.phase_accumulation = "aten_cpu"
    dec.noise_generator = fixed_noise(T(g["noise"]).to(DEV))
    A = lambda k, hop: AudioTensor(T(g[k]).to(DEV), hop_length=hop)
    H = int(g["hop"])
    with torch.no_grad():
        out = dec(phase=A("phase", int(g["phase_hop"])), harm_oscillator_params=(A("w", int(g["w_hop"])),),
                  noise_generator_params=(), noise_filter_params=(A("log_mag", H),),
                  end_filter_params=(A("gain", H), A("a", H)))
    assert out.hop_length == 1 and out.shape == g["out"].shape
    assert rel_rms(out, T(g["out"])) < REL_TOL


def test_modules_accept_precise_forward_and_backward():
    from golf_b200 import filters
    from golf_b200.audiotensor import AudioTensor

    g = golden("grads_ss")
    H = int(g["hop"])
    f = filters.LTVMinimumPhaseFilterPrecise(lpc_order=22)
    ex = T(g["ex"]).to(DEV).requires_grad_()
    gain = T(g["gain"]).to(DEV).requires_grad_()
    a = T(g["a"]).to(DEV).requires_grad_()
    y = f(AudioTensor(ex), AudioTensor(gain, hop_length=H), AudioTensor(a, hop_length=H))
    assert y.hop_length == 1
    (y.as_tensor() * T(g["ss_up"]).to(DEV)).sum().backward()
    assert rel_rms(ex.grad, T(g["ss_dex"])) < REL_TOL
    assert rel_rms(gain.grad, T(g["ss_dgain"])) < REL_TOL
    assert rel_rms(a.grad.flatten(1), T(g["ss_da"]).flatten(1)) < REL_TOL


def test_ff_module_not_centred_matches_golden():
    from golf_b200 import filters
    from golf_b200.audiotensor import AudioTensor

    g = golden("filters_rand")
    H = int(g["hop"])
    f = filters.LTVMinimumPhaseFilter(window="hanning", window_length=4 * H, centred=False, lpc_order=22).to(DEV)
    with torch.no_grad():
        y = f(AudioTensor(T(g["ex_22"]).to(DEV)), AudioTensor(T(g["gain_22"]).to(DEV), hop_length=H),
              AudioTensor(T(g["a_22"]).to(DEV), hop_length=H))
    assert y.shape == g["ffnc_22"].shape and rel_rms(y, T(g["ffnc_22"])) < 1e-5


def test_library_is_what_ran():
    """the .so must be mapped into this process and have launched kernels"""
    import golf_b200

    assert golf_b200.launch_count() > 0
    assert any("libgolf_b200.so" in line for line in open("/proc/self/maps"))


@pytest.mark.parametrize("variant", ["ss", "ff"])
def test_decoder_backward_directional_derivatives(variant):
    """config-4 decoder part: loss.backward() through the whole GOLF decoder (oscillator weight,
    noise FIR, LPC filter, room FIR) against central differences of the same forward along random
    directions.  y is linear in gain and in the room taps, so for a quadratic loss those two checks
    are exact up to rounding; the others carry an O(eps^2) truncation term."""
    from golf_b200.audiotensor import AudioTensor

    g = golden(f"stages_{variant}")
    dec = build_decoder(variant, g)
    dec.noise_generator = fixed_noise(T(g["noise"]).to(DEV))
    H = int(g["hop"])
    base = {k: T(g[k]).to(DEV) for k in ("phase", "w", "log_mag", "gain", "a")}
    base["room"] = dec.room_filter.kernel.data.clone()

    def loss_of(v):
        dec.room_filter.kernel.data = v["room"].detach() if not v["room"].requires_grad else dec.room_filter.kernel.data
        out = dec(phase=AudioTensor(v["phase"], hop_length=int(g["phase_hop"])),
                  harm_oscillator_params=(AudioTensor(v["w"], hop_length=int(g["w_hop"])),), noise_generator_params=(),
                  noise_filter_params=(AudioTensor(v["log_mag"], hop_length=H),),
                  end_filter_params=(AudioTensor(v["gain"], hop_length=H), AudioTensor(v["a"], hop_length=H)))
        return (out.as_tensor().double() ** 2).mean()

    leaves = {k: base[k].clone().requires_grad_() for k in ("w", "log_mag", "gain", "a")}
    dec.room_filter.kernel.data = base["room"].clone()
    dec.room_filter.kernel.requires_grad_(True)
    dec.zero_grad()
    loss_of({**base, **leaves}).backward()
    grads = {k: v.grad for k, v in leaves.items()}
    grads["room"] = dec.room_filter.kernel.grad.clone()
    dec.room_filter.kernel.requires_grad_(False)
    for k, gk in grads.items():
        assert gk is not None and torch.isfinite(gk).all(), k

    gen = torch.Generator().manual_seed(3)
    with torch.no_grad():
        # (`a` is left out: with pole radii ~0.996 a random perturbation large enough to beat float32
        #  rounding is far outside the linear regime; d_a is pinned against the reference's own
        #  autograd in test_gpu_parity.py::test_lpc_{ss,ff}_gradients_reference_golden)
        for k, eps, tol in (("gain", 2e-2, 2e-3), ("room", 2e-2, 2e-3), ("log_mag", 2e-2, 3e-2), ("w", 5e-3, 5e-2)):
            d = torch.randn(base[k].shape, generator=gen).to(DEV)
            if k == "gain":
                d = d * base[k]  # relative perturbation of a positive quantity
            plus, minus = dict(base), dict(base)
            plus[k], minus[k] = base[k] + eps * d, base[k] - eps * d
            fd = float((loss_of(plus) - loss_of(minus)) / (2 * eps))
            an = float((grads[k].double() * d.double()).sum())
            assert abs(fd - an) <= tol * max(abs(fd), abs(an)) + 1e-12, (k, fd, an)


def _ss_params(g, batch=None):
    from golf_b200.audiotensor import AudioTensor

    H = int(g["hop"])
    A = lambda k, hop: AudioTensor(T(g[k]).to(DEV), hop_length=hop)
    return dict(phase=A("phase", int(g["phase_hop"])), harm_oscillator_params=(A("w", int(g["w_hop"])),),
                noise_generator_params=(), noise_filter_params=(A("log_mag", H),), end_filter_params=(A("gain", H), A("a", H)))


def test_two_call_lpc_ss_matches_one_call():
    """responses on `a` alone + finish (z by a solve from rest) == the one-call filter up to float32
    rounding (the zero-state responses are summed in a different order)"""
    from golf_b200 import functional as G

    g = golden("grads_ss")
    H = int(g["hop"])
    ex, gain, a = (T(g[k]).to(DEV) for k in ("ex", "gain", "a"))
    one = G.lpc_ss(ex, gain, a, H)
    ws = G.lpc_ss_responses(a, G.lpc_ss_length(ex.shape[1], a.shape[1], H), H)
    two = G.lpc_ss_finish(ex, gain, a, H, ws)
    assert rel_rms(two, one) < 1e-5
    assert rel_rms(two, T(g["ss_y"])) < REL_TOL


def test_concurrent_decoder_path_same_arithmetic():
    """the multi-stream inference path only regroups the batch and reorders launches: with the noise
    branch silenced (log_mag -> -inf is not representable, so use a very low magnitude) the waveform of
    every utterance is bit-identical to the single-stream path, for every SPLIT"""
    from golf_b200 import sf

    g = golden("stages_ss")
    dec = build_decoder("ss", g)
    params = _ss_params(g)
    lm = params["noise_filter_params"][0]
    params["noise_filter_params"] = (type(lm)(torch.full_like(lm.as_tensor(), -200.0), hop_length=lm.hop_length),)  # exp -> 0
    outs = {}
    try:
        for mode, split in (("off", 1), ("auto", 1), ("auto", 2), ("auto", 4)):
            sf.CONCURRENT, sf.SPLIT = mode, split
            with torch.no_grad():
                outs[(mode, split)] = dec(**params).as_tensor().clone()
    finally:
        sf.CONCURRENT, sf.SPLIT = "auto", 1
    assert dec._can_run_concurrent(params["phase"], params["noise_filter_params"], params["end_filter_params"]) is False  # grad mode on
    base = outs[("off", 1)]
    assert base.shape == g["out"].shape and torch.isfinite(base).all() and float(base.abs().max()) > 0
    for k, v in outs.items():
        assert torch.equal(base, v), k


def test_programmatic_dependent_launch_same_results():
    """PDL only changes WHEN the kernels of the chain may start (each waits on-device for its
    predecessor): eager and CUDA-graph replays of the concurrent inference path give the same bits with
    the attribute on and off, for the same noise seed."""
    from golf_b200 import _lib
    from golf_b200.graphs import GraphedSynth

    g = golden("stages_ss")
    dec = build_decoder("ss", g)
    params = _ss_params(g)
    outs = {}
    try:
        for on in (0, 1):
            _lib.lib().golf_set_pdl(on)
            with torch.no_grad():
                for rep in range(3):  # back-to-back calls: the chain of one call follows the chain of the previous one
                    torch.manual_seed(7)
                    outs[(on, "eager", rep)] = dec(**params).as_tensor().clone()
                graphed = GraphedSynth(dec, params)
                for rep in range(3):
                    outs[(on, "graph", rep)] = graphed(**params).as_tensor().clone()
    finally:
        _lib.lib().golf_set_pdl(1)
    torch.cuda.synchronize()
    for rep in range(3):
        assert torch.equal(outs[(0, "eager", rep)], outs[(1, "eager", rep)]), rep
        assert torch.equal(outs[(0, "eager", 0)], outs[(1, "eager", rep)]), rep
    for on in (0, 1):  # graph replays draw fresh noise: same deterministic part -> same rms to a few percent
        for rep in range(3):
            o = outs[(on, "graph", rep)]
            assert torch.isfinite(o).all()
            assert abs(float(o.pow(2).mean().sqrt()) / float(outs[(0, "eager", 0)].pow(2).mean().sqrt()) - 1) < 0.05


def test_pipelined_synth_matches_graph_replay():
    """H2D / replay / D2H pipeline over 3 slots returns, per step, what a plain replay returns"""
    from golf_b200.graphs import GraphedSynth, PipelinedSynth

    g = golden("stages_ss")
    dec = build_decoder("ss", g)
    dev_params = _ss_params(g)
    host = []
    for s in range(5):  # five different control sets in pinned host memory
        p = _ss_params(g)
        gain = p["end_filter_params"][0]
        p["end_filter_params"] = (type(gain)((gain.as_tensor() * (1 + 0.1 * s)).cpu().pin_memory(), hop_length=gain.hop_length),
                                  type(gain)(p["end_filter_params"][1].as_tensor().cpu().pin_memory(), hop_length=gain.hop_length))
        p["phase"] = type(gain)(p["phase"].as_tensor().cpu().pin_memory(), hop_length=p["phase"].hop_length)
        host.append(p)
    with torch.no_grad():
        ref = GraphedSynth(dec, dev_params)
        pipe = PipelinedSynth(dec, dev_params, depth=3)
    outs = [torch.zeros(ref._out.shape[0], pipe.out_len).pin_memory() for _ in host]
    # the noise draw advances with every replay: compare on the deterministic part by fixing the seed per step
    # is not possible inside a graph, so check the linear dependence on gain instead: out is finite, differs
    # between steps, and a second pass over the same controls reproduces the slot bookkeeping (no torn reads)
    tickets = [pipe.submit(o, **p) for o, p in zip(outs, host)]
    for t in tickets:
        pipe.wait(t)
    torch.cuda.synchronize()
    for o in outs:
        assert torch.isfinite(o).all() and float(o.abs().max()) > 0
    r = [float(o.pow(2).mean().sqrt()) for o in outs]
    for s in range(1, 5):  # rms scales with the gain factor (noise differs per draw: 5 % slack)
        assert abs(r[s] / r[0] - (1 + 0.1 * s)) < 0.05 * (1 + 0.1 * s), r


def test_solver_variants_agree():
    """four-lanes-per-chunk (systolic) solve vs lane-per-chunk solve: same recurrence, different
    summation order -> equal to float32 rounding, and both within tolerance of the reference golden"""
    from golf_b200 import _lib, functional as G

    g = golden("grads_ss")
    H = int(g["hop"])
    ex, gain, a = (T(g[k]).to(DEV) for k in ("ex", "gain", "a"))
    outs = {}
    try:
        for mode in (0, 1):
            _lib.lib().golf_lpc_ss_set_solver(mode)
            outs[mode] = G.lpc_ss(ex, gain, a, H)
            ws = G.lpc_ss_responses(a, outs[mode].shape[1], H)
            outs[(mode, "two-call")] = G.lpc_ss_finish(ex, gain, a, H, ws)
    finally:
        _lib.lib().golf_lpc_ss_set_solver(1)
    ref = T(g["ss_y"])
    for k, v in outs.items():
        assert rel_rms(v, ref) < REL_TOL, k
        assert rel_rms(v, outs[0]) < 1e-5, k


def test_precise_zero_phase_fir_reference_golden():
    """drop-in LTVZeroPhaseFIRFilterPrecise vs the reference module (golden), forward and autograd"""
    from golf_b200 import filters
    from golf_b200.audiotensor import AudioTensor

    g = golden("fir_precise")
    H = int(g["hop"])
    f = filters.LTVZeroPhaseFIRFilterPrecise("hanning", n_mag=g["log_mag"].shape[-1]).to(DEV)
    lm = T(g["log_mag"]).to(DEV)
    for tag in ("long", "short"):
        with torch.no_grad():
            y = f(AudioTensor(T(g[f"ex_{tag}"]).to(DEV)), AudioTensor(lm, hop_length=H))
        assert y.hop_length == 1 and tuple(y.shape) == g[f"y_{tag}"].shape
        assert rel_rms(y.as_tensor(), T(g[f"y_{tag}"])) < 1e-5
    ex = T(g["ex_long"]).to(DEV).requires_grad_()
    lmg = lm.clone().requires_grad_()
    y = f(AudioTensor(ex), AudioTensor(lmg, hop_length=H)).as_tensor()
    d_ex, d_lm = torch.autograd.grad(y, (ex, lmg), T(g["up"]).to(DEV))
    assert rel_rms(d_ex, T(g["d_ex"])) < 1e-5 and rel_rms(d_lm.flatten(1), T(g["d_log_mag"]).flatten(1)) < 1e-4
    cfg = {"class_path": "golf_b200.filters.LTVZeroPhaseFIRFilter", "init_args": {"window": "hanning", "conv_method": "direct", "n_mag": 256}}
    assert filters.convert2samplewise({"noise_filter": cfg})["noise_filter"] == {
        "class_path": "golf_b200.filters.LTVZeroPhaseFIRFilterPrecise", "init_args": {"window": "hanning", "n_mag": 256}}


def test_replay_ring_concurrent_passes_do_not_interfere():
    """several decoder passes in flight on different streams (golf_b200.graphs.ReplayRing): every graph owns its
    workspaces, so concurrent replays must give exactly what one-at-a-time replays give (deterministic noise)"""
    from golf_b200.graphs import GraphedSynth, ReplayRing

    g = golden("stages_ss")
    dec = build_decoder("ss", g)
    dec.noise_generator = fixed_noise(T(g["noise"]).to(DEV))
    graphs = []
    with torch.no_grad():
        for s in range(4):
            p = _ss_params(g)
            gain = p["end_filter_params"][0]
            p["end_filter_params"] = (type(gain)(gain.as_tensor() * (1 + 0.25 * s), hop_length=gain.hop_length), p["end_filter_params"][1])
            graphs.append(GraphedSynth(dec, p))
        ref = [gr.replay().as_tensor().clone() for gr in graphs]
        torch.cuda.synchronize()
        for gr in graphs:
            gr._out.as_tensor().zero_()
        ring = ReplayRing(graphs, streams=4)
        ring.fork_from()
        for i in range(12):
            ring.submit(i)
        ring.join_into()
        torch.cuda.synchronize()
    for gr, r in zip(graphs, ref):
        assert torch.equal(gr._out.as_tensor(), r)
    assert not torch.equal(ref[0], ref[1])


def test_harmonic_plus_noise_synth_matches_oracle_composition(oracle):
    """HarmonicPlusNoiseSynth (GOLF-v1, cfg/ae/decoder/golf-v1.yaml; models/hpn.py:31-57): the harmonic branch goes
    through the frame-wise LPC filter alone, the FIR-filtered noise is added after it, then the room filter --
    checked against the same composition of the CPU restatements, on the reference's controls"""
    from golf_b200 import filters, hpn, synth
    from golf_b200.audiotensor import AudioTensor

    g = golden("stages_ff")
    H, W = int(g["hop"]), int(g["window_length"])
    dec = hpn.HarmonicPlusNoiseSynth(
        synth.DownsampledIndexedGlottalFlowTable(hop_rate=10, in_channels=64, oversampling=4, equal_energy=True, lf_v2=True, points=2048),
        fixed_noise(T(g["noise"]).to(DEV)), filters.LTVMinimumPhaseFilter(window="hanning", window_length=W, lpc_order=22),
        filters.LTVZeroPhaseFIRFilter("hanning", n_mag=256), filters.LTIAcousticFilter(128, "fft")).to(DEV).eval()
    dec.end_filter.kernel.data = T(g["room_kernel"]).to(DEV)
    dec.harm_oscillator.phase_accumulation = "aten_cpu"
    A = lambda k, hop: AudioTensor(T(g[k]).to(DEV), hop_length=hop)
    with torch.no_grad():
        out = dec(phase=A("phase", int(g["phase_hop"])), harm_oscillator_params=(A("w", int(g["w_hop"])),), noise_generator_params=(),
                  harm_filter_params=(A("gain", H), A("a", H)), noise_filter_params=(A("log_mag", H),))
    harm = T(g["harm"])  # the reference oscillator's own output (pinned by test_oscillator_reference_golden)
    hf = oracle.lpc_ff(harm, T(g["gain"]), T(g["a"]), H, W)
    nf = T(g["noise_filtered"])  # the reference noise branch (same module, same noise)
    n = min(hf.shape[1], nf.shape[1])
    ref = oracle.room_fir(hf[:, :n] + nf[:, :n], T(g["room_kernel"]))
    assert out.hop_length == 1 and tuple(out.shape) == tuple(ref.shape)
    assert rel_rms(out.as_tensor(), ref) < REL_TOL
    # voicing scales the phase (models/hpn.py:43-46)
    with torch.no_grad():
        out0 = dec(phase=A("phase", int(g["phase_hop"])), harm_oscillator_params=(A("w", int(g["w_hop"])),), noise_generator_params=(),
                   harm_filter_params=(A("gain", H), A("a", H)), noise_filter_params=(A("log_mag", H),),
                   voicing=AudioTensor(torch.full_like(T(g["phase"]), 0.5).to(DEV), hop_length=int(g["phase_hop"])))
    assert torch.isfinite(out0.as_tensor()).all() and not torch.equal(out0.as_tensor(), out.as_tensor())


def test_trainable_glottal_table_trains_through_the_decoder():
    """GlottalFlowTable(trainable=True) (models/synth.py:117-118): the table is an nn.Parameter of the decoder and receives
    its gradient from the CUDA adjoint; a small gradient step lowers a waveform loss by the first-order amount"""
    from golf_b200 import synth
    from golf_b200.audiotensor import AudioTensor

    osc = synth.IndexedGlottalFlowTable(oversampling=4, equal_energy=True, table_type="derivative",
                                        normalize_method="constant_power", align_peak=True, trainable=True, lf_v2=True,
                                        points=2048).to(DEV)
    assert isinstance(osc.table, torch.nn.Parameter) and "table" in dict(osc.named_parameters())
    gen = torch.Generator().manual_seed(3)
    phase = AudioTensor((torch.full((2, 100), 150.0) / 24000).to(DEV), hop_length=240)
    w = AudioTensor(torch.rand(2, 11, generator=gen).to(DEV), hop_length=2400)
    target = torch.randn(2, 24000, generator=gen).to(DEV)

    def loss_fn():
        y = osc(phase, w).as_tensor()
        return ((y - target[:, : y.shape[1]]) ** 2).mean()

    loss = loss_fn()
    loss.backward()
    g = osc.table.grad
    assert g is not None and g.shape == osc.table.shape and torch.isfinite(g).all() and g.abs().sum() > 0
    with torch.no_grad():  # a step sized for a 1 % first-order decrease; the loss is quadratic in the table
        osc.table -= 0.01 * loss / (g * g).sum() * g
        drop = float(loss - loss_fn()) / float(loss)
    assert 0.005 < drop < 0.0101


def test_ff_module_training_geometries():
    """window 1024 / hop 256 / order 22 trains (the adjoint runs at the padded order 32, which divides both); a geometry no
    compiled order divides fails in forward() with the reason when gradients are requested, and runs under no_grad"""
    from golf_b200 import GolfError, filters
    from golf_b200.audiotensor import AudioTensor

    gen = torch.Generator().manual_seed(5)

    def run(W, hop, grad):
        filt = filters.LTVMinimumPhaseFilter(window="hanning", window_length=W, lpc_order=22).to(DEV)
        Fr = 12
        ex = AudioTensor(torch.randn(2, (Fr - 1) * hop, generator=gen).to(DEV), hop_length=1)
        lg = AudioTensor(torch.randn(2, Fr, generator=gen).to(DEV) * 0.1, hop_length=hop)
        raw = (0.3 * torch.randn(2, Fr, 22, generator=gen)).to(DEV).requires_grad_(grad)
        logits = AudioTensor(raw, hop_length=hop)
        _, trsfms = filt.ctrl(lambda sizes, fns: (sizes, fns))((), ())  # the module's own .ctrl transform (exp, rc2lpc)
        gain, a = trsfms[0](lg, logits)
        y = filt(ex, gain, a).as_tensor()
        if grad:
            (g,) = torch.autograd.grad(y.square().mean(), raw)
            assert torch.isfinite(g).all() and g.abs().sum() > 0
        return y

    run(1024, 256, True)
    with torch.no_grad():
        assert torch.isfinite(run(1000, 250, False)).all()
    with pytest.raises(GolfError, match="divides both"):
        run(1000, 250, True)
