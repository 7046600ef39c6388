"""GPU: the two shipped decoder families beside GOLF-ss / GOLF-ff, whole decoder, against outputs of the UNMODIFIED reference
running the real checkpoints on the CPU (tests/golden/make_golden_decoders.py):

  GOLF-v1 (ckpts/interspeech24/golf-v1)   HarmonicPlusNoiseSynth: harmonic branch through LTVMinimumPhaseFilter (rc2lpc), noise
                                          through the zero-phase FIR, room FIR behind both           (a2, a4, a8, a11, a18)
  ISMIR-23 (ckpts/ismir23/glottal_d_f1)   HarmonicPlusNoiseSynth with both branches through LTVMinimumPhaseFilter(coef, 0.99,
                                          window 480, hop 120, centred=False), lf v1 table, no oversampling   (a2, a6, a11, a16)

Inputs are raw encoder logits: the test pushes them through the golf_b200 decoder's own .ctrl transforms (downsampler MLP with
the checkpoint's weights, rc2lpc / biquad kernels) exactly as VocoderParameterEncoderInterface does, so the control side is
part of what is compared.  The modules are constructed from the init_args of the checkpoints' config.yaml, written out here
because the GPU box has no reference tree."""
import pytest
import torch

from conftest import REL_TOL, T, golden, rel_rms

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _fixed_noise(noise):
    from golf_b200 import noise as gnoise
    from golf_b200.audiotensor import AudioTensor

    class FixedNoise(gnoise.NoiseInterface):
        def forward(self, ref, *a):
            return AudioTensor(noise[:, : ref.shape[1]])

    return FixedNoise()


def _build(which):
    from golf_b200 import ctrl, filters, hpn, synth

    if which == "v1":  # ckpts/interspeech24/golf-v1/config.yaml
        osc = synth.DownsampledIndexedGlottalFlowTable(hop_rate=10, in_channels=64, oversampling=4, equal_energy=True, table_size=100,
                                                       table_type="derivative", normalize_method="constant_power", align_peak=True,
                                                       trainable=False, min_R_d=0.3, max_R_d=2.7, lf_v2=True, points=2048)
        harm_f = filters.LTVMinimumPhaseFilter(window="hanning", window_length=960, lpc_order=22, lpc_parameterisation="rc2lpc", max_abs_value=1.0)
        noise_f = filters.LTVZeroPhaseFIRFilter(window="hanning", conv_method="direct", n_mag=256)
        end = filters.LTIAcousticFilter(length=128, conv_method="fft")
    else:  # ckpts/ismir23/glottal_d_f1/config.yaml
        osc = synth.DownsampledIndexedGlottalFlowTable(hop_rate=10, in_channels=64, table_size=100, table_type="derivative",
                                                       normalize_method="constant_power", align_peak=True, trainable=False, min_R_d=0.3,
                                                       max_R_d=2.7, T_0=5.0, n_iter_eps=5, n_iter_a=100, points=2048)
        mk = lambda: filters.LTVMinimumPhaseFilter(window="hanning", window_length=480, centred=False, lpc_order=22,
                                                   lpc_parameterisation="coef", max_abs_value=0.99)
        harm_f, noise_f, end = mk(), mk(), ctrl.PassThrough()
    return hpn.HarmonicPlusNoiseSynth(osc, _fixed_noise(None), harm_f, noise_f, end)


@pytest.mark.parametrize("which", ["v1", "ismir"])
def test_decoder_family_reference_golden(which):
    from golf_b200.audiotensor import AudioTensor

    g = golden(f"decoder_{which}")
    hop = int(g["hop"])
    dec = _build(which)
    sd = {k[3:]: T(g[k]) for k in g.files if k.startswith("sd_")}
    missing = dec.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys and all(k.endswith("table") or "_kernel" in k or "_window" in k for k in missing.missing_keys), missing
    dec = dec.to(DEV).eval()
    dec.noise_generator = _fixed_noise(T(g["noise"]).to(DEV))
    dec.harm_oscillator.phase_accumulation = "aten_cpu"
    # encoder logits -> the decoder's own control transforms (models/enc.py:73-98)
    sizes, trsfms, keys = dec.split_sizes_and_trsfms
    assert list(keys) == [str(k) for k in g["keys"]] and [len(s) for s in sizes] == [int(n) for n in g["sizes"]]
    params = {}
    with torch.no_grad():
        for key, grp, fn in zip(keys, sizes, trsfms):
            args = []
            for i, n in enumerate(grp):
                lg = T(g[f"{key}_logits{i}"]).to(DEV)
                args.append(AudioTensor(lg.squeeze(2) if n == 1 else lg, hop_length=hop))
            vals = fn(*args)
            for i, v in enumerate(vals):  # the transformed controls themselves (a3, a15, a16 through the modules)
                ref = T(g[f"{key}_{i}"])
                assert v.hop_length == int(g[f"{key}_{i}_hop"]) and tuple(v.shape) == tuple(ref.shape)
                assert rel_rms(v.as_tensor().reshape(ref.shape[0], -1), ref.reshape(ref.shape[0], -1)) < 1e-5, (key, i)
            params[key] = vals
        out = dec(phase=AudioTensor(T(g["phase"]).to(DEV), hop_length=hop), **params)
    assert out.hop_length == 1 and tuple(out.shape) == tuple(g["out"].shape)
    assert rel_rms(out.as_tensor(), T(g["out"])) < REL_TOL
    if which == "ismir":  # and with the filters in libtorchaudio's summation order
        from golf_b200 import _lib

        L = _lib.lib()
        L.golf_lpc_ff_set_exact_order(1)
        try:
            with torch.no_grad():
                out_exact = dec(phase=AudioTensor(T(g["phase"]).to(DEV), hop_length=hop), **params)
        finally:
            L.golf_lpc_ff_set_exact_order(0)
        assert rel_rms(out_exact.as_tensor(), T(g["out"])) < REL_TOL


def test_ff_exact_order_is_bitwise_torchaudio_per_frame(oracle):
    """golf_lpc_ff_set_exact_order(1): every frame's recurrence is libtorchaudio's CPU loop bit for bit (through
    lpc_synthesis-shaped frames: window length == hop count of one frame is not expressible, so the check goes through
    golf_lpc_frames_fwd with a rectangular window and non-overlapping frames, where the OLA is the identity)"""
    from golf_b200 import _lib, functional as G

    gen = torch.Generator().manual_seed(11)
    B, hop, M = 3, 240, 22
    Fr = 12
    ex = torch.randn(B, Fr * hop, generator=gen)
    gain = torch.rand(B, Fr, generator=gen) + 0.5
    a = oracle.rc2lpc(torch.tanh(0.4 * torch.randn(B, Fr, M, generator=gen)))
    # frames of 2*hop every hop, rectangular window: out = (y_k second half + y_{k+1} first half) / 2 in the interior;
    # compare the kernel in both orders with the same composition of the bit-exact oracle loop
    win = torch.ones(2 * hop)
    pad = hop // 2
    fr = torch.nn.functional.pad(ex, (pad, pad)).unfold(1, 2 * hop, hop) * gain[:, : Fr, None][:, : (ex.shape[1] + 2 * pad - 2 * hop) // hop + 1]
    nfr = fr.shape[1]
    yk = oracle.allpole_lti(fr.reshape(B * nfr, 2 * hop).contiguous(), a[:, :nfr].reshape(B * nfr, M)).view(B, nfr, 2 * hop)
    full = torch.zeros(B, (nfr - 1) * hop + 2 * hop)
    norm = torch.zeros((nfr - 1) * hop + 2 * hop)
    for k in range(nfr):
        full[:, k * hop : k * hop + 2 * hop] += yk[:, k]
        norm[k * hop : k * hop + 2 * hop] += 1
    ref = (full / norm)[:, pad : pad + (nfr - 1) * hop + 2 * hop - 2 * pad]
    L = _lib.lib()
    L.golf_lpc_ff_set_exact_order(1)
    try:
        y_exact = G.lpc_frames(ex.to(DEV), gain.to(DEV), a.to(DEV), win.to(DEV), hop)
    finally:
        L.golf_lpc_ff_set_exact_order(0)
    y_fast = G.lpc_frames(ex.to(DEV), gain.to(DEV), a.to(DEV), win.to(DEV), hop)
    assert y_exact.shape == ref.shape
    # single-frame regions (the first and last half hop... none here) aside, sums of two bit-identical frames divided by 2
    assert torch.equal(y_exact.cpu(), ref), float((y_exact.cpu() - ref).abs().max())
    assert rel_rms(y_fast, ref) < 1e-5 and not torch.equal(y_fast.cpu(), ref)
