"""CPU: the oracle's restatements composed into the GOLF-v1 and ISMIR-23 decoders (HarmonicPlusNoiseSynth, models/hpn.py:31-57)
against the outputs of the unmodified reference running the real checkpoints (tests/golden/decoder_{v1,ismir}.npz).  Pins the
frame-wise filter with `centred=False` at hop 120 / window 480, the biquad (`coef`) control transform composed with it, and the
harmonic-plus-noise wiring of the oracle -- the checker the GPU tests of these families rely on."""
import torch

from conftest import T, golden, rel_rms


def test_golf_v1_decoder_composition(oracle):
    g = golden("decoder_v1")
    H = int(g["hop"])
    table, _ = oracle.glottal_table()
    harm = oracle.glottal_osc(T(g["phase"]), H, T(g["harm_oscillator_params_0"]), int(g["harm_oscillator_params_0_hop"]), table, 4, True, "fp32")
    assert rel_rms(harm, T(g["harm"])) < 1e-4
    hf = oracle.lpc_ff(T(g["harm"]), T(g["harm_filter_params_0"]), T(g["harm_filter_params_1"]), H, 960)
    nf = oracle.noise_fir(T(g["noise"])[:, : g["harm"].shape[1]], T(g["noise_filter_params_0"]), H)
    n = min(hf.shape[1], nf.shape[1])
    out = oracle.room_fir(hf[:, :n] + nf[:, :n], T(g["sd_end_filter.kernel"]))
    assert out.shape == g["out"].shape and rel_rms(out, T(g["out"])) < 1e-5
    # the control transform of the harmonic filter: rc2lpc on tanh(logits)
    a = oracle.rc2lpc(torch.tanh(T(g["harm_filter_params_logits1"])))
    assert torch.allclose(a, T(g["harm_filter_params_1"]), atol=2e-6)


def test_ismir23_decoder_composition(oracle):
    g = golden("decoder_ismir")
    H = int(g["hop"])
    for br in ("harm", "noise"):  # `coef` parameterisation, 11 sections, rho 0.99 (ckpts/ismir23/glottal_d_f1/config.yaml:103-110)
        lg = T(g[f"{br}_filter_params_logits1"])
        a = oracle.biquads2lpc(oracle.logits2biquads(lg.view(*lg.shape[:2], 11, 2), "coef", 0.99))
        ref = T(g[f"{br}_filter_params_1"])
        assert float((a - ref).abs().max()) <= 3e-6 * float(ref.abs().max())
        assert torch.allclose(torch.exp(T(g[f"{br}_filter_params_logits0"]).squeeze(-1)), T(g[f"{br}_filter_params_0"]), rtol=1e-6)
    hf = oracle.lpc_ff(T(g["harm"]), T(g["harm_filter_params_0"]), T(g["harm_filter_params_1"]), H, 480, centred=False)
    nf = oracle.lpc_ff(T(g["noise"])[:, : g["harm"].shape[1]], T(g["noise_filter_params_0"]), T(g["noise_filter_params_1"]), H, 480, centred=False)
    n = min(hf.shape[1], nf.shape[1])
    out = hf[:, :n] + nf[:, :n]
    assert out.shape == g["out"].shape and rel_rms(out, T(g["out"])) < 1e-5
