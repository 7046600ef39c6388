"""GPU: block-wise (streaming) synthesis equals the one-shot GOLF-ss decoder on the same controls and noise."""
import pytest
import torch

from conftest import rel_rms, smooth, synthetic_controls

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SR, HOP, W_HOP, M = 24000, 240, 2400, 22


def _decoder():
    from golf_b200 import filters, noise, sf, synth

    dec = sf.SourceFilterSynth(
        synth.DownsampledIndexedGlottalFlowTable(hop_rate=10, in_channels=64, oversampling=4, equal_energy=True, lf_v2=True, points=2048),
        noise.StandardNormalNoise(), filters.LTVZeroPhaseFIRFilter("hanning", n_mag=256), filters.LTVMinimumPhaseFilterPrecise(lpc_order=M),
        filters.LTIAcousticFilter(128, "fft"), subtract_harmonics=False).to(DEV).eval()
    dec.room_filter.kernel.data = 0.05 * torch.randn(127, generator=torch.Generator().manual_seed(1)).to(DEV)
    return dec


def _controls(B, n_w, phase_hop, seed):
    g = torch.Generator().manual_seed(seed)
    T = n_w * W_HOP
    Fr = T // HOP
    gain, a = synthetic_controls(B, Fr, M, seed=seed)
    log_mag = smooth(torch.randn(B, Fr, 256, generator=g)) - 4
    w = torch.rand(B, n_w + 1, generator=g)
    f0 = (180 + 3 * torch.cumsum(torch.randn(B, T // phase_hop, generator=g), 1)).clamp(80, 400)
    return [t.to(DEV) for t in (f0 / SR, w, log_mag, gain, a)], T


@pytest.mark.parametrize("phase_hop,block_frames", [(1, 5), (1, 1), (1, 12), (120, 4)])
def test_streaming_equals_one_shot(phase_hop, block_frames):
    from golf_b200 import streaming
    from golf_b200.audiotensor import AudioTensor
    from golf_b200.ctrl import Controllable

    B, n_w = 2, 12
    (phase, w, log_mag, gain, a), T = _controls(B, n_w, phase_hop, seed=phase_hop + block_frames)
    noise = torch.randn(B, T, generator=torch.Generator().manual_seed(9)).to(DEV)
    dec = _decoder()

    class FixedNoise(Controllable):
        def forward(self, ref, *args):
            n = ref.shape[1]
            return type(ref)(torch.nn.functional.pad(noise, (0, max(n - T, 0)))[:, :n], hop_length=1)

    generator = dec.noise_generator
    dec.noise_generator = FixedNoise()
    with torch.no_grad():
        ref = dec(phase=AudioTensor(phase, hop_length=phase_hop), harm_oscillator_params=(AudioTensor(w, hop_length=W_HOP),),
                  noise_generator_params=(), noise_filter_params=(AudioTensor(log_mag, hop_length=HOP),),
                  end_filter_params=(AudioTensor(gain, hop_length=HOP), AudioTensor(a, hop_length=HOP))).as_tensor()
    dec.noise_generator = generator
    pos = [0]

    def noise_fn(b, n, dev):
        out = noise[:, pos[0] : pos[0] + n]
        pos[0] += n
        return out

    y = streaming.synthesize_long(dec, phase, w, log_mag, gain, a, block_frames=block_frames, hop=HOP, w_hop=W_HOP, phase_hop=phase_hop,
                                  noise_fn=noise_fn)
    assert y.shape == (B, T) and torch.isfinite(y).all()
    # the one-shot decoder stops a hop early (and, with frame-rate f0, its source ends 119 samples early, after
    # which it sees zero padding where the stream still has noise): compare where both have full context
    n_cmp = ref.shape[1] - (0 if phase_hop == 1 else 2 * HOP)
    assert n_cmp > T - 4 * HOP
    err = rel_rms(y[:, :n_cmp], ref[:, :n_cmp])
    assert err < 1e-5, err


def test_oscillator_initial_phase_continues_a_stream():
    """phase_offset (models/synth.py:251-252) as the kernel's initial phase: the second half of a signal,
    started from the running phase the first half ended on, equals the second half of the whole"""
    from golf_b200 import functional as G
    from oracle import golf_oracle as O

    g = torch.Generator().manual_seed(2)
    B, T = 2, 4 * W_HOP
    ph = ((150 + 100 * torch.rand(B, T + 1, generator=g)) / SR).to(DEV)
    w = torch.rand(B, 5, generator=g).to(DEV)
    table, _ = O.glottal_table()
    dk = O.decimate_kernel(4).to(DEV)
    whole = G.glottal_osc(ph, 1, w, W_HOP, table.to(DEV), dk, 4, True, "exact")
    half = 2 * W_HOP
    # running phase before sample `half`: closed form of the knot intervals, float64 (4x oversampled: hp = 4)
    x = ph[:, : half + 1].double() / 4
    cum = (4 * x[:, :-1] + (x[:, 1:] - x[:, :-1]) * 1.5).sum(1)
    second = G.glottal_osc(ph[:, half:].contiguous(), 1, w[:, 2:].contiguous(), W_HOP, table.to(DEV), dk, 4, True, "exact", phase0=cum)
    # away from the cut (the decimator sees nothing before the second call's first sample)
    assert rel_rms(second[:, 64:], whole[:, half + 64 :]) < 1e-5
