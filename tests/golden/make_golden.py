"""Generate the committed golden fixtures by running the UNMODIFIED reference.

Run once in the build container (needs /root/reference; never on the GPU box):

    python tests/golden/make_golden.py

What it does
  1. imports the reference's models.* through oracle/refimport.py (stubs for the
     absent third-party modules are restatements -- see that file),
  2. builds the GOLF-ss / GOLF-ff decoders + encoder from
     ckpts/interspeech24/golf-{ss,ff}/config.yaml with the checkpoint weights,
  3. runs the real encoder on medias/samples/gt_{f1,m1}_{1,2,3}.wav (first 2 s) to get
     realistic, demanding control trajectories (pole radius up to ~0.996),
  4. runs the reference decoder stage by stage on a 1 s crop and stores every stage's
     input and output.

Files written (float32 npz, a few MB in total):
  controls_gt.npz   encoder-derived controls, 6 utterances x 2 s (inputs only; expected
                    outputs for these are computed by the oracle at test time -- the
                    oracle itself is pinned by the files below)
  stages_ss.npz     reference stage outputs, GOLF-ss, 2 utterances x 1 s
  stages_ff.npz     same for GOLF-ff (own checkpoint)
  filters_rand.npz  filter-only cases on seeded synthetic controls: ss, ff (centred and
                    not), biquad cascade (a14), inverse filter (a17), M in {8, 20, 22}
  grads_ss.npz      autograd of the reference ss/ff modules for a fixed upstream gradient
  table.npz         decoder.harm_oscillator.table rows [::9] + R_d_values + room kernels
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refimport  # noqa: E402

SR = 24000


def np32(t):
    return t.detach().cpu().to(torch.float32).numpy()


def build_model(name):
    from importlib import import_module

    from models.enc import VocoderParameterEncoderInterface

    base = os.path.join(refimport.REF_ROOT, "ckpts", "interspeech24", name)
    cfg = yaml.safe_load(open(os.path.join(base, "config.yaml")))["model"]["init_args"]

    def inst(c):
        if isinstance(c, dict):
            c = {k: inst(v) for k, v in c.items()}
            if "class_path" in c:
                mod, cls = c["class_path"].rsplit(".", 1)
                return getattr(import_module(mod), cls)(**c.get("init_args", {}))
        return c

    decoder = inst(cfg["decoder"])
    split_sizes, trsfms, args_keys = decoder.split_sizes_and_trsfms
    encoder = VocoderParameterEncoderInterface(
        split_sizes=split_sizes, trsfms=trsfms, args_keys=args_keys, **cfg["encoder_init_args"]
    )
    ck = os.path.join(base, "checkpoints")
    ck = os.path.join(ck, sorted(os.listdir(ck))[0])
    sd = torch.load(ck, map_location="cpu", weights_only=True)["state_dict"]
    print(name, decoder.load_state_dict({k[8:]: v for k, v in sd.items() if k.startswith("decoder.")}, strict=False))
    print(name, encoder.load_state_dict({k[8:]: v for k, v in sd.items() if k.startswith("encoder.")}, strict=False))
    return encoder.eval(), decoder.eval()


def load_wavs(seconds=2.0):
    from scipy.io import wavfile

    out = []
    for n in ("gt_f1_1", "gt_f1_2", "gt_f1_3", "gt_m1_1", "gt_m1_2", "gt_m1_3"):
        sr, x = wavfile.read(os.path.join(refimport.REF_ROOT, "medias", "samples", n + ".wav"))
        assert sr == SR
        x = x.astype(np.float32) if x.dtype.kind == "f" else x.astype(np.float32) / 32768.0
        start = SR  # skip the first second (often silence)
        out.append(x[start : start + int(seconds * SR)])
    return torch.tensor(np.stack(out))


def smooth(x, n=16):
    """moving average over n frames along dim 1, variance-preserving (SURVEY 8d recipe)."""
    xt = x.transpose(1, -1) if x.ndim > 2 else x
    flat = xt.reshape(-1, 1, xt.shape[-1])
    y = torch.nn.functional.conv1d(torch.nn.functional.pad(flat, (n - 1, 0), mode="replicate"), torch.ones(1, 1, n) / n)
    y = (y * n**0.5).view(xt.shape)
    return y.transpose(1, -1) if x.ndim > 2 else y


@torch.no_grad()
def main():
    refimport.import_reference()
    from models.audiotensor import AudioTensor
    from models.filters import LTVMinimumPhaseFilter, LTVMinimumPhaseFilterPrecise
    from models.lpc import BatchSecondOrderLPCSynth
    from models.utils import biquads2lpc, get_logits2biquads, rc2lpc

    torch.manual_seed(2434)
    x = load_wavs(2.0)
    B, T = x.shape
    f0_hop = SR // 200
    f0 = torch.full((B, T // f0_hop + 1), 150.0)

    # ---------------------------------------------------------------- controls + stages
    for name in ("golf-ss", "golf-ff"):
        enc, dec = build_model(name)
        params = enc(AudioTensor(x), f0=AudioTensor(f0, hop_length=f0_hop))
        (w,) = params["harm_oscillator_params"]
        (logmag,) = params["noise_filter_params"]
        gain, a = params["end_filter_params"]
        print(name, "w", tuple(w.shape), w.hop_length, "logmag", tuple(logmag.shape), logmag.hop_length,
              "gain", tuple(gain.shape), gain.hop_length, "a", tuple(a.shape), a.hop_length)
        if name == "golf-ss":
            np.savez(
                os.path.join(HERE, "controls_gt.npz"),
                w=np32(w), w_hop=w.hop_length, log_mag=np32(logmag), gain=np32(gain), a=np32(a),
                hop=gain.hop_length, f0_hz=150.0, f0_hop=f0_hop, sr=SR,
            )
            np.savez(
                os.path.join(HERE, "table.npz"),
                table_rows=np32(dec.harm_oscillator.table[::9]), row_index=np.arange(100)[::9],
                table_sum=np32(dec.harm_oscillator.table.double().sum(1)),
                R_d_values=np32(dec.harm_oscillator.R_d_values),
                room_kernel_ss=np32(dec.room_filter.kernel),
            )
        # stage-by-stage on a 1 s crop of two utterances (f1_1, m1_1)
        sel = [0, 3]
        Tc = SR
        nf = Tc // gain.hop_length + 1
        ph = AudioTensor(f0[sel, : Tc // f0_hop + 1] / SR, hop_length=f0_hop)
        w_c = AudioTensor(w.as_tensor()[sel, : Tc // w.hop_length + 1], hop_length=w.hop_length)
        lm_c = AudioTensor(logmag.as_tensor()[sel, :nf], hop_length=logmag.hop_length)
        g_c = AudioTensor(gain.as_tensor()[sel, :nf], hop_length=gain.hop_length)
        a_c = AudioTensor(a.as_tensor()[sel, :nf], hop_length=a.hop_length)
        harm = dec.harm_oscillator(ph, w_c)
        noise = torch.randn_like(harm.as_tensor())
        nz = dec.noise_filter(AudioTensor(noise), lm_c)
        src = harm + nz
        lpc = dec.end_filter(src, g_c, a_c)
        out = dec.room_filter(lpc)
        np.savez(
            os.path.join(HERE, f"stages_{name[-2:]}.npz"),
            phase=np32(ph.as_tensor()), phase_hop=f0_hop, w=np32(w_c.as_tensor()), w_hop=w.hop_length,
            log_mag=np32(lm_c.as_tensor()), gain=np32(g_c.as_tensor()), a=np32(a_c.as_tensor()), hop=gain.hop_length,
            noise=np32(noise), harm=np32(harm.as_tensor()),
            noise_filtered=np32(nz.as_tensor()), lpc=np32(lpc.as_tensor()),
            out=np32(out.as_tensor()), room_kernel=np32(dec.room_filter.kernel),
            window_length=960,
        )
        print(name, "stage lengths", harm.shape, nz.shape, src.shape, lpc.shape, out.shape)

    # ------------------------------------------------------------- filter-only, synthetic
    torch.manual_seed(2434)
    H, Tf = 240, 12000
    out = {}
    for M in (8, 20, 22):
        Fr = Tf // H + 1
        ex = torch.randn(2, Tf)
        gain = torch.exp(smooth(torch.randn(2, Fr)) - 6)
        if M == 8:
            bq = get_logits2biquads("coef", 0.99)(0.6 * smooth(torch.randn(2, Fr, 4, 2)))
            a = biquads2lpc(bq)
            out["biquads_8"] = np32(bq)
            lp = BatchSecondOrderLPCSynth(hop_length=H, window="hanning")
            out["bq_cascade_8"] = np32(lp(ex, gain[:, : Fr], bq))
        else:
            a = rc2lpc(torch.tanh(0.15 * smooth(torch.randn(2, Fr, M))))
        A = (AudioTensor(ex), AudioTensor(gain, hop_length=H), AudioTensor(a, hop_length=H))
        out[f"ex_{M}"], out[f"gain_{M}"], out[f"a_{M}"] = np32(ex), np32(gain), np32(a)
        out[f"ss_{M}"] = np32(LTVMinimumPhaseFilterPrecise(lpc_order=M)(*A).as_tensor())
        ff = LTVMinimumPhaseFilter(window="hanning", window_length=4 * H, lpc_order=M)
        out[f"ff_{M}"] = np32(ff(*A).as_tensor())
        ffn = LTVMinimumPhaseFilter(window="hanning", window_length=4 * H, centred=False, lpc_order=M)
        out[f"ffnc_{M}"] = np32(ffn(*A).as_tensor())
        tgt = torch.randn(2, Tf)
        _, resid = ff.reverse(A[0], AudioTensor(tgt), A[1], A[2])
        out[f"target_{M}"], out[f"inverse_{M}"] = np32(tgt), np32(resid.as_tensor())
    out["hop"] = H
    np.savez(os.path.join(HERE, "filters_rand.npz"), **out)

    # ---------------------------------------------------------------------- gradients
    with torch.enable_grad():
        torch.manual_seed(2434)
        M, Tg = 22, 4800
        Fr = Tg // H + 1
        ex = torch.randn(2, Tg, requires_grad=True)
        gain = torch.exp(smooth(torch.randn(2, Fr)) - 6).requires_grad_()
        a = rc2lpc(torch.tanh(0.15 * smooth(torch.randn(2, Fr, M)))).requires_grad_()
        g = {"ex": np32(ex), "gain": np32(gain), "a": np32(a), "hop": H}
        for tag, mod in (
            ("ss", LTVMinimumPhaseFilterPrecise(lpc_order=M)),
            ("ff", LTVMinimumPhaseFilter(window="hanning", window_length=4 * H, lpc_order=M)),
        ):
            y = mod(AudioTensor(ex), AudioTensor(gain, hop_length=H), AudioTensor(a, hop_length=H)).as_tensor()
            torch.manual_seed(7)
            up = torch.randn_like(y)
            dex, dgain, da = torch.autograd.grad(y, (ex, gain, a), up)
            g.update({f"{tag}_y": np32(y), f"{tag}_up": np32(up), f"{tag}_dex": np32(dex),
                      f"{tag}_dgain": np32(dgain), f"{tag}_da": np32(da)})
        np.savez(os.path.join(HERE, "grads_ss.npz"), **g)

    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
