#!/usr/bin/env python
"""Golden vectors for the rows VERDICT r1 marked partial (a3, a14, a16), produced by the UNMODIFIED
reference imported from /root/reference (build container only; oracle/refimport.py).

    python tests/golden/make_golden_ctrl_lpc.py   ->  tests/golden/ctrl_lpc.npz

* a3  DownsampledIndexedGlottalFlowTable.ctrl (models/synth.py:297-340) with the golf-ss checkpoint's
      downsampler weights: h [2,200,64] @240 -> w [2,21] @2400
* a16 get_logits2biquads / biquads2lpc (models/utils.py:444-525) for coef | conj | real, K = 11 sections,
      max_abs_pole 0.99 (the ISMIR-23 configuration), and the reference's autograd for a fixed upstream
* a14 models/lpc.py: LPCSynth, BatchLPCSynth, BatchSecondOrderLPCSynth forward (torchaudio lfilter inside) and
      autograd w.r.t. (ex, gain, a | biquads) for a fixed upstream
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refimport  # noqa: E402


def smooth(x, n=16):
    xt = x.transpose(1, -1) if x.ndim > 2 else x
    flat = xt.reshape(-1, 1, xt.shape[-1])
    y = torch.nn.functional.conv1d(torch.nn.functional.pad(flat, (n - 1, 0), mode="replicate"), torch.ones(1, 1, n) / n)
    y = (y * n**0.5).view(xt.shape)
    return y.transpose(1, -1) if x.ndim > 2 else y


def main():
    refimport.import_reference()
    from models.audiotensor import AudioTensor
    from models.lpc import BatchLPCSynth, BatchSecondOrderLPCSynth, LPCSynth
    from models.synth import DownsampledIndexedGlottalFlowTable
    from models.utils import biquads2lpc, get_logits2biquads, rc2lpc

    torch.manual_seed(2434)
    out = {}

    # ---- a3: the checkpoint's downsampler MLP
    base = os.path.join(refimport.REF_ROOT, "ckpts", "interspeech24", "golf-ss")
    cfg = yaml.safe_load(open(os.path.join(base, "config.yaml")))["model"]["init_args"]["decoder"]["init_args"]["harm_oscillator"]["init_args"]
    osc = DownsampledIndexedGlottalFlowTable(**cfg)
    ck = os.path.join(base, "checkpoints")
    sd = torch.load(os.path.join(ck, sorted(os.listdir(ck))[0]), map_location="cpu", weights_only=True)["state_dict"]
    pre = "decoder.harm_oscillator."
    print(osc.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}, strict=False))
    h = torch.randn(2, 200, 64)
    sizes, trsfms = osc.ctrl(lambda s, t: (s, t))((), ())
    with torch.no_grad():
        (w,) = trsfms[0](AudioTensor(h, hop_length=240))
    assert sizes == ((64,),) and w.hop_length == 2400
    out.update(ds_h=h.numpy(), ds_w=w.as_tensor().numpy(), ds_w_hop=w.hop_length)
    for k in ("model.1.weight", "model.1.bias", "model.3.weight", "model.3.bias"):
        out["ds_" + k.replace(".", "_")] = osc.state_dict()[k].numpy()

    # ---- a16: logits -> biquads -> polynomial, forward and autograd
    K = 11
    lg = 0.8 * torch.randn(2, 30, 2 * K)
    up = torch.randn(2, 30, 2 * K)
    out.update(bq_logits=lg.numpy(), bq_up=up.numpy(), bq_rho=0.99)
    for rep in ("coef", "conj", "real"):
        x = lg.clone().requires_grad_()
        a = biquads2lpc(get_logits2biquads(rep, 0.99)(x.view(2, 30, K, 2)))
        (d,) = torch.autograd.grad(a, x, up)
        out[f"bq_a_{rep}"], out[f"bq_dlogits_{rep}"] = a.detach().numpy(), d.numpy()

    # ---- a14: models/lpc.py modules
    H, M, B, Tn = 120, 8, 2, 2400
    Fr = (Tn + 2 * ((4 * H - H) // 2) - 4 * H) // H + 1
    a = rc2lpc(torch.tanh(0.2 * smooth(torch.randn(B, Fr, M))))
    gain = torch.exp(0.3 * smooth(torch.randn(B, Fr)) - 2)
    ex = torch.randn(B, Tn)
    Kb = 4
    bq = get_logits2biquads("coef", 0.9)(0.5 * smooth(torch.randn(B, Fr, Kb, 2).flatten(2)).view(B, Fr, Kb, 2))
    out.update(lpc_hop=H, lpc_ex=ex.numpy(), lpc_gain=gain.numpy(), lpc_a=a.numpy(), lpc_biquads=bq.numpy())
    exg, gg, ag, bg = (t.clone().requires_grad_() for t in (ex, gain, a, bq))
    y = BatchLPCSynth(H, window="hanning")(exg, gg, ag)
    upy = torch.randn_like(y)
    d = torch.autograd.grad(y, (exg, gg, ag), upy)
    out.update(lpc_batch_y=y.detach().numpy(), lpc_up=upy.numpy(), lpc_batch_dex=d[0].numpy(), lpc_batch_dgain=d[1].numpy(),
               lpc_batch_da=d[2].numpy())
    y1 = LPCSynth(H, window="hanning")(ex[0], torch.cat([gain[0, :, None], a[0]], -1))
    out["lpc_single_y"] = y1.detach().numpy()
    y2 = BatchSecondOrderLPCSynth(H, window="hanning")(exg, gg, bg)
    d2 = torch.autograd.grad(y2, (exg, gg, bg), upy)
    out.update(lpc_bq_y=y2.detach().numpy(), lpc_bq_dex=d2[0].numpy(), lpc_bq_dgain=d2[1].numpy(), lpc_bq_dbiquads=d2[2].numpy())

    path = os.path.join(HERE, "ctrl_lpc.npz")
    np.savez(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB", {k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
