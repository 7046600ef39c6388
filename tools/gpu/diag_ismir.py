import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import T, golden, rel_rms
from test_gpu_decoders_golden import _build
from golf_b200.audiotensor import AudioTensor
from golf_b200 import functional as G
from oracle import golf_oracle as O
O.build()
g = golden("decoder_ismir"); dev = "cuda:0"; H = 120
dec = _build("ismir")
sd = {k[3:]: T(g[k]) for k in g.files if k.startswith("sd_")}
dec.load_state_dict(sd, strict=False); dec = dec.to(dev).eval()
for mode in ("aten_cpu", "exact"):
    dec.harm_oscillator.phase_accumulation = mode
    with torch.no_grad():
        harm = dec.harm_oscillator(AudioTensor(T(g["phase"]).to(dev), hop_length=H), AudioTensor(T(g["harm_oscillator_params_0"]).to(dev), hop_length=1200))
    print(mode, "harm rel", rel_rms(harm.as_tensor(), T(g["harm"])), harm.shape)
for br, src in (("harm", T(g["harm"])), ("noise", T(g["noise"])[:, : g["harm"].shape[1]])):
    gain, a = T(g[f"{br}_filter_params_0"]), T(g[f"{br}_filter_params_1"])
    ref = O.lpc_ff(src, gain, a, H, 480, centred=False)
    ref64 = O.lpc_ff(src.double(), gain.double(), a.double(), H, 480, centred=False) if False else None
    with torch.no_grad():
        y = dec.harm_filter(AudioTensor(src.to(dev)), AudioTensor(gain.to(dev), hop_length=H), AudioTensor(a.to(dev), hop_length=H))
    print(br, "filter rel vs oracle", rel_rms(y.as_tensor(), ref), "max|ref|", float(ref.abs().max()), "rms", float(ref.square().mean().sqrt()))
    e = (y.as_tensor().cpu() - ref).abs()
    print("   worst sample", int(e.argmax()) % ref.shape[1], float(e.max()), " per-row rel", [(float(((y.as_tensor().cpu()[b]-ref[b])**2).mean().sqrt()/ (ref[b]**2).mean().sqrt())) for b in range(2)])
