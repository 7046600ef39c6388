"""Drop-in check against the reference's shipped checkpoints (build container only: needs /root/reference).

A user of the reference switches to golf_b200 by rewriting `class_path: models.X` to `golf_b200.X` in the
checkpoint's own config.yaml.  For GOLF-ss / GOLF-ff / GOLF-v1 (Interspeech-24) and one ISMIR-23 checkpoint
(`coef` parameterisation, lf v1 table, hop 120 / window 480, centred: false) this test does exactly that:
the golf_b200 decoder must build from the unmodified init_args, `load_state_dict(strict=True)` the checkpoint's
`decoder.*` entries (minus the reference's non-persistent-in-spirit `_kernel` diag buffers of old checkpoints),
present the same control layout to the encoder as the reference decoder (split sizes, argument names), and apply
the same control transforms (checked on CPU: they are frame-rate torch code on both sides).
"""
import copy
import os
from importlib import import_module

import pytest
import torch
import yaml

from conftest import rel_rms

pytestmark = pytest.mark.reference

CKPTS = [
    ("interspeech24/golf-ss", None),
    ("interspeech24/golf-ff", None),
    ("interspeech24/golf-v1", None),
    ("ismir23/glottal_d_f1", "epoch=2669-step=792990_converted.ckpt"),
]


def _inst(c, rewrite):
    if isinstance(c, dict):
        c = {k: _inst(v, rewrite) for k, v in c.items()}
        if "class_path" in c:
            mod, cls = rewrite(c["class_path"]).rsplit(".", 1)
            return getattr(import_module(mod), cls)(**c.get("init_args", {}))
    return c


def _decoder_cfg(root, name):
    cfg = yaml.safe_load(open(os.path.join(root, "ckpts", name, "config.yaml")))["model"]
    cfg = cfg.get("init_args", cfg)
    return cfg["decoder"]


def _state(root, name, fname):
    base = os.path.join(root, "ckpts", name)
    if fname is None:
        base = os.path.join(base, "checkpoints")
        fname = sorted(os.listdir(base))[0]
    sd = torch.load(os.path.join(base, fname), map_location="cpu", weights_only=True)["state_dict"]
    return {k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")}


@pytest.mark.parametrize("name,fname", CKPTS)
def test_checkpoint_loads_into_golf_b200_decoder(reference, name, fname):
    from oracle import refimport

    cfg = _decoder_cfg(refimport.REF_ROOT, name)
    ours = _inst(copy.deepcopy(cfg), lambda p: p.replace("models.", "golf_b200.", 1))
    theirs = _inst(copy.deepcopy(cfg), lambda p: p)
    sd = _state(refimport.REF_ROOT, name, fname)
    assert sd, "no decoder.* entries in the checkpoint"
    res = ours.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    theirs.load_state_dict(sd, strict=True)
    assert type(ours).__module__.startswith("golf_b200.")
    # the encoder-facing contract: same splits, same argument names, in the same order
    s_o, t_o, k_o = ours.split_sizes_and_trsfms
    s_r, t_r, k_r = theirs.split_sizes_and_trsfms
    assert s_o == s_r and tuple(k_o) == tuple(k_r)
    # same persistent state (table, R_d_values, downsampler MLP, room kernel)
    sd_o, sd_r = ours.state_dict(), theirs.state_dict()
    assert set(sd_o) == set(sd_r)
    for k in sd_o:
        assert torch.equal(sd_o[k], sd_r[k]), k
    # same control transforms (frame-rate torch code on both sides; CPU)
    from golf_b200.audiotensor import AudioTensor as OurAT
    from models.audiotensor import AudioTensor as RefAT

    g = torch.Generator().manual_seed(0)
    hop = 240 if "interspeech" in name else 120
    for sizes, f_o, f_r in zip(s_o, t_o, t_r):
        xs = [0.5 * torch.randn(2, 40, n, generator=g) for n in sizes]
        with torch.no_grad():
            o = f_o(*[OurAT(x.squeeze(-1) if n == 1 else x, hop_length=hop) for x, n in zip(xs, sizes)])
            r = f_r(*[RefAT(x.squeeze(-1) if n == 1 else x, hop_length=hop) for x, n in zip(xs, sizes)])
        assert len(o) == len(r)
        pl = lambda t: t.as_subclass(torch.Tensor)  # (the two sides may use different AudioTensor classes)
        for a, b in zip(o, r):
            assert getattr(a, "hop_length", None) == getattr(b, "hop_length", None)
            assert rel_rms(pl(a).reshape(2, -1), pl(b).reshape(2, -1)) < 1e-6


def test_freshly_built_tables_match_checkpoint_buffers(reference):
    """a6: the table the constructor builds (lf_v2 for Interspeech-24, the iterative lf v1 fit for ISMIR-23) against
    the persistent `table` buffer of the checkpoints"""
    from oracle import refimport

    for name, fname in (CKPTS[0], CKPTS[3]):
        cfg = _decoder_cfg(refimport.REF_ROOT, name)
        ours = _inst(copy.deepcopy(cfg), lambda p: p.replace("models.", "golf_b200.", 1))
        sd = _state(refimport.REF_ROOT, name, fname)
        assert rel_rms(ours.harm_oscillator.table, sd["harm_oscillator.table"]) < 2e-6, name
        # (the ISMIR-23 buffer was written by an older torch: linspace/exp differ in the last ulp)
        assert torch.allclose(ours.harm_oscillator.R_d_values, sd["harm_oscillator.R_d_values"], rtol=3e-7, atol=0)
