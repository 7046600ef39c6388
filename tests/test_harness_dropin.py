"""SURVEY 8f rank 4, second half: the reference's Lightning-side harness runs UNMODIFIED with a golf_b200 YAML.

Build container only (needs /root/reference).  `lightning` is absent from the image; oracle/lightning_standin.py
provides the few names `ltng/ae.py` and `test_rtf.py` import.  The test follows test_rtf.py:main step by step with the
reference's own functions: `dict2object` on the checkpoint's config.yaml with the decoder's class paths rewritten to
golf_b200, `VoiceAutoEncoder.load_from_checkpoint`, the `analysis` stage (the reference's encoder on
the CPU) and the `synthesis` call.  There is no GPU here and golf_b200 has no CPU path, so the synthesis call must raise
GolfError -- on a box that has both a GPU and the reference tree, `python tools/run_test_rtf.py ... --cuda` runs the same
script end to end."""
import copy
import os

import pytest
import torch
import yaml

pytestmark = pytest.mark.reference


def _rewrite(c):
    if isinstance(c, dict):
        return {k: (v.replace("models.", "golf_b200.", 1) if k == "class_path" else _rewrite(v)) for k, v in c.items()}
    return c


@pytest.mark.parametrize("name", ["golf-ss", "golf-ff"])
def test_test_rtf_flow_with_golf_b200_yaml(reference, name):
    from oracle import refimport

    ae, rtf = refimport.import_harness()
    from models.audiotensor import AudioTensor

    base = os.path.join(refimport.REF_ROOT, "ckpts", "interspeech24", name)
    cfg = yaml.safe_load(open(os.path.join(base, "config.yaml")))["model"]["init_args"]
    cfg = copy.deepcopy(cfg)
    cfg["decoder"] = _rewrite(cfg["decoder"])
    ck = os.path.join(base, "checkpoints", sorted(os.listdir(os.path.join(base, "checkpoints")))[0])
    model = ae.VoiceAutoEncoder.load_from_checkpoint(ck, map_location=torch.device("cpu"), **rtf.dict2object(cfg))
    assert not model.load_result.missing_keys and not model.load_result.unexpected_keys
    assert type(model.decoder).__module__ == "golf_b200.sf"
    assert type(model.decoder.end_filter).__module__ == "golf_b200.filters"
    model.eval()

    sr = 24000
    x = AudioTensor(0.05 * torch.randn(1, sr, generator=torch.Generator().manual_seed(0)))
    f0_hop = sr // 200
    f0_in_hz = AudioTensor(torch.full((1, x.shape[1] // f0_hop + 1), 150.0), f0_hop)

    def analysis():  # test_rtf.py:219-229
        params = model.encoder(x, f0=f0_in_hz if model.train_with_true_f0 else None)
        f0_hat = params.pop("f0", None)
        params["phase"] = (f0_hat if f0_hat is not None else f0_in_hz) / sr
        return params

    with torch.no_grad():
        measurements, params = rtf.bench(analysis, 3)
    assert len(measurements) == 1 and set(params) >= {"phase", "harm_oscillator_params", "noise_filter_params", "end_filter_params"}
    gain, a = params["end_filter_params"]
    assert a.shape[-1] == 22 and gain.hop_length == 240

    from golf_b200._lib import GolfError

    with torch.no_grad(), pytest.raises(GolfError, match="CUDA"):
        model.decoder(**params)  # test_rtf.py:236-238 -- no CPU path in golf_b200


def test_standin_runs_the_reference_decoder_end_to_end(reference):
    """the stand-in itself: the same flow with the reference's own decoder runs through synthesis on the CPU"""
    from oracle import refimport

    ae, rtf = refimport.import_harness()
    from models.audiotensor import AudioTensor

    base = os.path.join(refimport.REF_ROOT, "ckpts", "interspeech24", "golf-ss")
    cfg = copy.deepcopy(yaml.safe_load(open(os.path.join(base, "config.yaml")))["model"]["init_args"])
    ck = os.path.join(base, "checkpoints", sorted(os.listdir(os.path.join(base, "checkpoints")))[0])
    model = ae.VoiceAutoEncoder.load_from_checkpoint(ck, map_location=torch.device("cpu"), **rtf.dict2object(cfg)).eval()
    sr = 24000
    x = AudioTensor(0.05 * torch.randn(1, sr // 2, generator=torch.Generator().manual_seed(0)))
    f0 = AudioTensor(torch.full((1, x.shape[1] // 120 + 1), 150.0), 120)
    with torch.no_grad():
        params = model.encoder(x, f0=f0)
        params.pop("f0", None)
        params["phase"] = f0 / sr
        y = model.decoder(**params)
    assert y.shape[0] == 1 and y.shape[1] > sr // 2 - 600 and torch.isfinite(y.as_tensor()).all()
