"""Control-side rows against vectors the UNMODIFIED reference produced (tests/golden/make_golden_ctrl_lpc.py):

* a3  DownsampledIndexedGlottalFlowTable.ctrl with the GOLF-ss checkpoint's downsampler weights (models/synth.py:297-340)
* a16 get_logits2biquads -> biquads2lpc for coef | conj | real at the ISMIR-23 configuration (models/utils.py:444-525)

Frame-rate torch code, so these run on the CPU here and on the GPU box alike (the CUDA twins of a16 are checked in
tests/test_gpu_lpc_modules.py against the same vectors)."""
import pytest
import torch

from conftest import T, golden, rel_rms


def test_downsampler_ctrl_matches_checkpoint_mlp():
    from golf_b200.audiotensor import AudioTensor
    from golf_b200.synth import DownsampledIndexedGlottalFlowTable

    g = golden("ctrl_lpc")
    osc = DownsampledIndexedGlottalFlowTable(hop_rate=10, in_channels=64, oversampling=4, equal_energy=True, lf_v2=True, points=2048)
    sd = {k: T(g["ds_" + k.replace(".", "_")]) for k in ("model.1.weight", "model.1.bias", "model.3.weight", "model.3.bias")}
    res = osc.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys and set(res.missing_keys) <= {"table", "R_d_values"}
    sizes, trsfms = osc.ctrl(lambda s, t: (s, t))((), ())
    assert sizes == ((64,),)
    with torch.no_grad():
        (w,) = trsfms[0](AudioTensor(T(g["ds_h"]), hop_length=240))
    assert w.hop_length == int(g["ds_w_hop"]) == 2400
    assert w.shape == g["ds_w"].shape
    assert rel_rms(w.as_tensor(), T(g["ds_w"])) < 1e-6


@pytest.mark.parametrize("rep", ["coef", "conj", "real"])
def test_logits2biquads2lpc_matches_reference(rep):
    from golf_b200 import utils as U

    g = golden("ctrl_lpc")
    lg = T(g["bq_logits"]).requires_grad_()
    a = U.biquads2lpc(U.get_logits2biquads(rep, float(g["bq_rho"]))(lg.view(2, 30, -1, 2)))
    assert rel_rms(a, T(g[f"bq_a_{rep}"])) < 1e-6
    (d,) = torch.autograd.grad(a, lg, T(g["bq_up"]))
    assert rel_rms(d, T(g[f"bq_dlogits_{rep}"])) < 1e-5
