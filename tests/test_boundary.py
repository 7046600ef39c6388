"""CPU: the drop-in boundary -- C ABI exports, module surface, state-dict keys, error behaviour."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

from conftest import ROOT


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "golf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(golf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from golf_b200 import _lib

    assert os.path.exists(_lib.SO_PATH), "run `python -m golf_b200.build` (or __graft_entry__.build())"
    L = ctypes.CDLL(_lib.SO_PATH)
    declared = _header_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/golf_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == declared, "ctypes binding and header disagree"
    L.golf_abi_version.restype = ctypes.c_int
    assert L.golf_abi_version() == _lib.ABI_VERSION
    L.golf_strerror.restype = ctypes.c_char_p
    assert b"workspace" in L.golf_strerror(-3)


def test_library_is_sm100a_only():
    from golf_b200 import _lib

    out = subprocess.run(["cuobjdump", "--list-elf", _lib.SO_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_product_never_touches_the_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "golf_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(base, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "golf_oracle" in txt:
                    bad.append(f)
    assert not bad, f"product files reference the test oracle: {bad}"


def test_no_cpu_fallback():
    from golf_b200 import GolfError, functional as G

    x = torch.randn(1, 480)
    with pytest.raises(GolfError):
        G.lpc_ss(x, torch.ones(1, 3), torch.zeros(1, 3, 4), 240)
    with pytest.raises(GolfError):
        G.room_fir(x, torch.zeros(127))


def test_workspace_query_and_argument_checks():
    from golf_b200 import _lib

    L = _lib.lib()
    assert L.golf_lpc_ss_workspace_bytes(32, 47760, 22, 240, 0) > 0
    assert L.golf_lpc_ss_workspace_bytes(32, 47760, 99, 240, 0) == 0  # order outside the compiled range
    assert L.golf_lpc_ss_fwd(0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 0, 0, 0, 0) == -1  # null pointers -> GOLF_ERR_INVALID


def test_decoder_surface_matches_reference_layout():
    from golf_b200 import filters, noise, sf, synth

    dec = sf.SourceFilterSynth(
        synth.DownsampledIndexedGlottalFlowTable(hop_rate=10, in_channels=64, oversampling=4, equal_energy=True, lf_v2=True, points=2048),
        noise.StandardNormalNoise(),
        filters.LTVZeroPhaseFIRFilter("hanning", n_mag=256),
        filters.LTVMinimumPhaseFilter(window="hanning", window_length=960, lpc_order=22, lpc_parameterisation="rc2lpc"),
        filters.LTIAcousticFilter(128, "fft"),
        subtract_harmonics=False,
    )
    sizes, trsfms, names = dec.split_sizes_and_trsfms
    assert sizes == ((64,), (), (256,), (1, 22), ())  # 343 logits, SURVEY 8b
    assert names == ("harm_oscillator_params", "noise_generator_params", "noise_filter_params", "end_filter_params", "room_filter_params")
    # checkpoint keys of ckpts/interspeech24/golf-*/ (decoder.* prefix stripped)
    assert sorted(dec.state_dict()) == [
        "harm_oscillator.R_d_values", "harm_oscillator.model.1.bias", "harm_oscillator.model.1.weight",
        "harm_oscillator.model.3.bias", "harm_oscillator.model.3.weight", "harm_oscillator.table", "room_filter.kernel"]
    assert dec.harm_oscillator.table.shape == (100, 2048)
    assert dec.room_filter.kernel.shape == (127,)
    with pytest.raises(ValueError):
        filters.LTVMinimumPhaseFilterPrecise(lpc_order=4, lpc_parameterisation="nope")
    with pytest.raises(ValueError):
        synth.GlottalFlowTable(table_type="nope", lf_v2=True, points=64)


def test_ctrl_transforms_produce_filter_inputs():
    from golf_b200 import filters
    from golf_b200.audiotensor import AudioTensor

    f = filters.LTVMinimumPhaseFilterPrecise(lpc_order=6)
    (sizes,), (trsfm,) = f.ctrl(lambda s, t: (s, t))((), ())
    assert sizes == (1, 6)
    lg = AudioTensor(torch.randn(2, 5, generator=torch.Generator().manual_seed(0)), hop_length=240)
    logits = AudioTensor(torch.randn(2, 5, 6, generator=torch.Generator().manual_seed(1)), hop_length=240)
    gain, a = trsfm(lg, logits)
    assert gain.hop_length == 240 and a.hop_length == 240 and a.shape == (2, 5, 6)
    assert torch.all(gain > 0)


def test_table_matches_checkpoint_buffer():
    import numpy as np

    from conftest import golden
    from golf_b200 import synth

    osc = synth.IndexedGlottalFlowTable(table_size=100, lf_v2=True, points=2048, oversampling=4, equal_energy=True)
    tb = golden("table")
    assert np.abs(osc.table[::9].numpy() - tb["table_rows"]).max() < 2e-6
    assert np.abs(osc.R_d_values.numpy() - tb["R_d_values"]).max() == 0
    assert "decimater.kernel" not in osc.state_dict()  # non-persistent like the reference
    assert osc.decimater.kernel.shape == (129,)


def test_audiotensor_rate_semantics():
    """length rules documented by the reference's tests/test_time_tensor.py:18-28"""
    from golf_b200.audiotensor import AudioTensor

    a = AudioTensor(torch.randn(2, 100), hop_length=10)
    b = AudioTensor(torch.randn(2, 100), hop_length=5)
    assert a.reduce_hop_length().shape == (2, 991) and a.reduce_hop_length().hop_length == 1
    c = a + b
    assert c.shape == (2, 100) and c.hop_length == 5  # a upsampled to 199 steps, both cut to 100
    d = AudioTensor(torch.randn(2, 496), hop_length=1) * b
    assert d.shape == (2, 496) and d.hop_length == 1
    e = AudioTensor(torch.randn(2, 7, 3), hop_length=4) * AudioTensor(torch.randn(2, 40))
    assert e.shape == (2, 25, 3)
    assert a.unfold(20, 5).hop_length == 50 and a.increase_hop_length(2).shape == (2, 50)
    assert a.sum(1).hop_length == -1
    with pytest.raises(AssertionError):
        AudioTensor(torch.randn(5))


@pytest.mark.reference
def test_audiotensor_agrees_with_reference(reference):
    from golf_b200.audiotensor import _AudioTensor as Mine
    from models.audiotensor import AudioTensor as Ref

    g = torch.Generator().manual_seed(0)
    x, y, z = torch.randn(2, 21, generator=g), torch.randn(2, 4800, generator=g), torch.randn(2, 41, 3, generator=g)
    for op in (torch.mul, torch.add, torch.sub):
        r = op(Ref(x, hop_length=240), Ref(y))
        m = op(Mine(x, hop_length=240), Mine(y))
        assert torch.equal(r.as_tensor(), m.as_tensor()) and r.hop_length == m.hop_length
    r = Ref(z, hop_length=120) * Ref(x, hop_length=240)
    m = Mine(z, hop_length=120) * Mine(x, hop_length=240)
    assert torch.equal(r.as_tensor(), m.as_tensor()) and r.hop_length == m.hop_length == 120


@pytest.mark.reference
def test_interop_inside_the_reference_process(reference):
    """with the reference importable, golf_b200 modules ARE reference Controllables and can be
    mixed into the reference's own Synth"""
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from oracle import refimport; refimport.import_reference()\n"
        "import models.ctrl, models.sf, models.synth, models.noise, models.filters, models.audiotensor\n"
        "import golf_b200, golf_b200.filters as F, golf_b200.audiotensor as A, golf_b200.ctrl as C\n"
        "assert C.Controllable is models.ctrl.Controllable and A.AudioTensor is models.audiotensor.AudioTensor\n"
        "dec = models.sf.SourceFilterSynth(\n"
        "    models.synth.DownsampledIndexedGlottalFlowTable(hop_rate=10, in_channels=64, oversampling=4, equal_energy=True, lf_v2=True, points=2048),\n"
        "    models.noise.StandardNormalNoise(), F.LTVZeroPhaseFIRFilter('hanning', n_mag=256),\n"
        "    F.LTVMinimumPhaseFilterPrecise(lpc_order=22), F.LTIAcousticFilter(128, 'fft'), subtract_harmonics=False)\n"
        "sizes, _, names = dec.split_sizes_and_trsfms\n"
        "assert sizes == ((64,), (), (256,), (1, 22), ()), sizes\n"
        "print('ok')\n" % ROOT
    )
    env = dict(os.environ, GOLF_B200_INTEROP="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=ROOT)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def test_bench_reference_arm_prints_one_json_line():
    """bench.py contract, CPU arm: stdout carries exactly one JSON line with the driver's keys (library chatter
    goes to stderr), and it runs without a GPU"""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype",
              "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in d["config"]


@pytest.mark.parametrize("lf_v2", [True, False])
@pytest.mark.parametrize("table_type", ["flow", "derivative"])
@pytest.mark.parametrize("normalize_method", [None, "constant_power", "peak"])
@pytest.mark.parametrize("align_peak", [True, False])
def test_glottal_construction_options_match_the_reference(lf_v2, table_type, normalize_method, align_peak):
    """the sweep of the reference's tests/test_glottal.py:7-15, with values: every option builds the reference's table
    (golden from the unmodified GlottalFlowTable, tests/golden/make_golden_table_options.py)"""
    import numpy as np

    from conftest import golden
    from golf_b200 import synth

    g = golden("table_options")
    osc = synth.GlottalFlowTable(table_size=6, table_type=table_type, normalize_method=normalize_method, align_peak=align_peak,
                                 lf_v2=lf_v2, points=128)
    ref = g[f"{'v2' if lf_v2 else 'v1'}|{table_type}|{normalize_method}|{int(align_peak)}"]
    assert osc.table.shape == ref.shape
    assert np.abs(osc.R_d_values.numpy() - g["R_d_values"]).max() == 0
    assert np.abs(osc.table.numpy() - ref).max() <= 2e-6 * max(1.0, float(np.abs(ref).max()))


@pytest.mark.parametrize("hop,W,M,grad,ok", [
    (240, 960, 22, True, True),      # shipped GOLF-ff
    (120, 480, 22, True, True),      # ISMIR-23
    (256, 1024, 22, True, True),     # adjoint at the padded order 32
    (250, 1000, 22, False, True),    # inference only needs the forward constraints
    (250, 1000, 22, True, False),    # no compiled order >= 22 divides 250
    (32, 128, 22, False, False),     # hop < 40
    (240, 2400, 22, False, False),   # window / hop > 8
    (240, 1000, 22, False, False),   # window not a multiple of the hop
    (240, 960, 41, False, False),    # order > 40
])
def test_ff_geometry_validation_mirrors_the_kernels(hop, W, M, grad, ok):
    """filters._check_ff_geometry is what LTVMinimumPhaseFilter.forward consults: it must say in forward() what the C entry
    points (golf_lpc_ff_fwd / _bwd) would refuse later (csrc/lpc_ff.cu: fill_geometry, ff_adjoint_order)"""
    from golf_b200 import GolfError
    from golf_b200.filters import _check_ff_geometry

    if ok:
        _check_ff_geometry(hop, W, M, grad)
    else:
        with pytest.raises(GolfError, match="LTVMinimumPhaseFilter"):
            _check_ff_geometry(hop, W, M, grad)
