"""Golden wavetables for every construction option the reference's own test sweeps (tests/test_glottal.py:7-15:
table_type x normalize_method x align_peak), for both LF generators (lf_v2 closed form, lf v1 iterative fit), produced by
the UNMODIFIED reference's GlottalFlowTable (models/synth.py:58-120).  Small tables (6 rows x 128 points) keep the file small.

    python tests/golden/make_golden_table_options.py          (build container only: needs /root/reference)
"""
import itertools
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refimport  # noqa: E402

models = refimport.import_reference()
out = {}
for lf_v2, table_type, norm, align in itertools.product((True, False), ("flow", "derivative"), (None, "constant_power", "peak"),
                                                         (True, False)):
    g = models.synth.GlottalFlowTable(table_size=6, table_type=table_type, normalize_method=norm, align_peak=align, lf_v2=lf_v2,
                                      points=128)
    out[f"{'v2' if lf_v2 else 'v1'}|{table_type}|{norm}|{int(align)}"] = g.table.detach().numpy().astype(np.float32)
    out["R_d_values"] = g.R_d_values.numpy().astype(np.float32)
np.savez_compressed(os.path.join(HERE, "table_options.npz"), **out)
print(len(out) - 1, "tables", out["v2|derivative|constant_power|1"].shape)
