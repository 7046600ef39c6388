"""A/B of the response pass on packed FP32 (variant build: GOLF_B200_VARIANT=f2 GOLF_NVCC_DEFS=-DGOLF_RESP_F2=1
python -m golf_b200.build) against the product library, at several refinement tolerances: runs bench.py in a child
process per configuration and prints value (8 passes in flight), one pass at a time, and the bench's parity self-check.

    python tools/gpu/ab_f2.py [tol ...]
"""
import json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CHILD = """
import sys
sys.path.insert(0, %r)
from golf_b200 import _lib
_lib.lib().golf_lpc_ss_set_refine_tolerance(%g)
import bench
sys.argv = ['bench.py', '--steps', '50', '--warmup', '5']
bench.main()
"""
tols = [float(t) for t in sys.argv[1:]] or [1e-4]
for so in ("", os.path.join(ROOT, "golf_b200", "_lib", "libgolf_b200_f2.so")):
    for tol in tols:
        env = dict(os.environ)
        if so:
            env["GOLF_B200_SO"] = so
        r = subprocess.run([sys.executable, "-c", CHILD % (ROOT, tol)], env=env, capture_output=True, text=True, cwd=ROOT)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            print(json.dumps({"lib": os.path.basename(so) or "product", "tol": tol, "value": d["value"], "ms_per_step": d["ms_per_step"],
                              "one_at_a_time_ms": d["ms_per_step_one_at_a_time"], "e2e": d["e2e"]["value"], "parity": d.get("parity")}))
        except Exception as e:  # noqa: BLE001
            print("FAILED", so, tol, e, r.stderr[-400:])
