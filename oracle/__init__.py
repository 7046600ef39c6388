"""CPU oracle for the GOLF synthesis hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package.  Nothing under
``golf_b200/`` does (tests/test_boundary.py greps for it).

``oracle.golf_oracle``  -- restatement of the reference algorithms (C for the
sample recurrences, torch-CPU/numpy for the frame-rate and FIR stages), each
function citing the reference file:line it follows.
``oracle.refimport``    -- makes the *unmodified* reference importable from
``/root/reference`` in the build container (stubs for absent third-party
modules); used by ``tests/golden/make_golden.py`` and by the CPU tests that pin
the oracle against the real reference when that tree is present.
"""
