#!/usr/bin/env python
"""How fast is the hand-written oracle (bench.py's CPU arm, kind "port") next to the reference's own code?

Times one `tracer` step (isopyc -> vmixc -> tracer with MOBI, all 37 tracers) of the mechanically translated reference
(oracle/_ref/libref_s.so, see oracle/refgen) and of the oracle on the same 34x26x8 case, same compiler flags, one core.
Test infrastructure only: needs oracle/_ref (built where /root/reference exists)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import reflib  # noqa: E402
import test_cpu_refpin as T  # noqa: E402
from conftest import load_pkg  # noqa: E402
from helpers import oracle_set_step  # noqa: E402

pkg = load_pkg()
ref = reflib.RefLib("s")
case, o = T.setup_pair(pkg, ref, seed=3)
oracle_set_step(o, case, True)
T.ref_set_step(ref, o, case, True)
ref.set("first", 1)
cells = (case.imt - 2) * (case.jmt - 2) * case.km * case.nt
out = {"grid": [case.imt, case.jmt, case.km], "nt": case.nt, "flags": "gcc -O2 -ffp-contract=off (both)", "steps": 20}
for key, name, fn in (("port_ms", "oracle (port)", lambda: o.call("ora_step")), ("reference_ms", "translated reference", lambda: T.ref_step(ref))):
    fn()
    best = 1e30
    for _ in range(3):                       # best of three blocks of 20 steps
        t0 = time.perf_counter()
        for _ in range(out["steps"]):
            fn()
        best = min(best, (time.perf_counter() - t0) / out["steps"])
    out[key] = round(best * 1e3, 3)
    if "--json" not in sys.argv:
        print(f"{name:22s} {best * 1e3:7.2f} ms/step  {cells / best / 1e6:6.2f} M cell.tracer/s")
if "--json" in sys.argv:
    import json

    print(json.dumps(out))
