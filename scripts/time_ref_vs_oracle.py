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
for name, fn in (("oracle (port)", lambda: o.call("ora_step")), ("translated reference", lambda: T.ref_step(ref))):
    fn()
    n = 20
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    dt = (time.perf_counter() - t0) / n
    print(f"{name:22s} {dt * 1e3:7.2f} ms/step  {cells / dt / 1e6:6.2f} M cell.tracer/s")
