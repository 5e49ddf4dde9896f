#!/usr/bin/env python
"""A few calls of uvic_b200_clinic on one synthetic grid (default: 0.5 degree x 40 levels), for ncu captures of the
k_clinic_* kernels:  python scripts/clinic_once.py [imt jmt km] ; prints the CUDA-event-free wall time per call."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg  # noqa: E402

pkg = load_pkg()
imt, jmt, km = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (722, 362, 40)
case = pkg.synthetic.make_case(imt=imt, jmt=jmt, km=km, nt=2, seed=2901)
pkg.synthetic.add_momentum(case)
ctx = pkg.TracerContext(case)
ctx.load_state()
ctx.upload_u_level(0, case["u"])
ctx.adv_vel()
ctx.clinic_setup(case, fourfil=os.environ.get("FOURFIL") == "1")
ctx.upload_u_level(-1, case["um1"])
ctx.upload_smf(np.stack([case["taux"], case["tauy"]]) * case["umask"][None, :, 0, :])
c2 = float(case.scalars["c2dtuv"])
for _ in range(3):
    ctx.clinic(c2)
ctx.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    ctx.clinic(c2)
ctx.synchronize()
print(f"clinic {imt}x{jmt}x{km}: {(time.perf_counter() - t0) / 10 * 1e3:.3f} ms per call (host wall clock, 10 calls)")
u = ctx.download_u(+1)
print("finite", bool(np.isfinite(u).all()), "max", float(np.abs(u).max()))
ctx.close()
