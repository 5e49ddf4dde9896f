#!/usr/bin/env python
"""Writes tests/golden/ref_step_t.npz: t(tau+1) of three steps (leapfrog, leapfrog, mixing) of the 37-tracer MOBI
configuration on the 20x16x6 grid, computed by THE REFERENCE'S OWN CODE (oracle/_ref/libref_t.so: the cpp-expanded Fortran
translated mechanically by oracle/refgen; isopyc -> vmixc -> tracer as source/mom/mom.F:340-389 sequences them).
Inputs are not stored: they are the seeded synthetic case `make_case(imt=20, jmt=16, km=6, nt=37, seed=29)`.

Run in the build container (needs /root/reference):  python tests/golden/make_ref_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import reflib  # noqa: E402
import test_cpu_refpin as T  # noqa: E402
from conftest import load_pkg  # noqa: E402

SEED = 29
SCHEDULE = (True, True, False)


def main():
    pkg = load_pkg()
    ref = reflib.RefLib("t")
    case, o = T.setup_pair(pkg, ref, seed=SEED)
    out = {}
    for itt, lf in enumerate(SCHEDULE):
        T.ref_set_step(ref, o, case, lf)
        ref.set("first", 1 if itt == 0 else 0)
        T.ref_step(ref)
        out[f"t_p1_step{itt}"] = ref.view("t")[2].copy()
        T.ref_rotate(ref)
    out["kmt"] = np.asarray(case["kmt"])
    out["schedule"] = np.array(SCHEDULE)
    out["seed"] = np.array(SEED)
    np.savez_compressed(os.path.join(HERE, "ref_step_t.npz"), **out)
    print("wrote ref_step_t.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
