"""Shared helpers for the parity tests: run the oracle and the CUDA path on one Case."""
from __future__ import annotations

import numpy as np

from oracle_ffi import Oracle


def make_oracle(case, do_convect=1, timavgperts=0, do_mobi=0):
    o = Oracle(case.imt, case.jmt, case.km, case.nt, max(case.nsrc, 1))
    o.load_case(case)
    o.set_scalar("do_mobi", do_mobi)
    o.set_scalar("timavgperts", timavgperts)
    o.set_scalar("do_convect", do_convect)
    o.set_scalar("fct", 1)
    o.set_scalar("isopycmix", 1)
    o.set_scalar("tidal_kv", 1)
    return o


def oracle_rotate(o, leapfrog_next=True):
    """tau-1 <- tau <- tau+1 (source/mom/mom.F:210-212)."""
    t = o.t()
    t[0] = t[1]
    t[1] = t[2]


def oracle_set_step(o, case, leapfrog):
    dtts = case.scalars["dtts"]
    o.set_scalar("dtts", dtts)
    o.set_scalar("c2dtts", 2.0 * dtts if leapfrog else dtts)
    if not leapfrog:
        t = o.t()
        t[0] = t[1]          # both levels read from the tau slot (09/mom/loadmw.F:109-111)


def relerr(a, b):
    """max |a-b| / max |b| (per-field normalisation)."""
    den = np.abs(b).max()
    return np.abs(a - b).max() / (den if den > 0 else 1.0)


def interior(x):
    """drop the cyclic boundary columns"""
    return x[..., 1:-1]
