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


def oracle_load_momentum(o, case):
    """wind stress into the coupler slots the oracle's setvbc reads (09/mom/setvbc.F:163-164); slots 1 and 2"""
    imt, jmt = case.imt, case.jmt
    sbc = o.arr("sbc").reshape(-1, jmt, imt)
    sbc[0] = case["taux"]
    sbc[1] = case["tauy"]
    o.set_scalar("itaux", 1)
    o.set_scalar("itauy", 2)


def oracle_clinic(o):
    """adv_vel -> state -> setvbc -> clinic as mom sequences them (source/mom/mom.F:300-390)"""
    for fn in ("ora_adv_vel", "ora_adv_vel_u", "ora_state", "ora_setvbc_mom", "ora_clinic"):
        o.call(fn)


def pointwise_relerr(a, b, rel_floor=1e-3):
    """max over cells of |a-b| / |b|, taken over the cells with |b| > rel_floor * max|b| (a point-wise relative
    difference has no meaning where the field passes through zero: sums of opposite-signed flux terms).
    Reported beside `relerr`, the field-normalised metric."""
    den = np.abs(b)
    big = den > rel_floor * (den.max() if den.size else 0.0)
    if not big.any():
        return 0.0
    return float((np.abs(a - b)[big] / den[big]).max())


def make_oracle_o3(case, **kw):
    """An oracle context from the -O3 build of the same sources (FMA contraction allowed: the reference's own optimisation
    level, run/mk.ver).  Run beside the strict build it measures how far two legitimate builds of the reference's arithmetic
    drift apart on a given state -- the conditioning of the step, used to scale gates where 1e-12 is not attainable by ANY
    pair of builds (calcite dissolution near saturation, see test_calcite_is_ill_conditioned_on_deep_grids)."""
    import os
    import oracle_ffi

    old_lib, old_path = oracle_ffi._lib, oracle_ffi.LIB
    oracle_ffi.build_oracle()
    oracle_ffi._lib, oracle_ffi.LIB = None, os.path.join(oracle_ffi.ORACLE_DIR, "_build", "liboracle_o3.so")
    try:
        o = make_oracle(case, **kw)
    finally:
        oracle_ffi._lib, oracle_ffi.LIB = old_lib, old_path
    return o
