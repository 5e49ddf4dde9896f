"""GPU parity where the benchmarks run: oracle comparisons with km > 20 (the k-tiled marching FCT
k_fct_march<16>, k_invtri<8>, k_mobi_column) and nt = 40 (37 MOBI + 3 passive: the tracer batching of the 0.5 / 0.1 degree
workloads), a convection stress case, the diagt1 inventories against the oracle, two slabs WITH MOBI, and conservation at
1e-14.  Both metrics are reported: the field-normalised one the 1e-12 gate uses and the point-wise relative one."""
import numpy as np
import pytest

from conftest import load_pkg
from helpers import make_oracle, make_oracle_o3, oracle_rotate, oracle_set_step, pointwise_relerr, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    return load_pkg()


# Calcite on deep grids: its source is dissk0 * max(0, 1 - Omega_c) * caco3 with Omega_c -> 1 at depth, so a last-bit
# difference in the carbonate constants is amplified by 1 / (1 - Omega_c).  The reference's OWN arithmetic shows it: the
# oracle compiled with and without FMA contraction differs by 1e-11 in caco3 / caco3c13 at km = 61 and by 1e-15 in every
# other tracer (tests/test_cpu_oracle.py::test_calcite_is_ill_conditioned_on_deep_grids).  Gate: 1e-10 for these two.
def _tol(nm):
    return 1e-10 if nm in ("caco3", "caco3c13") else 1e-12


def _names40(pkg):
    return pkg.synthetic.default_tracer_names(37) + ["passive0", "passive1", "passive2"]


@pytest.mark.parametrize("imt,jmt,km", [(34, 26, 40), (23, 30, 61)])
def test_step_parity_at_benchmark_depths_40_tracers(pkg, imt, jmt, km):
    """two leapfrog steps + one mixing step + one leapfrog step, nt = 40, km = 40 / 61; per-step gate (every step starts from
    the oracle's state).  Gate per tracer and step: 1e-12, or -- where the step itself is ill conditioned -- 20 x the difference
    between two builds of the ORACLE's own source (strict and -O3 with FMA contraction) on the same state.  On these deep grids
    that only widens the carbon tracers fed by calcite dissolution (caco3, caco3c13; alk, dic, dic13, c14 on the later steps)."""
    case = pkg.synthetic.make_case(imt=imt, jmt=jmt, km=km, nt=40, names=_names40(pkg), seed=km)
    assert case.has_mobi and case.nsrc == 35
    o = make_oracle(case, do_mobi=1)
    o3 = make_oracle_o3(case, do_mobi=1)
    ctx = pkg.TracerContext(case, mobi=1)
    ctx.load_state()
    worst, worst_pw, widened = 0.0, 0.0, set()
    for itt, lf in enumerate((True, True, False, True)):
        for q in (o, o3):
            oracle_set_step(q, case, lf)
            q.call("ora_step")
        ctx.step(leapfrog=lf)
        got, ref, alt = ctx.download_t(+1), o.t()[2], o3.t()[2]
        assert np.array_equal(got[:, 1:-1] == 0, ref[:, 1:-1] == 0)                  # kmt / land-mask indexing: bit exact
        for n, nm in enumerate(case.tracer_names):
            e = relerr(got[n, 1:-1], ref[n, 1:-1])
            cond = relerr(alt[n, 1:-1], ref[n, 1:-1])
            tol = max(1e-12, 20.0 * cond)
            if tol > 1e-12:
                widened.add(nm)
            else:
                worst, worst_pw = max(worst, e), max(worst_pw, pointwise_relerr(got[n, 1:-1], ref[n, 1:-1]))
            assert e <= tol, (itt, nm, e, cond)
        ctx.upload_t(+1, ref)      # per-step gate (north_star): the next step starts from the oracle's state on all sides
        o3.t()[2][:] = ref
        for q in (o, o3):
            oracle_rotate(q)
        ctx.rotate()
    print(f"{imt}x{jmt}x{km} nt=40: worst normalised {worst:.2e}, worst point-wise relative (|ref| > 1e-3 max) {worst_pw:.2e}; "
          f"gate widened by the oracle's own build-to-build difference for {sorted(widened)}")
    assert widened <= {"caco3", "caco3c13", "alk", "dic", "dic13", "c14"}, widened
    assert worst_pw <= 1e-9
    ctx.close()
    o.close()
    o3.close()


def test_convection_stress_parity(pkg):
    """a cold, salty surface anomaly over 40 % of the ocean: convct2 mixes in well over 20 % of the columns"""
    names = ["temp", "salt", "passive0", "passive1"]
    case = pkg.synthetic.make_case(imt=62, jmt=54, km=19, nt=4, names=names, seed=77)
    rng = np.random.default_rng(77)
    cold = rng.random((case.jmt, case.imt)) < 0.4
    t = case.arrays["t"]
    for lev in (0, 1):
        t[lev, 0, :, 0, :][cold] -= 14.0
        t[lev, 0, :, 1, :][cold] -= 7.0
        t[lev, 1, :, 0, :][cold] += 1.5e-3
    t *= case["tmask"][None, None]
    t[..., 0] = t[..., -2]      # cyclic boundary columns stay copies of the interior (setbcx)
    t[..., -1] = t[..., 1]
    o, o_off = make_oracle(case), make_oracle(case, do_convect=0)
    ctx = pkg.TracerContext(case)
    ctx.load_state()
    for itt in range(3):
        for q in (o, o_off):
            oracle_set_step(q, case, True)
            q.call("ora_step")
        ctx.step(True)
        got, ref = ctx.download_t(+1), o.t()[2]
        for n, nm in enumerate(names):
            assert relerr(got[n, 1:-1], ref[n, 1:-1]) <= 1e-12, (itt, nm)
        if itt == 0:
            mixed = (o.t()[2][0] != o_off.t()[2][0]).any(axis=1)
            wet = case["kmt"] > 1
            frac = mixed[wet].mean()
            assert frac >= 0.2, frac
        for q in (o, o_off):
            oracle_rotate(q)
        ctx.rotate()
    ctx.close()
    o.close()
    o_off.close()


def test_tbar_sumbk_against_oracle(pkg):
    """uvic_b200_tbar / uvic_b200_sumbk (09/mom/tracer.F:1516-1565) against ora_diag_tbar (itself bitwise equal to the
    translated reference's diagt1, tests/test_cpu_refpin.py)"""
    names = ["temp", "salt", "passive0", "passive1"]
    case = pkg.synthetic.make_case(imt=42, jmt=38, km=12, nt=4, names=names, seed=5)
    o = make_oracle(case)
    oracle_set_step(o, case, True)
    o.call("ora_step")
    ctx = pkg.TracerContext(case)
    ctx.load_state()
    ctx.step(True, diag=True)
    tb, sb = ctx.tbar(), ctx.sumbk()
    tv, da = ctx.travar_dtabs()
    for n in range(1, case.nt + 1):
        o.call("ora_diag_tbar", n)
    ref_tb = o.arr("tbar", (case.jmt, case.nt, case.km))[1:-1]
    ref_tv, ref_da = o.arr("travar", (case.jmt, case.nt, case.km))[1:-1], o.arr("dtabs", (case.jmt, case.nt, case.km))[1:-1]
    assert np.abs(ref_tv).max() > 0 and np.abs(ref_da).max() > 0
    for n in range(case.nt):
        assert relerr(tv[:, n], ref_tv[:, n]) <= 1e-13 and relerr(da[:, n], ref_da[:, n]) <= 1e-13, (n, relerr(tv[:, n], ref_tv[:, n]), relerr(da[:, n], ref_da[:, n]))
    ref_sb = o.arr("sumbk", (case.nt, case.km, 3))
    assert np.abs(ref_tb).max() > 0 and np.abs(ref_sb).max() > 0
    for n in range(case.nt):
        assert relerr(tb[:, n], ref_tb[:, n]) <= 1e-13, (n, relerr(tb[:, n], ref_tb[:, n]))
        assert relerr(sb[n], ref_sb[n]) <= 1e-13, (n, relerr(sb[n], ref_sb[n]))
    ctx.close()
    o.close()


def test_two_slabs_with_mobi_match_single_context(pkg):
    """latitude slabs with the full MOBI tracer set: bit-identical to one context over leapfrog + mixing steps"""
    case = pkg.synthetic.make_case(imt=34, jmt=30, km=10, nt=37, seed=19)
    one = pkg.TracerContext(case, mobi=1)
    lo = pkg.TracerContext(case, jlo=2, jhi=14, mobi=1)
    hi = pkg.TracerContext(case, jlo=15, jhi=case.jmt - 1, mobi=1)
    for c in (one, lo, hi):
        c.load_state()
    for itt, lf in enumerate((True, True, False, True)):
        for c in (one, lo, hi):
            c.step(leapfrog=lf)
        ref = one.download_t(+1)
        a, b = lo.download_t(+1), hi.download_t(+1)
        assert np.array_equal(a[:, lo.jlo - lo.jbase: lo.jhi - lo.jbase + 1], ref[:, lo.jlo - 1: lo.jhi]), itt
        assert np.array_equal(b[:, hi.jlo - hi.jbase: hi.jhi - hi.jbase + 1], ref[:, hi.jlo - 1: hi.jhi]), itt
        a[:, lo.jhi + 1 - lo.jbase: lo.jhi + 3 - lo.jbase] = b[:, hi.jlo - hi.jbase: hi.jlo - hi.jbase + 2]
        b[:, hi.jlo - 2 - hi.jbase: hi.jlo - hi.jbase] = a[:, lo.jhi - 1 - lo.jbase: lo.jhi + 1 - lo.jbase]
        lo.upload_t(+1, a)
        hi.upload_t(+1, b)
        for c in (one, lo, hi):
            c.rotate()
    for c in (one, lo, hi):
        c.close()


def test_inventory_conserved_to_1e14_per_step(pkg):
    """north_star: global inventories conserved to 1e-14 relative per step (zero surface / bottom flux; passive tracers and
    T, S through FCT + isopycnal mixing + implicit solve + convection)"""
    names = ["temp", "salt", "passive0", "passive1"]
    case = pkg.synthetic.make_case(imt=102, jmt=102, km=19, nt=4, names=names, seed=3)
    ctx = pkg.TracerContext(case)
    ctx.load_state()
    worst = 0.0
    for itt in range(4):
        inv0 = ctx.inventory(-1)
        ctx.step(True)
        inv1 = ctx.inventory(+1)
        rel = np.abs(inv1 - inv0) / np.abs(inv0)
        worst = max(worst, rel.max())
        ctx.rotate()
    print(f"worst relative inventory change per step: {worst:.2e}")
    assert worst <= 1e-14, worst
    ctx.close()


@pytest.mark.parametrize("options", [(), ("O_carbon_13",), ("O_carbon_14", "O_mobi_nitrogen_15")])
def test_mobi_option_subsets(pkg, options):
    """BASELINE config 2 and the other subsets run/mk.in can select: MOBI without (some of) the isotope options.  The oracle in
    the same subset mode is bitwise equal to the reference BUILT without the options (tests/test_cpu_refpin.py)."""
    from uvic29_b200 import mobi_params as mp

    names = mp.tracer_names_for(options=options)
    case = pkg.synthetic.make_case(imt=42, jmt=34, km=10, nt=len(names), names=names, seed=23)
    assert case.has_mobi and case.nsrc == len(names) - 2
    if not options:
        assert case.nt == 21
    o = make_oracle(case, do_mobi=1)
    ctx = pkg.TracerContext(case, mobi=1)
    ctx.load_state()
    for itt, lf in enumerate((True, True, False, True)):
        oracle_set_step(o, case, lf)
        o.call("ora_step")
        ctx.step(leapfrog=lf, next_leapfrog=(True, False, True, True)[itt])
        got, ref = ctx.download_t(+1), o.t()[2]
        for n, nm in enumerate(case.tracer_names):
            e = relerr(got[n, 1:-1], ref[n, 1:-1])
            assert e <= 1e-12, (itt, nm, e)
        ctx.upload_t(+1, ref)
        oracle_rotate(o)
        ctx.rotate()
    ctx.close()
    o.close()
