"""The CUDA path against the REFERENCE'S OWN CODE: against the golden vectors the translated reference wrote
(tests/golden/ref_step_t.npz) and, when the built library travelled to this box, against oracle/_ref/libref_s.so run side
by side (stress case: convection in a third of the columns, concentrations below trcmin, surface fluxes)."""
import os

import numpy as np
import pytest

import reflib
from conftest import load_pkg
from helpers import pointwise_relerr, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    return load_pkg()


def test_cuda_step_matches_reference_golden_vectors(pkg):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_step_t.npz"))
    case = pkg.synthetic.make_case(imt=20, jmt=16, km=6, nt=37, seed=int(g["seed"]))
    ctx = pkg.TracerContext(case, mobi=1, fourfil=0)
    ctx.load_state()
    for itt, lf in enumerate(g["schedule"]):
        ctx.step(leapfrog=bool(lf))
        got, ref = ctx.download_t(+1), g[f"t_p1_step{itt}"]
        assert np.array_equal(got[:, 1:-1] == 0, ref[:, 1:-1] == 0), "land / below-bottom cells"      # mask rule, bit exact
        for n, nm in enumerate(case.tracer_names):
            assert relerr(got[n, 1:-1], ref[n, 1:-1]) <= 1e-12, (itt, nm, relerr(got[n, 1:-1], ref[n, 1:-1]))
        ctx.upload_t(+1, np.ascontiguousarray(ref))     # per-step gate: continue from the reference's state
        ctx.rotate()
    ctx.close()


@pytest.mark.skipif(not os.path.exists(os.path.join(reflib.REFDIR, "libref_s.so")), reason="oracle/_ref/libref_s.so did not travel")
def test_cuda_step_matches_translated_reference_side_by_side(pkg):
    import test_cpu_refpin as T

    ref = reflib.RefLib("s")
    case, o = T.setup_pair(pkg, ref, seed=11, fourfil=True)
    T._stress(case, o, ref, np.random.default_rng(11))
    # the stressed state into the case the CUDA context loads
    case.arrays["t"] = o.t()[:2].copy()
    case.arrays["stf"] = o.arr("stf").reshape(case.nt, case.jmt, case.imt).copy()
    case.arrays["btf"] = o.arr("btf").reshape(case.nt, case.jmt, case.imt).copy()
    ctx = pkg.TracerContext(case, mobi=1, fourfil=1)
    ctx.load_state()
    worst = 0.0
    for itt, lf in enumerate((True, True, False, True)):
        T.ref_set_step(ref, o, case, lf)
        ref.set("first", 1 if itt == 0 else 0)
        T.ref_step(ref)
        ctx.step(leapfrog=lf)
        got, want = ctx.download_t(+1), ref.view("t")[2]
        for n, nm in enumerate(case.tracer_names):
            e = relerr(got[n, 1:-1], want[n, 1:-1])
            worst = max(worst, e)
            assert e <= 1e-12, (itt, nm, e)
        ctx.upload_t(+1, np.ascontiguousarray(want))    # per-step gate: continue from the reference's state
        T.ref_rotate(ref)
        ctx.rotate()
    print(f"CUDA vs translated reference, 4 steps, 37 tracers: worst normalised difference {worst:.2e}")
    ctx.close()
    o.close()


@pytest.mark.skipif(not os.path.exists(os.path.join(reflib.REFDIR, "libref_s.so")), reason="oracle/_ref/libref_s.so did not travel")
def test_one_model_year_against_translated_reference(pkg):
    """north_star: basin-mean T / S / DIC within 1e-8 relative after one model year -- 292 ocean steps (dtts = 108000 s,
    a mixing step every 16th, relyr advancing as `mom` advances it) of the 37-tracer MOBI configuration, the CUDA path against
    THE REFERENCE'S OWN CODE (oracle/_ref/libref_s.so) evolving freely side by side on the 34x26x8 grid."""
    import test_cpu_refpin as T

    ref = reflib.RefLib("s")
    case, o = T.setup_pair(pkg, ref, seed=12)
    o.close()
    ctx = pkg.TracerContext(case, mobi=1, fourfil=0)
    ctx.load_state()
    relyr0, dty = float(case.scalars["relyr"]), float(case.scalars["dtts"]) / (365.0 * 86400.0)
    itt = 0
    for _ in range(292):
        itt += 1
        lf = pkg.timestep.is_leapfrog(itt, 16)
        ry = relyr0 + (itt - 1) * dty
        ref.set("relyr", ry)
        T.ref_set_step(ref, None, case, lf)
        ref.set("first", 0)
        T.ref_step(ref)
        T.ref_rotate(ref)
        ctx.set_time(ry, ry + dty)      # the hint's time is formed like the Fortran shim forms it: not bit-equal to the next ry
        ctx.step(leapfrog=lf, next_leapfrog=pkg.timestep.is_leapfrog(itt + 1, 16))
        ctx.rotate()
    gt, rt = ctx.download_t(0), ref.view("t")[1]
    assert np.isfinite(gt).all() and np.isfinite(rt).all()
    h, m = ctx.lookahead_stats()
    assert h >= 290 and m <= 1, (h, m)          # the look-ahead MOBI was computed for the right model time every step
    a = case.arrays
    kmt, msk = np.asarray(case["kmt"]), np.asarray(case["mskhr"])
    vol = (np.asarray(a["dzt"])[None, :, None] * (np.asarray(a["cst"]) * np.asarray(a["dyt"]))[:, None, None] * np.asarray(a["dxt"])[None, None, :])
    wet = (np.arange(1, case.km + 1)[None, :, None] <= kmt[:, None, :])
    worst = {}
    for nm in ("temp", "salt", "dic", "alk", "o2", "po4", "no3"):
        n = case.tracer_names.index(nm)
        for reg in (1, 2, 3):
            w = (vol * wet * (msk == reg)[:, None, :])[1:-1, :, 1:-1]
            if w.sum() == 0:
                continue
            mg, mr = (gt[n, 1:-1, :, 1:-1] * w).sum() / w.sum(), (rt[n, 1:-1, :, 1:-1] * w).sum() / w.sum()
            scale = max(abs(mr), np.abs(rt[n]).max() * 1e-6)
            worst[nm] = max(worst.get(nm, 0.0), abs(mg - mr) / scale)
    print("basin means after one model year, CUDA vs translated reference, worst relative difference:", {k: f"{v:.1e}" for k, v in worst.items()})
    for nm in ("temp", "salt", "dic"):
        assert worst[nm] <= 1e-8, (nm, worst[nm])
    ctx.close()
