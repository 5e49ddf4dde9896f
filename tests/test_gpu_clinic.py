"""GPU parity of the baroclinic momentum step: uvic_b200_clinic through the C ABI against the
oracle's adv_vel (U part) + setvbc (momentum part) + clinic on the same seeded inputs.  The device keeps the reference's
operation order and is compiled without FMA contraction, so every field is compared bit for bit."""
import numpy as np
import pytest

from conftest import load_pkg
from helpers import make_oracle, oracle_clinic, oracle_load_momentum, oracle_rotate, oracle_set_step, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    return load_pkg()


def _case(pkg, imt, jmt, km, seed):
    case = pkg.synthetic.make_case(imt=imt, jmt=jmt, km=km, nt=2, seed=seed)
    return pkg.synthetic.add_momentum(case)


def _sbc_with_stress(case, numsbc):
    sbc = np.zeros((numsbc, case.jmt, case.imt))
    sbc[0], sbc[1] = case["taux"], case["tauy"]
    return sbc


def _device_clinic(pkg, case, jlo=None, jhi=None, use_slots=True, smf=None, fourfil=False):
    """one context (optionally a slab), loaded from the global case; returns the context after clinic has run"""
    ctx = pkg.TracerContext(case, jlo=jlo, jhi=jhi)
    ctx.load_state()
    sl = lambda n: pkg.api.slab_slice(n, case[n], ctx.jbase, ctx.jl, case)
    ctx.upload_u_level(0, sl("u"))
    ctx.adv_vel()
    ctx.clinic_setup(case, fourfil=fourfil)
    ctx.upload_u_level(-1, sl("um1"))
    if use_slots:
        numsbc = 2 * case.nt + 4
        ctx.sbc_setup(numsbc, np.zeros(case.nt, dtype=np.int32), np.zeros(case.nt, dtype=np.int32))
        ctx.upload_sbc(_sbc_with_stress(case, numsbc)[:, ctx.jbase - 1:ctx.jbase - 1 + ctx.jl], None)
        ctx.clinic(case.scalars["c2dtuv"], 1, 2)
    else:
        ctx.upload_smf(smf[:, ctx.jbase - 1:ctx.jbase - 1 + ctx.jl])
        ctx.clinic(case.scalars["c2dtuv"], 0, 0)
    ctx.synchronize()
    return ctx


def _oracle_clinic(case):
    o = make_oracle(case)
    oracle_load_momentum(o, case)
    oracle_clinic(o)
    return o


@pytest.mark.parametrize("dims", [(34, 30, 8, 33), (102, 102, 19, 2901)])
def test_clinic_bit_exact(pkg, dims):
    imt, jmt, km, seed = dims
    case = _case(pkg, imt, jmt, km, seed)
    o = _oracle_clinic(case)
    ctx = _device_clinic(pkg, case)
    s3, s3z, s2 = (jmt, km, imt), (jmt, km + 1, imt), (2, jmt, imt)
    bad = []

    def check(name, got, ref):
        if not np.array_equal(got, ref):
            den = max(np.abs(ref).max(), 1e-300)
            bad.append((name, float(np.abs(got - ref).max() / den), int((got != ref).sum())))

    # U-cell advective velocities: rows the reference computes (adv_vel.F:168,196,226)
    check("adv_vnu", ctx.fetch("adv_vnu", s3)[0:jmt - 1], o.arr("adv_vnu", s3)[0:jmt - 1])
    check("adv_veu", ctx.fetch("adv_veu", s3)[1:jmt - 1], o.arr("adv_veu", s3)[1:jmt - 1])
    check("adv_vbu", ctx.fetch("adv_vbu", s3z)[1:jmt - 1], o.arr("adv_vbu", s3z)[1:jmt - 1])
    check("smf", ctx.fetch("smf", s2), o.arr("smf", s2))
    check("bmf", ctx.fetch("bmf", s2), o.arr("bmf", s2))
    check("rho", ctx.fetch("rho", s3), o.arr("rho", s3))
    wet = case["umask"][1:-1, :, 1:-1] > 0
    gp_d = ctx.fetch("grad_p", (2,) + s3)[:, 1:-1, :, 1:-1]
    gp_o = o.arr("grad_p", (2,) + s3)[:, 1:-1, :, 1:-1]
    check("grad_p", gp_d * wet[None], gp_o * wet[None])
    up_d, up_o = ctx.download_u(+1), o.arr("up1", (2,) + s3)
    check("u(tau+1)", up_d[:, 1:-1], up_o[:, 1:-1])
    check("zu", ctx.download_zu()[:, 1:-1, 1:-1], o.arr("zu", s2)[:, 1:-1, 1:-1])
    assert not bad, bad
    assert np.abs(up_d).max() > 0
    # size-independent properties of the result: pure internal mode, zero on land, cyclic
    a = case.arrays
    mean = (up_d * a["dzt"][None, None, :, None]).sum(axis=2) * a["hr"][None]
    assert np.abs(mean[:, 1:-1]).max() < 1e-12 * np.abs(up_d).max()
    assert np.all((up_d[:, 1:-1] - a["um1"][:, 1:-1]) * (1.0 - a["umask"][None, 1:-1]) == 0.0)
    assert np.array_equal(up_d[..., 0], up_d[..., -2]) and np.array_equal(up_d[..., -1], up_d[..., 1])
    # the uploaded-smf entry and the time-level rotation
    ctx2 = _device_clinic(pkg, case, use_slots=False, smf=o.arr("smf", s2))
    assert np.array_equal(ctx2.download_u(+1)[:, 1:-1], up_d[:, 1:-1])
    u_tau = ctx2.download_u(0)
    ctx2.rotate_u()
    assert np.array_equal(ctx2.download_u(0)[:, 1:-1], up_d[:, 1:-1]) and np.array_equal(ctx2.download_u(-1), u_tau)
    ctx2.close()
    ctx.close()
    o.close()


def test_clinic_two_slabs_match_one_context(pkg):
    """rows 2..jmt-1 split over two contexts with 2-row halos give the single-context result bit for bit"""
    case = _case(pkg, 34, 30, 8, 33)
    jmt = case.jmt
    full = _device_clinic(pkg, case)
    ref_u, ref_zu = full.download_u(+1), full.download_zu()
    full.close()
    mid = 14
    for jlo, jhi in ((2, mid), (mid + 1, jmt - 1)):
        ctx = _device_clinic(pkg, case, jlo=jlo, jhi=jhi)
        u, zu = ctx.download_u(+1), ctx.download_zu()
        r0 = jlo - ctx.jbase
        n = jhi - jlo + 1
        assert np.array_equal(u[:, r0:r0 + n], ref_u[:, jlo - 1:jhi]), (jlo, jhi)
        assert np.array_equal(zu[:, r0:r0 + n, 1:-1], ref_zu[:, jlo - 1:jhi, 1:-1]), (jlo, jhi)
        ctx.close()


@pytest.mark.parametrize("geo", [dict(imt=42, jmt=48, km=6, seed=31, land_lat=86.0),
                                 dict(imt=26, jmt=44, km=5, seed=12, land_lat=86.0, land_frac=0.04)])
def test_clinic_with_polar_filter(pkg, geo):
    """O_fourfil: filuv (source/common/filuv.F) after the momentum step -- rotation to polar stereographic components,
    land-bounded strips (m = 2, also across the cyclic seam) and, on the second geography, full cyclic rows (m = 3);
    vertical mean removed again, mask.  One context and two slabs, against the oracle."""
    case = pkg.synthetic.make_case(nt=2, **geo)
    pkg.synthetic.add_momentum(case)
    jmt = case.jmt
    o = make_oracle(case)
    o.set_scalar("do_filter", 1)
    oracle_load_momentum(o, case)
    oracle_clinic(o)
    ref = o.arr("up1", (2, jmt, case.km, case.imt)).copy()
    o.set_scalar("do_filter", 0)
    oracle_clinic(o)
    plain = o.arr("up1", ref.shape).copy()
    changed = np.nonzero(np.abs(ref - plain).max(axis=(0, 2, 3)) > 0)[0]
    assert len(changed) >= 4, changed            # the filter did something on both polar caps
    assert changed.min() < jmt // 2 < changed.max()
    ctx = _device_clinic(pkg, case, fourfil=True)
    got = ctx.download_u(+1)
    ctx.close()
    rel = np.abs(got[:, 1:-1] - ref[:, 1:-1]).max() / np.abs(ref).max()
    assert np.array_equal(got[:, 1:-1], ref[:, 1:-1]), rel
    mid = jmt // 2
    for jlo, jhi in ((2, mid), (mid + 1, jmt - 1)):
        ctx = _device_clinic(pkg, case, jlo=jlo, jhi=jhi, fourfil=True)
        u = ctx.download_u(+1)
        r0 = jlo - ctx.jbase
        assert np.array_equal(u[:, r0:r0 + jhi - jlo + 1], ref[:, jlo - 1:jhi]), (jlo, jhi)
        ctx.close()
    o.close()


def test_ocean_steps_tracer_and_clinic_together(pkg):
    """Four ocean steps as mom sequences them (without the barotropic solver): adv_vel from u(tau), isopyc / vmixc /
    tracer, clinic, both sets of time levels rotated -- the device keeps u and t resident throughout.  The tracer step is
    within 1e-12 of the oracle per step (not bit-exact), its density feeds the pressure gradient, so u is compared to
    1e-11 of its maximum."""
    case = pkg.synthetic.make_case(imt=42, jmt=36, km=8, nt=3, names=["temp", "salt", "passive0"], seed=77)
    pkg.synthetic.add_momentum(case)
    jmt, km, imt = case.jmt, case.km, case.imt
    o = make_oracle(case)
    oracle_load_momentum(o, case)
    ctx = pkg.TracerContext(case)
    ctx.load_state()
    ctx.upload_u_level(0, case["u"])
    ctx.clinic_setup(case)
    ctx.upload_u_level(-1, case["um1"])
    numsbc = 2 * case.nt + 4
    ctx.sbc_setup(numsbc, np.zeros(case.nt, dtype=np.int32), np.zeros(case.nt, dtype=np.int32))
    ctx.upload_sbc(_sbc_with_stress(case, numsbc), None)
    c2dtuv = case.scalars["c2dtuv"]
    shu = (2, jmt, km, imt)
    for step in range(4):
        oracle_set_step(o, case, True)
        o.call("ora_adv_vel")
        o.call("ora_step")
        oracle_clinic(o)
        ctx.adv_vel()
        ctx.step(leapfrog=True)
        ctx.clinic(c2dtuv, 1, 2)
        t_d, u_d = ctx.download_t(+1), ctx.download_u(+1)
        t_o, u_o = o.t()[2], o.arr("up1", shu)
        for n in range(case.nt):
            assert relerr(t_d[n][1:-1], t_o[n][1:-1]) < 1e-11, (step, n)
        for n in range(2):
            assert relerr(u_d[n][1:-1], u_o[n][1:-1]) < 1e-11, (step, n)
        assert np.isfinite(u_d).all()
        # tau-1 <- tau <- tau+1 for both
        oracle_rotate(o)
        o.arr("um1", shu)[...] = o.arr("u", shu)
        o.arr("u", shu)[...] = o.arr("up1", shu)
        ctx.rotate()
        ctx.rotate_u()
    # the velocities moved: this was not a fixed point
    assert np.abs(ctx.download_u(0) - case["u"]).max() > 1e-3
    ctx.close()
    o.close()


def test_clinic_against_committed_vectors(pkg):
    """The device against tests/golden/tiny_clinic.npz (written from the oracle by tests/golden/make_golden.py): u(tau+1)
    without and with the polar filter, zu and the pressure gradient, bit for bit."""
    import importlib.util
    import os

    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    ref = np.load(os.path.join(here, "tiny_clinic.npz"))
    case = mg.clinic_case()
    assert np.array_equal(np.asarray(case["kmu"]), ref["kmu"])
    s3 = (case.jmt, case.km, case.imt)
    for tag, filt in (("", False), ("_filuv", True)):
        ctx = _device_clinic(pkg, case, fourfil=filt)
        assert np.array_equal(ctx.download_u(+1)[:, 1:-1], ref["u_p1" + tag][:, 1:-1]), tag
        if not filt:
            assert np.array_equal(ctx.download_zu()[:, 1:-1, 1:-1], ref["zu"][:, 1:-1, 1:-1])
            gp = ctx.fetch("grad_p", (2,) + s3) * np.asarray(case["umask"])[None]
            assert np.array_equal(gp[:, 1:-1, :, 1:-1], ref["grad_p"][:, 1:-1, :, 1:-1])
        ctx.close()


def test_clinic_needs_setup(pkg):
    case = _case(pkg, 34, 30, 8, 33)
    ctx = pkg.TracerContext(case)
    with pytest.raises(pkg.api.UvicError):
        ctx.clinic(1.0)
    ctx.clinic_setup(case)
    with pytest.raises(pkg.api.UvicError):
        ctx.clinic(1.0, 1, 2)      # wind stress slots without uvic_b200_sbc_setup
    ctx.close()
