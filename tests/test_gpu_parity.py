"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded synthetic inputs.  Tolerances: masks / integer work bit-exact; FP64 fields within
1e-12 of the field maximum (BASELINE.json north_star); inventories conserved to 1e-14."""
import numpy as np
import pytest

from conftest import load_pkg
from helpers import make_oracle, oracle_rotate, oracle_set_step, relerr

pytestmark = pytest.mark.gpu

TOL = 1e-12


@pytest.fixture(scope="module")
def pkg():
    return load_pkg()


def _case(pkg, imt=102, jmt=102, km=19, nt=4, seed=2901):
    names = ["temp", "salt"] + [f"passive{m}" for m in range(nt - 2)]
    return pkg.synthetic.make_case(imt=imt, jmt=jmt, km=km, nt=nt, names=names, seed=seed)


def _ctx(pkg, case, **kw):
    ctx = pkg.TracerContext(case, **kw)
    ctx.load_state()
    return ctx


@pytest.fixture(scope="module")
def pair(pkg):
    case = _case(pkg)
    o = make_oracle(case)
    oracle_set_step(o, case, True)
    o.call("ora_isopyc")
    o.call("ora_vmixc")
    ctx = _ctx(pkg, case)
    ctx.isopyc()
    ctx.vmixc(True)
    ctx.synchronize()
    yield case, o, ctx
    ctx.close()
    o.close()


def test_elements_bit_exact(pair):
    case, o, ctx = pair
    s3, s3z = ctx.shape3(), ctx.shape3z()
    for name, shape in (("alphai", s3), ("betai", s3), ("ddxt", (2,) + s3), ("ddyt", (2,) + s3), ("ddzt", (2,) + s3z)):
        got = ctx.fetch(name, shape)
        ref = o.arr(name, shape)
        assert np.array_equal(got, ref), (name, np.abs(got - ref).max())


def test_isopyc_coefficients(pair):
    case, o, ctx = pair
    s3 = ctx.shape3()
    for name in ("K11", "K22", "K33"):
        got, ref = ctx.fetch(name, s3)[..., 1:-1], o.arr(name, s3)[..., 1:-1]
        assert np.array_equal(got, ref), (name, relerr(got, ref))


def test_gm_velocities(pair):
    case, o, ctx = pair
    s3, s3z = ctx.shape3(), ctx.shape3z()
    for name, shape in (("adv_vetiso", s3), ("adv_vntiso", s3), ("adv_vbtiso", s3z)):
        got, ref = ctx.fetch(name, shape), o.arr(name, shape)
        if name == "adv_vntiso":
            got, ref = got[..., 1:-1], ref[..., 1:-1]   # columns 1, imt are never used (overwritten by setbcx downstream)
        assert np.array_equal(got, ref), (name, relerr(got, ref))


def test_vmixc_diff_cbt(pair):
    case, o, ctx = pair
    s3 = ctx.shape3()
    got, ref = ctx.fetch("diff_cbt", s3)[1:-1, :, 1:-1], o.arr("diff_cbt", s3)[1:-1, :, 1:-1]
    assert np.array_equal(got, ref), relerr(got, ref)


def test_tracer_step_parity(pair):
    case, o, ctx = pair
    o.call("ora_tracer")
    ctx.tracer(True)
    got = ctx.download_t(+1)
    ref = o.t()[2]
    for n in range(case.nt):
        e = relerr(got[n, 1:-1], ref[n, 1:-1])
        assert e <= TOL, (n, e)
    # land stays exactly where the reference leaves it (masks bit-exact)
    tm = case["tmask"][1:-1]
    assert np.array_equal(got[:, 1:-1][:, tm == 0], ref[:, 1:-1][:, tm == 0])


def test_multi_step_parity_with_mixing_steps(pkg):
    case = _case(pkg, nt=3, seed=7)
    o = make_oracle(case)
    ctx = _ctx(pkg, case)
    itt = 0
    for _ in range(20):
        itt += 1
        lf = pkg.timestep.is_leapfrog(itt, 16)
        oracle_set_step(o, case, lf)
        o.call("ora_step")
        ctx.step(leapfrog=lf)
        got = ctx.download_t(+1)
        ref = o.t()[2]
        for n in range(case.nt):
            e = relerr(got[n, 1:-1], ref[n, 1:-1])
            assert e <= 1e-11, (itt, n, e)
        oracle_rotate(o)
        ctx.rotate()
    ctx.close()
    o.close()


@pytest.mark.parametrize("shape", [(14, 12, 5), (34, 20, 7), (66, 40, 12)])
def test_small_and_ragged_grids(pkg, shape):
    imt, jmt, km = shape
    case = _case(pkg, imt=imt, jmt=jmt, km=km, nt=3, seed=11)
    o = make_oracle(case)
    oracle_set_step(o, case, True)
    o.call("ora_step")
    ctx = _ctx(pkg, case)
    ctx.step(True)
    got, ref = ctx.download_t(+1), o.t()[2]
    for n in range(case.nt):
        assert relerr(got[n, 1:-1], ref[n, 1:-1]) <= TOL
    ctx.close()
    o.close()


def test_all_land_and_flat_bottom(pkg):
    case = _case(pkg, imt=22, jmt=18, km=6, nt=3, seed=5)
    # flat bottom everywhere except the walls
    kmt = case["kmt"]
    kmt[1:-1, :] = case.km
    kmt[0, :] = 0
    kmt[-1, :] = 0
    k = np.arange(1, case.km + 1)[None, :, None]
    case.arrays["tmask"] = (kmt[:, None, :] >= k).astype(np.float64)
    o = make_oracle(case)
    oracle_set_step(o, case, True)
    o.call("ora_step")
    ctx = _ctx(pkg, case)
    ctx.step(True)
    got, ref = ctx.download_t(+1), o.t()[2]
    for n in range(case.nt):
        assert relerr(got[n, 1:-1], ref[n, 1:-1]) <= TOL
    ctx.close()
    o.close()


def test_conservation_and_inventory(pkg):
    case = _case(pkg, nt=4, seed=3)
    ctx = _ctx(pkg, case)
    inv0 = ctx.inventory(-1)
    ctx.step(True)
    inv1 = ctx.inventory(+1)
    # zero surface / bottom flux: advection + diffusion + convection conserve sum(t dV)
    rel = np.abs(inv1 - inv0) / np.abs(inv0)
    assert (rel <= 1e-13).all(), rel
    # the device reduction agrees with a float64 numpy sum and is reproducible bit for bit
    a = case.arrays
    dv = (a["dzt"][None, :, None] * a["dxt"][None, None, :] * (a["cst"] * a["dyt"])[:, None, None]) * a["tmask"]
    dv[..., 0] = 0
    dv[..., -1] = 0
    dv[0] = 0
    dv[-1] = 0
    ref = (a["t"][0] * dv[None]).sum(axis=(1, 2, 3))
    assert np.allclose(inv0, ref, rtol=1e-13, atol=0)
    assert np.array_equal(ctx.inventory(+1), inv1)
    ctx.close()


def test_slab_decomposition_matches_single_context(pkg):
    """Two latitude slabs with 2-row halos (exchanged here by host copies) reproduce the
    single-context result exactly; the NCCL exchange is covered by the gloo test on CPU."""
    case = _case(pkg, imt=42, jmt=38, km=8, nt=3, seed=13)
    one = _ctx(pkg, case)
    mid = 19
    lo = _ctx(pkg, case, jlo=2, jhi=mid)
    hi = _ctx(pkg, case, jlo=mid + 1, jhi=case.jmt - 1)
    for step in range(3):
        for c in (one, lo, hi):
            c.step(True)
        ref = one.download_t(+1)
        a, b = lo.download_t(+1), hi.download_t(+1)
        # owned rows
        assert np.array_equal(a[:, lo.jlo - lo.jbase: lo.jhi - lo.jbase + 1], ref[:, lo.jlo - 1: lo.jhi])
        assert np.array_equal(b[:, hi.jlo - hi.jbase: hi.jhi - hi.jbase + 1], ref[:, hi.jlo - 1: hi.jhi])
        # halo exchange of t(tau+1): 2 rows each way
        a[:, lo.jhi + 1 - lo.jbase: lo.jhi + 3 - lo.jbase] = b[:, hi.jlo - hi.jbase: hi.jlo - hi.jbase + 2]
        b[:, hi.jlo - 2 - hi.jbase: hi.jlo - hi.jbase] = a[:, lo.jhi - 1 - lo.jbase: lo.jhi + 1 - lo.jbase]
        lo.upload_t(+1, a)
        hi.upload_t(+1, b)
        for c in (one, lo, hi):
            c.rotate()
    for c in (one, lo, hi):
        c.close()


@pytest.mark.parametrize("opts", [dict(fct=0), dict(tidal_kv=0), dict(fullconvect=0), dict(fct=0, tidal_kv=0, fullconvect=0)])
def test_option_switches(pkg, opts):
    """cpp options of run/mk.in as run-time switches: O_fct off (2nd-order centred advection +
    explicit GM advective terms, 09/mom/tracer_adv_flx.F:1030-1082), O_tidal_kv off, O_fullconvect off."""
    case = _case(pkg, imt=42, jmt=34, km=9, nt=3, seed=17)
    o = make_oracle(case, do_convect=opts.get("fullconvect", 1))
    o.set_scalar("fct", opts.get("fct", 1))
    o.set_scalar("tidal_kv", opts.get("tidal_kv", 1))
    ctx = pkg.TracerContext(case, **opts)
    ctx.load_state()
    for step in range(2):
        oracle_set_step(o, case, True)
        o.call("ora_step")
        ctx.step(True)
        got, ref = ctx.download_t(+1), o.t()[2]
        for n in range(case.nt):
            assert relerr(got[n, 1:-1], ref[n, 1:-1]) <= TOL, (opts, step, n)
        oracle_rotate(o)
        ctx.rotate()
    ctx.close()
    o.close()


def test_adv_vel_bit_exact(pkg):
    """adv_vel (tracer part, source/mom/adv_vel.F:60-131) on the device against the oracle."""
    case = _case(pkg, imt=42, jmt=34, km=9, nt=3, seed=19)
    o = make_oracle(case)
    o.call("ora_adv_vel")
    ctx = pkg.TracerContext(case)
    ctx.upload_u(case["u"])
    ctx.adv_vel()
    ctx.synchronize()
    s3, s3z = ctx.shape3(), ctx.shape3z()
    got_e, got_n, got_b = ctx.fetch("adv_vet", s3), ctx.fetch("adv_vnt", s3), ctx.fetch("adv_vbt", s3z)
    assert np.array_equal(got_n, o.arr("adv_vnt", s3))
    assert np.array_equal(got_e[1:], o.arr("adv_vet", s3)[1:])
    assert np.array_equal(got_b[1:], o.arr("adv_vbt", s3z)[1:])
    # and they agree with the numpy generator used for the synthetic inputs to round-off
    assert np.allclose(got_n, case["adv_vnt"], rtol=0, atol=1e-13 * np.abs(case["adv_vnt"]).max())
    ctx.close()
    o.close()


def test_host_buffer_step_matches_resident_step(pkg):
    """uvic_b200_tracer_step (host buffers, what the Fortran shim calls) == resident step."""
    case = _case(pkg, imt=34, jmt=26, km=8, nt=3, seed=23)
    a = _ctx(pkg, case)
    b = pkg.TracerContext(case)
    t = case["t"]
    out = np.empty(b.shape_t())
    b.tracer_step_host(np.ascontiguousarray(t[0]), np.ascontiguousarray(t[1]), case["adv_vet"], case["adv_vnt"], case["adv_vbt"],
                       case["stf"], case["btf"], out, leapfrog=True)
    a.step(True)
    assert np.array_equal(out, a.download_t(+1))
    a.close()
    b.close()


def test_fourier_filter_parity(pkg):
    """O_fourfil: filt/filtr on the polar rows (source/common/filt.F, filtr.F) inside the full step,
    on a case whose ocean reaches 86 degrees (land-bounded strips, m = 1, and full cyclic rows, m = 3)."""
    names = ["temp", "salt", "passive0"]
    case = pkg.synthetic.make_case(imt=42, jmt=48, km=6, nt=3, names=names, seed=31, land_lat=86.0)
    o = make_oracle(case)
    o.set_scalar("do_filter", 1)
    ctx = pkg.TracerContext(case, fourfil=1)
    ctx.load_state()
    ref_nofilter = None
    for step in range(3):
        oracle_set_step(o, case, True)
        o.call("ora_step")
        ctx.step(True)
        got, ref = ctx.download_t(+1), o.t()[2]
        for n in range(case.nt):
            assert relerr(got[n, 1:-1], ref[n, 1:-1]) <= TOL, (step, n, relerr(got[n, 1:-1], ref[n, 1:-1]))
        assert np.array_equal(got[..., 0], got[..., -2]) and np.array_equal(got[..., -1], got[..., 1])
        oracle_rotate(o)
        ctx.rotate()
    # the filter did something: compare with an unfiltered context after one step
    a = pkg.TracerContext(case, fourfil=1)
    b = pkg.TracerContext(case, fourfil=0)
    for c in (a, b):
        c.load_state()
        c.step(True)
    assert np.abs(a.download_t(+1) - b.download_t(+1)).max() > 0
    for c in (a, b, ctx):
        c.close()
    o.close()


@pytest.mark.parametrize("shape", [dict(imt=34, jmt=26, km=8), dict(imt=70, jmt=37, km=19), dict(imt=23, jmt=50, km=33),
                                   dict(imt=95, jmt=21, km=45), dict(imt=12, jmt=90, km=61)])
def test_fct_variants_agree_bitwise(pkg, shape, monkeypatch):
    """The marching FCT kernel (k_fct_march: shared-memory staged rows, every face flux formed once),
    the two-pass version (k_fct_rfac + k_update<3>, ratios through HBM) and the merged k_update<1>
    evaluate every expression with the same operands in the same order.  Since round 2 the flux kernels are
    compiled with FMA contraction (the gate is 1e-12 against the oracle, not bit equality), and the compiler
    contracts differently in different kernels: the three variants agree to 1e-13 of the field maximum on
    leapfrog and mixing steps, for one and for several k tiles / row chunks.  What must stay BITWISE is the
    same kernel on a two-slab decomposition of the same grid."""
    case = pkg.synthetic.make_case(nt=4, names=["temp", "salt", "p0", "p1"], seed=3, **shape)
    out = {}
    for mode in ("merged", "split", "march"):
        monkeypatch.setenv("UVIC_B200_FCT", mode)
        ctx = pkg.TracerContext(case)
        ctx.load_state()
        res = []
        for lf in (True, False, True):
            ctx.step(leapfrog=lf)
            res.append(ctx.download_t(+1).copy())
            ctx.rotate()
        out[mode] = res
        ctx.close()
    for a, b, c in zip(out["merged"], out["split"], out["march"]):
        for n in range(case.nt):
            assert relerr(b[n], a[n]) <= 1e-13 and relerr(c[n], a[n]) <= 1e-13, (n, relerr(b[n], a[n]), relerr(c[n], a[n]))
    # slabs: rows 2..jm and jm+1..jmt-1 computed by two contexts from the same state
    jm = case.jmt // 2
    ref = out["march"][0]
    for jlo, jhi in ((2, jm), (jm + 1, case.jmt - 1)):
        ctx = pkg.TracerContext(case, jlo=jlo, jhi=jhi)
        ctx.load_state()
        ctx.step(leapfrog=True)
        got = ctx.download_t(+1)
        lo = jlo - ctx.jbase
        assert np.array_equal(got[:, lo:lo + (jhi - jlo + 1)], ref[:, jlo - 1:jhi])
        ctx.close()


def test_setvbc_and_set_sbc_on_device(pkg):
    """setvbc (09/mom/setvbc.F:60-140) and set_sbc (09/mom/set_sbc.F:36-83 via
    09/mom/tracer.F:1270-1288) on the device.  stf / btf and the accumulators built from an uploaded t(tau+1) are
    bit-exact; a three-step ocean segment driven by the device-side fluxes stays within 1e-12 of the oracle."""
    case = _case(pkg, imt=42, jmt=30, km=8, nt=5, seed=21)
    nt, jmt, imt = case.nt, case.jmt, case.imt
    numsbc = 2 * nt + 4
    rng = np.random.default_rng(3)
    sbc0 = rng.standard_normal((numsbc, jmt, imt)) * 1e-6
    bhf = rng.standard_normal((jmt, imt)) * 1e-6
    flx = np.array([1, 2, 0, 4, 5], dtype=np.int32)
    acc = np.array([6, 7, 8, 0, 9], dtype=np.int32)
    o = make_oracle(case)
    o.call("ora_make_masks")
    o.arr("sbc", (numsbc, jmt, imt))[...] = sbc0
    o.arr("bhf", (jmt, imt))[...] = bhf
    o.set("sbc_flx_index", flx)
    o.set("trsbcindex", acc)
    ctx = _ctx(pkg, case)
    ctx.sbc_setup(numsbc, flx, acc)
    ctx.upload_sbc(sbc0, bhf)
    # setvbc: bit-exact
    o.call("ora_setvbc")
    ctx.setvbc()
    for name in ("stf", "btf"):
        assert np.array_equal(ctx.fetch(name, (nt, jmt, imt)), o.arr(name, (nt, jmt, imt))), name
    # set_sbc on a given t(tau+1): bit-exact, all switch combinations of a segment
    tp1 = rng.standard_normal(ctx.shape_t())
    ctx.upload_t(+1, tp1)
    o.t()[2] = tp1
    for eots, osegs, osege in ((1, 1, 0), (1, 0, 0), (0, 0, 0), (1, 0, 1), (1, 1, 1)):
        for k, v in (("eots", eots), ("osegs", osegs), ("osege", osege), ("ntspos", 3)):
            o.set_scalar(k, v)
        o.call("ora_set_sbc")
        ctx.set_sbc(eots, osegs, osege, 3)
        got, ref = ctx.download_sbc(), o.arr("sbc", (numsbc, jmt, imt))
        assert np.array_equal(got[:, 1:-1], ref[:, 1:-1]), (eots, osegs, osege)
    # a three-step segment with the surface fluxes in play
    ctx.load_state()
    o.load_case(case)
    o.arr("sbc", (numsbc, jmt, imt))[...] = sbc0
    ctx.upload_sbc(sbc0, bhf)
    for step in range(3):
        oracle_set_step(o, case, True)
        o.call("ora_setvbc")
        o.call("ora_step")
        ctx.setvbc()
        ctx.step(True)
        for k, v in (("eots", 1), ("osegs", int(step == 0)), ("osege", int(step == 2)), ("ntspos", 3)):
            o.set_scalar(k, v)
        o.call("ora_set_sbc")
        ctx.set_sbc(1, step == 0, step == 2, 3)
        oracle_rotate(o)
        ctx.rotate()
    got, ref = ctx.download_sbc(), o.arr("sbc", (numsbc, jmt, imt))
    for n in range(nt):
        if acc[n]:
            assert relerr(got[acc[n] - 1][1:-1, 1:-1], ref[acc[n] - 1][1:-1, 1:-1]) <= TOL, n
    ctx.close()
    o.close()


def test_time_averages_on_device(pkg):
    """the tracer part of avgvar / avgout (09/mom/timeavgs.F:206-375, 398-420) on the device,
    bit-exact against the oracle for uploaded t(tau) / stf, for one context and for two slabs."""
    case = _case(pkg, imt=30, jmt=26, km=7, nt=5, seed=4)
    nt, jmt, km, imt = case.nt, case.jmt, case.km, case.imt
    rng = np.random.default_rng(5)
    vflux = rng.standard_normal((jmt, imt))
    gaost = rng.standard_normal(nt)
    o = make_oracle(case)
    o.arr("vflux", (jmt, imt))[...] = vflux
    o.set("gaost", gaost)
    ctxs = [_ctx(pkg, case), _ctx(pkg, case, jlo=2, jhi=12), _ctx(pkg, case, jlo=13, jhi=jmt - 1)]
    for step in range(3):
        tt = rng.standard_normal((nt, jmt, km, imt))
        ff = rng.standard_normal((nt, jmt, imt))
        o.t()[1] = tt
        o.arr("stf", (nt, jmt, imt))[...] = ff
        o.call("ora_avgvar")
        for c in ctxs:
            sl = slice(c.jbase - 1, c.jbase - 1 + c.jl)
            c.upload_t(0, tt[:, sl])
            c._ck(c.L.uvic_b200_upload_vbc(c.h, np.ascontiguousarray(ff[:, sl]).ctypes.data, None))
            c.tavg_accumulate(vflux[sl], gaost)
    o.call("ora_avgout")
    ref_t, ref_f = o.arr("avg_t", (nt, jmt, km, imt)), o.arr("avg_stf", (nt, jmt, imt))
    for c in ctxs:
        avg_t, avg_f, n = c.tavg_fetch()
        assert n == 3
        lo, hi = c.jlo - c.jbase, c.jhi - c.jbase + 1
        assert np.array_equal(avg_t[:, lo:hi], ref_t[:, c.jlo - 1:c.jhi])
        assert np.array_equal(avg_f[:, lo:hi], ref_f[:, c.jlo - 1:c.jhi])
        c.close()
    o.close()


def test_cuda_path_matches_committed_vectors(pkg):
    """The CUDA path against tests/golden/tiny_step.npz (written from the oracle by tests/golden/make_golden.py): masks
    bit-exact, K33 / diff_cbt bit-exact, t(tau+1) of two leapfrog steps and a mixing step within 1e-12."""
    import importlib.util
    import os

    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    ref = np.load(os.path.join(here, "tiny_step.npz"))
    case = pkg.synthetic.make_case(**mg.CASE)
    assert np.array_equal(np.asarray(case["kmt"]), ref["kmt"])
    ctx = _ctx(pkg, case)
    for step, lf in enumerate((True, True, False)):
        ctx.step(leapfrog=lf)
        got = ctx.download_t(+1)
        want = ref[f"t_p1_step{step}"]
        for n in range(case.nt):
            assert relerr(got[n, 1:-1], want[n, 1:-1]) <= TOL, (step, n)
        if step == 0:
            for name in ("K33", "diff_cbt"):
                assert np.array_equal(ctx.fetch(name, ctx.shape3())[..., 1:-1], ref[name][..., 1:-1]), name
        ctx.rotate()
    ctx.close()


def test_state_density_bit_exact(pkg):
    """source/mom/state.F (called from 09/mom/loadmw.F:150-155): rho = dens(T - to, S - so, k) of t(tau), bit-exact."""
    case = _case(pkg, imt=38, jmt=30, km=9, nt=3, seed=8)
    o = make_oracle(case)
    o.call("ora_state")
    ctx = _ctx(pkg, case)
    assert np.array_equal(ctx.state(0), o.arr("rho", ctx.shape3()))
    ctx.close()
    o.close()


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [dict(imt=34, jmt=26, km=8), dict(imt=70, jmt=37, km=19), dict(imt=95, jmt=45, km=45)])
def test_launch_geometry_does_not_change_a_bit(pkg, shape, monkeypatch):
    """The CTA tile of the diffusion kernel (linear cell order, 32 i x 4 levels, 32 i x 2 levels x 2 rows) and the length
    of the march's row chunks (warm-up rows recomputed per chunk) are launch geometry: t(tau+1) must be bit-identical
    whatever they are, over leapfrog and mixing steps with isopycnal mixing and the FCT."""
    case = pkg.synthetic.make_case(nt=4, names=["temp", "salt", "p0", "p1"], seed=17, **shape)

    def run(env):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        ctx = pkg.TracerContext(case)
        ctx.load_state()
        res = []
        for lf in (True, False, True):
            ctx.step(leapfrog=lf)
            res.append(ctx.download_t(+1).copy())
            ctx.rotate()
        ctx.close()
        for k in env:
            monkeypatch.delenv(k)
        return res

    base = run({"UVIC_B200_UPD_TILE": "0", "UVIC_B200_FCT_CHUNK": "64"})
    assert np.abs(base[-1]).max() > 0
    for env in ({"UVIC_B200_UPD_TILE": "1"}, {"UVIC_B200_UPD_TILE": "2"}, {"UVIC_B200_FCT_CHUNK": "8"}, {"UVIC_B200_FCT_CHUNK": "1000"}, {}):
        got = run(env)
        for a, b in zip(base, got):
            assert np.array_equal(a, b), env
