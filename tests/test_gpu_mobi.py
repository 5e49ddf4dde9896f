"""GPU parity of the MOBI source terms and of the full 37-tracer step against the oracle."""
import numpy as np
import pytest

from conftest import load_pkg
from helpers import make_oracle, oracle_rotate, oracle_set_step, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    return load_pkg()


def _mobi_pair(pkg, **kw):
    case = pkg.synthetic.make_case(nt=37, **kw)
    o = make_oracle(case, do_mobi=1)
    ctx = pkg.TracerContext(case, mobi=1)
    ctx.load_state()
    return case, o, ctx


@pytest.mark.parametrize("ws", ["0", "1"])
def test_mobi_sources_parity(pkg, ws, monkeypatch):
    # both column kernels: one thread per column ("0") and warp specialised ("1")
    monkeypatch.setenv("UVIC_B200_MOBI_WS", ws)
    case, o, ctx = _mobi_pair(pkg)
    oracle_set_step(o, case, True)
    o.call("ora_step")
    ctx.step(True)
    shp = (case.nsrc, case.jmt, case.km, case.imt)
    got, ref = ctx.fetch("src", shp), o.arr("src", shp)
    from uvic29_b200 import mobi_params as mp
    for s, nm in enumerate(mp.SOURCE_ORDER):
        e = relerr(got[s][1:-1, :, 1:-1], ref[s][1:-1, :, 1:-1])
        assert e <= 1e-10, (nm, e)
    gt, rt = ctx.download_t(+1), o.t()[2]
    for n, nm in enumerate(case.tracer_names):
        e = relerr(gt[n, 1:-1], rt[n, 1:-1])
        assert e <= 1e-12, (nm, e)
    ctx.close()
    o.close()


@pytest.mark.parametrize("ws", ["0", "1"])
def test_mobi_multi_step_with_mixing(pkg, ws, monkeypatch):
    monkeypatch.setenv("UVIC_B200_MOBI_WS", ws)
    case, o, ctx = _mobi_pair(pkg, imt=42, jmt=34, km=10, seed=5)
    itt = 0
    for _ in range(6):
        itt += 1
        lf = pkg.timestep.is_leapfrog(itt, 4)
        oracle_set_step(o, case, lf)
        o.call("ora_step")
        ctx.step(leapfrog=lf)
        gt, rt = ctx.download_t(+1), o.t()[2]
        for n, nm in enumerate(case.tracer_names):
            e = relerr(gt[n, 1:-1], rt[n, 1:-1])
            assert e <= 1e-11, (itt, nm, e)
        oracle_rotate(o)
        ctx.rotate()
    ctx.close()
    o.close()


def test_mobi_kernels_agree_bitwise(pkg, monkeypatch):
    """The warp-specialised kernel evaluates the same expressions in the same order as the one-thread-per-column kernel.
    With FMA contraction (round 2: the gate is the oracle at 1e-10 / 1e-12, not bit equality) the compiler fuses
    differently in the two kernels: sources agree to 1e-12, tracers to 1e-13 of the field maximum."""
    case = pkg.synthetic.make_case(nt=37, imt=50, jmt=40, km=12, seed=11)
    out = []
    for ws in ("0", "1"):
        monkeypatch.setenv("UVIC_B200_MOBI_WS", ws)
        ctx = pkg.TracerContext(case, mobi=1)
        ctx.load_state()
        ctx.step(True)
        out.append((ctx.fetch("src", (case.nsrc, case.jmt, case.km, case.imt)).copy(), ctx.download_t(+1).copy()))
        ctx.close()
    for q in range(case.nsrc):
        assert relerr(out[1][0][q], out[0][0][q]) <= 1e-12, (q, relerr(out[1][0][q], out[0][0][q]))
    for n in range(case.nt):
        assert relerr(out[1][1][n], out[0][1][n]) <= 1e-13, (n, relerr(out[1][1][n], out[0][1][n]))


def test_mobi_lookahead_is_bitwise_neutral(pkg):
    """Hints (right and wrong ones) never change results: a mispredicted look-ahead is discarded."""
    case = pkg.synthetic.make_case(nt=37, imt=42, jmt=34, km=10, seed=7)
    sched = [True, True, False, True, True]
    runs = []
    for hints in (None, "right", "wrong"):
        ctx = pkg.TracerContext(case, mobi=1)
        ctx.load_state()
        out = []
        for s, lf in enumerate(sched):
            nxt = None
            if hints is not None and s + 1 < len(sched):
                nxt = sched[s + 1] if hints == "right" else (not sched[s + 1])
            ctx.step(leapfrog=lf, next_leapfrog=nxt)
            out.append(ctx.download_t(+1).copy())
            ctx.rotate()
        runs.append(out)
        ctx.close()
    for other in runs[1:]:
        for a, b in zip(runs[0], other):
            assert np.array_equal(a, b)


def test_host_buffer_step_with_lookahead_matches_resident_step(pkg):
    """uvic_b200_tracer_step (copy streams, batched D2H, look-ahead) against the resident call sequence."""
    case = pkg.synthetic.make_case(nt=37, imt=42, jmt=34, km=10, seed=9)
    a = case.arrays
    ref = pkg.TracerContext(case, mobi=1)
    ref.load_state()
    ctx = pkg.TracerContext(case, mobi=1)
    ctx.load_state()
    sched = [True, True, False, True]
    vet, vnt, vbt = (np.ascontiguousarray(a[n]) for n in ("adv_vet", "adv_vnt", "adv_vbt"))
    stf, btf = np.ascontiguousarray(a["stf"]), np.ascontiguousarray(a["btf"])
    out = np.empty(ctx.shape_t())
    for s, lf in enumerate(sched):
        ref.step(leapfrog=lf)
        want = ref.download_t(+1)
        nxt = sched[s + 1] if s + 1 < len(sched) else None
        ctx.tracer_step_host(None, None, vet, vnt, vbt, stf, btf, out, leapfrog=lf, next_leapfrog=nxt)
        assert np.array_equal(out[:, 1:-1], want[:, 1:-1])
        ref.rotate()
        ctx.rotate()
    ref.close()
    ctx.close()


def test_mobi_matches_committed_vectors(pkg):
    """The CUDA MOBI path against tests/golden/tiny_mobi.npz (written from the oracle by tests/golden/make_golden.py):
    sources within 1e-10, t(tau+1) of all 37 tracers within 1e-12."""
    import importlib.util
    import os

    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    ref = np.load(os.path.join(here, "tiny_mobi.npz"))
    case = pkg.synthetic.make_case(**mg.MOBI_CASE)
    ctx = pkg.TracerContext(case, mobi=1)
    ctx.load_state()
    ctx.step(True)
    got = ctx.fetch("src", (case.nsrc, case.jmt, case.km, case.imt))
    for s in range(case.nsrc):
        assert relerr(got[s][1:-1, :, 1:-1], ref["src"][s][1:-1, :, 1:-1]) <= 1e-10, s
    gt = ctx.download_t(+1)
    for n in range(case.nt):
        assert relerr(gt[n, 1:-1], ref["t_p1"][n, 1:-1]) <= 1e-12, n
    ctx.close()


def test_one_model_year_drift(pkg):
    """BASELINE.json north_star: basin-mean T / S / DIC within 1e-8 relative after one model year.  292 ocean steps
    (dtts = 108000 s, run/control.in:3; a mixing step every 16th) of the full 37-tracer MOBI configuration on a small
    grid, CUDA against the oracle: volume-weighted basin means (the three horizontal regions of mskhr, as `sumbk`
    forms them, 09/mom/tracer.F:1548-1565) of every tracer, and the whole fields."""
    case, o, ctx = _mobi_pair(pkg, imt=26, jmt=22, km=8, seed=12)
    kmt = np.asarray(case["kmt"])
    msk = np.asarray(case["mskhr"])
    a = case.arrays
    vol = (np.asarray(a["dzt"])[None, :, None] * (np.asarray(a["cst"]) * np.asarray(a["dyt"]))[:, None, None] *
           np.asarray(a["dxt"])[None, None, :])
    wet = (np.arange(1, case.km + 1)[None, :, None] <= kmt[:, None, :])
    itt = 0
    for _ in range(292):
        itt += 1
        lf = pkg.timestep.is_leapfrog(itt, 16)
        oracle_set_step(o, case, lf)
        o.call("ora_step")
        ctx.step(leapfrog=lf)
        oracle_rotate(o)
        ctx.rotate()
    gt, rt = ctx.download_t(0), o.t()[1]
    assert np.isfinite(gt).all()
    worst = 0.0
    for n, nm in enumerate(case.tracer_names):
        for reg in (1, 2, 3):
            w = (vol * wet * (msk == reg)[:, None, :])[1:-1, :, 1:-1]
            if w.sum() == 0:
                continue
            mg, mr = (gt[n, 1:-1, :, 1:-1] * w).sum() / w.sum(), (rt[n, 1:-1, :, 1:-1] * w).sum() / w.sum()
            scale = max(abs(mr), np.abs(rt[n]).max() * 1e-6)
            worst = max(worst, abs(mg - mr) / scale)
            if nm in ("temp", "salt", "dic"):
                assert abs(mg - mr) <= 1e-8 * scale, (nm, reg, mg, mr)
        assert relerr(gt[n, 1:-1], rt[n, 1:-1]) <= 1e-7, (nm, relerr(gt[n, 1:-1], rt[n, 1:-1]))
    assert worst <= 1e-7, worst
    ctx.close()
    o.close()


def _gasbc_setup(case, rng):
    """a plausible sbc array: surface state from the case's tracers, winds 3-12 m/s, land carbon fluxes"""
    nt, jmt, imt = case.nt, case.jmt, case.imt
    numsbc = 2 * nt + 4
    slots = dict(isst=1, isss=2, issdic=3, issalk=4, issdic13=5, issc14=6, isso2=7, iws=8, inpp=9, isr=10, iburn=11,
                 idicflx=12, idic13flx=13, ic14flx=14, io2flx=15)
    t = np.asarray(case.arrays["t"])[1]            # tau
    name = {n: q for q, n in enumerate(case.tracer_names)}
    sbc = np.zeros((numsbc, jmt, imt))
    for key, tr in (("isst", "temp"), ("isss", "salt"), ("issdic", "dic"), ("issalk", "alk"), ("issdic13", "dic13"),
                    ("issc14", "c14"), ("isso2", "o2")):
        sbc[slots[key] - 1] = t[name[tr], :, 0, :]
    sbc[slots["iws"] - 1] = rng.uniform(300.0, 1200.0, (jmt, imt))
    for key in ("inpp", "isr", "iburn"):
        sbc[slots[key] - 1] = rng.uniform(0.0, 1e-8, (jmt, imt))
    for key in ("idicflx", "idic13flx", "ic14flx", "io2flx"):
        sbc[slots[key] - 1] = 7.0                  # stale values: must be overwritten where the routine writes
    return numsbc, slots, sbc


def test_gasbc_air_sea_exchange(pkg):
    """the flux loop of gasbc (09/common/gasbc.F:148-266) on the device against the oracle: CO2 /
    13C / 14C / O2 fluxes within 1e-10 (co2calc_SWS stops its Newton solve at |dx| < 1e-10), land fluxes and the cyclic
    boundary exact, with and without the land carbon slots."""
    case, o, ctx = _mobi_pair(pkg, imt=30, jmt=24, km=6, seed=31)
    rng = np.random.default_rng(2)
    numsbc, slots, sbc0 = _gasbc_setup(case, rng)
    o.call("ora_make_masks")
    order = ["isst", "isss", "issdic", "issalk", "issdic13", "issc14", "isso2", "iws", "inpp", "isr", "iburn", "idicflx",
             "idic13flx", "ic14flx", "io2flx"]
    ctx.sbc_setup(numsbc, np.zeros(case.nt, np.int32), np.zeros(case.nt, np.int32))
    for land in (True, False):
        sl = dict(slots)
        if not land:
            sl["inpp"] = 0
        o.arr("sbc", (numsbc, case.jmt, case.imt))[...] = sbc0
        o.set("gas_idx", np.array([sl[k] for k in order], dtype=np.int32))
        for k, v in (("co2ccn", 283.0), ("dc13ccn", -6.5), ("dc14ccn", 0.0)):
            o.set_scalar(k, v)
        o.call("ora_gasbc")
        ctx.upload_sbc(sbc0, None)
        ctx.gasbc(sl, 283.0, -6.5, 0.0)
        got, ref = ctx.download_sbc(), o.arr("sbc", (numsbc, case.jmt, case.imt))
        ocean = np.asarray(case["kmt"]) > 0
        for key in ("idicflx", "idic13flx", "ic14flx", "io2flx"):
            g, r = got[sl[key] - 1], ref[sl[key] - 1]
            assert relerr(g[1:-1][ocean[1:-1]], r[1:-1][ocean[1:-1]]) <= 1e-10, (land, key)
            assert np.array_equal(g[1:-1][~ocean[1:-1]], r[1:-1][~ocean[1:-1]]), (land, key)     # land: exact
            assert np.array_equal(g[1:-1, 0], g[1:-1, -2]) and np.array_equal(g[1:-1, -1], g[1:-1, 1])
            assert np.abs(r[1:-1][ocean[1:-1]]).max() > 0
        for m in range(numsbc):                                   # nothing else is touched
            if m + 1 not in (sl["idicflx"], sl["idic13flx"], sl["ic14flx"], sl["io2flx"]):
                assert np.array_equal(got[m], sbc0[m]), m
    ctx.close()
    o.close()
