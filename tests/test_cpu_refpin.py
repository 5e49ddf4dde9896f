"""The oracle pinned against the REFERENCE'S OWN CODE.

oracle/_ref/libref_<tag>.so is the reference Fortran (cpp-expanded with the options of run/mk.in, exactly as `mk` does)
translated to C by a mechanical translator (oracle/refgen/f2c.py: no hand editing, one rule set for every routine) and
compiled with the flags of the hand-written oracle.  These tests drive both with the same inputs and demand BITWISE
equality: the mask rule, every coefficient of isopyc / vmixc, the FCT and isopycnal fluxes, mobi_init's parameter set,
mobi_src on 10^4 random cells (including concentrations below trcmin), co2calc_SWS, state, adv_vel, the Fourier filter,
the diagt1 inventories, and the complete `tracer` call (MOBI + FCT + isoflux + ivdift/invtri + convct2 + filt) over
leapfrog and mixing steps on all 37 tracers.
"""
import ctypes
import os

import numpy as np
import pytest

import reflib
from conftest import load_pkg
from helpers import make_oracle, oracle_rotate, oracle_set_step

HAVE_REF = os.path.isdir(reflib.REFERENCE) or os.path.exists(os.path.join(reflib.REFDIR, "libref_s.so"))
pytestmark = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built and /root/reference not present")


@pytest.fixture(scope="module")
def pkg():
    return load_pkg()


@pytest.fixture(scope="module")
def ref():
    return reflib.RefLib("s")


def oget(o, n):
    return o.L.ora_get_scalar(o.h, n.encode())


def setup_pair(pkg, ref, seed=3, mobi=1, fourfil=False, names=None, momentum=False, **kw):
    """one synthetic case in the oracle and in the translated reference's COMMON blocks"""
    from uvic29_b200 import mobi_params as mp

    d = ref.dims
    names = names or mp.tracer_names_for()
    case = pkg.synthetic.make_case(imt=d["imt"], jmt=d["jmt"], km=d["km"], nt=len(names), names=names, seed=seed, **kw)
    if momentum:
        pkg.synthetic.add_momentum(case)
    o = make_oracle(case, do_mobi=mobi)
    o.set_scalar("do_filter", 1 if fourfil else 0)
    imt, jmt, km = d["imt"], d["jmt"], d["km"]
    # every COMMON scalar / array back to zero (static storage): the library is shared by the tests
    for nm, ents in ref.man["commons"].items():
        for e in ents:
            ref.view(nm, e["block"])[...] = 0
    o.call("ora_make_masks")
    reflib.oracle_to_ref(o, ref)
    for n in ["ahisop", "athkdf", "slmxr", "c2dtts", "dtts", "aidif", "kappa_h", "zetar", "ogamma", "gravrho0r", "diff_cet",
              "diff_cnt", "relyr", "co2ccn"]:
        ref.set(n, oget(o, n))
    ref.set("taum1", -1), ref.set("tau", 0), ref.set("taup1", 1)
    ref.set("grav", 980.6)                                    # source/common/pconst.h via setcom
    pi = 4.0 * np.arctan(1.0)
    ref.set("pi", pi), ref.set("radian", 360. / (2. * pi))   # source/common/setcom.F
    # mobi_init: namelist defaults, &mobi of run/control.in (through the namelist hook), unit conversion, sinking speeds
    ref.set_namelist_values()
    ref.set("daylen", 86400.0)
    ref.call("mobi_init")
    reflib.oracle_to_ref(o, ref, only=["fe_hydr", "fe_atmdep"])   # mobi_init zeroes them before its (dropped) file reads
    # tracer / source index maps: what tracer_init assigns (09/common/UVic_ESCM.F:1282-1483)
    for n, nm in enumerate(case.tracer_names):
        ref.set({"temp": "itemp", "salt": "isalt"}.get(nm, "i" + nm), n + 1)
    for s, nm in enumerate([q for q in mp.SOURCE_ORDER if q in case.tracer_names]):
        ref.set("is" + nm, s + 1)
    for nm in ("aice", "hice", "hsno"):                       # (imt,jmt,2): the tracer step reads time level 2
        ref.view(nm)[1] = o.arr(nm).reshape(jmt, imt)
    if fourfil:
        for n in ("jfrst", "jft0", "jft1", "jft2"):
            ref.set(n, int(oget(o, n)))
        # jskpt as source/common/setcom.F:83 sets it; the strip tables from the reference's own findex
        ref.set("jskpt", int(oget(o, "jft2")) - int(oget(o, "jft1")))
        ref.call("findex", ref.view("kmt"), 50, km, int(oget(o, "jft1")), int(oget(o, "jft2")), imt, ref.view("istf"), ref.view("ietf"))
    else:
        ref.set("jfrst", jmt + 10), ref.set("jft1", 0), ref.set("jft2", jmt + 10)
    return case, o


def ref_set_step(ref, o, case, leapfrog):
    dtts = case.scalars["dtts"]
    ref.set("c2dtts", 2.0 * dtts if leapfrog else dtts)        # source/mom/mom.F:107-137
    if not leapfrog:
        t = ref.view("t")
        t[0] = t[1]                                              # 09/mom/loadmw.F:109-111


def ref_step(ref):
    imt, jmt = ref.dims["imt"], ref.dims["jmt"]
    ref.call("isopyc", 0, 1, jmt, 1, imt)                        # source/mom/mom.F:340
    ref.call("vmixc", 0, 1, jmt, 1, imt)                         # :347
    ref.call("tracer", 0, 2, jmt - 1, 2, imt - 1)                # :389


def ref_rotate(ref):
    t = ref.view("t")
    t[0] = t[1]
    t[1] = t[2]


def assert_same(o, ref, names, what=""):
    for nm in names:
        dmax, ndiff = reflib.compare(o, ref, nm)
        assert ndiff == 0, (what, nm, dmax, ndiff)


def test_translation_manifest(ref):
    """what was translated, and that nothing but I/O was dropped"""
    rt = ref.man["routines"]
    for r in ("tracer", "adv_flux", "isoflux", "ivdift", "invtri", "convct2", "elements", "ai_east", "ai_north", "ai_bottom",
              "isopyc_adv", "vmixc", "mobi_init", "mobi_driver", "mobi_src", "co2calc_sws", "drtsafe", "ta_iter_sws", "state",
              "adv_vel", "setbcx", "filt", "filtr", "findex", "diagt1", "set_sbc", "setvbc", "clinic", "filuv", "gasbc", "areaavg"):
        assert r in rt, r
    assert all(d["why"] in ("I/O", "I/O helper", "CHARACTER assignment", "CHARACTER expression") for d in ref.man["dropped"] if d["unit"] != "gasbc")
    assert "-DO_mobi" in ref.man["cpp_options"] and "-DO_fct" in ref.man["cpp_options"] and "-DO_isopycmix" in ref.man["cpp_options"]


def test_mask_rule(pkg, ref):
    """09/mom/loadmw.F:60-77 is not a separate routine; the translated `tracer` consumes tmask as loaded.  The oracle's
    masks against the kmt rule, bit exact (both sides then use the same array)."""
    case, o = setup_pair(pkg, ref)
    k = np.arange(1, case.km + 1)[None, :, None]
    assert np.array_equal(o.arr("tmask", (case.jmt, case.km, case.imt)), (case["kmt"][:, None, :] >= k).astype(np.float64))
    o.close()


def test_mobi_init_parameters(pkg, ref):
    """mobi_init (09/mom/mobi.F:40-438) with &mobi of run/control.in == uvic2.9_b200/mobi_params.py, every parameter"""
    from uvic29_b200 import mobi_params as mp

    case, o = setup_pair(pkg, ref)
    P = o.raw("mobi_par")
    for i, nm in enumerate(mp.PAR_ORDER):
        multi = len(ref.man["commons"][nm.lower()]) > 1
        r = ref.get(nm, block="npzd_r") if multi else ref.get(nm)
        assert r == P[i], (nm, r, P[i])
    off = mp.N_SCALAR
    for q, nm in enumerate(("wd", "wc", "wo", "ztt")):
        assert np.array_equal(ref.view(nm), P[off + q * mp.KMAX: off + q * mp.KMAX + case.km]), nm
    for m, nm in enumerate(mp.MOBI_STATE):
        assert ref.get("imobi" + nm) == m + 1, nm                # the setimobi sequence (:440-497)
    # &mobi as recorded from run/control.in by gen.py == the hand-copied table in mobi_params.py
    nl = ref.man["namelists"]["mobi"]
    assert {k.lower(): float(v) for k, v in mp.CONTROL_IN.items()} == {k: float(v) for k, v in nl.items()}
    o.close()


def test_isopyc_vmixc_bitwise(pkg, ref):
    case, o = setup_pair(pkg, ref, seed=5)
    o.call("ora_isopyc")
    ref.call("isopyc", 0, 1, case.jmt, 1, case.imt)
    names = ["alphai", "betai", "ddxt", "ddyt", "ddzt", "Ai_ez", "Ai_nz", "Ai_bx", "Ai_by", "K11", "K22", "K33", "adv_vetiso",
             "adv_vntiso", "adv_vbtiso", "drodxte", "drodxbe", "drodytn", "drodybn", "drodzte", "drodzbe", "drodztn", "drodzbn"]
    assert_same(o, ref, names, "isopyc")
    assert np.abs(ref.view("k33")).max() > 0 and np.abs(ref.view("adv_vbtiso")).max() > 0 and np.abs(ref.view("ai_bx")).max() > 0
    o.call("ora_vmixc")
    ref.call("vmixc", 0, 1, case.jmt, 1, case.imt)
    assert_same(o, ref, ["diff_cbt"], "vmixc")
    assert ref.view("diff_cbt").max() > 0.35
    o.close()


def test_adv_flux_isoflux_bitwise(pkg, ref):
    case, o = setup_pair(pkg, ref, seed=7)
    o.call("ora_isopyc"), ref.call("isopyc", 0, 1, case.jmt, 1, case.imt)
    for n in (1, 2, 3, 9, 17, 30):
        o.call("ora_adv_flux", n)
        ref.call("adv_flux", 0, 2, case.jmt - 1, 2, case.imt - 1, n)
        assert_same(o, ref, ["adv_fe", "adv_fn", "adv_fb"], f"adv_flux n={n}")
        assert np.abs(ref.view("adv_fe")).max() > 0
        for nm in ("diff_fe", "diff_fn", "diff_fbiso"):          # tracer sets these before isoflux adds to them
            o.raw(nm)[:] = 0
            ref.view(nm)[...] = 0
        o.call("ora_isoflux", n)
        ref.call("isoflux", 0, 2, case.jmt - 1, 2, case.imt - 1, n)
        assert_same(o, ref, ["diff_fe", "diff_fn", "diff_fbiso"], f"isoflux n={n}")
        assert np.abs(ref.view("diff_fe")).max() > 0
    o.close()


def _stress(case, o, ref, rng):
    """make the step exercise what a smooth start does not: statically unstable columns (a cold, salty surface anomaly over
    a third of the ocean), concentrations below trcmin and exactly zero in scattered cells, surface and bottom fluxes"""
    t = o.t()
    imt, jmt, km, nt = case.imt, case.jmt, case.km, case.nt
    cold = rng.random((jmt, imt)) < 0.35
    for lev in (0, 1):
        t[lev, 0, :, 0, :][cold] -= 12.0
        t[lev, 1, :, 0, :][cold] += 1.5e-3
        t[lev, 0, :, 1, :][cold] -= 6.0
    for n in range(2, nt):
        if case.tracer_names[n] in ("dic", "alk"):               # the carbonate solve has no root for DIC << ALK
            continue
        m = rng.random((jmt, km, imt)) < 0.04
        t[0, n][m] *= 1e-9
        t[1, n][m] *= 1e-9
        # exact zeros, except where the reference itself divides 0/0: ptn_P = phyt_phos/phyt, ptn_detr (09/mom/mobi.F:1781-1784),
        # and O2 = 0 (o2flag = tanh(0) = 0 makes the iron speciation 0/0, :2210-2222)
        if case.tracer_names[n] not in ("phyt", "detr", "o2"):
            z = rng.random((jmt, km, imt)) < 0.01
            t[0, n][z] = 0.0
    t[0] *= case["tmask"][None]
    t[1] *= case["tmask"][None]
    # the cyclic boundary columns are copies of the interior (setbcx, source/common/util.F:789-812): keep them consistent
    t[..., 0] = t[..., -2]
    t[..., -1] = t[..., 1]
    stf, btf = o.arr("stf").reshape(nt, jmt, imt), o.arr("btf").reshape(nt, jmt, imt)
    stf[:] = 1e-7 * rng.standard_normal(stf.shape) * (case["kmt"] > 0)
    btf[:] = 1e-8 * rng.standard_normal(btf.shape) * (case["kmt"] > 0)
    for f in (stf, btf):
        f[..., 0] = f[..., -2]
        f[..., -1] = f[..., 1]
    reflib.oracle_to_ref(o, ref, only=["t", "stf", "btf"])


@pytest.mark.parametrize("seed,stress,fourfil", [(3, False, False), (11, True, False), (4, True, True), (7, True, True), (23, True, False),
                                                 (31, False, True)])
def test_full_tracer_step_bitwise(pkg, ref, seed, stress, fourfil):
    """isopyc -> vmixc -> tracer (MOBI, FCT, isoflux, ivdift/invtri, convct2, filt) as mom sequences them: leapfrog, leapfrog,
    mixing step, leapfrog -- all 37 tracers of t(tau+1) bit for bit after every step"""
    case, o = setup_pair(pkg, ref, seed=seed, fourfil=fourfil)
    if stress:
        _stress(case, o, ref, np.random.default_rng(seed))
    # the segment accumulators for the atmosphere (09/mom/set_sbc.F, called at the end of the reference's `tracer`,
    # 09/mom/tracer.F:1270-1288): every tracer owns slot n of the coupler array; a four-step segment
    nt, jmt, imt = case.nt, case.jmt, case.imt
    o.set("trsbcindex", np.arange(1, nt + 1, dtype=np.int32))
    ref.view("trsbcindex")[...] = np.arange(1, nt + 1)
    o.set_scalar("ntspos", 4), ref.set("ntspos", 4)
    for itt, lf in enumerate((True, True, False, True)):
        oracle_set_step(o, case, lf)
        ref_set_step(ref, o, case, lf)
        ref.set("first", 1 if itt == 0 else 0)                   # source/common/switch.h: filtr builds its tables when `first`
        for k, v in (("eots", 1), ("osegs", int(itt == 0)), ("osege", int(itt == 3))):
            o.set_scalar(k, v), ref.set(k, v)
        o.call("ora_step")
        o.call("ora_set_sbc")
        ref_step(ref)
        osbc, rsbc = o.arr("sbc").reshape(-1, jmt, imt)[:nt], ref.view("sbc")[:nt]
        assert np.array_equal(osbc[..., 1:-1], rsbc[..., 1:-1]) and np.abs(osbc[8]).max() > 0, itt
        to, tr = o.t()[2], ref.view("t")[2]
        for n, nm in enumerate(case.tracer_names):
            nd = int((to[n, 1:-1] != tr[n, 1:-1]).sum())
            assert nd == 0, (itt, nm, nd, float(np.abs(to[n, 1:-1] - tr[n, 1:-1]).max()))
        assert np.isfinite(tr).all() and np.abs(tr[8]).max() > 0
        oracle_rotate(o)
        ref_rotate(ref)
    if stress:
        # the stress case must really convect: compare with a run of the oracle without convct2
        o2 = make_oracle(case, do_mobi=1, do_convect=0)
        o2.t()[:] = o.t()
        o2.arr("stf")[:] = o.arr("stf")
        oracle_set_step(o2, case, True), oracle_set_step(o, case, True)
        o2.call("ora_step"), o.call("ora_step")
        changed = (o2.t()[2][0] != o.t()[2][0]).any(axis=1)
        wet = case["kmt"] > 1
        assert changed[wet].mean() > 0.2, changed[wet].mean()
        o2.close()
    o.close()


def test_mobi_src_random_cells_bitwise(pkg, ref):
    """mobi_src (09/mom/mobi.F:1485-3313) on 10^4 random cells, every one of the 32 state variables random over six decades,
    a fifth of them below trcmin, some exactly zero; increments and all twelve export / rate outputs bit for bit"""
    from uvic29_b200 import mobi_params as mp

    case, o = setup_pair(pkg, ref)
    L = o.L
    L.ora_test_mobi_src.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_double] + [ctypes.c_void_p] * 4
    rng = np.random.default_rng(2901)
    typical = {"po4": 1.0, "dic": 2.2, "dic13": 2.2 * 0.0112, "no3": 15.0, "din15": 15 * 0.0037, "sil": 0.03, "dfe": 5e-4, "don": 3.0, "dop": 0.1}
    in_names = ["gl", "bct", "impo", "dzt", "impo_phos", "dayfrac", "wwd", "nud", "impocaco3", "wwc", "dissk1", "impoopl", "wwo",
                "opl_disk1", "nudop", "nudon", "bctz", "rn15impo", "rc13impo", "ac13b", "rcaco3c13impo", "impofe", "o2", "aou"]
    out_names = ["nfixout", "expoout", "expo_phosout", "calproout", "disslout", "expocaco3out", "expooplout", "rn15expoout",
                 "rc13expoout", "rcaco3c13expoout", "expofeout", "remifeout"]
    spec = [a["name"] for a in ref.man["routines"]["mobi_src"]["args"]]
    ncell, nsub = 10000, 0
    for c in range(ncell):
        bio = np.array([typical.get(nm, 0.1) for nm in mp.MOBI_STATE]) * 10.0 ** rng.uniform(-3, 1, 32)
        low = rng.random(32) < 0.2
        bio[low] = 5e-12 * 10.0 ** rng.uniform(-3, 0.2, low.sum())
        bio[rng.random(32) < 0.02] = 0.0
        bio[mp.MOBI_STATE.index("phyt")] = max(bio[mp.MOBI_STATE.index("phyt")], 1e-30)   # ptn_P = phyt_phos/phyt (:1781)
        bio[mp.MOBI_STATE.index("detr")] = max(bio[mp.MOBI_STATE.index("detr")], 1e-30)
        nsub += int((bio < 5e-12).sum())
        temp = rng.uniform(-1.8, 30.0)
        bct = 1.066 ** temp
        vin = dict(gl=rng.uniform(0, 80.0), bct=bct, impo=rng.uniform(0, 1e-7), dzt=rng.choice([5000., 13000., 40000.]), impo_phos=rng.uniform(0, 1e-8),
                   dayfrac=rng.uniform(1e-12, 1.0), wwd=rng.uniform(1e-7, 4e-6), nud=rng.uniform(1e-7, 3e-6), impocaco3=rng.uniform(0, 1e-8),
                   wwc=rng.uniform(1e-7, 8e-6), dissk1=rng.uniform(0, 2e-7), impoopl=rng.uniform(0, 1e-8), wwo=rng.uniform(1e-7, 1e-5),
                   opl_disk1=rng.uniform(0, 1e-6), nudop=rng.uniform(0, 1e-9) * bct, nudon=rng.uniform(0, 1e-9) * bct,
                   bctz=1.066 ** min(temp, 20.0), rn15impo=rng.uniform(0.003, 0.004), rc13impo=rng.uniform(0.010, 0.012),
                   ac13b=rng.uniform(-0.03, -0.01), rcaco3c13impo=rng.uniform(0.010, 0.012), impofe=rng.uniform(0, 1e-12),
                   o2=rng.uniform(0.0, 350.0) * (rng.random() > 0.1), aou=rng.uniform(-20.0, 250.0))
        nbio = int(rng.integers(1, 9))
        dtbio = 216000.0 / nbio
        capr = rng.uniform(0.0, 0.035)
        # oracle
        bo, bout, vo, vout = bio.copy(), np.zeros(32), np.array([vin[k] for k in in_names]), np.zeros(12)
        L.ora_test_mobi_src(o.h, nbio, dtbio, capr, bo.ctypes.data, vo.ctypes.data, bout.ctypes.data, vout.ctypes.data)
        # translated reference: nbio, dtbio, capr travel through COMMON (09/mom/mobi.h)
        ref.set("nbio", nbio), ref.set("dtbio", dtbio), ref.set("capr", capr)
        br, brout = bio.copy(), np.zeros(32)
        args = []
        for a in spec:
            args.append(br if a == "bioin" else brout if a == "bioout" else vin.get(a, 0.0))
        ref.call("mobi_src", *args)
        assert np.array_equal(bo, br), (c, "bioin is clipped in place (:1894)")
        assert np.array_equal(bout, brout), (c, np.abs(bout - brout).max(), [mp.MOBI_STATE[i] for i in np.nonzero(bout != brout)[0]])
        got = np.array([ref.last[k] for k in out_names])
        assert np.array_equal(vout, got), (c, [out_names[i] for i in np.nonzero(vout != got)[0]])
        assert np.isfinite(brout).all()
    assert nsub > ncell * 3
    o.close()


def test_co2calc_bitwise_and_literature_values(pkg, ref):
    """co2calc_SWS / drtsafe / ta_iter_SWS (09/common/co2calc.F): 10^4 random states bit for bit, then the equilibrium constants
    the routine leaves in COMMON /const/ at S = 35, t = 25 degC, p = 0 against the check values printed in the DOE (1994)
    handbook / Dickson, Sabine & Christian (2007), Guide to best practices, chapter 5"""
    case, o = setup_pair(pkg, ref)
    L = o.L
    L.ora_co2calc_SWS.argtypes = [ctypes.c_double] * 7 + [ctypes.c_void_p] * 8
    rng = np.random.default_rng(7)
    outs = ["ph", "co2star", "dco2star", "pco2", "dpco2", "co3", "omega_c", "omega_a"]
    for c in range(10000):
        t, s = rng.uniform(-1.9, 32.0), rng.uniform(20.0, 40.0)
        dic, ta = rng.uniform(1.8, 2.5), rng.uniform(2.1, 2.6)
        co2, atm, depth = rng.uniform(180.0, 900.0), rng.uniform(0.95, 1.05), rng.uniform(0.0, 5500.0)
        res = (ctypes.c_double * 8)()
        L.ora_co2calc_SWS(t, s, dic, ta, co2, atm, depth, *[ctypes.addressof(res) + 8 * i for i in range(8)])
        ref.call("co2calc_sws", t, s, dic, ta, co2, atm, depth, *([0.0] * 8))
        got = [ref.last[k] for k in outs]
        assert list(res) == got, (c, list(res), got)
    # Literature check values, ln K at S = 35, t = 25 degC, p = 0 (DOE 1994 handbook ch. 5 / Dickson et al. 2007 print them on the
    # TOTAL hydrogen scale; the routine works on the SEAWATER scale, ln K_sws = ln K_tot + ln(1 + (FT/KF)/(1 + ST/KS)) = +0.0223)
    ref.call("co2calc_sws", 25.0, 35.0, 2.0, 2.3, 280.0, 1.0, 0.0, *([0.0] * 8))
    g = lambda n: float(ref.get(n, block="const"))
    st, ft = float(ref.get("st", block="species")), float(ref.get("ft", block="species"))
    conv = np.log(1.0 + (ft / g("kf")) / (1.0 + st / g("ks")))
    assert abs(conv - 0.0223) < 2e-4
    assert abs(np.log(g("k0")) - (-3.5617)) < 1e-4             # Weiss (1974)
    assert abs(np.log(g("ks")) - (-2.30)) < 1e-3               # Dickson (1990), free scale
    # Dickson & Riley (1979), free scale, evaluated by hand: 1590.2/T - 12.641 + 1.525 sqrt(I) + ln(1 - 0.001005 S), I = 0.72276
    assert abs(np.log(g("kf")) - (5.33356 - 12.641 + 1.29648 - 0.035808)) < 1e-4
    assert abs(np.log(g("kb")) - (-19.7964 + conv)) < 5e-4     # Dickson (1990)
    assert abs(np.log(g("kw")) - (-30.434 + conv)) < 1e-2      # Millero (1995); the handbook value is printed to 3 decimals
    assert abs(np.log(g("k1p")) - (-3.71 + conv)) < 1e-2       # two decimals printed
    assert abs(np.log(g("k2p")) - (-13.727 + conv)) < 1e-2
    assert abs(np.log(g("k3p")) - (-20.24 + conv)) < 1e-2
    assert abs(np.log(g("ksi")) - (-21.61 + conv)) < 1e-2
    # Mehrbach et al. (1973) as refit by Dickson & Millero (1987) on the seawater scale (Millero 1995, eqs 35-36), by hand:
    # pK1 = 3670.7/T - 62.008 + 9.7944 ln T - 0.0118 S + 0.000116 S^2 = 12.31159 - 62.008 + 55.80454 - 0.413 + 0.1421
    assert abs(-np.log10(g("k1")) - 5.83723) < 2e-5
    assert abs(-np.log10(g("k2")) - 8.955) < 1e-3
    # the solve itself: surface water of DIC 2.0, ALK 2.3 mol m-3 at 25 degC is pH(sws) 8.03, Omega_calcite 4.97
    assert abs(ref.last["ph"] - 8.0305) < 1e-3 and abs(ref.last["omega_c"] - 4.969) < 2e-3 and abs(ref.last["omega_a"] - 3.275) < 2e-3
    o.close()


def test_state_and_adv_vel_bitwise(pkg, ref):
    case, o = setup_pair(pkg, ref, seed=9)
    imt, jmt, km = case.imt, case.jmt, case.km
    # state (source/mom/state.F), called as 09/mom/loadmw.F:150-155 does for the density clinic uses
    o.call("ora_state")
    t = ref.view("t")
    rho = np.zeros((jmt - 1, km, imt))                            # rho(imt,km,jsmw:jmw): rows 2..jmt
    ref.call("state", np.ascontiguousarray(t[1, 0]), np.ascontiguousarray(t[1, 1]), rho, 2, jmt, 1, imt)
    assert np.array_equal(rho, o.arr("rho", (jmt, km, imt))[1:]) and np.abs(rho).max() > 0
    # adv_vel (source/mom/adv_vel.F:60-131): T-cell advective velocities from u(tau)
    pkg.synthetic.add_momentum(case)
    o.set("u", case["u"])
    ref.view("u")[1] = case["u"]                                  # u(imt,km,jmw,2,-1:1): tau is index 0 -> slot 1
    for nm in ("adv_vet", "adv_vnt", "adv_vbt"):
        o.raw(nm)[:] = 0
        ref.view(nm)[...] = 0
    o.call("ora_adv_vel")
    ref.call("adv_vel", 0, 1, jmt, 1, imt)
    assert_same(o, ref, ["adv_vet", "adv_vnt", "adv_vbt"], "adv_vel")
    assert np.abs(ref.view("adv_vbt")).max() > 0
    o.close()


def test_diagt1_inventories_bitwise(pkg, ref):
    """tbar / travar / dtabs (09/mom/tracer.F:1516-1539) and the basin sums sumbk (:1548-1565)"""
    case, o = setup_pair(pkg, ref, seed=13)
    oracle_set_step(o, case, True), ref_set_step(ref, o, case, True)
    o.call("ora_step"), ref_step(ref)
    ref.set("tsiperts", 1), ref.set("tavgts", 1), ref.set("eots", 1)
    twodt = np.full(case.km, 2.0 * case.scalars["dtts"]) * np.asarray(case["dtxcel"])
    for n in (1, 2, 8, 20):
        ref.view("sumbk")[...] = 0
        ref.view("sumbf")[...] = 0
        o.arr("sumbk")[:] = 0
        ref.call("diagt1", 0, 2, case.jmt - 1, 2, case.imt - 1, n, twodt)
        o.call("ora_diag_tbar", n)
        shp = (case.jmt, case.nt, case.km)
        for nm in ("tbar", "travar", "dtabs"):
            a, b = o.arr(nm, shp)[:, n - 1], ref.view(nm)[:, n - 1, 1:]       # tbar(0:km,nt,jmt): level 0 is the column total
            assert np.array_equal(a, b), (nm, n, np.abs(a - b).max())
        assert np.abs(ref.view("tbar")[:, n - 1, 1:]).max() > 0
        a, b = o.arr("sumbk", (case.nt, case.km, 3))[n - 1], ref.view("sumbk")[n - 1]
        assert np.array_equal(a, b), ("sumbk", n, np.abs(a - b).max())
    o.close()


def test_oracle_reproduces_reference_golden_vectors(pkg):
    """tests/golden/ref_step_t.npz was written by the reference's own (translated) code (tests/golden/make_ref_golden.py);
    the hand-written oracle reproduces it bit for bit.  Needs neither /root/reference nor oracle/_ref."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_step_t.npz"))
    case = pkg.synthetic.make_case(imt=20, jmt=16, km=6, nt=37, seed=int(g["seed"]))
    assert np.array_equal(case["kmt"], g["kmt"])
    o = make_oracle(case, do_mobi=1)
    o.set_scalar("do_filter", 0)
    for itt, lf in enumerate(g["schedule"]):
        oracle_set_step(o, case, bool(lf))
        o.call("ora_step")
        assert np.array_equal(o.t()[2][:, 1:-1], g[f"t_p1_step{itt}"][:, 1:-1]), itt
        oracle_rotate(o)
    o.close()


def test_no_isotope_configuration_bitwise(pkg):
    """BASELINE config 2 (full MOBI tracer set, no isotopes, nt = 21): the reference built WITHOUT O_carbon_13, O_carbon_14 and
    O_mobi_nitrogen_15 (oracle/_ref/libref_n.so) against the oracle run with the isotope tracers marked absent (index 0 in the
    MOBI maps): whole tracer steps, all 21 tracers bit for bit"""
    from uvic29_b200 import mobi_params as mp

    ref = reflib.RefLib("n")
    assert "-DO_carbon_13" not in ref.man["cpp_options"] and "-DO_mobi" in ref.man["cpp_options"]
    names = mp.tracer_names_for(options=())
    assert len(names) == 21 and ref.view("t").shape[1] == 21
    case, o = setup_pair(pkg, ref, seed=17, names=names)
    assert case.nsrc == 19 and (np.asarray(case["mobi_idx"])[:64] == 0).sum() == 2 * 15
    _stress(case, o, ref, np.random.default_rng(17))
    for itt, lf in enumerate((True, True, False, True)):
        oracle_set_step(o, case, lf)
        ref_set_step(ref, o, case, lf)
        o.call("ora_step")
        ref_step(ref)
        to, tr = o.t()[2], ref.view("t")[2]
        for n, nm in enumerate(case.tracer_names):
            nd = int((to[n, 1:-1] != tr[n, 1:-1]).sum())
            assert nd == 0, (itt, nm, nd, float(np.abs(to[n, 1:-1] - tr[n, 1:-1]).max()))
        assert np.isfinite(tr).all()
        oracle_rotate(o)
        ref_rotate(ref)
    o.close()


def _ref_momentum_inputs(ref, o, case):
    """the inputs of setvbc / clinic the by-name copy cannot place: time levels of u, the coupler's stress slots, scalars"""
    imt, jmt = case.imt, case.jmt
    u = ref.view("u")                                              # u(imt,km,jmw,2,-1:...): tau-1, tau are slots 0, 1
    u[0], u[1] = case["um1"], case["u"]
    sbc = ref.view("sbc")
    sbc[0], sbc[1] = case["taux"], case["tauy"]
    ref.set("itaux", 1), ref.set("itauy", 2)
    for n in ("c2dtuv", "kappa_m", "cdbot"):
        ref.set(n, oget(o, n))
    ref.set("grav", 980.6), ref.set("rho0r", 1.0 / 1.035)          # source/common/pconst.h; clinic forms grav*rho0r itself


def _ref_clinic(ref, case):
    imt, jmt = case.imt, case.jmt
    ref.call("adv_vel", 0, 1, jmt, 1, imt)                         # source/mom/mom.F:300-390
    t = ref.view("t")
    ref.call("state", np.ascontiguousarray(t[1, 0]), np.ascontiguousarray(t[1, 1]), ref.view("rho"), 2, jmt, 1, imt)
    ref.call("vmixc", 0, 1, jmt, 1, imt)
    ref.call("setvbc", 0, 1, jmt, 1, imt)
    ref.call("clinic", 0, 2, jmt - 1, 1, imt)


@pytest.mark.parametrize("fourfil", [False, True])
def test_setvbc_and_clinic_bitwise(pkg, ref, fourfil):
    """§8 rows f-2 / f-4: the reference's setvbc (09/mom/setvbc.F) and clinic (09/mom/clinic.F, with filuv and the ice
    coupling hooks it calls) against ora_setvbc / ora_setvbc_mom / ora_clinic / ora_filuv on the same inputs."""
    from helpers import oracle_clinic, oracle_load_momentum

    case, o = setup_pair(pkg, ref, seed=13, fourfil=fourfil, momentum=True)
    imt, jmt, km, nt = case.imt, case.jmt, case.km, case.nt
    rng = np.random.default_rng(5)
    oracle_load_momentum(o, case)
    _ref_momentum_inputs(ref, o, case)
    # ---- tracer surface fluxes: one coupler slot per tracer (slots 3..nt+2), random fluxes, a bottom heat flux
    osbc = o.arr("sbc").reshape(-1, jmt, imt)
    rsbc = ref.view("sbc")
    flx = o.raw("sbc_flx_index")
    for n, nm in enumerate(case.tracer_names):
        slot = n + 3
        f = rng.standard_normal((jmt, imt)) * 1e-5
        osbc[slot - 1] = f
        rsbc[slot - 1] = f
        flx[n] = slot
        fname = "i" + nm[:-5] + "flx_phos" if nm.endswith("_phos") else "i" + nm + "flx"   # 09/common/csbc.h
        ref.set({"temp": "ihflx", "salt": "isflx"}.get(nm, fname), slot)
    bhf = rng.standard_normal((jmt, imt)) * 1e-6
    o.raw("bhf")[:] = bhf.ravel()
    ref.view("bhf")[...] = bhf
    o.call("ora_setvbc")
    if fourfil:
        for n in ("jfu0", "jfu1", "jfu2"):
            ref.set(n, int(oget(o, n)))
        ref.set("jskpu", int(oget(o, "jfu2")) - int(oget(o, "jfu1")))     # source/common/setcom.F:84
        ref.call("findex", ref.view("kmu"), 50, km, int(oget(o, "jfu1")), int(oget(o, "jfu2")), imt, ref.view("isuf"), ref.view("ieuf"))
    else:
        ref.set("jfu1", 0), ref.set("jfu2", jmt + 10)
    ref.set("first", 1)                                           # source/common/switch.h: filtr builds its tables when `first`
    oracle_clinic(o)
    _ref_clinic(ref, case)
    # setvbc: interior columns (the reference also sets the cyclic columns of stf/btf, which nothing reads)
    ostf, obtf = o.arr("stf", (nt, jmt, imt)), o.arr("btf", (nt, jmt, imt))
    assert np.array_equal(ostf[..., 1:-1], ref.view("stf")[..., 1:-1]) and np.abs(ostf).max() > 0
    assert np.array_equal(obtf[..., 1:-1], ref.view("btf")[..., 1:-1]) and np.abs(obtf[0]).max() > 0
    assert np.array_equal(o.arr("smf", (2, jmt, imt)), ref.view("smf"))
    assert np.array_equal(o.arr("bmf", (2, jmt, imt)), ref.view("bmf")) and np.abs(ref.view("bmf")).max() > 0
    # adv_vel on U cells, density, pressure gradient
    assert np.array_equal(o.arr("adv_veu", (jmt, km, imt))[1:-1], ref.view("adv_veu"))
    assert np.array_equal(o.arr("adv_vnu", (jmt, km, imt))[:-1], ref.view("adv_vnu"))
    assert np.array_equal(o.arr("adv_vbu", (jmt, km + 1, imt))[1:-1], ref.view("adv_vbu")[:, :km + 1])
    assert np.array_equal(o.arr("rho", (jmt, km, imt))[1:], ref.view("rho"))
    assert np.array_equal(o.arr("grad_p", (2, jmt, km, imt))[:, 1:-1], ref.view("grad_p"))
    # the step itself: u(tau+1) (internal mode) and the forcing of the barotropic equation
    up1, rup1 = o.arr("up1", (2, jmt, km, imt)), ref.view("u")[2]
    assert np.array_equal(up1[:, 1:-1], rup1[:, 1:-1]), np.abs(up1 - rup1).max()
    assert np.abs(up1).max() > 0
    assert np.array_equal(o.arr("zu", (2, jmt, imt))[:, 1:-1, 1:-1], ref.view("zu")[:, 1:-1, 1:-1])
    if fourfil:
        # the filter did act on the polar rows
        o2 = make_oracle(case, do_mobi=0)
        o2.set_scalar("do_filter", 0)
        oracle_load_momentum(o2, case)
        oracle_clinic(o2)
        assert np.abs(o2.arr("up1", (2, jmt, km, imt)) - up1).max() > 0
        o2.close()
    o.close()


def test_gasbc_flux_loop_bitwise(pkg, ref):
    """§8 row f-2: the air-sea gas exchange of 09/common/gasbc.F (the other caller of co2calc_SWS) against ora_gasbc.
    The reference routine is the atmosphere's coupling step: its forcing-data readers and the atmosphere's own physics
    are dropped by the translation (listed in the manifest), the flux loop, the land carbon fluxes, setbcx and areaavg run."""
    case, o = setup_pair(pkg, ref, seed=21)
    imt, jmt, nt = case.imt, case.jmt, case.nt
    rng = np.random.default_rng(8)
    order = ["isst", "isss", "issdic", "issalk", "issdic13", "issc14", "isso2", "iws", "inpp", "isr", "iburn", "idicflx",
             "idic13flx", "ic14flx", "io2flx"]
    # every slot index of csbc.h gets a valid slot (gasbc zeroes ~50 flux slots by name); the 15 the flux loop uses are 1..15
    q = 0
    for nm, ents in ref.man["commons"].items():
        if ents[0]["block"] == "csbc_i" and ents[0]["dims"] is None and nm.startswith("i") and nm not in order:
            ref.set(nm, 16 + q % 87)
            q += 1
    for s, nm in enumerate(order):
        ref.set(nm, s + 1)
    o.set("gas_idx", np.arange(1, 16, dtype=np.int32))
    osbc, rsbc = o.arr("sbc").reshape(-1, jmt, imt), ref.view("sbc")
    ocean = np.asarray(case["kmt"]) > 0
    vals = {"isst": rng.uniform(-3.0, 36.0, (jmt, imt)),                   # beyond the clamps at -2 and 35 (:152)
            "isss": rng.uniform(-0.006, 0.004, (jmt, imt)),                # (S-35)/1000
            "issdic": rng.uniform(1.8, 2.4, (jmt, imt)), "issalk": rng.uniform(2.2, 2.5, (jmt, imt)),
            "isso2": rng.uniform(0.0, 0.4, (jmt, imt)), "iws": rng.uniform(0.0, 2000.0, (jmt, imt)),
            "inpp": rng.uniform(0, 5e-8, (jmt, imt)), "isr": rng.uniform(0, 3e-8, (jmt, imt)), "iburn": rng.uniform(0, 1e-8, (jmt, imt))}
    vals["issdic13"] = vals["issdic"] * rng.uniform(0.004, 0.03, (jmt, imt))    # both clamps of r13dic are reached (:170-171)
    vals["issc14"] = vals["issdic"] * rng.uniform(0.8e-12, 1.3e-12, (jmt, imt))
    for s, nm in enumerate(order[:11]):
        osbc[s], rsbc[s] = vals[nm], vals[nm]
    ice = np.clip(rng.uniform(-0.5, 1.2, (jmt, imt)), 0.0, 1.0)
    o.arr("aice", (jmt, imt))[...] = ice
    ref.view("aice")[1] = ice
    ref.view("tmsk")[...] = ocean.astype(np.float64)
    for k, v in (("co2ccn", 283.0), ("dc13ccn", -6.5), ("dc14ccn", 12.0)):
        o.set_scalar(k, v), ref.set(k, v)
    o.call("ora_gasbc")
    ref.call("gasbc", 1, imt, 1, jmt)
    for s in range(11, 15):
        assert np.array_equal(osbc[s][1:-1], rsbc[s][1:-1]), order[s]
        assert np.abs(osbc[s][1:-1][ocean[1:-1]]).max() > 0
    assert np.isfinite(rsbc[11:15]).all()
    dropped = {d["why"] for d in ref.man["dropped"] if d["unit"] == "gasbc"}
    assert dropped <= {"forcing data reader", "atmosphere model (outside the path)", "I/O"}
    o.close()
