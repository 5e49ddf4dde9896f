"""The C ABI exercised from C: tests/c_abi_harness.c is compiled with gcc against include/uvic_b200.h and linked with
libuvic_b200.so, keeps its arrays in static storage the way the reference's COMMON blocks do, and drives create ->
tracer_step_coupled x 4 -> download (and the uvic_b200_group_* interface).  Its result must equal the Python path's bit for
bit -- a check through the header, not through the prototypes api.py retypes."""
import os
import subprocess

import numpy as np
import pytest

from conftest import load_pkg

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NSTEP, NTSPOS = 4, 4


def _lf(itt):
    return (itt % 3) != 0


@pytest.fixture(scope="module")
def pkg():
    return load_pkg()


@pytest.fixture(scope="module")
def built(pkg, tmp_path_factory):
    from uvic29_b200 import mobi_params as mp

    case = pkg.synthetic.make_case(imt=34, jmt=26, km=8, nt=37, seed=31)
    tmp = tmp_path_factory.mktemp("cabi")
    exe = str(tmp / "c_abi_harness")
    libdir = os.path.dirname(pkg.api.LIB_PATH)
    cmd = ["gcc", "-O1", "-std=gnu11", "-mcmodel=medium", f"-DIMT={case.imt}", f"-DJMT={case.jmt}", f"-DKM={case.km}", f"-DNT={case.nt}",
           f"-DNSRC={case.nsrc}", f"-DNMOBIPAR={mp.N_PAR}", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c_abi_harness.c"),
           "-o", exe, "-L", libdir, "-luvic_b200", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    a, s = case.arrays, case.scalars
    inp = str(tmp / "in.bin")
    with open(inp, "wb") as f:
        def w(x, dt=np.float64):
            f.write(np.ascontiguousarray(x, dtype=dt).tobytes())
        for n in pkg.api._GRID_FIELDS:
            w(a[n])
        w([s[n] for n in ("aidif", "kappa_h", "ahisop", "athkdf", "slmxr", "diff_cet", "diff_cnt", "zetar", "ogamma", "gravrho0r")])
        w(a["itrc"], np.int32), w(a["mobi_idx"], np.int32), w(a["mobi_par"]), w(a["kmt"], np.int32), w(a["mskhr"], np.int32)
        for n in ("fisop", "addisop", "edrm2", "edrs2", "edrk1", "edro1", "sg_bathy", "fe_hydr", "fe_atmdep"):
            w(a[n])
        w(a["t"][0]), w(a["t"][1])
        for n in ("adv_vet", "adv_vnt", "adv_vbt", "dnswr", "aice", "hice", "hsno"):
            w(a[n])
        w([s["dtts"], s["relyr"], s["co2ccn"]])
    return case, exe, inp, tmp


def test_c_harness_coupled_steps_equal_python_path(pkg, built):
    case, exe, inp, tmp = built
    out = str(tmp / "out.bin")
    r = subprocess.run([exe, inp, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    n3 = case.imt * case.km * case.jmt
    raw = np.fromfile(out)
    t_c = raw[: n3 * case.nt].reshape(case.nt, case.jmt, case.km, case.imt)
    ts_c = raw[n3 * case.nt: n3 * (case.nt + 2)].reshape(2, case.jmt, case.km, case.imt)
    # the same calls through the Python binding
    ctx = pkg.TracerContext(case, mobi=1)
    ctx.load_state()
    nsbc = 2 * case.nt + 4
    ctx.sbc_setup(nsbc, np.arange(1, case.nt + 1, dtype=np.int32), np.arange(case.nt + 1, 2 * case.nt + 1, dtype=np.int32))
    sbc, bhf = np.zeros((nsbc, case.jmt, case.imt)), np.zeros((case.jmt, case.imt))
    sbc_out, ts = np.zeros_like(sbc), np.zeros((2, case.jmt, case.km, case.imt))
    vet, vnt, vbt = (np.ascontiguousarray(case[n]) for n in ("adv_vet", "adv_vnt", "adv_vbt"))
    for itt in range(1, NSTEP + 1):
        p = (itt - 1) % NTSPOS
        ctx.tracer_step_coupled(vet, vnt, vbt, sbc if p == 0 else None, bhf if p == 0 else None, True, p == 0, p == NTSPOS - 1, NTSPOS,
                                ts, sbc_out, leapfrog=_lf(itt), next_leapfrog=_lf(itt + 1))
        ctx.rotate()
    t_py = ctx.download_t(0)
    assert np.isfinite(t_c).all() and np.abs(t_c[8]).max() > 0
    assert np.array_equal(t_c, t_py), np.abs(t_c - t_py).max()
    assert np.array_equal(ts_c, ts)
    ctx.close()


def test_c_harness_group_mode_equals_single_context(pkg, built):
    case, exe, inp, tmp = built
    out = str(tmp / "out_group.bin")
    r = subprocess.run([exe, inp, out, "group", "3"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    t_c = np.fromfile(out).reshape(case.nt, case.jmt, case.km, case.imt)
    ctx = pkg.TracerContext(case, mobi=1)
    ctx.load_state()
    for itt in range(1, NSTEP + 1):
        ctx.step(leapfrog=_lf(itt), next_leapfrog=_lf(itt + 1))
        ctx.rotate()
    t_py = ctx.download_t(0)
    assert np.array_equal(t_c[:, 1:-1], t_py[:, 1:-1]), np.abs(t_c - t_py).max()
    ctx.close()
