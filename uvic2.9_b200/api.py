"""Host-side mirror of the reference's tracer-step call sites, over the C ABI.

The reference's interface for this path is a set of Fortran subroutine calls inside `mom`
(source/mom/mom.F:340-389): ``call isopyc``, ``call vmixc``, ``call tracer``.  `TracerContext`
exposes the same three calls (plus `step`, which sequences them as `mom` does) on top of
libuvic_b200.so (include/uvic_b200.h).  There is no CPU fallback: if the CUDA library is
missing or no device is present, construction fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# UVIC_B200_LIB selects another build of the same library (A/B experiments: scripts/build_variants.py); the default is the
# in-tree product build
LIB_PATH = os.environ.get("UVIC_B200_LIB") or os.path.join(HERE, "libuvic_b200.so")

_c_double_p = C.POINTER(C.c_double)
_c_int_p = C.POINTER(C.c_int32)


class Dims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("imt", "jmt", "km", "nt", "nsrc", "jrow_lo", "jrow_hi")]


_GRID_FIELDS = (
    "dxt dxtr dxt2r dxt4r dxu dxur dyt dytr dyt2r dyt4r dyu dyur cst cstr csu csur cstdytr cstdyt2r csu_dyur "
    "dzt dztr dzt2r dztur dztlr zt zw dzw dzwr dtxcel dtxsqr dztxcl dzwxcl tlat duw due dus dun eosc to so"
).split()


class Grid(C.Structure):
    _fields_ = [(n, _c_double_p) for n in _GRID_FIELDS]


class Params(C.Structure):
    _fields_ = (
        [(n, C.c_double) for n in ("aidif", "kappa_h", "ahisop", "athkdf", "slmxr", "diff_cet", "diff_cnt",
                                   "zetar", "ogamma", "gravrho0r")]
        + [(n, C.c_int32) for n in ("fct", "isopycmix", "tidal_kv", "fullconvect", "mobi", "fourfil",
                                    "jfrst", "jft0", "jft1", "jft2")]
        + [("itrc", _c_int_p), ("mobi_index", _c_int_p), ("mobi_par", _c_double_p),
           ("n_mobi_index", C.c_int32), ("n_mobi_par", C.c_int32)]
    )


class Static(C.Structure):
    _fields_ = [("kmt", _c_int_p), ("mskhr", _c_int_p)] + [
        (n, _c_double_p) for n in ("fisop", "addisop", "edrm2", "edrs2", "edrk1", "edro1", "sg_bathy", "fe_hydr", "fe_atmdep")]


class StepInfo(C.Structure):
    _fields_ = [("dtts", C.c_double), ("leapfrog", C.c_int32), ("diag", C.c_int32), ("relyr", C.c_double),
                ("co2ccn", C.c_double)]


class GasbcPar(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("isst", "isss", "issdic", "issalk", "issdic13", "issc14", "isso2", "iws", "inpp", "isr",
                                         "iburn", "idicflx", "idic13flx", "ic14flx", "io2flx")] + \
               [("co2ccn", C.c_double), ("dc13ccn", C.c_double), ("dc14ccn", C.c_double)]


class ClinicStatic(C.Structure):
    _fields_ = [("kmu", _c_int_p)] + [(n, _c_double_p) for n in (
        "hr", "cori", "advmet", "am3", "am4", "dxmetr", "dxu2r", "dyu2r", "dyu4r", "csudyu2r", "visc_ceu", "amc_north",
        "amc_south")] + [(n, C.c_double) for n in ("kappa_m", "cdbot", "grav_rho0r")] + \
        [(n, C.c_int32) for n in ("fourfil", "jfrst", "jfu0", "jfu1", "jfu2")] + [(n, _c_double_p) for n in ("spsin", "spcos", "phi")]


# every symbol include/uvic_b200.h declares
ABI_SYMBOLS = [
    "uvic_b200_create", "uvic_b200_destroy", "uvic_b200_last_error", "uvic_b200_set_stream", "uvic_b200_synchronize",
    "uvic_b200_upload_t", "uvic_b200_download_t", "uvic_b200_download_tracer", "uvic_b200_upload_adv_vel",
    "uvic_b200_upload_u", "uvic_b200_adv_vel", "uvic_b200_upload_vbc", "uvic_b200_upload_forcing", "uvic_b200_rotate",
    "uvic_b200_isopyc", "uvic_b200_vmixc", "uvic_b200_tracer", "uvic_b200_step", "uvic_b200_tracer_step",
    "uvic_b200_inventory", "uvic_b200_tbar", "uvic_b200_sumbk", "uvic_b200_fetch", "uvic_b200_device_ptr",
    "uvic_b200_t_ptr", "uvic_b200_kernel_launches", "uvic_b200_local_rows", "uvic_b200_version",
    "uvic_b200_profile_enable", "uvic_b200_profile_count", "uvic_b200_profile_get", "uvic_b200_profile_reset",
    "uvic_b200_hint_next_step", "uvic_b200_pin_host", "uvic_b200_unpin_host",
    "uvic_b200_sbc_setup", "uvic_b200_upload_sbc", "uvic_b200_upload_sbc_slot", "uvic_b200_download_sbc",
    "uvic_b200_download_sbc_slot", "uvic_b200_setvbc", "uvic_b200_set_sbc", "uvic_b200_tracer_step_coupled",
    "uvic_b200_tavg_accumulate", "uvic_b200_tavg_fetch", "uvic_b200_state", "uvic_b200_gasbc", "uvic_b200_wait_before_advection",
    "uvic_b200_clinic_setup", "uvic_b200_upload_u_level", "uvic_b200_download_u", "uvic_b200_upload_smf", "uvic_b200_clinic",
    "uvic_b200_download_zu", "uvic_b200_rotate_u",
    "uvic_b200_travar_dtabs", "uvic_b200_set_host_window", "uvic_b200_lookahead_stats", "uvic_b200_invalidate_lookahead", "uvic_b200_join_streams", "uvic_b200_measure_fp64_peak",
    "uvic_b200_group_create", "uvic_b200_group_destroy", "uvic_b200_group_last_error", "uvic_b200_group_size", "uvic_b200_group_ctx",
    "uvic_b200_group_rows", "uvic_b200_group_upload_t", "uvic_b200_group_download_t", "uvic_b200_group_upload_adv_vel",
    "uvic_b200_group_upload_vbc", "uvic_b200_group_upload_forcing", "uvic_b200_group_step", "uvic_b200_group_rotate",
    "uvic_b200_group_inventory", "uvic_b200_group_synchronize",
]

_lib = None


def load_library():
    """dlopen libuvic_b200.so and declare the prototypes.  Fails loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
            "There is no CPU fallback for the tracer step.")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.uvic_b200_create.argtypes = [C.POINTER(Dims), C.POINTER(Grid), C.POINTER(Params), C.POINTER(Static), C.c_int, C.POINTER(vp)]
    L.uvic_b200_last_error.restype = C.c_char_p
    L.uvic_b200_last_error.argtypes = [vp]
    L.uvic_b200_version.restype = C.c_char_p
    for fn in ("destroy", "synchronize", "adv_vel", "rotate", "isopyc"):
        getattr(L, "uvic_b200_" + fn).argtypes = [vp]
    L.uvic_b200_set_stream.argtypes = [vp, vp]
    L.uvic_b200_upload_t.argtypes = [vp, C.c_int, vp]
    L.uvic_b200_download_t.argtypes = [vp, C.c_int, vp]
    L.uvic_b200_download_tracer.argtypes = [vp, C.c_int, C.c_int, vp]
    L.uvic_b200_upload_adv_vel.argtypes = [vp, vp, vp, vp]
    L.uvic_b200_upload_u.argtypes = [vp, vp]
    L.uvic_b200_upload_vbc.argtypes = [vp, vp, vp]
    L.uvic_b200_upload_forcing.argtypes = [vp, vp, vp, vp, vp]
    for fn in ("vmixc", "tracer", "step"):
        getattr(L, "uvic_b200_" + fn).argtypes = [vp, C.POINTER(StepInfo)]
    L.uvic_b200_tracer_step.argtypes = [vp, C.POINTER(StepInfo)] + [vp] * 8
    L.uvic_b200_hint_next_step.argtypes = [vp, C.POINTER(StepInfo)]
    L.uvic_b200_lookahead_stats.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.uvic_b200_invalidate_lookahead.argtypes = [vp]
    L.uvic_b200_set_host_window.argtypes = [vp, C.c_int]
    L.uvic_b200_join_streams.argtypes = [vp]
    L.uvic_b200_measure_fp64_peak.argtypes = [C.c_int] + [C.POINTER(C.c_double)] * 4
    L.uvic_b200_group_create.argtypes = [C.POINTER(Dims), C.POINTER(Grid), C.POINTER(Params), C.POINTER(Static), C.c_int, vp, C.POINTER(vp)]
    L.uvic_b200_group_last_error.restype = C.c_char_p
    L.uvic_b200_group_last_error.argtypes = [vp]
    L.uvic_b200_group_ctx.restype = vp
    L.uvic_b200_group_ctx.argtypes = [vp, C.c_int]
    L.uvic_b200_group_rows.argtypes = [vp, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    for fn in ("destroy", "size", "rotate", "synchronize"):
        getattr(L, "uvic_b200_group_" + fn).argtypes = [vp]
    L.uvic_b200_group_upload_t.argtypes = [vp, C.c_int, vp]
    L.uvic_b200_group_download_t.argtypes = [vp, C.c_int, vp]
    L.uvic_b200_group_upload_adv_vel.argtypes = [vp, vp, vp, vp]
    L.uvic_b200_group_upload_vbc.argtypes = [vp, vp, vp]
    L.uvic_b200_group_upload_forcing.argtypes = [vp, vp, vp, vp, vp]
    L.uvic_b200_group_step.argtypes = [vp, C.POINTER(StepInfo), C.POINTER(StepInfo)]
    L.uvic_b200_group_inventory.argtypes = [vp, C.c_int, vp]
    L.uvic_b200_pin_host.argtypes = [vp, C.c_size_t]
    L.uvic_b200_unpin_host.argtypes = [vp]
    L.uvic_b200_inventory.argtypes = [vp, C.c_int, vp]
    L.uvic_b200_tbar.argtypes = [vp, vp]
    L.uvic_b200_travar_dtabs.argtypes = [vp, vp, vp]
    L.uvic_b200_sumbk.argtypes = [vp, vp]
    L.uvic_b200_fetch.argtypes = [vp, C.c_char_p, vp, C.POINTER(C.c_size_t)]
    L.uvic_b200_device_ptr.restype = vp
    L.uvic_b200_device_ptr.argtypes = [vp, C.c_char_p, C.POINTER(C.c_size_t)]
    L.uvic_b200_t_ptr.restype = vp
    L.uvic_b200_t_ptr.argtypes = [vp, C.c_int]
    L.uvic_b200_kernel_launches.restype = C.c_int64
    L.uvic_b200_kernel_launches.argtypes = [vp]
    L.uvic_b200_local_rows.argtypes = [vp, _c_int_p, _c_int_p]
    L.uvic_b200_profile_enable.argtypes = [vp, C.c_int]
    L.uvic_b200_profile_count.argtypes = [vp]
    L.uvic_b200_profile_get.argtypes = [vp, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    L.uvic_b200_profile_reset.argtypes = [vp]
    L.uvic_b200_sbc_setup.argtypes = [vp, C.c_int, vp, vp]
    L.uvic_b200_upload_sbc.argtypes = [vp, vp, vp]
    L.uvic_b200_upload_sbc_slot.argtypes = [vp, C.c_int, vp]
    L.uvic_b200_download_sbc.argtypes = [vp, vp]
    L.uvic_b200_download_sbc_slot.argtypes = [vp, C.c_int, vp]
    L.uvic_b200_setvbc.argtypes = [vp]
    L.uvic_b200_set_sbc.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
    L.uvic_b200_wait_before_advection.argtypes = [vp, vp]
    L.uvic_b200_state.argtypes = [vp, C.c_int, vp]
    L.uvic_b200_gasbc.argtypes = [vp, C.POINTER(GasbcPar)]
    L.uvic_b200_tavg_accumulate.argtypes = [vp, vp, vp]
    L.uvic_b200_tavg_fetch.argtypes = [vp, vp, vp, _c_int_p, C.c_int]
    L.uvic_b200_tracer_step_coupled.argtypes = [vp, C.POINTER(StepInfo)] + [vp] * 5 + [C.c_int] * 4 + [vp, vp]
    L.uvic_b200_clinic_setup.argtypes = [vp, C.POINTER(ClinicStatic)]
    L.uvic_b200_upload_u_level.argtypes = [vp, C.c_int, vp]
    L.uvic_b200_download_u.argtypes = [vp, C.c_int, vp]
    L.uvic_b200_upload_smf.argtypes = [vp, vp]
    L.uvic_b200_clinic.argtypes = [vp, C.c_double, C.c_int, C.c_int]
    L.uvic_b200_download_zu.argtypes = [vp, vp]
    L.uvic_b200_rotate_u.argtypes = [vp]
    _lib = L
    return L


def measure_fp64_peak(device=0):
    """thread-level FP64 instruction rates of the device (dependent DFMA / DADD / DMUL chains, no memory traffic)"""
    L = load_library()
    a, b, c, clk = C.c_double(), C.c_double(), C.c_double(), C.c_double()
    if L.uvic_b200_measure_fp64_peak(int(device), C.byref(a), C.byref(b), C.byref(c), C.byref(clk)) != 0:
        raise RuntimeError("uvic_b200_measure_fp64_peak failed")
    return {"dfma_per_s": a.value, "dadd_per_s": b.value, "dmul_per_s": c.value, "sm_clock_mhz_nominal": clk.value,
            "tflops_fma": 2.0 * a.value / 1e12}


class UvicError(RuntimeError):
    pass


def _dp(a):
    return a.ctypes.data_as(_c_double_p)


def _vp(a):
    return None if a is None else C.c_void_p(a.ctypes.data if isinstance(a, np.ndarray) else int(a))


# j axis (counted from the front of the C-ordered numpy array) of every array with a j extent
_JAXIS = {
    "kmt": 0, "mskhr": 0, "tlat": 0, "fisop": 1, "sg_bathy": 1, "fe_hydr": 1, "fe_atmdep": 1,
    "addisop": 0, "edrm2": 0, "edrs2": 0, "edrk1": 0, "edro1": 0,
    "adv_vet": 0, "adv_vnt": 0, "adv_vbt": 0, "stf": 1, "btf": 1, "u": 1, "t": 2,
    "dnswr": 0, "aice": 0, "hice": 0, "hsno": 0,
    "kmu": 0, "hr": 0, "cori": 1, "visc_ceu": 0, "amc_north": 0, "amc_south": 0, "um1": 1, "taux": 0, "tauy": 0,
    "umask": 0, "tmask": 0,
}


def slab_rows(jmt, jlo, jhi):
    jbase = max(1, jlo - 2)
    jtop = min(jmt, jhi + 2)
    return jbase, jtop - jbase + 1


def slab_slice(name, arr, jbase, jl, case=None):
    """Rows jbase..jbase+jl-1 (global, 1-based) of a global array.  A lazily stacked weak-scaling case
    (synthetic.stack_bands(..., lazy=True)) keeps its 3-D arrays at the size of one band: their rows are gathered through
    case.row_map instead of being sliced from a materialised global array."""
    ax = _JAXIS[name]
    if case is not None and name in getattr(case, "lazy", ()):
        rows = case.row_map[jbase - 1:jbase - 1 + jl]
        return np.ascontiguousarray(np.take(np.asarray(arr), rows, axis=ax))
    sl = [slice(None)] * arr.ndim
    sl[ax] = slice(jbase - 1, jbase - 1 + jl)
    return np.ascontiguousarray(arr[tuple(sl)])


class TracerContext:
    """Device-resident tracer step for one latitude slab (rows jlo..jhi of the global grid)."""

    def __init__(self, case, jlo=None, jhi=None, device=0, fct=1, isopycmix=1, tidal_kv=1, fullconvect=1, mobi=0,
                 fourfil=0):
        self.L = load_library()
        self.case = case
        imt, jmt, km, nt, nsrc = case.imt, case.jmt, case.km, case.nt, case.nsrc
        self.jlo = 2 if jlo is None else jlo
        self.jhi = jmt - 1 if jhi is None else jhi
        self.jbase, self.jl = slab_rows(jmt, self.jlo, self.jhi)
        self.imt, self.jmt, self.km, self.nt, self.nsrc = imt, jmt, km, nt, nsrc
        a, s = case.arrays, case.scalars
        self._keep = []

        def f64(x):
            x = np.ascontiguousarray(x, dtype=np.float64)
            self._keep.append(x)
            return x

        def loc(name):
            return f64(slab_slice(name, a[name], self.jbase, self.jl, case))

        d = Dims(imt, jmt, km, nt, max(nsrc, 0), self.jlo, self.jhi)
        g = Grid()
        for n in _GRID_FIELDS:
            arr = loc(n) if n == "tlat" else f64(a[n])
            setattr(g, n, _dp(arr))
        p = Params()
        for n in ("aidif", "kappa_h", "ahisop", "athkdf", "slmxr", "diff_cet", "diff_cnt", "zetar", "ogamma", "gravrho0r"):
            setattr(p, n, float(s[n]))
        p.fct, p.isopycmix, p.tidal_kv, p.fullconvect, p.mobi, p.fourfil = fct, isopycmix, tidal_kv, fullconvect, mobi, fourfil
        if fourfil:
            p.jfrst, p.jft0, p.jft1, p.jft2 = (int(s[n]) for n in ("jfrst", "jft0", "jft1", "jft2"))
        itrc = np.ascontiguousarray(a["itrc"], dtype=np.int32)
        self._keep.append(itrc)
        p.itrc = itrc.ctypes.data_as(_c_int_p)
        if mobi:
            if not getattr(case, "has_mobi", False):
                raise UvicError("O_mobi needs the full MOBI tracer set (see mobi_params.mobi_index_maps)")
            midx = np.ascontiguousarray(a["mobi_idx"], dtype=np.int32)
            mpar = f64(a["mobi_par"])
            self._keep.append(midx)
            p.mobi_index = midx.ctypes.data_as(_c_int_p)
            p.mobi_par = _dp(mpar)
            p.n_mobi_index, p.n_mobi_par = midx.size, mpar.size
        self.mobi = bool(mobi)
        st = Static()
        kmt = np.ascontiguousarray(slab_slice("kmt", a["kmt"], self.jbase, self.jl), dtype=np.int32)
        msk = np.ascontiguousarray(slab_slice("mskhr", a["mskhr"], self.jbase, self.jl), dtype=np.int32)
        self._keep += [kmt, msk]
        st.kmt = kmt.ctypes.data_as(_c_int_p)
        st.mskhr = msk.ctypes.data_as(_c_int_p)
        for n in ("fisop", "addisop", "edrm2", "edrs2", "edrk1", "edro1", "sg_bathy", "fe_hydr", "fe_atmdep"):
            if n in a:
                setattr(st, n, _dp(loc(n)))
        h = C.c_void_p()
        rc = self.L.uvic_b200_create(C.byref(d), C.byref(g), C.byref(p), C.byref(st), device, C.byref(h))
        if rc != 0:
            raise UvicError(self.L.uvic_b200_last_error(None).decode())
        self.h = h
        self.dtts = float(s["dtts"])
        self.relyr = float(s.get("relyr", 0.0))
        self.co2ccn = float(s.get("co2ccn", 280.0))
        self.relyr_next = self.co2ccn_next = None

    # ---- plumbing ----------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise UvicError(self.L.uvic_b200_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.uvic_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_handle):
        self._ck(self.L.uvic_b200_set_stream(self.h, C.c_void_p(int(cuda_stream_handle))))

    def synchronize(self):
        self._ck(self.L.uvic_b200_synchronize(self.h))

    def stepinfo(self, leapfrog=True, diag=False):
        return StepInfo(self.dtts, 1 if leapfrog else 0, 1 if diag else 0, self.relyr, self.co2ccn)

    # ---- state movement ------------------------------------------------------------
    def shape_t(self):
        return (self.nt, self.jl, self.km, self.imt)

    def upload_t(self, level, t_local):
        t_local = np.ascontiguousarray(t_local, dtype=np.float64)
        assert t_local.shape == self.shape_t(), (t_local.shape, self.shape_t())
        self._ck(self.L.uvic_b200_upload_t(self.h, level, _vp(t_local)))
        self.synchronize()

    def download_t(self, level, out=None):
        out = np.empty(self.shape_t()) if out is None else out
        self._ck(self.L.uvic_b200_download_t(self.h, level, _vp(out)))
        return out

    def load_state(self, case=None):
        """Upload t(tau-1), t(tau), velocities and vertical b.c. of a (global) Case."""
        case = case or self.case
        a = case.arrays
        sl = lambda n, x: slab_slice(n, x, self.jbase, self.jl, case)
        self.upload_t(-1, sl("t", a["t"])[0])
        self.upload_t(0, sl("t", a["t"])[1])
        vet, vnt, vbt = (np.ascontiguousarray(sl(n, a[n])) for n in ("adv_vet", "adv_vnt", "adv_vbt"))
        self._ck(self.L.uvic_b200_upload_adv_vel(self.h, _vp(vet), _vp(vnt), _vp(vbt)))
        stf, btf = np.ascontiguousarray(sl("stf", a["stf"])), np.ascontiguousarray(sl("btf", a["btf"]))
        self._ck(self.L.uvic_b200_upload_vbc(self.h, _vp(stf), _vp(btf)))
        if self.mobi:
            f = [np.ascontiguousarray(sl(n, a[n]), dtype=np.float64) for n in ("dnswr", "aice", "hice", "hsno")]
            self._ck(self.L.uvic_b200_upload_forcing(self.h, *[_vp(x) for x in f]))
        self.synchronize()

    def upload_u(self, u_local):
        u_local = np.ascontiguousarray(u_local, dtype=np.float64)
        self._ck(self.L.uvic_b200_upload_u(self.h, _vp(u_local)))
        self.synchronize()

    def adv_vel(self):
        self._ck(self.L.uvic_b200_adv_vel(self.h))

    def state(self, level=0):
        """rho (jl, km, imt) of a time level: source/mom/state.F."""
        out = np.empty(self.shape3())
        self._ck(self.L.uvic_b200_state(self.h, level, _vp(out)))
        return out

    def rotate(self):
        self._ck(self.L.uvic_b200_rotate(self.h))

    def wait_before_advection(self, cuda_event_handle):
        """The next step's first advection kernel waits for this cudaEvent_t (end of the halo exchange)."""
        self._ck(self.L.uvic_b200_wait_before_advection(self.h, C.c_void_p(int(cuda_event_handle))))

    # ---- surface boundary conditions on the device (09/mom/setvbc.F, 09/mom/set_sbc.F) ----
    def sbc_setup(self, numsbc, flx_index, acc_index):
        """flx_index[n] / acc_index[n]: 1-based sbc slot of tracer n's surface flux / accumulator (0 = none)."""
        self.numsbc = int(numsbc)
        f = np.ascontiguousarray(flx_index, dtype=np.int32)
        a = np.ascontiguousarray(acc_index, dtype=np.int32)
        assert f.shape == (self.nt,) and a.shape == (self.nt,)
        self._ck(self.L.uvic_b200_sbc_setup(self.h, self.numsbc, _vp(f), _vp(a)))

    def upload_sbc(self, sbc_local=None, bhf_local=None):
        """sbc_local: (numsbc, jl, imt) slab of the coupler's array; bhf_local: (jl, imt)."""
        if sbc_local is not None:
            sbc_local = np.ascontiguousarray(sbc_local, dtype=np.float64)
            assert sbc_local.shape == (self.numsbc, self.jl, self.imt)
        if bhf_local is not None:
            bhf_local = np.ascontiguousarray(bhf_local, dtype=np.float64)
            assert bhf_local.shape == (self.jl, self.imt)
        self._ck(self.L.uvic_b200_upload_sbc(self.h, _vp(sbc_local), _vp(bhf_local)))
        self.synchronize()

    def download_sbc(self):
        out = np.empty((self.numsbc, self.jl, self.imt))
        self._ck(self.L.uvic_b200_download_sbc(self.h, _vp(out)))
        return out

    def gasbc(self, slots, co2ccn, dc13ccn, dc14ccn):
        """slots: dict of the 15 sbc slot indices (1-based) of GasbcPar; 09/common/gasbc.F flux loop on the device."""
        gp = GasbcPar(co2ccn=co2ccn, dc13ccn=dc13ccn, dc14ccn=dc14ccn, **{k: int(v) for k, v in slots.items()})
        self._ck(self.L.uvic_b200_gasbc(self.h, C.byref(gp)))

    def setvbc(self):
        self._ck(self.L.uvic_b200_setvbc(self.h))

    def tracer_step_coupled(self, adv_vet, adv_vnt, adv_vbt, sbc_in, bhf, eots, osegs, osege, ntspos, ts_taup1, sbc_out,
                            leapfrog=True, next_leapfrog=None):
        """uvic_b200_tracer_step_coupled with host (ideally pinned) numpy buffers; None = NULL."""
        if next_leapfrog is not None:
            self.hint_next_step(next_leapfrog)
        si = self.stepinfo(leapfrog)
        self._ck(self.L.uvic_b200_tracer_step_coupled(self.h, C.byref(si), _vp(adv_vet), _vp(adv_vnt), _vp(adv_vbt), _vp(sbc_in),
                                                      _vp(bhf), int(eots), int(osegs), int(osege), int(ntspos), _vp(ts_taup1),
                                                      _vp(sbc_out)))

    # ---- baroclinic momentum step (09/mom/clinic.F) ----
    def clinic_setup(self, case=None, fourfil=False):
        """Time-invariant inputs of clinic from a Case prepared by synthetic.add_momentum (or any dict-like with the same
        arrays / scalars)."""
        case = case or self.case
        a, s = case.arrays, case.scalars
        cs = ClinicStatic()
        keep = []
        kmu = np.ascontiguousarray(slab_slice("kmu", a["kmu"], self.jbase, self.jl), dtype=np.int32)
        keep.append(kmu)
        cs.kmu = kmu.ctypes.data_as(_c_int_p)
        for n in ("hr", "cori", "visc_ceu", "amc_north", "amc_south"):
            x = np.ascontiguousarray(slab_slice(n, a[n], self.jbase, self.jl, case), dtype=np.float64)
            keep.append(x)
            setattr(cs, n, _dp(x))
        for n in ("advmet", "am3", "am4", "dxmetr", "dxu2r", "dyu2r", "dyu4r", "csudyu2r"):
            x = np.ascontiguousarray(a[n], dtype=np.float64)
            keep.append(x)
            setattr(cs, n, _dp(x))
        cs.kappa_m, cs.cdbot, cs.grav_rho0r = float(s["kappa_m"]), float(s["cdbot"]), float(s["grav_rho0r"])
        if fourfil:
            cs.fourfil = 1
            cs.jfrst, cs.jfu0, cs.jfu1, cs.jfu2 = (int(s[n]) for n in ("jfrst", "jfu0", "jfu1", "jfu2"))
            for n in ("spsin", "spcos", "phi"):
                x = np.ascontiguousarray(a[n], dtype=np.float64)
                keep.append(x)
                setattr(cs, n, _dp(x))
        self._ck(self.L.uvic_b200_clinic_setup(self.h, C.byref(cs)))

    def shape_u(self):
        return (2, self.jl, self.km, self.imt)

    def upload_u_level(self, level, u_local):
        u_local = np.ascontiguousarray(u_local, dtype=np.float64)
        assert u_local.shape == self.shape_u(), (u_local.shape, self.shape_u())
        self._ck(self.L.uvic_b200_upload_u_level(self.h, int(level), _vp(u_local)))
        self.synchronize()

    def download_u(self, level):
        out = np.empty(self.shape_u())
        self._ck(self.L.uvic_b200_download_u(self.h, int(level), _vp(out)))
        return out

    def upload_smf(self, smf_local):
        smf_local = np.ascontiguousarray(smf_local, dtype=np.float64)
        assert smf_local.shape == (2, self.jl, self.imt)
        self._ck(self.L.uvic_b200_upload_smf(self.h, _vp(smf_local)))
        self.synchronize()

    def clinic(self, c2dtuv, itaux=0, itauy=0):
        self._ck(self.L.uvic_b200_clinic(self.h, float(c2dtuv), int(itaux), int(itauy)))

    def download_zu(self):
        out = np.empty((2, self.jl, self.imt))
        self._ck(self.L.uvic_b200_download_zu(self.h, _vp(out)))
        return out

    def rotate_u(self):
        self._ck(self.L.uvic_b200_rotate_u(self.h))

    # ---- time averages (09/mom/timeavgs.F avgvar / avgout, tracer part) ----
    def tavg_accumulate(self, vflux_local=None, gaost=None):
        if vflux_local is not None:
            vflux_local = np.ascontiguousarray(vflux_local, dtype=np.float64)
            assert vflux_local.shape == (self.jl, self.imt)
        if gaost is not None:
            gaost = np.ascontiguousarray(gaost, dtype=np.float64)
            assert gaost.shape == (self.nt,)
        self._ck(self.L.uvic_b200_tavg_accumulate(self.h, _vp(vflux_local), _vp(gaost)))

    def tavg_fetch(self, reset=True):
        avg_t = np.empty(self.shape_t())
        avg_stf = np.empty((self.nt, self.jl, self.imt))
        n = C.c_int(0)
        self._ck(self.L.uvic_b200_tavg_fetch(self.h, _vp(avg_t), _vp(avg_stf), C.byref(n), int(reset)))
        return avg_t, avg_stf, n.value

    def set_sbc(self, eots=True, osegs=False, osege=False, ntspos=1):
        self._ck(self.L.uvic_b200_set_sbc(self.h, int(eots), int(osegs), int(osege), int(ntspos)))

    # ---- the reference's call sites ---------------------------------------------------
    def isopyc(self):
        self._ck(self.L.uvic_b200_isopyc(self.h))

    def vmixc(self, leapfrog=True):
        si = self.stepinfo(leapfrog)
        self._ck(self.L.uvic_b200_vmixc(self.h, C.byref(si)))

    def tracer(self, leapfrog=True, diag=False):
        si = self.stepinfo(leapfrog, diag)
        self._ck(self.L.uvic_b200_tracer(self.h, C.byref(si)))

    def hint_next_step(self, leapfrog=True):
        """Tell the library what the step AFTER the next call looks like (MOBI look-ahead): its leapfrog flag and ITS model
        time / CO2 (`set_time(relyr, relyr_next)`; a driver advances relyr every step, source/mom/mom.F)."""
        si = self.stepinfo(leapfrog)
        si.relyr = self.relyr_next if self.relyr_next is not None else self.relyr
        si.co2ccn = self.co2ccn_next if self.co2ccn_next is not None else self.co2ccn
        self._ck(self.L.uvic_b200_hint_next_step(self.h, C.byref(si)))

    def set_time(self, relyr, relyr_next=None, co2ccn=None, co2ccn_next=None):
        """Model time (years) and atmospheric CO2 of the next step call, and of the step after it (for the look-ahead hint)."""
        self.relyr, self.relyr_next = float(relyr), (None if relyr_next is None else float(relyr_next))
        if co2ccn is not None:
            self.co2ccn = float(co2ccn)
        self.co2ccn_next = None if co2ccn_next is None else float(co2ccn_next)

    def set_host_window(self, jrow_first):
        """host velocity arrays start at global row `jrow_first` (2 for the reference's (imt,km,jsmw:jmw) COMMON arrays)"""
        self._ck(self.L.uvic_b200_set_host_window(self.h, int(jrow_first)))

    def lookahead_stats(self):
        h, m = C.c_int64(), C.c_int64()
        self._ck(self.L.uvic_b200_lookahead_stats(self.h, C.byref(h), C.byref(m)))
        return h.value, m.value

    def invalidate_lookahead(self):
        self._ck(self.L.uvic_b200_invalidate_lookahead(self.h))

    def join_side_streams(self):
        self._ck(self.L.uvic_b200_join_streams(self.h))

    def step(self, leapfrog=True, diag=False, next_leapfrog=None):
        if next_leapfrog is not None:
            self.hint_next_step(next_leapfrog)
        si = self.stepinfo(leapfrog, diag)
        self._ck(self.L.uvic_b200_step(self.h, C.byref(si)))

    def tracer_step_host(self, t_taum1, t_tau, adv_vet, adv_vnt, adv_vbt, stf, btf, t_taup1, leapfrog=True, next_leapfrog=None):
        """One synchronous step with host buffers (the call the Fortran shim makes)."""
        if next_leapfrog is not None:
            self.hint_next_step(next_leapfrog)
        si = self.stepinfo(leapfrog)
        self._ck(self.L.uvic_b200_tracer_step(self.h, C.byref(si), _vp(t_taum1), _vp(t_tau), _vp(adv_vet), _vp(adv_vnt),
                                              _vp(adv_vbt), _vp(stf), _vp(btf), _vp(t_taup1)))

    # ---- diagnostics / introspection ----------------------------------------------------
    def inventory(self, level):
        out = np.empty(self.nt)
        self._ck(self.L.uvic_b200_inventory(self.h, level, _vp(out)))
        return out

    def tbar(self):
        out = np.empty((self.jhi - self.jlo + 1, self.nt, self.km))
        self._ck(self.L.uvic_b200_tbar(self.h, _vp(out)))
        return out

    def travar_dtabs(self):
        shp = (self.jhi - self.jlo + 1, self.nt, self.km)
        a, b = np.empty(shp), np.empty(shp)
        self._ck(self.L.uvic_b200_travar_dtabs(self.h, _vp(a), _vp(b)))
        return a, b

    def sumbk(self):
        out = np.empty((self.nt, self.km, 3))
        self._ck(self.L.uvic_b200_sumbk(self.h, _vp(out)))
        return out

    def fetch(self, name, shape=None):
        n = C.c_size_t()
        self._ck(self.L.uvic_b200_fetch(self.h, name.encode(), None, C.byref(n)))
        out = np.empty(n.value)
        self._ck(self.L.uvic_b200_fetch(self.h, name.encode(), _vp(out), C.byref(n)))
        return out.reshape(shape) if shape is not None else out

    def t_ptr(self, level):
        return self.L.uvic_b200_t_ptr(self.h, level)

    def device_ptr(self, name):
        n = C.c_size_t()
        p = self.L.uvic_b200_device_ptr(self.h, name.encode(), C.byref(n))
        return p, n.value

    def profile_enable(self, on=True):
        self._ck(self.L.uvic_b200_profile_enable(self.h, 1 if on else 0))

    def profile_reset(self):
        self._ck(self.L.uvic_b200_profile_reset(self.h))

    def profile(self):
        """{kernel name: (total ms, launches)} accumulated since the last reset (CUDA events)."""
        n = self.L.uvic_b200_profile_count(self.h)
        out = {}
        for q in range(max(n, 0)):
            name = C.create_string_buffer(64)
            ms, cnt = C.c_double(), C.c_int64()
            self._ck(self.L.uvic_b200_profile_get(self.h, q, name, 64, C.byref(ms), C.byref(cnt)))
            out[name.value.decode()] = (ms.value, cnt.value)
        return out

    @property
    def kernel_launches(self):
        return int(self.L.uvic_b200_kernel_launches(self.h))

    def shape3(self):
        return (self.jl, self.km, self.imt)

    def shape3z(self):
        return (self.jl, self.km + 1, self.imt)


class TracerGroup:
    """Several devices driven by one host thread (uvic_b200_group_*): the mode the serial Fortran host uses.  Takes the GLOBAL
    arrays of a Case; the library cuts the latitude slabs, exchanges the halos with peer copies and sums the inventories."""

    def __init__(self, case, devices, fct=1, isopycmix=1, tidal_kv=1, fullconvect=1, mobi=0, fourfil=0):
        self.L = load_library()
        self.case = case
        imt, jmt, km, nt, nsrc = case.imt, case.jmt, case.km, case.nt, case.nsrc
        self.imt, self.jmt, self.km, self.nt = imt, jmt, km, nt
        a, s = case.arrays, case.scalars
        self._keep = []

        def f64(x):
            x = np.ascontiguousarray(x, dtype=np.float64)
            self._keep.append(x)
            return x

        d = Dims(imt, jmt, km, nt, max(nsrc, 0), 2, jmt - 1)
        g = Grid()
        for n in _GRID_FIELDS:
            setattr(g, n, _dp(f64(a[n])))
        p = Params()
        for n in ("aidif", "kappa_h", "ahisop", "athkdf", "slmxr", "diff_cet", "diff_cnt", "zetar", "ogamma", "gravrho0r"):
            setattr(p, n, float(s[n]))
        p.fct, p.isopycmix, p.tidal_kv, p.fullconvect, p.mobi, p.fourfil = fct, isopycmix, tidal_kv, fullconvect, mobi, fourfil
        if fourfil:
            p.jfrst, p.jft0, p.jft1, p.jft2 = (int(s[n]) for n in ("jfrst", "jft0", "jft1", "jft2"))
        itrc = np.ascontiguousarray(a["itrc"], dtype=np.int32)
        self._keep.append(itrc)
        p.itrc = itrc.ctypes.data_as(_c_int_p)
        if mobi:
            midx = np.ascontiguousarray(a["mobi_idx"], dtype=np.int32)
            mpar = f64(a["mobi_par"])
            self._keep.append(midx)
            p.mobi_index = midx.ctypes.data_as(_c_int_p)
            p.mobi_par = _dp(mpar)
            p.n_mobi_index, p.n_mobi_par = midx.size, mpar.size
        self.mobi = bool(mobi)
        st = Static()
        kmt = np.ascontiguousarray(a["kmt"], dtype=np.int32)
        msk = np.ascontiguousarray(a["mskhr"], dtype=np.int32)
        self._keep += [kmt, msk]
        st.kmt = kmt.ctypes.data_as(_c_int_p)
        st.mskhr = msk.ctypes.data_as(_c_int_p)
        for n in ("fisop", "addisop", "edrm2", "edrs2", "edrk1", "edro1", "sg_bathy", "fe_hydr", "fe_atmdep"):
            if n in a:
                setattr(st, n, _dp(f64(a[n])))
        devs = np.ascontiguousarray(devices, dtype=np.int32)
        h = C.c_void_p()
        rc = self.L.uvic_b200_group_create(C.byref(d), C.byref(g), C.byref(p), C.byref(st), len(devs), devs.ctypes.data, C.byref(h))
        if rc != 0:
            raise UvicError(self.L.uvic_b200_group_last_error(None).decode())
        self.h = h
        self.n = len(devs)
        self.dtts = float(s["dtts"])
        self.relyr = float(s.get("relyr", 0.0))
        self.co2ccn = float(s.get("co2ccn", 280.0))

    def _ck(self, rc):
        if rc != 0:
            raise UvicError(self.L.uvic_b200_group_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.uvic_b200_group_destroy(self.h)
            self.h = None

    def rows(self, r):
        lo, hi = C.c_int32(), C.c_int32()
        self._ck(self.L.uvic_b200_group_rows(self.h, r, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def load_state(self):
        a = self.case.arrays
        t = np.ascontiguousarray(a["t"], dtype=np.float64)
        self._ck(self.L.uvic_b200_group_upload_t(self.h, -1, _vp(t[0])))
        self._ck(self.L.uvic_b200_group_upload_t(self.h, 0, _vp(t[1])))
        vet, vnt, vbt = (np.ascontiguousarray(a[n], dtype=np.float64) for n in ("adv_vet", "adv_vnt", "adv_vbt"))
        self._ck(self.L.uvic_b200_group_upload_adv_vel(self.h, _vp(vet), _vp(vnt), _vp(vbt)))
        stf, btf = np.ascontiguousarray(a["stf"], dtype=np.float64), np.ascontiguousarray(a["btf"], dtype=np.float64)
        self._ck(self.L.uvic_b200_group_upload_vbc(self.h, _vp(stf), _vp(btf)))
        if self.mobi:
            f = [np.ascontiguousarray(a[n], dtype=np.float64) for n in ("dnswr", "aice", "hice", "hsno")]
            self._ck(self.L.uvic_b200_group_upload_forcing(self.h, *[_vp(x) for x in f]))
        self._ck(self.L.uvic_b200_group_synchronize(self.h))

    def step(self, leapfrog=True, next_leapfrog=None):
        si = StepInfo(self.dtts, 1 if leapfrog else 0, 0, self.relyr, self.co2ccn)
        nx = None
        if next_leapfrog is not None:
            nx = StepInfo(self.dtts, 1 if next_leapfrog else 0, 0, self.relyr, self.co2ccn)
        self._ck(self.L.uvic_b200_group_step(self.h, C.byref(si), C.byref(nx) if nx is not None else None))

    def rotate(self):
        self._ck(self.L.uvic_b200_group_rotate(self.h))

    def download_t(self, level):
        out = np.zeros((self.nt, self.jmt, self.km, self.imt))
        self._ck(self.L.uvic_b200_group_download_t(self.h, level, _vp(out)))
        return out

    def inventory(self, level):
        out = np.empty(self.nt)
        self._ck(self.L.uvic_b200_group_inventory(self.h, level, _vp(out)))
        return out
