"""ctypes access to the CPU oracle (oracle/_build/liboracle.so).  Test infrastructure only."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
# UVIC_ORACLE_VARIANT=o3 selects the -O3 timing build (bench.py's CPU baseline); the parity tests use the default,
# which is compiled without FMA contraction
LIB = os.path.join(ORACLE_DIR, "_build", "liboracle_o3.so" if os.environ.get("UVIC_ORACLE_VARIANT") == "o3" else "liboracle.so")


def build_oracle(force=False):
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        L = ctypes.CDLL(LIB)
        L.ora_create.restype = ctypes.c_void_p
        L.ora_create.argtypes = [ctypes.c_int] * 5
        L.ora_destroy.argtypes = [ctypes.c_void_p]
        L.ora_array.restype = ctypes.c_void_p
        L.ora_array.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_int)]
        L.ora_set_scalar.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_double]
        L.ora_get_scalar.restype = ctypes.c_double
        L.ora_get_scalar.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        for fn in ("ora_make_masks", "ora_adv_vel", "ora_isopyc", "ora_vmixc", "ora_tracer", "ora_step",
                   "ora_mobi_columns", "ora_filt", "ora_setvbc", "ora_set_sbc", "ora_avgvar", "ora_avgout", "ora_state", "ora_gasbc",
                   "ora_adv_vel_u", "ora_setvbc_mom", "ora_clinic"):
            getattr(L, fn).argtypes = [ctypes.c_void_p]
            getattr(L, fn).restype = None
        for fn in ("ora_adv_flux", "ora_isoflux", "ora_diag_tbar"):
            getattr(L, fn).argtypes = [ctypes.c_void_p, ctypes.c_int]
            getattr(L, fn).restype = None
        _lib = L
    return _lib


class Oracle:
    """One oracle context; arrays are numpy views onto the C storage (Fortran layouts,
    dims reversed: t[3, nt, jmt, km, imt])."""

    def __init__(self, imt, jmt, km, nt, nsrc):
        self.L = lib()
        self.dims = (imt, jmt, km, nt, nsrc)
        self.h = self.L.ora_create(imt, jmt, km, nt, nsrc)
        self._views = {}

    def close(self):
        if self.h:
            self._views.clear()
            self.L.ora_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def raw(self, name):
        if name not in self._views:
            n = ctypes.c_size_t()
            isint = ctypes.c_int()
            p = self.L.ora_array(self.h, name.encode(), ctypes.byref(n), ctypes.byref(isint))
            if not p:
                raise KeyError(name)
            ct = ctypes.c_int32 if isint.value else ctypes.c_double
            buf = (ct * n.value).from_address(p)
            self._views[name] = np.frombuffer(buf, dtype=np.int32 if isint.value else np.float64)
        return self._views[name]

    def arr(self, name, shape=None):
        a = self.raw(name)
        return a.reshape(shape) if shape is not None else a

    def set(self, name, value):
        a = self.raw(name)
        v = np.ascontiguousarray(value).reshape(-1)
        assert v.size == a.size, (name, v.size, a.size)
        a[:] = v

    def set_scalar(self, name, v):
        r = self.L.ora_set_scalar(self.h, name.encode(), float(v))
        assert r == 0, name

    # shapes -----------------------------------------------------------------
    def shape3(self):
        imt, jmt, km, nt, nsrc = self.dims
        return (jmt, km, imt)

    def shape3z(self):
        imt, jmt, km, nt, nsrc = self.dims
        return (jmt, km + 1, imt)

    def t(self):
        imt, jmt, km, nt, nsrc = self.dims
        return self.arr("t", (3, nt, jmt, km, imt))

    def load_case(self, case):
        """Copy every array / scalar of a synthetic Case that the oracle knows."""
        for k, v in case.arrays.items():
            if k.startswith("_"):
                continue
            try:
                self.set(k, v)
            except KeyError:
                pass
        for k, v in case.scalars.items():
            self.L.ora_set_scalar(self.h, k.encode(), float(v))

    def call(self, fn, *args):
        getattr(self.L, fn)(self.h, *args)
