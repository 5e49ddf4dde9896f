"""ctypes access to oracle/_ref/libref*.so -- the reference's own Fortran, cpp-expanded and translated to C mechanically
by oracle/refgen (see gen.py / f2c.py).  Test infrastructure only: the hand-written oracle is pinned against it."""
from __future__ import annotations

import ctypes
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
GEN = os.path.join(ROOT, "oracle", "refgen", "gen.py")
REFERENCE = "/root/reference"

# the variants the tests use: tag -> size.h overrides (nt = 37, nsrc = 35 come from the options of run/mk.in)
VARIANTS = {
    "s": {"imt": 34, "jmt": 26, "km": 8},
    "t": {"imt": 20, "jmt": 16, "km": 6},      # tests/golden/ref_step_t.npz is made from this one
    # BASELINE config 2: "full MOBI tracer set, no isotopes" -- the reference built WITHOUT O_carbon_13, O_carbon_14,
    # O_mobi_nitrogen_15 (nt = 21)
    "n": {"imt": 34, "jmt": 26, "km": 8},
}
UNDEF = {"n": ["O_carbon_13", "O_carbon_14", "O_mobi_nitrogen_15"]}


def build_variant(tag, force=False):
    """(re)generate oracle/_ref/libref_<tag>.so when /root/reference is present; otherwise use the prebuilt file"""
    so = os.path.join(REFDIR, f"libref_{tag}.so")
    srcs = [GEN, os.path.join(os.path.dirname(GEN), "f2c.py")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if (force or stale) and os.path.isdir(REFERENCE):
        sets = [f"{k}={v}" for k, v in VARIANTS[tag].items()]
        undef = ["--undef", *UNDEF[tag]] if tag in UNDEF else []
        r = subprocess.run([sys.executable, GEN, "--tag", tag, "--set", *sets, *undef], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle/refgen/gen.py failed:\n" + r.stdout[-2000:] + r.stderr[-4000:])
    return so if os.path.exists(so) else None


_HOOK_T = ctypes.CFUNCTYPE(None, ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p)


class RefLib:
    def __init__(self, tag):
        so = build_variant(tag)
        if so is None:
            raise FileNotFoundError(f"oracle/_ref/libref_{tag}.so is missing and /root/reference is not here to build it")
        # a private copy per instance would be needed for two independent states; the tests use one at a time
        self.L = ctypes.CDLL(so)
        self.man = json.load(open(os.path.join(REFDIR, f"ref_gen_{tag}.json")))
        self.dims = dict(VARIANTS[tag])
        self._views = {}
        self._hook = None

    # ---- COMMON members -------------------------------------------------------------------
    def _entry(self, name, block=None):
        ents = self.man["commons"].get(name.lower())
        if not ents:
            raise KeyError(name)
        if block is not None:
            ents = [e for e in ents if e["block"] == block]
        if len(ents) != 1:
            raise KeyError(f"{name}: ambiguous, blocks {[e['block'] for e in ents]}")
        return ents[0]

    def has(self, name):
        return name.lower() in self.man["commons"] and len(self.man["commons"][name.lower()]) == 1

    def view(self, name, block=None):
        """numpy view: arrays in C order with the Fortran dimensions reversed; scalars as 1-element arrays"""
        key = (name.lower(), block)
        if key not in self._views:
            e = self._entry(name, block)
            ct = ctypes.c_double if e["type"] == "real" else ctypes.c_int
            dt = np.float64 if e["type"] == "real" else np.int32
            if e["dims"] is None:
                buf = (ct * 1).in_dll(self.L, e["symbol"])
                self._views[key] = np.frombuffer(buf, dtype=dt)
            else:
                shape = [d[1] for d in e["dims"]][::-1]
                n = int(np.prod(shape))
                buf = (ct * n).in_dll(self.L, e["symbol"])
                self._views[key] = np.frombuffer(buf, dtype=dt).reshape(shape)
        return self._views[key]

    def lower_bounds(self, name, block=None):
        e = self._entry(name, block)
        return [d[0] for d in e["dims"]][::-1] if e["dims"] else None

    def set(self, name, value, block=None):
        v = self.view(name, block)
        v[...] = value

    def get(self, name, block=None):
        v = self.view(name, block)
        return v[0] if v.shape == (1,) and self._entry(name, block)["dims"] is None else v

    # ---- routines ---------------------------------------------------------------------------
    def call(self, routine, *args):
        r = self.man["routines"][routine.lower()]
        fn = getattr(self.L, routine.lower() + "_")
        fn.restype = {None: None, "real": ctypes.c_double, "int": ctypes.c_int, "logical": ctypes.c_int}[r["returns"]]
        assert len(args) == len(r["args"]), (routine, len(args), len(r["args"]))
        keep, cargs = [], []
        for a, spec in zip(args, r["args"]):
            if isinstance(a, np.ndarray):
                want = np.float64 if spec["type"] == "real" else np.int32
                assert a.dtype == want and a.flags.c_contiguous, (routine, spec["name"], a.dtype)
                cargs.append(ctypes.c_void_p(a.ctypes.data))
            elif spec["type"] == "real":
                c = ctypes.c_double(float(a))
                keep.append(c)
                cargs.append(ctypes.byref(c))
            else:
                c = ctypes.c_int(int(a))
                keep.append(c)
                cargs.append(ctypes.byref(c))
        out = fn(*cargs)
        # scalar dummies after the call (Fortran passes by reference: outputs come back here)
        self.last = {spec["name"]: c.value for spec, c in zip([sp for a, sp in zip(args, r["args"]) if not isinstance(a, np.ndarray)], keep)}
        return out

    def set_namelist_values(self, groups=None):
        """feed the dropped NAMELIST reads: {group: {member: value}}; default = run/control.in as recorded by gen.py"""
        groups = groups if groups is not None else self.man["namelists"]

        def hook(grp, n, names, ptrs, types):
            vals = groups.get(grp.decode(), {})
            for i in range(n):
                nm = names[i].decode()
                if nm in vals and not isinstance(vals[nm], (list, str)):
                    if types[i:i + 1] == b"r":
                        ctypes.c_double.from_address(ptrs[i]).value = float(vals[nm])
                    else:
                        ctypes.c_int.from_address(ptrs[i]).value = int(vals[nm])

        self._hook = _HOOK_T(hook)
        ctypes.c_void_p.in_dll(self.L, "f2c_namelist_hook").value = ctypes.cast(self._hook, ctypes.c_void_p).value


# ---- moving state between the hand-written oracle (tests/oracle_ffi.py) and the translated reference -------------------

def _match(ref, name, n_oracle, jmt):
    """how the oracle's flat array `name` (allocated 1:jmt in j) maps onto the reference's COMMON member:
    -> (shape of the oracle array, slice into it) or None when the layouts cannot be matched"""
    shape = list(ref.view(name).shape)
    lows = ref.lower_bounds(name)
    if lows is None:
        return None
    if int(np.prod(shape)) == n_oracle:
        return shape, tuple(slice(None) for _ in shape)
    for d, (ext, lo) in enumerate(zip(shape, lows)):
        if ext in (jmt - 1, jmt - 2, jmt - 3) and lo in (1, 2):
            full = shape[:d] + [jmt] + shape[d + 1:]
            if int(np.prod(full)) == n_oracle:
                return full, tuple(slice(lo - 1, lo - 1 + ext) if i == d else slice(None) for i in range(len(shape)))
    return None


def oracle_array_names(o):
    o.L.ora_narrays.argtypes = [ctypes.c_void_p]
    o.L.ora_array_name.argtypes = [ctypes.c_void_p, ctypes.c_int]
    o.L.ora_array_name.restype = ctypes.c_char_p
    return [o.L.ora_array_name(o.h, i).decode() for i in range(o.L.ora_narrays(o.h))]


# oracle name -> reference COMMON name where they differ
RENAME = {"eosc": "c", "R_plusY": "r_plusy", "R_minusY": "r_minusy"}


def oracle_to_ref(o, ref, only=None, skip=()):
    """copy every array the two sides share by name; returns the names copied"""
    jmt = o.dims[1]
    done = []
    for nm in oracle_array_names(o):
        rn = RENAME.get(nm, nm).lower()
        if (only is not None and nm not in only) or nm in skip or not ref.has(rn):
            continue
        a = o.raw(nm)
        m = _match(ref, rn, a.size, jmt)
        if m is None:
            continue
        full, sl = m
        ref.view(rn)[...] = a.reshape(full)[sl]
        done.append(nm)
    return done


def compare(o, ref, name, interior_j=None, rname=None):
    """-> (max |difference|, number of differing elements) between the oracle's array and the reference's, on the
    reference's index range"""
    rn = (rname or RENAME.get(name, name)).lower()
    a = o.raw(name)
    m = _match(ref, rn, a.size, o.dims[1])
    assert m is not None, name
    full, sl = m
    x, y = a.reshape(full)[sl], ref.view(rn)
    d = np.abs(x - y)
    return float(d.max()), int((x != y).sum())
