#!/usr/bin/env python
"""Builds oracle/_ref/libref.so: the reference's OWN Fortran, cpp-expanded with the options of
run/mk.in exactly as `mk` does it (mk:2397-2410: `#include` -> Fortran include, then
`cpp -traditional -P` with the -D list; include files get the same treatment), then translated
to C by the mechanical translator f2c.py (no hand editing) and compiled with
`gcc -O2 -ffp-contract=off -fno-fast-math` -- the same flags as the hand-written oracle, so the
two can be compared BIT FOR BIT (tests/test_cpu_refpin.py).

TEST INFRASTRUCTURE.  Reads /root/reference where it lies; writes only into oracle/_ref/
(git-ignored; travels to the GPU box as a built .so).  Nothing of the reference is copied
into the repository.

    python oracle/refgen/gen.py [--ref /root/reference] [--set imt=34 jmt=26 km=8 ...] [--tag NAME]

Routines translated (file, as `mk` would resolve it: updates/09/source/<dir>/ first, then source/<dir>/):
see UNITS below.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import f2c  # noqa: E402

OUT = os.path.normpath(os.path.join(HERE, "..", "_ref"))

# mk search order (run/mk.in Source_Directory(1..10)): updates/09 shadows source/
SUBDIRS = ["common", "netcdf", "embm", "ice", "mtlm", "mom", "sed"]

# file -> program units taken from it.  Everything else in those files is ignored.
UNITS = {
    "mobi.F": ["mobi_init", "setimobi", "mobi_driver", "mobi_src"],
    "co2calc.F": ["co2calc_sws", "drtsafe", "ta_iter_sws"],
    "invtri.F": ["invtri"],
    "state.F": ["state", "statec", "state_ref"],
    "convect.F": ["convct2"],
    "isopyc.F": ["isopyc", "elements", "ai_east", "ai_north", "ai_bottom", "isoflux", "isopyc_adv"],
    "tracer_adv_flx.F": ["adv_flux"],
    "vmixc.F": ["vmixc"],
    "adv_vel.F": ["adv_vel"],
    "util.F": ["setbcx", "areaavg"],
    "tracer.F": ["tracer", "diagt1", "diagt2", "ivdift"],
    "filt.F": ["filt"],
    "filtr.F": ["filtr"],
    "findex.F": ["findex"],
    "set_sbc.F": ["set_sbc"],
    "setvbc.F": ["setvbc"],
    "clinic.F": ["clinic", "diagc1", "diagc2", "asbcu", "isbcu"],
    "filuv.F": ["filuv"],
    "gasbc.F": ["gasbc"],
}

# I/O helpers of the reference (unit management, netCDF reads, name mangling): calls are dropped, not translated
SKIP_CALLS = {n: "I/O helper" for n in (
    "getunit", "relunit", "openfile", "closefile", "getvara", "getvars", "putvara", "putvars", "defvar", "defdim",
    "new_file_name", "file_names", "opennew", "opennext", "redef", "enddef")}
# gasbc (09/common/gasbc.F) is the atmosphere's coupling routine: its air-sea gas exchange loop is on the path (the other
# caller of co2calc_SWS); the forcing-data readers and the atmosphere's own physics it also calls are not
SKIP_CALLS.update({n: "forcing data reader" for n in ("solardata", "co2ccndata", "data")})
SKIP_CALLS.update({n: "atmosphere model (outside the path)" for n in ("decl", "zenith", "co2forc", "ta_embm_tavg", "ta_embm_tsi")})


# diagnostics behind run-time switches that the pin tests leave off (gyrets, trmbts, ...): a call aborts if ever reached
STUB_CALLS = {"gyre", "ttb1", "ttb2", "ge1", "ge2", "utb1", "utb2"}


def find(ref, name):
    for base in ("updates/09/source", "source"):
        for d in SUBDIRS:
            p = os.path.join(ref, base, d, name)
            if os.path.exists(p):
                return p
    return None


def mk_options(ref):
    opts = []
    for ln in open(os.path.join(ref, "run", "mk.in")):
        m = re.match(r"\s*(O_[A-Za-z0-9_]+)\s*$", ln)
        if m:
            opts.append("-D" + m.group(1))
    return opts


def cpp_expand(path, opts):
    """what mk does to one file: #include -> include, then cpp -traditional -P"""
    src = open(path, errors="replace").read()
    lines = []
    for ln in src.split("\n"):
        if "include" in ln:
            ln = ln.replace("<", '"').replace(">", '"') if ln.lstrip().startswith("#") else ln
        if ln.lstrip().startswith("#"):
            ln = re.sub(r"#.*include", "      include", ln)
            ln = re.sub(r"endif.*", "endif", ln)
            ln = re.sub(r"else.*", "else", ln)
        lines.append(ln)
    r = subprocess.run(["cpp", "-traditional", "-P", *opts], input="\n".join(lines), capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"cpp failed on {path}: {r.stderr[:500]}")
    return r.stdout


def inline_includes(text, ref, opts, seen=()):
    out = []
    for ln in text.split("\n"):
        m = re.match(r"\s+include\s+[\"']([^\"']+)[\"']\s*$", ln, re.I)
        if m and not ln[:1] in "cC*!":
            p = find(ref, m.group(1))
            if p is None:
                raise RuntimeError("include not found: " + m.group(1))
            if p in seen:
                raise RuntimeError("recursive include " + p)
            out.append(inline_includes(cpp_expand(p, opts), ref, opts, seen + (p,)))
        else:
            out.append(ln)
    return "\n".join(out)


def read_namelists(ref):
    """run/control.in -> {group: {member: value}} (numbers and logicals; what the dropped NAMELIST reads would have set)"""
    text = open(os.path.join(ref, "run", "control.in")).read()
    out = {}
    for m in re.finditer(r"&([A-Za-z0-9_]+)(.*?)/", text, re.S):
        grp, body = m.group(1).lower(), m.group(2)
        vals = {}
        for mm in re.finditer(r"([A-Za-z0-9_]+)\s*=\s*([^=]*?)(?=(?:[A-Za-z0-9_]+\s*=)|$)", body, re.S):
            raw = [x.strip() for x in mm.group(2).replace("\n", " ").split(",") if x.strip()]
            conv = []
            for x in raw:
                xl = x.lower()
                if xl in (".true.", ".false."):
                    conv.append(xl == ".true.")
                else:
                    try:
                        conv.append(int(x) if re.fullmatch(r"[-+]?\d+", x) else float(xl.replace("d", "e")))
                    except ValueError:
                        conv.append(x.strip("'\""))
            vals[mm.group(1).lower()] = conv[0] if len(conv) == 1 else conv
        out[grp] = vals
    return out


def generate(ref, overrides, only_files=None, undef=()):
    opts = [o for o in mk_options(ref) if o[2:] not in set(undef)]
    tr = f2c.Translator(skip_calls=SKIP_CALLS, overrides=overrides, stub_calls=STUB_CALLS)
    wanted = set()
    for fn, units in UNITS.items():
        if only_files and fn not in only_files:
            continue
        wanted |= set(units)
    tr.known_units |= wanted
    for fn, units in UNITS.items():
        if only_files and fn not in only_files:
            continue
        p = find(ref, fn)
        if p is None:
            raise RuntimeError("source not found: " + fn)
        text = inline_includes(cpp_expand(p, opts), ref, opts)
        tr.load(f2c.read_fixed_form(text), only=set(units))
    return tr, opts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--set", nargs="*", default=[], help="override size.h parameters, e.g. imt=34 jmt=26 km=8")
    ap.add_argument("--tag", default=None, help="output name: oracle/_ref/libref_<tag>.so (default: libref.so)")
    ap.add_argument("--files", nargs="*", default=None)
    ap.add_argument("--undef", nargs="*", default=[], help="options of run/mk.in to leave out, e.g. O_carbon_13 O_carbon_14 O_mobi_nitrogen_15")
    ap.add_argument("--keep-going", action="store_true")
    a = ap.parse_args()
    if not os.path.isdir(a.ref):
        print(f"gen.py: {a.ref} not present (GPU box): keeping the prebuilt oracle/_ref", file=sys.stderr)
        return 0
    overrides = {}
    for kv in a.set:
        k, v = kv.split("=")
        overrides[k.strip().lower()] = int(v)
    tr, opts = generate(a.ref, overrides, a.files, a.undef)
    csrc = tr.emit()
    os.makedirs(OUT, exist_ok=True)
    suffix = f"_{a.tag}" if a.tag else ""
    cpath = os.path.join(OUT, f"ref_gen{suffix}.c")
    open(cpath, "w").write(csrc)
    man = tr.manifest()
    man["overrides"] = overrides
    man["cpp_options"] = opts
    man["namelists"] = read_namelists(a.ref)
    json.dump(man, open(os.path.join(OUT, f"ref_gen{suffix}.json"), "w"), indent=1)
    so = os.path.join(OUT, f"libref{suffix}.so")
    cmd = ["gcc", "-O2", "-march=x86-64-v3", "-fno-fast-math", "-ffp-contract=off", "-fPIC", "-std=gnu11", "-w", "-shared", "-o", so, cpath, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stderr[:4000])
        return 1
    print(f"{so}: {len(man['routines'])} routines, {len(man['commons'])} COMMON members, {len(man['dropped'])} statements dropped (I/O)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
