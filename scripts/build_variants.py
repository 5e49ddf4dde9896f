#!/usr/bin/env python
"""Experiment builds of the library beside the product build (selected at run time with UVIC_B200_LIB=<path>):
    python scripts/build_variants.py fmad_fct=k_fct.cu fmad_all=k_fct.cu,k_mobi.cu,k_tracer.cu
writes uvic2.9_b200/variants/libuvic_b200_<name>.so with -fmad=true for the named translation units."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("uvic_build", os.path.join(ROOT, "uvic2.9_b200", "build.py"))
b = importlib.util.module_from_spec(spec)
spec.loader.exec_module(b)
os.makedirs(os.path.join(ROOT, "uvic2.9_b200", "variants"), exist_ok=True)
for arg in sys.argv[1:]:
    name, files = arg.split("=", 1)
    defines = [f for f in files.split(",") if f.startswith("-D")]
    fm = {f: "true" for f in files.split(",") if f and not f.startswith("-D")} or None   # no file named: the default FMA set
    out = os.path.join(ROOT, "uvic2.9_b200", "variants", f"libuvic_b200_{name}.so")
    print(b.build(force=True, fmad=fm, out=out, objdir_name=f"build_{name}", defines=defines))
