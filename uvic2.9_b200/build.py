"""Build the CUDA library (libuvic_b200.so, sm_100a only) in-tree with nvcc."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libuvic_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

# -fmad=false (the default for a translation unit): no FMA contraction, so the operation order written in the kernels
# (which follows the reference's Fortran) is what executes -- the CPU oracle is built the same way (-ffp-contract=off)
# and the coefficient kernels (elements, Redi / GM coefficients, vmixc, adv_vel, state, clinic) agree with it to the last
# bit.  The three FP64-issue bound translation units -- the FCT, the diffusion / implicit solve and MOBI -- are compiled
# WITH contraction since round 2: their gate is 1e-12 (tracers) / 1e-10 (MOBI sources) against the oracle, which they
# keep, and a*b+c costs one FP64 issue slot instead of two (measured on B200, 0.5 degree x 40 tracers: march 11.35 ->
# 11.08 ms, MOBI column 6.63 -> 5.96 ms, diffusion 3.88 -> 3.57 ms, step 28.5 -> 27.2 ms).
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
]
FMAD = {"k_fct.cu": "true", "k_tracer.cu": "true", "k_mobi.cu": "true"}
if "UVIC_B200_FMAD" in os.environ:   # experiment switch: the comma separated list replaces the default ("" = none)
    FMAD = {f: "true" for f in os.environ["UVIC_B200_FMAD"].split(",") if f}


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "uvic_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra=(), fmad=None, out=None, objdir_name="build", defines=()):
    """fmad: {file: "true"} overrides; out: another library path (experiment builds keep the product library untouched)"""
    global FMAD
    lib = out or LIB
    if fmad is not None:
        FMAD = dict(fmad)
    if not force and not needs_build() and out is None:
        return LIB
    objdir = os.path.join(HERE, objdir_name)
    os.makedirs(objdir, exist_ok=True)

    def cc(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [NVCC, *FLAGS, f"-fmad={FMAD.get(src, 'false')}", *extra, *defines, *os.environ.get("UVIC_B200_NVCC_EXTRA", "").split(), "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose and (r.stdout or r.stderr):
            print(r.stdout, r.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(cc, sources()))
    r = subprocess.run([NVCC, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True,
                extra=("-Xptxas", "-v") if "--ptxas" in sys.argv else ()))
