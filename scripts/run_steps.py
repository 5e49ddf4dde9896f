#!/usr/bin/env python
"""Run a few device-resident steps of a bench workload (profiling target for ncu)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="uvic100_mobi37")
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
pkg = bench.load_pkg()
case = bench.make_case(pkg, a.workload, 1)
ctx = pkg.TracerContext(case, mobi=bench.WORKLOADS[a.workload]["mobi"])
ctx.load_state()
for s in range(a.steps):
    ctx.step(leapfrog=pkg.timestep.is_leapfrog(s + 1, 16))
    ctx.rotate()
ctx.synchronize()
print("ok", ctx.kernel_launches)
ctx.close()
