"""Dump the MOBI pre-pass fields and sources of one step (A/B check of a kernel change: run before and after, compare)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import load_pkg
pkg = load_pkg()
case = pkg.synthetic.make_case(imt=62, jmt=50, km=19, nt=37)
ctx = pkg.TracerContext(case, mobi=1)
ctx.load_state()
ctx.step(True)
ctx.synchronize()
out = {n: ctx.fetch(n) for n in ("mobi_pre", "src", "mobi_day")}
out["t"] = ctx.download_t(+1)
np.savez_compressed(sys.argv[1], **out)
print("saved", sys.argv[1], {k: v.shape for k, v in out.items()})
