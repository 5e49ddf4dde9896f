"""Size-independent properties of the CUDA path at BASELINE.json's full single-GPU size (0.5 degree: 720 x 360 x 40),
where the oracle would need minutes: conservation, affine equivariance, masks and the cyclic boundary.  The same
properties are checked on the oracle at small size in tests/test_cpu_oracle.py (and were checked once on the oracle at
this size: equivariance 2e-8 of the field maximum, conservation 2e-16).  The file sorts last on purpose: it is the
slowest GPU test (the synthetic case takes ~10 s to build on the host)."""
import numpy as np
import pytest

from conftest import load_pkg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    return load_pkg()


def test_half_degree_step_properties(pkg):
    names = ["temp", "salt", "passive0", "passive1"]
    case = pkg.synthetic.make_case(imt=722, jmt=362, km=40, nt=4, names=names, seed=2901)
    t = case["t"]
    t[:, 3] = (2.0 * t[:, 2] + 3.0) * case["tmask"][None]
    tmask = np.asarray(case["tmask"])
    ctx = pkg.TracerContext(case)
    ctx.load_state()
    for step, lf in enumerate((True, False, True)):
        inv_m1 = ctx.inventory(-1 if lf else 0)      # a mixing step starts from t(tau) (09/mom/loadmw.F:109-111)
        ctx.step(leapfrog=lf)
        inv_p1 = ctx.inventory(+1)
        tp = ctx.download_t(+1)
        assert np.isfinite(tp).all(), step
        # zero surface / bottom flux: sum(t dV) of every tracer is conserved
        rel = np.abs(inv_p1 - inv_m1) / np.abs(inv_m1)
        assert (rel <= 1e-11).all(), (step, rel)
        # 2 p + 3 stays 2 p' + 3 (to the bottom closure of the synthetic flow's continuity)
        d = (tp[3] - (2.0 * tp[2] + 3.0)) * tmask
        assert np.abs(d[1:-1, :, 1:-1]).max() < 1e-6 * np.abs(tp[3]).max(), step
        # land stays zero, the cyclic columns are copies
        assert np.all(tp * (1.0 - tmask)[None] == 0.0)
        assert np.array_equal(tp[..., 0], tp[..., -2]) and np.array_equal(tp[..., -1], tp[..., 1])
        # the step did something
        assert np.abs(tp[2] - ctx.download_t(0)[2]).max() > 1e-6
        ctx.rotate()
    ctx.close()
