"""uvic_b200_group_*: several contexts driven by ONE host thread (the serial Fortran host's multi-GPU mode).  The group cuts
the global arrays into latitude slabs, exchanges the 2-row halos with peer copies ordered by events, and must reproduce the
single-context result BIT FOR BIT.  On a one-GPU box both slabs live on device 0 (same code path, same copies); with two or
more GPUs the second test uses distinct devices (NVLink peer copies)."""
import numpy as np
import pytest

from conftest import load_pkg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    return load_pkg()


def _run(pkg, devices, mobi, **kw):
    case = pkg.synthetic.make_case(**kw)
    one = pkg.TracerContext(case, mobi=mobi, device=devices[0])
    one.load_state()
    grp = pkg.api.TracerGroup(case, devices, mobi=mobi)
    grp.load_state()
    rows = [grp.rows(r) for r in range(len(devices))]
    assert rows[0][0] == 2 and rows[-1][1] == case.jmt - 1 and all(rows[r][1] + 1 == rows[r + 1][0] for r in range(len(rows) - 1))
    for itt, lf in enumerate((True, True, False, True, True)):
        nxt = (True, False, True, True, True)[itt]
        one.step(leapfrog=lf, next_leapfrog=nxt)
        grp.step(leapfrog=lf, next_leapfrog=nxt)
        a, b = one.download_t(+1), grp.download_t(+1)
        assert np.array_equal(a[:, 1:-1], b[:, 1:-1]), (itt, np.abs(a - b).max())
        ia, ib = one.inventory(+1), grp.inventory(+1)
        # different summation order (per-slab partial sums added in device order): equal to rounding, not to the bit
        assert np.allclose(ia, ib, rtol=1e-13, atol=0), (itt, np.abs(ia - ib).max())
        one.rotate()
        grp.rotate()
    one.close()
    grp.close()


def test_group_of_three_slabs_on_one_device_matches_single_context(pkg):
    _run(pkg, [0, 0, 0], 0, imt=42, jmt=38, km=12, nt=4, names=["temp", "salt", "passive0", "passive1"], seed=13)


def test_group_with_mobi_matches_single_context(pkg):
    _run(pkg, [0, 0], 1, imt=34, jmt=30, km=10, nt=37, seed=19)


def test_group_on_distinct_devices(pkg):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _run(pkg, list(range(min(4, torch.cuda.device_count()))), 1, imt=62, jmt=54, km=19, nt=37, seed=5)
