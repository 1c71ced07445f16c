"""GPU parity of the chained hot path: the routines run back to back on device-resident
state for several steps (time levels swapping) and every registered array is compared with
the oracle after each routine.  Tolerance: 1e-10 relative to the field max-norm after N
chained steps, parity build."""
import numpy as np
import pytest

from blom_b200.driver import STEP_SEQUENCE, available_routines
from blom_b200.lib import time_levels
from util import Case, interior, max_rel_err

pytestmark = pytest.mark.gpu

SKIP = {"depths"}


def run_routine(b, r, lv):
    m, n, mm, nn, k1m, k1n = lv
    if r == "tmsmt1":
        b.tmsmt1(nn)
    elif r == "tmsmt2":
        b.tmsmt2(m, mm, nn, k1m)
    else:
        getattr(b, r)(m, n, mm, nn, k1m, k1n)


@pytest.mark.parametrize("cfg", ["tiny0", "tiny1", "tiny2", "tiny3", "tiny4", "fuk95"])
def test_chained_steps(cfg):
    c = Case(cfg, ntr=1, nstep=1)
    o = c.new_oracle(); g = c.new_gpu(parity=True)
    try:
        for b in (o, g):
            b.inieos(); b.numerical_bounds(); b.init_cppm()
        routines = [r for r in STEP_SEQUENCE if r in available_routines()]
        kk = c.dims[2]
        for nstep in (1, 2, 3):
            lv = time_levels(nstep, kk)
            for b in (o, g):
                b.set_scalar("nstep", nstep)
            for r in routines:
                run_routine(o, r, lv); run_routine(g, r, lv)
                g.download_all()
                bad = []
                for nm, a in g.arrays.items():
                    if nm in SKIP or a.dtype != np.float64:
                        continue
                    err = max_rel_err(interior(a), interior(o.arrays[nm]))
                    if not err <= 1e-10:
                        bad.append((nm, err))
                assert not bad, (cfg, nstep, r, sorted(bad, key=lambda t: -t[1])[:6])
        assert np.isfinite(g.arrays["dp"]).all()
    finally:
        g.finalize()
