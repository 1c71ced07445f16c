"""GPU parity of advect with advmth='remap' (incremental remapping, phy/mod_remap.F90 +
phy/mod_advect.F90:96-153) against the oracle on identical inputs, through the C ABI.

Tolerances (float64): parity build (-fmad=false, the reference's operation order) <= 1e-13 of the
field's max-norm on the interior plus the one halo ring remap updates (mrg=1); the dpeps offset the
reference leaves on halo rings 2..3 of dp must be reproduced bit for bit; performance build (FMA
contraction) <= 1e-10 (the departure-polygon moments difference nearly equal products)."""
import numpy as np
import pytest

from util import Case, interior, max_rel_err

pytestmark = pytest.mark.gpu

FLUXES = ["uflx", "vflx", "utflx", "vtflx", "usflx", "vsflx"]


def run_pair(cfg, ntr, nstep, parity):
    c = Case(cfg, ntr=ntr, nstep=nstep)
    o = c.new_oracle(); g = c.new_gpu(parity=parity)
    o.set_option("advmth", "remap"); g.set_option("advmth", "remap")
    o.advect(*c.levels); g.advect(*c.levels)
    g.download_all()
    return c, o, g


@pytest.mark.parametrize("cfg", ["tiny0", "tiny1", "tiny2", "tiny3", "tiny4"])
@pytest.mark.parametrize("nstep", [1, 2])
@pytest.mark.parametrize("ntr", [0, 1, 2])
def test_remap_parity_build(cfg, nstep, ntr):
    c, o, g = run_pair(cfg, ntr, nstep, parity=True)
    try:
        kk, nn = c.dims[2], c.levels[3]
        for nm in ["dp", "temp", "saln"] + (["trc"] if ntr else []) + FLUXES:
            err = max_rel_err(interior(g.arrays[nm], halo=1), interior(o.arrays[nm], halo=1))
            assert err <= 1e-13, (nm, err)
        for nm in ("cau", "cav"):
            assert max_rel_err(interior(g.arrays[nm], halo=3), interior(o.arrays[nm], halo=3)) <= 1e-15, nm
        # rings 2..3 of dp: untouched by the update, left at max(0,dp)+dpeps (mod_remap.F90:304-311)
        a, b = interior(g.arrays["dp"], halo=3), interior(o.arrays["dp"], halo=3)
        ring = np.ones(a.shape[-2:], bool); ring[2:-2, 2:-2] = False
        assert np.array_equal(a[:, ring], b[:, ring])
        assert np.abs(interior(g.arrays["dp"][nn:nn + kk]) - interior(c.state["dp"][nn:nn + kk])).max() > 1.0
    finally:
        g.finalize()


@pytest.mark.parametrize("cfg", ["tiny2", "tiny3", "fuk95"])
def test_remap_perf_build(cfg):
    c, o, g = run_pair(cfg, 1, 1, parity=False)
    try:
        for nm in ["dp", "temp", "saln", "trc"] + FLUXES:
            err = max_rel_err(interior(g.arrays[nm], halo=1), interior(o.arrays[nm], halo=1))
            assert err <= 1e-10, (nm, err)
    finally:
        g.finalize()


def test_remap_conserves_on_device():
    """Inventories through the device path alone (no oracle): round-off conservation on the
    tripolar grid, rows below the duplicated fold row."""
    c = Case("tiny2", ntr=1, nstep=1)
    g = c.new_gpu(parity=False)
    try:
        g.set_option("advmth", "remap")
        kk, nn = c.dims[2], c.levels[3]
        a = g.arrays
        scp2 = interior(a["scp2"][0])[:-1]

        def inv(nm):
            f = interior(a["dp"][nn:nn + kk])[:, :-1]
            if nm:
                f = f * interior(a[nm][nn:nn + kk])[:, :-1]
            return (f * scp2).sum()
        before = {nm: inv(nm) for nm in ("", "temp", "saln", "trc")}
        g.advect(*c.levels); g.download_all()
        for nm, v0 in before.items():
            assert abs(inv(nm) - v0) <= 4e-15 * abs(v0), nm
    finally:
        g.finalize()
