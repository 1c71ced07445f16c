"""Pins the oracle's incremental-remapping restatement (phy/mod_remap.F90, advmth='remap') on CPU
through the scheme's own guarantees (the reference ships no golden outputs):
 - flux form: mass, heat, salt and tracer inventories conserved to round-off, tripolar fold included
   (remap has no counterpart of the cppm hel/her half-row quirk);
 - a spatially uniform tracer stays uniform to round-off (compatibility of mass and tracer fluxes);
 - monotonicity: with limited gradients no new extrema appear (Dukowicz & Baumgardner 2000), checked
   against the 3x3 wet neighbourhood of every cell;
 - the reference's flux bookkeeping quirk: u fluxes are accumulated, v fluxes assigned
   (phy/mod_remap.F90:1054-1056 vs :1455-1457)."""
import numpy as np
import pytest

from util import Case, interior
from test_oracle_cppm import inventories


def run(cfg, ntr=1, nstep=1, mutate=None):
    c = Case(cfg, ntr=ntr, nstep=nstep)
    if mutate:
        mutate(c)
    o = c.new_oracle(); o.set_option("advmth", "remap")
    return c, o


@pytest.mark.parametrize("cfg", ["tiny0", "tiny1", "tiny2", "tiny3", "tiny4"])
@pytest.mark.parametrize("nstep", [1, 2])
def test_conservation_roundoff(cfg, nstep):
    c, o = run(cfg, nstep=nstep)
    nn, kk = c.levels[3], c.dims[2]
    rows = slice(0, -1) if cfg == "tiny2" else slice(None)  # row jj duplicates row jj-1 on the fold
    inv0 = inventories(o.arrays, c, nn, rows)
    o.advect(*c.levels)
    inv1 = inventories(o.arrays, c, nn, rows)
    for k in inv0:
        assert abs(inv1[k] - inv0[k]) <= 4e-15 * abs(inv0[k]), (k, inv0[k], inv1[k])
    assert o.arrays["dp"].min() >= 0.0
    assert np.abs(o.arrays["dp"][nn:nn + kk] - c.state["dp"][nn:nn + kk]).max() > 1.0


@pytest.mark.parametrize("cfg", ["tiny1", "tiny2", "tiny3"])
def test_uniform_tracer_stays_uniform(cfg):
    def mutate(c):
        c.state["trc"][:] = 3.25
        c.state["temp"][:] = 7.5
    c, o = run(cfg, mutate=mutate)
    nn, kk = c.levels[3], c.dims[2]
    o.advect(*c.levels)
    ip = interior(c.masks["ip"]) == 1
    assert np.abs(interior(o.arrays["temp"][nn:nn + kk])[:, ip] - 7.5).max() <= 1e-13
    assert np.abs(interior(o.arrays["trc"][nn:nn + kk])[:, ip] - 3.25).max() <= 1e-13


@pytest.mark.parametrize("cfg", ["tiny1", "tiny3"])
def test_no_new_extrema(cfg):
    c, o = run(cfg)
    nn, kk = c.levels[3], c.dims[2]
    ipf = c.masks["ip"] == 1
    for nm in ("temp", "saln"):
        old = c.state[nm][nn:nn + kk]
        lo = np.where(ipf, old, np.inf); hi = np.where(ipf, old, -np.inf)
        nlo, nhi = lo.copy(), hi.copy()
        for dj in (-1, 0, 1):
            for di in (-1, 0, 1):
                nlo = np.minimum(nlo, np.roll(lo, (dj, di), axis=(1, 2)))
                nhi = np.maximum(nhi, np.roll(hi, (dj, di), axis=(1, 2)))
        if nm == "temp":
            o.advect(*c.levels)
        new = interior(o.arrays[nm][nn:nn + kk]); ip = interior(ipf)
        # only layers that still hold water carry a meaningful tracer value
        wet = ip & (interior(o.arrays["dp"][nn:nn + kk]) > 1e-6)
        slack = 1e-10 * np.abs(old).max()
        assert (new[wet] >= interior(nlo)[wet] - slack).all(), nm
        assert (new[wet] <= interior(nhi)[wet] + slack).all(), nm


def test_v_fluxes_assigned_u_fluxes_accumulated():
    big = 1.0e18  # well above the fluxes of one call (~1e16)

    def mutate(c):
        c.state["uflx"][:] = big
        c.state["vflx"][:] = big
    # (the oracle library holds one global state: finish one run before creating the next)
    c, o = run("tiny1", ntr=0, mutate=mutate)
    o.advect(*c.levels)
    c2, o2 = run("tiny1", ntr=0)
    o2.advect(*c2.levels)
    mm, kk = c.levels[2], c.dims[2]
    iu = interior(c.masks["iu"]) == 1; iv = interior(c.masks["iv"]) == 1
    sl = slice(mm, mm + kk)
    du = (interior(o.arrays["uflx"][sl]) - interior(o2.arrays["uflx"][sl]))[:, iu]
    assert np.abs(du - (big - interior(c2.state["uflx"][sl])[:, iu])).max() <= 1e-12 * big
    assert np.array_equal(interior(o.arrays["vflx"][sl])[:, iv], interior(o2.arrays["vflx"][sl])[:, iv])


def test_zero_velocity_is_identity():
    def mutate(c):
        for nm in ("u", "v", "umfltd", "vmfltd", "ubflxs_p", "vbflxs_p", "umflsm", "vmflsm"):
            c.state[nm][:] = 0.0
    c, o = run("tiny1", ntr=0, mutate=mutate)
    nn, mm, kk = c.levels[3], c.levels[2], c.dims[2]
    o.advect(*c.levels)
    d0 = interior(c.state["dp"][nn:nn + kk]); d1 = interior(o.arrays["dp"][nn:nn + kk])
    assert np.abs(d1 - d0).max() <= 1e-9
    assert np.abs(interior(o.arrays["vflx"][mm:mm + kk])).max() == 0.0
