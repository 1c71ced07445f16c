"""fuk95 (the reference's idealized stand-alone test, fuk95/mod_fuk95.F90) as an analytic case:
geometry and initial conditions restated in blom_b200/fuk95.py, and the physical known answer the
set-up is built around - a density front at rest that adjusts geostrophically to a jet of maximum
speed u0 = 0.3 m/s - recovered by the ORACLE running only the hot-path routines."""
import numpy as np

from blom_b200 import fuk95
from blom_b200.driver import run_step, step_routines
from blom_b200.lib import time_levels
from util import Case, interior

# library defaults (layer diffusion inside diffus): with ltedtp='neutral' the diffused scalars would only reach
# T,S through the out-of-scope ALE step
ROUTINES = step_routines({})


def test_geometry_and_initial_state():
    c = Case("fuk95_analytic", ntr=1, nstep=1)
    assert c.dims == (156, 32, 12, 4) and c.nreg == 4
    ip = interior(c.masks["ip"])
    assert ip[:, 0].sum() == 0 and ip[:, -1].sum() == 0 and ip[:, 1:-1].all()      # walls at i=1, itdm
    for nm in ("scpx", "scpy", "scux", "scvy", "scqx"):
        assert np.all(interior(c.grid[nm]) == 20.8e3 / 32)                            # gs = lambda/jtdm
    assert np.all(interior(c.grid["corioq"]) == 1.0e-4)
    assert np.all(interior(c.grid["depths"])[0][:, 1:-1] == 200.0)
    sg = c.syn.initial_sigma()
    assert np.all(np.diff(sg, axis=0) >= 0) and np.all(np.diff(sg[:7], axis=0) > 0)   # stable; uniform below h1
    assert abs(sg.max() - 27.0) < 1e-12 and 25.4 < sg.min() < 25.6                   # rhob-rho0 below h1
    # tofsig inverts sig (phy/mod_eos.F90:191-203, :346-364) at the namelist's pref = 0
    th = fuk95.tofsig(sg, fuk95.SALN0, 0.0)
    assert np.abs(fuk95.sig(th, fuk95.SALN0, 0.0) - sg).max() < 1e-12
    # the front: lighter water on the low-i side, the meander displaces it by +-0.1 grid cells
    assert np.all(sg[0, :, 40] < sg[0, :, 115])
    assert np.ptp(sg[0, :, 78]) > 0
    assert not interior(c.state["u"]).any() and not interior(c.state["v"]).any()
    kk = 12
    assert np.allclose(interior(c.state["dp"])[:kk].sum(axis=0)[:, 1:-1], 200.0 * 9806.0)


def test_geostrophic_adjustment_known_answer():
    """160 steps of 180 s are half an inertial period (pi/f = 8.7 h): the along-channel velocity
    must have grown to the geostrophic jet speed u0 that the density front was constructed for
    (fuk95/mod_fuk95.F90:47,379-385), mass and salt must be conserved to round-off."""
    c = Case("fuk95_analytic", ntr=1, nstep=1)
    o = c.new_oracle()
    o.inieos(); o.numerical_bounds(); o.init_cppm()
    kk = c.dims[2]
    scp2 = interior(o.arrays["scp2"])[0]

    def inventory(nm, nn):
        dp = interior(o.arrays["dp"])[nn:nn + kk]
        a = 1.0 if nm is None else interior(o.arrays[nm])[nn:nn + kk]
        return float((a * dp * scp2).sum())

    m0 = s0 = None
    vmax = []
    for nstep in range(1, 161):
        m, n, mm, nn, k1m, k1n = time_levels(nstep, kk)
        o.set_scalar("nstep", nstep)
        run_step(o, ROUTINES, (m, n, mm, nn, k1m, k1n))
        if nstep == 1:
            m0, s0 = inventory(None, nn), inventory("saln", nn)
        vmax.append(np.abs(interior(o.arrays["v"])).max())
    assert np.isfinite(interior(o.arrays["dp"])).all() and interior(o.arrays["dp"]).min() >= 0.0
    assert abs(inventory(None, nn) / m0 - 1.0) < 1e-14
    assert abs(inventory("saln", nn) / s0 - 1.0) < 1e-14
    assert vmax[9] < 0.01 < vmax[79]                       # starts at rest, spins up
    assert 0.9 * fuk95.U0 < max(vmax) < 1.15 * fuk95.U0    # 0.318 m/s at step ~160
    assert np.abs(interior(o.arrays["u"])).max() < fuk95.U0


def test_level_isopycnals_stay_at_rest():
    """The same channel without the front (u0 = 0: horizontally uniform stratification, no forcing):
    the pressure gradient force must vanish identically, so the state of rest is a fixed point of
    pgforc + momtum + barotp and the thickness/tracer fields must not move."""
    c = Case("fuk95_analytic", ntr=1, nstep=1, u0=0.0)
    o = c.new_oracle()
    o.inieos(); o.numerical_bounds(); o.init_cppm()
    kk = c.dims[2]
    dp0 = interior(o.arrays["dp"]).copy()
    t0 = interior(o.arrays["temp"]).copy()
    for nstep in range(1, 11):
        m, n, mm, nn, k1m, k1n = time_levels(nstep, kk)
        o.set_scalar("nstep", nstep)
        run_step(o, ROUTINES, (m, n, mm, nn, k1m, k1n))
    for nm in ("u", "v", "pgfx", "pgfy", "ubflxs_p", "vbflxs_p"):
        assert np.abs(interior(o.arrays[nm])).max() < 1e-12, nm
    assert np.abs(interior(o.arrays["dp"]) - dp0).max() < 1e-6          # Pa, of 1.6e5
    assert np.abs(interior(o.arrays["temp"]) - t0).max() < 1e-12
