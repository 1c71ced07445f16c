"""GPU parity of the chained hot path: the routines run back to back on device-resident
state for several steps (time levels swapping) and every registered array is compared with
the oracle after each routine.  Tolerance: 1e-10 relative to the field max-norm after N
chained steps, parity build."""
import numpy as np
import pytest

from blom_b200.driver import run_step
from blom_b200.lib import time_levels
from util import Case, compare_all, interior, max_rel_err, prepare_step

pytestmark = pytest.mark.gpu

SKIP = {"depths"}


# the reference's option set for the hybrid coordinate (neutral diffusion through ndiff, dluc, bod23) and
# the layer-diffusion / uc / fox08 set (the isopycnic defaults, which exercise diffus' flux branch)
OPTION_SETS = {"reference": None, "layer": {"ltedtp": "layer", "bmcmth": "uc", "mlrmth": "fox08"}}


@pytest.mark.parametrize("optset", ["reference", "layer"])
@pytest.mark.parametrize("cfg", ["tiny0", "tiny1", "tiny2", "tiny3", "tiny4", "fuk95", "fuk95_analytic"])
def test_chained_steps(cfg, optset):
    c = Case(cfg, ntr=1, nstep=1)
    o = c.new_oracle(); g = c.new_gpu(parity=True)
    try:
        routines, _ = prepare_step(c, (o, g), OPTION_SETS[optset])
        kk = c.dims[2]
        for nstep in (1, 2, 3):
            lv = time_levels(nstep, kk)
            for b in (o, g):
                b.set_scalar("nstep", nstep)
            for r in routines:
                run_step(o, [r], lv); run_step(g, [r], lv)
                compare_all(g, o, (cfg, nstep, r))
        assert np.isfinite(g.arrays["dp"]).all()
    finally:
        g.finalize()


@pytest.mark.parametrize("ntr", [0, 1])
def test_pipelined_step_matches_plain_step(ntr):
    """HotPath.step_pipelined (asynchronous upload of the new time level, u,v halo refresh moved to its
    first reader, downloads started after each field's last writer) must leave exactly the host arrays
    that upload / step / download leave, and so must the early-download variant of step()."""
    from blom_b200.driver import HotPath, IO_FIELDS
    res = []
    for mode in ("plain", "early", "pipelined"):
        hp = HotPath("tiny2", ntr=ntr, nstep=1, parity=True)
        try:
            for _ in range(3):
                if mode == "pipelined":
                    hp.step_pipelined()
                    hp.set_step(hp.nstep + 1)
                else:
                    hp.upload_inputs()
                    hp.advance(early_download=(mode == "early"))
                    hp.download_outputs()
            res.append({nm: hp.arrays[nm].copy() for nm in IO_FIELDS if nm in hp.arrays})
        finally:
            hp.finalize()
    for other in res[1:]:
        for nm in res[0]:
            assert np.array_equal(res[0][nm], other[nm]), nm
    assert np.abs(res[0]["u"]).max() > 0


def test_fuk95_geostrophic_adjustment_on_gpu():
    """The analytic fuk95 front (blom_b200/fuk95.py) stepped by the CUDA hot path alone: after half
    an inertial period the along-channel jet has the speed u0 the front was built for, mass is
    conserved to round-off, and the state after 160 steps stays close to the oracle's (tolerance 1e-6
    of the field maxima: 160 chained steps of a developing instability amplify rounding differences;
    the 3-step chain above holds 1e-10)."""
    from blom_b200 import fuk95
    c = Case("fuk95_analytic", ntr=1, nstep=1)
    o = c.new_oracle(); g = c.new_gpu(parity=True)
    try:
        kk = c.dims[2]
        # layer diffusion inside diffus: with ltedtp='neutral' the diffused scalars would only reach T,S
        # through the out-of-scope ALE step, and the undiffused front overshoots u0
        routines, _ = prepare_step(c, (o, g), {"ltedtp": "layer"})
        for nstep in range(1, 161):
            lv = time_levels(nstep, kk)
            for b in (o, g):
                b.set_scalar("nstep", nstep)
                run_step(b, routines, lv)
        g.download_all()
        nn = lv[3]
        scp2 = interior(g.arrays["scp2"])[0]
        mass = float((interior(g.arrays["dp"])[nn:nn + kk] * scp2).sum())
        mass0 = float((interior(c.state["dp"])[:kk] * scp2).sum())
        assert abs(mass / mass0 - 1.0) < 1e-13
        vmax = np.abs(interior(g.arrays["v"])).max()
        assert 0.9 * fuk95.U0 < vmax < 1.15 * fuk95.U0, vmax
        for nm in ("dp", "temp", "saln", "u", "v", "pb"):
            err = max_rel_err(interior(g.arrays[nm]), interior(o.arrays[nm]))
            assert err <= 1e-6, (nm, err)
    finally:
        g.finalize()


@pytest.mark.parametrize("cfg", ["tiny2", "tiny4"])
def test_chained_step_two_passive_tracers(cfg):
    """ntr = 2 (the largest tracer count the transport kernels are instantiated for): one chained
    step, every registered array against the oracle after every routine, parity build, 1e-10."""
    c = Case(cfg, ntr=2, nstep=1)
    o = c.new_oracle(); g = c.new_gpu(parity=True)
    try:
        routines, _ = prepare_step(c, (o, g))
        lv = time_levels(1, c.dims[2])
        for r in routines:
            run_step(o, [r], lv); run_step(g, [r], lv)
            compare_all(g, o, (cfg, r))
        assert np.abs(interior(g.arrays["trc"])).max() > 0
    finally:
        g.finalize()


@pytest.mark.parametrize("cfg,optset", [("tiny2", "reference"), ("tiny4", "layer"), ("tiny1", "layer")])
def test_chained_steps_five_passive_tracers(cfg, optset):
    """ntr = 5, more than any kernel carries in one launch (cppm: T, S + 2 tracers, then groups of 2; pbcor and
    diffus: groups of 4; the reference loops nt = 1..ntr freely, phy/mod_cppm.F90:1599-1618): two chained steps,
    every registered array against the oracle after every routine, parity build, 1e-10."""
    c = Case(cfg, ntr=5, nstep=1)
    o = c.new_oracle(); g = c.new_gpu(parity=True)
    try:
        routines, _ = prepare_step(c, (o, g), OPTION_SETS[optset])
        kk = c.dims[2]
        for nstep in (1, 2):
            lv = time_levels(nstep, kk)
            for b in (o, g):
                b.set_scalar("nstep", nstep)
            for r in routines:
                run_step(o, [r], lv); run_step(g, [r], lv)
                compare_all(g, o, (cfg, nstep, r))
        trc = interior(g.arrays["trc"])
        assert np.abs(trc).max() > 0 and np.isfinite(trc).all()
        # the five tracers carry different fields (the generator offsets them), so a mixed-up group would show
        assert not np.array_equal(trc[:2 * kk], trc[8 * kk:10 * kk])
    finally:
        g.finalize()
