"""GPU parity of neutral diffusion (phy/mod_ndiff.F90:959-1175 driven in the order of
phy/mod_ale_regrid_remap.F90:1607-1690) against the oracle on identical synthetic ALE products,
through the C ABI (blomgpu_ndiff).

Tolerances (float64): the face fields (u|v t|s flld, u|v t|s flx, nslpx, nslpy) follow the reference's
operation order exactly, so the parity build (-fmad=false) must match to <= 1e-13 of the field's
max-norm (in practice bit for bit); the updated tracers trc_rm sum each face's contributions to one
destination layer before adding them to the cell (the one documented reassociation, ndiff.cu header):
<= 1e-13 of the field's max-norm.  Performance build (FMA contraction): <= 1e-9 — the neutral-interface
search is a discrete decision tree, so this bound also asserts that no column took another branch."""
import numpy as np
import pytest

from util import Case, interior, max_rel_err
from blom_b200 import synth

pytestmark = pytest.mark.gpu

FACE = ["utflld", "usflld", "vtflld", "vsflld", "utflx", "usflx", "vtflx", "vsflx", "nslpx", "nslpy"]


def run_pair(cfg, ntr, nstep, parity, align="1"):
    c = Case(cfg, ntr=ntr, nstep=nstep)
    nd_o = {k: v.copy() for k, v in synth.ndiff_inputs(c.syn, c.state, c.levels, ntr=ntr).items()}
    nd_g = {k: v.copy() for k, v in nd_o.items()}
    o = c.new_oracle(); g = c.new_gpu(parity=parity)
    o.register_all(nd_o); g.register_all(nd_g)
    for b in (o, g):
        b.inieos()
        b.set_option("ndiff_surface_align", align)
        b.pgforc(*c.levels)      # pu, pv: the face interface pressures the fluxes are binned on
        b.ndiff(*c.levels)
    g.download_all()
    return c, o, g, nd_o, nd_g


@pytest.mark.parametrize("cfg", ["tiny0", "tiny1", "tiny2", "tiny3", "tiny4", "fuk95"])
@pytest.mark.parametrize("align", ["1", "0"])
@pytest.mark.parametrize("ntr", [0, 1, 2])
def test_ndiff_parity_build(cfg, align, ntr):
    c, o, g, nd_o, nd_g = run_pair(cfg, ntr, 1, True, align)
    try:
        before = synth.ndiff_inputs(c.syn, c.state, c.levels, ntr=ntr)["nd_trc_rm"]
        assert np.abs(interior(nd_o["nd_trc_rm"]) - interior(before)).max() > 0.0
        for nm in FACE:
            err = max_rel_err(interior(g.arrays[nm]), interior(o.arrays[nm]))
            assert err <= 1e-13, (nm, err)
        err = max_rel_err(interior(nd_g["nd_trc_rm"]), interior(nd_o["nd_trc_rm"]))
        assert err <= 1e-13, ("nd_trc_rm", err)
        # the change itself (not only the field) must agree: relative to the largest increment
        inc_o = interior(nd_o["nd_trc_rm"]) - interior(before)
        inc_g = interior(nd_g["nd_trc_rm"]) - interior(before)
        assert max_rel_err(inc_g, inc_o) <= 1e-10
    finally:
        g.finalize()


@pytest.mark.parametrize("nstep", [1, 2])
def test_ndiff_time_levels(nstep):
    c, o, g, nd_o, nd_g = run_pair("tiny3", 1, nstep, True)
    try:
        for nm in FACE:
            assert max_rel_err(interior(g.arrays[nm]), interior(o.arrays[nm])) <= 1e-13, nm
        assert max_rel_err(interior(nd_g["nd_trc_rm"]), interior(nd_o["nd_trc_rm"])) <= 1e-13
    finally:
        g.finalize()


@pytest.mark.parametrize("cfg", ["tiny2", "fuk95"])
def test_ndiff_perf_build(cfg):
    c, o, g, nd_o, nd_g = run_pair(cfg, 1, 1, False)
    try:
        for nm in FACE:
            err = max_rel_err(interior(g.arrays[nm]), interior(o.arrays[nm]))
            assert err <= 1e-9, (nm, err)
        assert max_rel_err(interior(nd_g["nd_trc_rm"]), interior(nd_o["nd_trc_rm"])) <= 1e-9
    finally:
        g.finalize()


def test_ndiff_conserves_on_device():
    """Thickness-weighted inventories through the device path alone: antisymmetric face fluxes."""
    c = Case("tiny3", ntr=1, nstep=1)
    nd = {k: v.copy() for k, v in synth.ndiff_inputs(c.syn, c.state, c.levels, ntr=1).items()}
    g = c.new_gpu(parity=False)
    try:
        g.register_all(nd)
        g.inieos(); g.pgforc(*c.levels)
        kk = c.dims[2]
        dpd = np.maximum(np.diff(nd["nd_p_dst"], axis=0), 1e-5)

        def inv(nt):
            return interior(nd["nd_trc_rm"][nt * kk:(nt + 1) * kk] * dpd * c.grid["scp2"][0]).sum()
        before = [inv(nt) for nt in range(3)]
        g.ndiff(*c.levels); g.download_all()
        for nt in range(3):
            assert abs(inv(nt) - before[nt]) <= 1e-14 * abs(before[nt]), nt
    finally:
        g.finalize()
