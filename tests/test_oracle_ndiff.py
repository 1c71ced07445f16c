"""Pins the oracle's neutral-diffusion restatement (phy/mod_ndiff.F90) on CPU through the
scheme's own guarantees (the reference ships no golden outputs):
 - antisymmetric face fluxes: thickness-weighted inventories of T, S and tracers are conserved to
   round-off;
 - a horizontally and vertically uniform scalar is left untouched, bit for bit;
 - diffusion is down-gradient along neutral layers: the thickness-weighted variance of a passive
   tracer does not grow;
 - analytic neutral slope: with S uniform and T a function of (p - i*delta) only, isotherms are
   neutral surfaces, so every neutral interface is displaced by delta between neighbouring columns
   and nslpx = -alpha0*scuxi/grav*delta (mod_ndiff.F90:268-269, :1064) however the layers are cut;
   the meridional slope vanishes and so do all fluxes between columns of equal T on a neutral layer."""
import numpy as np
import pytest

from util import Case, interior
from blom_b200 import synth

ONEM = 9806.0


def prepared(cfg, ntr=1, mutate=None, **opts):
    c = Case(cfg, ntr=ntr)
    if mutate:
        mutate(c)
    nd = {k: v.copy() for k, v in synth.ndiff_inputs(c.syn, c.state, c.levels, ntr=ntr).items()}
    o = c.new_oracle(); o.inieos()
    o.register_all(nd)
    for k, v in opts.items():
        o.set_option(k, v)
    o.pgforc(*c.levels)   # pu, pv: the interface pressures at the faces the fluxes are binned on
    return c, o, nd


def inventory(c, nd, trm, nt, rows=slice(None), power=1):
    kk = c.dims[2]
    dpd = np.maximum(np.diff(nd["nd_p_dst"], axis=0), 1e-5)  # dp_eps floor of the update (:1168)
    f = trm[nt * kk:(nt + 1) * kk] ** power * dpd * c.grid["scp2"][0]
    return interior(f)[:, rows].sum()


@pytest.mark.parametrize("cfg", ["tiny1", "tiny3", "tiny4", "fuk95"])
@pytest.mark.parametrize("align", ["1", "0"])
def test_conservation_and_variance(cfg, align):
    c, o, nd = prepared(cfg, ndiff_surface_align=align)
    before = nd["nd_trc_rm"].copy()
    o.ndiff(*c.levels)
    after = nd["nd_trc_rm"]
    kk = c.dims[2]
    assert np.abs(after - before).max() > 0.0
    for nt in range(3):
        i0, i1 = inventory(c, nd, before, nt), inventory(c, nd, after, nt)
        assert abs(i1 - i0) <= 1e-14 * abs(i0), (nt, i0, i1)
    v0, v1 = inventory(c, nd, before, 2, power=2), inventory(c, nd, after, 2, power=2)
    assert v1 <= v0 * (1 + 1e-14)
    mm = c.levels[2]
    assert np.abs(interior(o.arrays["utflld"][mm:mm + kk])).max() > 0.0
    assert np.abs(interior(o.arrays["nslpy"])).max() > 0.0


def test_uniform_scalar_untouched():
    def mutate(c):
        c.state["trc"][:] = 2.5
    c, o, nd = prepared("tiny3", mutate=mutate)
    kk = c.dims[2]
    before = nd["nd_trc_rm"].copy()
    o.ndiff(*c.levels)
    ip = interior(c.masks["ip"]) == 1
    assert np.array_equal(interior(nd["nd_trc_rm"][2 * kk:])[:, ip], interior(before[2 * kk:])[:, ip])


def tilted_case(delta):
    """fuk95-sized domain (12 layers) without land; T = f(p - i*delta), S uniform; each column cut into
    layers differently (the source interfaces are jittered) with exact linear reconstructions."""
    c = Case("fuk95", ntr=0, land=False)
    kk = c.dims[2]
    ldj, ldi = c.syn.ldj, c.syn.ldi
    ig = (np.arange(ldi) - 4)[None, :] * np.ones((ldj, 1))
    pbot = 4000.0 * ONEM
    jit = 0.3 * (c.syn._hash_uniform_padded(kk + 1, 777) - 0.5)
    frac = (np.arange(kk + 1)[:, None, None] + np.where((np.arange(kk + 1) % kk == 0)[:, None, None], 0.0, jit)) / kk
    p_src = frac * pbot
    # (kept within -3..19 degC: colder water would get lighter on cooling and the column unstable)
    tfun = lambda p: 12.0 - 8.0 * (p - (ig[None] - c.dims[0] / 2) * delta) / pbot
    T = 2
    tsd = np.zeros((2 * kk * T, ldj, ldi)); tpc = np.zeros((5 * kk * T, ldj, ldi)); trm = np.zeros((kk * T, ldj, ldi))
    for k in range(kk):
        top, bot = tfun(p_src[k])[0], tfun(p_src[k + 1])[0]
        tsd[k * 2], tsd[k * 2 + 1] = top, bot
        tpc[k * 5], tpc[k * 5 + 1] = top, bot - top
        trm[k] = 0.5 * (top + bot)
        tsd[(kk + k) * 2] = tsd[(kk + k) * 2 + 1] = tpc[(kk + k) * 5] = trm[kk + k] = 35.0
    nn = c.levels[3]
    c.state["dp"][nn:nn + kk] = np.diff(p_src, axis=0)
    c.state["temp"][nn:nn + kk] = trm[:kk]; c.state["saln"][nn:nn + kk] = 35.0
    nd = {"nd_p_src": p_src, "nd_p_dst": p_src.copy(), "nd_ksmx": np.full((1, ldj, ldi), kk, np.int32),
          "nd_t_srcdi": tsd, "nd_tpc_src": tpc, "nd_trc_rm": trm, "dpml": np.full((1, ldj, ldi), 10.0 * ONEM)}
    o = c.new_oracle(); o.inieos()
    o.register_all(nd)
    o.set_option("ndiff_surface_align", "0")
    return c, o, nd


@pytest.mark.parametrize("delta_m", [0.0, 15.0, -40.0])
def test_analytic_neutral_slope(delta_m):
    delta = delta_m * ONEM
    c, o, nd = tilted_case(delta)
    kk = c.dims[2]
    before = nd["nd_trc_rm"].copy()
    o.ndiff(*c.levels)
    a = o.arrays
    expect = -1.0e-3 * interior(a["scuxi"][0]) / 9.806 * delta   # -alpha0*scuxi/grav*delta
    got = interior(a["nslpx"])
    # interfaces 2..kk-1: away from the surface/bottom where the displaced partner leaves the column;
    # skip the two columns next to the periodic seam, where i jumps by -itdm
    scale = np.abs(expect).max()
    # interfaces whose density difference is below rho_eps = 1e-5 count as neutral (:224-226): with
    # d(rho)/dp = drhodt*dT/dp ~ 4e-8 kg m-3 Pa-1 here (less in cold water) that is a pressure slack
    # of a few hundred Pa
    atol = 1.0e-3 * interior(a["scuxi"][0]).max() / 9.806 * 1000.0
    assert np.abs(got[1:kk - 1, :, 2:-2] - expect[None, :, 2:-2]).max() <= 2e-3 * scale + atol
    assert np.abs(interior(a["nslpy"])[1:kk - 1, :, 2:-2]).max() <= 2e-3 * scale + atol
    if delta_m != 0.0:
        assert scale > 50 * atol   # the analytic slope is far above the slack: the check has teeth
    # T is constant along every neutral layer: no flux, fields unchanged to the root-finder tolerance
    assert np.abs(interior(nd["nd_trc_rm"][:kk] - before[:kk])[:, :, 2:-2]).max() <= 1e-7
