"""Pins the oracle's restatement of cmnfld2's hybrid branch (phy/mod_cmnfld_routines.F90:229-350,
:654-883, :1158-1238) on CPU.  The reference ships no golden outputs, so the pins are
independent restatements and analytic limits:
 - bfsqf solves (I - d/dp sls0^2 d/dp) bfsqf = max(bfsqmn, bfsq): compared column by column with a dense
   numpy solve of the same tridiagonal system built from p, T, S with the scalar EOS, and with the
   flux-form identity sum(delp*bfsqf) = sum(delp*bfsq);
 - a horizontally uniform state has exactly zero neutral slope;
 - tilted isotherms: with S uniform and T = F(p - a*i), isotherms are neutral surfaces, so the slope in
   height coordinates is the finite-difference slope of an isotherm's geopotential height between
   neighbouring columns; the reference converts the pressure displacement with rho0 instead of the
   in-situ density (g*rho_x/(rho0*N2), :727), so the ratio to the exact slope must be rho/rho0 (+-0.5 %);
 - on a flat bottom with every layer massive, cmnfld_nnslope_ale (slope known) reproduces the
   slope x N product of cmnfld_nslope_ale bit for bit."""
import numpy as np
import pytest

from util import Case, interior
from blom_b200 import synth

ONEM, ONEMM, GRAV = 9806.0, 9.806, 9.806
BFSQMN, SLS0 = 1.0e-7, 10.0 * ONEM


def prepared(cfg="tiny2", **kw):
    c = Case(cfg, **kw)
    extra = synth.cmnfld_arrays(c.syn)
    o = c.new_oracle(); o.inieos()
    o.register_all(extra)
    return c, o, extra


def column_reference(o, p, t, s, dp1):
    """numpy/dense restatement of one column of cmnfld_bfsqf_ale (p: kk+1 interfaces; t, s: kk layers)"""
    kk = len(t)
    delp = np.zeros(kk); bfsq = np.zeros(kk); bi = np.zeros(kk + 1)
    bi[0] = BFSQMN
    pup, tup, sup = 0.5 * (p[0] + p[1]), t[0], s[0]
    for k in range(1, kk):
        if p[kk] - p[k] < 1e-12:
            delp[k], bi[k], bfsq[k] = ONEMM, bi[k - 1], BFSQMN
            continue
        plo = p[kk] if p[kk] - p[k + 1] < 1e-12 else 0.5 * (p[k] + p[k + 1])
        delp[k] = max(ONEMM, plo - pup)
        b = GRAV * GRAV * (o.eos("rho", p[k], t[k], s[k]) - o.eos("rho", p[k], tup, sup)) / delp[k]
        bfsq[k] = max(BFSQMN, b)
        b = b * delp[k] / max(ONEM, delp[k])
        bi[k] = bi[k - 1] if p[kk] - p[k] < ONEM else b
        pup, tup, sup = plo, t[k], s[k]
    delp[0] = dp1
    bi[0] = bi[1]
    bfsq[0] = max(BFSQMN, bi[0])
    bi[kk] = bi[kk - 1]
    A = np.zeros((kk, kk))
    for k in range(kk):
        a = -2 * SLS0 ** 2 / (delp[k] * (delp[k - 1] + delp[k])) if k > 0 else 0.0
        cc = -2 * SLS0 ** 2 / (delp[k] * (delp[k] + delp[k + 1])) if k < kk - 1 else 0.0
        if k > 0:
            A[k, k - 1] = a
        if k < kk - 1:
            A[k, k + 1] = cc
        A[k, k] = 1.0 - a - cc
    return bi, np.linalg.solve(A, bfsq), delp, bfsq


@pytest.mark.parametrize("cfg", ["tiny0", "tiny2", "fuk95"])
def test_bfsqf_matches_dense_solve(cfg):
    c, o, ex = prepared(cfg)
    m, n, mm, nn, k1m, k1n = c.levels
    kk = c.dims[2]
    o.cmnfld_bfsqf_ale(*c.levels)
    a = o.arrays
    ip = c.masks["ip"]
    pts = np.argwhere(ip == 1)
    rng = np.random.default_rng(3)
    checked = 0
    for jj_, ii_ in pts[rng.choice(len(pts), size=min(40, len(pts)), replace=False)]:
        if not (3 <= jj_ < ip.shape[0] - 3 and 3 <= ii_ < ip.shape[1] - 3):
            continue  # cmnfld_bfsqf_ale covers -1..ii+2 only
        p = a["p"][:, jj_, ii_]; t = a["temp"][nn:nn + kk, jj_, ii_]; s = a["saln"][nn:nn + kk, jj_, ii_]
        bi, f, delp, bfsq = column_reference(o, p, t, s, a["dp"][nn, jj_, ii_])
        np.testing.assert_allclose(ex["bfsqi"][:, jj_, ii_], bi, rtol=1e-12, atol=1e-30)
        # the system is stiff (sls0^2/delp^2 up to 1e8 for millimetre layers): dense LU and Thomas agree to 1e-6
        np.testing.assert_allclose(ex["bfsqf"][:kk, jj_, ii_], f, rtol=1e-6)
        assert ex["bfsqf"][kk, jj_, ii_] == ex["bfsqf"][kk - 1, jj_, ii_]
        lay = 0.5 * (bi[:-2] + bi[1:-1]); lay = np.append(lay, bi[kk - 1])
        np.testing.assert_allclose(ex["bfsql"][:, jj_, ii_], lay, rtol=1e-12, atol=1e-30)
        # flux form: the filter redistributes, it does not create buoyancy frequency
        assert abs((delp * ex["bfsqf"][:kk, jj_, ii_]).sum() - (delp * bfsq).sum()) <= 1e-9 * (delp * bfsq).sum()
        assert ex["bfsqf"][:kk, jj_, ii_].min() >= bfsq.min() * (1 - 1e-12)
        assert ex["bfsqf"][:kk, jj_, ii_].max() <= bfsq.max() * (1 + 1e-12)
        checked += 1
    assert checked >= 5
    # land and the outermost halo ring stay at the zero fill (:247-248)
    assert np.all(ex["bfsqi"][:, ip != 1] == 0.0) and np.all(ex["bfsql"][:, 0, :] == 0.0)


def flat_case(cfg="tiny1", tilt=0.0):
    """all-ocean-interior case overwritten with a horizontally uniform (or uniformly tilted) state"""
    c = Case(cfg, land=False, metric="uniform")
    kk = c.dims[2]
    st, gr = c.state, c.grid
    m, n, mm, nn, k1m, k1n = c.levels
    D = 50.0 * ONEM
    ldj, ldi = st["p"].shape[1:]
    icol = np.arange(ldi)[None, :] * np.ones((ldj, 1))
    for k in range(kk + 1):
        st["p"][k] = k * D
    for k in range(kk):
        st["dp"][k + nn] = D
        st["dp"][k + mm] = D
        pmid = (k + 0.5) * D
        st["temp"][k + nn] = 20.0 - 15.0 * (pmid - tilt * icol) / (kk * D)
        st["saln"][k + nn] = 35.0
    st["phi"][kk] = -GRAV * kk * 50.0
    return c


def test_uniform_state_has_zero_slope():
    c = flat_case()
    ex = synth.cmnfld_arrays(c.syn)
    o = c.new_oracle(); o.inieos(); o.register_all(ex)
    o.cmnfld_bfsqf_ale(*c.levels)
    o.arrays["nslpx"][:] = 7.0; o.arrays["nslpy"][:] = 7.0
    o.cmnfld_nslope_ale(*c.levels)
    iu, iv = c.masks["iu"], c.masks["iv"]
    assert np.all(interior(o.arrays["nslpx"])[:, interior(iu) == 1] == 0.0)
    assert np.all(interior(o.arrays["nslpy"])[:, interior(iv) == 1] == 0.0)
    assert np.all(interior(ex["nnslpx"]) == 0.0) and np.all(interior(ex["nnslpy"]) == 0.0)


def test_tilted_isotherms_give_the_isotherm_slope():
    a_tilt = 2.0 * ONEM       # isotherms deepen by 2 m of pressure per grid cell
    c = flat_case("fuk95", tilt=a_tilt)
    ex = synth.cmnfld_arrays(c.syn)
    o = c.new_oracle(); o.inieos(); o.register_all(ex)
    kk = c.dims[2]
    nn = c.levels[3]
    o.cmnfld_bfsqf_ale(*c.levels)
    o.cmnfld_nslope_ale(*c.levels)
    a = o.arrays
    j, i = 8, 12
    assert c.masks["iu"][j, i] == 1
    dx = 1.0 / c.grid["scuxi"][0, j, i]
    for k in range(2, kk - 2):      # interfaces away from the surface and bottom half layers
        # height of the isotherm that crosses interface k in column i-1, found in column i by linear
        # interpolation of phi in pressure (it sits a_tilt deeper there)
        z_l = a["phi"][k, j, i - 1] / GRAV
        pk = a["p"][k, j, i]
        z_r = (a["phi"][k, j, i] + (a["phi"][k + 1, j, i] - a["phi"][k, j, i]) * a_tilt /
               (a["p"][k + 1, j, i] - pk)) / GRAV
        expect = (z_r - z_l) / dx
        got = a["nslpx"][k, j, i]
        rho = o.eos("rho", pk, a["temp"][k + nn, j, i], a["saln"][k + nn, j, i])
        assert got < 0.0 and abs(got / expect - rho / 1000.0) <= 5e-3, (k, got, expect, rho)
    assert np.abs(interior(a["nslpy"])).max() <= 1e-12 * np.abs(interior(a["nslpx"])).max()


def test_nnslope_reproduces_nslope_product_on_flat_bottom():
    c = flat_case("tiny1", tilt=1.0 * ONEM)
    ex = synth.cmnfld_arrays(c.syn)
    o = c.new_oracle(); o.inieos(); o.register_all(ex)
    o.cmnfld_bfsqf_ale(*c.levels)
    o.cmnfld_nslope_ale(*c.levels)
    ref_x, ref_y = ex["nnslpx"].copy(), ex["nnslpy"].copy()
    assert np.abs(interior(ref_x)).max() > 0.0
    ex["nnslpx"][:] = -1.0; ex["nnslpy"][:] = -1.0
    o.cmnfld_nnslope_ale(*c.levels)
    iu, iv = interior(c.masks["iu"]) == 1, interior(c.masks["iv"]) == 1
    assert np.array_equal(interior(ex["nnslpx"])[:, iu], interior(ref_x)[:, iu])
    assert np.array_equal(interior(ex["nnslpy"])[:, iv], interior(ref_y)[:, iv])


def test_cmnfld2_options():
    c, o, ex = prepared("tiny2")
    o.cmnfld2(*c.levels)
    nsl = o.arrays["nslpx"].copy()
    assert np.abs(interior(nsl)).max() > 0.0 and np.isfinite(nsl).all()
    # ltedtp='neutral': the slope is an input (ndiff produced it) and only slope x N is rebuilt
    c2, o2, ex2 = prepared("tiny2")
    o2.set_option("ltedtp", "neutral")
    before = o2.arrays["nslpx"].copy()
    o2.cmnfld2(*c2.levels)
    # (row jj of a vector field is rewritten by the tripolar fold of xctilr, phy/mod_xc.F90:4275-4358)
    assert np.array_equal(interior(o2.arrays["nslpx"])[:, :-1], interior(before)[:, :-1])
    assert np.abs(interior(ex2["nnslpx"])).max() > 0.0
    c3, o3, ex3 = prepared("tiny2")
    o3.set_option("vcoord", "isopyc_bulkml")
    with pytest.raises(Exception, match="unsupported"):
        o3.cmnfld2(*c3.levels)
