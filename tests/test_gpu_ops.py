"""GPU parity (through the C ABI) of diffus, tmsmt1/2, pgforc and barotp against the
oracle on identical seeded inputs.

Tolerances (float64, stated per routine): parity build (-fmad=false) <= 1e-13 relative
to the field's max-norm for the streaming routines, <= 1e-12 for pgforc (CUDA exp/log-free
but long recurrences) and <= 1e-11 for barotp (125+ chained substeps, one libm exp);
performance build (FMA contraction) 1e-10."""
import numpy as np
import pytest

from util import Case, interior, max_rel_err

pytestmark = pytest.mark.gpu


def pair(cfg, ntr=1, nstep=1, parity=True, opts=None):
    c = Case(cfg, ntr=ntr, nstep=nstep)
    o = c.new_oracle(); g = c.new_gpu(parity=parity)
    for k, v in (opts or {}).items():
        o.set_option(k, v); g.set_option(k, v)
    o.inieos(); g.inieos()
    return c, o, g


def check(g, o, names, tol, halo=0):
    g.download_all()
    for nm in names:
        err = max_rel_err(interior(g.arrays[nm], halo=halo), interior(o.arrays[nm], halo=halo))
        assert err <= tol, (nm, err)


@pytest.mark.parametrize("cfg", ["tiny0", "tiny1", "tiny2", "tiny3", "tiny4"])
@pytest.mark.parametrize("parity", [True, False])
def test_diffus(cfg, parity):
    c, o, g = pair(cfg, parity=parity)
    try:
        o.diffus(*c.levels); g.diffus(*c.levels)
        check(g, o, ["temp", "saln", "trc", "sigma", "usflld", "utflld", "vsflld", "vtflld", "usflx", "utflx",
                     "vsflx", "vtflx"], 1e-13 if parity else 1e-11, halo=1)
        assert np.abs(interior(g.arrays["temp"]) - interior(c.state["temp"])).max() > 1e-6
    finally:
        g.finalize()


def test_diffus_neutral_and_bad_option():
    from blom_b200.lib import BlomGpuError
    c, o, g = pair("tiny1", opts={"ltedtp": "neutral"})
    try:
        o.diffus(*c.levels); g.diffus(*c.levels)
        check(g, o, ["temp", "saln", "dp"], 0.0, halo=1)
        g.set_option("ltedtp", "bogus")
        with pytest.raises(BlomGpuError, match="ltedtp = bogus is unsupported"):
            g.diffus(*c.levels)
    finally:
        g.finalize()


@pytest.mark.parametrize("cfg", ["tiny0", "tiny2", "tiny3"])
@pytest.mark.parametrize("vcoord", ["cntiso_hybrid", "isopyc_bulkml"])
def test_tmsmt(cfg, vcoord):
    c, o, g = pair(cfg, opts={"vcoord": vcoord})
    try:
        m, n, mm, nn, k1m, k1n = c.levels
        o.tmsmt1(nn); g.tmsmt1(nn)
        check(g, o, ["dpold", "told", "sold", "trcold", "dpuold", "dpvold"], 0.0)
        o.tmsmt2(m, mm, nn, k1m); g.tmsmt2(m, mm, nn, k1m)
        check(g, o, ["dp", "temp", "saln", "trc", "dpu", "dpv"], 1e-14)
        check(g, o, ["p"], 1e-15, halo=2)
    finally:
        g.finalize()


@pytest.mark.parametrize("pgfmth", ["dynamic enthalpy", "geopotential"])  # phy/mod_pgforc.F90:524-534
@pytest.mark.parametrize("cfg", ["tiny0", "tiny1", "tiny2", "tiny3", "tiny4"])
@pytest.mark.parametrize("parity", [True, False])
def test_pgforc(cfg, parity, pgfmth):
    c, o, g = pair(cfg, parity=parity, opts={"pgfmth": pgfmth})
    try:
        o.pgforc(*c.levels); g.pgforc(*c.levels)
        tol = 1e-12 if parity else 1e-10
        check(g, o, ["pgfx", "pgfy", "pgfx_o", "pgfy_o", "pgfxm", "pgfym", "xixp", "xixm", "xiyp", "xiym", "sealv",
                     "phi"], tol)
        check(g, o, ["p"], 1e-15, halo=2)
        check(g, o, ["dpu", "dpv", "pu", "pv", "xixp_o", "pgfxm_o", "xiym_o"], 1e-15, halo=1)
    finally:
        g.finalize()


BT_FIELDS = ["pb", "pbu", "pbv", "ub", "vb", "ubflx", "vbflx", "ubflxs", "vbflxs", "ubflxs_p", "vbflxs_p", "pb_p",
             "pbu_p", "pbv_p", "ubcors_p", "vbcors_p", "pb_mn", "ubflx_mn", "vbflx_mn"]


@pytest.mark.parametrize("cfg,mommth", [("tiny0", "enscon"), ("tiny1", "enecon"), ("tiny2", "enscon"),
                                        ("tiny3", "enscon"), ("tiny4", "enedis"), ("fuk95", "enscon")])
def test_barotp(cfg, mommth):
    c, o, g = pair(cfg, opts={"mommth": mommth})
    try:
        o.pgforc(*c.levels); g.pgforc(*c.levels)
        o.barotp(*c.levels); g.barotp(*c.levels)
        check(g, o, BT_FIELDS, 1e-11)
        check(g, o, ["pvtrop"], 1e-14, halo=1)
        # the routine-local save arrays carry over to the next call: run a second step
        o.barotp(*c.levels); g.barotp(*c.levels)
        check(g, o, BT_FIELDS, 1e-10)
    finally:
        g.finalize()


def test_barotp_perf_build_and_bad_option():
    from blom_b200.lib import BlomGpuError
    c, o, g = pair("tiny2", parity=False)
    try:
        o.pgforc(*c.levels); g.pgforc(*c.levels)
        o.barotp(*c.levels); g.barotp(*c.levels)
        check(g, o, BT_FIELDS, 1e-9)
        g.set_option("mommth", "bogus")
        with pytest.raises(BlomGpuError, match="mommth = bogus is unsupported"):
            g.barotp(*c.levels)
    finally:
        g.finalize()


MT_FIELDS = ["u", "v", "utotn", "vtotn", "pu", "pv", "ustarb"]


@pytest.mark.parametrize("cfg,mommth,vcoord", [("tiny0", "enscon", "cntiso_hybrid"), ("tiny1", "enecon", "cntiso_hybrid"),
                                               ("tiny2", "enscon", "cntiso_hybrid"), ("tiny2", "enedis", "isopyc_bulkml"),
                                               ("tiny3", "enedis", "cntiso_hybrid"), ("tiny4", "enscon", "isopyc_bulkml"),
                                               ("fuk95", "enscon", "cntiso_hybrid")])
def test_momtum(cfg, mommth, vcoord):
    """tolerance 1e-11 of the field max-norm (parity build): one sqrt-heavy routine, ~300 flops/cell"""
    c, o, g = pair(cfg, ntr=0, opts={"mommth": mommth, "vcoord": vcoord})
    try:
        for b in (o, g):
            b.numerical_bounds()
            b.pgforc(*c.levels)
            b.momtum(*c.levels)
        check(g, o, MT_FIELDS, 1e-11)
        check(g, o, ["p"], 1e-15, halo=1)
        kk = c.dims[2]
        g.download_all()
        iq = interior(c.masks["iq"]) == 1
        for nm in ("absvor", "dpvor"):
            a, b = interior(g.arrays[nm][:kk])[:, iq], interior(o.arrays[nm][:kk])[:, iq]
            assert max_rel_err(a, b) <= 1e-11, nm
        assert np.abs(interior(g.arrays["u"]) - interior(c.state["u"])).max() > 1e-4
        # second call: module work arrays / stale halos carry over identically
        for b in (o, g):
            b.momtum(*c.levels)
        check(g, o, MT_FIELDS, 1e-10)
    finally:
        g.finalize()


@pytest.mark.parametrize("cfg,mommth,vcoord", [("tiny1", "enecon", "cntiso_hybrid"), ("tiny2", "enscon", "cntiso_hybrid"),
                                               ("tiny3", "enedis", "cntiso_hybrid"), ("tiny4", "enscon", "isopyc_bulkml"),
                                               ("fuk95", "enedis", "cntiso_hybrid")])
def test_momtum_fused_equals_staged(cfg, mommth, vcoord):
    """The shared-memory tile form (one launch, momtum_form=fused) evaluates the same
    expressions in the same order as the staged form (one launch per stage, scratch in HBM, the default): the two
    must agree BIT FOR BIT in the parity build, and the staged form stays within 1e-11 of the oracle."""
    c = Case(cfg, ntr=0)
    o = c.new_oracle()
    o.set_option("mommth", mommth); o.set_option("vcoord", vcoord)
    o.inieos(); o.numerical_bounds(); o.pgforc(*c.levels)
    res = {}
    for form in ("fused", "staged"):  # one library context at a time
        g = c.new_gpu(parity=True)
        try:
            for k, v in {"mommth": mommth, "vcoord": vcoord, "momtum_form": form}.items():
                g.set_option(k, v)
            g.inieos(); g.numerical_bounds(); g.pgforc(*c.levels)
            for rep in range(2):
                g.momtum(*c.levels)
                if form == "staged":
                    o.momtum(*c.levels)
                g.download_all()
                res[form, rep] = {nm: g.arrays[nm].copy() for nm in MT_FIELDS + ["absvor", "dpvor"]}
            res[form, "launches"] = g.launch_count()
            if form == "staged":
                check(g, o, MT_FIELDS, 1e-10)
        finally:
            g.finalize()
    for rep in range(2):
        for nm in MT_FIELDS + ["absvor", "dpvor"]:
            assert np.array_equal(res["fused", rep][nm], res["staged", rep][nm]), (nm, rep)
    assert res["fused", "launches"] < res["staged", "launches"]


def test_momtum_perf_build():
    c, o, g = pair("tiny2", ntr=0, parity=False)
    try:
        for b in (o, g):
            b.numerical_bounds(); b.pgforc(*c.levels); b.momtum(*c.levels)
        check(g, o, MT_FIELDS, 1e-9)
    finally:
        g.finalize()


ED_FIELDS = ["umfltd", "vmfltd", "umflsm", "vmflsm", "utfltd", "vtfltd", "utflsm", "vtflsm", "usfltd", "vsfltd",
             "usflsm", "vsflsm", "hbl_tf", "wpup_tf", "hml_tf1", "hml_tf", "hml_tfbnd", "util1"]


@pytest.mark.parametrize("cfg,mlrmth,slope,parity", [
    ("tiny0", "none", 1.0, True), ("tiny1", "fox08", 1.0, True), ("tiny2", "bod23", 1.0, True),
    ("tiny3", "fox08", 3.0e3, True), ("tiny4", "bod23", 3.0e3, True), ("fuk95", "bod23", 1.0e3, True),
    ("tiny2", "fox08", 3.0e3, False), ("fuk95", "fox08", 1.0, False)])
def test_eddtra(cfg, mlrmth, slope, parity):
    """eddtra_ale + heat/salt flux diagnosis.  Tolerance 1e-13 of the field max-norm for the parity
    build (only pow() of the bod23 filter input differs from host libm, <=1 ulp), 1e-10 for the FMA
    build; steep synthetic slopes (x3e3) drive the iterative limiter through several sweeps."""
    c = Case(cfg, ntr=0)
    c.state["nslpx"] *= slope; c.state["nslpy"] *= slope
    o = c.new_oracle(); g = c.new_gpu(parity=parity)
    try:
        for b in (o, g):
            b.set_option("mlrmth", mlrmth); b.inieos()
        for rep in range(2):   # second call: the running-mean filter state carries over
            o.eddtra(*c.levels); g.eddtra(*c.levels)
            g.sync()
            check(g, o, ED_FIELDS, 1e-13 if parity else 1e-10)
        kk = c.dims[2]; mm = c.levels[2]
        assert np.abs(interior(g.arrays["umfltd"][mm:mm + kk])).max() > 0.0
    finally:
        g.finalize()


@pytest.mark.parametrize("cfg,eitmth,slope,parity", [
    ("tiny0", "gm", 1.0, True), ("tiny1", "gm", 3.0e3, True), ("tiny2", "gm", 1.0, True), ("tiny3", "gm", 3.0e3, True),
    ("tiny4", "gm", 1.0, True), ("fuk95", "gm", 1.0e3, True), ("fuk95", "gm", 1.0, False),
    ("tiny1", "intdif", 1.0, True), ("tiny2", "intdif", 1.0, True), ("fuk95", "intdif", 1.0, True),
    ("fuk95", "intdif", 1.0, False)])
def test_eddtra_isopycnic(cfg, eitmth, slope, parity):
    """vcoord='isopyc_bulkml': eddtra_gm_isopyc_bulkml / eddtra_intdif_isopyc_bulkml + the heat/salt
    diagnosis (phy/mod_eddtra.F90:153-999, :1818-1857) on a state with a consistent kfpla.  Tolerance
    1e-13 of the field max-norm (parity build), 1e-10 (FMA build); steep slopes exercise the limiter."""
    c = Case(cfg, ntr=0, isopycnic=True)
    c.state["nslpx"] *= slope; c.state["nslpy"] *= slope
    o = c.new_oracle(); g = c.new_gpu(parity=parity)
    try:
        for b in (o, g):
            b.set_option("vcoord", "isopyc_bulkml"); b.set_option("eitmth", eitmth); b.inieos()
        o.eddtra(*c.levels); g.eddtra(*c.levels)
        g.sync()
        check(g, o, ["umfltd", "vmfltd", "utfltd", "vtfltd", "usfltd", "vsfltd"], 1e-13 if parity else 1e-10)
        kk = c.dims[2]; mm = c.levels[2]
        assert np.abs(interior(g.arrays["umfltd"][mm:mm + kk])).max() > 0.0
    finally:
        g.finalize()


def test_eddtra_bad_option():
    from blom_b200.lib import BlomGpuError
    c, o, g = pair("tiny0", ntr=0)
    try:
        g.set_option("mlrmth", "bogus")
        with pytest.raises(BlomGpuError, match="mlrmth = bogus is unsupported"):
            g.eddtra(*c.levels)
        g.set_option("mlrmth", "fox08"); g.set_option("eitmth", "intdif")
        with pytest.raises(BlomGpuError, match="eitmth_opt is unsupported for vcoord = 'cntiso_hybrid'"):
            g.eddtra(*c.levels)
        g.set_option("vcoord", "isopyc_bulkml"); g.set_option("eitmth", "bogus")
        with pytest.raises(BlomGpuError, match="eitmth_opt is unsupported for vcoord = 'isopyc_bulkml'"):
            g.eddtra(*c.levels)
    finally:
        g.finalize()


PBC_FIELDS = ["dp", "temp", "saln", "trc", "uflx", "vflx", "utflx", "vtflx", "usflx", "vsflx", "sigma"]


@pytest.mark.parametrize("cfg,bmcmth,parity", [("tiny0", "uc", True), ("tiny1", "dluc", True), ("tiny2", "uc", True),
                                               ("tiny2", "dluc", True), ("tiny3", "uc", True), ("tiny4", "dluc", True),
                                               ("fuk95", "uc", True), ("tiny2", "uc", False), ("fuk95", "dluc", False)])
def test_pbcor(cfg, bmcmth, parity):
    """pbcor1 then pbcor2 on the same state.  Tolerance: 1e-13 of the field max-norm (parity
    build; only + - * / and max/min) and 1e-10 for the FMA build."""
    c, o, g = pair(cfg, parity=parity, opts={"bmcmth": bmcmth})
    try:
        m, n, mm, nn, k1m, k1n = c.levels
        for b in (o, g):
            b.arrays["ubflxs"][n - 1] = b.arrays["ubflxs_p"][m - 1] * 1.01
            b.arrays["vbflxs"][n - 1] = b.arrays["vbflxs_p"][m - 1] * 1.01
        g.upload("ubflxs"); g.upload("vbflxs")
        tol = 1e-13 if parity else 1e-10
        o.pbcor1(*c.levels); g.pbcor1(*c.levels)
        check(g, o, PBC_FIELDS, tol)
        check(g, o, ["p"], tol, halo=1)
        assert np.abs(interior(g.arrays["dp"]) - interior(c.state["dp"])).max() > 0.0
        o.pbcor2(*c.levels); g.pbcor2(*c.levels)
        check(g, o, PBC_FIELDS + ["utotn", "vtotn"], tol)
        check(g, o, ["p", "dp"], tol, halo=1)
    finally:
        g.finalize()


def test_pbcor_bad_option():
    from blom_b200.lib import BlomGpuError
    c, o, g = pair("tiny0", ntr=0, opts={"bmcmth": "bogus"})
    try:
        with pytest.raises(BlomGpuError, match="bmcmth = bogus is unsupported"):
            g.pbcor1(*c.levels)
        with pytest.raises(BlomGpuError, match="bmcmth = bogus is unsupported"):
            g.pbcor2(*c.levels)
    finally:
        g.finalize()


@pytest.mark.parametrize("cfg,ntr,parity", [("tiny0", 0, True), ("tiny2", 1, True), ("fuk95", 1, True), ("tiny2", 1, False)])
def test_budget_sums(cfg, ntr, parity):
    """budget_init / budget_sums (phy/mod_budget.F90:74-196): k-ordered column sums + strip-ordered
    xcsum.  Bit-exact against the oracle in the parity build; 1e-14 relative with FMA contraction."""
    import math
    c = Case(cfg, ntr=ntr)
    o = c.new_oracle(); g = c.new_gpu(parity=parity)
    try:
        m, n, mm, nn, k1m, k1n = c.levels
        a, b = g.budget_init(), o.budget_init()
        assert a == b
        for ncall in (1, 4):
            ga, oa = g.budget_sums(ncall, n, nn), o.budget_sums(ncall, n, nn)
            for x, y in zip(ga, oa):
                if math.isnan(y):
                    assert math.isnan(x)
                elif parity:
                    assert x == y, (ga, oa)
                else:
                    assert abs(x - y) <= 1e-14 * abs(y), (ga, oa)
        g.download_all()
        ip = interior(c.masks["ip"]) == 1
        if parity:
            assert np.array_equal(interior(g.arrays["util1"])[0][ip], interior(o.arrays["util1"])[0][ip])
    finally:
        g.finalize()


@pytest.mark.parametrize("cfg", ["tiny0", "tiny1", "tiny3"])
def test_budget_conserved_through_transport_and_diffusion(cfg):
    """the global inventories sum(S dp scp2), sum(T dp scp2), sum(trc dp scp2) as the reference's own
    diagnostic (budget_sums) measures them are conserved to round-off (1e-13 relative) across
    advect -> diffus on closed / periodic domains (no fold), all on device."""
    c = Case(cfg, ntr=1)
    g = c.new_gpu(parity=True)
    try:
        g.inieos(); g.numerical_bounds(); g.init_cppm()
        m, n, mm, nn, k1m, k1n = c.levels
        before = g.budget_sums(1, n, nn)
        g.advect(*c.levels)
        mid = g.budget_sums(2, n, nn)
        g.diffus(*c.levels)
        after = g.budget_sums(3, n, nn)
        for i in range(3):
            assert abs(mid[i] - before[i]) <= 1e-13 * abs(before[i]), (i, before, mid)
            assert abs(after[i] - before[i]) <= 1e-13 * abs(before[i]), (i, before, after)
        assert mid != before   # the fields did move
    finally:
        g.finalize()


def test_init_again_without_finalize_starts_clean():
    """A host that aborts half way and calls blomgpu_init again (no finalize) must not inherit device arrays sized
    for the previous tile: the second, LARGER tile runs pgforc + momtum and matches the oracle (a stale, too small
    allocation would be a device heap overflow)."""
    small = Case("tiny0", ntr=1)
    g0 = small.new_gpu(parity=True)
    g0.inieos(); g0.pgforc(*small.levels)
    # no g0.finalize(): same library context, new geometry
    c, o, g = pair("mid2", ntr=1, opts={"vcoord": "cntiso_hybrid"})
    try:
        for b in (o, g):
            b.numerical_bounds()
            b.pgforc(*c.levels)
            b.momtum(*c.levels)
        check(g, o, ["u", "v", "pgfx", "pgfy", "dpu", "dpv"], 1e-11)
    finally:
        g.finalize()


def test_comm_errors_reach_last_error():
    """blomgpu_comm_* report through blomgpu_last_error like every other entry (a 1-rank communicator from a
    garbage id must fail, not hang: NCCL rejects the id during bootstrap or the rank count)."""
    from blom_b200.lib import load_library
    lib = load_library(True)
    rc = lib.blomgpu_comm_init(b"\0" * 128, 3, 2)      # rank 3 of 2: invalid argument, returns at once
    assert rc != 0
    assert b"NCCL" in lib.blomgpu_last_error() or b"nccl" in lib.blomgpu_last_error()
