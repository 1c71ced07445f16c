"""CPU parity of the PRODUCT's neutral-diffusion kernel source (blom_b200/csrc/ndiff.cu) against the oracle.

tests/emul compiles ndiff.cu itself for the host (g++, -ffp-contract=off; every kernel of that file is a
one-thread-per-column program without cross-thread communication) and runs one emulated thread after the
other.  The face fields follow the reference's operation order exactly and must agree with the oracle bit for
bit; the updated tracers carry the one documented reassociation (ndiff.cu header), <= 1e-13 of the max-norm.
This is the same comparison tests/test_gpu_ndiff.py makes on the B200, so a change of the kernel source can be
checked without a GPU.  (The emulation is test infrastructure; the product never runs on the CPU.)"""
import numpy as np
import pytest

from util import Case, interior, max_rel_err
from blom_b200 import synth
from emul import ndiff_emul

FACE = ["utflld", "usflld", "vtflld", "vsflld", "utflx", "usflx", "vtflx", "vsflx", "nslpx", "nslpy"]
ND = {"ksmx": "nd_ksmx", "p_src": "nd_p_src", "tsd": "nd_t_srcdi", "tpc": "nd_tpc_src", "p_dst": "nd_p_dst",
      "trc_rm": "nd_trc_rm"}
GRID = ["dpml", "difiso", "temp", "saln", "trc", "scuy", "scuxi", "scvx", "scvyi", "scp2", "pu", "pv"]


def run_pair(cfg, ntr, nstep, align, variant=3):
    c = Case(cfg, ntr=ntr, nstep=nstep)
    nd = {k: v.copy() for k, v in synth.ndiff_inputs(c.syn, c.state, c.levels, ntr=ntr).items()}
    o = c.new_oracle()
    o.register_all(nd)
    o.inieos()
    o.set_option("ndiff_surface_align", align)
    o.pgforc(*c.levels)
    if align == "1":
        o.xctilr("dpml", 1, 1, 1, 1, 1)     # the halo update ndiff_dev issues (idempotent for the oracle's own)
    emu = {k: o.arrays[v].copy() for k, v in ND.items()}
    emu.update({k: o.arrays[k].copy() for k in GRID + FACE if k in o.arrays})
    emu.update({k: np.ascontiguousarray(c.masks[k]) for k in ("ip", "iu", "iv")})
    o.ndiff(*c.levels)
    itdm, jtdm, kdm, _ = c.dims
    ndiff_emul.run((itdm, jtdm, kdm, 4, c.syn.ldi, c.syn.ldj, ntr), c.levels, c.scalars["delt1"], emu,
                   surface_align=align == "1", variant=variant)
    return c, o, emu


@pytest.mark.parametrize("cfg", ["tiny0", "tiny1", "tiny2", "tiny3", "tiny4", "fuk95"])
@pytest.mark.parametrize("align", ["1", "0"])
@pytest.mark.parametrize("ntr", [0, 1, 2])
def test_kernel_source_against_oracle(cfg, align, ntr):
    c, o, emu = run_pair(cfg, ntr, 1, align)
    for nm in FACE:
        a, b = interior(emu[nm]), interior(o.arrays[nm])
        assert np.abs(b).max() > 0.0 or nm.endswith("flld") or cfg == "tiny0", nm
        assert np.array_equal(a, b), (nm, max_rel_err(a, b))
    assert max_rel_err(interior(emu["trc_rm"]), interior(o.arrays["nd_trc_rm"])) <= 1e-13


@pytest.mark.parametrize("cfg,ntr", [("tiny2", 0), ("tiny4", 1), ("fuk95", 2)])
@pytest.mark.parametrize("variant", [0, 1, 2, 4, 9])
def test_kernel_source_other_instantiations(cfg, ntr, variant):
    """ndiff_stage = 0 / 1 / 2 (records read in place instead of staged, other prefetches) and the 64-bit index instantiation that
    ndiff_dev falls back to on very large tiles are the same code paths with other accessors"""
    c, o, emu = run_pair(cfg, ntr, 1, "1", variant=variant)
    for nm in FACE:
        assert np.array_equal(interior(emu[nm]), interior(o.arrays[nm])), nm
    assert max_rel_err(interior(emu["trc_rm"]), interior(o.arrays["nd_trc_rm"])) <= 1e-13


@pytest.mark.parametrize("align", ["1", "0"])
@pytest.mark.parametrize("ntr", [0, 1])
def test_kernel_source_53_layers(align, ntr):
    """kdm = 53 as in the production grids: the partner-table masks use their second word, columns are deep"""
    c, o, emu = run_pair("tiny2k53", ntr, 1, align)
    for nm in FACE:
        assert np.array_equal(interior(emu[nm]), interior(o.arrays[nm])), nm
    assert max_rel_err(interior(emu["trc_rm"]), interior(o.arrays["nd_trc_rm"])) <= 1e-13


@pytest.mark.parametrize("cfg,ntr", [("tiny2", 5), ("tiny2k53", 6)])
def test_kernel_source_many_passive_tracers(cfg, ntr):
    """T = 7 and the compiled maximum T = 8: the generic instantiation, tracer parts of the records read in place"""
    c, o, emu = run_pair(cfg, ntr, 1, "1")
    for nm in FACE:
        assert np.array_equal(interior(emu[nm]), interior(o.arrays[nm])), nm
    assert max_rel_err(interior(emu["trc_rm"]), interior(o.arrays["nd_trc_rm"])) <= 1e-13


def test_kernel_source_full_size_tnx1v4():
    """the whole 360 x 385 x 53 grid (tripolar fold, 53 layers, ~116 000 wet faces per direction): about a minute"""
    c, o, emu = run_pair("tnx1v4", 0, 1, "1")
    for nm in FACE:
        assert np.array_equal(interior(emu[nm]), interior(o.arrays[nm])), nm
    assert max_rel_err(interior(emu["trc_rm"]), interior(o.arrays["nd_trc_rm"])) <= 1e-13


def test_kernel_source_second_time_level():
    c, o, emu = run_pair("tiny3", 1, 2, "1")
    for nm in FACE:
        assert np.array_equal(interior(emu[nm]), interior(o.arrays[nm])), nm
    assert max_rel_err(interior(emu["trc_rm"]), interior(o.arrays["nd_trc_rm"])) <= 1e-13
