#!/usr/bin/env python
"""csdiag_log.py - the reference's checksum diagnostics for the hot path, through the C ABI.

With CSDIAG = .true. (namelist LIMITS, tests/fuk95/limits:197) every BLOM routine ends with a block of
    ' chksum: <name>: 0x%08X'
lines (phy/mod_checksum.F90:41-74): the CRC-32 of the masked field, bit for bit.  This tool runs the analytic
fuk95 case (blom_b200/fuk95.py: the geometry, stratification and namelist scalars of the reference's own
stand-alone test, fuk95/mod_fuk95.F90:117-445) through the hot-path routines on cuda:0 (libblomgpu_parity.so)
and prints exactly the blocks the reference prints for them, in the reference's order and format:

    python tools/csdiag_log.py --steps 3

What it is for.  The Fortran reference cannot be built in this image, so nothing here pins the results to the
reference itself.  A maintainer with gfortran + meson can: the README next to the CPU
checker gives the meson commands and the small patch that makes `fuk95_blom` skip the routines that are out of
scope here (ALE regrid, column physics, forcing); run it with CSDIAG = .true. and `diff` its log against the
output of this tool - every line that matches is a bit-exact statement about a whole 3-D field.  The same log
produced by the CPU checker is committed as tests/golden/fuk95_csdiag.txt (tests/dev/csdiag_cpu.py), and
tests/test_gpu_csdiag.py requires this tool to reproduce it to the bit.
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

PS, US, VS, UV, VV = 1, 3, 4, 13, 14

# the chksum calls at the end of each routine: (array, first level or None, levels, itype, text)
#   'kk2' = 2*kk, 'kk1' = kk+1, k1m = first level of the mid time level
BLOCKS = {
    "tmsmt1": [("dpold", None, "kk2", PS, "dpold"), ("told", None, "kk", PS, "told"),          # phy/mod_tmsmt.F90:260-275
               ("sold", None, "kk", PS, "sold"), ("TRC", None, "kk", PS, "trcold")],
    "eddtra": [("hbl_tf", None, 1, PS, "hbl_tf"), ("wpup_tf", None, 1, PS, "wpup_tf"),         # phy/mod_eddtra.F90:1906-1925
               ("hml_tf1", None, 1, PS, "hml_tf1"), ("hml_tf", None, 1, PS, "hml_tf"),
               ("umfltd", "k1m", "kk", UV, "umfltd"), ("vmfltd", "k1m", "kk", VV, "vmfltd"),
               ("umflsm", "k1m", "kk", UV, "umflsm"), ("vmflsm", "k1m", "kk", VV, "vmflsm"),
               ("utfltd", "k1m", "kk", UV, "utfltd"), ("vtfltd", "k1m", "kk", VV, "vtfltd"),
               ("utflsm", "k1m", "kk", UV, "utflsm"), ("vtflsm", "k1m", "kk", VV, "vtflsm"),
               ("usfltd", "k1m", "kk", UV, "usfltd"), ("vsfltd", "k1m", "kk", VV, "vsfltd"),
               ("usflsm", "k1m", "kk", UV, "usflsm"), ("vsflsm", "k1m", "kk", VV, "vsflsm")],
    "advect": [("dp", None, "kk2", PS, "dp"), ("temp", None, "kk2", PS, "temp"),               # phy/mod_advect.F90:174-187
               ("saln", None, "kk2", PS, "saln"), ("uflx", None, "kk2", UV, "uflx"),
               ("vflx", None, "kk2", VV, "vflx"), ("TRC", None, "kk2", PS, "trc")],
    "pbcor1": [("dp", None, "kk2", PS, "dp"), ("temp", None, "kk2", PS, "temp"),               # phy/mod_pbcor.F90:395-410
               ("saln", None, "kk2", PS, "saln"), ("uflx", None, "kk2", UV, "uflx"),
               ("vflx", None, "kk2", VV, "vflx"), ("TRC", None, "kk2", PS, "trc")],
    "diffus": [("temp", None, "kk2", PS, "temp"), ("saln", None, "kk2", PS, "saln"),           # phy/mod_diffus.F90:165-183
               ("TRC", None, "kk2", PS, "trc"),
               ("utflld", "k1m", "kk", UV, "utflld"), ("vtflld", "k1m", "kk", VV, "vtflld"),
               ("usflld", "k1m", "kk", UV, "usflld"), ("vsflld", "k1m", "kk", VV, "vsflld"),
               ("utflx", "k1m", "kk", UV, "utflx"), ("vtflx", "k1m", "kk", VV, "vtflx"),
               ("usflx", "k1m", "kk", UV, "usflx"), ("vsflx", "k1m", "kk", VV, "vsflx")],
    "pgforc": [("phi", None, "kk1", PS, "phi"), ("pgfx", None, "kk2", UV, "pgfx"),             # phy/mod_pgforc.F90:600-613
               ("pgfy", None, "kk2", VV, "pgfy"), ("pgfxm", None, 2, UV, "pgfxm"),
               ("pgfym", None, 2, VV, "pgfym"), ("xixp", None, 2, US, "xixp"), ("xixm", None, 2, US, "xixm"),
               ("xiyp", None, 2, VS, "xiyp"), ("xiym", None, 2, VS, "xiym")],
    "momtum": [("dpu", None, "kk2", US, "dpu"), ("dpv", None, "kk2", VS, "dpv"),               # phy/mod_momtum.F90:1270-1280
               ("u", None, "kk2", UV, "u"), ("v", None, "kk2", VV, "v"),
               ("utotn", None, 1, UV, "utotn"), ("vtotn", None, 1, VV, "vtotn")],
    "barotp": [("pb", None, 2, PS, "pb"), ("pbu", None, 2, US, "pbu"), ("ubflx", None, 2, UV, "ubflx"),   # phy/mod_barotp.F90:981-1001
               ("ub", None, 2, UV, "ub"), ("ubflxs", None, 3, UV, "ubflxs"), ("pbv", None, 2, VS, "pbv"),
               ("vbflx", None, 2, VV, "vbflx"), ("vb", None, 2, VV, "vb"), ("vbflxs", None, 3, VV, "vbflxs"),
               ("pb_p", None, 1, PS, "pb_p"), ("pbu_p", None, 1, US, "pbu_p"),
               ("ubflxs_p", None, 2, UV, "ubflxs_p"), ("ubcors_p", None, 1, UV, "ubcors_p"),
               ("pbv_p", None, 1, VS, "pbv_p"), ("vbflxs_p", None, 2, VV, "vbflxs_p"),
               ("vbcors_p", None, 1, VV, "vbcors_p")],
    "pbcor2": [("dp", None, "kk2", PS, "dp"), ("temp", None, "kk2", PS, "temp"),               # phy/mod_pbcor.F90:726-741
               ("saln", None, "kk2", PS, "saln"), ("p", None, "kk1", PS, "p"), ("sigma", None, "kk2", PS, "sigma"),
               ("uflx", None, "kk2", UV, "uflx"), ("vflx", None, "kk2", VV, "vflx"), ("TRC", None, "kk2", PS, "trc")],
    "tmsmt2": [("dp", None, "kk2", PS, "dp"), ("temp", None, "kk2", PS, "temp"),               # phy/mod_tmsmt.F90:395-408
               ("saln", None, "kk2", PS, "saln"), ("dpu", None, "kk2", US, "dpu"), ("dpv", None, "kk2", VS, "dpv"),
               ("TRC", None, "kk2", PS, "trc")],
}


def csdiag_lines(b, routine, levels, kk, ntr):
    """The lines the reference writes at the end of `routine`; b: any backend with chksum_at (BlomGpu)."""
    m, n, mm, nn, k1m, k1n = levels
    out = [f" {routine}:"]
    for arr, first, nlev, itype, text in BLOCKS.get(routine, []):
        nl = {"kk": kk, "kk2": 2 * kk, "kk1": kk + 1}.get(nlev, nlev)
        if arr == "TRC":     # do nt = 1,ntr: chksum(trc(1-nbdy,1-nbdy,1,nt), ...) with text 'trc'//'01'
            name = "trcold" if text == "trcold" else "trc"
            for nt in range(ntr):
                out.append(" chksum: %s%02d: 0x%08X" % (text, nt + 1, b.chksum_at(name, nt * nl + 1, nl, itype)))
            continue
        koff = k1m if first == "k1m" else 1
        out.append(" chksum: %s: 0x%08X" % (text, b.chksum_at(arr, koff, nl, itype)))
    return out


def run(steps=3, config="fuk95_analytic", ntr=1, options=None, out=sys.stdout):
    """The log of `steps` baroclinic steps of the CUDA path (device-resident state, C ABI)."""
    from blom_b200.driver import HotPath, run_step
    # layer diffusion inside diffus (the fuk95 namelist has no neutral diffusion; its inputs are ALE products,
    # which a hot-path-only run does not have)
    hp = HotPath(config, ntr=ntr, nstep=1, parity=True, options={"ltedtp": "layer", **(options or {})})
    try:
        lines = []
        for nstep in range(1, steps + 1):
            hp.set_step(nstep)
            lines.append(f" step {nstep:6d}")
            for r in hp.routines:
                run_step(hp.gpu, [r], hp.levels)
                if r in BLOCKS:
                    lines += csdiag_lines(hp.gpu, r, hp.levels, hp.kdm, ntr)
        for ln in lines:
            print(ln, file=out)
        return lines
    finally:
        hp.finalize()


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--config", default="fuk95_analytic")
    ap.add_argument("--ntr", type=int, default=1)
    a = ap.parse_args()
    run(a.steps, a.config, a.ntr)
