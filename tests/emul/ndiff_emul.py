"""TEST INFRASTRUCTURE ONLY: builds tests/emul/ndiff_emul.cpp (the product's neutral-diffusion kernels compiled
for the host, see cuda_host_shim.hpp) and runs it on numpy arrays."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SRC = [HERE / "ndiff_emul.cpp", HERE / "cuda_host_shim.hpp", HERE.parents[1] / "blom_b200/csrc/ndiff.cu",
       HERE.parents[1] / "blom_b200/csrc/common.cuh", HERE.parents[1] / "blom_b200/csrc/eos.cuh"]
LIB = HERE / "_build" / "libndiff_emul.so"

_PD, _PI = C.POINTER(C.c_double), C.POINTER(C.c_int)
_INTS = ["ii", "jj", "kdm", "nb", "ldi", "ldj", "ntr", "mm", "nn", "surface_align", "ix64"]
_IN_I = ["ip", "iu", "iv", "ksmx"]
_IN_D = ["p_src", "tsd", "tpc", "p_dst", "dpml", "difiso", "temp", "saln", "trc",
         "scuy", "scuxi", "scvx", "scvyi", "scp2", "pu", "pv"]
OUT = ["utflld", "usflld", "vtflld", "vsflld", "utflx", "usflx", "vtflx", "vsflx", "nslpx", "nslpy", "trc_rm"]


class EmuNdiff(C.Structure):
    _fields_ = ([(n, C.c_int) for n in _INTS] + [("delt1", C.c_double)] + [(n, _PI) for n in _IN_I] +
                [(n, _PD) for n in _IN_D] + [(n, _PD) for n in OUT])


def build(force=False):
    if not force and LIB.exists() and all(LIB.stat().st_mtime >= s.stat().st_mtime for s in SRC):
        return LIB
    LIB.parent.mkdir(exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-w",
           "-I/usr/local/cuda/include", "-o", str(LIB), str(SRC[0])]
    subprocess.run(cmd, check=True)
    return LIB


def run(dims, levels, delt1, arrays, surface_align=True, variant=3):
    """dims = (ii, jj, kdm, nb, ldi, ldj, ntr); `arrays` maps the field names of EmuNdiff to C-contiguous numpy
    arrays in the common (level, j, i) layout; the OUT arrays are updated in place.  variant: 0..3 = option
    ndiff_stage of the product (32-bit index instantiation), 9 = the 64-bit index instantiation."""
    lib = C.CDLL(str(build()))
    lib.emu_ndiff.argtypes = [C.POINTER(EmuNdiff)]
    lib.emu_ndiff.restype = C.c_int
    e = EmuNdiff()
    ii, jj, kdm, nb, ldi, ldj, ntr = dims
    m, n, mm, nn, k1m, k1n = levels
    for k, v in dict(ii=ii, jj=jj, kdm=kdm, nb=nb, ldi=ldi, ldj=ldj, ntr=ntr, mm=mm, nn=nn,
                     surface_align=int(surface_align), ix64=int(variant)).items():
        setattr(e, k, v)
    e.delt1 = delt1
    keep = []
    for nm in _IN_I + _IN_D + OUT:
        a = arrays.get(nm)
        if a is None:
            assert nm == "trc" and ntr == 0, nm
            continue
        assert a.flags["C_CONTIGUOUS"], nm
        assert a.dtype == (np.int32 if nm in _IN_I else np.float64), (nm, a.dtype)
        keep.append(a)
        setattr(e, nm, a.ctypes.data_as(_PI if nm in _IN_I else _PD))
    rc = lib.emu_ndiff(C.byref(e))
    assert rc == 0, rc
