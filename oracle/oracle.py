"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes driver of the CPU restatement (oracle/*.cpp).  Imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
Same method names as blom_b200.lib.BlomGpu so parity tests read symmetrically;
it works in place on the numpy arrays it is given (the reference's routines
work in place on module arrays).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "liboracle.so"


def build(force: bool = False) -> Path:
    srcs = list(HERE.glob("*.cpp")) + list(HERE.glob("*.hpp")) + [HERE / "Makefile"]
    if force or not LIB.exists() or any(s.stat().st_mtime > LIB.stat().st_mtime for s in srcs):
        r = subprocess.run(["make", "-C", str(HERE), "-j", str(min(8, os.cpu_count() or 1))],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return LIB


class OracleError(RuntimeError):
    pass


class Oracle:
    def __init__(self, itdm, jtdm, kdm, nreg, ntr=0, nbdy=4, build_if_missing=True):
        if build_if_missing and not LIB.exists():
            build()
        self.lib = C.CDLL(str(LIB))
        self.lib.oracle_last_error.restype = C.c_char_p
        self.lib.oracle_get_int.restype = C.c_long
        self.lib.oracle_get_owned.restype = C.c_long
        self.lib.oracle_cppm_table.restype = C.c_long
        self.lib.oracle_cppm_stencil.restype = C.c_long
        self.lib.oracle_crc32.restype = C.c_uint32
        self.itdm, self.jtdm, self.kdm, self.ntr, self.nbdy = itdm, jtdm, kdm, ntr, nbdy
        self.idm, self.jdm = itdm, jtdm
        dims = (C.c_int * 8)(itdm, jtdm, kdm, itdm, jtdm, nbdy, ntr, nreg)
        self._ck(self.lib.oracle_init(dims))
        self.arrays = {}

    def _ck(self, rc):
        if rc != 0:
            raise OracleError(self.lib.oracle_last_error().decode())

    @property
    def shape2d(self):
        return (self.jdm + 2 * self.nbdy, self.idm + 2 * self.nbdy)

    @property
    def nreg(self):
        return self.lib.oracle_nreg()

    def register(self, name, a, upload=True):
        a = np.asarray(a)
        assert a.flags.c_contiguous
        lev = self.shape2d[0] * self.shape2d[1]
        assert a.size % lev == 0, name
        nlev = a.size // lev
        if a.dtype == np.float64:
            self._ck(self.lib.oracle_register(name.encode(), a.ctypes.data_as(C.c_void_p), nlev))
        elif a.dtype == np.int32:
            self._ck(self.lib.oracle_register_int(name.encode(), a.ctypes.data_as(C.c_void_p), nlev))
        else:
            raise OracleError(f"{name}: dtype {a.dtype}")
        self.arrays[name] = a

    def register_all(self, state, upload=True):
        for k, v in state.items():
            self.register(k, v)

    def set_option(self, key, value):
        self._ck(self.lib.oracle_set_option(key.encode(), str(value).encode()))

    def set_scalar(self, key, value):
        self._ck(self.lib.oracle_set_scalar(key.encode(), C.c_double(float(value))))

    def get_scalar(self, key):
        self.lib.oracle_get_scalar.restype = C.c_double
        return self.lib.oracle_get_scalar(key.encode())

    def set_scalars(self, **kw):
        for k, v in kw.items():
            self.set_scalar(k, v)

    # no-ops so the same driver code works for both
    def upload(self, name): pass
    def upload_all(self): pass
    def download(self, name): return self.arrays[name]
    def download_all(self): pass
    def sync(self): pass

    def get_int(self, name):
        n = self.lib.oracle_get_int(name.encode(), None, 0)
        if n < 0:
            raise OracleError(f"no owned int array {name}")
        out = np.zeros(n, dtype=np.int32)
        self.lib.oracle_get_int(name.encode(), out.ctypes.data_as(C.c_void_p), C.c_long(n))
        return out

    def fetch(self, name, nlev=1, dtype=np.float64):
        if dtype == np.int32:
            return self.get_int(name).reshape((-1,) + self.shape2d)
        n = self.lib.oracle_get_owned(name.encode(), None, 0)
        if n < 0:
            raise OracleError(f"no owned array {name}")
        out = np.zeros(n)
        self.lib.oracle_get_owned(name.encode(), out.ctypes.data_as(C.c_void_p), C.c_long(n))
        return out.reshape((-1,) + self.shape2d)

    def cppm_table(self, name):
        n = self.lib.oracle_cppm_table(name.encode(), None, 0)
        out = np.zeros(n)
        self.lib.oracle_cppm_table(name.encode(), out.ctypes.data_as(C.c_void_p), C.c_long(n))
        return out

    def cppm_stencil(self, name):
        n = self.lib.oracle_cppm_stencil(name.encode(), None, 0)
        out = np.zeros(n, dtype=np.int32)
        self.lib.oracle_cppm_stencil(name.encode(), out.ctypes.data_as(C.c_void_p), C.c_long(n))
        return out

    # mod_xc
    def xctilr(self, name, l1, ld, mh, nh, itype, koff=1):
        self._ck(self.lib.oracle_xctilr_at(name.encode(), koff, l1, ld, mh, nh, itype))

    def xcsum(self, name, mask="ip", lev=1):
        out = C.c_double()
        self._ck(self.lib.oracle_xcsum(name.encode(), lev, mask.encode(), C.byref(out)))
        return out.value

    def chksum(self, name, kcsd, itype):
        mask = {1: "ip", 11: "ip", 2: "iq", 12: "iq", 3: "iu", 13: "iu", 4: "iv", 14: "iv"}[itype]
        out = C.c_uint32()
        self._ck(self.lib.oracle_xccrc(name.encode(), kcsd, mask.encode(), C.byref(out)))
        return out.value

    def chksum_at(self, name, koff, kcsd, itype):
        """chksum(a(1-nbdy,1-nbdy,koff), kcsd, itype, text), phy/mod_checksum.F90:41-74"""
        mask = {1: "ip", 11: "ip", 2: "iq", 12: "iq", 3: "iu", 13: "iu", 4: "iv", 14: "iv"}[itype]
        out = C.c_uint32()
        self._ck(self.lib.oracle_xccrc_at(name.encode(), koff, kcsd, mask.encode(), C.byref(out)))
        return out.value

    def crc32(self, data: bytes, init=0):
        return self.lib.oracle_crc32(data, C.c_long(len(data)), C.c_uint32(init))

    def eos(self, fn, *args, nout=1):
        a = (C.c_double * len(args))(*args)
        out = (C.c_double * 3)()
        if self.lib.oracle_eos(fn.encode(), a, out) != 0:
            raise OracleError(f"unknown eos function {fn}")
        return out[0] if nout == 1 else tuple(out[:nout])

    def bigrid(self, depth="depths"):
        self._ck(self.lib.oracle_bigrid(depth.encode()))

    def init_cppm(self):
        self._ck(self.lib.oracle_init_cppm())

    def _call(self, name, *a):
        fn = getattr(self.lib, "oracle_" + name, None)
        if fn is None:
            raise OracleError(f"oracle_{name} not built")
        self._ck(fn(*a))

    def inieos(self): self._call("inieos")
    def numerical_bounds(self): self._call("numerical_bounds")
    def init_fluxes(self, *a): self._call("init_fluxes", *a)
    def tmsmt1(self, nn): self._call("tmsmt1", nn)
    def tmsmt2(self, m, mm, nn, k1m): self._call("tmsmt2", m, mm, nn, k1m)
    def eddtra(self, *a): self._call("eddtra", *a)
    def advect(self, *a): self._call("advect", *a)
    def pbcor1(self, *a): self._call("pbcor1", *a)
    def diffus(self, *a): self._call("diffus", *a)
    def pgforc(self, *a): self._call("pgforc", *a)
    def momtum(self, *a): self._call("momtum", *a)
    def barotp(self, *a): self._call("barotp", *a)
    def pbcor2(self, *a): self._call("pbcor2", *a)
    def ndiff(self, *a): self._call("ndiff", *a)

    def difest_halos(self, m, n, mm, nn, k1m, k1n):
        """xctilr calls of phy/mod_difest.F90:826-831 and phy/mod_cmnfld_routines.F90:1171-1172."""
        kk = self.kdm
        self.xctilr("u", 1, 2 * kk, 2, 2, 13); self.xctilr("v", 1, 2 * kk, 2, 2, 14)
        for nm, it in (("ubflxs_p", 13), ("vbflxs_p", 14), ("pbu", 3), ("pbv", 4)):
            self.xctilr(nm, 1, 2, 2, 2, it)
        self.xctilr("temp", 1, 2 * kk, 3, 3, 1); self.xctilr("saln", 1, 2 * kk, 3, 3, 1)
    def cmnfld2(self, *a): self._call("cmnfld2", *a)
    def cmnfld_bfsqf_ale(self, *a): self._call("cmnfld_bfsqf_ale", *a)
    def cmnfld_nslope_ale(self, *a): self._call("cmnfld_nslope_ale", *a)
    def cmnfld_nnslope_ale(self, *a): self._call("cmnfld_nnslope_ale", *a)

    def budget_init(self):
        """mass0 of budget_init (phy/mod_budget.F90:74-93)"""
        out = C.c_double()
        self._call("budget_init", C.byref(out))
        return out.value

    def budget_sums(self, ncall, n, nn):
        """(sdp, tdp, trdp, sc) of budget_sums (phy/mod_budget.F90:95-196); nan = not evaluated"""
        out = (C.c_double * 4)(*([float("nan")] * 4))
        self._call("budget_sums", ncall, n, nn, out)
        return tuple(out)
