"""Host-side mirror of the reference's module entry points over the C ABI.

The reference has no plugin API: `blom_step` calls module procedures
`X(m,n,mm,nn,k1m,k1n)` on module-global arrays (phy/mod_blom_step.F90:146-227).
`BlomGpu` exposes the same names / argument meaning on top of
`include/blomgpu.h`; arrays are numpy views of the caller's memory in the
reference layout a(1-nbdy:idm+nbdy,1-nbdy:jdm+nbdy,nlev) (numpy shape
(nlev, jdm+2nbdy, idm+2nbdy), C order == Fortran column-major).

There is no CPU fallback: constructing BlomGpu without the built CUDA library
or without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent

HALO_PS, HALO_QS, HALO_US, HALO_VS = 1, 2, 3, 4
HALO_PV, HALO_QV, HALO_UV, HALO_VV = 11, 12, 13, 14

# every symbol include/blomgpu.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "blomgpu_init", "blomgpu_finalize", "blomgpu_last_error", "blomgpu_parity_build",
    "blomgpu_comm_unique_id", "blomgpu_comm_init",
    "blomgpu_register", "blomgpu_register_int", "blomgpu_upload", "blomgpu_download",
    "blomgpu_upload_all", "blomgpu_download_all", "blomgpu_download_async", "blomgpu_download_levels_async",
    "blomgpu_upload_async", "blomgpu_wait_upload", "blomgpu_sync", "blomgpu_device_ptr",
    "blomgpu_set_option", "blomgpu_set_scalar", "blomgpu_get_scalar",
    "blomgpu_xctilr", "blomgpu_xcsum", "blomgpu_xcmax", "blomgpu_xcmin", "blomgpu_chksum", "blomgpu_chksum_at",
    "blomgpu_bigrid", "blomgpu_nreg", "blomgpu_init_cppm", "blomgpu_inieos",
    "blomgpu_numerical_bounds", "blomgpu_init_fluxes",
    "blomgpu_tmsmt1", "blomgpu_difest_halos", "blomgpu_eddtra", "blomgpu_advect", "blomgpu_pbcor1", "blomgpu_diffus",
    "blomgpu_pgforc", "blomgpu_momtum", "blomgpu_barotp", "blomgpu_pbcor2", "blomgpu_tmsmt2",
    "blomgpu_ndiff", "blomgpu_cmnfld2", "blomgpu_cmnfld_bfsqf_ale", "blomgpu_cmnfld_nslope_ale",
    "blomgpu_cmnfld_nnslope_ale", "blomgpu_budget_init", "blomgpu_budget_sums",
    "blomgpu_launch_count", "blomgpu_launch_count_reset", "blomgpu_timers_enable",
    "blomgpu_timers_get", "blomgpu_timers_reset", "blomgpu_stream",
    "blomgpu_ktimers_enable", "blomgpu_ktimers_get",
]


class BlomGpuError(RuntimeError):
    pass


def library_path(parity: bool = False) -> Path:
    return HERE / ("libblomgpu_parity.so" if parity else "libblomgpu.so")


def load_library(parity: bool = False) -> C.CDLL:
    path = library_path(parity)
    if not path.exists():
        raise BlomGpuError(
            f"{path} is missing: build it with `python -m blom_b200.build` "
            "(the hot path has no CPU fallback)")
    # RTLD_LOCAL, and the libraries are linked with -Bsymbolic: the two flavours export the same C++
    # symbols, and with a global scope the flavour loaded first would serve the internal calls of
    # the other (its blomgpu_* entry points would launch the other flavour's kernels)
    lib = C.CDLL(str(path), mode=C.RTLD_LOCAL)
    lib.blomgpu_last_error.restype = C.c_char_p
    lib.blomgpu_launch_count.restype = C.c_long
    lib.blomgpu_stream.restype = C.c_void_p
    return lib


class BlomGpu:
    """One tile (one GPU) of the horizontal stencil step."""

    def __init__(self, itdm, jtdm, kdm, nreg, ntr=0, nbdy=4, *, j0=0, jj=None, rank=0, nranks=1,
                 device=0, parity=False):
        self.lib = load_library(parity)
        self.itdm, self.jtdm, self.kdm, self.ntr, self.nbdy = itdm, jtdm, kdm, ntr, nbdy
        self.idm = itdm
        self.jdm = jtdm if jj is None else jj
        self.j0 = j0
        self.rank, self.nranks = rank, nranks
        dims = (C.c_int * 8)(itdm, jtdm, kdm, self.idm, self.jdm, nbdy, ntr, nreg)
        tile = (C.c_int * 6)(0, j0, self.idm, self.jdm, rank, nranks)
        self._ck(self.lib.blomgpu_init(dims, tile, device))
        self.arrays: dict[str, np.ndarray] = {}

    # -- plumbing ---------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise BlomGpuError(self.lib.blomgpu_last_error().decode())

    @property
    def shape2d(self):
        return (self.jdm + 2 * self.nbdy, self.idm + 2 * self.nbdy)

    @property
    def nreg(self):
        return self.lib.blomgpu_nreg()

    def register(self, name, a, upload=True):
        a = np.asarray(a)
        if not a.flags.c_contiguous:
            raise BlomGpuError(f"{name}: array must be contiguous")
        lev = self.shape2d[0] * self.shape2d[1]
        if a.size % lev:
            raise BlomGpuError(f"{name}: size {a.size} is not a multiple of a level ({lev})")
        nlev = a.size // lev
        if a.dtype == np.float64:
            self._ck(self.lib.blomgpu_register(name.encode(), a.ctypes.data_as(C.c_void_p), nlev))
        elif a.dtype == np.int32:
            self._ck(self.lib.blomgpu_register_int(name.encode(), a.ctypes.data_as(C.c_void_p), nlev))
        else:
            raise BlomGpuError(f"{name}: dtype {a.dtype} unsupported (float64/int32)")
        self.arrays[name] = a
        if upload:
            self.upload(name)

    def register_all(self, state: dict, upload=True):
        for k, v in state.items():
            self.register(k, v, upload=False)
        if upload:
            self.upload_all()

    def upload(self, name):
        self._ck(self.lib.blomgpu_upload(name.encode()))

    def download(self, name):
        self._ck(self.lib.blomgpu_download(name.encode()))
        return self.arrays[name]

    def download_async(self, name):
        """D2H on the copy stream, overlapping the routines called next; valid after sync()."""
        self._ck(self.lib.blomgpu_download_async(name.encode()))
        return self.arrays[name]

    def download_levels_async(self, name, koff, nlev):
        self._ck(self.lib.blomgpu_download_levels_async(name.encode(), koff, nlev))

    def upload_async(self, name, koff, nlev):
        """H2D of levels koff..koff+nlev-1 on the upload stream; wait_upload(name) before the first reader."""
        self._ck(self.lib.blomgpu_upload_async(name.encode(), koff, nlev))

    def wait_upload(self, name):
        self._ck(self.lib.blomgpu_wait_upload(name.encode()))

    def upload_all(self):
        self._ck(self.lib.blomgpu_upload_all())

    def download_all(self):
        self._ck(self.lib.blomgpu_download_all())

    def sync(self):
        self._ck(self.lib.blomgpu_sync())

    def fetch(self, name, nlev, dtype=np.float64):
        """Download a library-owned array (masks, tables) into a new numpy array."""
        a = np.zeros((nlev,) + self.shape2d, dtype=dtype)
        self.register(name, a, upload=False)
        return self.download(name)

    def set_option(self, key, value):
        self._ck(self.lib.blomgpu_set_option(key.encode(), str(value).encode()))

    def set_scalar(self, key, value):
        self._ck(self.lib.blomgpu_set_scalar(key.encode(), C.c_double(float(value))))

    def get_scalar(self, key):
        out = C.c_double()
        self._ck(self.lib.blomgpu_get_scalar(key.encode(), C.byref(out)))
        return out.value

    def set_scalars(self, **kw):
        for k, v in kw.items():
            self.set_scalar(k, v)

    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        self._ck(self.lib.blomgpu_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, uid: bytes):
        self._ck(self.lib.blomgpu_comm_init(C.c_char_p(uid), self.rank, self.nranks))

    def finalize(self):
        self._ck(self.lib.blomgpu_finalize())

    # -- mod_xc -----------------------------------------------------------------
    def xctilr(self, name, l1, ld, mh, nh, itype, koff=1):
        self._ck(self.lib.blomgpu_xctilr(name.encode(), koff, l1, ld, mh, nh, itype))

    def xcsum(self, name, mask="ip", lev=1):
        out = C.c_double()
        self._ck(self.lib.blomgpu_xcsum(name.encode(), lev, mask.encode(), C.byref(out)))
        return out.value

    def xcmax(self, name, mask="ip", lev=1):
        out = C.c_double()
        self._ck(self.lib.blomgpu_xcmax(name.encode(), lev, mask.encode(), C.byref(out)))
        return out.value

    def xcmin(self, name, mask="ip", lev=1):
        out = C.c_double()
        self._ck(self.lib.blomgpu_xcmin(name.encode(), lev, mask.encode(), C.byref(out)))
        return out.value

    def chksum(self, name, kcsd, itype):
        out = C.c_uint32()
        self._ck(self.lib.blomgpu_chksum(name.encode(), kcsd, itype, C.byref(out)))
        return out.value

    def chksum_at(self, name, koff, kcsd, itype):
        out = C.c_uint32()
        self._ck(self.lib.blomgpu_chksum_at(name.encode(), koff, kcsd, itype, C.byref(out)))
        return out.value

    # -- setup --------------------------------------------------------------------
    def bigrid(self, depth="depths"):
        self._ck(self.lib.blomgpu_bigrid(depth.encode()))

    def init_cppm(self):
        self._ck(self.lib.blomgpu_init_cppm())

    def inieos(self):
        self._ck(self.lib.blomgpu_inieos())

    def numerical_bounds(self):
        self._ck(self.lib.blomgpu_numerical_bounds())

    def init_fluxes(self, m, n, mm, nn, k1m, k1n):
        self._ck(self.lib.blomgpu_init_fluxes(m, n, mm, nn, k1m, k1n))

    # -- hot path, reference names ------------------------------------------------
    def tmsmt1(self, nn):
        self._ck(self.lib.blomgpu_tmsmt1(nn))

    def tmsmt2(self, m, mm, nn, k1m):
        self._ck(self.lib.blomgpu_tmsmt2(m, mm, nn, k1m))

    def _six(self, fn, m, n, mm, nn, k1m, k1n):
        self._ck(fn(m, n, mm, nn, k1m, k1n))

    def difest_halos(self, *a):
        self._six(self.lib.blomgpu_difest_halos, *a)

    def eddtra(self, *a):
        self._six(self.lib.blomgpu_eddtra, *a)

    def advect(self, *a):
        self._six(self.lib.blomgpu_advect, *a)

    def pbcor1(self, *a):
        self._six(self.lib.blomgpu_pbcor1, *a)

    def diffus(self, *a):
        self._six(self.lib.blomgpu_diffus, *a)

    def pgforc(self, *a):
        self._six(self.lib.blomgpu_pgforc, *a)

    def momtum(self, *a):
        self._six(self.lib.blomgpu_momtum, *a)

    def barotp(self, *a):
        self._six(self.lib.blomgpu_barotp, *a)

    def pbcor2(self, *a):
        self._six(self.lib.blomgpu_pbcor2, *a)

    def ndiff(self, *a):
        self._six(self.lib.blomgpu_ndiff, *a)

    def cmnfld2(self, *a):
        self._six(self.lib.blomgpu_cmnfld2, *a)

    def cmnfld_bfsqf_ale(self, *a):
        self._six(self.lib.blomgpu_cmnfld_bfsqf_ale, *a)

    def cmnfld_nslope_ale(self, *a):
        self._six(self.lib.blomgpu_cmnfld_nslope_ale, *a)

    def cmnfld_nnslope_ale(self, *a):
        self._six(self.lib.blomgpu_cmnfld_nnslope_ale, *a)

    def budget_init(self):
        """mass0 of budget_init (phy/mod_budget.F90:74-93)"""
        out = C.c_double()
        self._ck(self.lib.blomgpu_budget_init(C.byref(out)))
        return out.value

    def budget_sums(self, ncall, n, nn):
        """(sdp, tdp, trdp, sc) of budget_sums (phy/mod_budget.F90:95-196); nan = not evaluated"""
        out = (C.c_double * 4)(*([float("nan")] * 4))
        self._ck(self.lib.blomgpu_budget_sums(ncall, n, nn, out))
        return tuple(out)

    # -- instrumentation -----------------------------------------------------------
    def launch_count(self):
        return self.lib.blomgpu_launch_count()

    def launch_count_reset(self):
        self.lib.blomgpu_launch_count_reset()

    def timers_enable(self, on=True):
        self.lib.blomgpu_timers_enable(1 if on else 0)

    def timers_reset(self):
        self.lib.blomgpu_timers_reset()

    def timers(self):
        cap = 64
        names = ((C.c_char * 32) * cap)()
        ms = (C.c_double * cap)()
        calls = (C.c_long * cap)()
        launches = (C.c_long * cap)()
        n = self.lib.blomgpu_timers_get(cap, names, ms, calls, launches)
        return {names[i].value.decode(): {"ms": ms[i], "calls": calls[i], "launches": launches[i]}
                for i in range(n)}

    def ktimers_enable(self, on=True):
        self.lib.blomgpu_ktimers_enable(1 if on else 0)

    def ktimers(self):
        cap = 128
        names = ((C.c_char * 64) * cap)()
        ms = (C.c_double * cap)()
        launches = (C.c_long * cap)()
        n = self.lib.blomgpu_ktimers_get(cap, names, ms, launches)
        return {names[i].value.decode(): {"ms": ms[i], "launches": launches[i]} for i in range(n)}

    def stream(self):
        return self.lib.blomgpu_stream()


def time_levels(nstep: int, kk: int):
    """(m,n,mm,nn,k1m,k1n) of phy/mod_blom_step.F90:89-94."""
    m = nstep % 2 + 1
    n = (nstep + 1) % 2 + 1
    mm = (m - 1) * kk
    nn = (n - 1) * kk
    return m, n, mm, nn, 1 + mm, 1 + nn
