"""Host driver of the hot path: the part of `blom_step` (phy/mod_blom_step.F90:74-324)
that this package replaces, on synthetic state.

`HotPath` owns one BlomGpu tile, builds its band of the synthetic state, and runs
the reference's call order for the routines on the path:

    init_fluxes -> tmsmt1 -> [difest_halos] -> [ndiff] -> eddtra -> advect -> pbcor1 -> diffus
                -> pgforc -> momtum -> barotp -> pbcor2 -> tmsmt2

Routines that the reference runs in between (ALE regrid, cmnfld2, difest, column
physics, forcing) are out of scope; the halo refreshes those routines would have
done for the path (phy/mod_difest.F90:826-831, phy/mod_cmnfld_routines.F90:1171-1172)
are issued through `blomgpu_difest_halos` so that the chain stays valid on device.
`ndiff` (neutral diffusion, called from inside the ALE slice pipeline,
phy/mod_ale_regrid_remap.F90:1607-1690) is on the path whenever ltedtp='neutral',
which is the reference's default for the hybrid coordinate.

Option sets follow the reference's namelist defaults per vertical coordinate and grid
(cime_config/namelist_definition_blom.xml:367-460,1548-1575,1851-1862), see
`reference_options`.
"""
from __future__ import annotations

import os

import numpy as np

from . import synth
from .lib import BlomGpu, BlomGpuError, time_levels  # noqa: F401

# reference order of one baroclinic step (phy/mod_blom_step.F90:96-227); `ndiff` runs only with
# ltedtp='neutral', inside ale_regrid_remap (:128), i.e. between tmsmt1 and eddtra
STEP_SEQUENCE = ["init_fluxes", "tmsmt1", "ndiff", "eddtra", "advect", "pbcor1", "diffus", "pgforc", "momtum",
                 "barotp", "pbcor2", "tmsmt2"]
# prognostic arrays that cross the host<->device boundary in the end-to-end leg
IO_FIELDS = ("dp", "temp", "saln", "u", "v", "trc")


def reference_options(config: str, vcoord: str = "cntiso_hybrid") -> dict:
    """Namelist defaults the reference's build system writes for this grid and vertical coordinate
    (cime_config/namelist_definition_blom.xml: mommth :367, pgfmth :377, bmcmth :388, advmth :399,
    mlrmth :440, eitmth :1548, ltedtp :1851-1862)."""
    if vcoord == "isopyc_bulkml":
        return {"vcoord": vcoord, "mommth": "enscon", "pgfmth": "geopotential", "bmcmth": "uc", "advmth": "remap",
                "mlrmth": "fox08", "eitmth": "gm", "ltedtp": "layer"}
    return {"vcoord": vcoord, "mommth": "enscon", "pgfmth": "dynamic enthalpy", "bmcmth": "dluc", "advmth": "cppm",
            "mlrmth": "bod23", "eitmth": "gm", "ltedtp": "layer" if config == "tnx0.125v4" else "neutral"}


def step_routines(options: dict, routines=None):
    """The routines of STEP_SEQUENCE that run under this option set (optionally restricted)."""
    seq = [r for r in STEP_SEQUENCE if r != "ndiff" or options.get("ltedtp") == "neutral"]
    return [r for r in seq if routines is None or r in routines]


def run_step(b, routines, levels):
    """One pass of the hot path on backend `b` (BlomGpu, or the test oracle which mirrors its
    interface).  Same call order and arguments as phy/mod_blom_step.F90:96-227."""
    m, n, mm, nn, k1m, k1n = levels
    for r in routines:
        if r == "tmsmt1":
            b.tmsmt1(nn)
            b.difest_halos(m, n, mm, nn, k1m, k1n)
        elif r == "tmsmt2":
            b.tmsmt2(m, mm, nn, k1m)
        else:
            getattr(b, r)(m, n, mm, nn, k1m, k1n)


def band(jtdm: int, rank: int, nranks: int):
    """Contiguous j-band of `rank`: the first bands get the remainder rows."""
    base, rem = divmod(jtdm, nranks)
    jj = base + (1 if rank < rem else 0)
    j0 = rank * base + min(rank, rem)
    return j0, jj


def weighted_bands(row_cost, nranks: int, min_rows: int = 8):
    """Contiguous j-bands of (nearly) equal summed cost: [(j0, jj)] for rank 0..nranks-1.  The boundary after
    band r is the row where the running cost crosses (r+1)/nranks of the total; every band keeps at least
    `min_rows` rows (the halo of a band must come from its direct neighbour)."""
    c = np.cumsum(np.asarray(row_cost, dtype=np.float64))
    jtdm = len(c)
    cuts = [0]
    for r in range(1, nranks):
        j = int(np.searchsorted(c, c[-1] * r / nranks)) + 1
        j = max(j, cuts[-1] + min_rows)
        j = min(j, jtdm - (nranks - r) * min_rows)
        cuts.append(j)
    cuts.append(jtdm)
    return [(cuts[r], cuts[r + 1] - cuts[r]) for r in range(nranks)]


def balanced_band(config: str, rank: int, nranks: int):
    """The j-band of `rank` for a named grid.  One tile, or grids too small to matter: equal row counts.  Else
    the rows are weighted with their share of wet columns: neutral diffusion (60 % of the step under the
    reference's option set) works per wet face column and most other kernels skip land, so bands of equal
    height differ by up to 18 % in work on the 0.25 degree grid at 8 ranks.  The weights come from the
    synthetic bathymetry, which every rank can evaluate for the whole grid (a pure function of the indices)."""
    itdm, jtdm, kdm, nreg, baclin, batrop = synth.CONFIGS[config]
    if nranks == 1 or jtdm < 64 * nranks or config == "fuk95_analytic":
        return band(jtdm, rank, nranks)
    wet = (synth.Synth(itdm, jtdm, 1, nreg, baclin=baclin, batrop=batrop).depth_global > 0).sum(axis=1)
    cost = 0.3 + 0.7 * wet / max(wet.mean(), 1.0)
    return weighted_bands(cost, nranks)[rank]


class HotPath:
    def __init__(self, config="tnx1v4", *, ntr=0, nstep=1, rank=0, nranks=1, device=0, parity=True,
                 comm_uid: bytes | None = None, routines=None, seed=20240611, options=None,
                 pinned_alloc=None):
        itdm, jtdm, kdm, nreg, baclin, batrop = synth.CONFIGS[config]
        self.config, self.kdm, self.ntr = config, kdm, ntr
        self.itdm, self.jtdm, self.nreg = itdm, jtdm, nreg
        self.baclin = baclin
        self.rank, self.nranks = rank, nranks
        j0, jj = balanced_band(config, rank, nranks)
        self.j0, self.jj = j0, jj
        # namelist options: reference defaults for this grid, then BLOM_OPTIONS="key=value,..." (development
        # A/B switches, e.g. momtum_form=staged), then the caller's
        env_opts = dict(kv.split("=", 1) for kv in os.environ.get("BLOM_OPTIONS", "").split(",") if "=" in kv)
        self.options = {**reference_options(config), **env_opts, **(options or {})}
        self.routines = step_routines(self.options, routines)
        self.syn = synth.make_synth(config, ntr=ntr, j0=j0, jj=jj, seed=seed)
        self.grid = self.syn.grid()
        self.state = self.syn.state(self.grid)
        if pinned_alloc is not None:  # e2e leg: prognostic state lives in pinned host memory
            for nm in IO_FIELDS:
                if nm in self.state:
                    buf = pinned_alloc(self.state[nm].shape)
                    buf[...] = self.state[nm]
                    self.state[nm] = buf
        self.nstep = nstep
        self.scalars = self.syn.scalars(nstep)
        self.levels = time_levels(nstep, kdm)
        g = self.gpu = BlomGpu(itdm, jtdm, kdm, nreg, ntr=ntr, j0=j0, jj=jj, rank=rank, nranks=nranks,
                               device=device, parity=parity)
        if nranks > 1:
            if comm_uid is None:
                raise ValueError("multi-rank HotPath needs the NCCL unique id")
            g.comm_init(comm_uid)
        for k, v in self.options.items():
            g.set_option(k, v)
        self.arrays = {**self.grid, **self.state}
        g.register_all(self.arrays)
        g.set_scalars(**self.scalars)
        synth.fill_halos(g, self.arrays)
        g.bigrid("depths")
        g.download_all()
        self.masks = {k: g.fetch(k, 1, np.int32)[0].copy() for k in ("ip", "iu", "iv", "iq")}
        synth.derive(self.grid, self.state, self.masks, self.levels, self.scalars, g,
                     sync_in=lambda names: [g.upload(n) for n in names],
                     sync_out=lambda names: [g.download(n) for n in names])
        g.upload_all()
        if "ndiff" in self.routines:
            # products of the (out-of-scope) ALE slice pipeline that neutral diffusion consumes
            nd = synth.ndiff_inputs(self.syn, self.state, self.levels, ntr=ntr)
            g.register_all(nd)
            self.arrays.update(nd)
        self.setup()

    def setup(self):
        g = self.gpu
        g.inieos()
        g.numerical_bounds()
        if "advect" in self.routines:
            g.init_cppm()
        g.sync()

    @property
    def cells(self):
        """interior (i,j,k) cells of the GLOBAL grid"""
        return self.itdm * self.jtdm * self.kdm

    def set_step(self, nstep):
        self.nstep = nstep
        self.levels = time_levels(nstep, self.kdm)
        self.gpu.set_scalar("nstep", nstep)

    # fields no routine after `momtum` writes (barotp works on ub/vb/pb, pbcor2 and tmsmt2 on
    # dp/T/S/trc, phy/mod_blom_step.F90:169-227): their download can overlap the rest of the step
    FINAL_AFTER = {"momtum": ("u", "v")}

    def step(self, early_download=False):
        """One pass of the hot path over the resident state.  early_download: start the device ->
        host copy of a field on the copy stream as soon as its last writer has been enqueued."""
        g = self.gpu
        self._early = set()
        if not early_download:
            run_step(g, self.routines, self.levels)
            return
        for r in self.routines:
            run_step(g, [r], self.levels)
            for nm in self.FINAL_AFTER.get(r, ()):
                if nm in self.arrays:
                    g.download_async(nm)
                    self._early.add(nm)

    def step_pipelined(self):
        """One end-to-end pass with the host<->device copies overlapped with the kernels (the call sequence a
        host that keeps its own copy of the prognostic state would issue):
          H2D  the time level the host physics wrote (new level n) of T,S[,trc], dp, then u,v, on the upload
               stream; ndiff (which only reads T,S of that level and is independent of tmsmt1) starts as soon as
               T,S are there, tmsmt1 follows when dp has arrived, u,v are first read by momtum (advect only reads
               the resident mid level), so their copy hides behind ndiff..pgforc.  The u,v halo refresh of difest
               (phy/mod_difest.F90:826-827) moves with them to just before momtum - nothing reads or writes
               u,v in between, so the values are the ones run_step produces.
          D2H  on the copy stream as soon as a field's last writer has been enqueued: u,v (both levels,
               momtum also rewrites the mid level) after momtum; dp,T,S[,trc] new level after pbcor2, mid
               level after tmsmt2 (phy/mod_blom_step.F90:169-227).
        Ends with everything on the host (sync).  Falls back to upload / step / download when the routine set
        is restricted."""
        g, kk = self.gpu, self.kdm
        m, n, mm, nn, k1m, k1n = self.levels
        need = ("tmsmt1", "momtum", "pbcor2", "tmsmt2")
        if any(r not in self.routines for r in need):
            self.upload_inputs(); self.step(early_download=True); self.download_outputs()
            return
        from .lib import HALO_PS, HALO_UV, HALO_VV, HALO_US, HALO_VS
        # upload order = order of first use: T, S (neutral diffusion, the longest routine, reads only them), tracers,
        # then dp (tmsmt1), then u, v (momtum)
        scal = [("temp", 0), ("saln", 0)] + [("trc", nt * 2 * kk) for nt in range(self.ntr)] + [("dp", 0)]
        for nm, off in scal:
            g.upload_async(nm, off + k1n, kk)
        for nm in ("u", "v"):
            g.upload_async(nm, k1n, kk)
        nd_first = "ndiff" in self.routines
        for r in self.routines:
            if r == "tmsmt1":
                for nm in ("temp", "saln", "trc"):
                    if nm in self.arrays:
                        g.wait_upload(nm)
                g.xctilr("temp", 1, 2 * kk, 3, 3, HALO_PS)
                g.xctilr("saln", 1, 2 * kk, 3, 3, HALO_PS)
                if nd_first:
                    # ndiff and tmsmt1 touch disjoint arrays (tmsmt1 copies dp,T,S of the new level into the
                    # smoother's work arrays, phy/mod_tmsmt.F90:209-258), so ndiff can start while dp is on its way
                    g.ndiff(m, n, mm, nn, k1m, k1n)
                g.wait_upload("dp")
                g.tmsmt1(nn)
                for nm, it in (("ubflxs_p", HALO_UV), ("vbflxs_p", HALO_VV), ("pbu", HALO_US), ("pbv", HALO_VS)):
                    g.xctilr(nm, 1, 2, 2, 2, it)
            elif r == "ndiff" and nd_first:
                continue
            elif r == "momtum":
                g.wait_upload("u"); g.wait_upload("v")
                g.xctilr("u", 1, 2 * kk, 2, 2, HALO_UV)
                g.xctilr("v", 1, 2 * kk, 2, 2, HALO_VV)
                g.momtum(m, n, mm, nn, k1m, k1n)
                g.download_async("u"); g.download_async("v")
            elif r == "pbcor2":
                g.pbcor2(m, n, mm, nn, k1m, k1n)
                for nm, off in scal:
                    g.download_levels_async(nm, off + k1n, kk)
            elif r == "tmsmt2":
                g.tmsmt2(m, mm, nn, k1m)
                for nm, off in scal:
                    g.download_levels_async(nm, off + k1m, kk)
            else:
                run_step(g, [r], self.levels)
        g.sync()

    def io_bytes_pipelined(self):
        """(H2D, D2H) bytes of one step_pipelined(): the new level up, both levels down."""
        b = sum(self.arrays[nm].nbytes for nm in IO_FIELDS if nm in self.arrays)
        return b // 2, b

    def advance(self, early_download=False):
        """step() followed by the leap-frog role swap of the time levels."""
        self.step(early_download)
        self.set_step(self.nstep + 1)

    def upload_inputs(self):
        for nm in IO_FIELDS:
            if nm in self.arrays:
                self.gpu.upload(nm)

    def download_outputs(self):
        """Bring the prognostic state back; fields whose copy advance(early_download=True) already
        started are only waited for."""
        for nm in IO_FIELDS:
            if nm in self.arrays and nm not in getattr(self, "_early", ()):
                self.gpu.download(nm)
        self.gpu.sync()
        self._early = set()

    def io_bytes(self):
        b = sum(self.arrays[nm].nbytes for nm in IO_FIELDS if nm in self.arrays)
        return b, b

    def finalize(self):
        self.gpu.finalize()
