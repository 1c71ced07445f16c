"""Seeded synthetic ocean state on the reference's grid dimensions.

Every value is a pure function of (seed, field id, level, GLOBAL j, GLOBAL i),
so a j-band of a multi-GPU run holds exactly the rows of the one-tile state and
the same arrays can be fed to the CUDA path and to the test oracle
(SURVEY.md §8d).  Only interiors are generated; halos are filled afterwards with
xctilr of the proper grid type (`fill_halos`).

Grids (itdm, jtdm, kdm, nreg): bld/<grid>/ of the reference, see CONFIGS.
"""
from __future__ import annotations

import numpy as np

from .lib import (HALO_PS, HALO_PV, HALO_QS, HALO_QV, HALO_US, HALO_UV, HALO_VS, HALO_VV)

ONEM = 9806.0

# name: (itdm, jtdm, kdm, nreg, baclin, batrop)
CONFIGS = {
    "tiny0": (20, 18, 4, 0, 1800.0, 36.0),
    "tiny1": (24, 18, 5, 1, 1800.0, 36.0),
    "tiny2": (24, 20, 5, 2, 1800.0, 36.0),
    "tiny3": (20, 18, 4, 3, 1800.0, 36.0),
    "tiny4": (20, 22, 4, 4, 1800.0, 36.0),
    "tiny2k53": (24, 20, 53, 2, 1800.0, 36.0),      # tiny2 with the production layer count (kdm-dependent code paths)
    "mid1": (48, 44, 6, 1, 1800.0, 36.0),          # multi-band parity cases (periodic-i / tripolar)
    "mid2": (48, 46, 6, 2, 1800.0, 36.0),
    "fuk95": (156, 32, 12, 4, 180.0, 6.0),        # tests/fuk95/limits:131-143 (dims only, noisy state)
    "fuk95_analytic": (156, 32, 12, 4, 180.0, 6.0),  # analytic geometry + jet of fuk95/mod_fuk95.F90 (fuk95.py)
    "channel": (208, 512, 53, 1, 1800.0, 36.0),    # bld/channel/patch.input.1
    "tnx1v4": (360, 385, 53, 2, 3200.0, 64.0),     # namelist_definition_blom.xml:179-201
    "tnx0.25v4": (1440, 1153, 53, 2, 900.0, 15.0),
    "tnx0.25v4_band": (1440, 64, 53, 1, 900.0, 15.0),   # a closed 64-row band at full zonal width (parity at the tile widths
                                                        # the 0.25 degree kernels run with; also the CPU-baseline sample)
    "tnx0.125v4": (2880, 2165, 53, 2, 300.0, 6.0),
}

# xctilr grid/field type of every array the hot path touches
ITYPE = {
    # p-points
    "depths": HALO_PS, "scpx": HALO_PS, "scpy": HALO_PS, "scp2": HALO_PS, "scp2i": HALO_PS,
    "dp": HALO_PS, "temp": HALO_PS, "saln": HALO_PS, "sigma": HALO_PS, "p": HALO_PS, "phi": HALO_PS,
    "pb": HALO_PS, "pb_p": HALO_PS, "sealv": HALO_PS, "trc": HALO_PS, "difint": HALO_PS,
    "difiso": HALO_PS, "difwgt": HALO_PS, "coriop": HALO_PS, "pbath": HALO_PS,
    "hbl_tf": HALO_PS, "wpup_tf": HALO_PS, "hml_tf1": HALO_PS, "hml_tf": HALO_PS, "hml_tfbnd": HALO_PS,
    "ustar3": HALO_PS, "wstar3": HALO_PS, "util1": HALO_PS,
    "dpold": HALO_PS, "told": HALO_PS, "sold": HALO_PS, "trcold": HALO_PS, "pb_mn": HALO_PS, "mld": HALO_PS, "OBLdepth": HALO_PS,
    # u-points
    "scux": HALO_US, "scuy": HALO_US, "scu2": HALO_US, "scuxi": HALO_US, "scuyi": HALO_US,
    "u": HALO_UV, "dpu": HALO_US, "pu": HALO_US, "uflx": HALO_UV, "utflx": HALO_UV, "usflx": HALO_UV,
    "cau": HALO_UV, "ub": HALO_UV, "pbu": HALO_US, "pbu_p": HALO_US, "ubflx": HALO_UV,
    "ubflxs": HALO_UV, "ubflxs_p": HALO_UV, "ubcors_p": HALO_UV, "umfltd": HALO_UV, "umflsm": HALO_UV,
    "utfltd": HALO_UV, "utflsm": HALO_UV, "utflld": HALO_UV, "usfltd": HALO_UV, "usflsm": HALO_UV,
    "usflld": HALO_UV, "umax": HALO_US, "taux": HALO_UV, "nslpx": HALO_US, "pgfx": HALO_UV,
    "pgfxm": HALO_UV, "xixp": HALO_US, "xixm": HALO_US, "dpuold": HALO_US, "uja": HALO_UV, "ujb": HALO_UV,
    "pgfx_o": HALO_UV, "pgfxm_o": HALO_UV, "ubflx_mn": HALO_UV, "utotn": HALO_UV, "xixp_o": HALO_US, "xixm_o": HALO_US, "ubrhs": HALO_UV,
    "mu_nonloc": HALO_UV, "uflux": HALO_UV, "uflux2": HALO_UV, "uflux3": HALO_UV,
    # v-points
    "scvx": HALO_VS, "scvy": HALO_VS, "scv2": HALO_VS, "scvxi": HALO_VS, "scvyi": HALO_VS,
    "v": HALO_VV, "dpv": HALO_VS, "pv": HALO_VS, "vflx": HALO_VV, "vtflx": HALO_VV, "vsflx": HALO_VV,
    "cav": HALO_VV, "vb": HALO_VV, "pbv": HALO_VS, "pbv_p": HALO_VS, "vbflx": HALO_VV,
    "vbflxs": HALO_VV, "vbflxs_p": HALO_VV, "vbcors_p": HALO_VV, "vmfltd": HALO_VV, "vmflsm": HALO_VV,
    "vtfltd": HALO_VV, "vtflsm": HALO_VV, "vtflld": HALO_VV, "vsfltd": HALO_VV, "vsflsm": HALO_VV,
    "vsflld": HALO_VV, "vmax": HALO_VS, "tauy": HALO_VV, "nslpy": HALO_VS, "pgfy": HALO_VV,
    "pgfym": HALO_VV, "xiyp": HALO_VS, "xiym": HALO_VS, "dpvold": HALO_VS, "via": HALO_VV, "vib": HALO_VV,
    "pgfy_o": HALO_VV, "pgfym_o": HALO_VV, "vbflx_mn": HALO_VV, "vtotn": HALO_VV, "xiyp_o": HALO_VS, "xiym_o": HALO_VS, "vbrhs": HALO_VV,
    "mv_nonloc": HALO_VV, "vflux": HALO_VV, "vflux2": HALO_VV, "vflux3": HALO_VV,
    # q-points
    "scqx": HALO_QS, "scqy": HALO_QS, "scq2": HALO_QS, "scq2i": HALO_QS, "corioq": HALO_QS,
    "pvtrop": HALO_QS, "pvtrop_o": HALO_QS,
}


def _mix(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on uint64 arrays."""
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


class Synth:
    """Synthetic state of one j-band [j0+1, j0+jj] of a (itdm, jtdm, kdm) grid."""

    def __init__(self, itdm, jtdm, kdm, nreg, *, ntr=0, nbdy=4, j0=0, jj=None, seed=20240611,
                 baclin=1800.0, batrop=36.0, land=True, metric="tripolar"):
        self.itdm, self.jtdm, self.kdm, self.nreg, self.ntr, self.nb = itdm, jtdm, kdm, nreg, ntr, nbdy
        self.j0 = j0
        self.jj = jtdm if jj is None else jj
        self.seed = seed
        self.baclin, self.batrop = baclin, batrop
        self.land, self.metric = land, metric
        self.ldi, self.ldj = itdm + 2 * nbdy, self.jj + 2 * nbdy
        self._fid = 0
        self._cubes = {}
        with np.errstate(over="ignore"):
            self._build_geometry()

    @classmethod
    def from_config(cls, name, **kw):
        itdm, jtdm, kdm, nreg, baclin, batrop = CONFIGS[name]
        return cls(itdm, jtdm, kdm, nreg, baclin=baclin, batrop=batrop, **kw)

    # ---- helpers ---------------------------------------------------------------
    def zeros(self, nlev=1, dtype=np.float64):
        return np.zeros((nlev, self.ldj, self.ldi), dtype=dtype)

    def interior(self, a):
        nb = self.nb
        return a[..., nb:nb + self.jj, nb:nb + self.itdm]

    def _hash_uniform(self, nlev, fid):
        """U[0,1) on the band interior, shape (nlev, jj, itdm), keyed on global indices."""
        jg = np.arange(self.jj, dtype=np.uint64) + np.uint64(self.j0)
        ig = np.arange(self.itdm, dtype=np.uint64)
        k = np.arange(nlev, dtype=np.uint64)
        with np.errstate(over="ignore"):
            key = (np.uint64(self.seed) * np.uint64(0x9E3779B97F4A7C15)
                   + np.uint64(fid) * np.uint64(0xD1B54A32D192ED03))
            x = (key + jg[None, :, None] * np.uint64(0xA24BAED4963EE407)
                 + ig[None, None, :] * np.uint64(0x9FB21C651E98DF25))
            x = _mix(x)                                       # (1, jj, itdm): one hash per column
            x = _mix(x + k[:, None, None] * np.uint64(0x8CB92BA72F3D8DD7))
        return (x >> np.uint64(11)).astype(np.float64) * 2.0 ** -53

    def _hash_uniform_padded(self, nlev, fid):
        """U[0,1) on the whole padded band (nlev, ldj, ldi): the interior hash evaluated at global
        indices wrapped periodically, so halo points equal the interior points they image wherever
        the domain is periodic and every j-band sees its neighbours' values."""
        nb = self.nb
        jg = ((np.arange(self.ldj, dtype=np.int64) - nb + self.j0) % self.jtdm).astype(np.uint64)
        ig = ((np.arange(self.ldi, dtype=np.int64) - nb) % self.itdm).astype(np.uint64)
        k = np.arange(nlev, dtype=np.uint64)
        with np.errstate(over="ignore"):
            key = (np.uint64(self.seed) * np.uint64(0x9E3779B97F4A7C15)
                   + np.uint64(fid) * np.uint64(0xD1B54A32D192ED03))
            x = (key + jg[None, :, None] * np.uint64(0xA24BAED4963EE407)
                 + ig[None, None, :] * np.uint64(0x9FB21C651E98DF25))
            x = _mix(x)
            x = _mix(x + k[:, None, None] * np.uint64(0x8CB92BA72F3D8DD7))
        return (x >> np.uint64(11)).astype(np.float64) * 2.0 ** -53

    def _base(self, kind):
        """Base noise cubes (kdm levels), hashed once; every field is a cheap re-indexing
        of one of them (still a pure function of global indices)."""
        if kind not in self._cubes:
            if kind == "u":
                self._cubes[kind] = self._hash_uniform(self.kdm, 1)
            else:
                u1 = self._hash_uniform(self.kdm, 2)
                u2 = self._hash_uniform(self.kdm, 3)
                self._cubes[kind] = np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * np.pi * u2)
        return self._cubes[kind]

    def _variant(self, kind, nlev):
        self._fid += 1
        f = self._fid
        base = self._base(kind)
        reps = -(-nlev // self.kdm)
        out = []
        for r in range(reps):
            sh_i = (f * 37 + r * 101) % self.itdm
            sh_k = (f * 5 + r * 3) % self.kdm
            out.append(np.roll(base, (sh_k, sh_i), axis=(0, 2)))
        a = out[0] if reps == 1 else np.concatenate(out, axis=0)
        return a[:nlev]

    def _uniform(self, nlev):
        return self._variant("u", nlev)

    def _normal(self, nlev):
        return self._variant("n", nlev)

    def _put(self, a, vals):
        self.interior(a)[...] = vals
        return a

    # ---- geometry (global 2-D, then sliced) ---------------------------------------
    def _build_geometry(self):
        itdm, jtdm, nreg = self.itdm, self.jtdm, self.nreg
        ig = np.arange(1, itdm + 1)[None, :]
        jg = np.arange(1, jtdm + 1)[:, None]
        x = 2 * np.pi * (ig - 0.5) / itdm
        y = (jg - 0.5) / jtdm
        depth = 3200.0 + 1500.0 * np.sin(3 * x + 0.7) * np.cos(2 * np.pi * 1.5 * y) \
            + 250.0 * np.cos(7 * x - 1.1) * np.sin(2 * np.pi * 4 * y + 0.3)
        depth = np.clip(depth, 200.0, 5000.0)
        if self.land:
            blob = np.sin(2 * x + 0.4) * np.sin(2 * np.pi * 1.25 * y + 0.9) \
                + 0.35 * np.sin(5 * x - 0.3) * np.cos(2 * np.pi * 3 * y)
            depth = np.where(blob > 0.62, 0.0, depth)
        # closed boundaries per region type (phy/mod_bigrid.F90:59-107)
        if nreg in (0, 1, 2):
            depth[0, :] = 0.0
        if nreg in (0, 1):
            depth[-1, :] = 0.0
        if nreg in (0, 4):
            depth[:, 0] = 0.0
            depth[:, -1] = 0.0
        if nreg == 2:
            depth[-3:, :] = np.maximum(depth[-3:, :], 300.0)  # open water at the fold
        # remove wet cells with three or more dry neighbours until none is left; every pass dries at
        # least one cell, so the loop ends for any band height (a short band may need many passes)
        for _ in range(itdm * jtdm + 1):
            if nreg == 2:
                depth[-1, :] = depth[-2, ::-1]  # p-grid fold: a(i,jj) = a(ii+1-i,jj-1)
            d = self._pad_global(depth)
            wet = d[1:-1, 1:-1] > 0
            nzero = ((d[1:-1, :-2] <= 0).astype(int) + (d[1:-1, 2:] <= 0) + (d[:-2, 1:-1] <= 0)
                     + (d[2:, 1:-1] <= 0))
            bad = wet & (nzero >= 3)
            if not bad.any():
                break
            depth[bad] = 0.0
        else:
            raise RuntimeError("synthetic bathymetry did not converge")
        self.depth_global = depth

        lat = -78.0 + 166.0 * y  # degrees, tripolar-like
        gs = 1.0e5 * 360.0 / itdm
        if self.metric == "uniform":
            fx = np.ones_like(y)
        else:
            fx = np.maximum(0.2, np.cos(np.deg2rad(lat)))
        wob = 1.0 + 0.05 * np.sin(2 * x + 0.2) * np.cos(2 * np.pi * y)
        self._scpx_g = gs * fx * wob
        self._scpy_g = gs * (1.0 + 0.03 * np.cos(x - 0.5) * np.sin(2 * np.pi * y + 0.1)) * np.ones_like(x)
        self._lat_g = lat * np.ones_like(x)

    def _pad_global(self, a):
        """1-wide ring of a global 2-D array according to nreg (land where closed)."""
        nreg = self.nreg
        p = np.zeros((a.shape[0] + 2, a.shape[1] + 2), dtype=a.dtype)
        p[1:-1, 1:-1] = a
        if nreg in (1, 2, 3):
            p[1:-1, 0] = a[:, -1]
            p[1:-1, -1] = a[:, 0]
        if nreg in (3, 4):
            p[0, 1:-1] = a[-1, :]
            p[-1, 1:-1] = a[0, :]
        if nreg == 3:
            p[0, 0], p[0, -1], p[-1, 0], p[-1, -1] = a[-1, -1], a[-1, 0], a[0, -1], a[0, 0]
        if nreg == 2:
            p[-1, 1:-1] = a[-3, ::-1]  # row jj+1 <- row jj-2 mirrored
            p[-1, 0], p[-1, -1] = p[-1, -2], p[-1, 1]
        return p

    def _band(self, g2d):
        return g2d[self.j0:self.j0 + self.jj, :]

    # ---- fields ---------------------------------------------------------------------
    def grid(self):
        """Metric arrays of mod_grid (phy/mod_grid.F90:48-88) + depths."""
        out = {}
        scpx, scpy = self._band(self._scpx_g), self._band(self._scpy_g)
        # staggered metrics: same smooth functions evaluated half a cell away
        def shift_i(a):
            return 0.5 * (a + np.roll(a, 1, axis=1))

        def shift_j(gfun):
            g = 0.5 * (gfun + np.vstack([gfun[:1], gfun[:-1]]))
            return self._band(g)
        scux, scuy = shift_i(scpx), shift_i(scpy)
        scvx, scvy = shift_j(self._scpx_g), shift_j(self._scpy_g)
        scqx, scqy = shift_i(scvx), shift_i(scvy)
        vals = {
            "depths": self._band(self.depth_global),
            "scpx": scpx, "scpy": scpy, "scux": scux, "scuy": scuy, "scvx": scvx, "scvy": scvy,
            "scqx": scqx, "scqy": scqy,
            "scp2": scpx * scpy, "scu2": scux * scuy, "scv2": scvx * scvy, "scq2": scqx * scqy,
        }
        vals["scp2i"] = 1.0 / vals["scp2"]
        vals["scq2i"] = 1.0 / vals["scq2"]
        vals["scuxi"] = 1.0 / scux
        vals["scuyi"] = 1.0 / scuy
        vals["scvxi"] = 1.0 / scvx
        vals["scvyi"] = 1.0 / scvy
        omega2 = 2.0 * 7.2921e-5
        latq = shift_i(shift_j(self._lat_g))
        vals["corioq"] = omega2 * np.sin(np.deg2rad(latq))
        vals["coriop"] = omega2 * np.sin(np.deg2rad(self._band(self._lat_g)))
        for k, v in vals.items():
            out[k] = self._put(self.zeros(), v)
        return out

    def masks_np(self):
        """ip/iu/iv on the band interior (generator-side only; the product computes its
        own masks with blomgpu_bigrid)."""
        d = self._pad_global(self.depth_global)
        ip = (d > 0).astype(np.int32)
        iu = ip[1:-1, 1:-1] * ip[1:-1, :-2]
        iv = ip[1:-1, 1:-1] * ip[:-2, 1:-1]
        return self._band(ip[1:-1, 1:-1]), self._band(iu), self._band(iv)

    def state(self, grid):
        """Prognostic + auxiliary arrays for the whole hot path."""
        kk, nb = self.kdm, self.nb
        ipm, ium, ivm = self.masks_np()
        depth = self._band(self.depth_global)
        st = {}
        # --- layer thickness: positive partition of pb with massless layers near the bottom
        pb = depth * ONEM
        w = 0.2 + self._uniform(kk)
        kbot = np.floor(kk * (0.7 + 0.3 * self._uniform(1)[0])).astype(int)
        kbot = np.clip(kbot, 1, kk)
        kidx = np.arange(1, kk + 1)[:, None, None]
        w = np.where(kidx > kbot[None], 0.0, w)
        w[0] += 0.05
        dpm = pb[None] * w / w.sum(axis=0, keepdims=True)
        pert = 1.0 + 2.0e-3 * (self._uniform(kk) - 0.5)
        dpn = dpm * pert
        dpn *= np.where(pb > 0, pb / np.maximum(dpn.sum(axis=0), 1e-30), 0.0)[None]
        dp = self.zeros(2 * kk)
        self.interior(dp)[:kk] = dpm
        self.interior(dp)[kk:] = dpn
        st["dp"] = dp
        # --- temperature / salinity / tracers
        zfrac = (np.arange(kk) + 0.5)[:, None, None] / kk
        for nm, base, amp in (("temp", 25.0 * (1 - zfrac) ** 1.5, 0.1), ("saln", 34.0 + 2.0 * zfrac, 0.05)):
            a = self.zeros(2 * kk)
            lvl = base + amp * (self._uniform(kk) - 0.5) + 0.5 * np.sin(
                2 * np.pi * (np.arange(self.itdm)[None, None, :] / self.itdm))
            self.interior(a)[:kk] = lvl * ipm
            self.interior(a)[kk:] = (lvl + 0.02 * (self._uniform(kk) - 0.5)) * ipm
            st[nm] = a
        if self.ntr > 0:
            trc = self.zeros(2 * kk * self.ntr)
            for nt in range(self.ntr):
                lvl = np.maximum(0.0, 1.0 + 0.5 * np.cos(3 * np.pi * zfrac) + 0.3 * (self._uniform(kk) - 0.5))
                self.interior(trc)[nt * 2 * kk:nt * 2 * kk + kk] = lvl * ipm
                self.interior(trc)[nt * 2 * kk + kk:(nt + 1) * 2 * kk] = lvl * ipm
            st["trc"] = trc
        st["sigma"] = self.zeros(2 * kk)
        # --- velocities
        for nm, msk in (("u", ium), ("v", ivm)):
            a = self.zeros(2 * kk)
            base = 0.05 * self._normal(kk)
            self.interior(a)[:kk] = base * msk
            self.interior(a)[kk:] = (base + 0.005 * self._normal(kk)) * msk
            st[nm] = a
        # --- flux accumulators (init_fluxes zeroes them each step)
        for nm in ("uflx", "vflx", "utflx", "vtflx", "usflx", "vsflx"):
            st[nm] = self.zeros(2 * kk)
        for nm in ("umfltd", "vmfltd", "umflsm", "vmflsm", "utfltd", "vtfltd", "utflsm", "vtflsm",
                   "utflld", "vtflld", "usfltd", "vsfltd", "usflsm", "vsflsm", "usflld", "vsflld"):
            st[nm] = self.zeros(2 * kk)
        # small eddy-induced mass fluxes so that advect's (umfltd+umflsm)/dpu term is exercised
        scale = 1.0e-4 * ONEM * grid["scuy"][0, nb:nb + self.jj, nb:nb + self.itdm] * self.baclin
        self.interior(st["umfltd"])[:] = np.tile(scale * self._normal(kk) * ium, (2, 1, 1))
        scale = 1.0e-4 * ONEM * grid["scvx"][0, nb:nb + self.jj, nb:nb + self.itdm] * self.baclin
        self.interior(st["vmfltd"])[:] = np.tile(scale * self._normal(kk) * ivm, (2, 1, 1))
        st["cau"] = self.zeros(kk)
        st["cav"] = self.zeros(kk)
        for nm in ("p", "pu", "pv", "phi"):
            st[nm] = self.zeros(kk + 1)
        st["dpu"] = self.zeros(2 * kk)
        st["dpv"] = self.zeros(2 * kk)
        for nm in ("pb", "pbu", "pbv", "ub", "vb", "ubflxs_p", "vbflxs_p"):
            st[nm] = self.zeros(2)
        for nm in ("ubflxs", "vbflxs"):
            st[nm] = self.zeros(3)
        for nm in ("pb_p", "pbu_p", "pbv_p", "ubcors_p", "vbcors_p", "sealv", "umax", "vmax"):
            st[nm] = self.zeros(1)
        self.interior(st["pb"])[0] = dpm.sum(axis=0)
        self.interior(st["pb"])[1] = dpn.sum(axis=0)
        self.interior(st["pb_p"])[0] = dpn.sum(axis=0)
        # predicted barotropic mass flux sums: small, masked
        su = grid["scuy"][0, nb:nb + self.jj, nb:nb + self.itdm]
        sv = grid["scvx"][0, nb:nb + self.jj, nb:nb + self.itdm]
        self.interior(st["ubflxs_p"])[:] = (2.0e-3 * self._normal(2)) * su * pb[None] * ium / max(self.batrop, 1.0) * 0.1
        self.interior(st["vbflxs_p"])[:] = (2.0e-3 * self._normal(2)) * sv * pb[None] * ivm / max(self.batrop, 1.0) * 0.1
        # lateral diffusivities and friends
        st["difint"] = self._put(self.zeros(kk), 100.0 + 1400.0 * self._uniform(kk))
        st["difiso"] = self._put(self.zeros(kk), 100.0 + 1400.0 * self._uniform(kk))
        st["difwgt"] = self._put(self.zeros(1), self._uniform(1))
        st["nslpx"] = self._put(self.zeros(kk), 1.0e-4 * self._normal(kk) * ium)
        st["nslpy"] = self._put(self.zeros(kk), 1.0e-4 * self._normal(kk) * ivm)
        # --- eddy-induced transport: boundary/mixed layer inputs from the (out-of-scope) column
        # physics and the running-mean filter state of mod_eddtra (phy/mod_eddtra.F90:49-50)
        st["OBLdepth"] = self._put(self.zeros(1), (10.0 + 90.0 * self._uniform(1)) * ipm)
        st["mld"] = self._put(self.zeros(1), (10.0 + 140.0 * self._uniform(1)) * ipm)
        st["ustar3"] = self._put(self.zeros(1), 2.0e-6 * self._uniform(1) * ipm)
        st["wstar3"] = self._put(self.zeros(1), 1.0e-6 * self._uniform(1) * ipm)
        st["hbl_tf"] = self._put(self.zeros(1), (20.0 + 60.0 * self._uniform(1)) * ipm)
        st["wpup_tf"] = self._put(self.zeros(1), (1.0e-3 + 1.0e-4 * self._uniform(1)) * ipm)
        st["hml_tf1"] = self._put(self.zeros(1), (20.0 + 100.0 * self._uniform(1)) * ipm)
        st["hml_tf"] = self._put(self.zeros(1), (20.0 + 100.0 * self._uniform(1)) * ipm)
        st["hml_tfbnd"] = self.zeros(1)
        st["util1"] = self.zeros(1)
        st["taux"] = self._put(self.zeros(1), 0.1 * self._normal(1) * ium)
        st["tauy"] = self._put(self.zeros(1), 0.1 * self._normal(1) * ivm)
        # --- time smoother work arrays (phy/mod_tmsmt.F90:53-66)
        st["dpold"] = self.zeros(2 * kk)
        self.interior(st["dpold"])[:] = self.interior(dp) * (1.0 + 1.0e-3 * (self._uniform(2 * kk) - 0.5))
        st["told"] = self._put(self.zeros(kk), self.interior(st["temp"])[:kk] + 0.01)
        st["sold"] = self._put(self.zeros(kk), self.interior(st["saln"])[:kk] + 0.001)
        st["dpuold"] = self.zeros(kk)
        st["dpvold"] = self.zeros(kk)
        if self.ntr > 0:
            st["trcold"] = self.zeros(kk * self.ntr)
            for nt in range(self.ntr):
                self.interior(st["trcold"])[nt * kk:(nt + 1) * kk] = \
                    self.interior(st["trc"])[nt * 2 * kk:nt * 2 * kk + kk] * 1.01
        # --- pressure gradient force arrays (phy/mod_pgforc.F90:51-80); bottom geopotential
        self.interior(st["phi"])[kk] = -9.806 * depth
        for nm, msk in (("pgfx", ium), ("pgfy", ivm)):
            st[nm] = self._put(self.zeros(2 * kk), 1.0e-2 * self._normal(2 * kk) * msk)
        st["pgfx_o"] = self.zeros(kk)
        st["pgfy_o"] = self.zeros(kk)
        for nm, msk, amp in (("pgfxm", ium, 1e-2), ("pgfym", ivm, 1e-2), ("xixp", ium, 1e-7), ("xixm", ium, 1e-7),
                             ("xiyp", ivm, 1e-7), ("xiym", ivm, 1e-7)):
            base = 1.0e-3 / 9806.0 if nm.startswith("xi") else 0.0
            st[nm] = self._put(self.zeros(2), (base + amp * 1e-3 * self._normal(2)) * msk)
            st[nm + "_o"] = self._put(self.zeros(1), (base + amp * 1e-3 * self._normal(1)) * msk)
        # --- barotropic solver arrays (phy/mod_barotp.F90:57-67)
        pbm = dpm.sum(axis=0)
        st["pb_mn"] = self._put(self.zeros(2), np.stack([pbm, pbm * (1.0 + 1e-6 * (self._uniform(1)[0] - 0.5))]))
        for nm, msk, sc_ in (("ubflx", ium, su), ("vbflx", ivm, sv)):
            base = 1.0e-3 * self._normal(1)[0] * sc_ * pb * msk
            st[nm] = self._put(self.zeros(2), np.stack([base, base]))
            st[nm + "_mn"] = self._put(self.zeros(2), np.stack([base, base * 0.999]))
        omega2 = 2.0 * 7.2921e-5
        latq = self.interior(grid["corioq"])[0]
        pvq = np.where(pb > 0, latq / np.maximum(pb, 1.0), 0.0)
        st["pvtrop"] = self._put(self.zeros(2), np.stack([pvq, pvq]))
        st["pvtrop_o"] = self._put(self.zeros(1), pvq)
        # --- momentum: bounds filled by numerical_bounds, non-local momentum flux profile
        st["difmxp"] = self.zeros(1)
        st["difmxq"] = self.zeros(1)
        prof = np.clip(1.0 - np.arange(kk + 1) / max(3.0, 0.15 * kk), 0.0, 1.0)[:, None, None]
        st["mu_nonloc"] = self._put(self.zeros(kk + 1), prof * ium)
        st["mv_nonloc"] = self._put(self.zeros(kk + 1), prof * ivm)
        st["ustarb"] = self.zeros(1)
        st["absvor"] = self.zeros(2 * kk)
        st["dpvor"] = self.zeros(2 * kk)
        # hybrid coordinate: dpuold/dpvold come from the (out-of-scope) ALE step; any positive
        # thickness-like field exercises the velocity time filter
        self.interior(st["dpuold"])[:] = 0.5 * (dpn + np.roll(dpn, 1, axis=2)) * ium
        # (no j-neighbour here: a band must hold exactly the rows of the one-tile state)
        self.interior(st["dpvold"])[:] = 0.5 * (dpn + np.roll(dpn, -1, axis=2)) * ivm
        st["utotn"] = self._put(self.zeros(1), 1.0e-6 * self._normal(1) * ium)
        st["vtotn"] = self._put(self.zeros(1), 1.0e-6 * self._normal(1) * ivm)
        # --- isopycnic bulk-mixed-layer coordinate: index of the first physical layer below the mixed
        # layer (phy/mod_state.F90 kfpla, 2 time levels), 3..kk+1 (kk+1: mixed layer reaches the bottom).
        # Integer fields cannot go through the r8 xctilr, so the halo is generated directly: the hash
        # is keyed on wrapped global indices (rings outside a closed edge are land and never read).
        u = self._hash_uniform_padded(2, 9001)
        st["kfpla"] = np.minimum(kk + 1, 3 + np.floor(np.minimum(1.0, 1.2 * u) * (kk - 2))).astype(np.int32)
        return st

    def scalars(self, nstep=1):
        """Step-control scalars (phy/mod_time.F90:121-142, mod_blom_step.F90:300)."""
        lstep = 2 * int(np.ceil(0.5 * self.baclin / self.batrop))
        dlt = self.baclin / lstep
        return {"baclin": self.baclin, "batrop": self.batrop, "lstep": lstep, "dlt": dlt,
                "delt1": 2.0 * self.baclin, "nstep": nstep, "pref": 2000.0 * ONEM / 9.806 * 9.806,
                "cwbdts": 5.0e-5, "cwbdls": 25.0,
                # momentum dissipation / bottom friction (namelist LIMITS, tests/fuk95/limits:144-155;
                # biharmonic terms switched on so that every branch of momtum is exercised)
                "mdv2hi": 0.1, "mdv2lo": 0.05, "mdv4hi": 0.01, "mdv4lo": 0.005, "vsc2hi": 0.2, "vsc2lo": 0.15,
                "vsc4hi": 0.06, "vsc4lo": 0.05, "cbar": 0.05, "cb": 0.002}


def make_synth(config, **kw):
    """Generator of a named configuration: the seeded noise state, or the analytic fuk95 case."""
    itdm, jtdm, kdm, nreg, baclin, batrop = CONFIGS[config]
    if config == "fuk95_analytic":
        from .fuk95 import Fuk95
        kw = {k: v for k, v in kw.items() if k in ("ntr", "j0", "jj", "u0")}
        return Fuk95(itdm, jtdm, kdm, baclin=baclin, batrop=batrop, **kw)
    return Synth(itdm, jtdm, kdm, nreg, baclin=baclin, batrop=batrop, **kw)


def make_isopycnic(st):
    """Make a synthetic state consistent with vcoord='isopyc_bulkml' (what the reference's bulk mixed
    layer scheme guarantees before eddtra runs): layers 1-2 are the mixed layer, interior layers
    3..kfpla-1 are massless, layer kfpla holds mass, and kfpla = kk+1 where no interior layer does.
    The mass taken out of the emptied layers goes to layer 2, so the bottom pressure is unchanged.
    Works on whole arrays (halo included); both time levels get the same treatment."""
    kk = st["dp"].shape[0] // 2
    kf = st["kfpla"]
    dp = st["dp"]
    kidx = np.arange(1, kk + 1)[:, None, None]
    for lvl in range(2):
        d = dp[lvl * kk:(lvl + 1) * kk]
        k0 = kf[lvl][None].astype(np.int64)
        has = np.take_along_axis(d, np.minimum(k0, kk) - 1, axis=0)[0] > 1.0e-6
        kf[lvl][...] = np.where(has & (kf[lvl] <= kk), kf[lvl], kk + 1)
        empty = (kidx >= 3) & (kidx < kf[lvl][None])
        moved = np.where(empty, d, 0.0).sum(axis=0)
        d[empty] = 0.0
        d[1] += moved


def ndiff_inputs(syn, st, levels, ntr=0):
    """Synthetic products of the ALE regrid-remap pipeline that neutral diffusion consumes
    (phy/mod_ale_regrid_remap.F90:1614-1690 -> phy/mod_ndiff.F90): source interfaces and the deepest
    source layer with mass, a monotone-limited parabolic (+ zero-mean quartic) reconstruction of every
    scalar in every source layer with its interface values, regridded destination interfaces (some
    within the 1 % snapping distance of their source interface, some not), the remapped tracers and
    the mixed-layer pressure thickness.  Built level-wise from halo-valid dp/T/S/trc on the whole padded
    arrays, so halos stay consistent without integer halo exchanges.  Layouts: see oracle/ndiff.cpp."""
    m, n, mm, nn, k1m, k1n = levels
    kk = st["dp"].shape[0] // 2
    T = 2 + ntr
    dp = st["dp"][nn:nn + kk]
    out = {}
    p_src = np.zeros((kk + 1,) + dp.shape[1:])
    p_src[1:] = np.cumsum(dp, axis=0)
    out["nd_p_src"] = p_src
    has = dp > 1.0e-6
    kidx = np.arange(1, kk + 1)[:, None, None]
    out["nd_ksmx"] = np.maximum(1, np.where(has, kidx, 0).max(axis=0)).astype(np.int32)[None]
    u = syn._hash_uniform_padded(kk + 1, 9101)
    p_dst = p_src.copy()
    dmin = np.minimum(dp[:-1], dp[1:])
    # a third of the interfaces stay within the snapping distance (1 % of the thinner neighbour)
    shift = np.where(u[1:kk] < 0.33, 0.004, 0.6) * dmin * (u[1:kk] - 0.5)
    p_dst[1:kk] = p_src[1:kk] + shift
    out["nd_p_dst"] = p_dst
    scal = [st["temp"][nn:nn + kk], st["saln"][nn:nn + kk]]
    for nt in range(ntr):
        scal.append(st["trc"][nt * 2 * kk + nn:nt * 2 * kk + nn + kk])
    tsd = np.zeros((2 * kk * T,) + dp.shape[1:])
    tpc = np.zeros((5 * kk * T,) + dp.shape[1:])
    trm = np.zeros((kk * T,) + dp.shape[1:])
    w = syn._hash_uniform_padded(kk, 9102) - 0.5
    for nt, f in enumerate(scal):
        edge = np.zeros((kk + 1,) + dp.shape[1:])
        edge[1:kk] = 0.5 * (f[:-1] + f[1:])
        edge[0], edge[kk] = f[0], f[kk - 1]
        tl, tr = edge[:-1].copy(), edge[1:].copy()
        flat = (tr - f) * (f - tl) <= 0.0
        tl[flat] = f[flat]; tr[flat] = f[flat]
        c = 0.05 * w * np.abs(tr - tl)                  # zero-mean quartic c*(x^2(1-x)^2 - 1/30)
        a0 = tl - c / 30.0
        a1 = 6.0 * f - 4.0 * tl - 2.0 * tr
        a2 = 3.0 * (tl - 2.0 * f + tr) + c
        a3 = -2.0 * c
        a4 = c
        for k in range(kk):
            b = (nt * kk + k) * 5
            tpc[b + 0], tpc[b + 1], tpc[b + 2], tpc[b + 3], tpc[b + 4] = a0[k], a1[k], a2[k], a3[k], a4[k]
            tsd[(nt * kk + k) * 2 + 0] = a0[k]
            tsd[(nt * kk + k) * 2 + 1] = (((a4[k] + a3[k]) + a2[k]) + a1[k]) + a0[k]
        trm[nt * kk:(nt + 1) * kk] = f
    out["nd_t_srcdi"], out["nd_tpc_src"], out["nd_trc_rm"] = tsd, tpc, trm
    out["dpml"] = ((10.0 + 90.0 * syn._hash_uniform_padded(1, 9103)) * ONEM)
    return {k: np.ascontiguousarray(v) for k, v in out.items()}


def cmnfld_arrays(syn):
    """Output arrays of cmnfld2's hybrid branch (phy/mod_cmnfld.F90:51-75): bfsqi, bfsqf on kdm+1
    interfaces, bfsql on kdm layers, nnslpx/nnslpy on kdm interfaces.  nslpx/nslpy are part of state()."""
    kk = syn.kdm
    return {"bfsqi": syn.zeros(kk + 1), "bfsqf": syn.zeros(kk + 1), "bfsql": syn.zeros(kk),
            "nnslpx": syn.zeros(kk), "nnslpy": syn.zeros(kk)}


def fill_halos(backend, arrays: dict, nbdy=4, names=None):
    """xctilr(nbdy,nbdy) of every registered array with its grid type.  `backend` is a
    BlomGpu (product) or the test Oracle; arrays must already be registered."""
    for name, a in arrays.items():
        if names is not None and name not in names:
            continue
        it = ITYPE.get(name)
        if it is None:
            continue
        lev = a.shape[-1] * a.shape[-2]
        nlev = a.size // lev
        backend.xctilr(name, 1, nlev, nbdy, nbdy, it)


def derive(gr, st, masks, levels, scalars, backend, sync_in=None, sync_out=None):
    """Fields other routines of the model would have produced before the hot path runs:
    p from dp(kn), pbu/pbv, dpu/dpv (phy/mod_pgforc.F90:452-484), umax/vmax
    (phy/mod_blom_init.F90:514-523).  numpy on whole arrays, then halos refreshed."""
    m, n, mm, nn, k1m, k1n = levels
    kk = st["dp"].shape[0] // 2
    iu, iv = masks["iu"], masks["iv"]
    dp, p = st["dp"], st["p"]
    p[0] = 0.0
    for k in range(kk):
        p[k + 1] = p[k] + dp[k + nn]
    pb = st["pb"]
    pb[n - 1] = p[kk]
    pb[m - 1] = dp[mm:mm + kk].sum(axis=0)
    st["pb_p"][0] = pb[n - 1]
    for lvl in range(2):
        st["pbu"][lvl][:, 1:] = np.minimum(pb[lvl][:, 1:], pb[lvl][:, :-1])
        st["pbv"][lvl][1:, :] = np.minimum(pb[lvl][1:, :], pb[lvl][:-1, :])
    st["pbu_p"][0] = st["pbu"][n - 1]
    st["pbv_p"][0] = st["pbv"][n - 1]
    for (koff, pl) in ((nn, p),):
        for k in range(kk):
            q1 = np.minimum(st["pbu"][n - 1][:, 1:], 0.5 * (pl[k + 1][:, 1:] + pl[k + 1][:, :-1]))
            q0 = np.minimum(st["pbu"][n - 1][:, 1:], 0.5 * (pl[k][:, 1:] + pl[k][:, :-1]))
            st["dpu"][k + koff][:, 1:] = np.maximum(0.0, q1 - q0) * iu[:, 1:]
            q1 = np.minimum(st["pbv"][n - 1][1:, :], 0.5 * (pl[k + 1][1:, :] + pl[k + 1][:-1, :]))
            q0 = np.minimum(st["pbv"][n - 1][1:, :], 0.5 * (pl[k][1:, :] + pl[k][:-1, :]))
            st["dpv"][k + koff][1:, :] = np.maximum(0.0, q1 - q0) * iv[1:, :]
    # level m thicknesses at u/v points from dp(km)
    pm = np.zeros_like(p)
    for k in range(kk):
        pm[k + 1] = pm[k] + dp[k + mm]
    for k in range(kk):
        q1 = np.minimum(st["pbu"][m - 1][:, 1:], 0.5 * (pm[k + 1][:, 1:] + pm[k + 1][:, :-1]))
        q0 = np.minimum(st["pbu"][m - 1][:, 1:], 0.5 * (pm[k][:, 1:] + pm[k][:, :-1]))
        st["dpu"][k + mm][:, 1:] = np.maximum(0.0, q1 - q0) * iu[:, 1:]
        q1 = np.minimum(st["pbv"][m - 1][1:, :], 0.5 * (pm[k + 1][1:, :] + pm[k + 1][:-1, :]))
        q0 = np.minimum(st["pbv"][m - 1][1:, :], 0.5 * (pm[k][1:, :] + pm[k][:-1, :]))
        st["dpv"][k + mm][1:, :] = np.maximum(0.0, q1 - q0) * iv[1:, :]
    baclin = scalars["baclin"]
    scp2, scuy, scvx = gr["scp2"][0], gr["scuy"][0], gr["scvx"][0]
    with np.errstate(invalid="ignore", divide="ignore"):
        st["umax"][0][:, 1:] = 0.9 * 0.125 * np.minimum(scp2[:, 1:], scp2[:, :-1]) / (scuy[:, 1:] * baclin)
        st["vmax"][0][1:, :] = 0.9 * 0.125 * np.minimum(scp2[1:, :], scp2[:-1, :]) / (scvx[1:, :] * baclin)
    for nm in ("umax", "vmax"):
        st[nm][~np.isfinite(st[nm])] = 0.0
    names = {"p", "pb", "pb_p", "pbu", "pbv", "pbu_p", "pbv_p", "dpu", "dpv", "umax", "vmax"}
    if sync_in is not None:
        sync_in(names)      # e.g. upload the freshly derived arrays to the device
    fill_halos(backend, {**gr, **st}, names=names)
    if sync_out is not None:
        sync_out(names)     # e.g. download the halo-filled arrays
    # land points carry zero bottom pressure; keep divisions finite like the model does
    for nm in ("pbu", "pbv"):
        a = st[nm]
        a[a <= 0.0] = 0.0


