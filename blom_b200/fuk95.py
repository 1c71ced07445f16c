"""fuk95: the analytic geometry and initial conditions of the reference's idealized stand-alone test
(Fukamachi et al. 1995 channel, `fuk95/mod_fuk95.F90`), as a synthetic case for the hot path.

What is restated here (setup code, runs once on the host; the hot path itself is CUDA only):
  * geoenv_fuk95 (fuk95/mod_fuk95.F90:117-231): flat channel of depth h0 with closed walls at
    i=1 and i=itdm, periodic in j (nreg=4), uniform grid spacing gs = lambda/jtdm = 650 m,
    constant Coriolis parameter f;
  * inicon_fuk95, hybrid-coordinate branch (:336-445): equidistant z-level interfaces, potential
    density of a baroclinic jet of half-width l0 perturbed by a wavenumber-1 meander (x_nudge, psi,
    x_psi :72-110), constant salinity, temperature from tofsig (phy/mod_eos.F90:346-364), zero
    velocities (the reference overrides the jet velocity with v1=0, :417);
  * inifrc_fuk95 (:233-261): no surface forcing.
What is NOT restated: the conversion of the z interfaces to layer pressure thickness that the
reference does in `phy/mod_inicon.F90` (EOS iteration, ALE regridding, CVMix set-up: out of scope,
SURVEY.md §8f rank 4).  Here dp = g*rho0*dz (`onem` per metre), which is the same state to O(1e-3).
The column-physics products the path reads (diffusivities, neutral slopes, mixed-layer depths) are
set to quiet constants; the time step and dissipation parameters follow tests/fuk95/limits.
"""
from __future__ import annotations

import numpy as np

from .synth import ONEM, Synth

# fuk95/mod_fuk95.F90:47-60
U0, H1, H0, L0 = 0.3, 1.0e2, 2.0e2, 2.0e4
DRHO, RHOC, RHOB = 0.19, 1025.9, 1027.0
F0, LAT0, LAMBDA, SALN0 = 1.0e-4, 45.0, 20.8e3, 35.0
GRAV, RHO0, ALPHA0 = 9.806, 1.0e3, 1.0e-3

# phy/mod_eos.F90:36-54 (rational EOS) and :105-116 (inieos: coefficients at pref)
_A1 = (9.9985372432159340e+02, 1.0380621928183473e+01, 1.7073577195684715e+00,
       -3.6570490496333680e-02, -7.3677944503527477e-03, -3.5529175999643348e-03)
_B1 = (1.7083494994335439e-06, 7.1567921402953455e-09, 1.2821026080049485e-09)
_A2 = (1.0, 1.0316374535350838e-02, 8.9521792365142522e-04,
       -2.8438341552142710e-05, -1.1887778959461776e-05, -4.0163964812921489e-06)
_B2 = (1.1995545126831476e-09, 5.5234008384648383e-12, 8.4310335919950873e-13)


def eos_coefficients(pref):
    """ap11..ap16, ap21..ap26 of inieos (phy/mod_eos.F90:105-116)."""
    ap2 = [_A2[0] + _B2[0] * pref, _A2[1] + _B2[1] * pref, _A2[2] + _B2[2] * pref, _A2[3], _A2[4], _A2[5]]
    ap1 = [_A1[0] + _B1[0] * pref - ap2[0] / ALPHA0, _A1[1] + _B1[1] * pref - ap2[1] / ALPHA0,
           _A1[2] + _B1[2] * pref - ap2[2] / ALPHA0, _A1[3] - ap2[3] / ALPHA0, _A1[4] - ap2[4] / ALPHA0,
           _A1[5] - ap2[5] / ALPHA0]
    return ap1, ap2


def sig(th, s, pref):
    """phy/mod_eos.F90:191-203."""
    a1, a2 = eos_coefficients(pref)
    return (a1[0] + (a1[1] + a1[3] * th + a1[4] * s) * th + (a1[2] + a1[5] * s) * s) / \
           (a2[0] + (a2[1] + a2[3] * th + a2[4] * s) * th + (a2[2] + a2[5] * s) * s)


def tofsig(sg, s, pref):
    """Potential temperature from potential density and salinity (phy/mod_eos.F90:346-364)."""
    a1, a2 = eos_coefficients(pref)
    a = a1[3] - a2[3] * sg
    b = a1[1] - a2[1] * sg + (a1[4] - a2[4] * sg) * s
    c = a1[0] - a2[0] * sg + (a1[2] - a2[2] * sg + (a1[5] - a2[5] * sg) * s) * s
    return (-b - np.sqrt(b * b - 4.0 * a * c)) / (2.0 * a)


def psi(x):
    """Shape function of the jet (fuk95/mod_fuk95.F90:84-96)."""
    return np.where(np.abs(x) >= L0, 0.0, 0.5 * (1.0 + np.cos(np.pi * x / L0)))


def x_psi(x):
    """Integral of the shape function (:98-110)."""
    return np.where(x <= -L0, -0.5 * L0, np.where(x >= L0, 0.5 * L0, 0.5 * (x + L0 / np.pi * np.sin(np.pi * x / L0))))


class Fuk95(Synth):
    """The fuk95 case on the reference's grid (156 x 32 x 12 by default; any even itdm works)."""

    def __init__(self, itdm=156, jtdm=32, kdm=12, *, ntr=0, j0=0, jj=None, baclin=180.0, batrop=6.0, u0=U0):
        self.u0 = u0   # jet speed the front is built for; 0 gives level isopycnals (a state of rest)
        super().__init__(itdm, jtdm, kdm, 4, ntr=ntr, j0=j0, jj=jj, baclin=baclin, batrop=batrop, land=False,
                         metric="uniform")

    # geoenv_fuk95 (:117-231)
    def _build_geometry(self):
        itdm, jtdm = self.itdm, self.jtdm
        depth = np.full((jtdm, itdm), H0)
        depth[:, 0] = 0.0
        depth[:, -1] = 0.0
        self.depth_global = depth
        self.gs = LAMBDA / jtdm
        self._scpx_g = np.full((jtdm, itdm), self.gs)
        self._scpy_g = np.full((jtdm, itdm), self.gs)
        self._lat_g = np.full((jtdm, itdm), LAT0)

    def grid(self):
        out = super().grid()
        for nm in ("corioq", "coriop"):
            out[nm] = self._put(self.zeros(), np.full((self.jj, self.itdm), F0))
        return out

    def x_nudge(self, ri, rj):
        """Perturbed cross-channel position (:72-82); ri, rj are global 1-based grid indices."""
        return (ri - self.itdm // 2 - 0.5 + 0.1 * np.sin(2.0 * (rj - 1.0) * np.pi / self.jtdm)) * LAMBDA / self.jtdm

    def interfaces(self):
        """z(k), k=1..kk+1 of the hybrid branch (:365-377): equidistant levels."""
        kk = self.kdm
        return np.arange(kk + 1) * H0 / kk

    def initial_sigma(self):
        """Layer potential density [sigma units] of the hybrid branch (:379-396), band interior."""
        kk = self.kdm
        z = self.interfaces()
        ri = np.arange(1, self.itdm + 1, dtype=np.float64)[None, :]
        rj = (self.j0 + np.arange(1, self.jj + 1, dtype=np.float64))[:, None]
        x = self.x_nudge(ri, rj)
        s0 = RHOB - RHO0
        sg = np.zeros((kk, self.jj, self.itdm))
        for k in range(kk):
            s1 = RHOC * (1.0 + F0 * self.u0 * x_psi(x) / (GRAV * H1)) - RHO0 + 0.5 * DRHO * (z[k + 1] + z[k] - H1) / H1
            sg[k] = (s1 * max(0.0, min(z[k + 1], H1) - z[k]) + s0 * max(0.0, z[k + 1] - max(z[k], H1))) / (z[k + 1] - z[k])
        return sg

    def state(self, grid):
        st = super().state(grid)
        kk = self.kdm
        ipm, ium, ivm = self.masks_np()
        pref = self.scalars()["pref"]
        z = self.interfaces()
        sg = self.initial_sigma()
        saln = np.full_like(sg, SALN0)
        temp = tofsig(sg, saln, pref)
        dpk = (np.diff(z) * ONEM)[:, None, None] * ipm[None]

        def both_levels(a, val):
            self.interior(a)[:kk] = val
            self.interior(a)[kk:2 * kk] = val

        for nm, val in (("dp", dpk), ("temp", temp * ipm), ("saln", saln * ipm), ("sigma", sg * ipm)):
            st[nm][...] = 0.0
            both_levels(st[nm], val)
        self.interior(st["dpold"])[:] = np.concatenate([dpk, dpk])
        self.interior(st["told"])[:] = temp * ipm
        self.interior(st["sold"])[:] = saln * ipm
        pb = dpk.sum(axis=0)
        # at rest: velocities, mass fluxes, barotropic fields and their time-level copies
        for nm in ("u", "v", "umfltd", "vmfltd", "umflsm", "vmflsm", "ubflxs_p", "vbflxs_p", "ubflx", "vbflx",
                   "ubflx_mn", "vbflx_mn", "utotn", "vtotn", "taux", "tauy", "pgfx", "pgfy", "pgfxm", "pgfym",
                   "pgfxm_o", "pgfym_o", "xixp", "xixm", "xiyp", "xiym", "xixp_o", "xixm_o", "xiyp_o", "xiym_o",
                   "nslpx", "nslpy", "ustar3", "wstar3", "mu_nonloc", "mv_nonloc"):
            st[nm][...] = 0.0
        for nm in ("pb", "pb_mn"):
            st[nm][...] = 0.0
            self.interior(st[nm])[:] = pb
        self.interior(st["pb_p"])[0] = pb
        self.interior(st["phi"])[:] = (-GRAV * z)[:, None, None] * ipm[None]   # :430-441
        pvq = np.where(pb > 0, F0 / np.maximum(pb, 1.0), 0.0)
        for nm in ("pvtrop", "pvtrop_o"):
            st[nm][...] = 0.0
            self.interior(st[nm])[:] = pvq
        self.interior(st["dpuold"])[:] = dpk * ium[None]
        self.interior(st["dpvold"])[:] = dpk * ivm[None]
        # quiet column-physics products: constant diffusivities, shallow boundary layer
        for nm, val in (("difint", 50.0), ("difiso", 50.0)):
            st[nm][...] = 0.0
            self.interior(st[nm])[:] = val * ipm[None]
        for nm, val in (("difwgt", 0.0), ("OBLdepth", 10.0), ("mld", 10.0), ("hbl_tf", 10.0), ("hml_tf1", 10.0),
                        ("hml_tf", 10.0), ("wpup_tf", 1.0e-3)):
            st[nm][...] = 0.0
            self.interior(st[nm])[0] = val * ipm
        if self.ntr > 0:   # passive tracer: a dye in the light side of the front
            ri = np.arange(1, self.itdm + 1, dtype=np.float64)[None, :]
            rj = (self.j0 + np.arange(1, self.jj + 1, dtype=np.float64))[:, None]
            dye = psi(self.x_nudge(ri, rj) + 0.5 * L0) * ipm
            st["trc"][...] = 0.0
            for nt in range(self.ntr):
                for lvl in range(2):
                    o = nt * 2 * kk + lvl * kk
                    self.interior(st["trc"])[o:o + kk] = dye[None]
                self.interior(st["trcold"])[nt * kk:(nt + 1) * kk] = dye[None]
        return st

    def scalars(self, nstep=1):
        """&LIMITS of tests/fuk95/limits:111-168: pref=0, baclin=180 s, batrop=6 s, no background
        viscosity (mdv*=0), vsc2hi=vsc2lo=.2, vsc4*=0, cbar=.05, cb=.002, cwbdts=0, ce=0."""
        sc = super().scalars(nstep)
        sc.update({"pref": 0.0, "mdv2hi": 0.0, "mdv2lo": 0.0, "mdv4hi": 0.0, "mdv4lo": 0.0, "vsc2hi": 0.2,
                   "vsc2lo": 0.2, "vsc4hi": 0.0, "vsc4lo": 0.0, "cbar": 0.05, "cb": 0.002, "cwbdts": 0.0,
                   "cwbdls": 25.0, "ce": 0.0})
        return sc
