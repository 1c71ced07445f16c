// Baroclinic momentum tendencies (phy/mod_momtum.F90:215-1282).
//
// The reference sweeps ~25 masked 2-D loops per layer over shared work arrays.
// Here every layer is independent until the final column pass, so each stage is
// ONE launch over all (i,j,k):
//   mt_aux   : sidewall-weighted neighbour velocities uja/ujb/via/vib, del2 fields,
//              tension (defor1)                                  (:438-472, :549-559)
//   mt_vort  : q-point gather of vorticity / potential vorticity / dpvor / shear
//              (defor2) — the reference's span-endpoint scatter loops become a
//              priority gather (last writer of the sequential order wins)
//                                                                 (:360-396, :477-585)
//   mt_visc  : deformation-dependent viscosities at u and v points (:829-841, :988-1000)
//   mt_update: Coriolis/advection, stress fluxes, bottom drag, wind stress, time-averaged
//              pressure gradient, leap-frog update and time filter part 1 (:723-980, :1017-1143)
//   mt_column: massless-layer fill, clamp, depth mean, filter part 2 (:1154-1267)
// Cheap point-wise fields of the reference (utotm/utotn/uflux/vflux/dpmx/wgt*/ke/uflux1..3)
// are recomputed in registers instead of being staged through memory; 13 layer-sized
// scratch arrays remain (uja,ujb,via,vib,dl2u,dl2v,defor1,defor2,potvor,vsc2/4 at u and v).
#include "common.cuh"

namespace blom {

namespace {

constexpr double SLIP = -1., THKBOT = 10., WUV1 = .75, WUV2 = .125, WPGF = .25;

struct MtP {
  // state
  double *u, *v, *p, *pu, *pv, *absvor, *dpvor, *utotn, *vtotn, *ustarb;
  const double *dp, *dpu, *dpv, *pbu, *pbv, *ubflxs_p, *vbflxs_p, *ub, *vb, *pgfx, *pgfy, *pgfx_o, *pgfy_o,
      *dpuold, *dpvold, *mu_nonloc, *mv_nonloc, *ubcors_p, *vbcors_p, *pbu_p, *pbv_p, *difwgt, *difmxp, *difmxq,
      *taux, *tauy, *umax, *vmax;
  // grid
  const double *scuy, *scvx, *scux, *scvy, *scq2i, *scp2i, *scp2, *scu2, *scv2, *scpx, *scpy, *scqx, *scqy, *scuxi,
      *scvyi, *corioq;
  const int *ip, *iu, *iv, *iq;
  // scratch
  double *uja, *ujb, *via, *vib, *dl2u, *dl2v, *defor1, *defor2, *potvor, *vsc2u, *vsc4u, *vsc2v, *vsc4v, *drag;
  // updated velocities are staged so that both tendency kernels see the pre-update u,v
  double *su_m, *su_n, *sv_m, *sv_n;
  // level-independent barotropic part of the total velocities, ubflxs_p*tsfac/(pbu*scuy) at time
  // levels n and m, evaluated once per call (mt_pressures) instead of once per use and level
  double *ubn, *ubm, *vbn, *vbm;
  // kinetic energy and longitudinal stress fluxes at mass points, staged once per level because each
  // is needed by two update threads (and costs ~20 loads to rebuild)
  double *ke, *uflux1, *vflux1;
  double delt1, tsfac, mdv2hi, mdv2lo, mdv4hi, mdv4lo, vsc2hi, vsc2lo, vsc4hi, vsc4lo, cbar, cb;
  int m, n, mm, nn, mommth /*0 enscon 1 enecon 2 enedis*/, isopyc;
};

__device__ __forceinline__ double hfharm(double a, double b) { return a * b / (a + b); }
__device__ __forceinline__ double sq(double a) { return a * a; }

// ---- point-wise fields of the reference, recomputed on demand -------------------------------
// total velocities: in-range masked points carry the level-k value, every other point the
// (stale) content of the 2-D module array, exactly like the reference's shared work arrays.
__device__ __forceinline__ double utotn_at(const Geom& g, const MtP& P, int i, int j, int k) {
  const long x = ix2(g, i, j);
  if (i >= -1 && i <= g.ii + 2 && j >= -1 && j <= g.jj + 2 && P.iu[x] == 1) {
    return P.u[x + (long)(k + P.nn - 1) * g.lev] + P.ubn[x];
  }
  return P.utotn[x];
}
__device__ __forceinline__ double vtotn_at(const Geom& g, const MtP& P, int i, int j, int k) {
  const long x = ix2(g, i, j);
  if (i >= -1 && i <= g.ii + 2 && j >= -1 && j <= g.jj + 2 && P.iv[x] == 1) {
    return P.v[x + (long)(k + P.nn - 1) * g.lev] + P.vbn[x];
  }
  return P.vtotn[x];
}
__device__ __forceinline__ double utotm_at(const Geom& g, const MtP& P, int i, int j, int k) {
  const long x = ix2(g, i, j);
  if (i >= 0 && i <= g.ii + 1 && j >= 0 && j <= g.jj + 1 && P.iu[x] == 1) {
    return P.u[x + (long)(k + P.mm - 1) * g.lev] + P.ubm[x];
  }
  return 0.;
}
__device__ __forceinline__ double vtotm_at(const Geom& g, const MtP& P, int i, int j, int k) {
  const long x = ix2(g, i, j);
  if (i >= 0 && i <= g.ii + 1 && j >= 0 && j <= g.jj + 1 && P.iv[x] == 1) {
    return P.v[x + (long)(k + P.mm - 1) * g.lev] + P.vbm[x];
  }
  return 0.;
}
__device__ __forceinline__ double uflux_at(const Geom& g, const MtP& P, int i, int j, int k) {
  const long x = ix2(g, i, j);
  if (i >= 0 && i <= g.ii + 1 && j >= 0 && j <= g.jj + 1 && P.iu[x] == 1)
    return utotm_at(g, P, i, j, k) * fmax(P.dpu[x + (long)(k + P.mm - 1) * g.lev], onem);
  return 0.;
}
__device__ __forceinline__ double vflux_at(const Geom& g, const MtP& P, int i, int j, int k) {
  const long x = ix2(g, i, j);
  if (i >= 0 && i <= g.ii + 1 && j >= 0 && j <= g.jj + 1 && P.iv[x] == 1)
    return vtotm_at(g, P, i, j, k) * fmax(P.dpv[x + (long)(k + P.mm - 1) * g.lev], onem);
  return 0.;
}
// dpmx at q-point (i,j), 0<=i<=ii+2, 0<=j<=jj+2 (:360-396)
__device__ __forceinline__ double dpmx_at(const Geom& g, const MtP& P, int i, int j, int k) {
  const long x = ix2(g, i, j), s = g.ldi;
  const double* dpm = P.dp + (long)(k + P.mm - 1) * g.lev;
  double r = 8. * onem;
  if (P.iu[x] == 1) r = fmax(r, dpm[x] + dpm[x - 1]);
  if (P.iu[x - s] == 1) r = fmax(r, dpm[x - s] + dpm[x - s - 1]);
  if (P.iv[x] == 1) r = fmax(r, dpm[x] + dpm[x - s]);
  if (P.iv[x - 1] == 1) r = fmax(r, dpm[x - 1] + dpm[x - 1 - s]);
  return r;
}
__device__ __forceinline__ void wgtj_at(const Geom& g, const MtP& P, long x, int k, double& wa, double& wb) {
  const long s = g.ldi, m2 = (long)(P.m - 1) * g.lev;
  const double p1 = P.pu[x + (long)k * g.lev], p0 = P.pu[x + (long)(k - 1) * g.lev];
  const double den = fmax(p1 - p0, epsilp);
  wa = fmax(0., fmin(1., (p1 - P.pbu[x - s + m2]) / den));
  wb = fmax(0., fmin(1., (p1 - P.pbu[x + s + m2]) / den));
}
__device__ __forceinline__ void wgti_at(const Geom& g, const MtP& P, long x, int k, double& wa, double& wb) {
  const long m2 = (long)(P.m - 1) * g.lev;
  const double p1 = P.pv[x + (long)k * g.lev], p0 = P.pv[x + (long)(k - 1) * g.lev];
  const double den = fmax(p1 - p0, epsilp);
  wa = fmax(0., fmin(1., (p1 - P.pbv[x - 1 + m2]) / den));
  wb = fmax(0., fmin(1., (p1 - P.pbv[x + 1 + m2]) / den));
}

// ---- column pre-passes ---------------------------------------------------------------------
// p(k+1)=p(k)+dp(km) on -1..ii+2 (:244-255); pu,pv from dpu,dpv(km) (:322-338)
__global__ void mt_pressures(Geom g, MtP P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x - 1, j = (int)blockIdx.y - 1;
  if (i > g.ii + 2) return;
  const long x = ix2(g, i, j);
  if (P.ip[x] == 1) {
    double pk = P.p[x];
    for (int k = 1; k <= g.kdm; ++k) { pk = pk + P.dp[x + (long)(k + P.mm - 1) * g.lev]; P.p[x + (long)k * g.lev] = pk; }
  }
  const long xn2 = x + (long)(P.n - 1) * g.lev, xm2 = x + (long)(P.m - 1) * g.lev;
  if (P.iu[x] == 1) {
    double pk = P.pu[x];
    for (int k = 1; k <= g.kdm; ++k) { pk = pk + P.dpu[x + (long)(k + P.mm - 1) * g.lev]; P.pu[x + (long)k * g.lev] = pk; }
    P.ubn[x] = P.ubflxs_p[xn2] * P.tsfac / (P.pbu[xn2] * P.scuy[x]);
    P.ubm[x] = P.ubflxs_p[xm2] * P.tsfac / (P.pbu[xm2] * P.scuy[x]);
  }
  if (P.iv[x] == 1) {
    double pk = P.pv[x];
    for (int k = 1; k <= g.kdm; ++k) { pk = pk + P.dpv[x + (long)(k + P.mm - 1) * g.lev]; P.pv[x + (long)k * g.lev] = pk; }
    P.vbn[x] = P.vbflxs_p[xn2] * P.tsfac / (P.pbv[xn2] * P.scvx[x]);
    P.vbm[x] = P.vbflxs_p[xm2] * P.tsfac / (P.pbv[xm2] * P.scvx[x]);
  }
}
// bottom drag (:259-293) on 0..ii x 0..jj
__global__ void mt_drag(Geom g, MtP P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i > g.ii) return;
  const long x = ix2(g, i, j), s = g.ldi;
  if (P.ip[x] != 1) return;
  const double thkbop = THKBOT * onem;
  const double pb = P.p[x + (long)g.kdm * g.lev];
  double u1 = 0., u2 = 0., pk = P.p[x];
  for (int k = 1; k <= g.kdm; ++k) {
    const long xn = x + (long)(k + P.nn - 1) * g.lev;
    const double pk1 = P.p[x + (long)k * g.lev];
    const double pbotl = fmax(pk1, pb - thkbop), ptopl = fmax(pk, pb - thkbop);
    u1 = u1 + (P.u[xn] + P.u[xn + 1]) * (pbotl - ptopl);
    u2 = u2 + (P.v[xn] + P.v[xn + s]) * (pbotl - ptopl);
    pk = pk1;
  }
  const long x2 = x + (long)(P.n - 1) * g.lev;
  const double ubot = (P.ubflxs_p[x2] / fmax(epsilpl, P.pbu[x2] * P.scuy[x]) +
                       P.ubflxs_p[x2 + 1] / fmax(epsilpl, P.pbu[x2 + 1] * P.scuy[x + 1])) * P.tsfac + u1 / thkbop;
  const double vbot = (P.vbflxs_p[x2] / fmax(epsilpl, P.pbv[x2] * P.scvx[x]) +
                       P.vbflxs_p[x2 + s] / fmax(epsilpl, P.pbv[x2 + s] * P.scvx[x + s])) * P.tsfac + u2 / thkbop;
  const double ubbl = .5 * sqrt(ubot * ubot + vbot * vbot);
  const double q = P.cb * (ubbl + P.cbar);
  P.drag[x] = q * grav / (alpha0 * thkbop);
  P.ustarb[x] = sqrt(q * ubbl);
}

__device__ __forceinline__ double ke_at(const Geom& g, const MtP& P, int i, int j, int k);

// ---- stage 1: auxiliary velocities, del2, tension --------------------------------------------
__global__ void mt_aux(Geom g, MtP P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x - 1;  // -1..ii+2
  const int j = (int)blockIdx.y - 1, k = blockIdx.z + 1;    // -1..jj+2
  if (i > g.ii + 2) return;
  const long x = ix2(g, i, j), xk = x + (long)(k - 1) * g.lev;
  if (i >= 0 && P.iu[x] == 1) {
    double wa, wb;
    wgtj_at(g, P, x, k, wa, wb);
    const double un = utotn_at(g, P, i, j, k);
    const double a = (1. - wa) * utotn_at(g, P, i, j - 1, k) + wa * SLIP * un;
    const double b = (1. - wb) * utotn_at(g, P, i, j + 1, k) + wb * SLIP * un;
    P.uja[xk] = a; P.ujb[xk] = b;
    P.dl2u[xk] = un - .25 * (utotn_at(g, P, i + 1, j, k) + utotn_at(g, P, i - 1, j, k) + a + b);
  }
  if (j >= 0 && P.iv[x] == 1) {
    double wa, wb;
    wgti_at(g, P, x, k, wa, wb);
    const double vn = vtotn_at(g, P, i, j, k);
    const double a = (1. - wa) * vtotn_at(g, P, i - 1, j, k) + wa * SLIP * vn;
    const double b = (1. - wb) * vtotn_at(g, P, i + 1, j, k) + wb * SLIP * vn;
    P.via[xk] = a; P.vib[xk] = b;
    P.dl2v[xk] = vn - .25 * (vtotn_at(g, P, i, j + 1, k) + vtotn_at(g, P, i, j - 1, k) + a + b);
  }
  if (i >= 0 && i <= g.ii && j >= 0 && j <= g.jj) P.ke[xk] = ke_at(g, P, i, j, k);
  if (i <= g.ii + 1 && j <= g.jj + 1 && P.ip[x] == 1)
    P.defor1[xk] = sq((utotn_at(g, P, i + 1, j, k) * P.scuy[x + 1] - utotn_at(g, P, i, j, k) * P.scuy[x]) -
                      (vtotn_at(g, P, i, j + 1, k) * P.scvx[x + g.ldi] - vtotn_at(g, P, i, j, k) * P.scvx[x])) *
                   P.scp2i[x];
}

// ---- stage 2: q-point gather ---------------------------------------------------------------------
__global__ void mt_vort(Geom g, MtP P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // 0..ii+2
  const int j = blockIdx.y, k = blockIdx.z + 1;         // 0..jj+2
  if (i > g.ii + 2) return;
  const long x = ix2(g, i, j), s = g.ldi, xk = x + (long)(k - 1) * g.lev;
  const double* dpm = P.dp + (long)(k + P.mm - 1) * g.lev;
  const double q2i = P.scq2i[x];
  const bool vfirst = P.iv[x] == 1 && P.iv[x - 1] == 0, vlast = P.iv[x - 1] == 1 && P.iv[x] == 0;
  const bool ufirst = P.iu[x] == 1 && P.iu[x - s] == 0, ulast = P.iu[x - s] == 1 && P.iu[x] == 0;
  const bool isq = P.iq[x] == 1;
  // shear deformation (:498-509, :532-543, :577-585); later rules override earlier ones
  {
    bool set = false; double d2 = 0.;
    if (vfirst) { d2 = sq(vtotn_at(g, P, i, j, k) * (1. - SLIP) * P.scvy[x]) * q2i; set = true; }
    if (vlast) { d2 = sq(vtotn_at(g, P, i - 1, j, k) * (1. - SLIP) * P.scvy[x - 1]) * q2i; set = true; }
    if (ufirst) { d2 = sq(utotn_at(g, P, i, j, k) * (1. - SLIP) * P.scux[x]) * q2i; set = true; }
    if (ulast) { d2 = sq(utotn_at(g, P, i, j - 1, k) * (1. - SLIP) * P.scux[x - s]) * q2i; set = true; }
    if (isq) {
      d2 = sq(P.vib[xk - 1] * P.scvy[x] - P.via[xk] * P.scvy[x - 1] + P.ujb[xk - s] * P.scux[x] -
              P.uja[xk] * P.scux[x - s]) * q2i;
      set = true;
    }
    if (set) P.defor2[xk] = d2;
  }
  // vorticity, dpvor, potential vorticity on 1..ii+1 x 1..jj+1 (:477-496, :511-530, :561-575)
  if (i >= 1 && i <= g.ii + 1 && j >= 1 && j <= g.jj + 1) {
    bool set = false; double vort = 0., dpv = 1.;
    if (vfirst) {
      vort = vtotm_at(g, P, i, j, k) * (1. - SLIP) * P.scvy[x] * q2i;
      dpv = .125 * fmax(fmax(4. * (dpm[x] + dpm[x - s]), dpmx_at(g, P, i, j, k)), dpmx_at(g, P, i + 1, j, k));
      set = true;
    }
    if (vlast) {
      vort = -vtotm_at(g, P, i - 1, j, k) * (1. - SLIP) * P.scvy[x - 1] * q2i;
      dpv = .125 * fmax(fmax(4. * (dpm[x - 1] + dpm[x - 1 - s]), dpmx_at(g, P, i - 1, j, k)), dpmx_at(g, P, i, j, k));
      set = true;
    }
    if (ufirst) {
      vort = -utotm_at(g, P, i, j, k) * (1. - SLIP) * P.scux[x] * q2i;
      dpv = .125 * fmax(fmax(4. * (dpm[x] + dpm[x - 1]), dpmx_at(g, P, i, j, k)), dpmx_at(g, P, i, j + 1, k));
      set = true;
    }
    if (ulast) {
      vort = utotm_at(g, P, i, j - 1, k) * (1. - SLIP) * P.scux[x - s] * q2i;
      dpv = .125 * fmax(fmax(4. * (dpm[x - s] + dpm[x - s - 1]), dpmx_at(g, P, i, j - 1, k)), dpmx_at(g, P, i, j, k));
      set = true;
    }
    if (isq) {
      vort = (vtotm_at(g, P, i, j, k) * P.scvy[x] - vtotm_at(g, P, i - 1, j, k) * P.scvy[x - 1] -
              utotm_at(g, P, i, j, k) * P.scux[x] + utotm_at(g, P, i, j - 1, k) * P.scux[x - s]) * q2i;
      double mx = 2. * (dpm[x] + dpm[x - 1] + dpm[x - s] + dpm[x - s - 1]);
      mx = fmax(mx, dpmx_at(g, P, i, j, k)); mx = fmax(mx, dpmx_at(g, P, i - 1, j, k));
      mx = fmax(mx, dpmx_at(g, P, i + 1, j, k)); mx = fmax(mx, dpmx_at(g, P, i, j - 1, k));
      mx = fmax(mx, dpmx_at(g, P, i, j + 1, k));
      dpv = .125 * mx;
      set = true;
    }
    if (set) {
      const double av = vort + P.corioq[x];
      P.absvor[xk] = av;
      P.dpvor[xk] = dpv;
      P.potvor[xk] = av / dpv;
    }
  }
}

// ---- stage 3: viscosities (:829-841, :988-1000) ---------------------------------------------------
__global__ void mt_visc(Geom g, MtP P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // 0..ii+1
  const int j = blockIdx.y, k = blockIdx.z + 1;         // 0..jj+1
  if (i > g.ii + 1) return;
  const long x = ix2(g, i, j), s = g.ldi, xk = x + (long)(k - 1) * g.lev;
  if (P.iu[x] == 1) {
    const double q = .5 * (P.difwgt[x - 1] + P.difwgt[x]);
    const double deform = sqrt(.5 * (P.defor1[xk] + P.defor1[xk - 1] + P.defor2[xk] + P.defor2[xk + s]));
    P.vsc2u[xk] = fmax(q * P.mdv2hi + (1. - q) * P.mdv2lo, (q * P.vsc2hi + (1. - q) * P.vsc2lo) * deform);
    P.vsc4u[xk] = fmax(q * P.mdv4hi + (1. - q) * P.mdv4lo, (q * P.vsc4hi + (1. - q) * P.vsc4lo) * deform);
  }
  if (P.iv[x] == 1) {
    const double q = .5 * (P.difwgt[x - s] + P.difwgt[x]);
    const double deform = sqrt(.5 * (P.defor1[xk] + P.defor1[xk - s] + P.defor2[xk] + P.defor2[xk + 1]));
    P.vsc2v[xk] = fmax(q * P.mdv2hi + (1. - q) * P.mdv2lo, (q * P.vsc2hi + (1. - q) * P.vsc2lo) * deform);
    P.vsc4v[xk] = fmax(q * P.mdv4hi + (1. - q) * P.mdv4lo, (q * P.vsc4hi + (1. - q) * P.vsc4lo) * deform);
  }
}

// viscosity at u-position (i,j) as the reference's work array holds it after the span-end
// extension loops (:843-856): wet points carry their own value, a dry point next to a span
// takes the eastern span's first value if there is one, else the western span's last value.
__device__ __forceinline__ double viscu_ext(const Geom& g, const MtP& P, const double* vs, int i, int j, long xk) {
  const long x = ix2(g, i, j);
  if (P.iu[x] == 1) return vs[xk];
  if (P.iu[x + 1] == 1 && i + 1 > 0) return i + 1 <= g.ii + 1 ? vs[xk + 1] : 0.;
  if (P.iu[x - 1] == 1 && i - 1 < g.ii + 1) return i - 1 >= 0 ? vs[xk - 1] : 0.;
  return 0.;
}
__device__ __forceinline__ double viscv_ext(const Geom& g, const MtP& P, const double* vs, int i, int j, long xk) {
  const long x = ix2(g, i, j), s = g.ldi;
  if (P.iv[x] == 1) return vs[xk];
  if (P.iv[x + s] == 1 && j + 1 > 0) return j + 1 <= g.jj + 1 ? vs[xk + s] : 0.;
  if (P.iv[x - s] == 1 && j - 1 < g.jj + 1) return j - 1 >= 0 ? vs[xk - s] : 0.;
  return 0.;
}
// longitudinal stress flux at mass point (i,j) (:858-873)
__device__ __forceinline__ double uflux1_at(const Geom& g, const MtP& P, int i, int j, int k) {
  const long x = ix2(g, i, j), xk = x + (long)(k - 1) * g.lev, xm = x + (long)(k + P.mm - 1) * g.lev;
  const double dpxy = fmax(P.dpu[xm], onemm), dpib = fmax(P.dpu[xm + 1], onemm);
  const double v2 = viscu_ext(g, P, P.vsc2u, i, j, xk) + viscu_ext(g, P, P.vsc2u, i + 1, j, xk + 1);
  const double v4 = viscu_ext(g, P, P.vsc4u, i, j, xk) + viscu_ext(g, P, P.vsc4u, i + 1, j, xk + 1);
  const double hh = hfharm(dpxy, dpib);
  return fmin(P.difmxp[x], v2 * P.scpy[x]) * hh * (utotn_at(g, P, i, j, k) - utotn_at(g, P, i + 1, j, k)) +
         fmin(.125 * P.difmxp[x], v4 * P.scpy[x]) * hh * (P.dl2u[xk] - P.dl2u[xk + 1]);
}
__device__ __forceinline__ double vflux1_at(const Geom& g, const MtP& P, int i, int j, int k) {
  const long x = ix2(g, i, j), s = g.ldi, xk = x + (long)(k - 1) * g.lev, xm = x + (long)(k + P.mm - 1) * g.lev;
  const double dpxy = fmax(P.dpv[xm], onemm), dpjb = fmax(P.dpv[xm + s], onemm);
  const double v2 = viscv_ext(g, P, P.vsc2v, i, j, xk) + viscv_ext(g, P, P.vsc2v, i, j + 1, xk + s);
  const double v4 = viscv_ext(g, P, P.vsc4v, i, j, xk) + viscv_ext(g, P, P.vsc4v, i, j + 1, xk + s);
  const double hh = hfharm(dpxy, dpjb);
  return fmin(P.difmxp[x], v2 * P.scpx[x]) * hh * (vtotn_at(g, P, i, j, k) - vtotn_at(g, P, i, j + 1, k)) +
         fmin(.125 * P.difmxp[x], v4 * P.scpx[x]) * hh * (P.dl2v[xk] - P.dl2v[xk + s]);
}
__device__ __forceinline__ double ke_at(const Geom& g, const MtP& P, int i, int j, int k) {
  const long x = ix2(g, i, j), s = g.ldi;
  return .25 * (P.scu2[x] * sq(utotm_at(g, P, i, j, k)) + P.scu2[x + 1] * sq(utotm_at(g, P, i + 1, j, k)) +
                P.scv2[x] * sq(vtotm_at(g, P, i, j, k)) + P.scv2[x + s] * sq(vtotm_at(g, P, i, j + 1, k))) / P.scp2[x];
}
// Sadourny energy conserving scheme with dissipation: min/max transports (:664-719)
__device__ __forceinline__ void uh_minmax(const Geom& g, const MtP& P, int i, int j, int k, double& mn, double& mx) {
  const long x = ix2(g, i, j);
  mn = 0.; mx = 0.;
  if (!(i >= 0 && i <= g.ii + 1 && j >= 0 && j <= g.jj + 1 && P.iu[x] == 1)) return;
  const double* dpm = P.dp + (long)(k + P.mm - 1) * g.lev;
  double uhc = .5 * utotm_at(g, P, i, j, k) * (dpm[x] + dpm[x - 1]), uhm = uflux_at(g, P, i, j, k);
  const double c1 = 1. - 1.5 * .5, c2 = 1. - .5, c3 = 2., slope = .5;
  if (fabs(uhc) < .1 * fabs(uhm)) uhm = 10. * uhc;
  else if (fabs(uhc) > c1 * fabs(uhm)) {
    if (fabs(uhc) < c2 * fabs(uhm)) uhc = (3. * uhc + (1. - c2 * 3.) * uhm);
    else if (fabs(uhc) <= c3 * fabs(uhm)) uhc = uhm;
    else uhc = slope * uhc + (1. - c3 * slope) * uhm;
  }
  if (uhc > uhm) { mn = uhm; mx = uhc; } else { mx = uhm; mn = uhc; }
}
__device__ __forceinline__ void vh_minmax(const Geom& g, const MtP& P, int i, int j, int k, double& mn, double& mx) {
  const long x = ix2(g, i, j);
  mn = 0.; mx = 0.;
  if (!(i >= 0 && i <= g.ii + 1 && j >= 0 && j <= g.jj + 1 && P.iv[x] == 1)) return;
  const double* dpm = P.dp + (long)(k + P.mm - 1) * g.lev;
  double vhc = .5 * vtotm_at(g, P, i, j, k) * (dpm[x] + dpm[x - g.ldi]), vhm = vflux_at(g, P, i, j, k);
  const double c1 = 1. - 1.5 * .5, c2 = 1. - .5, c3 = 2., slope = .5;
  if (fabs(vhc) < .1 * fabs(vhm)) vhm = 10. * vhc;
  else if (fabs(vhc) > c1 * fabs(vhm)) {
    if (fabs(vhc) < c2 * fabs(vhm)) vhc = (3. * vhc + (1. - c2 * 3.) * vhm);
    else if (fabs(vhc) <= c3 * fabs(vhm)) vhc = vhm;
    else vhc = slope * vhc + (1. - c3 * slope) * vhm;
  }
  if (vhc > vhm) { mn = vhm; mx = vhc; } else { mx = vhm; mn = vhc; }
}

// ---- stage 3b: longitudinal stress fluxes at mass points (:858-873, :1017-1032) ------------------------
__global__ void mt_flux1(Geom g, MtP P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // 0..ii
  const int j = blockIdx.y, k = blockIdx.z + 1;         // 0..jj
  if (i > g.ii) return;
  const long xk = ix2(g, i, j) + (long)(k - 1) * g.lev;
  if (j >= 1) P.uflux1[xk] = uflux1_at(g, P, i, j, k);
  if (i >= 1) P.vflux1[xk] = vflux1_at(g, P, i, j, k);
}

// ---- stage 4: tendencies and leap-frog update ------------------------------------------------------
__global__ void __launch_bounds__(128)
mt_update(Geom g, MtP P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1, k = blockIdx.z + 1;
  if (i > g.ii) return;
  const long x = ix2(g, i, j), s = g.ldi, L = g.lev;
  const long xk = x + (long)(k - 1) * L, xm = x + (long)(k + P.mm - 1) * L, xn = x + (long)(k + P.nn - 1) * L;
  const long m2 = (long)(P.m - 1) * L;
  const double thkbop = THKBOT * onem;
  if (P.iu[x] == 1) {
    // coriolis / advection (:723-813)
    double cau;
    const double pv0 = P.potvor[xk], pv1 = P.potvor[xk + s];
    if (P.mommth == 0)
      cau = .125 * (vflux_at(g, P, i, j, k) + vflux_at(g, P, i, j + 1, k) + vflux_at(g, P, i - 1, j, k) +
                    vflux_at(g, P, i - 1, j + 1, k)) * (pv0 + pv1);
    else if (P.mommth == 1)
      cau = .25 * ((vflux_at(g, P, i, j, k) + vflux_at(g, P, i - 1, j, k)) * pv0 +
                   (vflux_at(g, P, i, j + 1, k) + vflux_at(g, P, i - 1, j + 1, k)) * pv1);
    else {
      const double um = utotm_at(g, P, i, j, k);
      double an, ax, bn, bx, temp1, temp2;
      vh_minmax(g, P, i, j + 1, k, an, ax); vh_minmax(g, P, i - 1, j + 1, k, bn, bx);
      if (pv1 * um == 0.) temp1 = pv1 * ((ax + bx) + (an + bn)) * .5;
      else if (pv1 * um < 0.) temp1 = pv1 * (ax + bx);
      else temp1 = pv1 * (an + bn);
      vh_minmax(g, P, i, j, k, an, ax); vh_minmax(g, P, i - 1, j, k, bn, bx);
      if (pv0 * um == 0.) temp2 = pv0 * ((ax + bx) + (an + bn)) * .5;
      else if (pv0 * um < 0.) temp2 = pv0 * (ax + bx);
      else temp2 = pv0 * (an + bn);
      cau = .25 * (temp1 + temp2);
    }
    // lateral stress fluxes with sidewalls (:879-914)
    double wa, wb;
    wgtj_at(g, P, x, k, wa, wb);
    const double un = utotn_at(g, P, i, j, k);
    const double dpxy = fmax(P.dpu[xm], onemm);
    double dpja = fmax(P.dpu[xm - s], onemm); dpja = dpja + wa * (dpxy - dpja);
    double dpjb = fmax(P.dpu[xm + s], onemm); dpjb = dpjb + wb * (dpxy - dpjb);
    const double v2 = P.vsc2u[xk], v4 = P.vsc4u[xk];
    const double v2a = P.iu[x - s] == 0 ? v2 : P.vsc2u[xk - s], v4a = P.iu[x - s] == 0 ? v4 : P.vsc4u[xk - s];
    const double v2b = P.iu[x + s] == 0 ? v2 : P.vsc2u[xk + s], v4b = P.iu[x + s] == 0 ? v4 : P.vsc4u[xk + s];
    const double d2 = P.dl2u[xk];
    const double dl2uja = (1. - wa) * P.dl2u[xk - s] + wa * SLIP * d2;
    const double dl2ujb = (1. - wb) * P.dl2u[xk + s] + wb * SLIP * d2;
    const double uflux2 = fmin(P.difmxq[x], (v2 + v2a) * P.scqx[x]) * hfharm(dpja, dpxy) * (P.uja[xk] - un) +
                          fmin(.125 * P.difmxq[x], (v4 + v4a) * P.scqx[x]) * hfharm(dpja, dpxy) * (dl2uja - d2);
    const double uflux3 = fmin(P.difmxq[x + s], (v2 + v2b) * P.scqx[x + s]) * hfharm(dpjb, dpxy) * (un - P.ujb[xk]) +
                          fmin(.125 * P.difmxq[x + s], (v4 + v4b) * P.scqx[x + s]) * hfharm(dpjb, dpxy) * (d2 - dl2ujb);
    const double uflux1c = P.uflux1[xk], uflux1w = P.uflux1[xk - 1];
    // wind stress (:919-946)
    double stress;
    if (P.isopyc) stress = k == 1 ? -2. * P.taux[x] * grav * P.scux[x] / (P.p[x + L] + P.p[x - 1 + L]) : 0.;
    else stress = -(P.mu_nonloc[xk] - P.mu_nonloc[xk + L]) * P.taux[x] * grav * P.scux[x] / fmax(onemm, P.dpu[xm]);
    // bottom stress, pressure gradient, update (:948-980)
    const double pbum = P.pbu[x + m2];
    const double ptopl = .5 * (fmin(pbum, P.p[xk]) + fmin(pbum, P.p[xk - 1]));
    const double pbotl = .5 * (fmin(pbum, P.p[xk + L]) + fmin(pbum, P.p[xk - 1 + L]));
    const double q = .5 * (P.drag[x] + P.drag[x - 1]) *
                     (fmax(pbum - thkbop, pbotl) - fmax(pbum - thkbop, fmin(ptopl, pbotl - onemm))) / fmax(P.dpu[xm], onemm);
    const double botstr = -un * q / (1. + P.delt1 * q);
    const double pgf = (1. - 2. * WPGF) * P.pgfx[xm] + WPGF * (P.pgfx_o[xk] + P.pgfx[xn]);
    const double ukm = P.u[xm], ukn = P.u[xn];
    P.su_m[xk] = ukm * (WUV1 * P.dpu[xm] + onemm) + ukn * WUV2 * P.dpuold[xk];
    P.su_n[xk] = ukn + P.delt1 * (-P.scuxi[x] * (-pgf + stress + (P.ke[xk] - P.ke[xk - 1])) + cau -
                               P.ubcors_p[x] * P.tsfac + botstr -
                               (uflux1c - uflux1w + uflux3 - uflux2) / (P.scu2[x] * fmax(P.dpu[xm], onemm)));
  }
}
__global__ void __launch_bounds__(128)
mt_update_v(Geom g, MtP P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1, k = blockIdx.z + 1;
  if (i > g.ii) return;
  const long x = ix2(g, i, j), s = g.ldi, L = g.lev;
  const long xk = x + (long)(k - 1) * L, xm = x + (long)(k + P.mm - 1) * L, xn = x + (long)(k + P.nn - 1) * L;
  const long m2 = (long)(P.m - 1) * L;
  const double thkbop = THKBOT * onem;
  if (P.iv[x] == 1) {
    double cav;
    const double pv0 = P.potvor[xk], pv1 = P.potvor[xk + 1];
    if (P.mommth == 0)
      cav = -.125 * (uflux_at(g, P, i, j, k) + uflux_at(g, P, i + 1, j, k) + uflux_at(g, P, i, j - 1, k) +
                     uflux_at(g, P, i + 1, j - 1, k)) * (pv0 + pv1);
    else if (P.mommth == 1)
      cav = -.25 * ((uflux_at(g, P, i, j, k) + uflux_at(g, P, i, j - 1, k)) * pv0 +
                    (uflux_at(g, P, i + 1, j, k) + uflux_at(g, P, i + 1, j - 1, k)) * pv1);
    else {
      const double vm = vtotm_at(g, P, i, j, k);
      double an, ax, bn, bx, temp1, temp2;
      uh_minmax(g, P, i + 1, j, k, an, ax); uh_minmax(g, P, i + 1, j - 1, k, bn, bx);
      if (pv1 * vm == 0.) temp1 = pv1 * ((ax + bx) + (an + bn)) * .5;
      else if (pv1 * vm > 0.) temp1 = pv1 * (ax + bx);
      else temp1 = pv1 * (an + bn);
      uh_minmax(g, P, i, j, k, an, ax); uh_minmax(g, P, i, j - 1, k, bn, bx);
      if (pv0 * vm == 0.) temp2 = pv0 * ((ax + bx) + (an + bn)) * .5;
      else if (pv0 * vm > 0.) temp2 = pv0 * (ax + bx);
      else temp2 = pv0 * (an + bn);
      cav = -.25 * (temp1 + temp2);
    }
    double wa, wb;
    wgti_at(g, P, x, k, wa, wb);
    const double vn = vtotn_at(g, P, i, j, k);
    const double dpxy = fmax(P.dpv[xm], onemm);
    double dpia = fmax(P.dpv[xm - 1], onemm); dpia = dpia + wa * (dpxy - dpia);
    double dpib = fmax(P.dpv[xm + 1], onemm); dpib = dpib + wb * (dpxy - dpib);
    const double v2 = P.vsc2v[xk], v4 = P.vsc4v[xk];
    const double v2a = P.iv[x - 1] == 0 ? v2 : P.vsc2v[xk - 1], v4a = P.iv[x - 1] == 0 ? v4 : P.vsc4v[xk - 1];
    const double v2b = P.iv[x + 1] == 0 ? v2 : P.vsc2v[xk + 1], v4b = P.iv[x + 1] == 0 ? v4 : P.vsc4v[xk + 1];
    const double d2 = P.dl2v[xk];
    const double dl2via = (1. - wa) * P.dl2v[xk - 1] + wa * SLIP * d2;
    const double dl2vib = (1. - wb) * P.dl2v[xk + 1] + wb * SLIP * d2;
    const double vflux2 = fmin(P.difmxq[x], (v2 + v2a) * P.scqy[x]) * hfharm(dpia, dpxy) * (P.via[xk] - vn) +
                          fmin(.125 * P.difmxq[x], (v4 + v4a) * P.scqy[x]) * hfharm(dpia, dpxy) * (dl2via - d2);
    const double vflux3 = fmin(P.difmxq[x + 1], (v2 + v2b) * P.scqy[x + 1]) * hfharm(dpib, dpxy) * (vn - P.vib[xk]) +
                          fmin(.125 * P.difmxq[x + 1], (v4 + v4b) * P.scqy[x + 1]) * hfharm(dpib, dpxy) * (d2 - dl2vib);
    const double vflux1c = P.vflux1[xk], vflux1s = P.vflux1[xk - s];
    double stress;
    if (P.isopyc) stress = k == 1 ? -2. * P.tauy[x] * grav * P.scvy[x] / (P.p[x + L] + P.p[x - s + L]) : 0.;
    else stress = -(P.mv_nonloc[xk] - P.mv_nonloc[xk + L]) * P.tauy[x] * grav * P.scvy[x] / fmax(onemm, P.dpv[xm]);
    const double pbvm = P.pbv[x + m2];
    const double ptopl = .5 * (fmin(pbvm, P.p[xk]) + fmin(pbvm, P.p[xk - s]));
    const double pbotl = .5 * (fmin(pbvm, P.p[xk + L]) + fmin(pbvm, P.p[xk - s + L]));
    const double q = .5 * (P.drag[x] + P.drag[x - s]) *
                     (fmax(pbvm - thkbop, pbotl) - fmax(pbvm - thkbop, fmin(ptopl, pbotl - onemm))) / fmax(P.dpv[xm], onemm);
    const double botstr = -vn * q / (1. + P.delt1 * q);
    const double pgf = (1. - 2. * WPGF) * P.pgfy[xm] + WPGF * (P.pgfy_o[xk] + P.pgfy[xn]);
    const double vkm = P.v[xm], vkn = P.v[xn];
    P.sv_m[xk] = vkm * (WUV1 * P.dpv[xm] + onemm) + vkn * WUV2 * P.dpvold[xk];
    P.sv_n[xk] = vkn + P.delt1 * (-P.scvyi[x] * (-pgf + stress + (P.ke[xk] - P.ke[xk - s])) + cav -
                               P.vbcors_p[x] * P.tsfac + botstr -
                               (vflux1c - vflux1s + vflux3 - vflux2) / (P.scv2[x] * fmax(P.dpv[xm], onemm)));
  }
}

// ---- stage 5: column pass (:1154-1267) ------------------------------------------------------------
__global__ void mt_column(Geom g, MtP P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.ii) return;
  const long x = ix2(g, i, j), L = g.lev, m2 = (long)(P.m - 1) * L;
  const double dt1inv = 1. / P.delt1;
  if (P.iu[x] == 1) {
    const double umx = P.umax[x], ubm = P.ub[x + m2];
    double tot = 0., uprev = 0.;
    for (int k = 1; k <= g.kdm; ++k) {
      const long xm = x + (long)(k + P.mm - 1) * L, xn = x + (long)(k + P.nn - 1) * L;
      const double q = fmin(fmin(P.dpu[xm], P.dpu[xn]), onem);
      double un = P.su_n[x + (long)(k - 1) * L];
      const double ua = k == 1 ? un : uprev;  // kan = max(1,k-1)+nn
      un = (un * q + ua * (onem - q)) / onem;
      un = fmax(-umx, fmin(umx, un + ubm)) - ubm;
      P.u[xn] = un;
      uprev = un;
      tot = tot + un * P.dpu[xn];
    }
    tot = tot / P.pbu_p[x];
    double pk = P.pu[x];
    for (int k = 1; k <= g.kdm; ++k) {
      const long xm = x + (long)(k + P.mm - 1) * L, xn = x + (long)(k + P.nn - 1) * L, xk = x + (long)(k - 1) * L;
      const double un = P.u[xn] - tot;
      P.u[xn] = un;
      P.u[xm] = (P.su_m[xk] + un * WUV2 * P.dpu[xn]) / (WUV1 * P.dpu[xm] + onemm + WUV2 * (P.dpuold[xk] + P.dpu[xn]));
      pk = pk + P.dpu[xn];
      P.pu[x + (long)k * L] = pk;
    }
    P.utotn[x] = tot * dt1inv;
  }
  if (P.iv[x] == 1) {
    const double vmx = P.vmax[x], vbm = P.vb[x + m2];
    double tot = 0., vprev = 0.;
    for (int k = 1; k <= g.kdm; ++k) {
      const long xm = x + (long)(k + P.mm - 1) * L, xn = x + (long)(k + P.nn - 1) * L;
      const double q = fmin(fmin(P.dpv[xm], P.dpv[xn]), onem);
      double vn = P.sv_n[x + (long)(k - 1) * L];
      const double va = k == 1 ? vn : vprev;
      vn = (vn * q + va * (onem - q)) / onem;
      vn = fmax(-vmx, fmin(vmx, vn + vbm)) - vbm;
      P.v[xn] = vn;
      vprev = vn;
      tot = tot + vn * P.dpv[xn];
    }
    tot = tot / P.pbv_p[x];
    double pk = P.pv[x];
    for (int k = 1; k <= g.kdm; ++k) {
      const long xm = x + (long)(k + P.mm - 1) * L, xn = x + (long)(k + P.nn - 1) * L, xk = x + (long)(k - 1) * L;
      const double vn = P.v[xn] - tot;
      P.v[xn] = vn;
      P.v[xm] = (P.sv_m[xk] + vn * WUV2 * P.dpv[xn]) / (WUV1 * P.dpv[xm] + onemm + WUV2 * (P.dpvold[xk] + P.dpv[xn]));
      pk = pk + P.dpv[xn];
      P.pv[x + (long)k * L] = pk;
    }
    P.vtotn[x] = tot * dt1inv;
  }
}

}  // namespace

void momtum_dev(int m, int n, int mm, int nn, int k1m, int k1n) {
  (void)k1m; (void)k1n;
  Ctx& c = C(); const Geom& g = c.g;
  const std::string mommth = c.option("mommth", "enscon");
  MtP P{};
  if (mommth == "enscon") P.mommth = 0;
  else if (mommth == "enecon") P.mommth = 1;
  else if (mommth == "enedis") P.mommth = 2;
  else throw std::runtime_error(" mommth = " + mommth + " is unsupported!");
  P.isopyc = c.option("vcoord", "cntiso_hybrid") == "isopyc_bulkml";
  P.m = m; P.n = n; P.mm = mm; P.nn = nn;
  P.delt1 = c.scalar("delt1"); P.tsfac = c.scalar("dlt") / P.delt1;
  P.mdv2hi = c.scalar("mdv2hi", 0.); P.mdv2lo = c.scalar("mdv2lo", 0.); P.mdv4hi = c.scalar("mdv4hi", 0.);
  P.mdv4lo = c.scalar("mdv4lo", 0.); P.vsc2hi = c.scalar("vsc2hi", 0.); P.vsc2lo = c.scalar("vsc2lo", 0.);
  P.vsc4hi = c.scalar("vsc4hi", 0.); P.vsc4lo = c.scalar("vsc4lo", 0.); P.cbar = c.scalar("cbar", 0.);
  P.cb = c.scalar("cb", 0.);
#define D(f) P.f = c.dev(#f)
  D(u); D(v); D(p); D(pu); D(pv); D(absvor); D(dpvor); D(utotn); D(vtotn); D(ustarb);
  D(dp); D(dpu); D(dpv); D(pbu); D(pbv); D(ubflxs_p); D(vbflxs_p); D(ub); D(vb); D(pgfx); D(pgfy); D(pgfx_o);
  D(pgfy_o); D(dpuold); D(dpvold); D(mu_nonloc); D(mv_nonloc); D(ubcors_p); D(vbcors_p); D(pbu_p); D(pbv_p);
  D(difwgt); D(difmxp); D(difmxq); D(taux); D(tauy); D(umax); D(vmax);
  D(scuy); D(scvx); D(scux); D(scvy); D(scq2i); D(scp2i); D(scp2); D(scu2); D(scv2); D(scpx); D(scpy); D(scqx);
  D(scqy); D(scuxi); D(scvyi); D(corioq);
#undef D
  P.ip = c.idev("ip"); P.iu = c.idev("iu"); P.iv = c.idev("iv"); P.iq = c.idev("iq");
#define S(f) P.f = c.owned("momtum_" #f, g.kdm)
  S(uja); S(ujb); S(via); S(vib); S(dl2u); S(dl2v); S(defor1); S(defor2); S(potvor); S(vsc2u); S(vsc4u); S(vsc2v);
  S(vsc4v); S(su_m); S(su_n); S(sv_m); S(sv_n); S(ke); S(uflux1); S(vflux1);
#undef S
  P.drag = c.owned("momtum_drag", 1);
  P.ubn = c.owned("momtum_ubn", 1); P.ubm = c.owned("momtum_ubm", 1);
  P.vbn = c.owned("momtum_vbn", 1); P.vbm = c.owned("momtum_vbm", 1);

  { dim3 grid(cdiv(g.ii + 4, 128), g.jj + 4); LAUNCH(mt_pressures, grid, 128, 0, g, P); }
  { dim3 grid(cdiv(g.ii + 1, 128), g.jj + 1); LAUNCH(mt_drag, grid, 128, 0, g, P); }
  halo_update(c.dev("difwgt"), 1, 2, 2, halo_ps);
  { dim3 grid(cdiv(g.ii + 4, 128), g.jj + 4, g.kdm); LAUNCH(mt_aux, grid, 128, 0, g, P); }
  { dim3 grid(cdiv(g.ii + 3, 128), g.jj + 3, g.kdm); LAUNCH(mt_vort, grid, 128, 0, g, P); }
  { dim3 grid(cdiv(g.ii + 2, 128), g.jj + 2, g.kdm); LAUNCH(mt_visc, grid, 128, 0, g, P); }
  { dim3 grid(cdiv(g.ii + 1, 128), g.jj + 1, g.kdm); LAUNCH(mt_flux1, grid, 128, 0, g, P); }
  { dim3 grid(cdiv(g.ii, 128), g.jj, g.kdm);
    LAUNCH(mt_update, grid, 128, 0, g, P);
    LAUNCH(mt_update_v, grid, 128, 0, g, P); }
  { dim3 grid(cdiv(g.ii, 128), g.jj); LAUNCH(mt_column, grid, 128, 0, g, P); }
}

}  // namespace blom
