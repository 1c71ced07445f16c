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
// The loads are unconditional (every address touched lies inside the halo-padded arrays) and the
// mask only selects: a thread's loads can then be issued back to back instead of as a chain of
// mask-load -> branch -> value-load round trips (ncu: long-scoreboard stalls dominated these kernels).
__device__ __forceinline__ double utotn_at(const Geom& g, const MtP& P, int i, int j, int k) {
  const long x = ix2(g, i, j);
  const int w = P.iu[x];
  const double a = P.u[x + (long)(k + P.nn - 1) * g.lev], b = P.ubn[x], c = P.utotn[x];
  return (i >= -1 && i <= g.ii + 2 && j >= -1 && j <= g.jj + 2 && w == 1) ? a + b : c;
}
__device__ __forceinline__ double vtotn_at(const Geom& g, const MtP& P, int i, int j, int k) {
  const long x = ix2(g, i, j);
  const int w = P.iv[x];
  const double a = P.v[x + (long)(k + P.nn - 1) * g.lev], b = P.vbn[x], c = P.vtotn[x];
  return (i >= -1 && i <= g.ii + 2 && j >= -1 && j <= g.jj + 2 && w == 1) ? a + b : c;
}
__device__ __forceinline__ double utotm_at(const Geom& g, const MtP& P, int i, int j, int k) {
  const long x = ix2(g, i, j);
  const int w = P.iu[x];
  const double a = P.u[x + (long)(k + P.mm - 1) * g.lev], b = P.ubm[x];
  return (i >= 0 && i <= g.ii + 1 && j >= 0 && j <= g.jj + 1 && w == 1) ? a + b : 0.;
}
__device__ __forceinline__ double vtotm_at(const Geom& g, const MtP& P, int i, int j, int k) {
  const long x = ix2(g, i, j);
  const int w = P.iv[x];
  const double a = P.v[x + (long)(k + P.mm - 1) * g.lev], b = P.vbm[x];
  return (i >= 0 && i <= g.ii + 1 && j >= 0 && j <= g.jj + 1 && w == 1) ? a + b : 0.;
}
__device__ __forceinline__ double uflux_at(const Geom& g, const MtP& P, int i, int j, int k) {
  const long x = ix2(g, i, j), xm = x + (long)(k + P.mm - 1) * g.lev;
  const int w = P.iu[x];
  const double a = P.u[xm], b = P.ubm[x], d = P.dpu[xm];
  return (i >= 0 && i <= g.ii + 1 && j >= 0 && j <= g.jj + 1 && w == 1) ? (a + b) * fmax(d, onem) : 0.;
}
__device__ __forceinline__ double vflux_at(const Geom& g, const MtP& P, int i, int j, int k) {
  const long x = ix2(g, i, j), xm = x + (long)(k + P.mm - 1) * g.lev;
  const int w = P.iv[x];
  const double a = P.v[xm], b = P.vbm[x], d = P.dpv[xm];
  return (i >= 0 && i <= g.ii + 1 && j >= 0 && j <= g.jj + 1 && w == 1) ? (a + b) * fmax(d, onem) : 0.;
}
// dpmx at q-point (i,j), 0<=i<=ii+2, 0<=j<=jj+2 (:360-396)
__device__ __forceinline__ double dpmx_at(const Geom& g, const MtP& P, int i, int j, int k) {
  const long x = ix2(g, i, j), s = g.ldi;
  const double* dpm = P.dp + (long)(k + P.mm - 1) * g.lev;
  const int u0 = P.iu[x], u1 = P.iu[x - s], v0 = P.iv[x], v1 = P.iv[x - 1];
  const double d00 = dpm[x], d10 = dpm[x - 1], d01 = dpm[x - s], d11 = dpm[x - s - 1];
  double r = 8. * onem;
  r = u0 == 1 ? fmax(r, d00 + d10) : r;
  r = u1 == 1 ? fmax(r, d01 + d11) : r;
  r = v0 == 1 ? fmax(r, d00 + d01) : r;
  r = v1 == 1 ? fmax(r, d10 + d11) : r;
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
template <int MINB>
__global__ void __launch_bounds__(128, MINB) mt_aux(Geom g, MtP P) {
  const Bid b_ = bid(g);
  const int i = b_.x * blockDim.x + threadIdx.x - 1;  // -1..ii+2
  const int j = b_.y - 1, k = b_.z + 1;    // -1..jj+2
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
template <int MINB>
__global__ void __launch_bounds__(128, MINB) mt_vort(Geom g, MtP P) {
  const Bid b_ = bid(g);
  const int i = b_.x * blockDim.x + threadIdx.x;  // 0..ii+2
  const int j = b_.y, k = b_.z + 1;         // 0..jj+2
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
  const Bid b_ = bid(g);
  const int i = b_.x * blockDim.x + threadIdx.x;  // 0..ii+1
  const int j = b_.y, k = b_.z + 1;         // 0..jj+1
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
  const int w0 = P.iu[x], wp = P.iu[x + 1], wm = P.iu[x - 1];
  const double c = vs[xk], e = vs[xk + 1], w = vs[xk - 1];
  // priority: own value, else the eastern span's first value, else the western span's last value
  double r = (wm == 1 && i - 1 < g.ii + 1 && i - 1 >= 0) ? w : 0.;
  r = (wp == 1 && i + 1 > 0) ? (i + 1 <= g.ii + 1 ? e : 0.) : r;
  return w0 == 1 ? c : r;
}
__device__ __forceinline__ double viscv_ext(const Geom& g, const MtP& P, const double* vs, int i, int j, long xk) {
  const long x = ix2(g, i, j), s = g.ldi;
  const int w0 = P.iv[x], wp = P.iv[x + s], wm = P.iv[x - s];
  const double c = vs[xk], n = vs[xk + s], so = vs[xk - s];
  double r = (wm == 1 && j - 1 < g.jj + 1 && j - 1 >= 0) ? so : 0.;
  r = (wp == 1 && j + 1 > 0) ? (j + 1 <= g.jj + 1 ? n : 0.) : r;
  return w0 == 1 ? c : r;
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
template <int MINB>
__global__ void __launch_bounds__(128, MINB) mt_flux1(Geom g, MtP P) {
  const Bid b_ = bid(g);
  const int i = b_.x * blockDim.x + threadIdx.x;  // 0..ii
  const int j = b_.y, k = b_.z + 1;         // 0..jj
  if (i > g.ii) return;
  const long xk = ix2(g, i, j) + (long)(k - 1) * g.lev;
  if (j >= 1) P.uflux1[xk] = uflux1_at(g, P, i, j, k);
  if (i >= 1) P.vflux1[xk] = vflux1_at(g, P, i, j, k);
}

// ---- stage 4: tendencies and leap-frog update ------------------------------------------------------
template <int MINB>
__global__ void __launch_bounds__(128, MINB)
mt_update(Geom g, MtP P) {
  const Bid b_ = bid(g);
  const int i = b_.x * blockDim.x + threadIdx.x + 1, j = b_.y + 1, k = b_.z + 1;
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
    const int iua = P.iu[x - s], iub = P.iu[x + s];
    const double l2a = P.vsc2u[xk - s], l4a = P.vsc4u[xk - s], l2b = P.vsc2u[xk + s], l4b = P.vsc4u[xk + s];
    const double v2a = iua == 0 ? v2 : l2a, v4a = iua == 0 ? v4 : l4a;
    const double v2b = iub == 0 ? v2 : l2b, v4b = iub == 0 ? v4 : l4b;
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
template <int MINB>
__global__ void __launch_bounds__(128, MINB)
mt_update_v(Geom g, MtP P) {
  const Bid b_ = bid(g);
  const int i = b_.x * blockDim.x + threadIdx.x + 1, j = b_.y + 1, k = b_.z + 1;
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
    const int iva = P.iv[x - 1], ivb = P.iv[x + 1];
    const double l2a = P.vsc2v[xk - 1], l4a = P.vsc4v[xk - 1], l2b = P.vsc2v[xk + 1], l4b = P.vsc4v[xk + 1];
    const double v2a = iva == 0 ? v2 : l2a, v4a = iva == 0 ? v4 : l4a;
    const double v2b = ivb == 0 ? v2 : l2b, v4b = ivb == 0 ? v4 : l4b;
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

// ---- fused per-level tile kernel ------------------------------------------------------------------
// Stages 1-4 for one (TX x TY) tile of one layer in ONE launch: every layer-sized scratch array of
// the staged form (uja..vib, dl2u/v, defor1/2, potvor, vsc2/4 at u and v, ke, uflux1, vflux1) lives
// in shared memory on a tile with a 3-point halo skirt (the widest dependency: update <- uflux1 <-
// span-extended viscosity at i+-2 <- defor2 <- uja/ujb <- utotn at +-3).  Values outside the
// reference's loop ranges are 0 exactly like the zero-initialised scratch arrays of the staged form,
// so both forms are bit-identical.  Global traffic per cell: ~24 R + 6 W words instead of ~95.
template <int TX, int TY>
struct MtTile {
  static constexpr int H = 3, SW = TX + 2 * H, SH = TY + 2 * H, N = SW * SH, NT = TX * TY, NARR = 21;
  static constexpr size_t bytes = (size_t)NARR * N * sizeof(double) + ((N + 15) / 16) * 16;
};
enum : unsigned { MK_P = 1, MK_U = 2, MK_V = 4, MK_Q = 8, MK_RM = 16 /*0..ii+1 x 0..jj+1*/, MK_RN = 32 /*-1..ii+2 x -1..jj+2*/ };

template <int SW>
__device__ __forceinline__ double t_dpmx(const double* dpm, const unsigned char* mk, int c) {
  double r = 8. * onem;
  if (mk[c] & MK_U) r = fmax(r, dpm[c] + dpm[c - 1]);
  if (mk[c - SW] & MK_U) r = fmax(r, dpm[c - SW] + dpm[c - SW - 1]);
  if (mk[c] & MK_V) r = fmax(r, dpm[c] + dpm[c - SW]);
  if (mk[c - 1] & MK_V) r = fmax(r, dpm[c - 1] + dpm[c - 1 - SW]);
  return r;
}
// viscu_ext / viscv_ext on the tile (i resp. j is the global index of cell c)
__device__ __forceinline__ double t_viscu_ext(const Geom& g, const unsigned char* mk, const double* vs, int i, int c) {
  if (mk[c] & MK_U) return vs[c];
  if ((mk[c + 1] & MK_U) && i + 1 > 0) return i + 1 <= g.ii + 1 ? vs[c + 1] : 0.;
  if ((mk[c - 1] & MK_U) && i - 1 < g.ii + 1) return i - 1 >= 0 ? vs[c - 1] : 0.;
  return 0.;
}
template <int SW>
__device__ __forceinline__ double t_viscv_ext(const Geom& g, const unsigned char* mk, const double* vs, int j, int c) {
  if (mk[c] & MK_V) return vs[c];
  if ((mk[c + SW] & MK_V) && j + 1 > 0) return j + 1 <= g.jj + 1 ? vs[c + SW] : 0.;
  if ((mk[c - SW] & MK_V) && j - 1 < g.jj + 1) return j - 1 >= 0 ? vs[c - SW] : 0.;
  return 0.;
}
__device__ __forceinline__ double t_flux(const unsigned char* mk, unsigned bit, const double* tot, const double* dpf, int c) {
  return ((mk[c] & bit) && (mk[c] & MK_RM)) ? tot[c] * fmax(dpf[c], onem) : 0.;
}
// uh_minmax / vh_minmax on the tile: tot = um|vm, dpf = dpum|dpvm, off = 1 | SW
__device__ __forceinline__ void t_minmax(const unsigned char* mk, unsigned bit, const double* tot, const double* dpf,
                                         const double* dpm, int off, int c, double& mn, double& mx) {
  mn = 0.; mx = 0.;
  if (!((mk[c] & bit) && (mk[c] & MK_RM))) return;
  double hc = .5 * tot[c] * (dpm[c] + dpm[c - off]), hm = tot[c] * fmax(dpf[c], onem);
  const double c1 = 1. - 1.5 * .5, c2 = 1. - .5, c3 = 2., slope = .5;
  if (fabs(hc) < .1 * fabs(hm)) hm = 10. * hc;
  else if (fabs(hc) > c1 * fabs(hm)) {
    if (fabs(hc) < c2 * fabs(hm)) hc = (3. * hc + (1. - c2 * 3.) * hm);
    else if (fabs(hc) <= c3 * fabs(hm)) hc = hm;
    else hc = slope * hc + (1. - c3 * slope) * hm;
  }
  if (hc > hm) { mn = hm; mx = hc; } else { mx = hm; mn = hc; }
}

template <int TX, int TY>
__global__ void __launch_bounds__(TX * TY, 2)
mt_level(Geom g, MtP P) {
  using T = MtTile<TX, TY>;
  constexpr int H = T::H, SW = T::SW, SH = T::SH, N = T::N, NT = T::NT;
  extern __shared__ double mt_sm[];
  double *un = mt_sm, *vn = un + N, *um = vn + N, *vm = um + N, *dpm = vm + N, *dpum = dpm + N, *dpvm = dpum + N,
         *uja = dpvm + N, *ujb = uja + N, *via = ujb + N, *vib = via + N, *dl2u = vib + N, *dl2v = dl2u + N,
         *d1 = dl2v + N, *d2 = d1 + N, *pvq = d2 + N, *v2u = pvq + N, *v4u = v2u + N, *v2v = v4u + N, *v4v = v2v + N,
         *ke = v4v + N;
  unsigned char* mk = reinterpret_cast<unsigned char*>(ke + N);
  const int ti = blockIdx.x * TX + 1, tj = blockIdx.y * TY + 1, k = blockIdx.z + 1;
  const int tid = threadIdx.x;
  const long L = g.lev, s = g.ldi;
  const long okn = (long)(k + P.nn - 1) * L, okm = (long)(k + P.mm - 1) * L, ok = (long)(k - 1) * L;
  const long m2 = (long)(P.m - 1) * L;

  // ---- A: total velocities, layer thicknesses and masks on the full tile (halo 3)
  for (int c = tid; c < N; c += NT) {
    const int lj = c / SW, li = c - lj * SW;
    const int i = ti - H + li, j = tj - H + lj;
    double a_un = 0., a_vn = 0., a_um = 0., a_vm = 0., a_dpm = 0., a_dpum = 0., a_dpvm = 0.;
    unsigned m = 0;
    if (i <= g.ii + g.nb && j <= g.jj + g.nb) {
      const long x = ix2(g, i, j);
      const bool rn = i >= -1 && i <= g.ii + 2 && j >= -1 && j <= g.jj + 2;
      const bool rm = i >= 0 && i <= g.ii + 1 && j >= 0 && j <= g.jj + 1;
      const bool wu = P.iu[x] == 1, wv = P.iv[x] == 1;
      m = (P.ip[x] == 1 ? MK_P : 0u) | (wu ? MK_U : 0u) | (wv ? MK_V : 0u) | (P.iq[x] == 1 ? MK_Q : 0u) |
          (rm ? MK_RM : 0u) | (rn ? MK_RN : 0u);
      a_un = (rn && wu) ? P.u[x + okn] + P.ubn[x] : P.utotn[x];
      a_vn = (rn && wv) ? P.v[x + okn] + P.vbn[x] : P.vtotn[x];
      if (rm && wu) a_um = P.u[x + okm] + P.ubm[x];
      if (rm && wv) a_vm = P.v[x + okm] + P.vbm[x];
      a_dpm = P.dp[x + okm]; a_dpum = P.dpu[x + okm]; a_dpvm = P.dpv[x + okm];
    }
    un[c] = a_un; vn[c] = a_vn; um[c] = a_um; vm[c] = a_vm; dpm[c] = a_dpm; dpum[c] = a_dpum; dpvm[c] = a_dpvm;
    mk[c] = (unsigned char)m;
  }
  __syncthreads();

  // ---- B: sidewall-weighted neighbours, del2, tension, kinetic energy (mt_aux)
  for (int c = tid; c < N; c += NT) {
    const int lj = c / SW, li = c - lj * SW;
    const int i = ti - H + li, j = tj - H + lj;
    const unsigned m = mk[c];
    const long x = ix2(g, i, j);
    double ua = 0., ub = 0., du = 0., va = 0., vb = 0., dv = 0., df1 = 0., kev = 0.;
    if ((m & MK_U) && (m & MK_RN) && i >= 0) {
      double wa, wb;
      wgtj_at(g, P, x, k, wa, wb);
      const double c0 = un[c];
      if (lj >= 1) ua = (1. - wa) * un[c - SW] + wa * SLIP * c0;
      if (lj <= SH - 2) ub = (1. - wb) * un[c + SW] + wb * SLIP * c0;
      if (li >= 1 && li <= SW - 2 && lj >= 1 && lj <= SH - 2) du = c0 - .25 * (un[c + 1] + un[c - 1] + ua + ub);
    }
    if ((m & MK_V) && (m & MK_RN) && j >= 0) {
      double wa, wb;
      wgti_at(g, P, x, k, wa, wb);
      const double c0 = vn[c];
      if (li >= 1) va = (1. - wa) * vn[c - 1] + wa * SLIP * c0;
      if (li <= SW - 2) vb = (1. - wb) * vn[c + 1] + wb * SLIP * c0;
      if (li >= 1 && li <= SW - 2 && lj >= 1 && lj <= SH - 2) dv = c0 - .25 * (vn[c + SW] + vn[c - SW] + va + vb);
    }
    if (li <= SW - 2 && lj <= SH - 2) {
      if ((m & MK_P) && (m & MK_RN) && i <= g.ii + 1 && j <= g.jj + 1)
        df1 = sq((un[c + 1] * P.scuy[x + 1] - un[c] * P.scuy[x]) - (vn[c + SW] * P.scvx[x + s] - vn[c] * P.scvx[x])) *
              P.scp2i[x];
      if (i >= 0 && i <= g.ii && j >= 0 && j <= g.jj)
        kev = .25 * (P.scu2[x] * sq(um[c]) + P.scu2[x + 1] * sq(um[c + 1]) + P.scv2[x] * sq(vm[c]) +
                     P.scv2[x + s] * sq(vm[c + SW])) / P.scp2[x];
    }
    uja[c] = ua; ujb[c] = ub; dl2u[c] = du; via[c] = va; vib[c] = vb; dl2v[c] = dv; d1[c] = df1; ke[c] = kev;
  }
  __syncthreads();

  // ---- C: q-point gather: shear deformation, vorticity, dpvor, potential vorticity (mt_vort)
  for (int c = tid; c < N; c += NT) {
    const int lj = c / SW, li = c - lj * SW;
    const int i = ti - H + li, j = tj - H + lj;
    double df2 = 0., pq = 0.;
    if (li >= 1 && lj >= 1 && i >= 0 && i <= g.ii + 2 && j >= 0 && j <= g.jj + 2) {
      const long x = ix2(g, i, j);
      const unsigned m = mk[c], mw = mk[c - 1], ms = mk[c - SW];
      const double q2i = P.scq2i[x];
      const bool vfirst = (m & MK_V) && !(mw & MK_V), vlast = (mw & MK_V) && !(m & MK_V);
      const bool ufirst = (m & MK_U) && !(ms & MK_U), ulast = (ms & MK_U) && !(m & MK_U);
      const bool isq = (m & MK_Q) != 0;
      if (vfirst) df2 = sq(vn[c] * (1. - SLIP) * P.scvy[x]) * q2i;
      if (vlast) df2 = sq(vn[c - 1] * (1. - SLIP) * P.scvy[x - 1]) * q2i;
      if (ufirst) df2 = sq(un[c] * (1. - SLIP) * P.scux[x]) * q2i;
      if (ulast) df2 = sq(un[c - SW] * (1. - SLIP) * P.scux[x - s]) * q2i;
      if (isq)
        df2 = sq(vib[c - 1] * P.scvy[x] - via[c] * P.scvy[x - 1] + ujb[c - SW] * P.scux[x] - uja[c] * P.scux[x - s]) * q2i;
      if (li >= H && li <= H + TX && lj >= H && lj <= H + TY && i >= 1 && i <= g.ii + 1 && j >= 1 && j <= g.jj + 1) {
        bool set = false; double vort = 0., dpv = 1.;
        if (vfirst) {
          vort = vm[c] * (1. - SLIP) * P.scvy[x] * q2i;
          dpv = .125 * fmax(fmax(4. * (dpm[c] + dpm[c - SW]), t_dpmx<SW>(dpm, mk, c)), t_dpmx<SW>(dpm, mk, c + 1));
          set = true;
        }
        if (vlast) {
          vort = -vm[c - 1] * (1. - SLIP) * P.scvy[x - 1] * q2i;
          dpv = .125 * fmax(fmax(4. * (dpm[c - 1] + dpm[c - 1 - SW]), t_dpmx<SW>(dpm, mk, c - 1)), t_dpmx<SW>(dpm, mk, c));
          set = true;
        }
        if (ufirst) {
          vort = -um[c] * (1. - SLIP) * P.scux[x] * q2i;
          dpv = .125 * fmax(fmax(4. * (dpm[c] + dpm[c - 1]), t_dpmx<SW>(dpm, mk, c)), t_dpmx<SW>(dpm, mk, c + SW));
          set = true;
        }
        if (ulast) {
          vort = um[c - SW] * (1. - SLIP) * P.scux[x - s] * q2i;
          dpv = .125 * fmax(fmax(4. * (dpm[c - SW] + dpm[c - SW - 1]), t_dpmx<SW>(dpm, mk, c - SW)), t_dpmx<SW>(dpm, mk, c));
          set = true;
        }
        if (isq) {
          vort = (vm[c] * P.scvy[x] - vm[c - 1] * P.scvy[x - 1] - um[c] * P.scux[x] + um[c - SW] * P.scux[x - s]) * q2i;
          double mx = 2. * (dpm[c] + dpm[c - 1] + dpm[c - SW] + dpm[c - SW - 1]);
          mx = fmax(mx, t_dpmx<SW>(dpm, mk, c)); mx = fmax(mx, t_dpmx<SW>(dpm, mk, c - 1));
          mx = fmax(mx, t_dpmx<SW>(dpm, mk, c + 1)); mx = fmax(mx, t_dpmx<SW>(dpm, mk, c - SW));
          mx = fmax(mx, t_dpmx<SW>(dpm, mk, c + SW));
          dpv = .125 * mx;
          set = true;
        }
        if (set) {
          const double av = vort + P.corioq[x];
          pq = av / dpv;
          // the tile that owns the q-point writes the diagnostics (last tiles also own ii+1 / jj+1)
          const bool own_i = li < H + TX || ti + TX == g.ii + 1, own_j = lj < H + TY || tj + TY == g.jj + 1;
          if (own_i && own_j) { P.absvor[x + ok] = av; P.dpvor[x + ok] = dpv; }
        }
      }
    }
    d2[c] = df2; pvq[c] = pq;
  }
  __syncthreads();

  // ---- D: deformation-dependent viscosities (mt_visc)
  for (int c = tid; c < N; c += NT) {
    const int lj = c / SW, li = c - lj * SW;
    const int i = ti - H + li, j = tj - H + lj;
    const unsigned m = mk[c];
    double a2 = 0., a4 = 0., b2 = 0., b4 = 0.;
    if (li >= 1 && li <= SW - 2 && lj >= 1 && lj <= SH - 2 && (m & MK_RM)) {
      const long x = ix2(g, i, j);
      if (m & MK_U) {
        const double q = .5 * (P.difwgt[x - 1] + P.difwgt[x]);
        const double deform = sqrt(.5 * (d1[c] + d1[c - 1] + d2[c] + d2[c + SW]));
        a2 = fmax(q * P.mdv2hi + (1. - q) * P.mdv2lo, (q * P.vsc2hi + (1. - q) * P.vsc2lo) * deform);
        a4 = fmax(q * P.mdv4hi + (1. - q) * P.mdv4lo, (q * P.vsc4hi + (1. - q) * P.vsc4lo) * deform);
      }
      if (m & MK_V) {
        const double q = .5 * (P.difwgt[x - s] + P.difwgt[x]);
        const double deform = sqrt(.5 * (d1[c] + d1[c - SW] + d2[c] + d2[c + 1]));
        b2 = fmax(q * P.mdv2hi + (1. - q) * P.mdv2lo, (q * P.vsc2hi + (1. - q) * P.vsc2lo) * deform);
        b4 = fmax(q * P.mdv4hi + (1. - q) * P.mdv4lo, (q * P.vsc4hi + (1. - q) * P.vsc4lo) * deform);
      }
    }
    v2u[c] = a2; v4u[c] = a4; v2v[c] = b2; v4v[c] = b4;
  }
  __syncthreads();

  // ---- E: longitudinal stress fluxes at mass points (mt_flux1); they replace defor1/defor2 in d1/d2
  for (int c = tid; c < N; c += NT) {
    const int lj = c / SW, li = c - lj * SW;
    const int i = ti - H + li, j = tj - H + lj;
    double f1 = 0., g1 = 0.;
    if (li >= H - 1 && li <= H + TX - 1 && lj >= H - 1 && lj <= H + TY - 1 && i >= 0 && i <= g.ii && j >= 0 && j <= g.jj) {
      const long x = ix2(g, i, j);
      const double dmx = P.difmxp[x];
      if (j >= 1) {
        const double dpxy = fmax(dpum[c], onemm), dpib = fmax(dpum[c + 1], onemm);
        const double v2 = t_viscu_ext(g, mk, v2u, i, c) + t_viscu_ext(g, mk, v2u, i + 1, c + 1);
        const double v4 = t_viscu_ext(g, mk, v4u, i, c) + t_viscu_ext(g, mk, v4u, i + 1, c + 1);
        const double hh = hfharm(dpxy, dpib);
        f1 = fmin(dmx, v2 * P.scpy[x]) * hh * (un[c] - un[c + 1]) +
             fmin(.125 * dmx, v4 * P.scpy[x]) * hh * (dl2u[c] - dl2u[c + 1]);
      }
      if (i >= 1) {
        const double dpxy = fmax(dpvm[c], onemm), dpjb = fmax(dpvm[c + SW], onemm);
        const double v2 = t_viscv_ext<SW>(g, mk, v2v, j, c) + t_viscv_ext<SW>(g, mk, v2v, j + 1, c + SW);
        const double v4 = t_viscv_ext<SW>(g, mk, v4v, j, c) + t_viscv_ext<SW>(g, mk, v4v, j + 1, c + SW);
        const double hh = hfharm(dpxy, dpjb);
        g1 = fmin(dmx, v2 * P.scpx[x]) * hh * (vn[c] - vn[c + SW]) +
             fmin(.125 * dmx, v4 * P.scpx[x]) * hh * (dl2v[c] - dl2v[c + SW]);
      }
    }
    d1[c] = f1; d2[c] = g1;  // stage D (the last reader of defor1/defor2) is behind the barrier above
  }
  __syncthreads();

  // ---- F: tendencies and leap-frog update of the tile's own cells (mt_update, mt_update_v)
  const int tx = tid % TX, ty = tid / TX;
  const int i = ti + tx, j = tj + ty;
  if (i > g.ii || j > g.jj) return;
  const int c = (ty + H) * SW + tx + H;
  const unsigned m = mk[c];
  const long x = ix2(g, i, j);
  const long xk = x + ok, xm = x + okm, xn = x + okn;
  const double thkbop = THKBOT * onem;
  if (m & MK_U) {
    double cau;
    const double pv0 = pvq[c], pv1 = pvq[c + SW];
    if (P.mommth == 0)
      cau = .125 * (t_flux(mk, MK_V, vm, dpvm, c) + t_flux(mk, MK_V, vm, dpvm, c + SW) + t_flux(mk, MK_V, vm, dpvm, c - 1) +
                    t_flux(mk, MK_V, vm, dpvm, c - 1 + SW)) * (pv0 + pv1);
    else if (P.mommth == 1)
      cau = .25 * ((t_flux(mk, MK_V, vm, dpvm, c) + t_flux(mk, MK_V, vm, dpvm, c - 1)) * pv0 +
                   (t_flux(mk, MK_V, vm, dpvm, c + SW) + t_flux(mk, MK_V, vm, dpvm, c - 1 + SW)) * pv1);
    else {
      const double umc = um[c];
      double an, ax, bn, bx, temp1, temp2;
      t_minmax(mk, MK_V, vm, dpvm, dpm, SW, c + SW, an, ax); t_minmax(mk, MK_V, vm, dpvm, dpm, SW, c - 1 + SW, bn, bx);
      if (pv1 * umc == 0.) temp1 = pv1 * ((ax + bx) + (an + bn)) * .5;
      else if (pv1 * umc < 0.) temp1 = pv1 * (ax + bx);
      else temp1 = pv1 * (an + bn);
      t_minmax(mk, MK_V, vm, dpvm, dpm, SW, c, an, ax); t_minmax(mk, MK_V, vm, dpvm, dpm, SW, c - 1, bn, bx);
      if (pv0 * umc == 0.) temp2 = pv0 * ((ax + bx) + (an + bn)) * .5;
      else if (pv0 * umc < 0.) temp2 = pv0 * (ax + bx);
      else temp2 = pv0 * (an + bn);
      cau = .25 * (temp1 + temp2);
    }
    double wa, wb;
    wgtj_at(g, P, x, k, wa, wb);
    const double unc = un[c];
    const double dpuxm = dpum[c];
    const double dpxy = fmax(dpuxm, onemm);
    double dpja = fmax(dpum[c - SW], onemm); dpja = dpja + wa * (dpxy - dpja);
    double dpjb = fmax(dpum[c + SW], onemm); dpjb = dpjb + wb * (dpxy - dpjb);
    const double v2 = v2u[c], v4 = v4u[c];
    const bool dry_a = !(mk[c - SW] & MK_U), dry_b = !(mk[c + SW] & MK_U);
    const double v2a = dry_a ? v2 : v2u[c - SW], v4a = dry_a ? v4 : v4u[c - SW];
    const double v2b = dry_b ? v2 : v2u[c + SW], v4b = dry_b ? v4 : v4u[c + SW];
    const double dd = dl2u[c];
    const double dl2uja = (1. - wa) * dl2u[c - SW] + wa * SLIP * dd;
    const double dl2ujb = (1. - wb) * dl2u[c + SW] + wb * SLIP * dd;
    const double uflux2 = fmin(P.difmxq[x], (v2 + v2a) * P.scqx[x]) * hfharm(dpja, dpxy) * (uja[c] - unc) +
                          fmin(.125 * P.difmxq[x], (v4 + v4a) * P.scqx[x]) * hfharm(dpja, dpxy) * (dl2uja - dd);
    const double uflux3 = fmin(P.difmxq[x + s], (v2 + v2b) * P.scqx[x + s]) * hfharm(dpjb, dpxy) * (unc - ujb[c]) +
                          fmin(.125 * P.difmxq[x + s], (v4 + v4b) * P.scqx[x + s]) * hfharm(dpjb, dpxy) * (dd - dl2ujb);
    const double uflux1c = d1[c], uflux1w = d1[c - 1];
    double stress;
    if (P.isopyc) stress = k == 1 ? -2. * P.taux[x] * grav * P.scux[x] / (P.p[x + L] + P.p[x - 1 + L]) : 0.;
    else stress = -(P.mu_nonloc[xk] - P.mu_nonloc[xk + L]) * P.taux[x] * grav * P.scux[x] / fmax(onemm, dpuxm);
    const double pbum = P.pbu[x + m2];
    const double ptopl = .5 * (fmin(pbum, P.p[xk]) + fmin(pbum, P.p[xk - 1]));
    const double pbotl = .5 * (fmin(pbum, P.p[xk + L]) + fmin(pbum, P.p[xk - 1 + L]));
    const double q = .5 * (P.drag[x] + P.drag[x - 1]) *
                     (fmax(pbum - thkbop, pbotl) - fmax(pbum - thkbop, fmin(ptopl, pbotl - onemm))) / fmax(dpuxm, onemm);
    const double botstr = -unc * q / (1. + P.delt1 * q);
    const double pgf = (1. - 2. * WPGF) * P.pgfx[xm] + WPGF * (P.pgfx_o[xk] + P.pgfx[xn]);
    const double ukm = P.u[xm], ukn = P.u[xn];
    P.su_m[xk] = ukm * (WUV1 * dpuxm + onemm) + ukn * WUV2 * P.dpuold[xk];
    P.su_n[xk] = ukn + P.delt1 * (-P.scuxi[x] * (-pgf + stress + (ke[c] - ke[c - 1])) + cau -
                               P.ubcors_p[x] * P.tsfac + botstr -
                               (uflux1c - uflux1w + uflux3 - uflux2) / (P.scu2[x] * fmax(dpuxm, onemm)));
  }
  if (m & MK_V) {
    double cav;
    const double pv0 = pvq[c], pv1 = pvq[c + 1];
    if (P.mommth == 0)
      cav = -.125 * (t_flux(mk, MK_U, um, dpum, c) + t_flux(mk, MK_U, um, dpum, c + 1) + t_flux(mk, MK_U, um, dpum, c - SW) +
                     t_flux(mk, MK_U, um, dpum, c + 1 - SW)) * (pv0 + pv1);
    else if (P.mommth == 1)
      cav = -.25 * ((t_flux(mk, MK_U, um, dpum, c) + t_flux(mk, MK_U, um, dpum, c - SW)) * pv0 +
                    (t_flux(mk, MK_U, um, dpum, c + 1) + t_flux(mk, MK_U, um, dpum, c + 1 - SW)) * pv1);
    else {
      const double vmc = vm[c];
      double an, ax, bn, bx, temp1, temp2;
      t_minmax(mk, MK_U, um, dpum, dpm, 1, c + 1, an, ax); t_minmax(mk, MK_U, um, dpum, dpm, 1, c + 1 - SW, bn, bx);
      if (pv1 * vmc == 0.) temp1 = pv1 * ((ax + bx) + (an + bn)) * .5;
      else if (pv1 * vmc > 0.) temp1 = pv1 * (ax + bx);
      else temp1 = pv1 * (an + bn);
      t_minmax(mk, MK_U, um, dpum, dpm, 1, c, an, ax); t_minmax(mk, MK_U, um, dpum, dpm, 1, c - SW, bn, bx);
      if (pv0 * vmc == 0.) temp2 = pv0 * ((ax + bx) + (an + bn)) * .5;
      else if (pv0 * vmc > 0.) temp2 = pv0 * (ax + bx);
      else temp2 = pv0 * (an + bn);
      cav = -.25 * (temp1 + temp2);
    }
    double wa, wb;
    wgti_at(g, P, x, k, wa, wb);
    const double vnc = vn[c];
    const double dpvxm = dpvm[c];
    const double dpxy = fmax(dpvxm, onemm);
    double dpia = fmax(dpvm[c - 1], onemm); dpia = dpia + wa * (dpxy - dpia);
    double dpib = fmax(dpvm[c + 1], onemm); dpib = dpib + wb * (dpxy - dpib);
    const double v2 = v2v[c], v4 = v4v[c];
    const bool dry_a = !(mk[c - 1] & MK_V), dry_b = !(mk[c + 1] & MK_V);
    const double v2a = dry_a ? v2 : v2v[c - 1], v4a = dry_a ? v4 : v4v[c - 1];
    const double v2b = dry_b ? v2 : v2v[c + 1], v4b = dry_b ? v4 : v4v[c + 1];
    const double dd = dl2v[c];
    const double dl2via = (1. - wa) * dl2v[c - 1] + wa * SLIP * dd;
    const double dl2vib = (1. - wb) * dl2v[c + 1] + wb * SLIP * dd;
    const double vflux2 = fmin(P.difmxq[x], (v2 + v2a) * P.scqy[x]) * hfharm(dpia, dpxy) * (via[c] - vnc) +
                          fmin(.125 * P.difmxq[x], (v4 + v4a) * P.scqy[x]) * hfharm(dpia, dpxy) * (dl2via - dd);
    const double vflux3 = fmin(P.difmxq[x + 1], (v2 + v2b) * P.scqy[x + 1]) * hfharm(dpib, dpxy) * (vnc - vib[c]) +
                          fmin(.125 * P.difmxq[x + 1], (v4 + v4b) * P.scqy[x + 1]) * hfharm(dpib, dpxy) * (dd - dl2vib);
    const double vflux1c = d2[c], vflux1s = d2[c - SW];
    double stress;
    if (P.isopyc) stress = k == 1 ? -2. * P.tauy[x] * grav * P.scvy[x] / (P.p[x + L] + P.p[x - s + L]) : 0.;
    else stress = -(P.mv_nonloc[xk] - P.mv_nonloc[xk + L]) * P.tauy[x] * grav * P.scvy[x] / fmax(onemm, dpvxm);
    const double pbvm = P.pbv[x + m2];
    const double ptopl = .5 * (fmin(pbvm, P.p[xk]) + fmin(pbvm, P.p[xk - s]));
    const double pbotl = .5 * (fmin(pbvm, P.p[xk + L]) + fmin(pbvm, P.p[xk - s + L]));
    const double q = .5 * (P.drag[x] + P.drag[x - s]) *
                     (fmax(pbvm - thkbop, pbotl) - fmax(pbvm - thkbop, fmin(ptopl, pbotl - onemm))) / fmax(dpvxm, onemm);
    const double botstr = -vnc * q / (1. + P.delt1 * q);
    const double pgf = (1. - 2. * WPGF) * P.pgfy[xm] + WPGF * (P.pgfy_o[xk] + P.pgfy[xn]);
    const double vkm = P.v[xm], vkn = P.v[xn];
    P.sv_m[xk] = vkm * (WUV1 * dpvxm + onemm) + vkn * WUV2 * P.dpvold[xk];
    P.sv_n[xk] = vkn + P.delt1 * (-P.scvyi[x] * (-pgf + stress + (ke[c] - ke[c - SW])) + cav -
                               P.vbcors_p[x] * P.tsfac + botstr -
                               (vflux1c - vflux1s + vflux3 - vflux2) / (P.scv2[x] * fmax(dpvxm, onemm)));
  }
}

// ---- stage 5: column pass (:1154-1267) ------------------------------------------------------------
// The pointers of MtP carry no aliasing information, so a store to u would fence every later load of
// dpu/su_n and the level loop would run as one memory round trip per level (ncu: 25 long-scoreboard
// stall cycles per issue).  col_velocity takes the column's arrays as __restrict__ parameters and the
// operands of four levels are fetched into registers before the four levels are computed, so the
// loads of a batch go out together.
__device__ __forceinline__ void col_velocity(long x, long L, int kdm, int mm, int nn, double delt1, double vmx,
                                             double vbm, double pbp, const double* __restrict__ dpf,
                                             const double* __restrict__ s_n, const double* __restrict__ s_m,
                                             const double* __restrict__ dpold, double* __restrict__ vel,
                                             double* __restrict__ pf, double* __restrict__ totn) {
  constexpr int B = 4;   // levels whose operands are fetched together
  double tot = 0., vprev = 0.;
  for (int k0 = 1; k0 <= kdm; k0 += B) {
    double dm[B], dn[B], sn[B];
#pragma unroll
    for (int b = 0; b < B; ++b) {
      const int k = min(k0 + b, kdm);
      dm[b] = dpf[x + (long)(k + mm - 1) * L]; dn[b] = dpf[x + (long)(k + nn - 1) * L];
      sn[b] = s_n[x + (long)(k - 1) * L];
    }
#pragma unroll
    for (int b = 0; b < B; ++b) {
      const int k = k0 + b;
      if (k <= kdm) {
        const double q = fmin(fmin(dm[b], dn[b]), onem);
        double vn = sn[b];
        const double va = k == 1 ? vn : vprev;  // kan = max(1,k-1)+nn
        vn = (vn * q + va * (onem - q)) / onem;
        vn = fmax(-vmx, fmin(vmx, vn + vbm)) - vbm;
        vel[x + (long)(k + nn - 1) * L] = vn;
        vprev = vn;
        tot = tot + vn * dn[b];
      }
    }
  }
  tot = tot / pbp;
  double pk = pf[x];
  for (int k0 = 1; k0 <= kdm; k0 += B) {
    double dm[B], dn[B], sm[B], dold[B], vo[B];
#pragma unroll
    for (int b = 0; b < B; ++b) {
      const int k = min(k0 + b, kdm);
      dm[b] = dpf[x + (long)(k + mm - 1) * L]; dn[b] = dpf[x + (long)(k + nn - 1) * L];
      sm[b] = s_m[x + (long)(k - 1) * L]; dold[b] = dpold[x + (long)(k - 1) * L];
      vo[b] = vel[x + (long)(k + nn - 1) * L];
    }
#pragma unroll
    for (int b = 0; b < B; ++b) {
      const int k = k0 + b;
      if (k <= kdm) {
        const double vn = vo[b] - tot;
        vel[x + (long)(k + nn - 1) * L] = vn;
        vel[x + (long)(k + mm - 1) * L] = (sm[b] + vn * WUV2 * dn[b]) / (WUV1 * dm[b] + onemm + WUV2 * (dold[b] + dn[b]));
        pk = pk + dn[b];
        pf[x + (long)k * L] = pk;
      }
    }
  }
  totn[x] = tot * (1. / delt1);
}

template <int MINB>
__global__ void __launch_bounds__(128, MINB) mt_column(Geom g, MtP P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.ii) return;
  const long x = ix2(g, i, j), L = g.lev, m2 = (long)(P.m - 1) * L;
  if (P.iu[x] == 1)
    col_velocity(x, L, g.kdm, P.mm, P.nn, P.delt1, P.umax[x], P.ub[x + m2], P.pbu_p[x], P.dpu, P.su_n, P.su_m,
                 P.dpuold, P.u, P.pu, P.utotn);
  if (P.iv[x] == 1)
    col_velocity(x, L, g.kdm, P.mm, P.nn, P.delt1, P.vmax[x], P.vb[x + m2], P.pbv_p[x], P.dpv, P.sv_n, P.sv_m,
                 P.dpvold, P.v, P.pv, P.vtotn);
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
  const bool fused = c.option("momtum_form", "staged") == "fused";
#define S(f) P.f = c.owned("momtum_" #f, g.kdm)
  S(su_m); S(su_n); S(sv_m); S(sv_n);
  if (!fused) {  // layer-sized scratch of the staged form; the fused form keeps these in shared memory
    S(uja); S(ujb); S(via); S(vib); S(dl2u); S(dl2v); S(defor1); S(defor2); S(potvor); S(vsc2u); S(vsc4u); S(vsc2v);
    S(vsc4v); S(ke); S(uflux1); S(vflux1);
  }
#undef S
  P.drag = c.owned("momtum_drag", 1);
  P.ubn = c.owned("momtum_ubn", 1); P.ubm = c.owned("momtum_ubm", 1);
  P.vbn = c.owned("momtum_vbn", 1); P.vbm = c.owned("momtum_vbm", 1);

  { dim3 grid(cdiv(g.ii + 4, 128), g.jj + 4); LAUNCH(mt_pressures, grid, 128, 0, g, P); }
  { dim3 grid(cdiv(g.ii + 1, 128), g.jj + 1); LAUNCH(mt_drag, grid, 128, 0, g, P); }
  halo_update(c.dev("difwgt"), 1, 2, 2, halo_ps);
  if (fused) {
    // momtum_form=fused: one launch per call, all per-layer stages on shared-memory tiles with a 3-point
    // halo skirt.  Bit-identical to the staged form, 3x less HBM traffic, but measured SLOWER on B200
    // (tnx0.25v4: 32.0 ms vs 28.1 ms for the six staged launches; ncu profiles/r01_ncu_mt_level.txt):
    // the routine is issue/latency-bound, not traffic-bound, and the tile form pays 1.6-2x redundant
    // halo work at 24 % occupancy (90 KB of shared memory per 256-thread block).  Kept as an option.
    constexpr int TX = 32, TY = 8;
    using T = MtTile<TX, TY>;
    static bool attr_set = false;
    if (!attr_set) {
      CUDA_CHECK(cudaFuncSetAttribute(mt_level<TX, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::bytes));
      attr_set = true;
    }
    dim3 grid(cdiv(g.ii, TX), cdiv(g.jj, TY), g.kdm);
    LAUNCH_NAMED("mt_level", (mt_level<TX, TY>), grid, T::NT, T::bytes, g, P);
  } else {
    // staged form (default): one launch per stage, layer-sized scratch arrays in HBM
    { const dim3 grid = lgrid(g, dim3(cdiv(g.ii + 4, 128), g.jj + 4, g.kdm)); OCC_DISPATCH3("mt_aux_minblk", 12, 12, 14, 16, LAUNCH_NAMED("mt_aux", mt_aux<OCC>, grid, 128, 0, g, P)); }
    { const dim3 grid = lgrid(g, dim3(cdiv(g.ii + 3, 128), g.jj + 3, g.kdm)); OCC_DISPATCH3("mt_vort_minblk", 12, 7, 12, 16, LAUNCH_NAMED("mt_vort", mt_vort<OCC>, grid, 128, 0, g, P)); }
    { const dim3 grid = lgrid(g, dim3(cdiv(g.ii + 2, 128), g.jj + 2, g.kdm)); LAUNCH(mt_visc, grid, 128, 0, g, P); }
    { const dim3 grid = lgrid(g, dim3(cdiv(g.ii + 1, 128), g.jj + 1, g.kdm)); OCC_DISPATCH3("mt_flux1_minblk", 16, 12, 14, 16, LAUNCH_NAMED("mt_flux1", mt_flux1<OCC>, grid, 128, 0, g, P)); }
    { const dim3 grid = lgrid(g, dim3(cdiv(g.ii, 128), g.jj, g.kdm));
      OCC_DISPATCH3("momtum_minblk", 12, 7, 12, 16,
                    LAUNCH_NAMED("mt_update", mt_update<OCC>, grid, 128, 0, g, P);
                    LAUNCH_NAMED("mt_update_v", mt_update_v<OCC>, grid, 128, 0, g, P)); }
  }
  { dim3 grid(cdiv(g.ii, 128), g.jj);
    OCC_DISPATCH3("mt_column_minblk", 7, 7, 10, 12, LAUNCH_NAMED("mt_column", mt_column<OCC>, grid, 128, 0, g, P)); }
}

}  // namespace blom
