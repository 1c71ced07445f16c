// ORACLE — TEST INFRASTRUCTURE ONLY (see core.hpp header).
// Restatement of phy/mod_ndiff.F90 (neutral diffusion of tracers): peval :62-74, pmeval :76-102,
// drhoroot :104-148, drho :150-158, ndiff_flx :160-953, ndiff_prep_jslice :959-1026,
// ndiff_uflx_jslice :1028-1088, ndiff_vflx_jslice :1090-1150, ndiff_update_trc_jslice :1152-1175.
//
// The reference runs these routines on rotating j-slices inside the ALE regrid-remap pipeline
// (phy/mod_ale_regrid_remap.F90:1614-1690); the slice arrays are that pipeline's products.  Here (and
// in the CUDA library) they are whole-domain arrays in the common (i,j,level) layout:
//   nd_p_src  (kk+1)      source interface pressures            p_src_js(k,i,js)
//   nd_ksmx   int (1)     deepest source layer with mass         ksmx_js(i,js)
//   nd_t_srcdi(2*kk*T)    tracer values at upper/lower interface  t_srcdi_js(is,k,nt,i,js) -> level ((nt-1)*kk+k-1)*2+is
//   nd_tpc_src(5*kk*T)    reconstruction polynomial coefficients  tpc_src_js(c,k,nt,i,js)  -> level ((nt-1)*kk+k-1)*5+c
//   nd_p_dst  (kk+1)      destination interface pressures        p_dst_js(k,i,js)
//   nd_trc_rm (kk*T)      remapped tracers to be updated          trc_rm(k,nt,i)           -> level (nt-1)*kk+k
// with T = 2+ntr scalars (1 temperature, 2 salinity, 3.. passive tracers).  The pipeline order of the
// reference (row j: u faces of row j, then v faces of row j+1) is kept, so every cell accumulates its
// flux convergence in the reference's order.
#include "core.hpp"
#include "eos.hpp"

namespace orc {

namespace {

constexpr double ndiff_dstsnp_fac = .01, rho_eps = 1.e-5, dp_eps = 1.e-5;  // :39-42
constexpr int it = 1, is_ = 2;                                           // :43-45 (`is` of the reference)
constexpr double mval = 1.e30;                                           // :210

// :220-241, :284-304 of phy/mod_eos.F90
inline double eos_drhodt(double p, double th, double s) {
  using namespace eos;
  double r1 = a11 + (a12 + a14 * th + a15 * s) * th + (a13 + a16 * s) * s + (b11 + b12 * th + b13 * s) * p;
  double r2i = 1. / (a21 + (a22 + a24 * th + a25 * s) * th + (a23 + a26 * s) * s + (b21 + b22 * th + b23 * s) * p);
  return (a12 + 2. * a14 * th + a15 * s + b12 * p - (a22 + 2. * a24 * th + a25 * s + b22 * p) * r1 * r2i) * r2i;
}
inline double eos_drhods(double p, double th, double s) {
  using namespace eos;
  double r1 = a11 + (a12 + a14 * th + a15 * s) * th + (a13 + a16 * s) * s + (b11 + b12 * th + b13 * s) * p;
  double r2i = 1. / (a21 + (a22 + a24 * th + a25 * s) * th + (a23 + a26 * s) * s + (b21 + b22 * th + b23 * s) * p);
  return (a13 + a15 * th + 2. * a16 * s + b13 * p - (a23 + a25 * th + 2. * a26 * s + b23 * p) * r1 * r2i) * r2i;
}

// one column of the slice data with the reference's 1-based indices
struct Col {
  const double *p_src, *tsd, *tpc, *drdt, *drds, *p_dst;
  size_t lev; int kk;
  double psd(int s, int k) const { return p_src[(size_t)(k + s - 2) * lev]; }  // p_srcdi(s,k) = p_src(k+s-1)
  double tsrcdi(int s, int k, int nt) const { return tsd[(size_t)(((nt - 1) * kk + k - 1) * 2 + s - 1) * lev]; }
  double tpcc(int c, int k, int nt) const { return tpc[(size_t)(((nt - 1) * kk + k - 1) * 5 + c - 1) * lev]; }
  double drhodt(int s, int k) const { return drdt[(size_t)((k - 1) * 2 + s - 1) * lev]; }
  double drhods(int s, int k) const { return drds[(size_t)((k - 1) * 2 + s - 1) * lev]; }
  double pdst(int k) const { return p_dst[(size_t)(k - 1) * lev]; }
};

// :62-74
inline double peval(const Col& c, int k, int nt, double x) {
  return (((c.tpcc(5, k, nt) * x + c.tpcc(4, k, nt)) * x + c.tpcc(3, k, nt)) * x + c.tpcc(2, k, nt)) * x +
         c.tpcc(1, k, nt);
}
// :76-102
inline double pmeval(const Col& c, int k, int nt, double x0, double x1) {
  const double c1_2 = 1. / 2., c1_3 = 1. / 3., c1_4 = 1. / 4., c1_5 = 1. / 5.;
  double b5 = c1_5 * c.tpcc(5, k, nt);
  double b4 = b5 * x1 + c1_4 * c.tpcc(4, k, nt);
  double b3 = b4 * x1 + c1_3 * c.tpcc(3, k, nt);
  double b2 = b3 * x1 + c1_2 * c.tpcc(2, k, nt);
  double b1 = b2 * x1 + c.tpcc(1, k, nt);
  return (((b5 * x0 + b4) * x0 + b3) * x0 + b2) * x0 + b1;
}
// :104-148: Newton search for the position x in layer k of column c where the density difference
// to the fixed point (tf,sf) vanishes
inline double drhoroot(const Col& c, int k, double tf, double sf, double drhodt_l, double drhodt_u,
                       double drhods_l, double drhods_u) {
  const double eps = 1.e-14, x_tol = 1.e-4;
  double x = .5;
  const double ddrdtdx = drhodt_l - drhodt_u, ddrdsdx = drhods_l - drhods_u;
  auto T = [&](int q) { return c.tpcc(q, k, it); };
  auto S = [&](int q) { return c.tpcc(q, k, is_); };
  for (int n = 1; n <= 10; ++n) {
    double dt = tf - (T(1) + (T(2) + (T(3) + (T(4) + T(5) * x) * x) * x) * x);
    double ds = sf - (S(1) + (S(2) + (S(3) + (S(4) + S(5) * x) * x) * x) * x);
    double drdt = drhodt_l * x + drhodt_u * (1. - x);
    double drds = drhods_l * x + drhods_u * (1. - x);
    double dtdx = -(T(2) + (2. * T(3) + (3. * T(4) + 4. * T(5) * x) * x) * x);
    double dsdx = -(S(2) + (2. * S(3) + (3. * S(4) + 4. * S(5) * x) * x) * x);
    double dr = drdt * dt + drds * ds;
    double ddrdx = ddrdtdx * dt + drdt * dtdx + ddrdsdx * ds + drds * dsdx;
    double x_old = x;
    x = std::max(0., std::min(1., x_old - dr / fsign(std::max(eps, std::fabs(ddrdx)), ddrdx)));
    if (std::fabs(x - x_old) < x_tol) return x;
  }
  return x;
}
// :150-158
inline double drho(double t1, double s1, double t2, double s2, double drhodt, double drhods) {
  return drhodt * (t2 - t1) + drhods * (s2 - s1);
}

// what ndiff_flx needs from the two cells besides the column data
struct CellRef {
  int ksmx, kdmx;
  double dpml;
  const double* difiso;   // level 1 at this cell
  const double* tlev[8];  // tlev[nt-1]: scalar nt at time level base `nn` (level 1+nn), this cell
  double* conv;           // flxconv(kd,nt) of this cell: level (nt-1)*kk + kd, stride lev
};
struct FaceRef {
  const double* puv;      // pu|pv at the face, level 1
  double *tflld, *sflld, *tflx, *sflx, *nslp;  // face arrays, level 1
};

// :160-953
void ndiff_flx(const Col& M, const Col& P, const CellRef& cm, const CellRef& cp, const FaceRef& F, double cdiff,
               double cnslp, int ntr_loc, int mm, bool surface_align) {
  const int kk = M.kk;
  const size_t lev = M.lev;
  const int ksmx_m = cm.ksmx, ksmx_p = cp.ksmx, kdmx_m = cm.kdmx, kdmx_p = cp.kdmx;
  std::vector<double> nslp_src(4 * (kk + 1) + 1), p_nslp_src(4 * (kk + 1) + 1);
  std::vector<double> pnm(2 * (kk + 1) + 2, mval), pnp(2 * (kk + 1) + 2, mval);  // p_ni_srcdi_m/p(is,k)
  auto PNM = [&](int s, int k) -> double& { return pnm[2 * k + s - 1]; };
  auto PNP = [&](int s, int k) -> double& { return pnp[2 * k + s - 1]; };
  std::vector<char> stab_m(kk + 2, 0), stab_p(kk + 2, 0);
  std::vector<double> p_dstsnp_m(kk + 3), p_dstsnp_p(kk + 3);
  std::vector<double> t_ni_m(2 * ntr_loc), t_ni_p(2 * ntr_loc), t_nl_m(ntr_loc), t_nl_p(ntr_loc);
  auto TNM = [&](int nt, int q) -> double& { return t_ni_m[(q - 1) * ntr_loc + nt - 1]; };
  auto TNP = [&](int nt, int q) -> double& { return t_ni_p[(q - 1) * ntr_loc + nt - 1]; };
  double x_ni_m[3], x_ni_p[3], p_ni_m[3], p_ni_p[3];
  double pml = 0, drho_curr = 0, p_ni_m_prev, p_ni_p_prev, drhodt_x0, drhodt_x1, drhods_x0, drhods_x1, x_ni, p_ni,
         drho_prev, p1_m, p2_m, p1_p, p2_p, dp_dst_u, dp_dst_l, pu_m, pl_m, pu_p, pl_p, pp1, pp2, dp_ni_m,
         dp_ni_p, dp_ni, q, dt, ds, tflx, sflx, p_ni_up, p_ni_lo, dp_ni_i, mlfrac, p_nslp_dst;
  int nns, issa_m, issa_p, kssa_m = 0, kssa_p = 0, is_m, is_p, ks_m, ks_p, ks_m_prev, ks_p_prev, kd_m, kd_p,
      isn_m = 1, isn_p = 1, ksn_m = 1, ksn_p = 1, nip, nic, kuv, case_m, case_p, kd, ks;
  bool drho_neg, drho_pos, drho_zero, advance_src_m, advance_src_p, advance_dst_m, advance_dst_p, found_ni;

  auto drho_at = [&]() {
    return drho(M.tsrcdi(is_m, ks_m, it), M.tsrcdi(is_m, ks_m, is_), P.tsrcdi(is_p, ks_p, it), P.tsrcdi(is_p, ks_p, is_),
                .5 * (M.drhodt(is_m, ks_m) + P.drhodt(is_p, ks_p)), .5 * (M.drhods(is_m, ks_m) + P.drhods(is_p, ks_p)));
  };

  // ---- first search: neutral interfaces anchored at source layer interfaces (:212-406)
  nns = 0;
  if (surface_align) {
    pml = .5 * (M.psd(1, 1) + cm.dpml + P.psd(1, 1) + cp.dpml);
    kssa_m = 2;
    while (kssa_m <= ksmx_m) {
      if (M.psd(1, kssa_m) > pml) break;
      kssa_m = kssa_m + 1;
    }
    kssa_p = 2;
    while (kssa_p <= ksmx_p) {
      if (P.psd(1, kssa_p) > pml) break;
      kssa_p = kssa_p + 1;
    }
    is_m = 1; ks_m = kssa_m; is_p = 1; ks_p = kssa_p;
    p_ni_m_prev = pml; p_ni_p_prev = pml;
  } else {
    is_m = 1; ks_m = 1; is_p = 1; ks_p = 1;
    p_ni_m_prev = M.psd(1, 1); p_ni_p_prev = P.psd(1, 1);
  }
  if (ks_m <= ksmx_m && ks_p <= ksmx_p) drho_curr = drho_at();

  [&]() {  // search_loop1
    while (ks_m <= ksmx_m && ks_p <= ksmx_p) {
      drho_neg = drho_curr <= -rho_eps;
      drho_pos = drho_curr >= rho_eps;
      drho_zero = !(drho_neg || drho_pos);
      if (is_m + ks_m > 2 && is_p + ks_p > 2) {
        if (drho_neg) {
          if (is_m == 2) {
            drhodt_x0 = .5 * (M.drhodt(1, ks_m) + P.drhodt(is_p, ks_p));
            drhodt_x1 = .5 * (M.drhodt(2, ks_m) + P.drhodt(is_p, ks_p));
            drhods_x0 = .5 * (M.drhods(1, ks_m) + P.drhods(is_p, ks_p));
            drhods_x1 = .5 * (M.drhods(2, ks_m) + P.drhods(is_p, ks_p));
            x_ni = drhoroot(M, ks_m, P.tsrcdi(is_p, ks_p, it), P.tsrcdi(is_p, ks_p, is_), drhodt_x1, drhodt_x0,
                            drhods_x1, drhods_x0);
            p_ni = M.psd(2, ks_m) * x_ni + M.psd(1, ks_m) * (1. - x_ni);
            if (p_ni > p_ni_m_prev) {
              p_ni_m_prev = p_ni;
              PNP(is_p, ks_p) = p_ni;
              nns = nns + 1;
              nslp_src[nns] = -cnslp * (P.psd(is_p, ks_p) - p_ni);
              p_nslp_src[nns] = .5 * (P.psd(is_p, ks_p) + p_ni);
            }
          }
        } else if (drho_pos) {
          if (is_p == 2) {
            drhodt_x0 = .5 * (M.drhodt(is_m, ks_m) + P.drhodt(1, ks_p));
            drhodt_x1 = .5 * (M.drhodt(is_m, ks_m) + P.drhodt(2, ks_p));
            drhods_x0 = .5 * (M.drhods(is_m, ks_m) + P.drhods(1, ks_p));
            drhods_x1 = .5 * (M.drhods(is_m, ks_m) + P.drhods(2, ks_p));
            x_ni = drhoroot(P, ks_p, M.tsrcdi(is_m, ks_m, it), M.tsrcdi(is_m, ks_m, is_), drhodt_x1, drhodt_x0,
                            drhods_x1, drhods_x0);
            p_ni = P.psd(2, ks_p) * x_ni + P.psd(1, ks_p) * (1. - x_ni);
            if (p_ni > p_ni_p_prev) {
              p_ni_p_prev = p_ni;
              PNM(is_m, ks_m) = p_ni;
              nns = nns + 1;
              nslp_src[nns] = -cnslp * (p_ni - M.psd(is_m, ks_m));
              p_nslp_src[nns] = .5 * (p_ni + M.psd(is_m, ks_m));
            }
          }
        } else {
          PNP(is_p, ks_p) = M.psd(is_m, ks_m);
          PNM(is_m, ks_m) = P.psd(is_p, ks_p);
          nns = nns + 1;
          nslp_src[nns] = -cnslp * (P.psd(is_p, ks_p) - M.psd(is_m, ks_m));
          p_nslp_src[nns] = .5 * (P.psd(is_p, ks_p) + M.psd(is_m, ks_m));
        }
      }
      if (drho_zero || drho_pos) {
        for (;;) {
          drho_prev = drho_curr;
          if (is_m == 1) is_m = 2;
          else {
            ks_m = ks_m + 1;
            if (ks_m > ksmx_m) return;
            is_m = 1;
          }
          drho_curr = drho_at();
          if (drho_prev - drho_curr > rho_eps) {
            if (is_m == 2 && M.psd(2, ks_m) - M.psd(1, ks_m) > onemm) stab_m[ks_m] = 1;
            break;
          }
          if (is_m == 1) PNM(is_m, ks_m) = PNM(2, ks_m - 1);
        }
      }
      if (drho_zero || drho_neg) {
        for (;;) {
          drho_prev = drho_curr;
          if (is_p == 1) is_p = 2;
          else {
            ks_p = ks_p + 1;
            if (ks_p > ksmx_p) return;
            is_p = 1;
          }
          drho_curr = drho_at();
          if (drho_curr - drho_prev > rho_eps) {
            if (is_p == 2 && P.psd(2, ks_p) - P.psd(1, ks_p) > onemm) stab_p[ks_p] = 1;
            break;
          }
          if (is_p == 1) PNP(is_p, ks_p) = PNP(2, ks_p - 1);
        }
      }
    }
  }();

  if (surface_align) {  // :408-479
    issa_m = 1;
    while (kssa_m <= ksmx_m) {
      if (PNM(issa_m, kssa_m) != mval) break;
      if (issa_m == 1) issa_m = 2;
      else { kssa_m = kssa_m + 1; issa_m = 1; }
    }
    issa_p = 1;
    while (kssa_p <= ksmx_p) {
      if (PNP(issa_p, kssa_p) != mval) break;
      if (issa_p == 1) issa_p = 2;
      else { kssa_p = kssa_p + 1; issa_p = 1; }
    }
    if (kssa_m > ksmx_m || kssa_p > ksmx_p) {
      PNM(1, 1) = M.psd(1, 1);
      for (ks_m = 1; ks_m <= ksmx_m - 1; ++ks_m) {
        if (M.psd(1, ks_m) > P.psd(2, ksmx_p)) break;
        p_ni = std::min(M.psd(2, ks_m), P.psd(2, ksmx_p));
        PNM(1, ks_m + 1) = p_ni;
        PNM(2, ks_m) = p_ni;
        stab_m[ks_m] = 1;
      }
      PNP(1, 1) = P.psd(1, 1);
      for (ks_p = 1; ks_p <= ksmx_p - 1; ++ks_p) {
        if (P.psd(1, ks_p) > M.psd(2, ksmx_m)) break;
        p_ni = std::min(P.psd(2, ks_p), M.psd(2, ksmx_m));
        PNP(1, ks_p + 1) = p_ni;
        PNP(2, ks_p) = p_ni;
        stab_p[ks_p] = 1;
      }
    } else {
      if (M.psd(issa_m, kssa_m) < PNP(issa_p, kssa_p)) {
        p1_m = M.psd(1, 1); p2_m = M.psd(issa_m, kssa_m);
        p1_p = P.psd(1, 1); p2_p = PNM(issa_m, kssa_m);
      } else {
        p1_m = M.psd(1, 1); p2_m = PNP(issa_p, kssa_p);
        p1_p = P.psd(1, 1); p2_p = P.psd(issa_p, kssa_p);
      }
      PNM(1, 1) = p1_p;
      for (ks_m = 1; ks_m <= kssa_m - 1; ++ks_m) {
        p_ni = ((M.psd(2, ks_m) - p1_m) * p2_p + (p2_m - M.psd(2, ks_m)) * p1_p) / (p2_m - p1_m);
        PNM(1, ks_m + 1) = p_ni;
        PNM(2, ks_m) = p_ni;
        stab_m[ks_m] = 1;
      }
      PNP(1, 1) = p1_m;
      for (ks_p = 1; ks_p <= kssa_p - 1; ++ks_p) {
        p_ni = ((P.psd(2, ks_p) - p1_p) * p2_m + (p2_p - P.psd(2, ks_p)) * p1_m) / (p2_p - p1_p);
        PNP(1, ks_p + 1) = p_ni;
        PNP(2, ks_p) = p_ni;
        stab_p[ks_p] = 1;
      }
    }
  }

  // ---- destination interfaces snapped to nearby source interfaces (:491-523)
  p_dstsnp_m[1] = M.pdst(1);
  dp_dst_u = M.pdst(2) - M.pdst(1);
  for (int k = 2; k <= std::min(ksmx_m, kdmx_m); ++k) {
    dp_dst_l = M.pdst(k + 1) - M.pdst(k);
    if (std::fabs(M.pdst(k) - M.psd(1, k)) < std::min(dp_dst_u, dp_dst_l) * ndiff_dstsnp_fac) p_dstsnp_m[k] = M.psd(1, k);
    else p_dstsnp_m[k] = M.pdst(k);
    dp_dst_u = dp_dst_l;
  }
  for (int k = std::min(ksmx_m, kdmx_m) + 1; k <= kdmx_m + 1; ++k) p_dstsnp_m[k] = M.pdst(k);
  p_dstsnp_p[1] = P.pdst(1);
  dp_dst_u = P.pdst(2) - P.pdst(1);
  for (int k = 2; k <= std::min(ksmx_p, kdmx_p); ++k) {
    dp_dst_l = P.pdst(k + 1) - P.pdst(k);
    if (std::fabs(P.pdst(k) - P.psd(1, k)) < std::min(dp_dst_u, dp_dst_l) * ndiff_dstsnp_fac) p_dstsnp_p[k] = P.psd(1, k);
    else p_dstsnp_p[k] = P.pdst(k);
    dp_dst_u = dp_dst_l;
  }
  for (int k = std::min(ksmx_p, kdmx_p) + 1; k <= kdmx_p + 1; ++k) p_dstsnp_p[k] = P.pdst(k);

  // ---- second search: neutral layers and their fluxes (:525-911)
  is_m = 2; ks_m = 0; is_p = 2; ks_p = 0; kd_m = 0; kd_p = 0;
  advance_src_m = true; advance_src_p = true; advance_dst_m = true; advance_dst_p = true;
  ks_m_prev = 0; ks_p_prev = 0;
  nip = 1; nic = 2;
  p_ni_m[nip] = -mval; p_ni_p[nip] = -mval;
  kuv = 1;
  auto puv = [&](int k) { return F.puv[(size_t)(k - 1) * lev]; };

  [&]() {  // search_loop2
    for (;;) {
      if (advance_src_m) {
        for (;;) {
          if (is_m == 1) {
            is_m = 2;
            if (stab_m[ks_m]) break;
          } else {
            ks_m = ks_m + 1;
            if (ks_m > ksmx_m) return;
            is_m = 1;
            if (stab_m[ks_m] && PNM(is_m, ks_m) != mval) break;
          }
        }
        isn_m = is_m; ksn_m = ks_m;
        while (PNM(isn_m, ksn_m) == mval) {
          if (isn_m == 1) isn_m = 2;
          else {
            if (ksn_m == ksmx_m) break;
            ksn_m = ksn_m + 1;
            isn_m = 1;
          }
        }
      }
      if (advance_src_p) {
        for (;;) {
          if (is_p == 1) {
            is_p = 2;
            if (stab_p[ks_p]) break;
          } else {
            ks_p = ks_p + 1;
            if (ks_p > ksmx_p) return;
            is_p = 1;
            if (stab_p[ks_p] && PNP(is_p, ks_p) != mval) break;
          }
        }
        isn_p = is_p; ksn_p = ks_p;
        while (PNP(isn_p, ksn_p) == mval) {
          if (isn_p == 1) isn_p = 2;
          else {
            if (ksn_p == ksmx_p) break;
            ksn_p = ksn_p + 1;
            isn_p = 1;
          }
        }
      }
      if (p_ni_m[nip] == -mval) {
        if ((PNM(isn_m, ksn_m) - P.psd(isn_p, ksn_p)) < (PNP(isn_p, ksn_p) - M.psd(isn_m, ksn_m))) {
          p_ni_m[nip] = M.psd(isn_m, ksn_m);
          p_ni_p[nip] = PNM(isn_m, ksn_m);
        } else {
          p_ni_m[nip] = PNP(isn_p, ksn_p);
          p_ni_p[nip] = P.psd(isn_p, ksn_p);
        }
      }
      if (advance_dst_m) {
        kd_m = kd_m + 1;
        if (kd_m > kdmx_m) return;
      }
      if (advance_dst_p) {
        kd_p = kd_p + 1;
        if (kd_p > kdmx_p) return;
      }
      while (p_dstsnp_m[kd_m + 1] <= std::max(M.psd(1, ks_m), p_ni_m[nip])) {
        kd_m = kd_m + 1;
        if (kd_m > kdmx_m) return;
      }
      while (p_dstsnp_p[kd_p + 1] <= std::max(P.psd(1, ks_p), p_ni_p[nip])) {
        kd_p = kd_p + 1;
        if (kd_p > kdmx_p) return;
      }
      advance_src_m = false; advance_src_p = false; advance_dst_m = false; advance_dst_p = false;

      case_m = 3;
      if (M.psd(is_m, ks_m) <= PNP(isn_p, ksn_p)) {
        if (M.psd(is_m, ks_m) <= p_dstsnp_m[kd_m + 1]) case_m = 1;
      } else if (PNP(isn_p, ksn_p) <= p_dstsnp_m[kd_m + 1]) {
        case_m = 2;
      }
      case_p = 3;
      if (P.psd(is_p, ks_p) <= PNM(isn_m, ksn_m)) {
        if (P.psd(is_p, ks_p) <= p_dstsnp_p[kd_p + 1]) case_p = 1;
      } else if (PNM(isn_m, ksn_m) <= p_dstsnp_p[kd_p + 1]) {
        case_p = 2;
      }
      found_ni = false;
      auto eval_both = [&]() {
        for (int nt = 1; nt <= ntr_loc; ++nt) {
          TNM(nt, nic) = peval(M, ks_m, nt, x_ni_m[nic]);
          TNP(nt, nic) = peval(P, ks_p, nt, x_ni_p[nic]);
        }
      };

      if (case_m == 3 && case_p == 3) {
        if (is_p == 2 && is_m == 2) {
          p_ni_m[nic] = p_dstsnp_m[kd_m + 1];
          p_ni_p[nic] = p_dstsnp_p[kd_p + 1];
          pu_m = p_ni_m[nip];
          pu_p = p_ni_p[nip];
          if ((PNM(isn_m, ksn_m) - P.psd(isn_p, ksn_p)) < (PNP(isn_p, ksn_p) - M.psd(isn_m, ksn_m))) {
            pl_m = M.psd(isn_m, ksn_m);
            pl_p = PNM(isn_m, ksn_m);
          } else {
            pl_m = PNP(isn_p, ksn_p);
            pl_p = P.psd(isn_p, ksn_p);
          }
          pp1 = (p_ni_m[nic] - pu_m) * (pl_p - pu_p);
          pp2 = (p_ni_p[nic] - pu_p) * (pl_m - pu_m);
          if (std::fabs(pp1 - pp2) < dp_eps * std::max(dp_eps, pl_m - pu_m + pl_p - pu_p)) {
            advance_dst_m = true;
            advance_dst_p = true;
          } else if (pp1 < pp2) {
            p_ni_p[nic] = pu_p + pp1 / (pl_m - pu_m);
            advance_dst_m = true;
          } else {
            p_ni_m[nic] = pu_m + pp2 / (pl_p - pu_p);
            advance_dst_p = true;
          }
          if (p_ni_m[nic] >= M.psd(1, ks_m) && p_ni_m[nic] <= M.psd(2, ks_m) && p_ni_p[nic] >= P.psd(1, ks_p) &&
              p_ni_p[nic] <= P.psd(2, ks_p)) {
            x_ni_m[nic] = (p_ni_m[nic] - M.psd(1, ks_m)) / (M.psd(2, ks_m) - M.psd(1, ks_m));
            x_ni_p[nic] = (p_ni_p[nic] - P.psd(1, ks_p)) / (P.psd(2, ks_p) - P.psd(1, ks_p));
            eval_both();
            found_ni = true;
          }
        } else {
          if (is_p != 2) advance_dst_m = true;
          if (is_m != 2) advance_dst_p = true;
        }
      } else if (case_m == 3) {
        if (is_p == 2) {
          p_ni_m[nic] = p_dstsnp_m[kd_m + 1];
          if (case_p == 1)
            p_ni_p[nic] = p_ni_p[nip] + (p_ni_m[nic] - p_ni_m[nip]) * (P.psd(isn_p, ksn_p) - p_ni_p[nip]) /
                                            (PNP(isn_p, ksn_p) - p_ni_m[nip]);
          else
            p_ni_p[nic] = p_ni_p[nip] + (p_ni_m[nic] - p_ni_m[nip]) * (PNM(isn_m, ksn_m) - p_ni_p[nip]) /
                                            (M.psd(isn_m, ksn_m) - p_ni_m[nip]);
          if (p_ni_p[nic] >= P.psd(1, ks_p) && p_ni_p[nic] <= P.psd(2, ks_p)) {
            x_ni_m[nic] = (p_dstsnp_m[kd_m + 1] - M.psd(1, ks_m)) / (M.psd(2, ks_m) - M.psd(1, ks_m));
            x_ni_p[nic] = (p_ni_p[nic] - P.psd(1, ks_p)) / (P.psd(2, ks_p) - P.psd(1, ks_p));
            eval_both();
            found_ni = true;
            advance_dst_m = true;
          } else {
            if (case_p == 1 && PNP(is_p, ks_p) == mval) advance_src_p = true;
            else advance_dst_m = true;
          }
        } else {
          advance_dst_m = true;
        }
      } else if (case_p == 3) {
        if (is_m == 2) {
          p_ni_p[nic] = p_dstsnp_p[kd_p + 1];
          if (case_m == 1)
            p_ni_m[nic] = p_ni_m[nip] + (p_ni_p[nic] - p_ni_p[nip]) * (M.psd(isn_m, ksn_m) - p_ni_m[nip]) /
                                            (PNM(isn_m, ksn_m) - p_ni_p[nip]);
          else
            p_ni_m[nic] = p_ni_m[nip] + (p_ni_p[nic] - p_ni_p[nip]) * (PNP(isn_p, ksn_p) - p_ni_m[nip]) /
                                            (P.psd(isn_p, ksn_p) - p_ni_p[nip]);
          if (p_ni_m[nic] >= M.psd(1, ks_m) && p_ni_m[nic] <= M.psd(2, ks_m)) {
            x_ni_p[nic] = (p_dstsnp_p[kd_p + 1] - P.psd(1, ks_p)) / (P.psd(2, ks_p) - P.psd(1, ks_p));
            x_ni_m[nic] = (p_ni_m[nic] - M.psd(1, ks_m)) / (M.psd(2, ks_m) - M.psd(1, ks_m));
            eval_both();
            found_ni = true;
            advance_dst_p = true;
          } else {
            if (case_m == 1 && PNM(is_m, ks_m) == mval) advance_src_m = true;
            else advance_dst_p = true;
          }
        } else {
          advance_dst_p = true;
        }
      } else if (case_m == 1 && case_p == 1) {
        if (PNM(is_m, ks_m) != mval && PNP(is_p, ks_p) != mval) {
          x_ni_m[nic] = (double)(is_m - 1);
          p_ni_m[nic] = M.psd(is_m, ks_m);
          x_ni_p[nic] = (double)(is_p - 1);
          p_ni_p[nic] = P.psd(is_p, ks_p);
          for (int nt = 1; nt <= ntr_loc; ++nt) {
            TNM(nt, nic) = M.tsrcdi(is_m, ks_m, nt);
            TNP(nt, nic) = P.tsrcdi(is_p, ks_p, nt);
          }
          found_ni = true;
          advance_src_m = true;
          advance_src_p = true;
        } else {
          if (PNM(is_m, ks_m) == mval) advance_src_m = true;
          if (PNP(is_p, ks_p) == mval) advance_src_p = true;
        }
      } else if (case_m == 1) {
        if (PNM(is_m, ks_m) != mval && PNM(is_m, ks_m) >= P.psd(1, ks_p)) {
          x_ni_m[nic] = (double)(is_m - 1);
          p_ni_m[nic] = M.psd(is_m, ks_m);
          p_ni_p[nic] = PNM(is_m, ks_m);
          x_ni_p[nic] = (p_ni_p[nic] - P.psd(1, ks_p)) / (P.psd(2, ks_p) - P.psd(1, ks_p));
          for (int nt = 1; nt <= ntr_loc; ++nt) {
            TNM(nt, nic) = M.tsrcdi(is_m, ks_m, nt);
            TNP(nt, nic) = peval(P, ks_p, nt, x_ni_p[nic]);
          }
          found_ni = true;
        }
        advance_src_m = true;
      } else if (case_p == 1) {
        if (PNP(is_p, ks_p) != mval && PNP(is_p, ks_p) >= M.psd(1, ks_m)) {
          x_ni_p[nic] = (double)(is_p - 1);
          p_ni_p[nic] = P.psd(is_p, ks_p);
          p_ni_m[nic] = PNP(is_p, ks_p);
          x_ni_m[nic] = (p_ni_m[nic] - M.psd(1, ks_m)) / (M.psd(2, ks_m) - M.psd(1, ks_m));
          for (int nt = 1; nt <= ntr_loc; ++nt) {
            TNP(nt, nic) = P.tsrcdi(is_p, ks_p, nt);
            TNM(nt, nic) = peval(M, ks_m, nt, x_ni_m[nic]);
          }
          found_ni = true;
        }
        advance_src_p = true;
      } else {
        advance_src_m = true;
        advance_src_p = true;
      }

      if (found_ni) {  // :795-907
        dp_ni_m = std::min(p_ni_m[nic] - p_ni_m[nip], M.pdst(kd_m + 1) - M.pdst(kd_m));
        dp_ni_p = std::min(p_ni_p[nic] - p_ni_p[nip], P.pdst(kd_p + 1) - P.pdst(kd_p));
        dp_ni = 2. * dp_ni_m * dp_ni_p / std::max(dp_ni_m + dp_ni_p, 2. * dp_eps);
        if (ks_m == ks_m_prev && ks_p == ks_p_prev && p_ni_m[nip] >= p_dstsnp_m[kd_m] &&
            p_ni_m[nic] <= p_dstsnp_m[kd_m + 1] && p_ni_p[nip] >= p_dstsnp_p[kd_p] &&
            p_ni_p[nic] <= p_dstsnp_p[kd_p + 1] && dp_ni > 2. * dp_eps) {
          for (int nt = 1; nt <= ntr_loc; ++nt) {
            t_nl_m[nt - 1] = pmeval(M, ks_m, nt, x_ni_m[nip], x_ni_m[nic]);
            t_nl_p[nt - 1] = pmeval(P, ks_p, nt, x_ni_p[nip], x_ni_p[nic]);
          }
          q = .5 * cdiff * (cm.difiso[(size_t)(ks_m - 1) * lev] + cp.difiso[(size_t)(ks_p - 1) * lev]) * dp_ni;
          dt = t_nl_m[it - 1] - t_nl_p[it - 1];
          ds = t_nl_m[is_ - 1] - t_nl_p[is_ - 1];
          auto cellv = [&](const CellRef& c, int nt, int ksl) { return c.tlev[nt - 1][(size_t)(ksl - 1) * lev]; };
          if (dt * (cellv(cm, it, ks_m) - cellv(cp, it, ks_p)) >= 0. && dt * (TNM(it, nip) - TNP(it, nip)) >= 0. &&
              dt * (TNM(it, nic) - TNP(it, nic)) >= 0. && ds * (cellv(cm, is_, ks_m) - cellv(cp, is_, ks_p)) >= 0. &&
              ds * (TNM(is_, nip) - TNP(is_, nip)) >= 0. && ds * (TNM(is_, nic) - TNP(is_, nic)) >= 0.) {
            tflx = q * dt;
            cm.conv[(size_t)((it - 1) * kk + kd_m - 1) * lev] += tflx;
            cp.conv[(size_t)((it - 1) * kk + kd_p - 1) * lev] -= tflx;
            sflx = q * ds;
            cm.conv[(size_t)((is_ - 1) * kk + kd_m - 1) * lev] += sflx;
            cp.conv[(size_t)((is_ - 1) * kk + kd_p - 1) * lev] -= sflx;
            p_ni_up = .5 * (p_ni_m[nip] + p_ni_p[nip]);
            p_ni_lo = .5 * (p_ni_m[nic] + p_ni_p[nic]);
            dp_ni_i = 1. / std::max(epsilp, p_ni_lo - p_ni_up);
            while (kuv <= kk) {
              const size_t o = (size_t)(kuv + mm - 1) * lev;
              if (puv(kuv + 1) < p_ni_lo) {
                mlfrac = std::max(0., puv(kuv + 1) - std::max(p_ni_up, puv(kuv))) * dp_ni_i;
                F.tflld[o] = F.tflld[o] + tflx * mlfrac;
                F.sflld[o] = F.sflld[o] + sflx * mlfrac;
                F.tflx[o] = F.tflx[o] + tflx * mlfrac;
                F.sflx[o] = F.sflx[o] + sflx * mlfrac;
                kuv = kuv + 1;
              } else {
                mlfrac = (p_ni_lo - std::max(p_ni_up, puv(kuv))) * dp_ni_i;
                F.tflld[o] = F.tflld[o] + tflx * mlfrac;
                F.sflld[o] = F.sflld[o] + sflx * mlfrac;
                F.tflx[o] = F.tflx[o] + tflx * mlfrac;
                F.sflx[o] = F.sflx[o] + sflx * mlfrac;
                break;
              }
            }
          }
          for (int nt = 3; nt <= ntr_loc; ++nt) {
            dt = t_nl_m[nt - 1] - t_nl_p[nt - 1];
            if (dt * (cellv(cm, nt, ks_m) - cellv(cp, nt, ks_p)) >= 0. && dt * (TNM(nt, nip) - TNP(nt, nip)) >= 0. &&
                dt * (TNM(nt, nic) - TNP(nt, nic)) >= 0.) {
              tflx = q * dt;
              cm.conv[(size_t)((nt - 1) * kk + kd_m - 1) * lev] += tflx;
              cp.conv[(size_t)((nt - 1) * kk + kd_p - 1) * lev] -= tflx;
            }
          }
        }
        ks_m_prev = ks_m;
        ks_p_prev = ks_p;
        nip = 3 - nip;
        nic = 3 - nic;
      }
    }
  }();

  // ---- neutral slope at the destination interfaces (:913-951)
  auto nslp = [&](int k) -> double& { return F.nslp[(size_t)(k - 1) * lev]; };
  if (nns == 0) {
    for (int k = 1; k <= kk; ++k) nslp(k) = 0.;
  } else {
    p_nslp_dst = 0.;
    for (kd = 1; kd <= kk; ++kd) {
      p_nslp_dst = .5 * (M.pdst(kd) + P.pdst(kd));
      if (p_nslp_dst > p_nslp_src[1]) break;
      nslp(kd) = nslp_src[1];
    }
    if (kd <= kk) {
      ks = 1;
      bool done = false;
      for (;;) {
        while (p_nslp_dst > p_nslp_src[ks]) {
          if (ks == nns) { done = true; break; }
          ks = ks + 1;
        }
        if (done) break;
        q = (p_nslp_src[ks] - p_nslp_dst) / std::max(p_nslp_src[ks] - p_nslp_src[ks - 1], epsilp);
        nslp(kd) = q * nslp_src[ks - 1] + (1. - q) * nslp_src[ks];
        kd = kd + 1;
        if (kd > kk) break;
        p_nslp_dst = .5 * (M.pdst(kd) + P.pdst(kd));
      }
      for (; kd <= kk; ++kd) nslp(kd) = nslp_src[nns];
    }
  }
}

}  // namespace

// Whole-domain driver in the reference's pipeline order (phy/mod_ale_regrid_remap.F90:1614-1690 with
// jofs2 = 1): prep on 0..ii+1 x 0..jj+1, then per row the u faces (i = 1..ii+1), the v faces of the
// next row (i = 1..ii) and the tracer update (i = 1..ii).
void ndiff(int m, int n, int mm, int nn, int k1m, int k1n) {
  (void)m; (void)n; (void)k1m; (void)k1n;
  Oracle& o = O(); const Dims& d = o.d;
  const int ii = d.ii, jj = d.jj, kk = d.kk, T = 2 + d.ntr;
  const size_t lev = d.lev;
  const double delt1 = o.scalar("delt1");
  const bool surface_align = o.option("ndiff_surface_align", "1") == "1";  // namelist default .true.
  I2 ip = o.i2("ip"), iu = o.i2("iu"), iv = o.i2("iv"), ksmx = o.i2("nd_ksmx");
  A3 p_src = o.a3("nd_p_src"), tsd = o.a3("nd_t_srcdi"), tpc = o.a3("nd_tpc_src"), p_dst = o.a3("nd_p_dst"),
     trc_rm = o.a3("nd_trc_rm");
  A3 temp = o.a3("temp"), saln = o.a3("saln"), difiso = o.a3("difiso"), pu = o.a3("pu"), pv = o.a3("pv");
  A3 utflld = o.a3("utflld"), usflld = o.a3("usflld"), vtflld = o.a3("vtflld"), vsflld = o.a3("vsflld");
  A3 utflx = o.a3("utflx"), usflx = o.a3("usflx"), vtflx = o.a3("vtflx"), vsflx = o.a3("vsflx");
  A3 nslpx = o.a3("nslpx"), nslpy = o.a3("nslpy");
  A2 dpml = o.a2("dpml"), scp2 = o.a2("scp2"), scuy = o.a2("scuy"), scuxi = o.a2("scuxi"), scvx = o.a2("scvx"),
     scvyi = o.a2("scvyi");
  A3 trc = d.ntr > 0 ? o.a3("trc") : A3{};
  A3 drdt = o.scratch("_nd_drhodt", 2 * kk), drds = o.scratch("_nd_drhods", 2 * kk), conv = o.scratch("_nd_flxconv", kk * T);
  I2 kdmx = o.iscratch("_nd_kdmx");
  if (surface_align) xctilr(dpml, 1, 1, halo_ps);   // mod_ale_regrid_remap.F90:1607

  auto off = [&](int i, int j) { return (size_t)(j + d.nbdy - 1) * d.ldi + (i + d.nbdy - 1); };
  // ndiff_prep_jslice (:959-1026)
  for (int j = 0; j <= jj + 1; ++j)
    for (int i = 0; i <= ii + 1; ++i) if (ip(i, j) == 1) {
      kdmx(i, j) = kk;
      for (int k = kk; k >= 1; --k)
        if (p_dst(i, j, k) == p_dst(i, j, kk + 1)) kdmx(i, j) = k - 1;
      for (int k = 1; k <= ksmx(i, j); ++k)
        for (int s = 1; s <= 2; ++s) {
          const double ps = p_src(i, j, k + s - 1);
          const double t = tsd(i, j, ((it - 1) * kk + k - 1) * 2 + s), sa = tsd(i, j, ((is_ - 1) * kk + k - 1) * 2 + s);
          drdt(i, j, (k - 1) * 2 + s) = eos_drhodt(ps, t, sa);
          drds(i, j, (k - 1) * 2 + s) = eos_drhods(ps, t, sa);
        }
      for (int q = 1; q <= kk * T; ++q) conv(i, j, q) = 0.;
    }
  for (int j = 0; j <= jj + 1; ++j)
    for (int k = 1; k <= kk; ++k) {
      const int km = k + mm;
      for (int i = 0; i <= ii + 1; ++i) {
        if (iu(i, j) == 1) { utflld(i, j, km) = 0.; usflld(i, j, km) = 0.; }
        if (iv(i, j) == 1) { vtflld(i, j, km) = 0.; vsflld(i, j, km) = 0.; }
      }
    }

  auto col = [&](int i, int j) {
    const size_t x = off(i, j);
    return Col{p_src.p + x, tsd.p + x, tpc.p + x, drdt.p + x, drds.p + x, p_dst.p + x, lev, kk};
  };
  auto cell = [&](int i, int j) {
    const size_t x = off(i, j);
    CellRef c{};
    c.ksmx = ksmx(i, j); c.kdmx = kdmx(i, j); c.dpml = dpml(i, j);
    c.difiso = difiso.p + x;
    c.tlev[0] = temp.p + x + (size_t)nn * lev;
    c.tlev[1] = saln.p + x + (size_t)nn * lev;
    for (int nt = 3; nt <= T; ++nt) c.tlev[nt - 1] = trc.p + x + (size_t)(nn + (nt - 3) * 2 * d.kdm) * lev;
    c.conv = conv.p + x;
    return c;
  };
  auto face = [&](int i, int j, A3 puv, A3 tflld, A3 sflld, A3 tflx, A3 sflx, A3 nslp) {
    const size_t x = off(i, j);
    return FaceRef{puv.p + x, tflld.p + x, sflld.p + x, tflx.p + x, sflx.p + x, nslp.p + x};
  };
  auto uflx_row = [&](int j) {  // :1028-1088
    for (int i = 1; i <= ii + 1; ++i) if (iu(i, j) == 1)
      ndiff_flx(col(i - 1, j), col(i, j), cell(i - 1, j), cell(i, j), face(i, j, pu, utflld, usflld, utflx, usflx, nslpx),
                delt1 * scuy(i, j) * scuxi(i, j), alpha0 * scuxi(i, j) / grav, T, mm, surface_align);
  };
  auto vflx_row = [&](int j) {  // :1090-1150
    for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1)
      ndiff_flx(col(i, j - 1), col(i, j), cell(i, j - 1), cell(i, j), face(i, j, pv, vtflld, vsflld, vtflx, vsflx, nslpy),
                delt1 * scvx(i, j) * scvyi(i, j), alpha0 * scvyi(i, j) / grav, T, mm, surface_align);
  };
  for (int j = 0; j <= jj; ++j) {
    if (j >= 1) uflx_row(j);
    vflx_row(j + 1);
    if (j >= 1)  // ndiff_update_trc_jslice (:1152-1175)
      for (int i = 1; i <= ii; ++i) if (ip(i, j) == 1)
        for (int k = 1; k <= kk; ++k) {
          const double q = 1. / (scp2(i, j) * std::max(p_dst(i, j, k + 1) - p_dst(i, j, k), dp_eps));
          for (int nt = 1; nt <= T; ++nt)
            trc_rm(i, j, (nt - 1) * kk + k) = trc_rm(i, j, (nt - 1) * kk + k) - q * conv(i, j, (nt - 1) * kk + k);
        }
  }
}

}  // namespace orc
