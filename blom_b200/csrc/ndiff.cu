// Neutral diffusion of tracers (phy/mod_ndiff.F90): ndiff_prep_jslice :959-1026, ndiff_flx :160-953
// (with peval :62-74, pmeval :76-102, drhoroot :104-148, drho :150-158), ndiff_uflx_jslice :1028-1088,
// ndiff_vflx_jslice :1090-1150 and ndiff_update_trc_jslice :1152-1175.
//
// The reference walks rotating j-slices inside the ALE regrid-remap pipeline
// (phy/mod_ale_regrid_remap.F90:1614-1690); its slice arrays are whole-domain arrays here, in the common
// (i,j,level) layout (names and level order: include/blomgpu.h, "neutral diffusion inputs").
//
// B200 design: the search for neutral sublayers between two columns is sequential and data dependent,
// so ONE THREAD OWNS ONE FACE COLUMN with i across lanes (every level access of a warp is a row
// segment).  Three launches instead of the reference's slice pipeline:
//   ndiff_prep    per cell column: kdmx, drhodt/drhods at the source interfaces, the snapped destination
//                 interfaces (a pure function of the cell, so it is evaluated once per cell instead of
//                 once per face as in the reference), zero of the face accumulators
//   ndiff_face<u|v>  per face column: both searches, fluxes, layer binning of the face fluxes and the
//                 neutral slope.  The reference scatters flux convergences into the two cells
//                 (flxconv(kd,nt,i-1|i)); scattering would race between faces, so each face writes
//                 its contribution to its minus-side and plus-side cell into two face-owned buffers.
//                 The destination index only moves down the column, so the running sum of the current
//                 destination layer sits in registers and every buffer level is written exactly once
//                 (no memset, no read-modify-write).
//   ndiff_update  per cell and level: gathers the four face contributions in the reference's pipeline
//                 order (south v face, west u face, east u face, north v face) and updates trc_rm.
// The only floating-point reassociation against the reference is that several contributions of one
// face to the same destination layer are summed before they meet the cell's running total.
#include "common.cuh"
#include "eos.cuh"

namespace blom {

namespace {

constexpr int KMN = 64;    // compile-time bound on kdm for the thread-local arrays
constexpr int NTMAX = 8;   // bound on the number of diffused scalars (2 + ntr)
constexpr double ndiff_dstsnp_fac = .01, rho_eps = 1.e-5, dp_eps = 1.e-5;  // :39-42
constexpr double mval = 1.e30;                                            // :210
constexpr int IT = 1, IS = 2;                                             // :43-45

// phy/mod_eos.F90:220-241, :284-304
__device__ __forceinline__ double eos_drhodt(double p, double th, double s) {
  const double r1 = eos::P1(p, th, s);
  const double r2i = 1. / eos::P2(p, th, s);
  return (EA12 + 2. * EA14 * th + EA15 * s + EB12 * p - (EA22 + 2. * EA24 * th + EA25 * s + EB22 * p) * r1 * r2i) * r2i;
}
__device__ __forceinline__ double eos_drhods(double p, double th, double s) {
  const double r1 = eos::P1(p, th, s);
  const double r2i = 1. / eos::P2(p, th, s);
  return (EA13 + EA15 * th + 2. * EA16 * s + EB13 * p - (EA23 + EA25 * th + 2. * EA26 * s + EB23 * p) * r1 * r2i) * r2i;
}

struct NdArgs {
  const double *p_src, *tsd, *tpc, *drdt, *drds, *p_dst, *snp;
  const int *ksmx, *kdmx, *mask;
  const double *dpml, *difiso;
  const double* tlev[NTMAX];   // scalar nt at time level nn, level 1
  const double *sca, *scbi;    // scuy,scuxi | scvx,scvyi
  const double* puv;           // pu | pv
  double *tflld, *sflld, *tflx, *sflx, *nslp;
  double *cvm, *cvp;           // face contributions to the minus / plus side cell, level (nt-1)*kk+kd
  double delt1;
  int mm, T, surface_align;
};

// one cell column with the reference's 1-based indices
struct Col {
  const double *p_src, *tsd, *tpc, *drdt, *drds, *p_dst, *snp;
  long lev; int kk;
  __device__ __forceinline__ double psd(int s, int k) const { return p_src[(long)(k + s - 2) * lev]; }
  __device__ __forceinline__ double tsrcdi(int s, int k, int nt) const { return tsd[(long)(((nt - 1) * kk + k - 1) * 2 + s - 1) * lev]; }
  __device__ __forceinline__ double tpcc(int c, int k, int nt) const { return tpc[(long)(((nt - 1) * kk + k - 1) * 5 + c - 1) * lev]; }
  __device__ __forceinline__ double drhodt(int s, int k) const { return drdt[(long)((k - 1) * 2 + s - 1) * lev]; }
  __device__ __forceinline__ double drhods(int s, int k) const { return drds[(long)((k - 1) * 2 + s - 1) * lev]; }
  __device__ __forceinline__ double pdst(int k) const { return p_dst[(long)(k - 1) * lev]; }
  __device__ __forceinline__ double dstsnp(int k) const { return snp[(long)(k - 1) * lev]; }
};

// :62-74
__device__ __forceinline__ double peval(const Col& c, int k, int nt, double x) {
  const double c5 = c.tpcc(5, k, nt), c4 = c.tpcc(4, k, nt), c3 = c.tpcc(3, k, nt), c2 = c.tpcc(2, k, nt),
               c1 = c.tpcc(1, k, nt);
  return (((c5 * x + c4) * x + c3) * x + c2) * x + c1;
}
// :76-102
__device__ __forceinline__ double pmeval(const Col& c, int k, int nt, double x0, double x1) {
  const double c1_2 = 1. / 2., c1_3 = 1. / 3., c1_4 = 1. / 4., c1_5 = 1. / 5.;
  const double b5 = c1_5 * c.tpcc(5, k, nt);
  const double b4 = b5 * x1 + c1_4 * c.tpcc(4, k, nt);
  const double b3 = b4 * x1 + c1_3 * c.tpcc(3, k, nt);
  const double b2 = b3 * x1 + c1_2 * c.tpcc(2, k, nt);
  const double b1 = b2 * x1 + c.tpcc(1, k, nt);
  return (((b5 * x0 + b4) * x0 + b3) * x0 + b2) * x0 + b1;
}
// :104-148: Newton search for the position in layer k of column c that is neutral to (tf,sf); the ten
// polynomial coefficients are loaded once instead of once per iteration
__device__ double drhoroot(const Col& c, int k, double tf, double sf, double drhodt_l, double drhodt_u,
                           double drhods_l, double drhods_u) {
  const double eps = 1.e-14, x_tol = 1.e-4;
  double x = .5;
  const double ddrdtdx = drhodt_l - drhodt_u, ddrdsdx = drhods_l - drhods_u;
  const double T1 = c.tpcc(1, k, IT), T2 = c.tpcc(2, k, IT), T3 = c.tpcc(3, k, IT), T4 = c.tpcc(4, k, IT),
               T5 = c.tpcc(5, k, IT);
  const double S1 = c.tpcc(1, k, IS), S2 = c.tpcc(2, k, IS), S3 = c.tpcc(3, k, IS), S4 = c.tpcc(4, k, IS),
               S5 = c.tpcc(5, k, IS);
  for (int n = 1; n <= 10; ++n) {
    const double dt = tf - (T1 + (T2 + (T3 + (T4 + T5 * x) * x) * x) * x);
    const double ds = sf - (S1 + (S2 + (S3 + (S4 + S5 * x) * x) * x) * x);
    const double drdt = drhodt_l * x + drhodt_u * (1. - x);
    const double drds = drhods_l * x + drhods_u * (1. - x);
    const double dtdx = -(T2 + (2. * T3 + (3. * T4 + 4. * T5 * x) * x) * x);
    const double dsdx = -(S2 + (2. * S3 + (3. * S4 + 4. * S5 * x) * x) * x);
    const double dr = drdt * dt + drds * ds;
    const double ddrdx = ddrdtdx * dt + drdt * dtdx + ddrdsdx * ds + drds * dsdx;
    const double x_old = x;
    x = fmax(0., fmin(1., x_old - dr / copysign(fmax(eps, fabs(ddrdx)), ddrdx)));
    if (fabs(x - x_old) < x_tol) return x;
  }
  return x;
}

// ndiff_prep_jslice (:959-1026) on 0..ii+1 x 0..jj+1 + the destination snapping of ndiff_flx (:491-523)
__global__ void __launch_bounds__(128)
ndiff_prep(Geom g, int mm, int T, const int* __restrict__ ip, const int* __restrict__ iu,
           const int* __restrict__ iv, const int* __restrict__ ksmx, const double* __restrict__ p_src,
           const double* __restrict__ tsd, const double* __restrict__ p_dst, int* __restrict__ kdmx,
           double* __restrict__ drdt, double* __restrict__ drds, double* __restrict__ snp,
           double* __restrict__ utflld, double* __restrict__ usflld, double* __restrict__ vtflld,
           double* __restrict__ vsflld) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i > g.ii + 1) return;
  const long x = ix2(g, i, j), lev = g.lev;
  const int kk = g.kdm;
  const bool wu = iu[x] == 1, wv = iv[x] == 1;
  if (wu || wv)
    for (int k = 1; k <= kk; ++k) {
      const long o = x + (long)(k + mm - 1) * lev;
      if (wu) { utflld[o] = 0.; usflld[o] = 0.; }
      if (wv) { vtflld[o] = 0.; vsflld[o] = 0.; }
    }
  if (ip[x] != 1) return;
  const double pbot = p_dst[x + (long)kk * lev];
  int kd = kk;
  for (int k = kk; k >= 1; --k)
    if (p_dst[x + (long)(k - 1) * lev] == pbot) kd = k - 1;
  kdmx[x] = kd;
  const int ks = ksmx[x];
  for (int k = 1; k <= ks; ++k)
    for (int s = 1; s <= 2; ++s) {
      const double ps = p_src[x + (long)(k + s - 2) * lev];
      const double t = tsd[x + (long)(((IT - 1) * kk + k - 1) * 2 + s - 1) * lev];
      const double sa = tsd[x + (long)(((IS - 1) * kk + k - 1) * 2 + s - 1) * lev];
      drdt[x + (long)((k - 1) * 2 + s - 1) * lev] = eos_drhodt(ps, t, sa);
      drds[x + (long)((k - 1) * 2 + s - 1) * lev] = eos_drhods(ps, t, sa);
    }
  // p_dstsnp(1..kdmx+1)
  double pk = p_dst[x], pk1 = p_dst[x + lev];
  snp[x] = pk;
  double dp_dst_u = pk1 - pk;
  const int kl = min(ks, kd);
  for (int k = 2; k <= kl; ++k) {
    pk = pk1; pk1 = p_dst[x + (long)k * lev];
    const double dp_dst_l = pk1 - pk;
    const double ps = p_src[x + (long)(k - 1) * lev];
    snp[x + (long)(k - 1) * lev] = fabs(pk - ps) < fmin(dp_dst_u, dp_dst_l) * ndiff_dstsnp_fac ? ps : pk;
    dp_dst_u = dp_dst_l;
  }
  for (int k = kl + 1; k <= kd + 1; ++k) snp[x + (long)(k - 1) * lev] = p_dst[x + (long)(k - 1) * lev];
}

// face-owned running sums of the flux convergence of the current destination layer of one side
template <int NT>
struct SideAcc {
  double a[NT > 0 ? NT : NTMAX];
  double* buf; long lev; int kk, T, cur;
  __device__ __forceinline__ void init(double* b, long l, int k, int t) {
    buf = b; lev = l; kk = k; T = t; cur = 0;
#pragma unroll
    for (int q = 0; q < (NT > 0 ? NT : NTMAX); ++q) a[q] = 0.;
  }
  __device__ __forceinline__ void advance(int kd) {   // kd never decreases
    if (kd == cur) return;
#pragma unroll
    for (int q = 0; q < (NT > 0 ? NT : NTMAX); ++q)
      if (q < T) {
        if (cur > 0) buf[(long)(q * kk + cur - 1) * lev] = a[q];
        for (int k = cur + 1; k < kd; ++k) buf[(long)(q * kk + k - 1) * lev] = 0.;
        a[q] = 0.;
      }
    cur = kd;
  }
  __device__ __forceinline__ void finish() { advance(kk + 1); }
};

// ndiff_flx (:160-953) for the face between cell M (i-1|j-1) and cell P (i,j)
template <int DIR, int NT>
__global__ void __launch_bounds__(128)
ndiff_face(Geom g, NdArgs A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.ii + (DIR == 0 ? 1 : 0)) return;
  const long x = ix2(g, i, j);
  if (A.mask[x] != 1) return;
  const long xm = x - (DIR == 0 ? 1 : g.ldi), lev = g.lev;
  const int kk = g.kdm, T = NT > 0 ? NT : A.T, mm = A.mm;
  const Col M{A.p_src + xm, A.tsd + xm, A.tpc + xm, A.drdt + xm, A.drds + xm, A.p_dst + xm, A.snp + xm, lev, kk};
  const Col P{A.p_src + x, A.tsd + x, A.tpc + x, A.drdt + x, A.drds + x, A.p_dst + x, A.snp + x, lev, kk};
  const int ksmx_m = A.ksmx[xm], ksmx_p = A.ksmx[x], kdmx_m = A.kdmx[xm], kdmx_p = A.kdmx[x];
  const double cdiff = A.delt1 * A.sca[x] * A.scbi[x];          // :1064 / :1126
  const double cnslp = alpha0 * A.scbi[x] / grav;

  double nslp_src[4 * (KMN + 1) + 1], p_nslp_src[4 * (KMN + 1) + 1];
  double pnm[2 * (KMN + 1) + 2], pnp[2 * (KMN + 1) + 2];
  for (int q = 0; q < 2 * (kk + 1) + 2; ++q) { pnm[q] = mval; pnp[q] = mval; }
#define PNM(s, k) pnm[2 * (k) + (s) - 1]
#define PNP(s, k) pnp[2 * (k) + (s) - 1]
  unsigned long long stab_m = 0ull, stab_p = 0ull;   // bit k-1 <-> stab(k), k = 1..64
  auto stabm = [&](int k) { return k >= 1 && ((stab_m >> (k - 1)) & 1ull); };
  auto stabp = [&](int k) { return k >= 1 && ((stab_p >> (k - 1)) & 1ull); };
  double t_ni_m[2][NT > 0 ? NT : NTMAX], t_ni_p[2][NT > 0 ? NT : NTMAX];
  double x_ni_m[2], x_ni_p[2], p_ni_m[2], p_ni_p[2];   // index nip-1 / nic-1
  double pml = 0., drho_curr = 0., p_ni_m_prev, p_ni_p_prev;
  int nns = 0, kssa_m = 0, kssa_p = 0, is_m, is_p, ks_m, ks_p;

  auto drho_at = [&]() {
    return .5 * (M.drhodt(is_m, ks_m) + P.drhodt(is_p, ks_p)) * (P.tsrcdi(is_p, ks_p, IT) - M.tsrcdi(is_m, ks_m, IT)) +
           .5 * (M.drhods(is_m, ks_m) + P.drhods(is_p, ks_p)) * (P.tsrcdi(is_p, ks_p, IS) - M.tsrcdi(is_m, ks_m, IS));
  };

  // ---- first search: neutral interfaces anchored at source layer interfaces (:212-406)
  if (A.surface_align) {
    pml = .5 * (M.psd(1, 1) + A.dpml[xm] + P.psd(1, 1) + A.dpml[x]);
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
      const bool drho_neg = drho_curr <= -rho_eps;
      const bool drho_pos = drho_curr >= rho_eps;
      const bool drho_zero = !(drho_neg || drho_pos);
      if (is_m + ks_m > 2 && is_p + ks_p > 2) {
        if (drho_neg) {
          if (is_m == 2) {
            const double dtp = P.drhodt(is_p, ks_p), dsp = P.drhods(is_p, ks_p);
            const double drhodt_x0 = .5 * (M.drhodt(1, ks_m) + dtp), drhodt_x1 = .5 * (M.drhodt(2, ks_m) + dtp);
            const double drhods_x0 = .5 * (M.drhods(1, ks_m) + dsp), drhods_x1 = .5 * (M.drhods(2, ks_m) + dsp);
            const double x_ni = drhoroot(M, ks_m, P.tsrcdi(is_p, ks_p, IT), P.tsrcdi(is_p, ks_p, IS), drhodt_x1,
                                         drhodt_x0, drhods_x1, drhods_x0);
            const double p_ni = M.psd(2, ks_m) * x_ni + M.psd(1, ks_m) * (1. - x_ni);
            if (p_ni > p_ni_m_prev) {
              p_ni_m_prev = p_ni;
              PNP(is_p, ks_p) = p_ni;
              nns = nns + 1;
              const double pp = P.psd(is_p, ks_p);
              nslp_src[nns] = -cnslp * (pp - p_ni);
              p_nslp_src[nns] = .5 * (pp + p_ni);
            }
          }
        } else if (drho_pos) {
          if (is_p == 2) {
            const double dtm = M.drhodt(is_m, ks_m), dsm = M.drhods(is_m, ks_m);
            const double drhodt_x0 = .5 * (dtm + P.drhodt(1, ks_p)), drhodt_x1 = .5 * (dtm + P.drhodt(2, ks_p));
            const double drhods_x0 = .5 * (dsm + P.drhods(1, ks_p)), drhods_x1 = .5 * (dsm + P.drhods(2, ks_p));
            const double x_ni = drhoroot(P, ks_p, M.tsrcdi(is_m, ks_m, IT), M.tsrcdi(is_m, ks_m, IS), drhodt_x1,
                                         drhodt_x0, drhods_x1, drhods_x0);
            const double p_ni = P.psd(2, ks_p) * x_ni + P.psd(1, ks_p) * (1. - x_ni);
            if (p_ni > p_ni_p_prev) {
              p_ni_p_prev = p_ni;
              PNM(is_m, ks_m) = p_ni;
              nns = nns + 1;
              const double pm = M.psd(is_m, ks_m);
              nslp_src[nns] = -cnslp * (p_ni - pm);
              p_nslp_src[nns] = .5 * (p_ni + pm);
            }
          }
        } else {
          const double pm = M.psd(is_m, ks_m), pp = P.psd(is_p, ks_p);
          PNP(is_p, ks_p) = pm;
          PNM(is_m, ks_m) = pp;
          nns = nns + 1;
          nslp_src[nns] = -cnslp * (pp - pm);
          p_nslp_src[nns] = .5 * (pp + pm);
        }
      }
      if (drho_zero || drho_pos) {
        for (;;) {
          const double drho_prev = drho_curr;
          if (is_m == 1) is_m = 2;
          else {
            ks_m = ks_m + 1;
            if (ks_m > ksmx_m) return;
            is_m = 1;
          }
          drho_curr = drho_at();
          if (drho_prev - drho_curr > rho_eps) {
            if (is_m == 2 && M.psd(2, ks_m) - M.psd(1, ks_m) > onemm) stab_m |= 1ull << (ks_m - 1);
            break;
          }
          if (is_m == 1) PNM(is_m, ks_m) = PNM(2, ks_m - 1);
        }
      }
      if (drho_zero || drho_neg) {
        for (;;) {
          const double drho_prev = drho_curr;
          if (is_p == 1) is_p = 2;
          else {
            ks_p = ks_p + 1;
            if (ks_p > ksmx_p) return;
            is_p = 1;
          }
          drho_curr = drho_at();
          if (drho_curr - drho_prev > rho_eps) {
            if (is_p == 2 && P.psd(2, ks_p) - P.psd(1, ks_p) > onemm) stab_p |= 1ull << (ks_p - 1);
            break;
          }
          if (is_p == 1) PNP(is_p, ks_p) = PNP(2, ks_p - 1);
        }
      }
    }
  }();

  if (A.surface_align) {  // :408-479
    int issa_m = 1;
    while (kssa_m <= ksmx_m) {
      if (PNM(issa_m, kssa_m) != mval) break;
      if (issa_m == 1) issa_m = 2;
      else { kssa_m = kssa_m + 1; issa_m = 1; }
    }
    int issa_p = 1;
    while (kssa_p <= ksmx_p) {
      if (PNP(issa_p, kssa_p) != mval) break;
      if (issa_p == 1) issa_p = 2;
      else { kssa_p = kssa_p + 1; issa_p = 1; }
    }
    if (kssa_m > ksmx_m || kssa_p > ksmx_p) {
      const double pbm = M.psd(2, ksmx_m), pbp = P.psd(2, ksmx_p);
      PNM(1, 1) = M.psd(1, 1);
      for (ks_m = 1; ks_m <= ksmx_m - 1; ++ks_m) {
        if (M.psd(1, ks_m) > pbp) break;
        const double p_ni = fmin(M.psd(2, ks_m), pbp);
        PNM(1, ks_m + 1) = p_ni;
        PNM(2, ks_m) = p_ni;
        stab_m |= 1ull << (ks_m - 1);
      }
      PNP(1, 1) = P.psd(1, 1);
      for (ks_p = 1; ks_p <= ksmx_p - 1; ++ks_p) {
        if (P.psd(1, ks_p) > pbm) break;
        const double p_ni = fmin(P.psd(2, ks_p), pbm);
        PNP(1, ks_p + 1) = p_ni;
        PNP(2, ks_p) = p_ni;
        stab_p |= 1ull << (ks_p - 1);
      }
    } else {
      double p1_m, p2_m, p1_p, p2_p;
      if (M.psd(issa_m, kssa_m) < PNP(issa_p, kssa_p)) {
        p1_m = M.psd(1, 1); p2_m = M.psd(issa_m, kssa_m);
        p1_p = P.psd(1, 1); p2_p = PNM(issa_m, kssa_m);
      } else {
        p1_m = M.psd(1, 1); p2_m = PNP(issa_p, kssa_p);
        p1_p = P.psd(1, 1); p2_p = P.psd(issa_p, kssa_p);
      }
      PNM(1, 1) = p1_p;
      for (ks_m = 1; ks_m <= kssa_m - 1; ++ks_m) {
        const double pl = M.psd(2, ks_m);
        const double p_ni = ((pl - p1_m) * p2_p + (p2_m - pl) * p1_p) / (p2_m - p1_m);
        PNM(1, ks_m + 1) = p_ni;
        PNM(2, ks_m) = p_ni;
        stab_m |= 1ull << (ks_m - 1);
      }
      PNP(1, 1) = p1_m;
      for (ks_p = 1; ks_p <= kssa_p - 1; ++ks_p) {
        const double pl = P.psd(2, ks_p);
        const double p_ni = ((pl - p1_p) * p2_m + (p2_p - pl) * p1_m) / (p2_p - p1_p);
        PNP(1, ks_p + 1) = p_ni;
        PNP(2, ks_p) = p_ni;
        stab_p |= 1ull << (ks_p - 1);
      }
    }
  }

  // ---- second search: neutral layers and their fluxes (:525-911)
  SideAcc<NT> accm, accp;
  accm.init(A.cvm + x, lev, kk, T);
  accp.init(A.cvp + x, lev, kk, T);
  {
    is_m = 2; ks_m = 0; is_p = 2; ks_p = 0;
    int kd_m = 0, kd_p = 0, isn_m = 1, isn_p = 1, ksn_m = 1, ksn_p = 1, ks_m_prev = 0, ks_p_prev = 0;
    bool advance_src_m = true, advance_src_p = true, advance_dst_m = true, advance_dst_p = true;
    int nip = 0, nic = 1;
    p_ni_m[nip] = -mval; p_ni_p[nip] = -mval;
    int kuv = 1;
    const double* puvx = A.puv + x;
    auto puv = [&](int k) { return puvx[(long)(k - 1) * lev]; };

    for (;;) {
      if (advance_src_m) {
        bool out = false;
        for (;;) {
          if (is_m == 1) {
            is_m = 2;
            if (stabm(ks_m)) break;
          } else {
            ks_m = ks_m + 1;
            if (ks_m > ksmx_m) { out = true; break; }
            is_m = 1;
            if (stabm(ks_m) && PNM(is_m, ks_m) != mval) break;
          }
        }
        if (out) break;
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
        bool out = false;
        for (;;) {
          if (is_p == 1) {
            is_p = 2;
            if (stabp(ks_p)) break;
          } else {
            ks_p = ks_p + 1;
            if (ks_p > ksmx_p) { out = true; break; }
            is_p = 1;
            if (stabp(ks_p) && PNP(is_p, ks_p) != mval) break;
          }
        }
        if (out) break;
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
      // the quantities every branch below looks at
      const double pnm_n = PNM(isn_m, ksn_m), pnp_n = PNP(isn_p, ksn_p);
      const double psm_n = M.psd(isn_m, ksn_m), psp_n = P.psd(isn_p, ksn_p);
      if (p_ni_m[nip] == -mval) {
        if ((pnm_n - psp_n) < (pnp_n - psm_n)) {
          p_ni_m[nip] = psm_n;
          p_ni_p[nip] = pnm_n;
        } else {
          p_ni_m[nip] = pnp_n;
          p_ni_p[nip] = psp_n;
        }
      }
      if (advance_dst_m) {
        kd_m = kd_m + 1;
        if (kd_m > kdmx_m) break;
      }
      if (advance_dst_p) {
        kd_p = kd_p + 1;
        if (kd_p > kdmx_p) break;
      }
      const double psm1 = M.psd(1, ks_m), psm2 = M.psd(2, ks_m), psp1 = P.psd(1, ks_p), psp2 = P.psd(2, ks_p);
      {
        bool out = false;
        const double lim_m = fmax(psm1, p_ni_m[nip]);
        while (M.dstsnp(kd_m + 1) <= lim_m) {
          kd_m = kd_m + 1;
          if (kd_m > kdmx_m) { out = true; break; }
        }
        if (out) break;
        const double lim_p = fmax(psp1, p_ni_p[nip]);
        while (P.dstsnp(kd_p + 1) <= lim_p) {
          kd_p = kd_p + 1;
          if (kd_p > kdmx_p) { out = true; break; }
        }
        if (out) break;
      }
      advance_src_m = false; advance_src_p = false; advance_dst_m = false; advance_dst_p = false;

      const double psm = is_m == 1 ? psm1 : psm2, psp = is_p == 1 ? psp1 : psp2;
      const double snp_m = M.dstsnp(kd_m + 1), snp_p = P.dstsnp(kd_p + 1);
      int case_m = 3;
      if (psm <= pnp_n) {
        if (psm <= snp_m) case_m = 1;
      } else if (pnp_n <= snp_m) {
        case_m = 2;
      }
      int case_p = 3;
      if (psp <= pnm_n) {
        if (psp <= snp_p) case_p = 1;
      } else if (pnm_n <= snp_p) {
        case_p = 2;
      }
      bool found_ni = false;
      auto eval_both = [&]() {
#pragma unroll
        for (int nt = 1; nt <= (NT > 0 ? NT : NTMAX); ++nt)
          if (nt <= T) {
            t_ni_m[nic][nt - 1] = peval(M, ks_m, nt, x_ni_m[nic]);
            t_ni_p[nic][nt - 1] = peval(P, ks_p, nt, x_ni_p[nic]);
          }
      };

      if (case_m == 3 && case_p == 3) {
        if (is_p == 2 && is_m == 2) {
          p_ni_m[nic] = snp_m;
          p_ni_p[nic] = snp_p;
          const double pu_m = p_ni_m[nip], pu_p = p_ni_p[nip];
          double pl_m, pl_p;
          if ((pnm_n - psp_n) < (pnp_n - psm_n)) {
            pl_m = psm_n;
            pl_p = pnm_n;
          } else {
            pl_m = pnp_n;
            pl_p = psp_n;
          }
          const double pp1 = (p_ni_m[nic] - pu_m) * (pl_p - pu_p);
          const double pp2 = (p_ni_p[nic] - pu_p) * (pl_m - pu_m);
          if (fabs(pp1 - pp2) < dp_eps * fmax(dp_eps, pl_m - pu_m + pl_p - pu_p)) {
            advance_dst_m = true;
            advance_dst_p = true;
          } else if (pp1 < pp2) {
            p_ni_p[nic] = pu_p + pp1 / (pl_m - pu_m);
            advance_dst_m = true;
          } else {
            p_ni_m[nic] = pu_m + pp2 / (pl_p - pu_p);
            advance_dst_p = true;
          }
          if (p_ni_m[nic] >= psm1 && p_ni_m[nic] <= psm2 && p_ni_p[nic] >= psp1 && p_ni_p[nic] <= psp2) {
            x_ni_m[nic] = (p_ni_m[nic] - psm1) / (psm2 - psm1);
            x_ni_p[nic] = (p_ni_p[nic] - psp1) / (psp2 - psp1);
            eval_both();
            found_ni = true;
          }
        } else {
          if (is_p != 2) advance_dst_m = true;
          if (is_m != 2) advance_dst_p = true;
        }
      } else if (case_m == 3) {
        if (is_p == 2) {
          p_ni_m[nic] = snp_m;
          if (case_p == 1)
            p_ni_p[nic] = p_ni_p[nip] + (p_ni_m[nic] - p_ni_m[nip]) * (psp_n - p_ni_p[nip]) / (pnp_n - p_ni_m[nip]);
          else
            p_ni_p[nic] = p_ni_p[nip] + (p_ni_m[nic] - p_ni_m[nip]) * (pnm_n - p_ni_p[nip]) / (psm_n - p_ni_m[nip]);
          if (p_ni_p[nic] >= psp1 && p_ni_p[nic] <= psp2) {
            x_ni_m[nic] = (snp_m - psm1) / (psm2 - psm1);
            x_ni_p[nic] = (p_ni_p[nic] - psp1) / (psp2 - psp1);
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
          p_ni_p[nic] = snp_p;
          if (case_m == 1)
            p_ni_m[nic] = p_ni_m[nip] + (p_ni_p[nic] - p_ni_p[nip]) * (psm_n - p_ni_m[nip]) / (pnm_n - p_ni_p[nip]);
          else
            p_ni_m[nic] = p_ni_m[nip] + (p_ni_p[nic] - p_ni_p[nip]) * (pnp_n - p_ni_m[nip]) / (psp_n - p_ni_p[nip]);
          if (p_ni_m[nic] >= psm1 && p_ni_m[nic] <= psm2) {
            x_ni_p[nic] = (snp_p - psp1) / (psp2 - psp1);
            x_ni_m[nic] = (p_ni_m[nic] - psm1) / (psm2 - psm1);
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
        const double pnm_c = PNM(is_m, ks_m), pnp_c = PNP(is_p, ks_p);
        if (pnm_c != mval && pnp_c != mval) {
          x_ni_m[nic] = (double)(is_m - 1);
          p_ni_m[nic] = psm;
          x_ni_p[nic] = (double)(is_p - 1);
          p_ni_p[nic] = psp;
#pragma unroll
          for (int nt = 1; nt <= (NT > 0 ? NT : NTMAX); ++nt)
            if (nt <= T) {
              t_ni_m[nic][nt - 1] = M.tsrcdi(is_m, ks_m, nt);
              t_ni_p[nic][nt - 1] = P.tsrcdi(is_p, ks_p, nt);
            }
          found_ni = true;
          advance_src_m = true;
          advance_src_p = true;
        } else {
          if (pnm_c == mval) advance_src_m = true;
          if (pnp_c == mval) advance_src_p = true;
        }
      } else if (case_m == 1) {
        const double pnm_c = PNM(is_m, ks_m);
        if (pnm_c != mval && pnm_c >= psp1) {
          x_ni_m[nic] = (double)(is_m - 1);
          p_ni_m[nic] = psm;
          p_ni_p[nic] = pnm_c;
          x_ni_p[nic] = (p_ni_p[nic] - psp1) / (psp2 - psp1);
#pragma unroll
          for (int nt = 1; nt <= (NT > 0 ? NT : NTMAX); ++nt)
            if (nt <= T) {
              t_ni_m[nic][nt - 1] = M.tsrcdi(is_m, ks_m, nt);
              t_ni_p[nic][nt - 1] = peval(P, ks_p, nt, x_ni_p[nic]);
            }
          found_ni = true;
        }
        advance_src_m = true;
      } else if (case_p == 1) {
        const double pnp_c = PNP(is_p, ks_p);
        if (pnp_c != mval && pnp_c >= psm1) {
          x_ni_p[nic] = (double)(is_p - 1);
          p_ni_p[nic] = psp;
          p_ni_m[nic] = pnp_c;
          x_ni_m[nic] = (p_ni_m[nic] - psm1) / (psm2 - psm1);
#pragma unroll
          for (int nt = 1; nt <= (NT > 0 ? NT : NTMAX); ++nt)
            if (nt <= T) {
              t_ni_p[nic][nt - 1] = P.tsrcdi(is_p, ks_p, nt);
              t_ni_m[nic][nt - 1] = peval(M, ks_m, nt, x_ni_m[nic]);
            }
          found_ni = true;
        }
        advance_src_p = true;
      } else {
        advance_src_m = true;
        advance_src_p = true;
      }

      if (found_ni) {  // :795-907
        const double dp_ni_m = fmin(p_ni_m[nic] - p_ni_m[nip], M.pdst(kd_m + 1) - M.pdst(kd_m));
        const double dp_ni_p = fmin(p_ni_p[nic] - p_ni_p[nip], P.pdst(kd_p + 1) - P.pdst(kd_p));
        const double dp_ni = 2. * dp_ni_m * dp_ni_p / fmax(dp_ni_m + dp_ni_p, 2. * dp_eps);
        if (ks_m == ks_m_prev && ks_p == ks_p_prev && p_ni_m[nip] >= M.dstsnp(kd_m) && p_ni_m[nic] <= snp_m &&
            p_ni_p[nip] >= P.dstsnp(kd_p) && p_ni_p[nic] <= snp_p && dp_ni > 2. * dp_eps) {
          accm.advance(kd_m);
          accp.advance(kd_p);
          const double q = .5 * cdiff * (A.difiso[xm + (long)(ks_m - 1) * lev] + A.difiso[x + (long)(ks_p - 1) * lev]) * dp_ni;
          double tflx = 0., sflx = 0.;
          bool ts_ok = true;
#pragma unroll
          for (int nt = 1; nt <= (NT > 0 ? NT : NTMAX); ++nt)
            if (nt <= T) {
              const double d = pmeval(M, ks_m, nt, x_ni_m[nip], x_ni_m[nic]) -
                               pmeval(P, ks_p, nt, x_ni_p[nip], x_ni_p[nic]);
              const double cm = A.tlev[nt - 1][xm + (long)(ks_m - 1) * lev], cp = A.tlev[nt - 1][x + (long)(ks_p - 1) * lev];
              const bool ok = d * (cm - cp) >= 0. && d * (t_ni_m[nip][nt - 1] - t_ni_p[nip][nt - 1]) >= 0. &&
                              d * (t_ni_m[nic][nt - 1] - t_ni_p[nic][nt - 1]) >= 0.;
              if (nt == IT) { tflx = q * d; ts_ok = ok; }
              else if (nt == IS) { sflx = q * d; ts_ok = ts_ok && ok; }
              else if (ok) {
                const double f = q * d;
                accm.a[nt - 1] += f;
                accp.a[nt - 1] -= f;
              }
            }
          if (ts_ok) {
            accm.a[IT - 1] += tflx; accp.a[IT - 1] -= tflx;
            accm.a[IS - 1] += sflx; accp.a[IS - 1] -= sflx;
            const double p_ni_up = .5 * (p_ni_m[nip] + p_ni_p[nip]);
            const double p_ni_lo = .5 * (p_ni_m[nic] + p_ni_p[nic]);
            const double dp_ni_i = 1. / fmax(epsilp, p_ni_lo - p_ni_up);
            while (kuv <= kk) {
              const long o = x + (long)(kuv + mm - 1) * lev;
              const double pk = puv(kuv), pk1 = puv(kuv + 1);
              const bool below = pk1 < p_ni_lo;
              const double mlfrac = below ? fmax(0., pk1 - fmax(p_ni_up, pk)) * dp_ni_i
                                          : (p_ni_lo - fmax(p_ni_up, pk)) * dp_ni_i;
              A.tflld[o] = A.tflld[o] + tflx * mlfrac;
              A.sflld[o] = A.sflld[o] + sflx * mlfrac;
              A.tflx[o] = A.tflx[o] + tflx * mlfrac;
              A.sflx[o] = A.sflx[o] + sflx * mlfrac;
              if (!below) break;
              kuv = kuv + 1;
            }
          }
        }
        ks_m_prev = ks_m;
        ks_p_prev = ks_p;
        nip = 1 - nip;
        nic = 1 - nic;
      }
    }
  }
  accm.finish();
  accp.finish();

  // ---- neutral slope at the destination interfaces (:913-951)
  double* nslp = A.nslp + x;
  if (nns == 0) {
    for (int k = 1; k <= kk; ++k) nslp[(long)(k - 1) * lev] = 0.;
  } else {
    double p_nslp_dst = 0.;
    int kd;
    for (kd = 1; kd <= kk; ++kd) {
      p_nslp_dst = .5 * (M.pdst(kd) + P.pdst(kd));
      if (p_nslp_dst > p_nslp_src[1]) break;
      nslp[(long)(kd - 1) * lev] = nslp_src[1];
    }
    if (kd <= kk) {
      int ks = 1;
      bool done = false;
      for (;;) {
        while (p_nslp_dst > p_nslp_src[ks]) {
          if (ks == nns) { done = true; break; }
          ks = ks + 1;
        }
        if (done) break;
        const double q = (p_nslp_src[ks] - p_nslp_dst) / fmax(p_nslp_src[ks] - p_nslp_src[ks - 1], epsilp);
        nslp[(long)(kd - 1) * lev] = q * nslp_src[ks - 1] + (1. - q) * nslp_src[ks];
        kd = kd + 1;
        if (kd > kk) break;
        p_nslp_dst = .5 * (M.pdst(kd) + P.pdst(kd));
      }
      for (; kd <= kk; ++kd) nslp[(long)(kd - 1) * lev] = nslp_src[nns];
    }
  }
#undef PNM
#undef PNP
}

// ndiff_update_trc_jslice (:1152-1175) with the gather of the four face contributions
__global__ void __launch_bounds__(256)
ndiff_update(Geom g, int T, const int* __restrict__ ip, const int* __restrict__ iu, const int* __restrict__ iv,
             const double* __restrict__ scp2, const double* __restrict__ p_dst, const double* __restrict__ ucm,
             const double* __restrict__ ucp, const double* __restrict__ vcm, const double* __restrict__ vcp,
             double* __restrict__ trc_rm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1, k = blockIdx.z + 1;
  if (i > g.ii) return;
  const long x = ix2(g, i, j), lev = g.lev;
  if (ip[x] != 1) return;
  const int kk = g.kdm;
  const double q = 1. / (scp2[x] * fmax(p_dst[x + (long)k * lev] - p_dst[x + (long)(k - 1) * lev], dp_eps));
  const bool ws = iv[x] == 1, ww = iu[x] == 1, we = iu[x + 1] == 1, wn = iv[x + g.ldi] == 1;
  for (int nt = 0; nt < T; ++nt) {
    const long o = (long)(nt * kk + k - 1) * lev;
    double conv = 0.;
    if (ws) conv += vcp[x + o];
    if (ww) conv += ucp[x + o];
    if (we) conv += ucm[x + 1 + o];
    if (wn) conv += vcm[x + g.ldi + o];
    trc_rm[x + o] = trc_rm[x + o] - q * conv;
  }
}

}  // namespace

// neutral diffusion over the whole tile in the order of the reference's slice pipeline
// (phy/mod_ale_regrid_remap.F90:1607-1690)
void ndiff_dev(int m, int n, int mm, int nn, int k1m, int k1n) {
  (void)m; (void)n; (void)k1m; (void)k1n;
  Ctx& c = C(); const Geom& g = c.g;
  const int kk = g.kdm, T = 2 + g.ntr;
  if (kk >= KMN) throw std::runtime_error("ndiff: kdm exceeds the compiled column bound (63)");
  if (T > NTMAX) throw std::runtime_error("ndiff: more than 8 diffused scalars are not compiled in");
  const bool surface_align = c.option("ndiff_surface_align", "1") == "1";   // namelist default .true.
  if (surface_align) halo_update(c.dev("dpml"), 1, 1, 1, halo_ps);
  double* drdt = c.owned("_nd_drhodt", 2 * kk);
  double* drds = c.owned("_nd_drhods", 2 * kk);
  double* snp = c.owned("_nd_dstsnp", kk + 1);
  int* kdmx = c.owned_int("_nd_kdmx", 1);
  double* ucm = c.owned("_nd_ucm", kk * T);
  double* ucp = c.owned("_nd_ucp", kk * T);
  double* vcm = c.owned("_nd_vcm", kk * T);
  double* vcp = c.owned("_nd_vcp", kk * T);

  LAUNCH(ndiff_prep, dim3(cdiv(g.ii + 2, 128), g.jj + 2), 128, 0, g, mm, T, c.idev("ip"), c.idev("iu"), c.idev("iv"),
         c.idev("nd_ksmx"), c.dev("nd_p_src"), c.dev("nd_t_srcdi"), c.dev("nd_p_dst"), kdmx, drdt, drds, snp,
         c.dev("utflld"), c.dev("usflld"), c.dev("vtflld"), c.dev("vsflld"));

  NdArgs A{};
  A.p_src = c.dev("nd_p_src"); A.tsd = c.dev("nd_t_srcdi"); A.tpc = c.dev("nd_tpc_src");
  A.drdt = drdt; A.drds = drds; A.p_dst = c.dev("nd_p_dst"); A.snp = snp;
  A.ksmx = c.idev("nd_ksmx"); A.kdmx = kdmx;
  A.dpml = c.dev("dpml"); A.difiso = c.dev("difiso");
  A.tlev[0] = c.dev("temp") + (long)nn * g.lev;
  A.tlev[1] = c.dev("saln") + (long)nn * g.lev;
  for (int nt = 3; nt <= T; ++nt) A.tlev[nt - 1] = c.dev("trc") + (long)(nn + (nt - 3) * 2 * kk) * g.lev;
  A.delt1 = c.scalar("delt1"); A.mm = mm; A.T = T; A.surface_align = surface_align ? 1 : 0;

  NdArgs U = A;
  U.mask = c.idev("iu"); U.sca = c.dev("scuy"); U.scbi = c.dev("scuxi"); U.puv = c.dev("pu");
  U.tflld = c.dev("utflld"); U.sflld = c.dev("usflld"); U.tflx = c.dev("utflx"); U.sflx = c.dev("usflx");
  U.nslp = c.dev("nslpx"); U.cvm = ucm; U.cvp = ucp;
  NdArgs V = A;
  V.mask = c.idev("iv"); V.sca = c.dev("scvx"); V.scbi = c.dev("scvyi"); V.puv = c.dev("pv");
  V.tflld = c.dev("vtflld"); V.sflld = c.dev("vsflld"); V.tflx = c.dev("vtflx"); V.sflx = c.dev("vsflx");
  V.nslp = c.dev("nslpy"); V.cvm = vcm; V.cvp = vcp;

  const dim3 gu(cdiv(g.ii + 1, 128), g.jj), gv(cdiv(g.ii, 128), g.jj + 1);
  if (T == 2) {
    LAUNCH_NAMED("ndiff_face<u>", (ndiff_face<0, 2>), gu, 128, 0, g, U);
    LAUNCH_NAMED("ndiff_face<v>", (ndiff_face<1, 2>), gv, 128, 0, g, V);
  } else if (T == 3) {
    LAUNCH_NAMED("ndiff_face<u>", (ndiff_face<0, 3>), gu, 128, 0, g, U);
    LAUNCH_NAMED("ndiff_face<v>", (ndiff_face<1, 3>), gv, 128, 0, g, V);
  } else {
    LAUNCH_NAMED("ndiff_face<u>", (ndiff_face<0, 0>), gu, 128, 0, g, U);
    LAUNCH_NAMED("ndiff_face<v>", (ndiff_face<1, 0>), gv, 128, 0, g, V);
  }
  LAUNCH(ndiff_update, dim3(cdiv(g.ii, 256), g.jj, kk), 256, 0, g, T, c.idev("ip"), c.idev("iu"), c.idev("iv"),
         c.dev("scp2"), c.dev("nd_p_dst"), ucm, ucp, vcm, vcp, c.dev("nd_trc_rm"));
}

}  // namespace blom
