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
//
// ndiff_face is issue- and latency-bound (every lane follows its own column; ~250 k thread instructions
// per face), not bandwidth-bound, so what makes it fast is fewer instructions and more lanes on the same
// path (profiles/r02_ncu_full_ndiff_face.txt): interface records of one sector,
// the neutral slope interpolated on the fly instead of from stored lists, polynomial coefficients of the
// current layers cached in shared memory, binned face fluxes summed in registers, the mirrored
// minus-/plus-column code blocks of the reference written once with the column chosen per lane, and
// compacted lists of the wet faces.
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

// Packed record of one source-layer interface of one cell column: {drhodt, drhods, T, S} at (is,k), 32 bytes =
// one memory sector.  The first search evaluates the density difference between two interfaces at every step
// and its lanes sit at different layers, so four separate level-strided arrays cost four sectors per lane
// where the record costs one.  Layout: record ((k-1)*2+is-1) of cell x at rec[(((k-1)*2+is-1)*lev + x)*4].
struct NdArgs {
  const double *p_src, *tsd, *tpc, *rec, *p_dst, *snp;
  const int *ksmx, *kdmx, *mask;
  const int* faces; int nfaces;   // offsets ix2(i,j) of the wet faces of this direction
  const double *dpml, *difiso;
  const double* tlev[NTMAX];   // scalar nt at time level nn, level 1
  const double *sca, *scbi;    // scuy,scuxi | scvx,scvyi
  const double* puv;           // pu | pv
  double *tflld, *sflld, *tflx, *sflx, *nslp;
  double *cvm, *cvp;           // face contributions to the minus / plus side cell, level (nt-1)*kk+kd
  double delt1;
  int mm, T, surface_align;
};

struct Rec { double drdt, drds, t, s; };

// :104-148: Newton search for the position in layer k of column c that is neutral to (tf,sf); the ten
// polynomial coefficients are loaded once instead of once per iteration
template <class IX>
__device__ __forceinline__ double drhoroot(const double* __restrict__ tpc, IX o, IX lev, int kk, int k, double tf, double sf,
                                           double drhodt_l, double drhodt_u, double drhods_l, double drhods_u) {
  const double eps = 1.e-14, x_tol = 1.e-4;
  double x = .5;
  const double ddrdtdx = drhodt_l - drhodt_u, ddrdsdx = drhods_l - drhods_u;
  const IX bt = o + (IX)(((IT - 1) * kk + k - 1) * 5) * lev, bs = o + (IX)(((IS - 1) * kk + k - 1) * 5) * lev;
  const double T1 = tpc[bt], T2 = tpc[bt + lev], T3 = tpc[bt + 2 * lev], T4 = tpc[bt + 3 * lev], T5 = tpc[bt + 4 * lev];
  const double S1 = tpc[bs], S2 = tpc[bs + lev], S3 = tpc[bs + 2 * lev], S4 = tpc[bs + 3 * lev], S5 = tpc[bs + 4 * lev];
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
           double* __restrict__ rec, double* __restrict__ snp,
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
      double2* r = reinterpret_cast<double2*>(rec + ((long)((k - 1) * 2 + s - 1) * lev + x) * 4);
      r[0] = make_double2(eos_drhodt(ps, t, sa), eos_drhods(ps, t, sa));
      r[1] = make_double2(t, sa);
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
template <int NT, class IX>
struct SideAcc {
  double a[NT > 0 ? NT : NTMAX];
  double* buf; IX x, lev; int kk, T, cur;
  __device__ __forceinline__ void init(double* b, IX x_, IX l, int k, int t) {
    buf = b; x = x_; lev = l; kk = k; T = t; cur = 0;
#pragma unroll
    for (int q = 0; q < (NT > 0 ? NT : NTMAX); ++q) a[q] = 0.;
  }
  __device__ __forceinline__ void advance(int kd) {   // kd never decreases
    if (kd == cur) return;
#pragma unroll
    for (int q = 0; q < (NT > 0 ? NT : NTMAX); ++q)
      if (q < T) {
        if (cur > 0) buf[x + (IX)(q * kk + cur - 1) * lev] = a[q];
        for (int k = cur + 1; k < kd; ++k) buf[x + (IX)(q * kk + k - 1) * lev] = 0.;
        a[q] = 0.;
      }
    cur = kd;
  }
  __device__ __forceinline__ void finish() { advance(kk + 1); }
};

// ndiff_flx (:160-953) for the face between cell M (i-1|j-1) and cell P (i,j)
// 128 threads per block at a register budget of 128 per thread (512 resident threads per SM).
//
// Index arithmetic: every access is `array[cell + level * lev]`.  IX is the type that arithmetic is done in:
// `unsigned` when the largest element index of the call ((5*kk*T + 1) * lev) fits 32 bits (one IMAD for the
// index and one IMAD.WIDE for the address instead of the six instructions of a 64-bit product; address
// arithmetic was a third of all instructions of this kernel), `long` otherwise (ndiff_dev chooses).  The cell
// offsets x, xm are part of the index, so no per-array cell pointers are held in registers.
constexpr int ND_BS = 128;
template <int DIR, int NT, class IX>
__global__ void __launch_bounds__(ND_BS, 512 / ND_BS)
ndiff_face(Geom g, NdArgs A) {
  // wet faces only: thread t owns face A.faces[t] (linear (i,j) offset of the face's plus-side cell).  A thread
  // of a land face would idle for the whole life of its warp - the kernel is issue- and latency-bound, so the
  // compacted list (built once, the masks are static) removes that share of the warps outright.
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= A.nfaces) return;
  const IX x = (IX)A.faces[t];
  const IX xm = x - (IX)(DIR == 0 ? 1 : g.ldi), lev = (IX)g.lev;
  const int kk = g.kdm, T = NT > 0 ? NT : A.T, mm = A.mm;
  // the slice data of a cell column o (= x or xm) with the reference's 1-based indices
  auto psd = [&](IX o, int s, int k) { return A.p_src[o + (IX)(k + s - 2) * lev]; };   // p_srcdi(s,k) = p_src(k+s-1)
  auto tsrcdi = [&](IX o, int s, int k, int nt) { return A.tsd[o + (IX)(((nt - 1) * kk + k - 1) * 2 + s - 1) * lev]; };
  auto tpcc = [&](IX o, int c, int k, int nt) { return A.tpc[o + (IX)(((nt - 1) * kk + k - 1) * 5 + c - 1) * lev]; };
  auto pdst = [&](IX o, int k) { return A.p_dst[o + (IX)(k - 1) * lev]; };
  auto dstsnp = [&](IX o, int k) { return A.snp[o + (IX)(k - 1) * lev]; };
  auto rec_at = [&](IX o, int is, int ks) {
    const double2* r = reinterpret_cast<const double2*>(A.rec) + (o + (IX)((ks - 1) * 2 + is - 1) * lev) * 2;
    const double2 a = r[0], b = r[1];
    return Rec{a.x, a.y, b.x, b.y};
  };
  // :62-74, :76-102 (only used when the coefficient cache is off)
  auto peval = [&](IX o, int k, int nt, double xx) {
    const double c5 = tpcc(o, 5, k, nt), c4 = tpcc(o, 4, k, nt), c3 = tpcc(o, 3, k, nt), c2 = tpcc(o, 2, k, nt),
                 c1 = tpcc(o, 1, k, nt);
    return (((c5 * xx + c4) * xx + c3) * xx + c2) * xx + c1;
  };
  auto pmeval = [&](IX o, int k, int nt, double x0, double x1) {
    const double c1_2 = 1. / 2., c1_3 = 1. / 3., c1_4 = 1. / 4., c1_5 = 1. / 5.;
    const double b5 = c1_5 * tpcc(o, 5, k, nt);
    const double b4 = b5 * x1 + c1_4 * tpcc(o, 4, k, nt);
    const double b3 = b4 * x1 + c1_3 * tpcc(o, 3, k, nt);
    const double b2 = b3 * x1 + c1_2 * tpcc(o, 2, k, nt);
    const double b1 = b2 * x1 + tpcc(o, 1, k, nt);
    return (((b5 * x0 + b4) * x0 + b3) * x0 + b2) * x0 + b1;
  };
  const int ksmx_m = A.ksmx[xm], ksmx_p = A.ksmx[x], kdmx_m = A.kdmx[xm], kdmx_p = A.kdmx[x];
  const double cdiff = A.delt1 * A.sca[x] * A.scbi[x];          // :1064 / :1126
  const double cnslp = alpha0 * A.scbi[x] / grav;

  double pnm[2 * (KMN + 1) + 2], pnp[2 * (KMN + 1) + 2];
  for (int q = 0; q < 2 * (kk + 1) + 2; ++q) { pnm[q] = mval; pnp[q] = mval; }
#define PNM(s, k) pnm[2 * (k) + (s) - 1]
#define PNP(s, k) pnp[2 * (k) + (s) - 1]
  unsigned long long stab_m = 0ull, stab_p = 0ull;   // bit k-1 <-> stab(k), k = 1..64
  double pml = 0., drho_curr = 0., p_ni_m_prev, p_ni_p_prev;
  int nns = 0, kssa_m = 0, kssa_p = 0, is_m, is_p, ks_m, ks_p;

  // records of the current interfaces (is_m,ks_m) and (is_p,ks_p); only the side that moved is reloaded
  Rec rm{0., 0., 0., 0.}, rp{0., 0., 0., 0.};
  auto drho_at = [&]() {
    return .5 * (rm.drdt + rp.drdt) * (rp.t - rm.t) + .5 * (rm.drds + rp.drds) * (rp.s - rm.s);
  };

  // Neutral slope at the destination interfaces (:913-951), evaluated while the first search produces the
  // (slope, pressure) pairs instead of from stored lists afterwards.  The reference scans the destination
  // interfaces kd = 1..kk in order and for each one advances a source pointer ks to the first pair at or
  // below it, never moving ks back; handing every new pair to the still-unfilled destination interfaces
  // visits exactly the same (kd, ks) combinations, so the interpolated values are identical and the two
  // lists of up to 4*(kk+1) doubles per thread never exist.
  int kd_sl = 1;
  double pd_sl = .5 * (pdst(xm, 1) + pdst(x, 1)), s_prev = 0., p_prev = 0.;
  auto emit_slope = [&](double sl, double pr) {
    nns = nns + 1;
    while (kd_sl <= kk && !(pd_sl > pr)) {
      if (nns == 1) A.nslp[x + (IX)(kd_sl - 1) * lev] = sl;
      else {
        const double q = (pr - pd_sl) / fmax(pr - p_prev, epsilp);
        A.nslp[x + (IX)(kd_sl - 1) * lev] = q * s_prev + (1. - q) * sl;
      }
      kd_sl = kd_sl + 1;
      if (kd_sl <= kk) pd_sl = .5 * (pdst(xm, kd_sl) + pdst(x, kd_sl));
    }
    s_prev = sl; p_prev = pr;
  };

  // ---- first search: neutral interfaces anchored at source layer interfaces (:212-406)
  if (A.surface_align) {
    pml = .5 * (psd(xm, 1, 1) + A.dpml[xm] + psd(x, 1, 1) + A.dpml[x]);
    kssa_m = 2;
    while (kssa_m <= ksmx_m) {
      if (psd(xm, 1, kssa_m) > pml) break;
      kssa_m = kssa_m + 1;
    }
    kssa_p = 2;
    while (kssa_p <= ksmx_p) {
      if (psd(x, 1, kssa_p) > pml) break;
      kssa_p = kssa_p + 1;
    }
    is_m = 1; ks_m = kssa_m; is_p = 1; ks_p = kssa_p;
    p_ni_m_prev = pml; p_ni_p_prev = pml;
  } else {
    is_m = 1; ks_m = 1; is_p = 1; ks_p = 1;
    p_ni_m_prev = psd(xm, 1, 1); p_ni_p_prev = psd(x, 1, 1);
  }
  if (ks_m <= ksmx_m && ks_p <= ksmx_p) { rm = rec_at(xm, is_m, ks_m); rp = rec_at(x, is_p, ks_p); drho_curr = drho_at(); }

  // search_loop1.  The reference handles the minus and the plus column in separate, mirrored code blocks
  // (root search in M when drho < 0, in P when drho > 0; advance M, then advance P).  Lanes of a warp sit in
  // different blocks at the same time, so the mirrored blocks are written ONCE with the column chosen per
  // lane (`side`: 0 = M, offset xm; 1 = P, offset x): lanes that advance M and lanes that advance P, or that
  // solve for a root in M and in P, then execute together instead of one after the other.  The arithmetic
  // per lane is the reference's (sums are commuted only where a + b == b + a exactly).
  {
    bool done1 = false;
    while (!done1 && ks_m <= ksmx_m && ks_p <= ksmx_p) {
      const bool drho_neg = drho_curr <= -rho_eps;
      const bool drho_pos = drho_curr >= rho_eps;
      const bool drho_zero = !(drho_neg || drho_pos);
      if (is_m + ks_m > 2 && is_p + ks_p > 2) {
        const bool rootm = drho_neg && is_m == 2, rootp = drho_pos && is_p == 2;
        if (rootm || rootp) {
          // layer kr of the searched column (offset o) against the fixed interface fx of the other column
          const IX o = rootm ? xm : x;
          const int kr = rootm ? ks_m : ks_p;
          const Rec fx = rootm ? rp : rm;
          const double2* r1 = reinterpret_cast<const double2*>(A.rec) + (o + (IX)((kr - 1) * 2) * lev) * 2;
          const double2 u1 = r1[0], l1 = r1[(IX)2 * lev];
          const double drhodt_x0 = .5 * (u1.x + fx.drdt), drhodt_x1 = .5 * (l1.x + fx.drdt);
          const double drhods_x0 = .5 * (u1.y + fx.drds), drhods_x1 = .5 * (l1.y + fx.drds);
          const double x_ni = drhoroot<IX>(A.tpc, o, lev, kk, kr, fx.t, fx.s, drhodt_x1, drhodt_x0, drhods_x1, drhods_x0);
          const double p_ni = psd(o, 2, kr) * x_ni + psd(o, 1, kr) * (1. - x_ni);
          if (p_ni > (rootm ? p_ni_m_prev : p_ni_p_prev)) {
            // pressure of the fixed interface in its own column
            const double pe = rootm ? psd(x, is_p, ks_p) : psd(xm, is_m, ks_m);
            if (rootm) { p_ni_m_prev = p_ni; PNP(is_p, ks_p) = p_ni; }
            else { p_ni_p_prev = p_ni; PNM(is_m, ks_m) = p_ni; }
            const double pa = rootm ? pe : p_ni, pb = rootm ? p_ni : pe;   // (plus side) - (minus side)
            emit_slope(-cnslp * (pa - pb), .5 * (pa + pb));
          }
        } else if (drho_zero) {
          const double pm = psd(xm, is_m, ks_m), pp = psd(x, is_p, ks_p);
          PNP(is_p, ks_p) = pm;
          PNM(is_m, ks_m) = pp;
          emit_slope(-cnslp * (pp - pm), .5 * (pp + pm));
        }
      }
      // advance the minus column (drho >= 0), then the plus column (drho <= 0)
      int side = (drho_zero || drho_pos) ? 0 : 1;
      bool then_p = drho_zero;
      for (;;) {
        const double drho_prev = drho_curr;
        int is = side ? is_p : is_m, ks = side ? ks_p : ks_m;
        if (is == 1) is = 2;
        else {
          ks = ks + 1;
          if (ks > (side ? ksmx_p : ksmx_m)) { if (side) ks_p = ks; else ks_m = ks; done1 = true; break; }
          is = 1;
        }
        const IX o = side ? x : xm;
        const Rec r = rec_at(o, is, ks);
        if (side) { rp = r; is_p = is; ks_p = ks; } else { rm = r; is_m = is; ks_m = ks; }
        drho_curr = drho_at();
        if ((side ? drho_curr - drho_prev : drho_prev - drho_curr) > rho_eps) {
          if (is == 2 && psd(o, 2, ks) - psd(o, 1, ks) > onemm) {
            if (side) stab_p |= 1ull << (ks - 1); else stab_m |= 1ull << (ks - 1);
          }
          if (then_p) { then_p = false; side = 1; continue; }
          break;
        }
        if (is == 1) { double* pn = side ? pnp : pnm; pn[2 * ks] = pn[2 * ks - 1]; }   // PN(1,ks) = PN(2,ks-1)
      }
    }
  }

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
      const double pbm = psd(xm, 2, ksmx_m), pbp = psd(x, 2, ksmx_p);
      PNM(1, 1) = psd(xm, 1, 1);
      for (ks_m = 1; ks_m <= ksmx_m - 1; ++ks_m) {
        if (psd(xm, 1, ks_m) > pbp) break;
        const double p_ni = fmin(psd(xm, 2, ks_m), pbp);
        PNM(1, ks_m + 1) = p_ni;
        PNM(2, ks_m) = p_ni;
        stab_m |= 1ull << (ks_m - 1);
      }
      PNP(1, 1) = psd(x, 1, 1);
      for (ks_p = 1; ks_p <= ksmx_p - 1; ++ks_p) {
        if (psd(x, 1, ks_p) > pbm) break;
        const double p_ni = fmin(psd(x, 2, ks_p), pbm);
        PNP(1, ks_p + 1) = p_ni;
        PNP(2, ks_p) = p_ni;
        stab_p |= 1ull << (ks_p - 1);
      }
    } else {
      double p1_m, p2_m, p1_p, p2_p;
      if (psd(xm, issa_m, kssa_m) < PNP(issa_p, kssa_p)) {
        p1_m = psd(xm, 1, 1); p2_m = psd(xm, issa_m, kssa_m);
        p1_p = psd(x, 1, 1); p2_p = PNM(issa_m, kssa_m);
      } else {
        p1_m = psd(xm, 1, 1); p2_m = PNP(issa_p, kssa_p);
        p1_p = psd(x, 1, 1); p2_p = psd(x, issa_p, kssa_p);
      }
      PNM(1, 1) = p1_p;
      for (ks_m = 1; ks_m <= kssa_m - 1; ++ks_m) {
        const double pl = psd(xm, 2, ks_m);
        const double p_ni = ((pl - p1_m) * p2_p + (p2_m - pl) * p1_p) / (p2_m - p1_m);
        PNM(1, ks_m + 1) = p_ni;
        PNM(2, ks_m) = p_ni;
        stab_m |= 1ull << (ks_m - 1);
      }
      PNP(1, 1) = p1_m;
      for (ks_p = 1; ks_p <= kssa_p - 1; ++ks_p) {
        const double pl = psd(x, 2, ks_p);
        const double p_ni = ((pl - p1_p) * p2_m + (p2_p - pl) * p1_m) / (p2_p - p1_p);
        PNP(1, ks_p + 1) = p_ni;
        PNP(2, ks_p) = p_ni;
        stab_p |= 1ull << (ks_p - 1);
      }
    }
  }

  // ---- second search: neutral layers and their fluxes (:525-911)
  // The reference keeps the previous/current neutral interface in two slots that swap (nip/nic); here they
  // are plain "prev"/"cur" registers and cur is copied to prev when an interface has been found (a slot
  // index would put them in local memory).  The polynomial coefficients of the current source layer of
  // each side (tpc_src, 5 per scalar) are cached while ks_m / ks_p do not change: peval and pmeval
  // are evaluated for every neutral interface found inside a layer, which took a quarter of all
  // instructions as strided loads and their address arithmetic.  The branches of the case analysis only
  // decide HOW the interface values are obtained (ev_m/ev_p); the evaluation itself and the flux
  // computation run after the branches have reconverged.
  SideAcc<NT, IX> accm, accp;
  accm.init(A.cvm, x, lev, kk, T);
  accp.init(A.cvp, x, lev, kk, T);
  {
    constexpr int NTC = NT > 0 ? NT : NTMAX;
    constexpr bool CACHE = NT > 0 && NT <= 3;     // register budget: 10 doubles per scalar
    // the cache lives in shared memory, one column of 5*NT doubles per thread and side ([..][threadIdx.x],
    // conflict-free): in registers its 40 values pushed the searches' state into spills at 128 registers
    // rows 0..2*NTC*5-1: coefficients (minus side, then plus side); then per side the layer's diffusivity and its
    // NTC layer means (difiso(ks), scalar(ks) at the new time level), which every flux of the layer reads as well
    constexpr int LM = 2 * NTC * 5;
    __shared__ double cf_sm[CACHE ? LM + 2 * (NTC + 1) : 1][ND_BS];
    struct CoefRef {   // the five coefficients of scalar nt of one side, as the polynomial helpers read them
      const double (*col)[ND_BS]; int t;
      __device__ __forceinline__ double operator[](int c5) const { return col[c5][t]; }
    };
    auto cfm = [&](int nt) { return CoefRef{&cf_sm[CACHE ? (nt - 1) * 5 : 0], (int)threadIdx.x}; };
    auto cfp = [&](int nt) { return CoefRef{&cf_sm[CACHE ? (NTC + nt - 1) * 5 : 0], (int)threadIdx.x}; };
    int kc_m = 0, kc_p = 0;                        // layers whose coefficients are cached
    auto coef = [&](IX o, int k, int nt, int row0) {
      const IX b5 = o + (IX)(((nt - 1) * kk + k - 1) * 5) * lev;
      const double c0 = A.tpc[b5], c1 = A.tpc[b5 + lev], c2 = A.tpc[b5 + 2 * lev], c3 = A.tpc[b5 + 3 * lev],
                   c4 = A.tpc[b5 + 4 * lev];
      double(*o5)[ND_BS] = &cf_sm[CACHE ? row0 : 0];
      const int tx = threadIdx.x;
      o5[0][tx] = c0; o5[1][tx] = c1; o5[2][tx] = c2; o5[3][tx] = c3; o5[4][tx] = c4;
    };
    auto layer_means = [&](IX o, int k, int row0) {
      const IX ol = o + (IX)(k - 1) * lev;
      const int tx = threadIdx.x;
      double v[NTC + 1];
      v[0] = A.difiso[ol];
#pragma unroll
      for (int nt = 1; nt <= NTC; ++nt) v[nt] = A.tlev[nt - 1][ol];
#pragma unroll
      for (int nt = 0; nt <= NTC; ++nt) cf_sm[CACHE ? row0 + nt : 0][tx] = v[nt];
    };
    auto need_m = [&]() {
      if (CACHE && kc_m != ks_m) {
#pragma unroll
        for (int nt = 1; nt <= NTC; ++nt) coef(xm, ks_m, nt, (nt - 1) * 5);
        layer_means(xm, ks_m, LM);
        kc_m = ks_m;
      }
    };
    auto need_p = [&]() {
      if (CACHE && kc_p != ks_p) {
#pragma unroll
        for (int nt = 1; nt <= NTC; ++nt) coef(x, ks_p, nt, (NTC + nt - 1) * 5);
        layer_means(x, ks_p, LM + NTC + 1);
        kc_p = ks_p;
      }
    };
    auto lmean_m = [&](int nt) { return cf_sm[CACHE ? LM + nt : 0][threadIdx.x]; };             // nt = 0: difiso
    auto lmean_p = [&](int nt) { return cf_sm[CACHE ? LM + NTC + 1 + nt : 0][threadIdx.x]; };
    auto pe = [&](const CoefRef c, double xx) { return (((c[4] * xx + c[3]) * xx + c[2]) * xx + c[1]) * xx + c[0]; };
    auto pme = [&](const CoefRef c, double x0, double x1) {
      const double c1_2 = 1. / 2., c1_3 = 1. / 3., c1_4 = 1. / 4., c1_5 = 1. / 5.;
      const double b5 = c1_5 * c[4];
      const double b4 = b5 * x1 + c1_4 * c[3];
      const double b3 = b4 * x1 + c1_3 * c[2];
      const double b2 = b3 * x1 + c1_2 * c[1];
      const double b1 = b2 * x1 + c[0];
      return (((b5 * x0 + b4) * x0 + b3) * x0 + b2) * x0 + b1;
    };

    is_m = 2; ks_m = 0; is_p = 2; ks_p = 0;
    int kd_m = 0, kd_p = 0, ks_m_prev = 0, ks_p_prev = 0;
    bool advance_src_m = true, advance_src_p = true, advance_dst_m = true, advance_dst_p = true;
    double p_prev_m = -mval, p_prev_p = -mval, p_cur_m = 0., p_cur_p = 0.;
    double x_prev_m = 0., x_prev_p = 0., x_cur_m = 0., x_cur_p = 0.;
    double t_prev_m[NTC], t_prev_p[NTC], t_cur_m[NTC], t_cur_p[NTC];
#pragma unroll
    for (int q = 0; q < NTC; ++q) { t_prev_m[q] = 0.; t_prev_p[q] = 0.; t_cur_m[q] = 0.; t_cur_p[q] = 0.; }
    // Values that only change when a column's source interface or destination layer moves are loaded at that
    // moment and kept: the interface pressures of the current source layer (ps1, ps2), pressure and neutral
    // partner of the next interface with a partner (ps_n, pn_n), the partner of the current interface (pn_c) and
    // the snapped lower interface of the current destination layer (snp).  An iteration of the search moves one
    // of the four pointers; re-reading all of these at its top cost ten loads where one to five are needed.
    double psm1 = 0., psm2 = 0., psp1 = 0., psp2 = 0., psm_n = 0., psp_n = 0., pnm_n = 0., pnp_n = 0.;
    double pnm_c = 0., pnp_c = 0., snp_m = 0., snp_p = 0.;
    int kuv = 1;
    auto puv = [&](int k) { return A.puv[x + (IX)(k - 1) * lev]; };
    // The face fluxes of a neutral sublayer are binned on the face's layers (:870-905).  A layer collects
    // the contributions of several sublayers one after the other; its four running sums (tflld, sflld, tflx,
    // sflx of layer kuv_acc) are kept in registers from the first contribution until the binning moves on
    // to the next layer: the additions and their order are the reference's, each array element is read
    // once and written once instead of once per contribution.
    int kuv_acc = 0;
    double a_tflld = 0., a_sflld = 0., a_tflx = 0., a_sflx = 0., pk_c = 0., pk1_c = 0.;
    auto flush_layer = [&]() {
      if (kuv_acc == 0) return;
      const IX o = x + (IX)(kuv_acc + mm - 1) * lev;
      A.tflld[o] = a_tflld; A.sflld[o] = a_sflld; A.tflx[o] = a_tflx; A.sflx[o] = a_sflx;
    };

    for (;;) {
      // advance to the next source interface of the minus and/or the plus column (mirrored blocks of the
      // reference, written once; lanes that advance different columns run together)
      if (advance_src_m || advance_src_p) {
        int side = advance_src_m ? 0 : 1;
        bool then_p = advance_src_m && advance_src_p, out = false;
        for (;;) {
          int is = side ? is_p : is_m, ks = side ? ks_p : ks_m;
          const int kmx = side ? ksmx_p : ksmx_m;
          const unsigned long long stab = side ? stab_p : stab_m;
          const double* pn = side ? pnp : pnm;
          for (;;) {
            if (is == 1) {
              is = 2;
              if (ks >= 1 && ((stab >> (ks - 1)) & 1ull)) break;
            } else {
              ks = ks + 1;
              if (ks > kmx) { out = true; break; }
              is = 1;
              if (((stab >> (ks - 1)) & 1ull) && pn[2 * ks + is - 1] != mval) break;
            }
          }
          if (out) break;
          int isn = is, ksn = ks;
          while (pn[2 * ksn + isn - 1] == mval) {
            if (isn == 1) isn = 2;
            else {
              if (ksn == kmx) break;
              ksn = ksn + 1;
              isn = 1;
            }
          }
          {
            const IX o = side ? x : xm;
            const double ps1 = psd(o, 1, ks), ps2 = psd(o, 2, ks), ps_n = psd(o, isn, ksn);
            const double pn_n = pn[2 * ksn + isn - 1], pn_c = pn[2 * ks + is - 1];
            if (side) { is_p = is; ks_p = ks; psp1 = ps1; psp2 = ps2; psp_n = ps_n; pnp_n = pn_n; pnp_c = pn_c; }
            else { is_m = is; ks_m = ks; psm1 = ps1; psm2 = ps2; psm_n = ps_n; pnm_n = pn_n; pnm_c = pn_c; }
          }
          if (then_p) { then_p = false; side = 1; continue; }
          break;
        }
        if (out) break;
      }
      // the quantities every branch below looks at
      if (p_prev_m == -mval) {
        if ((pnm_n - psp_n) < (pnp_n - psm_n)) {
          p_prev_m = psm_n;
          p_prev_p = pnm_n;
        } else {
          p_prev_m = pnp_n;
          p_prev_p = psp_n;
        }
      }
      if (advance_dst_m) {
        kd_m = kd_m + 1;
        if (kd_m > kdmx_m) break;
        snp_m = dstsnp(xm, kd_m + 1);
      }
      if (advance_dst_p) {
        kd_p = kd_p + 1;
        if (kd_p > kdmx_p) break;
        snp_p = dstsnp(x, kd_p + 1);
      }
      {
        bool out = false;
        const double lim_m = fmax(psm1, p_prev_m);
        while (snp_m <= lim_m) {
          kd_m = kd_m + 1;
          if (kd_m > kdmx_m) { out = true; break; }
          snp_m = dstsnp(xm, kd_m + 1);
        }
        if (out) break;
        const double lim_p = fmax(psp1, p_prev_p);
        while (snp_p <= lim_p) {
          kd_p = kd_p + 1;
          if (kd_p > kdmx_p) { out = true; break; }
          snp_p = dstsnp(x, kd_p + 1);
        }
        if (out) break;
      }
      advance_src_m = false; advance_src_p = false; advance_dst_m = false; advance_dst_p = false;

      const double psm = is_m == 1 ? psm1 : psm2, psp = is_p == 1 ? psp1 : psp2;
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
      // how the scalar values at the new interface are obtained: 1 = polynomial at x_cur, 2 = stored
      // interface value t_srcdi(is,ks)
      int ev_m = 0, ev_p = 0;

      // The reference spells the case analysis out once per column; the two columns' branches are mirror
      // images, so they are written once with the roles chosen per lane (a = the column that decides,
      // b = the other one) and lanes on mirrored branches stay together.  The local coordinates x of the
      // interface are evaluated after the branches from p_cur (same expression as in every branch).
      if (case_m == 3 && case_p == 3) {
        if (is_p == 2 && is_m == 2) {
          p_cur_m = snp_m;
          p_cur_p = snp_p;
          const double pu_m = p_prev_m, pu_p = p_prev_p;
          double pl_m, pl_p;
          if ((pnm_n - psp_n) < (pnp_n - psm_n)) {
            pl_m = psm_n;
            pl_p = pnm_n;
          } else {
            pl_m = pnp_n;
            pl_p = psp_n;
          }
          const double pp1 = (p_cur_m - pu_m) * (pl_p - pu_p);
          const double pp2 = (p_cur_p - pu_p) * (pl_m - pu_m);
          if (fabs(pp1 - pp2) < dp_eps * fmax(dp_eps, pl_m - pu_m + pl_p - pu_p)) {
            advance_dst_m = true;
            advance_dst_p = true;
          } else if (pp1 < pp2) {
            p_cur_p = pu_p + pp1 / (pl_m - pu_m);
            advance_dst_m = true;
          } else {
            p_cur_m = pu_m + pp2 / (pl_p - pu_p);
            advance_dst_p = true;
          }
          if (p_cur_m >= psm1 && p_cur_m <= psm2 && p_cur_p >= psp1 && p_cur_p <= psp2) {
            ev_m = 1; ev_p = 1;
            found_ni = true;
          }
        } else {
          if (is_p != 2) advance_dst_m = true;
          if (is_m != 2) advance_dst_p = true;
        }
      } else if (case_m == 3 || case_p == 3) {
        // exactly one column (a) meets its next destination interface first (:640-700)
        const bool am = case_m == 3;
        const int is_b = am ? is_p : is_m, case_b = am ? case_p : case_m;
        if (is_b == 2) {
          const double snp_a = am ? snp_m : snp_p;
          const double pa_prev = am ? p_prev_m : p_prev_p, pb_prev = am ? p_prev_p : p_prev_m;
          const double px = case_b == 1 ? (am ? psp_n : psm_n) : (am ? pnm_n : pnp_n);
          const double py = case_b == 1 ? (am ? pnp_n : pnm_n) : (am ? psm_n : psp_n);
          const double pb_cur = pb_prev + (snp_a - pa_prev) * (px - pb_prev) / (py - pa_prev);
          if (am) { p_cur_m = snp_a; p_cur_p = pb_cur; } else { p_cur_p = snp_a; p_cur_m = pb_cur; }
          if (pb_cur >= (am ? psp1 : psm1) && pb_cur <= (am ? psp2 : psm2)) {
            ev_m = 1; ev_p = 1;
            found_ni = true;
            if (am) advance_dst_m = true; else advance_dst_p = true;
          } else {
            const double pn_b = am ? pnp_c : pnm_c;
            if (case_b == 1 && pn_b == mval) { if (am) advance_src_p = true; else advance_src_m = true; }
            else { if (am) advance_dst_m = true; else advance_dst_p = true; }
          }
        } else {
          if (am) advance_dst_m = true; else advance_dst_p = true;
        }
      } else if (case_m == 1 && case_p == 1) {
        if (pnm_c != mval && pnp_c != mval) {
          p_cur_m = psm;
          p_cur_p = psp;
          ev_m = 2; ev_p = 2;
          found_ni = true;
          advance_src_m = true;
          advance_src_p = true;
        } else {
          if (pnm_c == mval) advance_src_m = true;
          if (pnp_c == mval) advance_src_p = true;
        }
      } else if (case_m == 1 || case_p == 1) {
        // one column (a) sits on a source interface whose neutral partner lies inside the other's layer (:745-790)
        const bool am = case_m == 1;
        const double pn_c = am ? pnm_c : pnp_c;
        if (pn_c != mval && pn_c >= (am ? psp1 : psm1)) {
          if (am) { p_cur_m = psm; p_cur_p = pn_c; ev_m = 2; ev_p = 1; }
          else { p_cur_p = psp; p_cur_m = pn_c; ev_p = 2; ev_m = 1; }
          found_ni = true;
        }
        if (am) advance_src_m = true; else advance_src_p = true;
      } else {
        advance_src_m = true;
        advance_src_p = true;
      }

      if (found_ni) {  // :795-907
        // NOTE: advance_src_* set above only act at the top of the next iteration: is/ks are still the
        // ones the interface was found for
        need_m(); need_p();
        {
          const double xm_ = (p_cur_m - psm1) / (psm2 - psm1), xp_ = (p_cur_p - psp1) / (psp2 - psp1);
          x_cur_m = ev_m == 2 ? (double)(is_m - 1) : xm_;
          x_cur_p = ev_p == 2 ? (double)(is_p - 1) : xp_;
        }
#pragma unroll
        for (int nt = 1; nt <= NTC; ++nt)
          if (nt <= T) {
            t_cur_m[nt - 1] = ev_m == 2 ? tsrcdi(xm, is_m, ks_m, nt)
                                        : (CACHE ? pe(cfm(nt), x_cur_m) : peval(xm, ks_m, nt, x_cur_m));
            t_cur_p[nt - 1] = ev_p == 2 ? tsrcdi(x, is_p, ks_p, nt)
                                        : (CACHE ? pe(cfp(nt), x_cur_p) : peval(x, ks_p, nt, x_cur_p));
          }
        const double dp_ni_m = fmin(p_cur_m - p_prev_m, pdst(xm, kd_m + 1) - pdst(xm, kd_m));
        const double dp_ni_p = fmin(p_cur_p - p_prev_p, pdst(x, kd_p + 1) - pdst(x, kd_p));
        const double dp_ni = 2. * dp_ni_m * dp_ni_p / fmax(dp_ni_m + dp_ni_p, 2. * dp_eps);
        if (ks_m == ks_m_prev && ks_p == ks_p_prev && p_prev_m >= dstsnp(xm, kd_m) && p_cur_m <= snp_m &&
            p_prev_p >= dstsnp(x, kd_p) && p_cur_p <= snp_p && dp_ni > 2. * dp_eps) {
          accm.advance(kd_m);
          accp.advance(kd_p);
          const double q = .5 * cdiff * (CACHE ? lmean_m(0) + lmean_p(0)
                                               : A.difiso[xm + (IX)(ks_m - 1) * lev] + A.difiso[x + (IX)(ks_p - 1) * lev]) * dp_ni;
          double tflx = 0., sflx = 0.;
          bool ts_ok = true;
#pragma unroll
          for (int nt = 1; nt <= NTC; ++nt)
            if (nt <= T) {
              const double d = CACHE ? pme(cfm(nt), x_prev_m, x_cur_m) - pme(cfp(nt), x_prev_p, x_cur_p)
                                     : pmeval(xm, ks_m, nt, x_prev_m, x_cur_m) - pmeval(x, ks_p, nt, x_prev_p, x_cur_p);
              const double cm = CACHE ? lmean_m(nt) : A.tlev[nt - 1][xm + (IX)(ks_m - 1) * lev];
              const double cp = CACHE ? lmean_p(nt) : A.tlev[nt - 1][x + (IX)(ks_p - 1) * lev];
              const bool ok = d * (cm - cp) >= 0. && d * (t_prev_m[nt - 1] - t_prev_p[nt - 1]) >= 0. &&
                              d * (t_cur_m[nt - 1] - t_cur_p[nt - 1]) >= 0.;
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
            const double p_ni_up = .5 * (p_prev_m + p_prev_p);
            const double p_ni_lo = .5 * (p_cur_m + p_cur_p);
            const double dp_ni_i = 1. / fmax(epsilp, p_ni_lo - p_ni_up);
            while (kuv <= kk) {
              if (kuv_acc != kuv) {   // bring layer kuv's four sums into registers (the previous layer's go out)
                flush_layer();
                const IX o = x + (IX)(kuv + mm - 1) * lev;
                a_tflld = A.tflld[o]; a_sflld = A.sflld[o]; a_tflx = A.tflx[o]; a_sflx = A.sflx[o];
                pk_c = puv(kuv); pk1_c = puv(kuv + 1);
                kuv_acc = kuv;
              }
              const double pk = pk_c, pk1 = pk1_c;
              const bool below = pk1 < p_ni_lo;
              const double mlfrac = below ? fmax(0., pk1 - fmax(p_ni_up, pk)) * dp_ni_i
                                          : (p_ni_lo - fmax(p_ni_up, pk)) * dp_ni_i;
              a_tflld = a_tflld + tflx * mlfrac;
              a_sflld = a_sflld + sflx * mlfrac;
              a_tflx = a_tflx + tflx * mlfrac;
              a_sflx = a_sflx + sflx * mlfrac;
              if (!below) break;
              kuv = kuv + 1;
            }
          }
        }
        ks_m_prev = ks_m;
        ks_p_prev = ks_p;
        p_prev_m = p_cur_m; p_prev_p = p_cur_p; x_prev_m = x_cur_m; x_prev_p = x_cur_p;
#pragma unroll
        for (int q2 = 0; q2 < NTC; ++q2) { t_prev_m[q2] = t_cur_m[q2]; t_prev_p[q2] = t_cur_p[q2]; }
      }
    }
    flush_layer();
  }
  accm.finish();
  accp.finish();

  // ---- neutral slope at the destination interfaces below the last pair (:913-951, tail of emit_slope)
  for (; kd_sl <= kk; ++kd_sl) A.nslp[x + (IX)(kd_sl - 1) * lev] = nns == 0 ? 0. : s_prev;
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

#ifndef BLOM_HOST_EMUL
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
  double* rec = c.owned("_nd_rec", 8 * kk);   // 2*kk interface records of 4 doubles per cell
  double* snp = c.owned("_nd_dstsnp", kk + 1);
  int* kdmx = c.owned_int("_nd_kdmx", 1);
  double* ucm = c.owned("_nd_ucm", kk * T);
  double* ucp = c.owned("_nd_ucp", kk * T);
  double* vcm = c.owned("_nd_vcm", kk * T);
  double* vcp = c.owned("_nd_vcp", kk * T);

  LAUNCH(ndiff_prep, dim3(cdiv(g.ii + 2, 128), g.jj + 2), 128, 0, g, mm, T, c.idev("ip"), c.idev("iu"), c.idev("iv"),
         c.idev("nd_ksmx"), c.dev("nd_p_src"), c.dev("nd_t_srcdi"), c.dev("nd_p_dst"), kdmx, rec, snp,
         c.dev("utflld"), c.dev("usflld"), c.dev("vtflld"), c.dev("vsflld"));

  NdArgs A{};
  A.p_src = c.dev("nd_p_src"); A.tsd = c.dev("nd_t_srcdi"); A.tpc = c.dev("nd_tpc_src");
  A.rec = rec; A.p_dst = c.dev("nd_p_dst"); A.snp = snp;
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

  // compacted lists of the wet u faces (1..ii+1 x 1..jj) and v faces (1..ii x 1..jj+1); the masks are static
  // after bigrid, so the lists are built once per tile
  int* list_u = c.owned_int("_nd_faces_u", 1);
  int* list_v = c.owned_int("_nd_faces_v", 1);
  if (!c.sc.count("_nd_nfaces_u")) {   // (bigrid erases the counts when the masks are rebuilt)
    std::vector<int> hu(g.lev), hv(g.lev), lu, lv;
    CUDA_CHECK(cudaMemcpyAsync(hu.data(), c.idev("iu"), sizeof(int) * g.lev, cudaMemcpyDeviceToHost, c.stream));
    CUDA_CHECK(cudaMemcpyAsync(hv.data(), c.idev("iv"), sizeof(int) * g.lev, cudaMemcpyDeviceToHost, c.stream));
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
    for (int j = 1; j <= g.jj; ++j)
      for (int i = 1; i <= g.ii + 1; ++i) {
        const long x = ix2(g, i, j);
        if (hu[x] == 1) lu.push_back((int)x);
      }
    // v faces in tiles of 32 (i) x 4 (j): a 128-thread block then owns four consecutive rows of a 32-wide strip, and
    // the cell column (i,j) that is the plus side of face (i,j) and the minus side of face (i,j+1) is fetched by one
    // block instead of by two blocks a whole grid row apart (ncu, row-major order: 75 GB of DRAM traffic per launch
    // for the v faces against 49 GB for the u faces, whose two columns sit in neighbouring lanes)
    for (int j0 = 1; j0 <= g.jj + 1; j0 += 4)
      for (int i0 = 1; i0 <= g.ii; i0 += 32)
        for (int j = j0; j < std::min(j0 + 4, g.jj + 2); ++j)
          for (int i = i0; i < std::min(i0 + 32, g.ii + 1); ++i) {
            const long x = ix2(g, i, j);
            if (hv[x] == 1) lv.push_back((int)x);
          }
    if (!lu.empty()) CUDA_CHECK(cudaMemcpyAsync(list_u, lu.data(), sizeof(int) * lu.size(), cudaMemcpyHostToDevice, c.stream));
    if (!lv.empty()) CUDA_CHECK(cudaMemcpyAsync(list_v, lv.data(), sizeof(int) * lv.size(), cudaMemcpyHostToDevice, c.stream));
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
    c.sc["_nd_nfaces_u"] = (double)lu.size();
    c.sc["_nd_nfaces_v"] = (double)lv.size();
  }
  U.faces = list_u; U.nfaces = (int)c.sc["_nd_nfaces_u"];
  V.faces = list_v; V.nfaces = (int)c.sc["_nd_nfaces_v"];
  // 32-bit index arithmetic in ndiff_face when every element index of the call fits (tpc_src is the largest array:
  // 5*kk*T levels; the interface records are addressed as double2, 4*kk levels of pairs): true for every tile that
  // fits a B200 with T <= 3, e.g. tnx0.25v4 on one GPU; the 64-bit instantiation covers the rest
  const bool ix32 = ((long)5 * kk * T + 2) * g.lev < (1l << 32) && ((long)8 * kk + 4) * g.lev < (1l << 32);
  const dim3 gu(std::max(1, cdiv(U.nfaces, ND_BS))), gv(std::max(1, cdiv(V.nfaces, ND_BS)));
#define ND_LAUNCH(NT_, IX_)                                                                        \
  do {                                                                                             \
    LAUNCH_NAMED("ndiff_face<u>", (ndiff_face<0, NT_, IX_>), gu, ND_BS, 0, g, U);                  \
    LAUNCH_NAMED("ndiff_face<v>", (ndiff_face<1, NT_, IX_>), gv, ND_BS, 0, g, V);                  \
  } while (0)
#define ND_FACE(NT_)                                                                               \
  do {                                                                                             \
    if (ix32) ND_LAUNCH(NT_, unsigned);                                                            \
    else ND_LAUNCH(NT_, long);                                                                     \
  } while (0)
  if (T == 2) { ND_FACE(2); }
  else if (T == 3) { ND_FACE(3); }
  else { ND_FACE(0); }
#undef ND_FACE
#undef ND_LAUNCH
  LAUNCH(ndiff_update, dim3(cdiv(g.ii, 256), g.jj, kk), 256, 0, g, T, c.idev("ip"), c.idev("iu"), c.idev("iv"),
         c.dev("scp2"), c.dev("nd_p_dst"), ucm, ucp, vcm, vcp, c.dev("nd_trc_rm"));
}
#endif  // BLOM_HOST_EMUL

}  // namespace blom
