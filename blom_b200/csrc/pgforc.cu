// Pressure-gradient force (phy/mod_pgforc.F90:438-615): pgfmth='dynamic enthalpy' (:262-408, the
// default) and pgfmth='geopotential' (:95-260).
//
// B200 design: the reference stages five kdm-level temporaries (pot_dynh,
// pot_dynh_pb, dynh_a, dynh_t, alpha_r) through memory between its column sweep
// and its gradient sweep.  Here one kernel marches every column bottom-up
// (k=kk..1, the direction of both the potential recurrences and the reference's
// accumulation order), a 32x8 thread tile shares the per-level column values
// with its west/south neighbours through double-buffered shared memory, and
// the layer gradients, pgfx_o/pgfy_o copies and the vertical sums pgfxm/xix*
// come out of the same pass: per cell it reads p,dp,T,S,dpu,dpv,pgfx,pgfy once
// and writes phi,pgfx,pgfy,pgfx_o,pgfy_o once.
#include "common.cuh"
#include "eos.cuh"

namespace blom {

namespace {

// p(k+1) = p(k) + dp(kn) on -2..ii+2 x -2..jj+2  (:452-461)
__global__ void pg_p_from_dp(Geom g, int nn, const int* __restrict__ ip, const double* __restrict__ dp,
                             double* __restrict__ p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x - 2;
  const int j = (int)blockIdx.y - 2;
  if (i > g.ii + 2) return;
  const long x = ix2(g, i, j);
  if (ip[x] != 1) return;
  double pk = p[x];
  for (int k = 1; k <= g.kdm; ++k) {
    pk = pk + dp[x + (long)(k + nn - 1) * g.lev];
    p[x + (long)k * g.lev] = pk;
  }
}

// dpu,dpv(kn), pu,pv(k+1) on -1..ii+2 x -1..jj+2 (:463-484); old 2-D fields (:488-504)
__global__ void pg_dpuv(Geom g, int n, int nn, const int* __restrict__ iu, const int* __restrict__ iv,
                        const double* __restrict__ p, double* __restrict__ dpu, double* __restrict__ dpv,
                        double* __restrict__ pu, double* __restrict__ pv, const double* __restrict__ xixp,
                        const double* __restrict__ xixm, const double* __restrict__ pgfxm,
                        const double* __restrict__ xiyp, const double* __restrict__ xiym,
                        const double* __restrict__ pgfym, double* __restrict__ xixp_o,
                        double* __restrict__ xixm_o, double* __restrict__ pgfxm_o, double* __restrict__ xiyp_o,
                        double* __restrict__ xiym_o, double* __restrict__ pgfym_o) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x - 1;
  const int j = (int)blockIdx.y - 1;
  if (i > g.ii + 2) return;
  const long x = ix2(g, i, j), s = g.ldi, xb = x + (long)g.kdm * g.lev, x2 = x + (long)(n - 1) * g.lev;
  const bool isu = iu[x] == 1, isv = iv[x] == 1;
  if (i >= 0 && i <= g.ii + 1) {
    if (isu) { xixp_o[x] = xixp[x2]; xixm_o[x] = xixm[x2]; pgfxm_o[x] = pgfxm[x2]; }
    if (isv) { xiyp_o[x] = xiyp[x2]; xiym_o[x] = xiym[x2]; pgfym_o[x] = pgfym[x2]; }
  }
  if (!isu && !isv) return;
  const double qu = isu ? fmin(p[xb], p[xb - 1]) : 0., qv = isv ? fmin(p[xb], p[xb - s]) : 0.;
  double pc0 = p[x], pw0 = isu ? p[x - 1] : 0., ps0 = isv ? p[x - s] : 0.;
  double puk = isu ? pu[x] : 0., pvk = isv ? pv[x] : 0.;
  for (int k = 1; k <= g.kdm; ++k) {
    const long x1 = x + (long)k * g.lev, xn = x + (long)(k + nn - 1) * g.lev;
    const double pc1 = p[x1];
    if (isu) {
      const double pw1 = p[x1 - 1];
      const double d = .5 * ((fmin(qu, pw1) - fmin(qu, pw0)) + (fmin(qu, pc1) - fmin(qu, pc0)));
      dpu[xn] = d;
      puk = puk + d;
      pu[x1] = puk;
      pw0 = pw1;
    }
    if (isv) {
      const double ps1 = p[x1 - s];
      const double d = .5 * ((fmin(qv, ps1) - fmin(qv, ps0)) + (fmin(qv, pc1) - fmin(qv, pc0)));
      dpv[xn] = d;
      pvk = pvk + d;
      pv[x1] = pvk;
      ps0 = ps1;
    }
    pc0 = pc1;
  }
}

constexpr int TX = 32, TY = 8, NV = 7;  // values shared per column and level

// bottom-up march; thread (tx,ty) owns p-column (i,j) = (bx*31+tx, by*7+ty), i.e. a one-column
// skirt on the west and south of the 31x7 output tile.
template <int MINB>
__global__ void __launch_bounds__(TX* TY, MINB)
pg_dynh_march(Geom g, eos::Coef ec, int n, int nn, const int* __restrict__ ip, const int* __restrict__ iu,
              const int* __restrict__ iv, const double* __restrict__ p, const double* __restrict__ dp,
              const double* __restrict__ temp, const double* __restrict__ saln, const double* __restrict__ dpu,
              const double* __restrict__ dpv, double* __restrict__ phi, double* __restrict__ pgfx,
              double* __restrict__ pgfy, double* __restrict__ pgfx_o, double* __restrict__ pgfy_o,
              double* __restrict__ pgfxm, double* __restrict__ xixp, double* __restrict__ xixm,
              double* __restrict__ pgfym, double* __restrict__ xiyp, double* __restrict__ xiym) {
  __shared__ double sm[2][NV][TY][TX + 1];
  const double p0_dynh = 0.0;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int i = blockIdx.x * (TX - 1) + tx, j = blockIdx.y * (TY - 1) + ty;  // from 0
  const bool inr = i <= g.ii && j <= g.jj;
  const long x = ix2(g, min(i, g.ii), min(j, g.jj)), s = g.ldi;
  const bool col = inr && ip[x] == 1;
  const bool own = inr && tx >= 1 && ty >= 1;  // i>=1, j>=1 follow from tx,ty>=1
  const bool isu = own && iu[x] == 1, isv = own && iv[x] == 1;
  const int kk = g.kdm;
  double pot = 0., potpb = 0., phik = col ? phi[x + (long)kk * g.lev] : 0.;
  double pk1 = col ? p[x + (long)kk * g.lev] : 0.;
  double tn = 0., sn = 0.;
  double a_pgfxm = 0., a_xixm = 0., a_xixp = 0., a_pgfym = 0., a_xiym = 0., a_xiyp = 0.;
  // level operands are fetched one level ahead of the (EOS-heavy) computation that uses them
  double t_f = 0., sa_f = 0., dpk_f = 0., pk_f = 0., gx_f = 0., du_f = 0., gy_f = 0., dv_f = 0.;
  auto fetch = [&](int k) {
    const long xn = x + (long)(k + nn - 1) * g.lev;
    if (col) { t_f = temp[xn]; sa_f = saln[xn]; dpk_f = dp[xn]; pk_f = p[x + (long)(k - 1) * g.lev]; }
    if (isu) { gx_f = pgfx[xn]; du_f = dpu[xn]; }
    if (isv) { gy_f = pgfy[xn]; dv_f = dpv[xn]; }
  };
  fetch(kk);
  for (int k = kk; k >= 1; --k) {
    const int b = k & 1;
    const long xn = x + (long)(k + nn - 1) * g.lev, xk = x + (long)(k - 1) * g.lev;
    double t = 0., sa = 0., dyn_a = 0., dyn_t = 0., ar = 0., dpk = 0.;
    const double pk = pk_f, gx = gx_f, du = du_f, gy = gy_f, dv = dv_f;
    t = t_f; sa = sa_f; dpk = dpk_f;
    if (k > 1) fetch(k - 1);
    if (col) {
      if (k == kk) {
        pot = phik + eos::p_alpha(p0_dynh, pk1, t, sa);
        potpb = eos::alp(pk1, t, sa) * pk1;
      } else {
        pot = pot + eos::p_alpha(p0_dynh, pk1, t, sa) - eos::p_alpha(p0_dynh, pk1, tn, sn);
        potpb = potpb + (eos::alp(pk1, t, sa) - eos::alp(pk1, tn, sn)) * pk1;
      }
      phik = phik + eos::p_alpha(pk, pk1, t, sa);
      phi[xk] = phik;
      if (!(dpk < onemm)) {
        double d_t, d_s;
        eos::dynh_derivatives(p0_dynh, pk, pk1, t, sa, d_t, d_s);
        dyn_a = d_s / eos::dalpds(ec.pref, t, sa);
        dyn_t = d_t - dyn_a * eos::dalpdt(ec.pref, t, sa);
      }
      ar = eos::alp(ec.pref, t, sa);
      pk1 = pk; tn = t; sn = sa;
    }
    sm[b][0][ty][tx] = pot; sm[b][1][ty][tx] = potpb; sm[b][2][ty][tx] = dyn_a; sm[b][3][ty][tx] = dyn_t;
    sm[b][4][ty][tx] = ar; sm[b][5][ty][tx] = t; sm[b][6][ty][tx] = dpk;
    __syncthreads();
    if (isu) {
      const double pot_w = sm[b][0][ty][tx - 1], potpb_w = sm[b][1][ty][tx - 1];
      double f = -(pot - pot_w);
      if (sm[b][6][ty][tx - 1] >= onemm && dpk >= onemm)
        f = f + .5 * ((sm[b][3][ty][tx - 1] + dyn_t) * (t - sm[b][5][ty][tx - 1]) +
                      (sm[b][2][ty][tx - 1] + dyn_a) * (ar - sm[b][4][ty][tx - 1]));
      pgfx_o[xk] = gx;
      pgfx[xn] = f;
      a_pgfxm = a_pgfxm + f * du;
      a_xixm = a_xixm + potpb_w * du;
      a_xixp = a_xixp + potpb * du;
    }
    if (isv) {
      const double pot_s = sm[b][0][ty - 1][tx], potpb_s = sm[b][1][ty - 1][tx];
      double f = -(pot - pot_s);
      if (sm[b][6][ty - 1][tx] >= onemm && dpk >= onemm)
        f = f + .5 * ((sm[b][3][ty - 1][tx] + dyn_t) * (t - sm[b][5][ty - 1][tx]) +
                      (sm[b][2][ty - 1][tx] + dyn_a) * (ar - sm[b][4][ty - 1][tx]));
      pgfy_o[xk] = gy;
      pgfy[xn] = f;
      a_pgfym = a_pgfym + f * dv;
      a_xiym = a_xiym + potpb_s * dv;
      a_xiyp = a_xiyp + potpb * dv;
    }
  }
  const long x2 = x + (long)(n - 1) * g.lev;
  if (isu) { pgfxm[x2] = a_pgfxm; xixm[x2] = a_xixm; xixp[x2] = a_xixp; }
  if (isv) { pgfym[x2] = a_pgfym; xiym[x2] = a_xiym; xiyp[x2] = a_xiyp; }
  (void)s;
}

// ---- pgfmth='geopotential' (:95-260) ---------------------------------------------------------
// Column pass: geopotential phi at the layer interfaces and phip = sum of p*alpha jumps, marched
// bottom-up on 0..ii x 0..jj (:114-137); one thread per column, lanes along i.
__global__ void pg_geop_column(Geom g, int nn, const int* __restrict__ ip, const double* __restrict__ p,
                               const double* __restrict__ dp, const double* __restrict__ temp,
                               const double* __restrict__ saln, double* __restrict__ phi,
                               double* __restrict__ phip) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i > g.ii) return;
  const long x = ix2(g, i, j);
  if (ip[x] != 1) return;
  const int kk = g.kdm;
  double phik = phi[x + (long)kk * g.lev], phipk = 0., pk1 = p[x + (long)kk * g.lev];
  phip[x + (long)kk * g.lev] = 0.;
  for (int k = kk; k >= 1; --k) {
    const long xn = x + (long)(k + nn - 1) * g.lev, xk = x + (long)(k - 1) * g.lev;
    const double pk = p[xk];
    if (!(dp[xn] < epsilp)) {
      double dphi, alpu, alpl;
      eos::delphi(pk, pk1, temp[xn], saln[xn], dphi, alpu, alpl);
      phik = phik - dphi;
      phipk = phipk + pk1 * alpl - pk * alpu;
    }
    phi[xk] = phik;
    phip[xk] = phipk;
    pk1 = pk;
  }
}

// Face pass (:141-256): for every u (DIR 0) or v (DIR 1) face column march bottom-up, track the
// layers kp/km of the two adjacent columns that contain the face's mid-layer pressure (monotone
// search, the reference's do-while), evaluate both geopotentials at that pressure and difference
// them.  Also takes the pgfx_o/pgfy_o copy of the caller (:507-521) so pgfx is touched once.
template <int DIR>
__global__ void pg_geop_face(Geom g, int n, int nn, const int* __restrict__ imask, const double* __restrict__ p,
                             const double* __restrict__ pd, const double* __restrict__ dpd,
                             const double* __restrict__ temp, const double* __restrict__ saln,
                             const double* __restrict__ phi, const double* __restrict__ phip,
                             double* __restrict__ pgf, double* __restrict__ pgf_o, double* __restrict__ pgfm,
                             double* __restrict__ xip, double* __restrict__ xim) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.ii) return;
  const long x = ix2(g, i, j);
  if (imask[x] != 1) return;
  const long xm = x - (DIR == 0 ? 1 : g.ldi), lev = g.lev;
  const int kk = g.kdm;
  int kp = kk, km = kk;
  double a_xip = 0., a_xim = 0., a_pgfm = 0.;
  double pp1 = p[x + (long)kk * lev], pm1 = p[xm + (long)kk * lev];   // p(k+1) of the two columns
  for (int k = kk; k >= 1; --k) {
    const long xn = x + (long)(k + nn - 1) * lev, xk = x + (long)(k - 1) * lev;
    const double dd = dpd[xn];
    const double prs = pd[x + (long)k * lev] - .5 * dd;
    // p(.,1) = 0 <= prs stops the search at the surface; the bound only guards corrupt input
    while (kp > 1 && p[x + (long)(kp - 1) * lev] > prs) kp = kp - 1;
    while (km > 1 && p[xm + (long)(km - 1) * lev] > prs) km = km - 1;
    double dphip, alpup, alplp, dphim, alpum, alplm;
    const double pkp1 = p[x + (long)kp * lev], pkm1 = p[xm + (long)km * lev];
    eos::delphi(prs, pkp1, temp[x + (long)(kp + nn - 1) * lev], saln[x + (long)(kp + nn - 1) * lev], dphip, alpup,
                alplp);
    eos::delphi(prs, pkm1, temp[xm + (long)(km + nn - 1) * lev], saln[xm + (long)(km + nn - 1) * lev], dphim, alpum,
                alplm);
    const double pp0 = p[xk], pm0 = p[xm + (long)(k - 1) * lev];
    double cp = .25 * (pp1 + pp0);
    double cm = .25 * (pm1 + pm0);
    const double q = prs / (cp + cm);
    cp = q * cp;
    cm = q * cm;
    const double phi_p = phi[x + (long)kp * lev] - dphip;
    a_xip = a_xip + (phip[x + (long)kp * lev] + pkp1 * alplp - cp * (alpup - alpum)) * dd;
    const double phi_m = phi[xm + (long)km * lev] - dphim;
    a_xim = a_xim + (phip[xm + (long)km * lev] + pkm1 * alplm - cm * (alpum - alpup)) * dd;
    const double f = -(phi_p - phi_m);
    pgf_o[xk] = pgf[xn];
    pgf[xn] = f;
    a_pgfm = a_pgfm + f * dd;
    pp1 = pp0; pm1 = pm0;
  }
  const long x2 = x + (long)(n - 1) * lev;
  pgfm[x2] = a_pgfm; xip[x2] = a_xip; xim[x2] = a_xim;
}

// depth-mean removal and normalisation (:543-597); one thread per interior column
__global__ void pg_finalize(Geom g, int n, int nn, const int* __restrict__ ip, const int* __restrict__ iu,
                            const int* __restrict__ iv, const double* __restrict__ pb_p,
                            const double* __restrict__ pbu_p, const double* __restrict__ pbv_p,
                            const double* __restrict__ phi, double* __restrict__ pgfx, double* __restrict__ pgfy,
                            double* __restrict__ pgfxm, double* __restrict__ xixp, double* __restrict__ xixm,
                            double* __restrict__ pgfym, double* __restrict__ xiyp, double* __restrict__ xiym,
                            double* __restrict__ sealv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y + 1;
  if (i > g.ii) return;
  const long x = ix2(g, i, j), x2 = x + (long)(n - 1) * g.lev, s = g.ldi;
  if (iu[x] == 1) {
    const double q = 1. / pbu_p[x];
    const double m_ = pgfxm[x2] * q, xp = xixp[x2] * q, xm = xixm[x2] * q;
    for (int k = 1; k <= g.kdm; ++k) {
      const long xn = x + (long)(k + nn - 1) * g.lev;
      pgfx[xn] = pgfx[xn] - m_;
    }
    pgfxm[x2] = m_ + xp - xm;
    xixp[x2] = xp / pb_p[x];
    xixm[x2] = xm / pb_p[x - 1];
  }
  if (iv[x] == 1) {
    const double q = 1. / pbv_p[x];
    const double m_ = pgfym[x2] * q, yp = xiyp[x2] * q, ym = xiym[x2] * q;
    for (int k = 1; k <= g.kdm; ++k) {
      const long xn = x + (long)(k + nn - 1) * g.lev;
      pgfy[xn] = pgfy[xn] - m_;
    }
    pgfym[x2] = m_ + yp - ym;
    xiyp[x2] = yp / pb_p[x];
    xiym[x2] = ym / pb_p[x - s];
  }
  if (ip[x] == 1) sealv[x] = phi[x] / grav;
}

}  // namespace

void pgforc_dev(int m, int n, int mm, int nn, int k1m, int k1n) {
  (void)m; (void)mm; (void)k1m; (void)k1n;
  Ctx& c = C(); const Geom& g = c.g;
  const std::string pgfmth = c.option("pgfmth", "dynamic enthalpy");
  if (pgfmth != "dynamic enthalpy" && pgfmth != "geopotential")
    throw std::runtime_error(" pgfmth = " + pgfmth + " is unsupported!");
  {
    dim3 grid(cdiv(g.ii + 5, 128), g.jj + 5);
    LAUNCH(pg_p_from_dp, grid, 128, 0, g, nn, c.idev("ip"), c.dev("dp"), c.dev("p"));
  }
  {
    dim3 grid(cdiv(g.ii + 4, 128), g.jj + 4);
    LAUNCH(pg_dpuv, grid, 128, 0, g, n, nn, c.idev("iu"), c.idev("iv"), c.dev("p"), c.dev("dpu"), c.dev("dpv"),
           c.dev("pu"), c.dev("pv"), c.dev("xixp"), c.dev("xixm"), c.dev("pgfxm"), c.dev("xiyp"), c.dev("xiym"),
           c.dev("pgfym"), c.dev("xixp_o"), c.dev("xixm_o"), c.dev("pgfxm_o"), c.dev("xiyp_o"), c.dev("xiym_o"),
           c.dev("pgfym_o"));
  }
  if (pgfmth == "geopotential") {
    double* phip = c.owned("pg_phip", g.kdm + 1);   // routine-local phip of the reference (:105)
    dim3 gridc(cdiv(g.ii + 1, 128), g.jj + 1), gridf(cdiv(g.ii, 128), g.jj);
    LAUNCH(pg_geop_column, gridc, 128, 0, g, nn, c.idev("ip"), c.dev("p"), c.dev("dp"), c.dev("temp"),
           c.dev("saln"), c.dev("phi"), phip);
    LAUNCH(pg_geop_face<0>, gridf, 128, 0, g, n, nn, c.idev("iu"), c.dev("p"), c.dev("pu"), c.dev("dpu"),
           c.dev("temp"), c.dev("saln"), c.dev("phi"), phip, c.dev("pgfx"), c.dev("pgfx_o"), c.dev("pgfxm"),
           c.dev("xixp"), c.dev("xixm"));
    LAUNCH(pg_geop_face<1>, gridf, 128, 0, g, n, nn, c.idev("iv"), c.dev("p"), c.dev("pv"), c.dev("dpv"),
           c.dev("temp"), c.dev("saln"), c.dev("phi"), phip, c.dev("pgfy"), c.dev("pgfy_o"), c.dev("pgfym"),
           c.dev("xiyp"), c.dev("xiym"));
  } else {
    dim3 grid(cdiv(g.ii + 1, TX - 1), cdiv(g.jj + 1, TY - 1)), block(TX, TY);
    OCC_DISPATCH3("pgforc_minblk", 2, 2, 3, 4,
    LAUNCH_NAMED("pg_dynh_march", pg_dynh_march<OCC>, grid, block, 0, g, eos::host_coef(), n, nn, c.idev("ip"), c.idev("iu"), c.idev("iv"),
           c.dev("p"), c.dev("dp"), c.dev("temp"), c.dev("saln"), c.dev("dpu"), c.dev("dpv"), c.dev("phi"),
           c.dev("pgfx"), c.dev("pgfy"), c.dev("pgfx_o"), c.dev("pgfy_o"), c.dev("pgfxm"), c.dev("xixp"),
           c.dev("xixm"), c.dev("pgfym"), c.dev("xiyp"), c.dev("xiym")));
  }
  halo_update(c.dev("pb_p"), 1, 1, 1, halo_ps);
  {
    dim3 grid(cdiv(g.ii, 128), g.jj);
    LAUNCH(pg_finalize, grid, 128, 0, g, n, nn, c.idev("ip"), c.idev("iu"), c.idev("iv"), c.dev("pb_p"),
           c.dev("pbu_p"), c.dev("pbv_p"), c.dev("phi"), c.dev("pgfx"), c.dev("pgfy"), c.dev("pgfxm"),
           c.dev("xixp"), c.dev("xixm"), c.dev("pgfym"), c.dev("xiyp"), c.dev("xiym"), c.dev("sealv"));
  }
}

}  // namespace blom
