// Layer-wise isopycnal tracer diffusion (diffus, phy/mod_diffus.F90:41-185), the
// time smoother (tmsmt1/tmsmt2, phy/mod_tmsmt.F90:209-410) and inieos
// (phy/mod_eos.F90:83-155).
//
// diffus is two streaming kernels per call over all levels at once (the
// reference loops levels outside three masked 2-D sweeps): `diffus_flux` writes
// the face fluxes (usflld.. are state arrays of the reference, so they go to HBM
// anyway) and accumulates them into utflx..; `diffus_update` applies the
// divergence in place and refreshes sigma.  tmsmt2 is one thread per column
// with i across lanes (coalesced), the two column sums kept in registers.
#include "common.cuh"
#include "eos.cuh"

namespace blom {

namespace eos {
static Coef g_coef;
static bool g_coef_set = false;
const Coef& host_coef() {
  if (!g_coef_set) throw std::runtime_error("blomgpu: inieos has not been called");
  return g_coef;
}
}  // namespace eos

void inieos_dev() {
  using namespace eos;
  Coef& c = g_coef;
  const double pref = C().scalar("pref", 0.0);
  c.pref = pref;
  c.ap21 = EA21 + EB21 * pref; c.ap22 = EA22 + EB22 * pref; c.ap23 = EA23 + EB23 * pref;
  c.ap24 = EA24; c.ap25 = EA25; c.ap26 = EA26;
  c.ap11 = EA11 + EB11 * pref - c.ap21 / alpha0;
  c.ap12 = EA12 + EB12 * pref - c.ap22 / alpha0;
  c.ap13 = EA13 + EB13 * pref - c.ap23 / alpha0;
  c.ap14 = EA14 - c.ap24 / alpha0; c.ap15 = EA15 - c.ap25 / alpha0; c.ap16 = EA16 - c.ap26 / alpha0;
  c.ap210 = EA21; c.ap220 = EA22; c.ap230 = EA23; c.ap240 = EA24; c.ap250 = EA25; c.ap260 = EA26;
  c.ap110 = EA11 - c.ap210 / alpha0; c.ap120 = EA12 - c.ap220 / alpha0; c.ap130 = EA13 - c.ap230 / alpha0;
  c.ap140 = EA14 - c.ap240 / alpha0; c.ap150 = EA15 - c.ap250 / alpha0; c.ap160 = EA16 - c.ap260 / alpha0;
  g_coef_set = true;
}

namespace {

constexpr int MAXTR = 4;
struct TrcPtrs { double* t[MAXTR]; double* fu[MAXTR]; double* fv[MAXTR]; int n; };

// face fluxes on i=0..ii+2, j=0..jj+2 (u: j<=jj+1; v: i<=ii+1)
__global__ void diffus_flux(Geom g, double delt1, int mm, int nn, const int* __restrict__ iu,
                            const int* __restrict__ iv, const double* __restrict__ dp,
                            const double* __restrict__ temp, const double* __restrict__ saln,
                            const double* __restrict__ difiso, const double* __restrict__ scuy,
                            const double* __restrict__ scuxi, const double* __restrict__ scvx,
                            const double* __restrict__ scvyi, double* __restrict__ usflld,
                            double* __restrict__ utflld, double* __restrict__ vsflld,
                            double* __restrict__ vtflld, double* __restrict__ usflx,
                            double* __restrict__ utflx, double* __restrict__ vsflx,
                            double* __restrict__ vtflx, TrcPtrs T, int extra /* 1: a further tracer group only */) {
  const double dpeps = 1.e-5;
  const Bid b_ = bid(g);
  const int i = b_.x * blockDim.x + threadIdx.x;  // 0..ii+2
  const int j = b_.y, k = b_.z + 1;         // j 0..jj+2
  if (i > g.ii + 2) return;
  const long x = ix2(g, i, j);
  const long xn = x + (long)(k + nn - 1) * g.lev, xm = x + (long)(k + mm - 1) * g.lev;
  const long xk = x + (long)(k - 1) * g.lev;
  const long s = g.ldi;
  // all operands are loaded unconditionally (x-1 and x-s stay inside the halo-padded arrays) so the loads
  // of a thread are issued as one batch; the masks only gate the stores
  const bool wu = j <= g.jj + 1 && iu[x] == 1, wv = i <= g.ii + 1 && iv[x] == 1;
  const double dpc = dp[xn], tc = temp[xn], sc = saln[xn], dc = difiso[xk];
  const double dpw = dp[xn - 1], tw = temp[xn - 1], sw = saln[xn - 1], dw = difiso[xk - 1];
  const double dps = dp[xn - s], ts = temp[xn - s], ss = saln[xn - s], ds = difiso[xk - s];
  const double ousf = usflx[xm], outf = utflx[xm], ovsf = vsflx[xm], ovtf = vtflx[xm];
  const double qu = delt1 * .5 * (dw + dc) * scuy[x] * scuxi[x] * fmax(fmin(dpw, dpc), dpeps);
  const double qv = delt1 * .5 * (ds + dc) * scvx[x] * scvyi[x] * fmax(fmin(dps, dpc), dpeps);
  if (wu && !extra) {
    const double fs = qu * (sw - sc), ft = qu * (tw - tc);
    usflld[xm] = fs; utflld[xm] = ft;
    usflx[xm] = ousf + fs;
    utflx[xm] = outf + ft;
  }
  if (wv && !extra) {
    const double fs = qv * (ss - sc), ft = qv * (ts - tc);
    vsflld[xm] = fs; vtflld[xm] = ft;
    vsflx[xm] = ovsf + fs;
    vtflx[xm] = ovtf + ft;
  }
  // static tracer index: keeps the pointer table in the constant bank (a run-time index spills it to local memory)
#pragma unroll
  for (int nt = 0; nt < MAXTR; ++nt)
    if (nt < T.n) {
      const double c0 = T.t[nt][xn];
      if (wu) T.fu[nt][xk] = qu * (T.t[nt][xn - 1] - c0);
      if (wv) T.fv[nt][xk] = qv * (T.t[nt][xn - s] - c0);
    }
}

__global__ void diffus_update(Geom g, eos::Coef ec, int mm, int nn, const int* __restrict__ ip,
                              const double* __restrict__ dp, double* __restrict__ temp,
                              double* __restrict__ saln, double* __restrict__ sigma,
                              const double* __restrict__ scp2, const double* __restrict__ usflld,
                              const double* __restrict__ utflld, const double* __restrict__ vsflld,
                              const double* __restrict__ vtflld, TrcPtrs T, int extra) {
  const double dpeps = 1.e-5;
  const Bid b_ = bid(g);
  const int i = b_.x * blockDim.x + threadIdx.x;  // 0..ii+1
  const int j = b_.y, k = b_.z + 1;         // 0..jj+1
  if (i > g.ii + 1) return;
  const long x = ix2(g, i, j);
  if (ip[x] != 1) return;
  const long xn = x + (long)(k + nn - 1) * g.lev, xm = x + (long)(k + mm - 1) * g.lev;
  const long xk = x + (long)(k - 1) * g.lev, s = g.ldi;
  const double q = 1. / (scp2[x] * fmax(dp[xn], dpeps));
  const double sn = saln[xn] - q * (usflld[xm + 1] - usflld[xm] + vsflld[xm + s] - vsflld[xm]);
  const double tn = temp[xn] - q * (utflld[xm + 1] - utflld[xm] + vtflld[xm + s] - vtflld[xm]);
  if (!extra) {
    saln[xn] = sn;
    temp[xn] = tn;
  }
#pragma unroll
  for (int nt = 0; nt < MAXTR; ++nt)
    if (nt < T.n)
      T.t[nt][xn] = T.t[nt][xn] - q * (T.fu[nt][xk + 1] - T.fu[nt][xk] + T.fv[nt][xk + s] - T.fv[nt][xk]);
  if (!extra) sigma[xn] = eos::sig(ec, tn, sn);
}

__global__ void tmsmt1_kernel(Geom g, int nn, bool isopyc, const int* __restrict__ ip,
                              const int* __restrict__ iu, const int* __restrict__ iv,
                              const double* __restrict__ dp, const double* __restrict__ temp,
                              const double* __restrict__ saln, double* __restrict__ dpold,
                              double* __restrict__ told, double* __restrict__ sold,
                              const double* __restrict__ trc, double* __restrict__ trcold,
                              const double* __restrict__ dpu, const double* __restrict__ dpv,
                              double* __restrict__ dpuold, double* __restrict__ dpvold) {
  const Bid b_ = bid(g);
  const int i = b_.x * blockDim.x + threadIdx.x + 1;
  const int j = b_.y + 1, k = b_.z + 1;
  if (i > g.ii) return;
  const long x = ix2(g, i, j);
  const long xn = x + (long)(k + nn - 1) * g.lev, xk = x + (long)(k - 1) * g.lev;
  if (ip[x] == 1) {
    dpold[xn] = dp[xn];
    told[xk] = temp[xn];
    sold[xk] = saln[xn];
    for (int nt = 0; nt < g.ntr; ++nt)
      trcold[xk + (long)nt * g.kdm * g.lev] = trc[xn + (long)nt * 2 * g.kdm * g.lev];
  }
  if (isopyc) {
    if (iu[x] == 1) dpuold[xk] = dpu[xn];
    if (iv[x] == 1) dpvold[xk] = dpv[xn];
  }
}

// one thread per wet column; pbfaco/pbfacn column sums in registers
__global__ void tmsmt2_kernel(Geom g, int m, int mm, int nn, const int* __restrict__ ip,
                              double* __restrict__ dp, double* __restrict__ temp, double* __restrict__ saln,
                              const double* __restrict__ dpold, const double* __restrict__ told,
                              const double* __restrict__ sold, const double* __restrict__ pb,
                              double* __restrict__ trc, const double* __restrict__ trcold) {
  const double wts1 = .875, wts2 = .0625;
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y + 1;
  if (i > g.ii) return;
  const long x = ix2(g, i, j);
  if (ip[x] != 1) return;
  double pbfaco = 0., pbfacn = 0.;
  for (int k = 1; k <= g.kdm; ++k) {
    const long xn = x + (long)(k + nn - 1) * g.lev;
    pbfaco = pbfaco + dpold[xn];
    pbfacn = pbfacn + dp[xn];
  }
  const double pbm = pb[x + (long)(m - 1) * g.lev];
  pbfaco = pbm / pbfaco;
  pbfacn = pbm / pbfacn;
  for (int k = 1; k <= g.kdm; ++k) {
    const long xn = x + (long)(k + nn - 1) * g.lev, xm = x + (long)(k + mm - 1) * g.lev;
    const long xk = x + (long)(k - 1) * g.lev;
    double pold = fmax(0., dpold[xn] * pbfaco);
    double pmid = fmax(0., dp[xm]);
    double pnew = fmax(0., dp[xn] * pbfacn);
    const double dpm = wts1 * pmid + wts2 * (pold + pnew);
    dp[xm] = dpm;
    pold = pold + epsilp; pmid = pmid + epsilp; pnew = pnew + epsilp;
    temp[xm] = (wts1 * pmid * temp[xm] + wts2 * (pold * told[xk] + pnew * temp[xn])) / (dpm + epsilp);
    saln[xm] = (wts1 * pmid * saln[xm] + wts2 * (pold * sold[xk] + pnew * saln[xn])) / (dpm + epsilp);
    for (int nt = 0; nt < g.ntr; ++nt) {
      const long o2 = (long)nt * 2 * g.kdm * g.lev, o1 = (long)nt * g.kdm * g.lev;
      trc[xm + o2] = (wts1 * pmid * trc[xm + o2] + wts2 * (pold * trcold[xk + o1] + pnew * trc[xn + o2])) /
                     (dpm + epsilp);
    }
  }
}

// p(k+1) = p(k) + dp(km) on -halo..ii+halo x -halo..jj+halo (wet columns)
__global__ void p_from_dp(Geom g, int mm, int halo, const int* __restrict__ ip, const double* __restrict__ dp,
                          double* __restrict__ p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x - halo;
  const int j = (int)blockIdx.y - halo;
  if (i > g.ii + halo) return;
  const long x = ix2(g, i, j);
  if (ip[x] != 1) return;
  double pk = p[x];
  for (int k = 1; k <= g.kdm; ++k) {
    pk = pk + dp[x + (long)(k + mm - 1) * g.lev];
    p[x + (long)k * g.lev] = pk;
  }
}

__global__ void dpuv_from_p(Geom g, int mm, const int* __restrict__ iu, const int* __restrict__ iv,
                            const double* __restrict__ p, double* __restrict__ dpu, double* __restrict__ dpv) {
  const Bid b_ = bid(g);
  const int i = b_.x * blockDim.x + threadIdx.x - 1;  // -1..ii+2
  const int j = b_.y - 1, k = b_.z + 1;    // -1..jj+2
  if (i > g.ii + 2) return;
  const long x = ix2(g, i, j), s = g.ldi;
  const long xb = x + (long)g.kdm * g.lev, x0 = x + (long)(k - 1) * g.lev, x1 = x0 + g.lev;
  const long xm = x + (long)(k + mm - 1) * g.lev;
  if (iu[x] == 1) {
    const double q = fmin(p[xb], p[xb - 1]);
    dpu[xm] = .5 * ((fmin(q, p[x1 - 1]) - fmin(q, p[x0 - 1])) + (fmin(q, p[x1]) - fmin(q, p[x0])));
  }
  if (iv[x] == 1) {
    const double q = fmin(p[xb], p[xb - s]);
    dpv[xm] = .5 * ((fmin(q, p[x1 - s]) - fmin(q, p[x0 - s])) + (fmin(q, p[x1]) - fmin(q, p[x0])));
  }
}

// passive tracers n0 .. n0+MAXTR-1 (the reference loops nt = 1..ntr inside its sweeps, phy/mod_diffus.F90:99-135;
// here the first group travels with T and S and further groups take extra launches)
TrcPtrs trc_ptrs(int n0) {
  Ctx& c = C(); const Geom& g = c.g;
  TrcPtrs T{}; T.n = std::max(0, std::min(MAXTR, g.ntr - n0));
  for (int q = 0; q < T.n; ++q) {
    T.t[q] = c.dev("trc") + (long)(n0 + q) * 2 * g.kdm * g.lev;
    T.fu[q] = c.owned("diffus_uflxtr" + std::to_string(q + 1), g.kdm);   // flux scratch is reused by every group
    T.fv[q] = c.owned("diffus_vflxtr" + std::to_string(q + 1), g.kdm);
  }
  return T;
}

}  // namespace

void diffus_dev(int m, int n, int mm, int nn, int k1m, int k1n) {
  (void)m; (void)n; (void)k1m; (void)k1n;
  Ctx& c = C(); const Geom& g = c.g;
  const long on = (long)nn * g.lev;
  const std::string ltedtp = c.option("ltedtp", "layer");
  halo_update(c.dev("dp") + on, g.kdm, 3, 3, halo_ps);
  std::vector<HaloReq> reqs{{c.dev("temp") + on, g.kdm, halo_ps}, {c.dev("saln") + on, g.kdm, halo_ps}};
  for (int nt = 0; nt < g.ntr; ++nt) reqs.push_back({c.dev("trc") + on + (long)nt * 2 * g.kdm * g.lev, g.kdm, halo_ps});
  if (ltedtp == "neutral") { halo_update(reqs, 1, 1); return; }
  if (ltedtp != "layer") throw std::runtime_error(" ltedtp = " + ltedtp + " is unsupported!");
  halo_update(reqs, 2, 2);
  for (int n0 = 0; n0 < std::max(1, g.ntr); n0 += MAXTR) {
    const TrcPtrs T = trc_ptrs(n0);
    const int extra = n0 > 0 ? 1 : 0;
    {
      const dim3 grid = lgrid(g, dim3(cdiv(g.ii + 3, 128), g.jj + 3, g.kdm));
      LAUNCH(diffus_flux, grid, 128, 0, g, c.scalar("delt1"), mm, nn, c.idev("iu"), c.idev("iv"), c.dev("dp"),
             c.dev("temp"), c.dev("saln"), c.dev("difiso"), c.dev("scuy"), c.dev("scuxi"), c.dev("scvx"),
             c.dev("scvyi"), c.dev("usflld"), c.dev("utflld"), c.dev("vsflld"), c.dev("vtflld"), c.dev("usflx"),
             c.dev("utflx"), c.dev("vsflx"), c.dev("vtflx"), T, extra);
    }
    {
      const dim3 grid = lgrid(g, dim3(cdiv(g.ii + 2, 128), g.jj + 2, g.kdm));
      LAUNCH(diffus_update, grid, 128, 0, g, eos::host_coef(), mm, nn, c.idev("ip"), c.dev("dp"), c.dev("temp"),
             c.dev("saln"), c.dev("sigma"), c.dev("scp2"), c.dev("usflld"), c.dev("utflld"), c.dev("vsflld"),
             c.dev("vtflld"), T, extra);
    }
  }
}

void tmsmt1_dev(int nn) {
  Ctx& c = C(); const Geom& g = c.g;
  const bool isopyc = c.option("vcoord", "cntiso_hybrid") == "isopyc_bulkml";
  // pure copy of three fields with two masks: nothing 2-D worth keeping in cache, and the plane order
  // measured faster (0.79 vs 0.89 ms at tnx0.25v4)
  Geom gp = g; gp.lf = 0;
  const dim3 grid(cdiv(g.ii, 128), g.jj, g.kdm);
  LAUNCH(tmsmt1_kernel, grid, 128, 0, gp, nn, isopyc, c.idev("ip"), c.idev("iu"), c.idev("iv"), c.dev("dp"),
         c.dev("temp"), c.dev("saln"), c.dev("dpold"), c.dev("told"), c.dev("sold"),
         g.ntr ? c.dev("trc") : nullptr, g.ntr ? c.dev("trcold") : nullptr, isopyc ? c.dev("dpu") : nullptr,
         isopyc ? c.dev("dpv") : nullptr, isopyc ? c.dev("dpuold") : nullptr, isopyc ? c.dev("dpvold") : nullptr);
}

void tmsmt2_dev(int m, int mm, int nn, int k1m) {
  (void)k1m;
  Ctx& c = C(); const Geom& g = c.g;
  {
    dim3 grid(cdiv(g.ii, 128), g.jj);
    LAUNCH(tmsmt2_kernel, grid, 128, 0, g, m, mm, nn, c.idev("ip"), c.dev("dp"), c.dev("temp"), c.dev("saln"),
           c.dev("dpold"), c.dev("told"), c.dev("sold"), c.dev("pb"), g.ntr ? c.dev("trc") : nullptr,
           g.ntr ? c.dev("trcold") : nullptr);
  }
  halo_update(c.dev("dp") + (long)mm * g.lev, g.kdm, 3, 3, halo_ps);
  {
    dim3 grid(cdiv(g.ii + 5, 128), g.jj + 5);
    LAUNCH(p_from_dp, grid, 128, 0, g, mm, 2, c.idev("ip"), c.dev("dp"), c.dev("p"));
  }
  if (c.option("vcoord", "cntiso_hybrid") == "isopyc_bulkml") {
    const dim3 grid = lgrid(g, dim3(cdiv(g.ii + 4, 128), g.jj + 4, g.kdm));
    LAUNCH(dpuv_from_p, grid, 128, 0, g, mm, c.idev("iu"), c.idev("iv"), c.dev("p"), c.dev("dpu"), c.dev("dpv"));
  }
}

}  // namespace blom
