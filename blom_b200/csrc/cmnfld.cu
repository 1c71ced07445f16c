// Producers of the neutral slope consumed by eddtra / ndiff, hybrid (ALE) branch
// of cmnfld2 (phy/mod_cmnfld_routines.F90:1158-1238):
//   cmnfld_bfsqf_ale   (:229-350)  buoyancy frequency squared on interfaces / layers + vertically filtered
//   cmnfld_nslope_ale  (:654-811)  interface geopotential, neutral slope vector and slope x N
//   cmnfld_nnslope_ale (:813-883)  slope x N where the slope is already known (ltedtp='neutral')
//
// B200 design: all three are column recurrences with no horizontal coupling beyond one face neighbour,
// so the layout is one thread per (p|u|v) column with lanes along i: `a[x + k*lev]` is a coalesced row
// segment per warp and level.  The tridiagonal filter keeps its three per-column work vectors
// (delp, rhs, gam) thread-local instead of the reference's eight; the face kernels carry the upper
// layer's T,S of both cells in registers so every layer value is loaded once per face, and fuse the
// zero fill, the kmax scan, the slope and the knnsl extrapolation of the reference's three loops.
#include "common.cuh"
#include "eos.cuh"

namespace blom {

namespace {

constexpr int KM = 64;  // compile-time bound on kdm for thread-local column vectors

// cmnfld_bfsqf_ale (:229-350): p-columns on -1..ii+2 x -1..jj+2
__global__ void __launch_bounds__(128)
cf_bfsq_column(Geom g, int nn, double sls0, double bfsqmn, const int* __restrict__ ip,
               const double* __restrict__ p, const double* __restrict__ dp, const double* __restrict__ temp,
               const double* __restrict__ saln, double* __restrict__ bfsqi, double* __restrict__ bfsql,
               double* __restrict__ bfsqf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x - 1;
  const int j = (int)blockIdx.y - 1;
  if (i > g.ii + 2) return;
  const long x = ix2(g, i, j), L = g.lev;
  if (ip[x] != 1) return;
  const int kk = g.kdm;
  double delp[KM + 1], bfsq[KM + 1], gam[KM + 1], bi[KM + 2];
  const double pbot = p[x + (long)kk * L];
  const double sls2 = sls0 * sls0;
  bi[1] = bfsqmn;
  double pk = p[x + L];  // p(k), k=2 first
  double pup = .5 * (p[x] + pk);
  double tup = temp[x + (long)nn * L], sup = saln[x + (long)nn * L];
  for (int k = 2; k <= kk; ++k) {
    const double pk1 = p[x + (long)k * L];  // p(k+1)
    if (pbot - pk < epsilp) {
      delp[k] = onemm;
      bi[k] = bi[k - 1];
      bfsq[k] = bfsqmn;
    } else {
      const double plo = (pbot - pk1 < epsilp) ? pbot : .5 * (pk + pk1);
      const double tlo = temp[x + (long)(k + nn - 1) * L], slo = saln[x + (long)(k + nn - 1) * L];
      const double dk = fmax(onemm, plo - pup);
      delp[k] = dk;
      double b = grav * grav * (eos::rho(pk, tlo, slo) - eos::rho(pk, tup, sup)) / dk;
      bfsq[k] = fmax(bfsqmn, b);
      b = b * dk / fmax(onem, dk);
      if (pbot - pk < onem) b = bi[k - 1];
      bi[k] = b;
      pup = plo; tup = tlo; sup = slo;
    }
    pk = pk1;
  }
  delp[1] = dp[x + (long)nn * L];
  bi[1] = bi[2];
  bfsq[1] = fmax(bfsqmn, bi[1]);
  bi[kk + 1] = bi[kk];
  for (int k = 1; k <= kk + 1; ++k) bfsqi[x + (long)(k - 1) * L] = bi[k];
  for (int k = 1; k <= kk - 1; ++k) bfsql[x + (long)(k - 1) * L] = .5 * (bi[k] + bi[k + 1]);
  bfsql[x + (long)(kk - 1) * L] = bi[kk];
  // implicit vertical diffusion of bfsq: tridiagonal coefficients (:300-314) and solve (:317-325);
  // the solution overwrites bfsq[]
  double ctd_prev = -2. * sls2 / (delp[1] * (delp[1] + delp[2]));
  double bei = 1. / (1. - ctd_prev);
  double fprev = bfsq[1] * bei;
  bfsq[1] = fprev;
  for (int k = 2; k <= kk; ++k) {
    const double atd = -2. * sls2 / (delp[k] * (delp[k - 1] + delp[k]));
    double btd, ctd = 0.;
    if (k < kk) {
      ctd = -2. * sls2 / (delp[k] * (delp[k] + delp[k + 1]));
      btd = 1. - atd - ctd;
    } else {
      btd = 1. - atd;
    }
    const double gk = ctd_prev * bei;
    gam[k] = gk;
    bei = 1. / (btd - atd * gk);
    fprev = (bfsq[k] - atd * fprev) * bei;
    bfsq[k] = fprev;
    ctd_prev = ctd;
  }
  bfsqf[x + (long)kk * L] = fprev;        // bfsqf(kk+1) = bfsqf(kk)
  bfsqf[x + (long)(kk - 1) * L] = fprev;
  for (int k = kk - 1; k >= 1; --k) {
    fprev = bfsq[k] - gam[k + 1] * fprev;
    bfsqf[x + (long)(k - 1) * L] = fprev;
  }
}

// interface geopotential (:669-685) on -1..ii+2 x -1..jj+2, bottom-up
__global__ void cf_phi_column(Geom g, int nn, const int* __restrict__ ip, const double* __restrict__ p,
                              const double* __restrict__ dp, const double* __restrict__ temp,
                              const double* __restrict__ saln, double* __restrict__ phi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x - 1;
  const int j = (int)blockIdx.y - 1;
  if (i > g.ii + 2) return;
  const long x = ix2(g, i, j), L = g.lev;
  if (ip[x] != 1) return;
  const int kk = g.kdm;
  double ph = phi[x + (long)kk * L], pk1 = p[x + (long)kk * L];
  for (int k = kk; k >= 1; --k) {
    const long xn = x + (long)(k + nn - 1) * L;
    const double pk = p[x + (long)(k - 1) * L];
    if (!(dp[xn] < epsilp)) ph = ph - eos::p_alpha(pk1, pk, temp[xn], saln[xn]);
    phi[x + (long)(k - 1) * L] = ph;
    pk1 = pk;
  }
}

// slope vector component at the faces (:696-747 x, :751-798 y).  DIR 0: u faces, i 0..ii+2, j -1..jj+2;
// DIR 1: v faces, i -1..ii+2, j 0..jj+2.
template <int DIR>
__global__ void __launch_bounds__(128)
cf_nslope_face(Geom g, int nn, const int* __restrict__ msk, const double* __restrict__ p,
               const double* __restrict__ dp, const double* __restrict__ temp, const double* __restrict__ saln,
               const double* __restrict__ phi, const double* __restrict__ bfsqf, const double* __restrict__ sci,
               double* __restrict__ nslp, double* __restrict__ nnslp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x - (DIR == 0 ? 0 : 1);
  const int j = (int)blockIdx.y - (DIR == 0 ? 1 : 0);
  if (i > g.ii + 2) return;
  const long x = ix2(g, i, j), L = g.lev, xo = x - (DIR == 0 ? 1 : g.ldi);
  if (msk[x] != 1) return;
  const int kk = g.kdm;
  int kmax = 1;
  for (int k = 2; k <= kk; ++k) {
    const long kn = (long)(k + nn - 1) * L;
    if (dp[xo + kn] > epsilp || dp[x + kn] > epsilp) kmax = k;
  }
  const double phb_c = phi[x + (long)kk * L], phb_o = phi[xo + (long)kk * L], sc = sci[x];
  nslp[x] = 0.; nnslp[x] = 0.;
  double tc0 = temp[x + (long)nn * L], sc0 = saln[x + (long)nn * L];
  double to0 = temp[xo + (long)nn * L], so0 = saln[xo + (long)nn * L];
  int knnsl = 2;
  double fill = 0.;
  for (int k = 2; k <= kk; ++k) {
    const long xk = (long)(k - 1) * L;
    double ns = 0., nns = 0.;
    if (k <= kmax) {
      const long kn = (long)(k + nn - 1) * L;
      const double tc1 = temp[x + kn], sc1 = saln[x + kn], to1 = temp[xo + kn], so1 = saln[xo + kn];
      const double pm = .5 * (p[xo + xk] + p[x + xk]);
      const double rho_d = .5 * (eos::rho(pm, tc0, sc0) - eos::rho(pm, to0, so0) + eos::rho(pm, tc1, sc1) -
                                 eos::rho(pm, to1, so1));
      const double ph_c = phi[x + xk], ph_o = phi[xo + xk];
      const double phi_d = ph_c - ph_o;
      const double bfsqm = .5 * (bfsqf[xo + xk] + bfsqf[x + xk]);
      ns = (grav * rho_d / (rho0 * bfsqm) + phi_d / grav) * sc;
      if (ph_c > phb_o && ph_o > phb_c) {
        nns = sqrt(bfsqm) * ns;
        knnsl = k;
        fill = nns;
      }
      tc0 = tc1; sc0 = sc1; to0 = to1; so0 = so1;
    }
    nslp[x + xk] = ns;
    nnslp[x + xk] = nns;
  }
  for (int k = knnsl + 1; k <= kmax; ++k) nnslp[x + (long)(k - 1) * L] = fill;
}

// cmnfld_nnslope_ale (:826-870): same face ranges
template <int DIR>
__global__ void cf_nnslope_face(Geom g, const int* __restrict__ msk, const double* __restrict__ p,
                                const double* __restrict__ bfsqf, const double* __restrict__ nslp,
                                double* __restrict__ nnslp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x - (DIR == 0 ? 0 : 1);
  const int j = (int)blockIdx.y - (DIR == 0 ? 1 : 0);
  if (i > g.ii + 2) return;
  const long x = ix2(g, i, j), L = g.lev, xo = x - (DIR == 0 ? 1 : g.ldi);
  if (msk[x] != 1) return;
  const int kk = g.kdm;
  const double pb_c = p[x + (long)kk * L], pb_o = p[xo + (long)kk * L];
  double fill = 0.;
  nnslp[x] = 0.;
  int k = 2;
  for (; k <= kk; ++k) {
    const long xk = (long)(k - 1) * L;
    if (p[x + xk] < pb_o && p[xo + xk] < pb_c) {
      const double bfsqm = .5 * (bfsqf[xo + xk] + bfsqf[x + xk]);
      fill = sqrt(bfsqm) * nslp[x + xk];
      nnslp[x + xk] = fill;
    } else {
      break;
    }
  }
  for (; k <= kk; ++k) nnslp[x + (long)(k - 1) * L] = fill;
}

void check_kdm(const Geom& g) {
  if (g.kdm > KM) throw std::runtime_error("cmnfld: kdm exceeds the compiled column bound (64)");
  if (g.kdm < 2) throw std::runtime_error("cmnfld: kdm >= 2 required");
}

}  // namespace

void cmnfld_bfsqf_ale_dev(int m, int n, int mm, int nn, int k1m, int k1n) {
  (void)m; (void)n; (void)mm; (void)k1m; (void)k1n;
  Ctx& c = C(); const Geom& g = c.g;
  check_kdm(g);
  const double sls0 = c.scalar("sls0", 10. * onem), bfsqmn = c.scalar("bfsqmn", 1.e-7);
  double* bfsqi = c.dev("bfsqi"); double* bfsql = c.dev("bfsql");
  if (c.nlev("bfsqi") < g.kdm + 1 || c.nlev("bfsqf") < g.kdm + 1 || c.nlev("bfsql") < g.kdm)
    throw std::runtime_error("cmnfld_bfsqf_ale: bfsqi/bfsqf need kdm+1 levels, bfsql kdm");
  CUDA_CHECK(cudaMemsetAsync(bfsqi, 0, sizeof(double) * g.lev * (size_t)(g.kdm + 1), c.stream));  // :247
  CUDA_CHECK(cudaMemsetAsync(bfsql, 0, sizeof(double) * g.lev * (size_t)g.kdm, c.stream));        // :248
  dim3 grid(cdiv(g.ii + 4, 128), g.jj + 4);
  LAUNCH(cf_bfsq_column, grid, 128, 0, g, nn, sls0, bfsqmn, c.idev("ip"), c.dev("p"), c.dev("dp"), c.dev("temp"),
         c.dev("saln"), bfsqi, bfsql, c.dev("bfsqf"));
}

void cmnfld_nslope_ale_dev(int m, int n, int mm, int nn, int k1m, int k1n) {
  (void)m; (void)n; (void)mm; (void)k1m; (void)k1n;
  Ctx& c = C(); const Geom& g = c.g;
  check_kdm(g);
  { dim3 grid(cdiv(g.ii + 4, 128), g.jj + 4);
    LAUNCH(cf_phi_column, grid, 128, 0, g, nn, c.idev("ip"), c.dev("p"), c.dev("dp"), c.dev("temp"), c.dev("saln"),
           c.dev("phi")); }
  { dim3 grid(cdiv(g.ii + 3, 128), g.jj + 4);
    LAUNCH_NAMED("cf_nslope_face<u>", cf_nslope_face<0>, grid, 128, 0, g, nn, c.idev("iu"), c.dev("p"), c.dev("dp"),
                 c.dev("temp"), c.dev("saln"), c.dev("phi"), c.dev("bfsqf"), c.dev("scuxi"), c.dev("nslpx"),
                 c.dev("nnslpx")); }
  { dim3 grid(cdiv(g.ii + 4, 128), g.jj + 3);
    LAUNCH_NAMED("cf_nslope_face<v>", cf_nslope_face<1>, grid, 128, 0, g, nn, c.idev("iv"), c.dev("p"), c.dev("dp"),
                 c.dev("temp"), c.dev("saln"), c.dev("phi"), c.dev("bfsqf"), c.dev("scvyi"), c.dev("nslpy"),
                 c.dev("nnslpy")); }
}

void cmnfld_nnslope_ale_dev(int m, int n, int mm, int nn, int k1m, int k1n) {
  (void)m; (void)n; (void)mm; (void)nn; (void)k1m; (void)k1n;
  Ctx& c = C(); const Geom& g = c.g;
  halo_update(std::vector<HaloReq>{HaloReq{c.dev("nslpx"), g.kdm, halo_uv}, HaloReq{c.dev("nslpy"), g.kdm, halo_vv}},
              2, 2);
  { dim3 grid(cdiv(g.ii + 3, 128), g.jj + 4);
    LAUNCH_NAMED("cf_nnslope_face<u>", cf_nnslope_face<0>, grid, 128, 0, g, c.idev("iu"), c.dev("p"), c.dev("bfsqf"),
                 c.dev("nslpx"), c.dev("nnslpx")); }
  { dim3 grid(cdiv(g.ii + 4, 128), g.jj + 3);
    LAUNCH_NAMED("cf_nnslope_face<v>", cf_nnslope_face<1>, grid, 128, 0, g, c.idev("iv"), c.dev("p"), c.dev("bfsqf"),
                 c.dev("nslpy"), c.dev("nnslpy")); }
}

// cmnfld2 (:1158-1238), vcoord /= 'isopyc_bulkml'
void cmnfld2_dev(int m, int n, int mm, int nn, int k1m, int k1n) {
  Ctx& c = C(); const Geom& g = c.g;
  if (c.option("vcoord", "cntiso_hybrid") == "isopyc_bulkml")
    throw std::runtime_error(" cmnfld2: vcoord = isopyc_bulkml is unsupported!");
  halo_update(std::vector<HaloReq>{HaloReq{c.dev("temp"), 2 * g.kdm, halo_ps}, HaloReq{c.dev("saln"), 2 * g.kdm, halo_ps}},
              3, 3);
  cmnfld_bfsqf_ale_dev(m, n, mm, nn, k1m, k1n);
  if (c.option("edritp", "large scale") == "large scale" || c.option("eitmth", "gm") == "gm") {
    if (c.option("ltedtp", "layer") == "neutral") cmnfld_nnslope_ale_dev(m, n, mm, nn, k1m, k1n);
    else cmnfld_nslope_ale_dev(m, n, mm, nn, k1m, k1n);
  }
}

}  // namespace blom
