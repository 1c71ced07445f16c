// numerical_bounds (phy/mod_blom_init.F90:446-555) and init_fluxes
// (phy/mod_state.F90:341-383): setup / per-step zero-fill kernels; budget_init / budget_sums
// (phy/mod_budget.F90:74-196): thickness-weighted column sums feeding the strip-ordered xcsum.
#include "common.cuh"

namespace blom {

namespace {

__global__ void nb_difmx(Geom g, double baclin, const double* __restrict__ scpx, const double* __restrict__ scpy,
                         const double* __restrict__ scqx, const double* __restrict__ scqy,
                         double* __restrict__ difmxp, double* __restrict__ difmxq) {
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= g.lev) return;
  double dx2 = scpx[t] * scpx[t], dy2 = scpy[t] * scpy[t];
  difmxp[t] = .9 * .5 * dx2 * dy2 / fmax(1., (dx2 + dy2) * (baclin + baclin));
  dx2 = scqx[t] * scqx[t]; dy2 = scqy[t] * scqy[t];
  difmxq[t] = .9 * .5 * dx2 * dy2 / fmax(1., (dx2 + dy2) * (baclin + baclin));
}

__global__ void nb_umax(Geom g, double baclin, const int* __restrict__ ip, const int* __restrict__ iu,
                        const int* __restrict__ iv, const double* __restrict__ scp2,
                        const double* __restrict__ scuy, const double* __restrict__ scvx,
                        const double* __restrict__ scpx, const double* __restrict__ scpy,
                        const double* __restrict__ depths, double* __restrict__ umax,
                        double* __restrict__ vmax, double* __restrict__ btdt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.ii) return;
  const long x = ix2(g, i, j);
  if (iu[x] == 1) umax[x] = .9 * .125 * fmin(scp2[x - 1], scp2[x]) / (scuy[x] * baclin);
  if (iv[x] == 1) vmax[x] = .9 * .125 * fmin(scp2[x - g.ldi], scp2[x]) / (scvx[x] * baclin);
  // CFL estimate of the barotropic step; land points carry the neutral value
  btdt[x] = ip[x] == 1 ? scpx[x] * scpy[x] / sqrt(grav * depths[x] * (scpx[x] * scpx[x] + scpy[x] * scpy[x]))
                       : 86400.;
}

__global__ void zero_fluxes(Geom g, int mm, const int* __restrict__ iu, const int* __restrict__ iv,
                            double* __restrict__ uflx, double* __restrict__ utflx, double* __restrict__ usflx,
                            double* __restrict__ vflx, double* __restrict__ vtflx, double* __restrict__ vsflx) {
  const Bid b_ = bid(g);
  const int i = b_.x * blockDim.x + threadIdx.x;  // 0..ii+2
  const int j = b_.y, k = b_.z + 1;         // 0..jj+2
  if (i > g.ii + 2) return;
  const long x = ix2(g, i, j), xm = x + (long)(k + mm - 1) * g.lev;
  if (iu[x] == 1) { uflx[xm] = 0.; utflx[xm] = 0.; usflx[xm] = 0.; }
  if (iv[x] == 1) { vflx[xm] = 0.; vtflx[xm] = 0.; vsflx[xm] = 0.; }
}

// column sums in k order (phy/mod_budget.F90:121-139,161-172): util1 = sum_k a(kn)*dp(kn)*scp2 (+ util2 for b)
__global__ void budget_columns(Geom g, int nn, const int* __restrict__ ip, const double* __restrict__ dp,
                               const double* __restrict__ scp2, const double* __restrict__ a,
                               const double* __restrict__ b, double* __restrict__ util1, double* __restrict__ util2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.ii) return;
  const long x = ix2(g, i, j);
  if (ip[x] != 1) return;
  const double area = scp2[x];
  double s1 = 0., s2 = 0.;
  for (int k = 1; k <= g.kdm; ++k) {
    const long xn = x + (long)(k + nn - 1) * g.lev;
    const double q = dp[xn] * area;
    s1 = s1 + a[xn] * q;
    if (b) s2 = s2 + b[xn] * q;
  }
  util1[x] = s1;
  if (b) util2[x] = s2;
}
// util1 = f(:,:)*scp2 on wet points (budget_init with f=pb(:,:,1); the salt_corr sum of budget_sums)
__global__ void budget_area_weight(Geom g, const int* __restrict__ ip, const double* __restrict__ f,
                                   const double* __restrict__ scp2, double* __restrict__ util1) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.ii) return;
  const long x = ix2(g, i, j);
  if (ip[x] == 1) util1[x] = f[x] * scp2[x];
}

}  // namespace

double budget_init_dev() {
  Ctx& c = C(); const Geom& g = c.g;
  double* util1 = c.has("util1") ? c.dev("util1") : c.owned("util1", 1);
  dim3 grid(cdiv(g.ii, 128), g.jj);
  LAUNCH(budget_area_weight, grid, 128, 0, g, c.idev("ip"), c.dev("pb"), c.dev("scp2"), util1);
  return xcsum_dev(util1, c.idev("ip"));
}

void budget_sums_dev(int ncall, int n, int nn, double* out) {
  (void)n;
  Ctx& c = C(); const Geom& g = c.g;
  double* util1 = c.has("util1") ? c.dev("util1") : c.owned("util1", 1);
  double* util2 = c.has("util2") ? c.dev("util2") : c.owned("util2", 1);
  const int* ip = c.idev("ip");
  dim3 grid(cdiv(g.ii, 128), g.jj);
  LAUNCH(budget_columns, grid, 128, 0, g, nn, ip, c.dev("dp"), c.dev("scp2"), c.dev("saln"), c.dev("temp"), util1, util2);
  out[0] = xcsum_dev(util1, ip);
  out[1] = xcsum_dev(util2, ip);
  if (g.ntr > 0) {
    LAUNCH(budget_columns, grid, 128, 0, g, nn, ip, c.dev("dp"), c.dev("scp2"), c.dev("trc"), (const double*)nullptr,
           util1, util2);
    out[2] = xcsum_dev(util1, ip);
  }
  const bool isopyc = c.option("vcoord", "cntiso_hybrid") == "isopyc_bulkml";
  if (((isopyc && ncall == 5) || (!isopyc && ncall == 4)) && c.has("salt_corr")) {
    LAUNCH(budget_area_weight, grid, 128, 0, g, ip, c.dev("salt_corr"), c.dev("scp2"), util1);
    out[3] = xcsum_dev(util1, ip);
  }
}

void numerical_bounds_dev() {
  Ctx& c = C(); const Geom& g = c.g;
  const double baclin = c.scalar("baclin");
  LAUNCH(nb_difmx, cdiv(g.lev, 256), 256, 0, g, baclin, c.dev("scpx"), c.dev("scpy"), c.dev("scqx"), c.dev("scqy"),
         c.dev("difmxp"), c.dev("difmxq"));
  double* btdt = c.owned("_btdt", 1);
  dim3 grid(cdiv(g.ii, 128), g.jj);
  LAUNCH(nb_umax, grid, 128, 0, g, baclin, c.idev("ip"), c.idev("iu"), c.idev("iv"), c.dev("scp2"), c.dev("scuy"),
         c.dev("scvx"), c.dev("scpx"), c.dev("scpy"), c.dev("depths"), c.dev("umax"), c.dev("vmax"), btdt);
  halo_update(std::vector<HaloReq>{{c.dev("umax"), 1, halo_us}, {c.dev("vmax"), 1, halo_vs}}, g.nb, g.nb);
  // xcmin over wet points (phy/mod_blom_init.F90:483-497); reported like the reference does
  c.sc["btdtmx"] = fmin(86400., xcmax_dev(btdt, c.idev("ip"), false)) / sqrt(2.);
}

void init_fluxes_dev(int m, int n, int mm, int nn, int k1m, int k1n) {
  (void)m; (void)n; (void)k1m; (void)k1n;
  Ctx& c = C(); const Geom& g = c.g;
  const dim3 grid = lgrid(g, dim3(cdiv(g.ii + 3, 128), g.jj + 3, g.kdm));
  LAUNCH(zero_fluxes, grid, 128, 0, g, mm, c.idev("iu"), c.idev("iv"), c.dev("uflx"), c.dev("utflx"), c.dev("usflx"),
         c.dev("vflx"), c.dev("vtflx"), c.dev("vsflx"));
  const long on = (long)nn * g.lev;
  halo_update(std::vector<HaloReq>{{c.dev("uflx") + on, g.kdm, halo_uv}, {c.dev("utflx") + on, g.kdm, halo_uv},
                                   {c.dev("usflx") + on, g.kdm, halo_uv}, {c.dev("vflx") + on, g.kdm, halo_vv},
                                   {c.dev("vtflx") + on, g.kdm, halo_vv}, {c.dev("vsflx") + on, g.kdm, halo_vv}}, 1, 1);
}

}  // namespace blom
