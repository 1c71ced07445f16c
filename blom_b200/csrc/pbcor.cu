// Baroclinic mass flux correction: pbcor1 (phy/mod_pbcor.F90:66-412) and pbcor2
// (:416-743), bmcmth 'uc' (upstream column) and 'dluc' (depth limited upstream
// column).  dpeps1=1e-5, dpeps2=1e-7 (:58-59).
//
// The reference loops k outside three masked 2-D sweeps (u faces, v faces, cell
// update) that communicate through the 2-D work arrays uflux,uflux2,uflux3,
// uflxtr of mod_utility.  Every level is in fact independent (the flux of level
// k depends only on the 2-D residual, the pre-update state of level k and the
// column total), so the GPU form is three launches over all levels at once:
//   pbcor_prep    per column: (pbcor2: dp=max(0,dp)+epsilp) p(k+1)=p(k)+dp on the
//                 1-wide ring; per face: residual utot=dlt*ub - sum_k uflx
//   pbcor_update  per (i,j,k): the four face fluxes of the cell are recomputed in
//                 registers (bit-identical on both sides of a face: products are
//                 taken with __dmul_rn so FMA contraction cannot split them),
//                 own faces accumulated into uflx/usflx/utflx, new dp/T/S/trc
//                 written to the ping-pong set shared with cppm (no in-place
//                 hazard with neighbours reading the pre-update state)
//   pbcor_finish  per column: p rebuilt, column rescaled to the barotropic bottom
//                 pressure, T/S/trc moved back
// The 2-D work arrays never exist; the residuals utotm/vtotm (pbcor1) and
// utotn/vtotn (pbcor2) are written like the reference does.
#include "common.cuh"
#include "eos.cuh"

namespace blom {

namespace {

constexpr int MAXTR = 4;
struct PbTr { double* t[MAXTR]; double* tb[MAXTR]; int n; };

template <int WHICH>
__global__ void pbcor_prep(Geom g, double dlt, int ks, int kf, int lt, const int* __restrict__ ip,
                           const int* __restrict__ iu, const int* __restrict__ iv, double* __restrict__ dp,
                           double* __restrict__ p, const double* __restrict__ ubt, const double* __restrict__ vbt,
                           const double* __restrict__ uflx, const double* __restrict__ vflx,
                           double* __restrict__ utot, double* __restrict__ vtot) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;  // 0..ii+1, 0..jj+1
  if (i > g.ii + 1) return;
  const long x = ix2(g, i, j), lev = g.lev;
  const int kk = g.kdm;
  if (ip[x] == 1) {
    double pk = p[x];
    for (int k = 1; k <= kk; ++k) {
      double d = dp[x + (long)(k + ks - 1) * lev];
      if (WHICH == 2) { d = fmax(0., d) + epsilp; dp[x + (long)(k + ks - 1) * lev] = d; }
      pk = pk + d;
      p[x + (long)k * lev] = pk;
    }
  }
  if (i >= 1 && j >= 1 && j <= g.jj && iu[x] == 1) {
    double t = dlt * ubt[x + (long)(lt - 1) * lev];
    for (int k = 1; k <= kk; ++k) t = t - uflx[x + (long)(k + kf - 1) * lev];
    utot[x] = t;
  }
  if (i >= 1 && i <= g.ii && j >= 1 && iv[x] == 1) {
    double t = dlt * vbt[x + (long)(lt - 1) * lev];
    for (int k = 1; k <= kk; ++k) t = t - vflx[x + (long)(k + kf - 1) * lev];
    vtot[x] = t;
  }
}

struct Flux3 { double f, f2, f3, ftr[MAXTR]; };

// flux through the face at x whose "minus" cell is x-s (mod_pbcor.F90:166-187 'uc', :240-265 'dluc')
template <bool DLUC>
__device__ __forceinline__ Flux3 face_flux(long x, long s, int k, int kk, long lev, long ol /* state level offset */,
                                           const int* __restrict__ mask, const double* __restrict__ tot,
                                           const double* __restrict__ dp, const double* __restrict__ p,
                                           const double* __restrict__ temp, const double* __restrict__ saln,
                                           const PbTr& T) {
  Flux3 r{};
  if (mask[x] != 1) return r;
  const double t = tot[x];
  const long up = t > 0. ? x - s : x;
  if (!DLUC) {
    r.f = __dmul_rn(t, dp[up + ol]) / p[up + (long)kk * lev];
  } else {
    const double pbt = fmin(p[x + (long)kk * lev], p[x - s + (long)kk * lev]);
    r.f = __dmul_rn(t, fmax(0., fmin(pbt, p[up + (long)k * lev]) - p[up + (long)(k - 1) * lev])) / pbt;
  }
  r.f2 = __dmul_rn(r.f, saln[up + ol]);
  r.f3 = __dmul_rn(r.f, temp[up + ol]);
  for (int nt = 0; nt < T.n; ++nt) r.ftr[nt] = __dmul_rn(r.f, T.t[nt][up + ol]);
  return r;
}

template <int WHICH, bool DLUC>
__global__ void __launch_bounds__(128)
pbcor_update(Geom g, eos::Coef ec, int ks, int kf, const int* __restrict__ ip, const int* __restrict__ iu,
             const int* __restrict__ iv, const double* __restrict__ utot, const double* __restrict__ vtot,
             const double* __restrict__ dp, const double* __restrict__ p, const double* __restrict__ temp,
             const double* __restrict__ saln, const double* __restrict__ scp2i, double* __restrict__ uflx,
             double* __restrict__ usflx, double* __restrict__ utflx, double* __restrict__ vflx,
             double* __restrict__ vsflx, double* __restrict__ vtflx, double* __restrict__ dpB,
             double* __restrict__ tB, double* __restrict__ sB, double* __restrict__ sigma, PbTr T) {
  const double dpeps1 = 1.e-5, dpeps2 = 1.e-7;
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1, k = blockIdx.z + 1;  // 1..ii+1, 1..jj+1
  if (i > g.ii + 1) return;
  const long x = ix2(g, i, j), lev = g.lev, s = g.ldi;
  const int kk = g.kdm;
  const long ol = (long)(k + ks - 1) * lev, oa = (long)(k + kf - 1) * lev, ok = (long)(k - 1) * lev;
  Flux3 fu{}, fv{};
  if (j <= g.jj) {
    fu = face_flux<DLUC>(x, 1, k, kk, lev, ol, iu, utot, dp, p, temp, saln, T);
    if (iu[x] == 1) {
      uflx[x + oa] = uflx[x + oa] + fu.f;
      usflx[x + oa] = usflx[x + oa] + fu.f2;
      utflx[x + oa] = utflx[x + oa] + fu.f3;
    }
  }
  if (i <= g.ii) {
    fv = face_flux<DLUC>(x, s, k, kk, lev, ol, iv, vtot, dp, p, temp, saln, T);
    if (iv[x] == 1) {
      vflx[x + oa] = vflx[x + oa] + fv.f;
      vsflx[x + oa] = vsflx[x + oa] + fv.f2;
      vtflx[x + oa] = vtflx[x + oa] + fv.f3;
    }
  }
  if (i > g.ii || j > g.jj || ip[x] != 1) return;
  const Flux3 fue = face_flux<DLUC>(x + 1, 1, k, kk, lev, ol, iu, utot, dp, p, temp, saln, T);
  const Flux3 fvn = face_flux<DLUC>(x + s, s, k, kk, lev, ol, iv, vtot, dp, p, temp, saln, T);
  const double a = scp2i[x];
  double dpo = dp[x + ol], dpn, dpni;
  const double dm = fue.f - fu.f + fvn.f - fv.f;
  const double ds = fue.f2 - fu.f2 + fvn.f2 - fv.f2;
  const double dt = fue.f3 - fu.f3 + fvn.f3 - fv.f3;
  if (WHICH == 1) {
    dpn = fmax(0., dpo - dm * a);
    dpo = dpo + dpeps1;
    dpni = 1. / (dpn + dpeps1);
  } else {
    dpn = dpo - a * dm;
    dpni = 1. / dpn;
  }
  const double sn = (dpo * saln[x + ol] - ds * a) * dpni;
  const double tn = (dpo * temp[x + ol] - dt * a) * dpni;
  for (int nt = 0; nt < T.n; ++nt) {
    const double dtr = fue.ftr[nt] - fu.ftr[nt] + fvn.ftr[nt] - fv.ftr[nt];
    T.tb[nt][x + ok] = (dpo * T.t[nt][x + ol] - dtr * a) * dpni;
  }
  if (WHICH == 2) {
    sigma[x + ol] = eos::sig(ec, tn, sn);
    dpn = dpn - epsilp;
  }
  if (dpn < dpeps2) dpn = 0.;
  dpB[x + ok] = dpn; tB[x + ok] = tn; sB[x + ok] = sn;
}

template <int WHICH>
__global__ void pbcor_finish(Geom g, int ks, const int* __restrict__ ip, const double* __restrict__ pbt,
                             const double* __restrict__ dpB, const double* __restrict__ tB,
                             const double* __restrict__ sB, double* __restrict__ dp, double* __restrict__ temp,
                             double* __restrict__ saln, double* __restrict__ p, PbTr T) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.ii) return;
  const long x = ix2(g, i, j), lev = g.lev;
  if (ip[x] != 1) return;
  const int kk = g.kdm;
  const double p1 = p[x];
  double pk = p1;
  for (int k = 1; k <= kk; ++k) {
    pk = pk + dpB[x + (long)(k - 1) * lev];
    if (WHICH == 1) p[x + (long)k * lev] = pk;
  }
  const double pbfac = pbt[x] / pk;
  pk = p1;
  for (int k = 1; k <= kk; ++k) {
    const long ok = (long)(k - 1) * lev, ol = (long)(k + ks - 1) * lev;
    const double d = dpB[x + ok] * pbfac;
    dp[x + ol] = d;
    if (WHICH == 2) { pk = pk + d; p[x + (long)k * lev] = pk; }
    temp[x + ol] = tB[x + ok];
    saln[x + ol] = sB[x + ok];
    for (int nt = 0; nt < T.n; ++nt) T.t[nt][x + ol] = T.tb[nt][x + ok];
  }
}

template <int WHICH>
void pbcor_run(int m, int n, int mm, int nn, int k1m) {
  Ctx& c = C(); const Geom& g = c.g;
  const std::string bmcmth = c.option("bmcmth", "uc");
  if (bmcmth != "uc" && bmcmth != "dluc")
    throw std::runtime_error(" bmcmth = " + bmcmth + " is unsupported! " + (WHICH == 1 ? "(pbcor1)" : "(pbcor2)"));
  if (g.ntr > MAXTR) throw std::runtime_error("pbcor: this build handles at most 4 passive tracers");
  const bool dluc = bmcmth == "dluc";
  const double dlt = c.scalar("dlt");
  const int ks = WHICH == 1 ? nn : mm, kf = WHICH == 1 ? mm : nn, lt = WHICH == 1 ? m : n;
  const int kk = g.kdm;
  PbTr T{}; T.n = g.ntr;
  for (int nt = 0; nt < g.ntr; ++nt) {
    T.t[nt] = c.dev("trc") + (long)nt * 2 * kk * g.lev;
    T.tb[nt] = c.owned("cppm_tmp_trc" + std::to_string(nt + 1), kk);
  }
  if (WHICH == 2) {  // :433-440
    halo_update(std::vector<HaloReq>{{c.dev("ubflxs") + (long)(n - 1) * g.lev, 1, halo_uv},
                                     {c.dev("vbflxs") + (long)(n - 1) * g.lev, 1, halo_vv}}, 1, 1);
    if (g.ntr > 0) {
      std::vector<HaloReq> r;
      for (int nt = 0; nt < g.ntr; ++nt) r.push_back({T.t[nt] + (long)(k1m - 1) * g.lev, kk, halo_ps});
      halo_update(r, 1, 1);
    }
  }
  const char* un = WHICH == 1 ? "utotm" : "utotn"; const char* vn = WHICH == 1 ? "vtotm" : "vtotn";
  double* utot = c.has(un) ? c.dev(un) : c.owned(un, 1);
  double* vtot = c.has(vn) ? c.dev(vn) : c.owned(vn, 1);
  double *dp = c.dev("dp"), *p = c.dev("p"), *temp = c.dev("temp"), *saln = c.dev("saln");
  double *dpB = c.owned("cppm_tmp_dp", kk), *tB = c.owned("cppm_tmp_temp", kk), *sB = c.owned("cppm_tmp_saln", kk);
  const int *ip = c.idev("ip"), *iu = c.idev("iu"), *iv = c.idev("iv");
  {
    dim3 grid(cdiv(g.ii + 2, 128), g.jj + 2);
    LAUNCH_NAMED(WHICH == 1 ? "pbcor_prep<1>" : "pbcor_prep<2>", pbcor_prep<WHICH>, grid, 128, 0, g, dlt, ks, kf, lt,
                 ip, iu, iv, dp, p, c.dev(WHICH == 1 ? "ubflxs_p" : "ubflxs"), c.dev(WHICH == 1 ? "vbflxs_p" : "vbflxs"),
                 c.dev("uflx"), c.dev("vflx"), utot, vtot);
  }
  {
    dim3 grid(cdiv(g.ii + 1, 128), g.jj + 1, kk);
    const char* nm = WHICH == 1 ? "pbcor_update<1>" : "pbcor_update<2>";
    const eos::Coef ec = WHICH == 2 ? eos::host_coef() : eos::Coef{};  // only pbcor2 refreshes sigma
    if (dluc)
      LAUNCH_NAMED(nm, (pbcor_update<WHICH, true>), grid, 128, 0, g, ec, ks, kf, ip, iu, iv, utot, vtot,
                   dp, p, temp, saln, c.dev("scp2i"), c.dev("uflx"), c.dev("usflx"), c.dev("utflx"), c.dev("vflx"),
                   c.dev("vsflx"), c.dev("vtflx"), dpB, tB, sB, c.dev("sigma"), T);
    else
      LAUNCH_NAMED(nm, (pbcor_update<WHICH, false>), grid, 128, 0, g, ec, ks, kf, ip, iu, iv, utot, vtot,
                   dp, p, temp, saln, c.dev("scp2i"), c.dev("uflx"), c.dev("usflx"), c.dev("utflx"), c.dev("vflx"),
                   c.dev("vsflx"), c.dev("vtflx"), dpB, tB, sB, c.dev("sigma"), T);
  }
  {
    dim3 grid(cdiv(g.ii, 128), g.jj);
    const double* pbt = WHICH == 1 ? c.dev("pb_p") : c.dev("pb") + (long)(m - 1) * g.lev;
    LAUNCH_NAMED(WHICH == 1 ? "pbcor_finish<1>" : "pbcor_finish<2>", pbcor_finish<WHICH>, grid, 128, 0, g, ks, ip, pbt,
                 dpB, tB, sB, dp, temp, saln, p, T);
  }
}

}  // namespace

void pbcor1_dev(int m, int n, int mm, int nn, int k1m, int k1n) { (void)k1n; pbcor_run<1>(m, n, mm, nn, k1m); }
void pbcor2_dev(int m, int n, int mm, int nn, int k1m, int k1n) { (void)k1n; pbcor_run<2>(m, n, mm, nn, k1m); }

}  // namespace blom
