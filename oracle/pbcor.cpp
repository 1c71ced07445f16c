// ORACLE — TEST INFRASTRUCTURE ONLY (see core.hpp header).
// Restatement of phy/mod_pbcor.F90: pbcor1 :66-412 and pbcor2 :416-743
// (bmcmth 'uc' and 'dluc'; dpeps1=1e-5, dpeps2=1e-7 :58-59).  The two routines
// share their structure, so one body is written once with the differences
// spelled out at each step (`which` = 1 or 2):
//                      pbcor1                         pbcor2
//   state level        kn (new)                       km (mid)
//   flux accumulators  uflx..(km)                     uflx..(kn)
//   total to match     dlt*ubflxs_p(m)                dlt*ubflxs(n)  (after (1,1) halos)
//   2-D totals         utotm, vtotm                   utotn, vtotn
//   dp pre-treatment   none                           max(0,dp)+epsilp on 0..ii+1
//   update             max(0,..), dpeps1 weights      plain, sigma refreshed, -epsilp
//   final rescale      pb_p/p(kk+1), p left unscaled  pb(m)/p(kk+1), p rebuilt
// The work arrays uflux.. of mod_utility are zero at land faces next to wet
// cells (phy/mod_utility.F90:86-115) and only written at wet faces here.
#include "core.hpp"
#include "eos.hpp"

namespace orc {

namespace {

void pbcor_body(int which, int m, int n, int mm, int nn, int k1m) {
  Oracle& o = O(); const Dims& d = o.d;
  const int ii = d.ii, jj = d.jj, kk = d.kk, ntr = d.ntr;
  const double dlt = o.scalar("dlt");
  const double dpeps1 = 1.e-5, dpeps2 = 1.e-7;
  const std::string bmcmth = o.option("bmcmth", "uc");
  const char* rname = which == 1 ? "(pbcor1)" : "(pbcor2)";
  if (bmcmth != "uc" && bmcmth != "dluc")
    throw std::runtime_error(" bmcmth = " + bmcmth + " is unsupported! " + rname);
  const bool dluc = bmcmth == "dluc";
  I2 ip = o.i2("ip"), iu = o.i2("iu"), iv = o.i2("iv");
  A3 dp = o.a3("dp"), temp = o.a3("temp"), saln = o.a3("saln"), sigma = o.a3("sigma"), p = o.a3("p");
  A3 uflx = o.a3("uflx"), vflx = o.a3("vflx"), utflx = o.a3("utflx"), vtflx = o.a3("vtflx"),
     usflx = o.a3("usflx"), vsflx = o.a3("vsflx");
  A3 trc = ntr > 0 ? o.a3("trc") : A3{};
  A2 scp2i = o.a2("scp2i");
  auto T = [&](int nt, int i, int j, int kl) -> double& { return trc(i, j, kl + (nt - 1) * 2 * d.kdm); };
  const int ks = which == 1 ? nn : mm;   // level offset of the state that is corrected
  const int kf = which == 1 ? mm : nn;   // level offset of the flux accumulators
  A2 utot = o.has(which == 1 ? "utotm" : "utotn") ? o.a2(which == 1 ? "utotm" : "utotn")
                                                 : o.scratch(which == 1 ? "utotm" : "utotn", 1).level(1);
  A2 vtot = o.has(which == 1 ? "vtotm" : "vtotn") ? o.a2(which == 1 ? "vtotm" : "vtotn")
                                                 : o.scratch(which == 1 ? "vtotm" : "vtotn", 1).level(1);
  // private copies of the mod_utility work arrays: pbcor only ever reads faces it has just
  // written or land faces (zero), so sharing them with momtum is not observable
  auto work = [&](const char* nm) { return o.scratch(std::string("_pbcor_") + nm, 1).level(1); };
  A2 uflux = work("uflux"), vflux = work("vflux"), uflux2 = work("uflux2"), vflux2 = work("vflux2"),
     uflux3 = work("uflux3"), vflux3 = work("vflux3");
  A2 pbu_t = o.scratch("_pbcor_pbu_t", 1).level(1), pbv_t = o.scratch("_pbcor_pbv_t", 1).level(1);
  std::vector<double> uflxtr((size_t)std::max(ntr, 1) * d.lev, 0.0), vflxtr((size_t)std::max(ntr, 1) * d.lev, 0.0);
  auto FX = [&](std::vector<double>& v, int nt, int i, int j) -> double& {
    return v[((size_t)(j + d.nbdy - 1) * d.ldi + (i + d.nbdy - 1)) * ntr + (nt - 1)];
  };

  if (which == 2) {  // :433-440
    xctilr(o.a3("ubflxs").from(n), 1, 1, 1, 1, halo_uv);
    xctilr(o.a3("vbflxs").from(n), 1, 1, 1, 1, halo_vv);
    for (int nt = 1; nt <= ntr; ++nt) xctilr(trc.from(k1m + (nt - 1) * 2 * d.kdm), 1, kk, 1, 1, halo_ps);
  }
  // interface pressures on the 1-wide ring (:83-93 / :442-454)
  for (int j = 0; j <= jj + 1; ++j)
    for (int k = 1; k <= kk; ++k)
      for (int i = 0; i <= ii + 1; ++i) {
        if (ip(i, j) != 1) continue;
        if (which == 2) dp(i, j, k + ks) = std::max(0., dp(i, j, k + ks)) + epsilp;
        p(i, j, k + 1) = p(i, j, k) + dp(i, j, k + ks);
      }
  // flux totals still to be distributed (:95-161 / :456-521)
  A3 ubt = which == 1 ? o.a3("ubflxs_p") : o.a3("ubflxs"), vbt = which == 1 ? o.a3("vbflxs_p") : o.a3("vbflxs");
  const int lt = which == 1 ? m : n;
  for (int j = 1; j <= jj; ++j) {
    for (int i = 1; i <= ii + 1; ++i) {
      if (iu(i, j) != 1) continue;
      utot(i, j) = dlt * ubt(i, j, lt);
      if (dluc) pbu_t(i, j) = std::min(p(i, j, kk + 1), p(i - 1, j, kk + 1));
    }
    for (int k = 1; k <= kk; ++k)
      for (int i = 1; i <= ii + 1; ++i)
        if (iu(i, j) == 1) utot(i, j) = utot(i, j) - uflx(i, j, k + kf);
  }
  for (int j = 1; j <= jj + 1; ++j) {
    for (int i = 1; i <= ii; ++i) {
      if (iv(i, j) != 1) continue;
      vtot(i, j) = dlt * vbt(i, j, lt);
      if (dluc) pbv_t(i, j) = std::min(p(i, j, kk + 1), p(i, j - 1, kk + 1));
    }
    for (int k = 1; k <= kk; ++k)
      for (int i = 1; i <= ii; ++i)
        if (iv(i, j) == 1) vtot(i, j) = vtot(i, j) - vflx(i, j, k + kf);
  }

  for (int k = 1; k <= kk; ++k) {
    const int kl = k + ks, ka = k + kf;
    // upstream-column distribution of the residual (:163-343 / :523-683)
    for (int j = 1; j <= jj; ++j)
      for (int i = 1; i <= ii + 1; ++i) {
        if (iu(i, j) != 1) continue;
        const int iup = utot(i, j) > 0. ? i - 1 : i;
        if (!dluc) uflux(i, j) = utot(i, j) * dp(iup, j, kl) / p(iup, j, kk + 1);
        else uflux(i, j) = utot(i, j) * std::max(0., std::min(pbu_t(i, j), p(iup, j, k + 1)) - p(iup, j, k)) /
                           pbu_t(i, j);
        uflux2(i, j) = uflux(i, j) * saln(iup, j, kl);
        uflux3(i, j) = uflux(i, j) * temp(iup, j, kl);
        for (int nt = 1; nt <= ntr; ++nt) FX(uflxtr, nt, i, j) = uflux(i, j) * T(nt, iup, j, kl);
        uflx(i, j, ka) = uflx(i, j, ka) + uflux(i, j);
        usflx(i, j, ka) = usflx(i, j, ka) + uflux2(i, j);
        utflx(i, j, ka) = utflx(i, j, ka) + uflux3(i, j);
      }
    for (int j = 1; j <= jj + 1; ++j)
      for (int i = 1; i <= ii; ++i) {
        if (iv(i, j) != 1) continue;
        const int jup = vtot(i, j) > 0. ? j - 1 : j;
        if (!dluc) vflux(i, j) = vtot(i, j) * dp(i, jup, kl) / p(i, jup, kk + 1);
        else vflux(i, j) = vtot(i, j) * std::max(0., std::min(pbv_t(i, j), p(i, jup, k + 1)) - p(i, jup, k)) /
                           pbv_t(i, j);
        vflux2(i, j) = vflux(i, j) * saln(i, jup, kl);
        vflux3(i, j) = vflux(i, j) * temp(i, jup, kl);
        for (int nt = 1; nt <= ntr; ++nt) FX(vflxtr, nt, i, j) = vflux(i, j) * T(nt, i, jup, kl);
        vflx(i, j, ka) = vflx(i, j, ka) + vflux(i, j);
        vsflx(i, j, ka) = vsflx(i, j, ka) + vflux2(i, j);
        vtflx(i, j, ka) = vtflx(i, j, ka) + vflux3(i, j);
      }
    // divergence update (:345-375 / :685-712)
    for (int j = 1; j <= jj; ++j)
      for (int i = 1; i <= ii; ++i) {
        if (ip(i, j) != 1) continue;
        double dpo = dp(i, j, kl), dpni;
        const double dm = uflux(i + 1, j) - uflux(i, j) + vflux(i, j + 1) - vflux(i, j);
        const double ds = uflux2(i + 1, j) - uflux2(i, j) + vflux2(i, j + 1) - vflux2(i, j);
        const double dt = uflux3(i + 1, j) - uflux3(i, j) + vflux3(i, j + 1) - vflux3(i, j);
        if (which == 1) {
          dp(i, j, kl) = std::max(0., dpo - dm * scp2i(i, j));
          dpo = dpo + dpeps1;
          dpni = 1. / (dp(i, j, kl) + dpeps1);
        } else {
          dp(i, j, kl) = dpo - scp2i(i, j) * dm;
          dpni = 1. / dp(i, j, kl);
        }
        saln(i, j, kl) = (dpo * saln(i, j, kl) - ds * scp2i(i, j)) * dpni;
        temp(i, j, kl) = (dpo * temp(i, j, kl) - dt * scp2i(i, j)) * dpni;
        for (int nt = 1; nt <= ntr; ++nt)
          T(nt, i, j, kl) = (dpo * T(nt, i, j, kl) - (FX(uflxtr, nt, i + 1, j) - FX(uflxtr, nt, i, j) +
                                                     FX(vflxtr, nt, i, j + 1) - FX(vflxtr, nt, i, j)) * scp2i(i, j)) * dpni;
        if (which == 2) {
          sigma(i, j, kl) = eos::sig(temp(i, j, kl), saln(i, j, kl));
          dp(i, j, kl) = dp(i, j, kl) - epsilp;
        }
        if (dp(i, j, kl) < dpeps2) dp(i, j, kl) = 0.;
      }
  }
  // rescale the column to the barotropic bottom pressure (:379-401 / :716-741)
  A2 pbt = which == 1 ? o.a2("pb_p") : o.a3("pb").level(m);
  for (int j = 1; j <= jj; ++j)
    for (int i = 1; i <= ii; ++i) {
      if (ip(i, j) != 1) continue;
      for (int k = 1; k <= kk; ++k) p(i, j, k + 1) = p(i, j, k) + dp(i, j, k + ks);
      const double pbfac = pbt(i, j) / p(i, j, kk + 1);
      for (int k = 1; k <= kk; ++k) {
        dp(i, j, k + ks) = dp(i, j, k + ks) * pbfac;
        if (which == 2) p(i, j, k + 1) = p(i, j, k) + dp(i, j, k + ks);
      }
    }
}

}  // namespace

void pbcor1(int m, int n, int mm, int nn, int k1m, int k1n) { (void)k1n; pbcor_body(1, m, n, mm, nn, k1m); }
void pbcor2(int m, int n, int mm, int nn, int k1m, int k1n) { (void)k1n; pbcor_body(2, m, n, mm, nn, k1m); }

}  // namespace orc
