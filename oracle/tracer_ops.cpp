// ORACLE — TEST INFRASTRUCTURE ONLY (see core.hpp header).
// Restatement of phy/mod_diffus.F90:41-185 (ltedtp='layer' and 'neutral'),
// phy/mod_tmsmt.F90:209-277 (tmsmt1), :281-410 (tmsmt2), phy/mod_eos.F90:83-155.
#include "core.hpp"
#include "eos.hpp"

namespace orc {

namespace eos {
Coef& K() { static Coef c; return c; }
void inieos_pref(double pref) {
  Coef& c = K();
  c.pref = pref;
  c.ap21 = a21 + b21 * pref; c.ap22 = a22 + b22 * pref; c.ap23 = a23 + b23 * pref;
  c.ap24 = a24; c.ap25 = a25; c.ap26 = a26;
  c.ap11 = a11 + b11 * pref - c.ap21 / alpha0;
  c.ap12 = a12 + b12 * pref - c.ap22 / alpha0;
  c.ap13 = a13 + b13 * pref - c.ap23 / alpha0;
  c.ap14 = a14 - c.ap24 / alpha0; c.ap15 = a15 - c.ap25 / alpha0; c.ap16 = a16 - c.ap26 / alpha0;
  c.ap210 = a21; c.ap220 = a22; c.ap230 = a23; c.ap240 = a24; c.ap250 = a25; c.ap260 = a26;
  c.ap110 = a11 - c.ap210 / alpha0; c.ap120 = a12 - c.ap220 / alpha0; c.ap130 = a13 - c.ap230 / alpha0;
  c.ap140 = a14 - c.ap240 / alpha0; c.ap150 = a15 - c.ap250 / alpha0; c.ap160 = a16 - c.ap260 / alpha0;
}
}  // namespace eos

void inieos() { eos::inieos_pref(O().scalar("pref", 0.0)); }

// phy/mod_diffus.F90:41-185
void diffus(int m, int n, int mm, int nn, int k1m, int k1n) {
  (void)m; (void)n; (void)k1m;
  Oracle& o = O(); const Dims& d = o.d;
  const int ii = d.ii, jj = d.jj, kk = d.kk, ntr = d.ntr;
  const double delt1 = o.scalar("delt1");
  const double dpeps = 1.e-5;
  A3 dp = o.a3("dp"), temp = o.a3("temp"), saln = o.a3("saln"), sigma = o.a3("sigma");
  A3 trc = ntr > 0 ? o.a3("trc") : A3{};
  A3 utflx = o.a3("utflx"), vtflx = o.a3("vtflx"), usflx = o.a3("usflx"), vsflx = o.a3("vsflx");
  A3 utflld = o.a3("utflld"), vtflld = o.a3("vtflld"), usflld = o.a3("usflld"), vsflld = o.a3("vsflld");
  A3 difiso = o.a3("difiso");
  A2 scuy = o.a2("scuy"), scvx = o.a2("scvx"), scp2 = o.a2("scp2"), scuxi = o.a2("scuxi"), scvyi = o.a2("scvyi");
  I2 ip = o.i2("ip"), iu = o.i2("iu"), iv = o.i2("iv");
  auto T = [&](int nt, int i, int j, int kn) -> double& { return trc(i, j, kn + (nt - 1) * 2 * d.kdm); };

  xctilr(dp.from(k1n), 1, kk, 3, 3, halo_ps);
  if (o.option("ltedtp", "layer") == "neutral") {
    xctilr(temp.from(k1n), 1, kk, 1, 1, halo_ps);
    xctilr(saln.from(k1n), 1, kk, 1, 1, halo_ps);
    for (int nt = 1; nt <= ntr; ++nt) xctilr(trc.from(k1n + (nt - 1) * 2 * d.kdm), 1, kk, 1, 1, halo_ps);
    return;
  }
  xctilr(temp.from(k1n), 1, kk, 2, 2, halo_ps);
  xctilr(saln.from(k1n), 1, kk, 2, 2, halo_ps);
  for (int nt = 1; nt <= ntr; ++nt) xctilr(trc.from(k1n + (nt - 1) * 2 * d.kdm), 1, kk, 2, 2, halo_ps);

  // uflxtr(nt,i,j), vflxtr(nt,i,j): tracer fastest (trc/mod_tracers.F90:226-235)
  std::vector<double> uflxtr((size_t)std::max(ntr, 1) * d.lev, 0.0), vflxtr((size_t)std::max(ntr, 1) * d.lev, 0.0);
  auto FX = [&](std::vector<double>& v, int nt, int i, int j) -> double& {
    return v[((size_t)(j + d.nbdy - 1) * d.ldi + (i + d.nbdy - 1)) * ntr + (nt - 1)];
  };
  for (int k = 1; k <= kk; ++k) {
    const int kn = k + nn, km = k + mm;
    for (int j = 0; j <= jj + 1; ++j)
      for (int i = 0; i <= ii + 2; ++i) {
        if (iu(i, j) != 1) continue;
        double q = delt1 * .5 * (difiso(i - 1, j, k) + difiso(i, j, k)) * scuy(i, j) * scuxi(i, j) *
                   std::max(std::min(dp(i - 1, j, kn), dp(i, j, kn)), dpeps);
        usflld(i, j, km) = q * (saln(i - 1, j, kn) - saln(i, j, kn));
        utflld(i, j, km) = q * (temp(i - 1, j, kn) - temp(i, j, kn));
        for (int nt = 1; nt <= ntr; ++nt) FX(uflxtr, nt, i, j) = q * (T(nt, i - 1, j, kn) - T(nt, i, j, kn));
        usflx(i, j, km) = usflx(i, j, km) + usflld(i, j, km);
        utflx(i, j, km) = utflx(i, j, km) + utflld(i, j, km);
      }
    for (int j = 0; j <= jj + 2; ++j)
      for (int i = 0; i <= ii + 1; ++i) {
        if (iv(i, j) != 1) continue;
        double q = delt1 * .5 * (difiso(i, j - 1, k) + difiso(i, j, k)) * scvx(i, j) * scvyi(i, j) *
                   std::max(std::min(dp(i, j - 1, kn), dp(i, j, kn)), dpeps);
        vsflld(i, j, km) = q * (saln(i, j - 1, kn) - saln(i, j, kn));
        vtflld(i, j, km) = q * (temp(i, j - 1, kn) - temp(i, j, kn));
        for (int nt = 1; nt <= ntr; ++nt) FX(vflxtr, nt, i, j) = q * (T(nt, i, j - 1, kn) - T(nt, i, j, kn));
        vsflx(i, j, km) = vsflx(i, j, km) + vsflld(i, j, km);
        vtflx(i, j, km) = vtflx(i, j, km) + vtflld(i, j, km);
      }
    for (int j = 0; j <= jj + 1; ++j)
      for (int i = 0; i <= ii + 1; ++i) {
        if (ip(i, j) != 1) continue;
        double q = 1. / (scp2(i, j) * std::max(dp(i, j, kn), dpeps));
        saln(i, j, kn) = saln(i, j, kn) -
                         q * (usflld(i + 1, j, km) - usflld(i, j, km) + vsflld(i, j + 1, km) - vsflld(i, j, km));
        temp(i, j, kn) = temp(i, j, kn) -
                         q * (utflld(i + 1, j, km) - utflld(i, j, km) + vtflld(i, j + 1, km) - vtflld(i, j, km));
        for (int nt = 1; nt <= ntr; ++nt)
          T(nt, i, j, kn) = T(nt, i, j, kn) - q * (FX(uflxtr, nt, i + 1, j) - FX(uflxtr, nt, i, j) +
                                                   FX(vflxtr, nt, i, j + 1) - FX(vflxtr, nt, i, j));
        sigma(i, j, kn) = eos::sig(temp(i, j, kn), saln(i, j, kn));
      }
  }
}

// phy/mod_tmsmt.F90:209-277
void tmsmt1(int nn) {
  Oracle& o = O(); const Dims& d = o.d;
  const int ii = d.ii, jj = d.jj, kk = d.kk, ntr = d.ntr;
  A3 dp = o.a3("dp"), temp = o.a3("temp"), saln = o.a3("saln");
  A3 dpold = o.a3("dpold"), told = o.a3("told"), sold = o.a3("sold");
  A3 trc = ntr > 0 ? o.a3("trc") : A3{}, trcold = ntr > 0 ? o.a3("trcold") : A3{};
  I2 ip = o.i2("ip"), iu = o.i2("iu"), iv = o.i2("iv");
  for (int j = 1; j <= jj; ++j)
    for (int k = 1; k <= kk; ++k) {
      const int kn = k + nn;
      for (int i = 1; i <= ii; ++i) {
        if (ip(i, j) != 1) continue;
        dpold(i, j, kn) = dp(i, j, kn);
        told(i, j, k) = temp(i, j, kn);
        sold(i, j, k) = saln(i, j, kn);
        for (int nt = 1; nt <= ntr; ++nt) trcold(i, j, k + (nt - 1) * d.kdm) = trc(i, j, kn + (nt - 1) * 2 * d.kdm);
      }
    }
  if (o.option("vcoord", "cntiso_hybrid") == "isopyc_bulkml") {
    A3 dpu = o.a3("dpu"), dpv = o.a3("dpv"), dpuold = o.a3("dpuold"), dpvold = o.a3("dpvold");
    for (int j = 1; j <= jj; ++j)
      for (int k = 1; k <= kk; ++k) {
        const int kn = k + nn;
        for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) dpuold(i, j, k) = dpu(i, j, kn);
        for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) dpvold(i, j, k) = dpv(i, j, kn);
      }
  }
}

// phy/mod_tmsmt.F90:281-410
void tmsmt2(int m, int mm, int nn, int k1m) {
  Oracle& o = O(); const Dims& d = o.d;
  const int ii = d.ii, jj = d.jj, kk = d.kk, ntr = d.ntr;
  const double wts1 = .875, wts2 = .0625;  // :49-50
  A3 dp = o.a3("dp"), temp = o.a3("temp"), saln = o.a3("saln"), p = o.a3("p"), pb = o.a3("pb");
  A3 dpold = o.a3("dpold"), told = o.a3("told"), sold = o.a3("sold");
  A3 trc = ntr > 0 ? o.a3("trc") : A3{}, trcold = ntr > 0 ? o.a3("trcold") : A3{};
  I2 ip = o.i2("ip"), iu = o.i2("iu"), iv = o.i2("iv");
  std::vector<double> pbfaco(d.ldi), pbfacn(d.ldi);
  const int nb = d.nbdy;
  for (int j = 1; j <= jj; ++j) {
    for (int i = 1; i <= ii; ++i) if (ip(i, j) == 1) { pbfaco[i + nb - 1] = 0.; pbfacn[i + nb - 1] = 0.; }
    for (int k = 1; k <= kk; ++k) {
      const int kn = k + nn;
      for (int i = 1; i <= ii; ++i) if (ip(i, j) == 1) {
        pbfaco[i + nb - 1] = pbfaco[i + nb - 1] + dpold(i, j, kn);
        pbfacn[i + nb - 1] = pbfacn[i + nb - 1] + dp(i, j, kn);
      }
    }
    for (int i = 1; i <= ii; ++i) if (ip(i, j) == 1) {
      pbfaco[i + nb - 1] = pb(i, j, m) / pbfaco[i + nb - 1];
      pbfacn[i + nb - 1] = pb(i, j, m) / pbfacn[i + nb - 1];
    }
    for (int k = 1; k <= kk; ++k) {
      const int km = k + mm, kn = k + nn;
      for (int i = 1; i <= ii; ++i) if (ip(i, j) == 1) {
        double pold = std::max(0., dpold(i, j, kn) * pbfaco[i + nb - 1]);
        double pmid = std::max(0., dp(i, j, km));
        double pnew = std::max(0., dp(i, j, kn) * pbfacn[i + nb - 1]);
        dp(i, j, km) = wts1 * pmid + wts2 * (pold + pnew);
        pold = pold + epsilp; pmid = pmid + epsilp; pnew = pnew + epsilp;
        temp(i, j, km) = (wts1 * pmid * temp(i, j, km) + wts2 * (pold * told(i, j, k) + pnew * temp(i, j, kn))) /
                         (dp(i, j, km) + epsilp);
        saln(i, j, km) = (wts1 * pmid * saln(i, j, km) + wts2 * (pold * sold(i, j, k) + pnew * saln(i, j, kn))) /
                         (dp(i, j, km) + epsilp);
        for (int nt = 1; nt <= ntr; ++nt) {
          const int o2 = (nt - 1) * 2 * d.kdm, o1 = (nt - 1) * d.kdm;
          trc(i, j, km + o2) = (wts1 * pmid * trc(i, j, km + o2) +
                                wts2 * (pold * trcold(i, j, k + o1) + pnew * trc(i, j, kn + o2))) /
                               (dp(i, j, km) + epsilp);
        }
      }
    }
  }
  xctilr(dp.from(k1m), 1, kk, 3, 3, halo_ps);
  for (int j = -2; j <= jj + 2; ++j)
    for (int k = 1; k <= kk; ++k) {
      const int km = k + mm;
      for (int i = -2; i <= ii + 2; ++i) if (ip(i, j) == 1) p(i, j, k + 1) = p(i, j, k) + dp(i, j, km);
    }
  if (o.option("vcoord", "cntiso_hybrid") == "isopyc_bulkml") {
    A3 dpu = o.a3("dpu"), dpv = o.a3("dpv");
    for (int j = -1; j <= jj + 2; ++j)
      for (int k = 1; k <= kk; ++k) {
        const int km = k + mm;
        for (int i = -1; i <= ii + 2; ++i) if (iu(i, j) == 1) {
          double q = std::min(p(i, j, kk + 1), p(i - 1, j, kk + 1));
          dpu(i, j, km) = .5 * ((std::min(q, p(i - 1, j, k + 1)) - std::min(q, p(i - 1, j, k))) +
                                (std::min(q, p(i, j, k + 1)) - std::min(q, p(i, j, k))));
        }
        for (int i = -1; i <= ii + 2; ++i) if (iv(i, j) == 1) {
          double q = std::min(p(i, j, kk + 1), p(i, j - 1, kk + 1));
          dpv(i, j, km) = .5 * ((std::min(q, p(i, j - 1, k + 1)) - std::min(q, p(i, j - 1, k))) +
                                (std::min(q, p(i, j, k + 1)) - std::min(q, p(i, j, k))));
        }
      }
  }
}

}  // namespace orc
