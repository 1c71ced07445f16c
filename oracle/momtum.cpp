// ORACLE — TEST INFRASTRUCTURE ONLY (see core.hpp header).
// Restatement of phy/mod_momtum.F90:215-1282 (serial order: the k loop runs
// ascending on the shared module work arrays, i.e. the no-OpenMP build).
#include "core.hpp"

namespace orc {

namespace {
inline double hfharm(double a, double b) { return a * b / (a + b); }  // :131-141
inline double sq(double a) { return a * a; }
struct Span {  // span tables of bigrid: first/last/count per row (or per column)
  const int *f, *l, *s; int nb, ms;
  int first(int r, int k) const { return f[(size_t)(r + nb - 1) * ms + k - 1]; }
  int last(int r, int k) const { return l[(size_t)(r + nb - 1) * ms + k - 1]; }
  int count(int r) const { return s[r + nb - 1]; }
};
}  // namespace

void momtum(int m, int n, int mm, int nn, int k1m, int k1n) {
  (void)k1m; (void)k1n;
  Oracle& o = O(); const Dims& d = o.d;
  const int ii = d.ii, jj = d.jj, kk = d.kk, nb = d.nbdy;
  const double c1 = 1. - 1.5 * .5, c2 = 1. - .5, c3 = 2., slope = .5;  // :221
  const double slip = -1., thkbot = 10.;                               // :93-97
  const double wuv1 = .75, wuv2 = .125, wpgf = .25;                    // mod_tmsmt.F90:47-48, mod_pgforc.F90:47
  const double delt1 = o.scalar("delt1"), dlt = o.scalar("dlt");
  const double mdv2hi = o.scalar("mdv2hi", 0.), mdv2lo = o.scalar("mdv2lo", 0.), mdv4hi = o.scalar("mdv4hi", 0.),
               mdv4lo = o.scalar("mdv4lo", 0.), vsc2hi = o.scalar("vsc2hi", 0.), vsc2lo = o.scalar("vsc2lo", 0.),
               vsc4hi = o.scalar("vsc4hi", 0.), vsc4lo = o.scalar("vsc4lo", 0.), cbar = o.scalar("cbar", 0.),
               cb = o.scalar("cb", 0.);
  const std::string mommth = o.option("mommth", "enscon");
  const bool isopyc = o.option("vcoord", "cntiso_hybrid") == "isopyc_bulkml";
  if (mommth != "enscon" && mommth != "enecon" && mommth != "enedis")
    throw std::runtime_error(" mommth = " + mommth + " is unsupported!");

  A3 p = o.a3("p"), dp = o.a3("dp"), u = o.a3("u"), v = o.a3("v"), dpu = o.a3("dpu"), dpv = o.a3("dpv"),
     pu = o.a3("pu"), pv = o.a3("pv"), pbu = o.a3("pbu"), pbv = o.a3("pbv"), ubflxs_p = o.a3("ubflxs_p"),
     vbflxs_p = o.a3("vbflxs_p"), ub = o.a3("ub"), vb = o.a3("vb"), pgfx = o.a3("pgfx"), pgfy = o.a3("pgfy"),
     pgfx_o = o.a3("pgfx_o"), pgfy_o = o.a3("pgfy_o"), dpuold = o.a3("dpuold"), dpvold = o.a3("dpvold"),
     mu_nonloc = o.a3("mu_nonloc"), mv_nonloc = o.a3("mv_nonloc"), absvor = o.a3("absvor"), dpvor = o.a3("dpvor");
  A2 ubcors_p = o.a2("ubcors_p"), vbcors_p = o.a2("vbcors_p"), pbu_p = o.a2("pbu_p"), pbv_p = o.a2("pbv_p"),
     difwgt = o.a2("difwgt"), difmxp = o.a2("difmxp"), difmxq = o.a2("difmxq"), taux = o.a2("taux"),
     tauy = o.a2("tauy"), ustarb = o.a2("ustarb"), umax = o.a2("umax"), vmax = o.a2("vmax"), utotn = o.a2("utotn"),
     vtotn = o.a2("vtotn");
  A2 scuy = o.a2("scuy"), scvx = o.a2("scvx"), scux = o.a2("scux"), scvy = o.a2("scvy"), scq2i = o.a2("scq2i"),
     scp2i = o.a2("scp2i"), scp2 = o.a2("scp2"), scu2 = o.a2("scu2"), scv2 = o.a2("scv2"), scpx = o.a2("scpx"),
     scpy = o.a2("scpy"), scqx = o.a2("scqx"), scqy = o.a2("scqy"), scuxi = o.a2("scuxi"), scvyi = o.a2("scvyi"),
     corioq = o.a2("corioq");
  I2 ip = o.i2("ip"), iu = o.i2("iu"), iv = o.i2("iv"), iq = o.i2("iq");
  // module work arrays (mod_utility, mod_momtum): persist between calls
  auto M = [&](const char* nm) { return o.scratch(std::string("momtum_") + nm, 1).level(1); };
  A2 utotm = M("utotm"), vtotm = M("vtotm"), uflux = M("uflux"), vflux = M("vflux"), uflux2 = M("uflux2"),
     uflux3 = M("uflux3"), vflux2 = M("vflux2"), vflux3 = M("vflux3"), uja = M("uja"), ujb = M("ujb"), via = M("via"),
     vib = M("vib"), defor1 = M("defor1"), defor2 = M("defor2"), util1 = M("util1"), util2 = M("util2");
  // routine locals (:223-226): fresh per call
  std::vector<std::vector<double>> loc(27, std::vector<double>(d.lev, 0.0));
  int nl = 0;
  auto Lc = [&]() { return A2{loc[nl++].data(), d.ldi, nb}; };
  A2 drag = Lc(), ubrhs = Lc(), vbrhs = Lc(), stress = Lc(), dpmx = Lc(), vsc2 = Lc(), vsc4 = Lc(), potvor = Lc(),
     vort = Lc(), wgtia = Lc(), wgtib = Lc(), wgtja = Lc(), wgtjb = Lc(), dl2u = Lc(), dl2uja = Lc(), dl2ujb = Lc(),
     dl2v = Lc(), dl2via = Lc(), dl2vib = Lc(), ke = Lc(), uh_min = Lc(), uh_max = Lc(), vh_min = Lc(), vh_max = Lc(),
     cau = Lc(), cav = Lc(), uflux1 = Lc();
  std::vector<double> vf1(d.lev, 0.0);
  A2 vflux1{vf1.data(), d.ldi, nb};
  auto& oi = o.owni;
  Span SU{oi["ifu"].data(), oi["ilu"].data(), oi["isu"].data(), nb, Oracle::ms};
  Span SV{oi["ifv"].data(), oi["ilv"].data(), oi["isv"].data(), nb, Oracle::ms};
  Span JU{oi["jfu"].data(), oi["jlu"].data(), oi["jsu"].data(), nb, Oracle::ms};
  Span JV{oi["jfv"].data(), oi["jlv"].data(), oi["jsv"].data(), nb, Oracle::ms};

  const double cutoff = onem, thkbop = thkbot * onem, tsfac = dlt / delt1, dt1inv = 1. / delt1;

  // :244-255
  for (int j = -1; j <= jj + 2; ++j)
    for (int k = 1; k <= kk; ++k)
      for (int i = -1; i <= ii + 2; ++i) if (ip(i, j) == 1) p(i, j, k + 1) = p(i, j, k) + dp(i, j, k + mm);
  // :259-293 bottom drag
  for (int j = 0; j <= jj; ++j) {
    for (int i = 0; i <= ii; ++i) if (ip(i, j) == 1) { util1(i, j) = 0.; util2(i, j) = 0.; }
    for (int k = 1; k <= kk; ++k) {
      const int kn = k + nn;
      for (int i = 0; i <= ii; ++i) if (ip(i, j) == 1) {
        double pbotl = std::max(p(i, j, k + 1), p(i, j, kk + 1) - thkbop);
        double ptopl = std::max(p(i, j, k), p(i, j, kk + 1) - thkbop);
        util1(i, j) = util1(i, j) + (u(i, j, kn) + u(i + 1, j, kn)) * (pbotl - ptopl);
        util2(i, j) = util2(i, j) + (v(i, j, kn) + v(i, j + 1, kn)) * (pbotl - ptopl);
      }
    }
    for (int i = 0; i <= ii; ++i) if (ip(i, j) == 1) {
      double ubot = (ubflxs_p(i, j, n) / std::max(epsilpl, pbu(i, j, n) * scuy(i, j)) +
                     ubflxs_p(i + 1, j, n) / std::max(epsilpl, pbu(i + 1, j, n) * scuy(i + 1, j))) * tsfac +
                    util1(i, j) / thkbop;
      double vbot = (vbflxs_p(i, j, n) / std::max(epsilpl, pbv(i, j, n) * scvx(i, j)) +
                     vbflxs_p(i, j + 1, n) / std::max(epsilpl, pbv(i, j + 1, n) * scvx(i, j + 1))) * tsfac +
                    util2(i, j) / thkbop;
      double ubbl = .5 * std::sqrt(ubot * ubot + vbot * vbot);
      double q = cb * (ubbl + cbar);
      drag(i, j) = q * grav / (alpha0 * thkbop);
      ustarb(i, j) = std::sqrt(q * ubbl);
    }
  }
  // :298-311
  for (int j = 1; j <= jj; ++j) {
    for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) ubrhs(i, j) = ubcors_p(i, j) * tsfac;
    for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) vbrhs(i, j) = vbcors_p(i, j) * tsfac;
  }
  for (int j = 0; j <= jj + 1; ++j)
    for (int i = 0; i <= ii + 1; ++i) { dl2u(i, j) = 0.; dl2v(i, j) = 0.; }
  // :322-338
  for (int k = 1; k <= kk; ++k) {
    const int km = k + mm;
    for (int j = -1; j <= jj + 2; ++j) {
      for (int i = -1; i <= ii + 2; ++i) if (iu(i, j) == 1) pu(i, j, k + 1) = pu(i, j, k) + dpu(i, j, km);
      for (int i = -1; i <= ii + 2; ++i) if (iv(i, j) == 1) pv(i, j, k + 1) = pv(i, j, k) + dpv(i, j, km);
    }
  }
  xctilr(difwgt, 2, 2, halo_ps);  // :340

  for (int k = 1; k <= kk; ++k) {
    const int km = k + mm, kn = k + nn;
    // :360-396 dpmx
    for (int j = 0; j <= jj + 2; ++j) for (int i = 0; i <= ii + 2; ++i) dpmx(i, j) = 8. * cutoff;
    for (int j = 0; j <= jj + 2; ++j)
      for (int i = 0; i <= ii + 2; ++i) if (iu(i, j) == 1) dpmx(i, j) = std::max(dpmx(i, j), dp(i, j, km) + dp(i - 1, j, km));
    for (int j = -1; j <= jj + 1; ++j)
      for (int i = 0; i <= ii + 2; ++i) if (iu(i, j) == 1) dpmx(i, j + 1) = std::max(dpmx(i, j + 1), dp(i, j, km) + dp(i - 1, j, km));
    for (int j = 0; j <= jj + 2; ++j)
      for (int i = 0; i <= ii + 2; ++i) if (iv(i, j) == 1) dpmx(i, j) = std::max(dpmx(i, j), dp(i, j, km) + dp(i, j - 1, km));
    for (int j = 0; j <= jj + 2; ++j)
      for (int i = -1; i <= ii + 1; ++i) if (iv(i, j) == 1) dpmx(i + 1, j) = std::max(dpmx(i + 1, j), dp(i, j, km) + dp(i, j - 1, km));
    // :398-431 total velocities
    for (int j = 0; j <= jj + 1; ++j)
      for (int i = 0; i <= ii + 1; ++i) if (iu(i, j) == 1) {
        utotm(i, j) = u(i, j, km) + ubflxs_p(i, j, m) * tsfac / (pbu(i, j, m) * scuy(i, j));
        uflux(i, j) = utotm(i, j) * std::max(dpu(i, j, km), cutoff);
      }
    for (int j = -1; j <= jj + 2; ++j)
      for (int i = -1; i <= ii + 2; ++i) if (iu(i, j) == 1)
        utotn(i, j) = u(i, j, kn) + ubflxs_p(i, j, n) * tsfac / (pbu(i, j, n) * scuy(i, j));
    for (int j = 0; j <= jj + 1; ++j)
      for (int i = 0; i <= ii + 1; ++i) if (iv(i, j) == 1) {
        vtotm(i, j) = v(i, j, km) + vbflxs_p(i, j, m) * tsfac / (pbv(i, j, m) * scvx(i, j));
        vflux(i, j) = vtotm(i, j) * std::max(dpv(i, j, km), cutoff);
      }
    for (int j = -1; j <= jj + 2; ++j)
      for (int i = -1; i <= ii + 2; ++i) if (iv(i, j) == 1)
        vtotn(i, j) = v(i, j, kn) + vbflxs_p(i, j, n) * tsfac / (pbv(i, j, n) * scvx(i, j));
    // :438-472 sidewall weights, auxiliary velocities, del2
    for (int j = -1; j <= jj + 2; ++j)
      for (int i = 0; i <= ii + 2; ++i) if (iu(i, j) == 1) {
        wgtja(i, j) = std::max(0., std::min(1., (pu(i, j, k + 1) - pbu(i, j - 1, m)) / std::max(pu(i, j, k + 1) - pu(i, j, k), epsilp)));
        wgtjb(i, j) = std::max(0., std::min(1., (pu(i, j, k + 1) - pbu(i, j + 1, m)) / std::max(pu(i, j, k + 1) - pu(i, j, k), epsilp)));
        uja(i, j) = (1. - wgtja(i, j)) * utotn(i, j - 1) + wgtja(i, j) * slip * utotn(i, j);
        ujb(i, j) = (1. - wgtjb(i, j)) * utotn(i, j + 1) + wgtjb(i, j) * slip * utotn(i, j);
        dl2u(i, j) = utotn(i, j) - .25 * (utotn(i + 1, j) + utotn(i - 1, j) + uja(i, j) + ujb(i, j));
      }
    for (int j = 0; j <= jj + 2; ++j)
      for (int i = -1; i <= ii + 2; ++i) if (iv(i, j) == 1) {
        wgtia(i, j) = std::max(0., std::min(1., (pv(i, j, k + 1) - pbv(i - 1, j, m)) / std::max(pv(i, j, k + 1) - pv(i, j, k), epsilp)));
        wgtib(i, j) = std::max(0., std::min(1., (pv(i, j, k + 1) - pbv(i + 1, j, m)) / std::max(pv(i, j, k + 1) - pv(i, j, k), epsilp)));
        via(i, j) = (1. - wgtia(i, j)) * vtotn(i - 1, j) + wgtia(i, j) * slip * vtotn(i, j);
        vib(i, j) = (1. - wgtib(i, j)) * vtotn(i + 1, j) + wgtib(i, j) * slip * vtotn(i, j);
        dl2v(i, j) = vtotn(i, j) - .25 * (vtotn(i, j + 1) + vtotn(i, j - 1) + via(i, j) + vib(i, j));
      }
    // :477-496 vorticity at lateral boundary points (v spans)
    for (int j = 1; j <= jj + 1; ++j)
      for (int l = 1; l <= SV.count(j); ++l) {
        int i = SV.first(j, l);
        if (i >= 1 && i <= ii + 1) {
          vort(i, j) = vtotm(i, j) * (1. - slip) * scvy(i, j) * scq2i(i, j);
          absvor(i, j, k) = vort(i, j) + corioq(i, j);
          dpvor(i, j, k) = .125 * std::max(std::max(4. * (dp(i, j, km) + dp(i, j - 1, km)), dpmx(i, j)), dpmx(i + 1, j));
          potvor(i, j) = absvor(i, j, k) / dpvor(i, j, k);
        }
        i = SV.last(j, l);
        if (i >= 0 && i <= ii) {
          vort(i + 1, j) = -vtotm(i, j) * (1. - slip) * scvy(i, j) * scq2i(i + 1, j);
          absvor(i + 1, j, k) = vort(i + 1, j) + corioq(i + 1, j);
          dpvor(i + 1, j, k) = .125 * std::max(std::max(4. * (dp(i, j, km) + dp(i, j - 1, km)), dpmx(i, j)), dpmx(i + 1, j));
          potvor(i + 1, j) = absvor(i + 1, j, k) / dpvor(i + 1, j, k);
        }
      }
    // :498-509
    for (int j = 0; j <= jj + 2; ++j)
      for (int l = 1; l <= SV.count(j); ++l) {
        int i = SV.first(j, l);
        if (i >= 0) defor2(i, j) = sq(vtotn(i, j) * (1. - slip) * scvy(i, j)) * scq2i(i, j);
        i = SV.last(j, l);
        if (i < ii + 2) defor2(i + 1, j) = sq(vtotn(i, j) * (1. - slip) * scvy(i, j)) * scq2i(i + 1, j);
      }
    // :511-530 (u spans in j)
    for (int i = 1; i <= ii + 1; ++i)
      for (int l = 1; l <= JU.count(i); ++l) {
        int j = JU.first(i, l);
        if (j >= 1 && j <= jj + 1) {
          vort(i, j) = -utotm(i, j) * (1. - slip) * scux(i, j) * scq2i(i, j);
          absvor(i, j, k) = vort(i, j) + corioq(i, j);
          dpvor(i, j, k) = .125 * std::max(std::max(4. * (dp(i, j, km) + dp(i - 1, j, km)), dpmx(i, j)), dpmx(i, j + 1));
          potvor(i, j) = absvor(i, j, k) / dpvor(i, j, k);
        }
        j = JU.last(i, l);
        if (j >= 0 && j <= jj) {
          vort(i, j + 1) = utotm(i, j) * (1. - slip) * scux(i, j) * scq2i(i, j + 1);
          absvor(i, j + 1, k) = vort(i, j + 1) + corioq(i, j + 1);
          dpvor(i, j + 1, k) = .125 * std::max(std::max(4. * (dp(i, j, km) + dp(i - 1, j, km)), dpmx(i, j)), dpmx(i, j + 1));
          potvor(i, j + 1) = absvor(i, j + 1, k) / dpvor(i, j + 1, k);
        }
      }
    // :532-543
    for (int i = 0; i <= ii + 2; ++i)
      for (int l = 1; l <= JU.count(i); ++l) {
        int j = JU.first(i, l);
        if (j >= 0) defor2(i, j) = sq(utotn(i, j) * (1. - slip) * scux(i, j)) * scq2i(i, j);
        j = JU.last(i, l);
        if (j < jj + 2) defor2(i, j + 1) = sq(utotn(i, j) * (1. - slip) * scux(i, j)) * scq2i(i, j + 1);
      }
    // :549-585 interior points
    for (int j = -1; j <= jj + 1; ++j)
      for (int i = -1; i <= ii + 1; ++i) if (ip(i, j) == 1)
        defor1(i, j) = sq((utotn(i + 1, j) * scuy(i + 1, j) - utotn(i, j) * scuy(i, j)) -
                          (vtotn(i, j + 1) * scvx(i, j + 1) - vtotn(i, j) * scvx(i, j))) * scp2i(i, j);
    for (int j = 1; j <= jj + 1; ++j)
      for (int i = 1; i <= ii + 1; ++i) if (iq(i, j) == 1) {
        vort(i, j) = (vtotm(i, j) * scvy(i, j) - vtotm(i - 1, j) * scvy(i - 1, j) - utotm(i, j) * scux(i, j) +
                      utotm(i, j - 1) * scux(i, j - 1)) * scq2i(i, j);
        absvor(i, j, k) = vort(i, j) + corioq(i, j);
        double mx = 2. * (dp(i, j, km) + dp(i - 1, j, km) + dp(i, j - 1, km) + dp(i - 1, j - 1, km));
        mx = std::max(mx, dpmx(i, j)); mx = std::max(mx, dpmx(i - 1, j)); mx = std::max(mx, dpmx(i + 1, j));
        mx = std::max(mx, dpmx(i, j - 1)); mx = std::max(mx, dpmx(i, j + 1));
        dpvor(i, j, k) = .125 * mx;
        potvor(i, j) = absvor(i, j, k) / dpvor(i, j, k);
      }
    for (int j = 0; j <= jj + 2; ++j)
      for (int i = 0; i <= ii + 2; ++i) if (iq(i, j) == 1)
        defor2(i, j) = sq(vib(i - 1, j) * scvy(i, j) - via(i, j) * scvy(i - 1, j) + ujb(i, j - 1) * scux(i, j) -
                          uja(i, j) * scux(i, j - 1)) * scq2i(i, j);
    // :591-608
    for (int j = 1; j <= jj; ++j) {
      for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) {
        dl2uja(i, j) = (1. - wgtja(i, j)) * dl2u(i, j - 1) + wgtja(i, j) * slip * dl2u(i, j);
        dl2ujb(i, j) = (1. - wgtjb(i, j)) * dl2u(i, j + 1) + wgtjb(i, j) * slip * dl2u(i, j);
      }
      for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) {
        dl2via(i, j) = (1. - wgtia(i, j)) * dl2v(i - 1, j) + wgtia(i, j) * slip * dl2v(i, j);
        dl2vib(i, j) = (1. - wgtib(i, j)) * dl2v(i + 1, j) + wgtib(i, j) * slip * dl2v(i, j);
      }
    }
    // :613-662 kinetic energy (GOLD version of Arakawa)
    for (int j = 0; j <= jj; ++j)
      for (int i = 0; i <= ii; ++i) if (ip(i, j) == 1)
        ke(i, j) = .25 * (scu2(i, j) * sq(utotm(i, j)) + scu2(i + 1, j) * sq(utotm(i + 1, j)) +
                          scv2(i, j) * sq(vtotm(i, j)) + scv2(i, j + 1) * sq(vtotm(i, j + 1))) / scp2(i, j);
    if (mommth == "enedis") {  // :664-719
      for (int j = 0; j <= jj + 1; ++j) {
        for (int i = 0; i <= ii + 1; ++i) if (iu(i, j) == 1) {
          double uhc = .5 * utotm(i, j) * (dp(i, j, km) + dp(i - 1, j, km)), uhm = uflux(i, j);
          if (std::fabs(uhc) < .1 * std::fabs(uhm)) uhm = 10. * uhc;
          else if (std::fabs(uhc) > c1 * std::fabs(uhm)) {
            if (std::fabs(uhc) < c2 * std::fabs(uhm)) uhc = (3. * uhc + (1. - c2 * 3.) * uhm);
            else if (std::fabs(uhc) <= c3 * std::fabs(uhm)) uhc = uhm;
            else uhc = slope * uhc + (1. - c3 * slope) * uhm;
          }
          if (uhc > uhm) { uh_min(i, j) = uhm; uh_max(i, j) = uhc; } else { uh_max(i, j) = uhm; uh_min(i, j) = uhc; }
        }
        for (int i = 0; i <= ii + 1; ++i) if (iv(i, j) == 1) {
          double vhc = .5 * vtotm(i, j) * (dp(i, j, km) + dp(i, j - 1, km)), vhm = vflux(i, j);
          if (std::fabs(vhc) < .1 * std::fabs(vhm)) vhm = 10. * vhc;
          else if (std::fabs(vhc) > c1 * std::fabs(vhm)) {
            if (std::fabs(vhc) < c2 * std::fabs(vhm)) vhc = (3. * vhc + (1. - c2 * 3.) * vhm);
            else if (std::fabs(vhc) <= c3 * std::fabs(vhm)) vhc = vhm;
            else vhc = slope * vhc + (1. - c3 * slope) * vhm;
          }
          if (vhc > vhm) { vh_min(i, j) = vhm; vh_max(i, j) = vhc; } else { vh_max(i, j) = vhm; vh_min(i, j) = vhc; }
        }
      }
    }
    // :723-821 coriolis / advection
    for (int j = 1; j <= jj; ++j) {
      for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) {
        if (mommth == "enscon")
          cau(i, j) = .125 * (vflux(i, j) + vflux(i, j + 1) + vflux(i - 1, j) + vflux(i - 1, j + 1)) *
                      (potvor(i, j) + potvor(i, j + 1));
        else if (mommth == "enecon")
          cau(i, j) = .25 * ((vflux(i, j) + vflux(i - 1, j)) * potvor(i, j) +
                             (vflux(i, j + 1) + vflux(i - 1, j + 1)) * potvor(i, j + 1));
        else {
          double temp1, temp2;
          if (potvor(i, j + 1) * utotm(i, j) == 0.)
            temp1 = potvor(i, j + 1) * ((vh_max(i, j + 1) + vh_max(i - 1, j + 1)) + (vh_min(i, j + 1) + vh_min(i - 1, j + 1))) * .5;
          else if (potvor(i, j + 1) * utotm(i, j) < 0.) temp1 = potvor(i, j + 1) * (vh_max(i, j + 1) + vh_max(i - 1, j + 1));
          else temp1 = potvor(i, j + 1) * (vh_min(i, j + 1) + vh_min(i - 1, j + 1));
          if (potvor(i, j) * utotm(i, j) == 0.)
            temp2 = potvor(i, j) * ((vh_max(i, j) + vh_max(i - 1, j)) + (vh_min(i, j) + vh_min(i - 1, j))) * .5;
          else if (potvor(i, j) * utotm(i, j) < 0.) temp2 = potvor(i, j) * (vh_max(i, j) + vh_max(i - 1, j));
          else temp2 = potvor(i, j) * (vh_min(i, j) + vh_min(i - 1, j));
          cau(i, j) = .25 * (temp1 + temp2);
        }
      }
      for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) {
        if (mommth == "enscon")
          cav(i, j) = -.125 * (uflux(i, j) + uflux(i + 1, j) + uflux(i, j - 1) + uflux(i + 1, j - 1)) *
                      (potvor(i, j) + potvor(i + 1, j));
        else if (mommth == "enecon")
          cav(i, j) = -.25 * ((uflux(i, j) + uflux(i, j - 1)) * potvor(i, j) +
                              (uflux(i + 1, j) + uflux(i + 1, j - 1)) * potvor(i + 1, j));
        else {
          double temp1, temp2;
          if (potvor(i + 1, j) * vtotm(i, j) == 0.)
            temp1 = potvor(i + 1, j) * ((uh_max(i + 1, j) + uh_max(i + 1, j - 1)) + (uh_min(i + 1, j) + uh_min(i + 1, j - 1))) * .5;
          else if (potvor(i + 1, j) * vtotm(i, j) > 0.) temp1 = potvor(i + 1, j) * (uh_max(i + 1, j) + uh_max(i + 1, j - 1));
          else temp1 = potvor(i + 1, j) * (uh_min(i + 1, j) + uh_min(i + 1, j - 1));
          if (potvor(i, j) * vtotm(i, j) == 0.)
            temp2 = potvor(i, j) * ((uh_max(i, j) + uh_max(i, j - 1)) + (uh_min(i, j) + uh_min(i, j - 1))) * .5;
          else if (potvor(i, j) * vtotm(i, j) > 0.) temp2 = potvor(i, j) * (uh_max(i, j) + uh_max(i, j - 1));
          else temp2 = potvor(i, j) * (uh_min(i, j) + uh_min(i, j - 1));
          cav(i, j) = -.25 * (temp1 + temp2);
        }
      }
    }
    // ---------- u equation ---------- :829-980
    for (int j = 0; j <= jj + 1; ++j)
      for (int i = 0; i <= ii + 1; ++i) if (iu(i, j) == 1) {
        double q = .5 * (difwgt(i - 1, j) + difwgt(i, j));
        double deform = std::sqrt(.5 * (defor1(i, j) + defor1(i - 1, j) + defor2(i, j) + defor2(i, j + 1)));
        vsc2(i, j) = std::max(q * mdv2hi + (1. - q) * mdv2lo, (q * vsc2hi + (1. - q) * vsc2lo) * deform);
        vsc4(i, j) = std::max(q * mdv4hi + (1. - q) * mdv4lo, (q * vsc4hi + (1. - q) * vsc4lo) * deform);
      }
    for (int j = 1; j <= jj; ++j) {
      for (int l = 1; l <= SU.count(j); ++l) {
        int i = SU.first(j, l);
        if (i > 0) { vsc2(i - 1, j) = vsc2(i, j); vsc4(i - 1, j) = vsc4(i, j); }
        i = SU.last(j, l);
        if (i < ii + 1) { vsc2(i + 1, j) = vsc2(i, j); vsc4(i + 1, j) = vsc4(i, j); }
      }
      for (int i = 0; i <= ii; ++i) if (ip(i, j) == 1) {
        if (iu(i, j) + iu(i + 1, j) > 0) {
          double dpxy = std::max(dpu(i, j, km), onemm), dpib = std::max(dpu(i + 1, j, km), onemm);
          uflux1(i, j) = std::min(difmxp(i, j), (vsc2(i, j) + vsc2(i + 1, j)) * scpy(i, j)) * hfharm(dpxy, dpib) *
                             (utotn(i, j) - utotn(i + 1, j)) +
                         std::min(.125 * difmxp(i, j), (vsc4(i, j) + vsc4(i + 1, j)) * scpy(i, j)) * hfharm(dpxy, dpib) *
                             (dl2u(i, j) - dl2u(i + 1, j));
        }
      }
      for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) {
        double dpxy = std::max(dpu(i, j, km), onemm);
        double dpja = std::max(dpu(i, j - 1, km), onemm);
        dpja = dpja + wgtja(i, j) * (dpxy - dpja);
        double dpjb = std::max(dpu(i, j + 1, km), onemm);
        dpjb = dpjb + wgtjb(i, j) * (dpxy - dpjb);
        double vsc2a, vsc4a, vsc2b, vsc4b;
        if (iu(i, j - 1) == 0) { vsc2a = vsc2(i, j); vsc4a = vsc4(i, j); } else { vsc2a = vsc2(i, j - 1); vsc4a = vsc4(i, j - 1); }
        if (iu(i, j + 1) == 0) { vsc2b = vsc2(i, j); vsc4b = vsc4(i, j); } else { vsc2b = vsc2(i, j + 1); vsc4b = vsc4(i, j + 1); }
        uflux2(i, j) = std::min(difmxq(i, j), (vsc2(i, j) + vsc2a) * scqx(i, j)) * hfharm(dpja, dpxy) * (uja(i, j) - utotn(i, j)) +
                       std::min(.125 * difmxq(i, j), (vsc4(i, j) + vsc4a) * scqx(i, j)) * hfharm(dpja, dpxy) *
                           (dl2uja(i, j) - dl2u(i, j));
        uflux3(i, j) = std::min(difmxq(i, j + 1), (vsc2(i, j) + vsc2b) * scqx(i, j + 1)) * hfharm(dpjb, dpxy) *
                           (utotn(i, j) - ujb(i, j)) +
                       std::min(.125 * difmxq(i, j + 1), (vsc4(i, j) + vsc4b) * scqx(i, j + 1)) * hfharm(dpjb, dpxy) *
                           (dl2u(i, j) - dl2ujb(i, j));
      }
    }
    for (int j = 1; j <= jj; ++j)
      for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) {
        if (isopyc) stress(i, j) = k == 1 ? -2. * taux(i, j) * grav * scux(i, j) / (p(i, j, 2) + p(i - 1, j, 2)) : 0.;
        else stress(i, j) = -(mu_nonloc(i, j, k) - mu_nonloc(i, j, k + 1)) * taux(i, j) * grav * scux(i, j) /
                            std::max(onemm, dpu(i, j, km));
      }
    for (int j = 1; j <= jj; ++j)
      for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) {
        double ptopl = .5 * (std::min(pbu(i, j, m), p(i, j, k)) + std::min(pbu(i, j, m), p(i - 1, j, k)));
        double pbotl = .5 * (std::min(pbu(i, j, m), p(i, j, k + 1)) + std::min(pbu(i, j, m), p(i - 1, j, k + 1)));
        double q = .5 * (drag(i, j) + drag(i - 1, j)) *
                   (std::max(pbu(i, j, m) - thkbop, pbotl) - std::max(pbu(i, j, m) - thkbop, std::min(ptopl, pbotl - onemm))) /
                   std::max(dpu(i, j, km), onemm);
        double botstr = -utotn(i, j) * q / (1. + delt1 * q);
        double pgf = (1. - 2. * wpgf) * pgfx(i, j, km) + wpgf * (pgfx_o(i, j, k) + pgfx(i, j, kn));
        u(i, j, km) = u(i, j, km) * (wuv1 * dpu(i, j, km) + onemm) + u(i, j, kn) * wuv2 * dpuold(i, j, k);
        u(i, j, kn) = u(i, j, kn) +
                      delt1 * (-scuxi(i, j) * (-pgf + stress(i, j) + (ke(i, j) - ke(i - 1, j))) + cau(i, j) - ubrhs(i, j) +
                               botstr -
                               (uflux1(i, j) - uflux1(i - 1, j) + uflux3(i, j) - uflux2(i, j)) /
                                   (scu2(i, j) * std::max(dpu(i, j, km), onemm)));
      }
    // ---------- v equation ---------- :988-1143
    for (int j = 0; j <= jj + 1; ++j)
      for (int i = 0; i <= ii + 1; ++i) if (iv(i, j) == 1) {
        double q = .5 * (difwgt(i, j - 1) + difwgt(i, j));
        double deform = std::sqrt(.5 * (defor1(i, j) + defor1(i, j - 1) + defor2(i, j) + defor2(i + 1, j)));
        vsc2(i, j) = std::max(q * mdv2hi + (1. - q) * mdv2lo, (q * vsc2hi + (1. - q) * vsc2lo) * deform);
        vsc4(i, j) = std::max(q * mdv4hi + (1. - q) * mdv4lo, (q * vsc4hi + (1. - q) * vsc4lo) * deform);
      }
    for (int i = 0; i <= ii + 1; ++i)
      for (int l = 1; l <= JV.count(i); ++l) {
        int j = JV.first(i, l);
        if (j > 0) { vsc2(i, j - 1) = vsc2(i, j); vsc4(i, j - 1) = vsc4(i, j); }
        j = JV.last(i, l);
        if (j < jj + 1) { vsc2(i, j + 1) = vsc2(i, j); vsc4(i, j + 1) = vsc4(i, j); }
      }
    for (int j = 0; j <= jj; ++j)
      for (int i = 1; i <= ii; ++i) if (ip(i, j) == 1) {
        if (iv(i, j) + iv(i, j + 1) > 0) {
          double dpxy = std::max(dpv(i, j, km), onemm), dpjb = std::max(dpv(i, j + 1, km), onemm);
          vflux1(i, j) = std::min(difmxp(i, j), (vsc2(i, j) + vsc2(i, j + 1)) * scpx(i, j)) * hfharm(dpxy, dpjb) *
                             (vtotn(i, j) - vtotn(i, j + 1)) +
                         std::min(.125 * difmxp(i, j), (vsc4(i, j) + vsc4(i, j + 1)) * scpx(i, j)) * hfharm(dpxy, dpjb) *
                             (dl2v(i, j) - dl2v(i, j + 1));
        }
      }
    for (int j = 1; j <= jj; ++j)
      for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) {
        double dpxy = std::max(dpv(i, j, km), onemm);
        double dpia = std::max(dpv(i - 1, j, km), onemm);
        dpia = dpia + wgtia(i, j) * (dpxy - dpia);
        double dpib = std::max(dpv(i + 1, j, km), onemm);
        dpib = dpib + wgtib(i, j) * (dpxy - dpib);
        double vsc2a, vsc4a, vsc2b, vsc4b;
        if (iv(i - 1, j) == 0) { vsc2a = vsc2(i, j); vsc4a = vsc4(i, j); } else { vsc2a = vsc2(i - 1, j); vsc4a = vsc4(i - 1, j); }
        if (iv(i + 1, j) == 0) { vsc2b = vsc2(i, j); vsc4b = vsc4(i, j); } else { vsc2b = vsc2(i + 1, j); vsc4b = vsc4(i + 1, j); }
        vflux2(i, j) = std::min(difmxq(i, j), (vsc2(i, j) + vsc2a) * scqy(i, j)) * hfharm(dpia, dpxy) * (via(i, j) - vtotn(i, j)) +
                       std::min(.125 * difmxq(i, j), (vsc4(i, j) + vsc4a) * scqy(i, j)) * hfharm(dpia, dpxy) *
                           (dl2via(i, j) - dl2v(i, j));
        vflux3(i, j) = std::min(difmxq(i + 1, j), (vsc2(i, j) + vsc2b) * scqy(i + 1, j)) * hfharm(dpib, dpxy) *
                           (vtotn(i, j) - vib(i, j)) +
                       std::min(.125 * difmxq(i + 1, j), (vsc4(i, j) + vsc4b) * scqy(i + 1, j)) * hfharm(dpib, dpxy) *
                           (dl2v(i, j) - dl2vib(i, j));
      }
    for (int j = 1; j <= jj; ++j)
      for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) {
        if (isopyc) stress(i, j) = k == 1 ? -2. * tauy(i, j) * grav * scvy(i, j) / (p(i, j, 2) + p(i, j - 1, 2)) : 0.;
        else stress(i, j) = -(mv_nonloc(i, j, k) - mv_nonloc(i, j, k + 1)) * tauy(i, j) * grav * scvy(i, j) /
                            std::max(onemm, dpv(i, j, km));
      }
    for (int j = 1; j <= jj; ++j)
      for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) {
        double ptopl = .5 * (std::min(pbv(i, j, m), p(i, j, k)) + std::min(pbv(i, j, m), p(i, j - 1, k)));
        double pbotl = .5 * (std::min(pbv(i, j, m), p(i, j, k + 1)) + std::min(pbv(i, j, m), p(i, j - 1, k + 1)));
        double q = .5 * (drag(i, j) + drag(i, j - 1)) *
                   (std::max(pbv(i, j, m) - thkbop, pbotl) - std::max(pbv(i, j, m) - thkbop, std::min(ptopl, pbotl - onemm))) /
                   std::max(dpv(i, j, km), onemm);
        double botstr = -vtotn(i, j) * q / (1. + delt1 * q);
        double pgf = (1. - 2. * wpgf) * pgfy(i, j, km) + wpgf * (pgfy_o(i, j, k) + pgfy(i, j, kn));
        v(i, j, km) = v(i, j, km) * (wuv1 * dpv(i, j, km) + onemm) + v(i, j, kn) * wuv2 * dpvold(i, j, k);
        v(i, j, kn) = v(i, j, kn) +
                      delt1 * (-scvyi(i, j) * (-pgf + stress(i, j) + (ke(i, j) - ke(i, j - 1))) + cav(i, j) - vbrhs(i, j) +
                               botstr -
                               (vflux1(i, j) - vflux1(i, j - 1) + vflux3(i, j) - vflux2(i, j)) /
                                   (scv2(i, j) * std::max(dpv(i, j, km), onemm)));
      }
  }  // k

  // :1154-1197 massless-layer fill, clamp, depth mean
  for (int j = 1; j <= jj; ++j) {
    for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) utotn(i, j) = 0.;
    for (int k = 1; k <= kk; ++k) {
      const int km = k + mm, kn = k + nn, kan = std::max(1, k - 1) + nn;
      for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) {
        double q = std::min(std::min(dpu(i, j, km), dpu(i, j, kn)), onem);
        u(i, j, kn) = (u(i, j, kn) * q + u(i, j, kan) * (onem - q)) / onem;
        u(i, j, kn) = std::max(-umax(i, j), std::min(umax(i, j), u(i, j, kn) + ub(i, j, m))) - ub(i, j, m);
        utotn(i, j) = utotn(i, j) + u(i, j, kn) * dpu(i, j, kn);
      }
    }
    for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) utotn(i, j) = utotn(i, j) / pbu_p(i, j);
    for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) vtotn(i, j) = 0.;
    for (int k = 1; k <= kk; ++k) {
      const int km = k + mm, kn = k + nn, kan = std::max(1, k - 1) + nn;
      for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) {
        double q = std::min(std::min(dpv(i, j, km), dpv(i, j, kn)), onem);
        v(i, j, kn) = (v(i, j, kn) * q + v(i, j, kan) * (onem - q)) / onem;
        v(i, j, kn) = std::max(-vmax(i, j), std::min(vmax(i, j), v(i, j, kn) + vb(i, j, m))) - vb(i, j, m);
        vtotn(i, j) = vtotn(i, j) + v(i, j, kn) * dpv(i, j, kn);
      }
    }
    for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) vtotn(i, j) = vtotn(i, j) / pbv_p(i, j);
  }
  // :1202-1232 time smoothing part 2
  for (int k = 1; k <= kk; ++k) {
    const int km = k + mm, kn = k + nn;
    for (int j = 1; j <= jj; ++j)
      for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) {
        u(i, j, kn) = u(i, j, kn) - utotn(i, j);
        u(i, j, km) = (u(i, j, km) + u(i, j, kn) * wuv2 * dpu(i, j, kn)) /
                      (wuv1 * dpu(i, j, km) + onemm + wuv2 * (dpuold(i, j, k) + dpu(i, j, kn)));
      }
    for (int j = 1; j <= jj; ++j)
      for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) {
        v(i, j, kn) = v(i, j, kn) - vtotn(i, j);
        v(i, j, km) = (v(i, j, km) + v(i, j, kn) * wuv2 * dpv(i, j, kn)) /
                      (wuv1 * dpv(i, j, km) + onemm + wuv2 * (dpvold(i, j, k) + dpv(i, j, kn)));
      }
  }
  for (int j = 1; j <= jj; ++j) {
    for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) utotn(i, j) = utotn(i, j) * dt1inv;
    for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) vtotn(i, j) = vtotn(i, j) * dt1inv;
  }
  // :1252-1267
  for (int j = 1; j <= jj; ++j)
    for (int k = 1; k <= kk; ++k) {
      const int kn = k + nn;
      for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) pu(i, j, k + 1) = pu(i, j, k) + dpu(i, j, kn);
      for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) pv(i, j, k + 1) = pv(i, j, k) + dpv(i, j, kn);
    }
}

}  // namespace orc
