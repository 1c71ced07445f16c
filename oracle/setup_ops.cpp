// ORACLE — TEST INFRASTRUCTURE ONLY (see core.hpp header).
// Restatement of numerical_bounds (phy/mod_blom_init.F90:446-555) and init_fluxes
// (phy/mod_state.F90:341-383, update_flux_halos=.true. as in phy/mod_blom_step.F90), and of the
// conservation diagnostics budget_init / budget_sums (phy/mod_budget.F90:74-196).
#include "core.hpp"

namespace orc {

void numerical_bounds() {
  Oracle& o = O(); const Dims& d = o.d;
  const int ii = d.ii, jj = d.jj, nb = d.nbdy;
  const double baclin = o.scalar("baclin");
  A2 scqx = o.a2("scqx"), scqy = o.a2("scqy"), scpx = o.a2("scpx"), scpy = o.a2("scpy"), scuy = o.a2("scuy"),
     scvx = o.a2("scvx"), scp2 = o.a2("scp2"), depths = o.a2("depths");
  A2 difmxp = o.a2("difmxp"), difmxq = o.a2("difmxq"), umax = o.a2("umax"), vmax = o.a2("vmax");
  I2 ip = o.i2("ip"), iu = o.i2("iu"), iv = o.i2("iv");
  for (int j = 1 - nb; j <= jj + nb; ++j)
    for (int i = 1 - nb; i <= ii + nb; ++i) {
      double dx2 = scpx(i, j) * scpx(i, j), dy2 = scpy(i, j) * scpy(i, j);
      difmxp(i, j) = .9 * .5 * dx2 * dy2 / std::max(1., (dx2 + dy2) * (baclin + baclin));
      dx2 = scqx(i, j) * scqx(i, j); dy2 = scqy(i, j) * scqy(i, j);
      difmxq(i, j) = .9 * .5 * dx2 * dy2 / std::max(1., (dx2 + dy2) * (baclin + baclin));
    }
  double btdtmx = 86400.;
  for (int j = 1; j <= jj; ++j)
    for (int i = 1; i <= ii; ++i) if (ip(i, j) == 1)
      btdtmx = std::min(btdtmx, scpx(i, j) * scpy(i, j) /
                                    std::sqrt(grav * depths(i, j) * (scpx(i, j) * scpx(i, j) + scpy(i, j) * scpy(i, j))));
  o.sc["btdtmx"] = btdtmx / std::sqrt(2.);
  for (int j = 1; j <= jj; ++j) {
    for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1)
      umax(i, j) = .9 * .125 * std::min(scp2(i - 1, j), scp2(i, j)) / (scuy(i, j) * baclin);
    for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1)
      vmax(i, j) = .9 * .125 * std::min(scp2(i, j - 1), scp2(i, j)) / (scvx(i, j) * baclin);
  }
  xctilr(umax, nb, nb, halo_us);
  xctilr(vmax, nb, nb, halo_vs);
}

void init_fluxes(int m, int n, int mm, int nn, int k1m, int k1n) {
  (void)m; (void)n; (void)nn; (void)k1m;
  Oracle& o = O(); const Dims& d = o.d;
  const int ii = d.ii, jj = d.jj, kk = d.kk;
  A3 uflx = o.a3("uflx"), utflx = o.a3("utflx"), usflx = o.a3("usflx"), vflx = o.a3("vflx"), vtflx = o.a3("vtflx"),
     vsflx = o.a3("vsflx");
  I2 iu = o.i2("iu"), iv = o.i2("iv");
  for (int j = 0; j <= jj + 2; ++j)
    for (int k = 1; k <= kk; ++k) {
      const int km = k + mm;
      for (int i = 0; i <= ii + 2; ++i) if (iu(i, j) == 1) { uflx(i, j, km) = 0.; utflx(i, j, km) = 0.; usflx(i, j, km) = 0.; }
      for (int i = 0; i <= ii + 2; ++i) if (iv(i, j) == 1) { vflx(i, j, km) = 0.; vtflx(i, j, km) = 0.; vsflx(i, j, km) = 0.; }
    }
  xctilr(uflx.from(k1n), 1, kk, 1, 1, halo_uv); xctilr(utflx.from(k1n), 1, kk, 1, 1, halo_uv);
  xctilr(usflx.from(k1n), 1, kk, 1, 1, halo_uv); xctilr(vflx.from(k1n), 1, kk, 1, 1, halo_vv);
  xctilr(vtflx.from(k1n), 1, kk, 1, 1, halo_vv); xctilr(vsflx.from(k1n), 1, kk, 1, 1, halo_vv);
}

// budget_init (phy/mod_budget.F90:74-93): global mass, xcsum of pb(:,:,1)*scp2
void budget_init(double* mass0) {
  Oracle& o = O(); const Dims& d = o.d;
  A2 util1 = (o.has("util1") ? o.a3("util1") : o.scratch("util1", 1)).level(1), scp2 = o.a2("scp2");
  A3 pb = o.a3("pb"); I2 ip = o.i2("ip");
  for (int j = 1; j <= d.jj; ++j)
    for (int i = 1; i <= d.ii; ++i) if (ip(i, j) == 1) util1(i, j) = pb(i, j, 1) * scp2(i, j);
  *mass0 = xcsum(util1, ip);
}

// budget_sums (phy/mod_budget.F90:95-196): column sums in k order, then the strip-ordered xcsum.
// out[0]=sdp  out[1]=tdp  out[2]=trdp (1st tracer, if ntr>0)  out[3]=sc (salt_corr, on the call the
// reference evaluates it: ncall 4, or 5 for vcoord='isopyc_bulkml'); entries not evaluated are untouched.
void budget_sums(int ncall, int n, int nn, double* out) {
  (void)n;
  Oracle& o = O(); const Dims& d = o.d;
  const int ii = d.ii, jj = d.jj, kk = d.kk;
  // util1/util2 are the module work arrays of mod_utility; bound host arrays are used when registered
  A2 util1 = (o.has("util1") ? o.a3("util1") : o.scratch("util1", 1)).level(1),
     util2 = (o.has("util2") ? o.a3("util2") : o.scratch("util2", 1)).level(1), scp2 = o.a2("scp2");
  A3 dp = o.a3("dp"), temp = o.a3("temp"), saln = o.a3("saln"); I2 ip = o.i2("ip");
  for (int j = 1; j <= jj; ++j)
    for (int i = 1; i <= ii; ++i) if (ip(i, j) == 1) { util1(i, j) = 0.; util2(i, j) = 0.; }
  for (int j = 1; j <= jj; ++j)
    for (int k = 1; k <= kk; ++k) {
      const int kn = k + nn;
      for (int i = 1; i <= ii; ++i) if (ip(i, j) == 1) {
        const double q = dp(i, j, kn) * scp2(i, j);
        util1(i, j) = util1(i, j) + saln(i, j, kn) * q;
        util2(i, j) = util2(i, j) + temp(i, j, kn) * q;
      }
    }
  out[0] = xcsum(util1, ip);
  out[1] = xcsum(util2, ip);
  if (d.ntr > 0) {
    A3 trc = o.a3("trc");
    for (int j = 1; j <= jj; ++j)
      for (int i = 1; i <= ii; ++i) if (ip(i, j) == 1) util1(i, j) = 0.;
    for (int j = 1; j <= jj; ++j)
      for (int k = 1; k <= kk; ++k) {
        const int kn = k + nn;
        for (int i = 1; i <= ii; ++i) if (ip(i, j) == 1) {
          const double q = dp(i, j, kn) * scp2(i, j);
          util1(i, j) = util1(i, j) + trc(i, j, kn) * q;   // trc(:,:,kn,1)
        }
      }
    out[2] = xcsum(util1, ip);
  }
  const bool isopyc = o.option("vcoord", "cntiso_hybrid") == "isopyc_bulkml";
  if (((isopyc && ncall == 5) || (!isopyc && ncall == 4)) && o.f.count("salt_corr")) {
    A2 salt_corr = o.a2("salt_corr");
    for (int j = 1; j <= jj; ++j)
      for (int i = 1; i <= ii; ++i) if (ip(i, j) == 1) util1(i, j) = salt_corr(i, j) * scp2(i, j);
    out[3] = xcsum(util1, ip);
  }
}

}  // namespace orc
