// ORACLE — TEST INFRASTRUCTURE ONLY (see core.hpp header).
// Restatement of numerical_bounds (phy/mod_blom_init.F90:446-555) and init_fluxes
// (phy/mod_state.F90:341-383, update_flux_halos=.true. as in phy/mod_blom_step.F90).
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

}  // namespace orc
