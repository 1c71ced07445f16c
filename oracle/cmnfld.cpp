// ORACLE — TEST INFRASTRUCTURE ONLY (see core.hpp header).
// Restatement of the producers of the neutral slope that eddtra / ndiff consume
// (SURVEY.md §8f rank 3):
//   cmnfld_bfsqf_ale    phy/mod_cmnfld_routines.F90:229-350
//   cmnfld_nslope_ale   phy/mod_cmnfld_routines.F90:654-811
//   cmnfld_nnslope_ale  phy/mod_cmnfld_routines.F90:813-883
//   cmnfld2             phy/mod_cmnfld_routines.F90:1158-1238 (hybrid / ALE branch)
// Constants sls0, bfsqmn: phy/mod_cmnfld.F90:36,46.
#include "core.hpp"
#include "eos.hpp"

namespace orc {

// phy/mod_cmnfld_routines.F90:229-350
void cmnfld_bfsqf_ale(int m, int n, int mm, int nn, int k1m, int k1n) {
  (void)m; (void)n; (void)mm; (void)k1m; (void)k1n;
  Oracle& o = O(); const Dims& d = o.d;
  const int ii = d.ii, jj = d.jj, kk = d.kk;
  if (kk < 2) throw std::runtime_error("cmnfld_bfsqf_ale: kk >= 2 required");
  const double sls0 = o.scalar("sls0", 10. * onem), bfsqmn = o.scalar("bfsqmn", 1.e-7);
  A3 p = o.a3("p"), dp = o.a3("dp"), temp = o.a3("temp"), saln = o.a3("saln");
  A3 bfsqi = o.a3("bfsqi"), bfsql = o.a3("bfsql"), bfsqf = o.a3("bfsqf");
  I2 ip = o.i2("ip");
  std::fill(bfsqi.p, bfsqi.p + d.lev * (size_t)(kk + 1), 0.0);  // :247
  std::fill(bfsql.p, bfsql.p + d.lev * (size_t)kk, 0.0);        // :248
  std::vector<double> delp(kk + 1), bfsq(kk + 1), sls2(kk + 1), atd(kk + 1), btd(kk + 1), ctd(kk + 1), rtd(kk + 1),
      gam(kk + 1);
  for (int j = -1; j <= jj + 2; ++j)
    for (int i = -1; i <= ii + 2; ++i) {
      if (ip(i, j) != 1) continue;
      bfsqi(i, j, 1) = bfsqmn;
      double pup = .5 * (p(i, j, 1) + p(i, j, 2));
      double tup = temp(i, j, 1 + nn), sup = saln(i, j, 1 + nn);
      for (int k = 2; k <= kk; ++k) {
        const int kn = k + nn;
        if (p(i, j, kk + 1) - p(i, j, k) < epsilp) {
          delp[k] = onemm;
          bfsqi(i, j, k) = bfsqi(i, j, k - 1);
          bfsq[k] = bfsqmn;
          sls2[k] = sls0 * sls0;
        } else {
          double plo;
          if (p(i, j, kk + 1) - p(i, j, k + 1) < epsilp) plo = p(i, j, kk + 1);
          else plo = .5 * (p(i, j, k) + p(i, j, k + 1));
          const double tlo = temp(i, j, kn), slo = saln(i, j, kn);
          delp[k] = std::max(onemm, plo - pup);
          bfsqi(i, j, k) = grav * grav * (eos::rho(p(i, j, k), tlo, slo) - eos::rho(p(i, j, k), tup, sup)) / delp[k];
          bfsq[k] = std::max(bfsqmn, bfsqi(i, j, k));
          bfsqi(i, j, k) = bfsqi(i, j, k) * delp[k] / std::max(onem, delp[k]);
          if (p(i, j, kk + 1) - p(i, j, k) < onem) bfsqi(i, j, k) = bfsqi(i, j, k - 1);
          sls2[k] = sls0 * sls0;
          pup = plo; tup = tlo; sup = slo;
        }
      }
      delp[1] = dp(i, j, 1 + nn);
      bfsqi(i, j, 1) = bfsqi(i, j, 2);
      bfsq[1] = std::max(bfsqmn, bfsqi(i, j, 1));
      sls2[1] = sls0 * sls0;
      for (int k = 1; k <= kk - 1; ++k) bfsql(i, j, k) = .5 * (bfsqi(i, j, k) + bfsqi(i, j, k + 1));
      bfsql(i, j, kk) = bfsqi(i, j, kk);
      int k = 1;
      ctd[k] = -2. * sls2[k] / (delp[k] * (delp[k] + delp[k + 1]));
      btd[k] = 1. - ctd[k];
      rtd[k] = bfsq[k];
      for (k = 2; k <= kk - 1; ++k) {
        atd[k] = -2. * sls2[k - 1] / (delp[k] * (delp[k - 1] + delp[k]));
        ctd[k] = -2. * sls2[k] / (delp[k] * (delp[k] + delp[k + 1]));
        btd[k] = 1. - atd[k] - ctd[k];
        rtd[k] = bfsq[k];
      }
      k = kk;
      atd[k] = -2. * sls2[k - 1] / (delp[k] * (delp[k - 1] + delp[k]));
      btd[k] = 1. - atd[k];
      rtd[k] = bfsq[k];
      double bei = 1. / btd[1];
      bfsqf(i, j, 1) = rtd[1] * bei;
      for (k = 2; k <= kk; ++k) {
        gam[k] = ctd[k - 1] * bei;
        bei = 1. / (btd[k] - atd[k] * gam[k]);
        bfsqf(i, j, k) = (rtd[k] - atd[k] * bfsqf(i, j, k - 1)) * bei;
      }
      for (k = kk - 1; k >= 1; --k) bfsqf(i, j, k) = bfsqf(i, j, k) - gam[k + 1] * bfsqf(i, j, k + 1);
      bfsqi(i, j, kk + 1) = bfsqi(i, j, kk);
      bfsqf(i, j, kk + 1) = bfsqf(i, j, kk);
    }
}

// phy/mod_cmnfld_routines.F90:654-811
void cmnfld_nslope_ale(int m, int n, int mm, int nn, int k1m, int k1n) {
  (void)m; (void)n; (void)mm; (void)k1m; (void)k1n;
  Oracle& o = O(); const Dims& d = o.d;
  const int ii = d.ii, jj = d.jj, kk = d.kk;
  A3 p = o.a3("p"), dp = o.a3("dp"), temp = o.a3("temp"), saln = o.a3("saln"), phi = o.a3("phi");
  A3 bfsqf = o.a3("bfsqf"), nslpx = o.a3("nslpx"), nslpy = o.a3("nslpy"), nnslpx = o.a3("nnslpx"),
     nnslpy = o.a3("nnslpy");
  A2 scuxi = o.a2("scuxi"), scvyi = o.a2("scvyi");
  I2 ip = o.i2("ip"), iu = o.i2("iu"), iv = o.i2("iv");
  // geopotential at layer interfaces (:669-685)
  for (int j = -1; j <= jj + 2; ++j)
    for (int k = kk; k >= 1; --k) {
      const int kn = k + nn;
      for (int i = -1; i <= ii + 2; ++i) {
        if (ip(i, j) != 1) continue;
        if (dp(i, j, kn) < epsilp) phi(i, j, k) = phi(i, j, k + 1);
        else phi(i, j, k) = phi(i, j, k + 1) - eos::p_alpha(p(i, j, k + 1), p(i, j, k), temp(i, j, kn), saln(i, j, kn));
      }
    }
  // x-component (:696-747)
  for (int j = -1; j <= jj + 2; ++j)
    for (int i = 0; i <= ii + 2; ++i) {
      if (iu(i, j) != 1) continue;
      for (int k = 1; k <= kk; ++k) { nslpx(i, j, k) = 0.; nnslpx(i, j, k) = 0.; }
      int kmax = 1;
      for (int k = 2; k <= kk; ++k) {
        const int kn = k + nn;
        if (dp(i - 1, j, kn) > epsilp || dp(i, j, kn) > epsilp) kmax = k;
      }
      int knnsl = 2;
      for (int k = 2; k <= kmax; ++k) {
        const int kn = k + nn;
        const double pm = .5 * (p(i - 1, j, k) + p(i, j, k));
        const double rho_x = .5 * (eos::rho(pm, temp(i, j, kn - 1), saln(i, j, kn - 1)) -
                                   eos::rho(pm, temp(i - 1, j, kn - 1), saln(i - 1, j, kn - 1)) +
                                   eos::rho(pm, temp(i, j, kn), saln(i, j, kn)) -
                                   eos::rho(pm, temp(i - 1, j, kn), saln(i - 1, j, kn)));
        const double phi_x = phi(i, j, k) - phi(i - 1, j, k);
        const double bfsqm = .5 * (bfsqf(i - 1, j, k) + bfsqf(i, j, k));
        nslpx(i, j, k) = (grav * rho_x / (rho0 * bfsqm) + phi_x / grav) * scuxi(i, j);
        if (phi(i, j, k) > phi(i - 1, j, kk + 1) && phi(i - 1, j, k) > phi(i, j, kk + 1)) {
          nnslpx(i, j, k) = std::sqrt(bfsqm) * nslpx(i, j, k);
          knnsl = k;
        }
      }
      for (int k = knnsl + 1; k <= kmax; ++k) nnslpx(i, j, k) = nnslpx(i, j, knnsl);
    }
  // y-component (:751-798)
  for (int j = 0; j <= jj + 2; ++j)
    for (int i = -1; i <= ii + 2; ++i) {
      if (iv(i, j) != 1) continue;
      for (int k = 1; k <= kk; ++k) { nslpy(i, j, k) = 0.; nnslpy(i, j, k) = 0.; }
      int kmax = 1;
      for (int k = 2; k <= kk; ++k) {
        const int kn = k + nn;
        if (dp(i, j - 1, kn) > epsilp || dp(i, j, kn) > epsilp) kmax = k;
      }
      int knnsl = 2;
      for (int k = 2; k <= kmax; ++k) {
        const int kn = k + nn;
        const double pm = .5 * (p(i, j - 1, k) + p(i, j, k));
        const double rho_y = .5 * (eos::rho(pm, temp(i, j, kn - 1), saln(i, j, kn - 1)) -
                                   eos::rho(pm, temp(i, j - 1, kn - 1), saln(i, j - 1, kn - 1)) +
                                   eos::rho(pm, temp(i, j, kn), saln(i, j, kn)) -
                                   eos::rho(pm, temp(i, j - 1, kn), saln(i, j - 1, kn)));
        const double phi_y = phi(i, j, k) - phi(i, j - 1, k);
        const double bfsqm = .5 * (bfsqf(i, j - 1, k) + bfsqf(i, j, k));
        nslpy(i, j, k) = (grav * rho_y / (rho0 * bfsqm) + phi_y / grav) * scvyi(i, j);
        if (phi(i, j, k) > phi(i, j - 1, kk + 1) && phi(i, j - 1, k) > phi(i, j, kk + 1)) {
          nnslpy(i, j, k) = std::sqrt(bfsqm) * nslpy(i, j, k);
          knnsl = k;
        }
      }
      for (int k = knnsl + 1; k <= kmax; ++k) nnslpy(i, j, k) = nnslpy(i, j, knnsl);
    }
}

// phy/mod_cmnfld_routines.F90:813-883
void cmnfld_nnslope_ale(int m, int n, int mm, int nn, int k1m, int k1n) {
  (void)m; (void)n; (void)mm; (void)nn; (void)k1m; (void)k1n;
  Oracle& o = O(); const Dims& d = o.d;
  const int ii = d.ii, jj = d.jj, kk = d.kk;
  A3 p = o.a3("p"), bfsqf = o.a3("bfsqf"), nslpx = o.a3("nslpx"), nslpy = o.a3("nslpy"),
     nnslpx = o.a3("nnslpx"), nnslpy = o.a3("nnslpy");
  I2 iu = o.i2("iu"), iv = o.i2("iv");
  xctilr(nslpx, 1, kk, 2, 2, halo_uv);
  xctilr(nslpy, 1, kk, 2, 2, halo_vv);
  for (int j = -1; j <= jj + 2; ++j)
    for (int i = 0; i <= ii + 2; ++i) {
      if (iu(i, j) != 1) continue;
      int knnsl = 1;
      nnslpx(i, j, 1) = 0.;
      for (int k = 2; k <= kk; ++k) {
        if (p(i, j, k) < p(i - 1, j, kk + 1) && p(i - 1, j, k) < p(i, j, kk + 1)) {
          const double bfsqm = .5 * (bfsqf(i - 1, j, k) + bfsqf(i, j, k));
          nnslpx(i, j, k) = std::sqrt(bfsqm) * nslpx(i, j, k);
          knnsl = k;
        } else {
          break;
        }
      }
      for (int k = knnsl + 1; k <= kk; ++k) nnslpx(i, j, k) = nnslpx(i, j, knnsl);
    }
  for (int j = 0; j <= jj + 2; ++j)
    for (int i = -1; i <= ii + 2; ++i) {
      if (iv(i, j) != 1) continue;
      int knnsl = 1;
      nnslpy(i, j, 1) = 0.;
      for (int k = 2; k <= kk; ++k) {
        if (p(i, j, k) < p(i, j - 1, kk + 1) && p(i, j - 1, k) < p(i, j, kk + 1)) {
          const double bfsqm = .5 * (bfsqf(i, j - 1, k) + bfsqf(i, j, k));
          nnslpy(i, j, k) = std::sqrt(bfsqm) * nslpy(i, j, k);
          knnsl = k;
        } else {
          break;
        }
      }
      for (int k = knnsl + 1; k <= kk; ++k) nnslpy(i, j, k) = nnslpy(i, j, knnsl);
    }
}

// phy/mod_cmnfld_routines.F90:1158-1238, vcoord /= 'isopyc_bulkml' (the hybrid default of every named
// grid).  Options: edritp ('large scale' default), eitmth ('gm' default), ltedtp ('layer' default).
void cmnfld2(int m, int n, int mm, int nn, int k1m, int k1n) {
  Oracle& o = O(); const Dims& d = o.d;
  if (o.option("vcoord", "cntiso_hybrid") == "isopyc_bulkml")
    throw std::runtime_error(" cmnfld2: vcoord = isopyc_bulkml is unsupported!");
  A3 temp = o.a3("temp"), saln = o.a3("saln");
  xctilr(temp, 1, 2 * d.kk, 3, 3, halo_ps);
  xctilr(saln, 1, 2 * d.kk, 3, 3, halo_ps);
  cmnfld_bfsqf_ale(m, n, mm, nn, k1m, k1n);
  if (o.option("edritp", "large scale") == "large scale" || o.option("eitmth", "gm") == "gm") {
    if (o.option("ltedtp", "layer") == "neutral") cmnfld_nnslope_ale(m, n, mm, nn, k1m, k1n);
    else cmnfld_nslope_ale(m, n, mm, nn, k1m, k1n);
  }
}

}  // namespace orc
