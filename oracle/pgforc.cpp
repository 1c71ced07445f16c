// ORACLE — TEST INFRASTRUCTURE ONLY (see core.hpp header).
// Restatement of phy/mod_pgforc.F90: pgforc :438-615, pgforc_dynamic_enthalpy :262-408
// (pgfmth='dynamic enthalpy') and pgforc_geopotential :95-260 (pgfmth='geopotential').
#include "core.hpp"
#include "eos.hpp"

namespace orc {

static void pgforc_dynamic_enthalpy(int m, int n, int mm, int nn, int k1m, int k1n) {
  (void)m; (void)mm; (void)k1m; (void)k1n;
  Oracle& o = O(); const Dims& d = o.d;
  const int ii = d.ii, jj = d.jj, kk = d.kk;
  const double p0_dynh = 0.0;  // :49
  const double pref = eos::K().pref;
  A3 p = o.a3("p"), phi = o.a3("phi"), temp = o.a3("temp"), saln = o.a3("saln"), dp = o.a3("dp");
  A3 dpu = o.a3("dpu"), dpv = o.a3("dpv"), pgfx = o.a3("pgfx"), pgfy = o.a3("pgfy");
  A3 pgfxm = o.a3("pgfxm"), pgfym = o.a3("pgfym"), xixp = o.a3("xixp"), xixm = o.a3("xixm"),
     xiyp = o.a3("xiyp"), xiym = o.a3("xiym");
  I2 ip = o.i2("ip"), iu = o.i2("iu"), iv = o.i2("iv");
  A3 pot_dynh = o.scratch("_pot_dynh", kk), pot_dynh_pb = o.scratch("_pot_dynh_pb", kk),
     dynh_a = o.scratch("_dynh_a", kk), dynh_t = o.scratch("_dynh_t", kk), alpha_r = o.scratch("_alpha_r", kk);
  using namespace eos;
  for (int j = 0; j <= jj; ++j) {
    int kn = kk + nn;
    for (int i = 0; i <= ii; ++i) if (ip(i, j) == 1) {
      pot_dynh(i, j, kk) = phi(i, j, kk + 1) + p_alpha(p0_dynh, p(i, j, kk + 1), temp(i, j, kn), saln(i, j, kn));
      pot_dynh_pb(i, j, kk) = alp(p(i, j, kk + 1), temp(i, j, kn), saln(i, j, kn)) * p(i, j, kk + 1);
      phi(i, j, kk) = phi(i, j, kk + 1) + p_alpha(p(i, j, kk), p(i, j, kk + 1), temp(i, j, kn), saln(i, j, kn));
    }
    for (int k = kk - 1; k >= 1; --k) {
      kn = k + nn;
      for (int i = 0; i <= ii; ++i) if (ip(i, j) == 1) {
        pot_dynh(i, j, k) = pot_dynh(i, j, k + 1) +
                            p_alpha(p0_dynh, p(i, j, k + 1), temp(i, j, kn), saln(i, j, kn)) -
                            p_alpha(p0_dynh, p(i, j, k + 1), temp(i, j, kn + 1), saln(i, j, kn + 1));
        pot_dynh_pb(i, j, k) = pot_dynh_pb(i, j, k + 1) +
                               (alp(p(i, j, k + 1), temp(i, j, kn), saln(i, j, kn)) -
                                alp(p(i, j, k + 1), temp(i, j, kn + 1), saln(i, j, kn + 1))) * p(i, j, k + 1);
        phi(i, j, k) = phi(i, j, k + 1) + p_alpha(p(i, j, k), p(i, j, k + 1), temp(i, j, kn), saln(i, j, kn));
      }
    }
    for (int k = 1; k <= kk; ++k) {
      kn = k + nn;
      for (int i = 0; i <= ii; ++i) if (ip(i, j) == 1) {
        if (dp(i, j, kn) < onemm) {
          dynh_a(i, j, k) = 0.; dynh_t(i, j, k) = 0.;
        } else {
          double dynh_ts_t, dynh_ts_s;
          dynh_derivatives(p0_dynh, p(i, j, k), p(i, j, k + 1), temp(i, j, kn), saln(i, j, kn), dynh_ts_t, dynh_ts_s);
          dynh_a(i, j, k) = dynh_ts_s / dalpds(pref, temp(i, j, kn), saln(i, j, kn));
          dynh_t(i, j, k) = dynh_ts_t - dynh_a(i, j, k) * dalpdt(pref, temp(i, j, kn), saln(i, j, kn));
        }
        alpha_r(i, j, k) = alp(pref, temp(i, j, kn), saln(i, j, kn));
      }
    }
  }
  for (int j = 1; j <= jj; ++j) {
    for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) { xixp(i, j, n) = 0.; xixm(i, j, n) = 0.; pgfxm(i, j, n) = 0.; }
    for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) { xiyp(i, j, n) = 0.; xiym(i, j, n) = 0.; pgfym(i, j, n) = 0.; }
    for (int k = kk; k >= 1; --k) {
      const int kn = k + nn;
      for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) {
        pgfx(i, j, kn) = -(pot_dynh(i, j, k) - pot_dynh(i - 1, j, k));
        if (dp(i - 1, j, kn) >= onemm && dp(i, j, kn) >= onemm)
          pgfx(i, j, kn) = pgfx(i, j, kn) +
                           .5 * ((dynh_t(i - 1, j, k) + dynh_t(i, j, k)) * (temp(i, j, kn) - temp(i - 1, j, kn)) +
                                 (dynh_a(i - 1, j, k) + dynh_a(i, j, k)) * (alpha_r(i, j, k) - alpha_r(i - 1, j, k)));
        pgfxm(i, j, n) = pgfxm(i, j, n) + pgfx(i, j, kn) * dpu(i, j, kn);
        xixm(i, j, n) = xixm(i, j, n) + pot_dynh_pb(i - 1, j, k) * dpu(i, j, kn);
        xixp(i, j, n) = xixp(i, j, n) + pot_dynh_pb(i, j, k) * dpu(i, j, kn);
      }
      for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) {
        pgfy(i, j, kn) = -(pot_dynh(i, j, k) - pot_dynh(i, j - 1, k));
        if (dp(i, j - 1, kn) >= onemm && dp(i, j, kn) >= onemm)
          pgfy(i, j, kn) = pgfy(i, j, kn) +
                           .5 * ((dynh_t(i, j - 1, k) + dynh_t(i, j, k)) * (temp(i, j, kn) - temp(i, j - 1, kn)) +
                                 (dynh_a(i, j - 1, k) + dynh_a(i, j, k)) * (alpha_r(i, j, k) - alpha_r(i, j - 1, k)));
        pgfym(i, j, n) = pgfym(i, j, n) + pgfy(i, j, kn) * dpv(i, j, kn);
        xiym(i, j, n) = xiym(i, j, n) + pot_dynh_pb(i, j - 1, k) * dpv(i, j, kn);
        xiyp(i, j, n) = xiyp(i, j, n) + pot_dynh_pb(i, j, k) * dpv(i, j, kn);
      }
    }
  }
}

// :95-260 (pgfmth='geopotential'): gradient of the geopotential on pressure surfaces.
static void pgforc_geopotential(int m, int n, int mm, int nn, int k1m, int k1n) {
  (void)m; (void)mm; (void)k1m; (void)k1n;
  Oracle& o = O(); const Dims& d = o.d;
  const int ii = d.ii, jj = d.jj, kk = d.kk;
  A3 p = o.a3("p"), phi = o.a3("phi"), temp = o.a3("temp"), saln = o.a3("saln"), dp = o.a3("dp");
  A3 dpu = o.a3("dpu"), dpv = o.a3("dpv"), pu = o.a3("pu"), pv = o.a3("pv"), pgfx = o.a3("pgfx"), pgfy = o.a3("pgfy");
  A3 pgfxm = o.a3("pgfxm"), pgfym = o.a3("pgfym"), xixp = o.a3("xixp"), xixm = o.a3("xixm"),
     xiyp = o.a3("xiyp"), xiym = o.a3("xiym");
  I2 ip = o.i2("ip"), iu = o.i2("iu"), iv = o.i2("iv");
  A3 phip = o.scratch("_phip", kk + 1);
  using namespace eos;
  double dphi, alpl, alpu;
  for (int j = 0; j <= jj; ++j) {
    for (int i = 0; i <= ii; ++i) if (ip(i, j) == 1) phip(i, j, kk + 1) = 0.;
    for (int k = kk; k >= 1; --k) {
      const int kn = k + nn;
      for (int i = 0; i <= ii; ++i) if (ip(i, j) == 1) {
        if (dp(i, j, kn) < epsilp) {
          phi(i, j, k) = phi(i, j, k + 1);
          phip(i, j, k) = phip(i, j, k + 1);
        } else {
          delphi(p(i, j, k), p(i, j, k + 1), temp(i, j, kn), saln(i, j, kn), dphi, alpu, alpl);
          phi(i, j, k) = phi(i, j, k + 1) - dphi;
          phip(i, j, k) = phip(i, j, k + 1) + p(i, j, k + 1) * alpl - p(i, j, k) * alpu;
        }
      }
    }
  }
  std::vector<int> kup(ii + 1), kum(ii + 1), kvp(ii + 1), kvm(ii + 1);
  double prs, dphip, dphim, alplp, alpup, alplm, alpum, cp, cm, phi_p, phi_m, q;
  for (int j = 1; j <= jj; ++j) {
    for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) {
      kup[i] = kk; kum[i] = kk;
      xixp(i, j, n) = 0.; xixm(i, j, n) = 0.; pgfxm(i, j, n) = 0.;
    }
    for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) {
      kvp[i] = kk; kvm[i] = kk;
      xiyp(i, j, n) = 0.; xiym(i, j, n) = 0.; pgfym(i, j, n) = 0.;
    }
    for (int k = kk; k >= 1; --k) {
      const int kn = k + nn;
      for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) {
        prs = pu(i, j, k + 1) - .5 * dpu(i, j, kn);
        while (p(i, j, kup[i]) > prs) kup[i] = kup[i] - 1;
        while (p(i - 1, j, kum[i]) > prs) kum[i] = kum[i] - 1;
        delphi(prs, p(i, j, kup[i] + 1), temp(i, j, kup[i] + nn), saln(i, j, kup[i] + nn), dphip, alpup, alplp);
        delphi(prs, p(i - 1, j, kum[i] + 1), temp(i - 1, j, kum[i] + nn), saln(i - 1, j, kum[i] + nn), dphim, alpum,
               alplm);
        cp = .25 * (p(i, j, k + 1) + p(i, j, k));
        cm = .25 * (p(i - 1, j, k + 1) + p(i - 1, j, k));
        q = prs / (cp + cm);
        cp = q * cp;
        cm = q * cm;
        phi_p = phi(i, j, kup[i] + 1) - dphip;
        xixp(i, j, n) = xixp(i, j, n) +
                        (phip(i, j, kup[i] + 1) + p(i, j, kup[i] + 1) * alplp - cp * (alpup - alpum)) * dpu(i, j, kn);
        phi_m = phi(i - 1, j, kum[i] + 1) - dphim;
        xixm(i, j, n) = xixm(i, j, n) +
                        (phip(i - 1, j, kum[i] + 1) + p(i - 1, j, kum[i] + 1) * alplm - cm * (alpum - alpup)) *
                            dpu(i, j, kn);
        pgfx(i, j, kn) = -(phi_p - phi_m);
        pgfxm(i, j, n) = pgfxm(i, j, n) + pgfx(i, j, kn) * dpu(i, j, kn);
      }
      for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) {
        prs = pv(i, j, k + 1) - .5 * dpv(i, j, kn);
        while (p(i, j, kvp[i]) > prs) kvp[i] = kvp[i] - 1;
        while (p(i, j - 1, kvm[i]) > prs) kvm[i] = kvm[i] - 1;
        delphi(prs, p(i, j, kvp[i] + 1), temp(i, j, kvp[i] + nn), saln(i, j, kvp[i] + nn), dphip, alpup, alplp);
        delphi(prs, p(i, j - 1, kvm[i] + 1), temp(i, j - 1, kvm[i] + nn), saln(i, j - 1, kvm[i] + nn), dphim, alpum,
               alplm);
        cp = .25 * (p(i, j, k + 1) + p(i, j, k));
        cm = .25 * (p(i, j - 1, k + 1) + p(i, j - 1, k));
        q = prs / (cp + cm);
        cp = q * cp;
        cm = q * cm;
        phi_p = phi(i, j, kvp[i] + 1) - dphip;
        xiyp(i, j, n) = xiyp(i, j, n) +
                        (phip(i, j, kvp[i] + 1) + p(i, j, kvp[i] + 1) * alplp - cp * (alpup - alpum)) * dpv(i, j, kn);
        phi_m = phi(i, j - 1, kvm[i] + 1) - dphim;
        xiym(i, j, n) = xiym(i, j, n) +
                        (phip(i, j - 1, kvm[i] + 1) + p(i, j - 1, kvm[i] + 1) * alplm - cm * (alpum - alpup)) *
                            dpv(i, j, kn);
        pgfy(i, j, kn) = -(phi_p - phi_m);
        pgfym(i, j, n) = pgfym(i, j, n) + pgfy(i, j, kn) * dpv(i, j, kn);
      }
    }
  }
}

// :438-615
void pgforc(int m, int n, int mm, int nn, int k1m, int k1n) {
  Oracle& o = O(); const Dims& d = o.d;
  const int ii = d.ii, jj = d.jj, kk = d.kk;
  A3 p = o.a3("p"), dp = o.a3("dp"), dpu = o.a3("dpu"), dpv = o.a3("dpv"), pu = o.a3("pu"), pv = o.a3("pv");
  A3 phi = o.a3("phi"), pgfx = o.a3("pgfx"), pgfy = o.a3("pgfy"), pgfx_o = o.a3("pgfx_o"), pgfy_o = o.a3("pgfy_o");
  A3 pgfxm = o.a3("pgfxm"), pgfym = o.a3("pgfym"), xixp = o.a3("xixp"), xixm = o.a3("xixm"),
     xiyp = o.a3("xiyp"), xiym = o.a3("xiym");
  A2 pgfxm_o = o.a2("pgfxm_o"), pgfym_o = o.a2("pgfym_o"), xixp_o = o.a2("xixp_o"), xixm_o = o.a2("xixm_o"),
     xiyp_o = o.a2("xiyp_o"), xiym_o = o.a2("xiym_o");
  A2 pb_p = o.a2("pb_p"), pbu_p = o.a2("pbu_p"), pbv_p = o.a2("pbv_p"), sealv = o.a2("sealv");
  I2 ip = o.i2("ip"), iu = o.i2("iu"), iv = o.i2("iv");

  for (int j = -2; j <= jj + 2; ++j)
    for (int k = 1; k <= kk; ++k) {
      const int kn = k + nn;
      for (int i = -2; i <= ii + 2; ++i) if (ip(i, j) == 1) p(i, j, k + 1) = p(i, j, k) + dp(i, j, kn);
    }
  for (int j = -1; j <= jj + 2; ++j)
    for (int k = 1; k <= kk; ++k) {
      const int kn = k + nn;
      for (int i = -1; i <= ii + 2; ++i) if (iu(i, j) == 1) {
        double q = std::min(p(i, j, kk + 1), p(i - 1, j, kk + 1));
        dpu(i, j, kn) = .5 * ((std::min(q, p(i - 1, j, k + 1)) - std::min(q, p(i - 1, j, k))) +
                              (std::min(q, p(i, j, k + 1)) - std::min(q, p(i, j, k))));
        pu(i, j, k + 1) = pu(i, j, k) + dpu(i, j, kn);
      }
      for (int i = -1; i <= ii + 2; ++i) if (iv(i, j) == 1) {
        double q = std::min(p(i, j, kk + 1), p(i, j - 1, kk + 1));
        dpv(i, j, kn) = .5 * ((std::min(q, p(i, j - 1, k + 1)) - std::min(q, p(i, j - 1, k))) +
                              (std::min(q, p(i, j, k + 1)) - std::min(q, p(i, j, k))));
        pv(i, j, k + 1) = pv(i, j, k) + dpv(i, j, kn);
      }
    }
  for (int j = -1; j <= jj + 2; ++j) {
    for (int i = 0; i <= ii + 1; ++i) if (iu(i, j) == 1) {
      xixp_o(i, j) = xixp(i, j, n); xixm_o(i, j) = xixm(i, j, n); pgfxm_o(i, j) = pgfxm(i, j, n);
    }
    for (int i = 0; i <= ii + 1; ++i) if (iv(i, j) == 1) {
      xiyp_o(i, j) = xiyp(i, j, n); xiym_o(i, j) = xiym(i, j, n); pgfym_o(i, j) = pgfym(i, j, n);
    }
  }
  for (int j = 1; j <= jj; ++j)
    for (int k = kk; k >= 1; --k) {
      const int kn = k + nn;
      for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) pgfx_o(i, j, k) = pgfx(i, j, kn);
      for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) pgfy_o(i, j, k) = pgfy(i, j, kn);
    }
  const std::string pgfmth = o.option("pgfmth", "dynamic enthalpy");
  if (pgfmth == "geopotential") pgforc_geopotential(m, n, mm, nn, k1m, k1n);
  else if (pgfmth == "dynamic enthalpy") pgforc_dynamic_enthalpy(m, n, mm, nn, k1m, k1n);
  else throw std::runtime_error(" pgfmth = " + pgfmth + " is unsupported!");

  xctilr(pb_p, 1, 1, halo_ps);
  for (int j = 1; j <= jj; ++j) {
    for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) {
      double q = 1. / pbu_p(i, j);
      pgfxm(i, j, n) = pgfxm(i, j, n) * q; xixp(i, j, n) = xixp(i, j, n) * q; xixm(i, j, n) = xixm(i, j, n) * q;
    }
    for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) {
      double q = 1. / pbv_p(i, j);
      pgfym(i, j, n) = pgfym(i, j, n) * q; xiyp(i, j, n) = xiyp(i, j, n) * q; xiym(i, j, n) = xiym(i, j, n) * q;
    }
    for (int k = 1; k <= kk; ++k) {
      const int kn = k + nn;
      for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) pgfx(i, j, kn) = pgfx(i, j, kn) - pgfxm(i, j, n);
      for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) pgfy(i, j, kn) = pgfy(i, j, kn) - pgfym(i, j, n);
    }
    for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) {
      pgfxm(i, j, n) = pgfxm(i, j, n) + xixp(i, j, n) - xixm(i, j, n);
      xixp(i, j, n) = xixp(i, j, n) / pb_p(i, j);
      xixm(i, j, n) = xixm(i, j, n) / pb_p(i - 1, j);
    }
    for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) {
      pgfym(i, j, n) = pgfym(i, j, n) + xiyp(i, j, n) - xiym(i, j, n);
      xiyp(i, j, n) = xiyp(i, j, n) / pb_p(i, j);
      xiym(i, j, n) = xiym(i, j, n) / pb_p(i, j - 1);
    }
    for (int i = 1; i <= ii; ++i) if (ip(i, j) == 1) sealv(i, j) = phi(i, j, 1) / grav;
  }
}

}  // namespace orc
