// ORACLE — TEST INFRASTRUCTURE ONLY (see core.hpp header).
// Restatement of phy/mod_barotp.F90:148-1003 (split-explicit barotropic subcycle).
#include "core.hpp"

namespace orc {

void barotp(int m, int n, int mm, int nn, int k1m, int k1n) {
  (void)mm; (void)k1m; (void)k1n;
  Oracle& o = O(); const Dims& d = o.d;
  const int ii = d.ii, jj = d.jj, kk = d.kk;
  const double wbaro = .125;  // phy/mod_tmsmt.F90:51
  const int lstep = (int)o.scalar("lstep");
  const double dlt = o.scalar("dlt");
  const double cwbdts = o.scalar("cwbdts", 0.0), cwbdls = o.scalar("cwbdls", 25.0);
  const std::string mommth = o.option("mommth", "enscon");
  const bool enscon = mommth == "enscon";
  if (!enscon && mommth != "enecon" && mommth != "enedis")
    throw std::runtime_error(" mommth = " + mommth + " is unsupported!");

  // routine-local `save` arrays (:155-167)
  A3 pb_t = o.scratch("barotp_pb_t", 2), ubflx_t = o.scratch("barotp_ubflx_t", 2), vbflx_t = o.scratch("barotp_vbflx_t", 2);
  A2 umaxb = o.scratch("barotp_umaxb", 1).level(1), uminb = o.scratch("barotp_uminb", 1).level(1),
     vmaxb = o.scratch("barotp_vmaxb", 1).level(1), vminb = o.scratch("barotp_vminb", 1).level(1),
     uglue = o.scratch("barotp_uglue", 1).level(1), vglue = o.scratch("barotp_vglue", 1).level(1),
     ubflxs_t = o.scratch("barotp_ubflxs_t", 1).level(1), vbflxs_t = o.scratch("barotp_vbflxs_t", 1).level(1),
     ubcors_t = o.scratch("barotp_ubcors_t", 1).level(1), vbcors_t = o.scratch("barotp_vbcors_t", 1).level(1);
  A3 u = o.a3("u"), v = o.a3("v"), ubflxs = o.a3("ubflxs"), vbflxs = o.a3("vbflxs"), ub = o.a3("ub"), vb = o.a3("vb"),
     pb = o.a3("pb"), pbu = o.a3("pbu"), pbv = o.a3("pbv"), ubflxs_p = o.a3("ubflxs_p"), vbflxs_p = o.a3("vbflxs_p");
  A2 pb_p = o.a2("pb_p"), pbu_p = o.a2("pbu_p"), pbv_p = o.a2("pbv_p"), ubcors_p = o.a2("ubcors_p"),
     vbcors_p = o.a2("vbcors_p");
  A3 pgfxm = o.a3("pgfxm"), pgfym = o.a3("pgfym"), xixp = o.a3("xixp"), xixm = o.a3("xixm"), xiyp = o.a3("xiyp"),
     xiym = o.a3("xiym");
  A2 pgfxm_o = o.a2("pgfxm_o"), pgfym_o = o.a2("pgfym_o"), xixp_o = o.a2("xixp_o"), xixm_o = o.a2("xixm_o"),
     xiyp_o = o.a2("xiyp_o"), xiym_o = o.a2("xiym_o");
  A2 utotn = o.a2("utotn"), vtotn = o.a2("vtotn"), umax = o.a2("umax"), vmax = o.a2("vmax");
  A3 ubflx = o.a3("ubflx"), vbflx = o.a3("vbflx"), pb_mn = o.a3("pb_mn"), ubflx_mn = o.a3("ubflx_mn"),
     vbflx_mn = o.a3("vbflx_mn"), pvtrop = o.a3("pvtrop");
  A2 pvtrop_o = o.a2("pvtrop_o");
  A2 scuy = o.a2("scuy"), scvx = o.a2("scvx"), scp2i = o.a2("scp2i"), scuxi = o.a2("scuxi"), scuyi = o.a2("scuyi"),
     scvxi = o.a2("scvxi"), scvyi = o.a2("scvyi"), corioq = o.a2("corioq");
  I2 ip = o.i2("ip"), iu = o.i2("iu"), iv = o.i2("iv"), iq = o.i2("iq");

  // :177-224
  for (int j = 1; j <= jj; ++j) {
    for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) {
      umaxb(i, j) = 0.; uminb(i, j) = 0.;
      uglue(i, j) = cwbdts * std::exp(1. - pbu(i, j, m) / (cwbdls * onem));
    }
    for (int k = 1; k <= kk; ++k) {
      const int kn = k + nn;
      for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) {
        umaxb(i, j) = std::max(umaxb(i, j), u(i, j, kn));
        uminb(i, j) = std::min(uminb(i, j), u(i, j, kn));
      }
    }
    for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) {
      umaxb(i, j) = (umax(i, j) - umaxb(i, j)) * pbu(i, j, m) * scuy(i, j);
      uminb(i, j) = (umax(i, j) + uminb(i, j)) * pbu(i, j, m) * scuy(i, j);
    }
    for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) {
      vmaxb(i, j) = 0.; vminb(i, j) = 0.;
      vglue(i, j) = cwbdts * std::exp(1. - pbv(i, j, m) / (cwbdls * onem));
    }
    for (int k = 1; k <= kk; ++k) {
      const int kn = k + nn;
      for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) {
        vmaxb(i, j) = std::max(vmaxb(i, j), v(i, j, kn));
        vminb(i, j) = std::min(vminb(i, j), v(i, j, kn));
      }
    }
    for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) {
      vmaxb(i, j) = (vmax(i, j) - vmaxb(i, j)) * pbv(i, j, m) * scvx(i, j);
      vminb(i, j) = (vmax(i, j) + vminb(i, j)) * pbv(i, j, m) * scvx(i, j);
    }
  }
  // :230-269 potential vorticity of barotropic flow
  for (int j = -2; j <= jj + 3; ++j)
    for (int i = 0; i <= ii + 1; ++i) pvtrop_o(i, j) = pvtrop(i, j, n);
  for (int j = 0; j <= jj; ++j)
    for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) {
      double q = 2. / (pb_p(i, j) + pb_p(i - 1, j));
      pvtrop(i, j, n) = corioq(i, j) * q;
      pvtrop(i, j + 1, n) = corioq(i, j + 1) * q;
    }
  for (int j = 1; j <= jj; ++j)
    for (int i = 0; i <= ii; ++i) if (iv(i, j) == 1) {
      double q = 2. / (pb_p(i, j) + pb_p(i, j - 1));
      pvtrop(i, j, n) = corioq(i, j) * q;
      pvtrop(i + 1, j, n) = corioq(i + 1, j) * q;
    }
  for (int j = 1; j <= jj; ++j)
    for (int i = 1; i <= ii; ++i) if (iq(i, j) == 1)
      pvtrop(i, j, n) = corioq(i, j) * 4. / (pb_p(i, j) + pb_p(i - 1, j) + pb_p(i, j - 1) + pb_p(i - 1, j - 1));
  // :271-285
  xctilr(uglue, 1, 2, halo_us); xctilr(utotn, 1, 2, halo_uv); xctilr(umaxb, 1, 2, halo_us); xctilr(uminb, 1, 2, halo_us);
  xctilr(vglue, 1, 2, halo_vs); xctilr(vtotn, 1, 2, halo_vv); xctilr(vmaxb, 1, 2, halo_vs); xctilr(vminb, 1, 2, halo_vs);
  xctilr(pvtrop.level(n), 1, 3, halo_qs);
  xctilr(pgfxm.level(n), 1, 2, halo_uv); xctilr(xixp.level(n), 1, 2, halo_us); xctilr(xixm.level(n), 1, 2, halo_us);
  xctilr(pgfym.level(n), 1, 2, halo_vv); xctilr(xiyp.level(n), 1, 2, halo_vs); xctilr(xiym.level(n), 1, 2, halo_vs);
  // :290-319 arctic switches
  if (d.nreg == 2) {
    for (int j = jj; j <= jj + 2; ++j)
      for (int i = 0; i <= ii + 1; ++i) {
        std::swap(umaxb(i, j), uminb(i, j));
        std::swap(xixp(i, j, n), xixm(i, j, n));
      }
    for (int i = std::max(0, d.itdm / 2 - d.i0 + 1); i <= ii + 1; ++i) {
      std::swap(vmaxb(i, jj), vminb(i, jj));
      std::swap(xiyp(i, jj, n), xiym(i, jj, n));
    }
    for (int j = jj + 1; j <= jj + 2; ++j)
      for (int i = 0; i <= ii + 1; ++i) {
        std::swap(vmaxb(i, j), vminb(i, j));
        std::swap(xiyp(i, j, n), xiym(i, j, n));
      }
  }

  int lll0 = 1, ml = 1, nl = 2;
  double woa = 0, wob = 0, wna = 0, wnb = 0;
  for (int nb = 1; nb <= 5; ++nb) {
    if (nb == 1) {
      lll0 = 1; ml = 1; nl = 2;
      woa = -1. / lstep;
      wob = .5 + (lll0 - .5) / lstep;
      wna = 0.; wnb = 0.;
      for (int j = 1; j <= jj; ++j)
        for (int i = 1; i <= ii; ++i) {
          pb_t(i, j, ml) = pb_mn(i, j, ml); pb_t(i, j, nl) = pb_mn(i, j, nl);
          ubflx_t(i, j, ml) = ubflx_mn(i, j, ml); ubflx_t(i, j, nl) = ubflx_mn(i, j, nl);
          vbflx_t(i, j, ml) = vbflx_mn(i, j, ml); vbflx_t(i, j, nl) = vbflx_mn(i, j, nl);
        }
    } else if (nb == 2) {
      woa = 0.; wob = 0.;
      wna = 1. / lstep;
      wnb = -(lll0 - .5) / lstep;
    } else if (nb == 4) {
      wna = 0.; wnb = 1.;
    }
    for (int j = -1; j <= jj + 2; ++j)
      for (int i = 0; i <= ii + 1; ++i) if (iu(i, j) == 1) { ubflxs_t(i, j) = 0.; ubcors_t(i, j) = 0.; }
    for (int j = 0; j <= jj + 2; ++j)
      for (int i = 0; i <= ii; ++i) if (iv(i, j) == 1) { vbflxs_t(i, j) = 0.; vbcors_t(i, j) = 0.; }

    auto continuity = [&](int j0, int j1, int i0, int i1) {
      for (int j = j0; j <= j1; ++j)
        for (int i = i0; i <= i1; ++i) if (ip(i, j) == 1)
          pb_t(i, j, nl) = (1. - wbaro) * pb_t(i, j, ml) + wbaro * pb_t(i, j, nl) -
                           (1. + wbaro) * dlt * (ubflx_t(i + 1, j, ml) - ubflx_t(i, j, ml) + vbflx_t(i, j + 1, ml) -
                                                 vbflx_t(i, j, ml)) * scp2i(i, j);
    };
    // lv: time level of vbflx_t used in the Coriolis term (ml on odd, nl on even substeps)
    auto ueq = [&](int j0, int j1, int i0, int i1, int lv, double wo, double wm, double wn) {
      for (int j = j0; j <= j1; ++j)
        for (int i = i0; i <= i1; ++i) if (iu(i, j) == 1) {
          ubflxs_t(i, j) = ubflxs_t(i, j) - wbaro * ubflx_t(i, j, nl) + (1. + wbaro) * ubflx_t(i, j, ml);
          double q;
          if (enscon)
            q = (vbflx_t(i, j, lv) * scvxi(i, j) + vbflx_t(i, j + 1, lv) * scvxi(i, j + 1) +
                 vbflx_t(i - 1, j, lv) * scvxi(i - 1, j) + vbflx_t(i - 1, j + 1, lv) * scvxi(i - 1, j + 1)) *
                (wo * (pvtrop_o(i, j) + pvtrop_o(i, j + 1)) + wm * (pvtrop(i, j, m) + pvtrop(i, j + 1, m)) +
                 wn * (pvtrop(i, j, n) + pvtrop(i, j + 1, n))) * .125;
          else
            q = .25 * ((vbflx_t(i, j, lv) * scvxi(i, j) + vbflx_t(i - 1, j, lv) * scvxi(i - 1, j)) *
                           (wo * pvtrop_o(i, j) + wm * pvtrop(i, j, m) + wn * pvtrop(i, j, n)) +
                       (vbflx_t(i, j + 1, lv) * scvxi(i, j + 1) + vbflx_t(i - 1, j + 1, lv) * scvxi(i - 1, j + 1)) *
                           (wo * pvtrop_o(i, j + 1) + wm * pvtrop(i, j + 1, m) + wn * pvtrop(i, j + 1, n)));
          ubcors_t(i, j) = ubcors_t(i, j) + q;
          double utndcy = q + (wo * (pgfxm_o(i, j) - (xixp_o(i, j) * pb_t(i, j, nl) - xixm_o(i, j) * pb_t(i - 1, j, nl))) +
                               wm * (pgfxm(i, j, m) - (xixp(i, j, m) * pb_t(i, j, nl) - xixm(i, j, m) * pb_t(i - 1, j, nl))) +
                               wn * (pgfxm(i, j, n) - (xixp(i, j, n) * pb_t(i, j, nl) - xixm(i, j, n) * pb_t(i - 1, j, nl)))) *
                                  scuxi(i, j);
          ubflx_t(i, j, nl) = (1. - wbaro) * ubflx_t(i, j, ml) + wbaro * ubflx_t(i, j, nl) +
                              (1. + wbaro) * dlt * ((utndcy + utotn(i, j)) * scuy(i, j) *
                                                        std::min(pb_t(i - 1, j, nl), pb_t(i, j, nl)) -
                                                    uglue(i, j) * ubflx_t(i, j, ml));
          ubflx_t(i, j, nl) = std::max(-uminb(i, j), std::min(umaxb(i, j), ubflx_t(i, j, nl)));
        }
    };
    auto veq = [&](int j0, int j1, int i0, int i1, int lu, double wo, double wm, double wn) {
      for (int j = j0; j <= j1; ++j)
        for (int i = i0; i <= i1; ++i) if (iv(i, j) == 1) {
          vbflxs_t(i, j) = vbflxs_t(i, j) - wbaro * vbflx_t(i, j, nl) + (1. + wbaro) * vbflx_t(i, j, ml);
          double q;
          if (enscon)
            q = -(ubflx_t(i, j, lu) * scuyi(i, j) + ubflx_t(i + 1, j, lu) * scuyi(i + 1, j) +
                  ubflx_t(i, j - 1, lu) * scuyi(i, j - 1) + ubflx_t(i + 1, j - 1, lu) * scuyi(i + 1, j - 1)) *
                (wo * (pvtrop_o(i, j) + pvtrop_o(i + 1, j)) + wm * (pvtrop(i, j, m) + pvtrop(i + 1, j, m)) +
                 wn * (pvtrop(i, j, n) + pvtrop(i + 1, j, n))) * .125;
          else
            q = -.25 * ((ubflx_t(i, j, lu) * scuyi(i, j) + ubflx_t(i, j - 1, lu) * scuyi(i, j - 1)) *
                            (wo * pvtrop_o(i, j) + wm * pvtrop(i, j, m) + wn * pvtrop(i, j, n)) +
                        (ubflx_t(i + 1, j, lu) * scuyi(i + 1, j) + ubflx_t(i + 1, j - 1, lu) * scuyi(i + 1, j - 1)) *
                            (wo * pvtrop_o(i + 1, j) + wm * pvtrop(i + 1, j, m) + wn * pvtrop(i + 1, j, n)));
          vbcors_t(i, j) = vbcors_t(i, j) + q;
          double vtndcy = q + (wo * (pgfym_o(i, j) - (xiyp_o(i, j) * pb_t(i, j, nl) - xiym_o(i, j) * pb_t(i, j - 1, nl))) +
                               wm * (pgfym(i, j, m) - (xiyp(i, j, m) * pb_t(i, j, nl) - xiym(i, j, m) * pb_t(i, j - 1, nl))) +
                               wn * (pgfym(i, j, n) - (xiyp(i, j, n) * pb_t(i, j, nl) - xiym(i, j, n) * pb_t(i, j - 1, nl)))) *
                                  scvyi(i, j);
          vbflx_t(i, j, nl) = (1. - wbaro) * vbflx_t(i, j, ml) + wbaro * vbflx_t(i, j, nl) +
                              (1. + wbaro) * dlt * ((vtndcy + vtotn(i, j)) * scvx(i, j) *
                                                        std::min(pb_t(i, j - 1, nl), pb_t(i, j, nl)) -
                                                    vglue(i, j) * vbflx_t(i, j, ml));
          vbflx_t(i, j, nl) = std::max(-vminb(i, j), std::min(vmaxb(i, j), vbflx_t(i, j, nl)));
        }
    };

    for (int lll = lll0; lll <= lll0 + lstep / 2 - 1; ++lll) {
      const double wo = woa * lll + wob, wn = wna * lll + wnb, wm = 1. - wo - wn;
      if (lll % 2 == 1) {
        xctilr(pb_t, 1, 2, 2, 2, halo_ps);
        xctilr(ubflx_t, 1, 2, 2, 2, halo_uv);
        xctilr(vbflx_t, 1, 2, 2, 3, halo_vv);
        continuity(-1, jj + 2, -1, ii + 1);
        ueq(-1, jj + 2, 0, ii + 1, ml, wo, wm, wn);
        veq(0, jj + 2, 0, ii, nl, wo, wm, wn);
      } else {
        continuity(0, jj + 1, 0, ii);
        veq(1, jj + 1, 0, ii, ml, wo, wm, wn);
        ueq(1, jj, 1, ii, nl, wo, wm, wn);
      }
      std::swap(ml, nl);
    }
    lll0 = lll0 + lstep / 2;

    // :847-977 harvest
    for (int j = 1; j <= jj; ++j) {
      if (nb == 1) {
        for (int i = 1; i <= ii; ++i) if (ip(i, j) == 1) pb(i, j, m) = pb_t(i, j, ml);
        for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) {
          pbu(i, j, m) = std::min(pb_t(i, j, ml), pb_t(i - 1, j, ml));
          ubflx(i, j, m) = ubflx_t(i, j, ml);
          ub(i, j, m) = ubflx(i, j, m) / (pbu(i, j, m) * scuy(i, j));
          ubflxs(i, j, n) = ubflxs(i, j, n) + ubflxs_t(i, j);
          ubflxs(i, j, m) = ubflxs(i, j, 3) + ubflxs_t(i, j);
        }
        for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) {
          pbv(i, j, m) = std::min(pb_t(i, j, ml), pb_t(i, j - 1, ml));
          vbflx(i, j, m) = vbflx_t(i, j, ml);
          vb(i, j, m) = vbflx(i, j, m) / (pbv(i, j, m) * scvx(i, j));
          vbflxs(i, j, n) = vbflxs(i, j, n) + vbflxs_t(i, j);
          vbflxs(i, j, m) = vbflxs(i, j, 3) + vbflxs_t(i, j);
        }
      } else if (nb == 2) {
        for (int i = 1; i <= ii; ++i) if (ip(i, j) == 1) { pb_mn(i, j, ml) = pb_t(i, j, ml); pb_mn(i, j, nl) = pb_t(i, j, nl); }
        for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) {
          ubflx_mn(i, j, ml) = ubflx_t(i, j, ml); ubflx_mn(i, j, nl) = ubflx_t(i, j, nl);
          ubflxs(i, j, m) = ubflxs(i, j, m) + ubflxs_t(i, j);
          ubflxs(i, j, 3) = ubflxs_t(i, j);
          ubflxs_p(i, j, n) = ubflxs_t(i, j);
          ubcors_p(i, j) = ubcors_t(i, j);
        }
        for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) {
          vbflx_mn(i, j, ml) = vbflx_t(i, j, ml); vbflx_mn(i, j, nl) = vbflx_t(i, j, nl);
          vbflxs(i, j, m) = vbflxs(i, j, m) + vbflxs_t(i, j);
          vbflxs(i, j, 3) = vbflxs_t(i, j);
          vbflxs_p(i, j, n) = vbflxs_t(i, j);
          vbcors_p(i, j) = vbcors_t(i, j);
        }
      } else if (nb == 3) {
        for (int i = 1; i <= ii; ++i) if (ip(i, j) == 1) pb(i, j, n) = pb_t(i, j, ml);
        for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) {
          pbu(i, j, n) = std::min(pb_t(i, j, ml), pb_t(i - 1, j, ml));
          ubflx(i, j, n) = ubflx_t(i, j, ml);
          ub(i, j, n) = ubflx(i, j, n) / (pbu(i, j, n) * scuy(i, j));
          ubflxs_p(i, j, m) = ubflxs(i, j, m) + ubflxs_t(i, j);
          ubflxs_p(i, j, n) = ubflxs_p(i, j, n) + ubflxs_t(i, j);
          ubcors_p(i, j) = ubcors_p(i, j) + ubcors_t(i, j);
        }
        for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) {
          pbv(i, j, n) = std::min(pb_t(i, j, ml), pb_t(i, j - 1, ml));
          vbflx(i, j, n) = vbflx_t(i, j, ml);
          vb(i, j, n) = vbflx(i, j, n) / (pbv(i, j, n) * scvx(i, j));
          vbflxs_p(i, j, m) = vbflxs(i, j, m) + vbflxs_t(i, j);
          vbflxs_p(i, j, n) = vbflxs_p(i, j, n) + vbflxs_t(i, j);
          vbcors_p(i, j) = vbcors_p(i, j) + vbcors_t(i, j);
        }
      } else if (nb == 4) {
        for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) {
          ubflxs_p(i, j, n) = ubflxs_p(i, j, n) + ubflxs_t(i, j);
          ubcors_p(i, j) = ubcors_p(i, j) + ubcors_t(i, j);
        }
        for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) {
          vbflxs_p(i, j, n) = vbflxs_p(i, j, n) + vbflxs_t(i, j);
          vbcors_p(i, j) = vbcors_p(i, j) + vbcors_t(i, j);
        }
      } else {
        for (int i = 1; i <= ii; ++i) if (ip(i, j) == 1) pb_p(i, j) = pb_t(i, j, ml);
        for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) {
          pbu_p(i, j) = std::min(pb_t(i, j, ml), pb_t(i - 1, j, ml));
          ubflxs_p(i, j, n) = ubflxs_p(i, j, n) + ubflxs_t(i, j);
          ubcors_p(i, j) = ubcors_p(i, j) + ubcors_t(i, j);
        }
        for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) {
          pbv_p(i, j) = std::min(pb_t(i, j, ml), pb_t(i, j - 1, ml));
          vbflxs_p(i, j, n) = vbflxs_p(i, j, n) + vbflxs_t(i, j);
          vbcors_p(i, j) = vbcors_p(i, j) + vbcors_t(i, j);
        }
      }
    }
  }
}

}  // namespace orc
