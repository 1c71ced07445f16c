// ORACLE — TEST INFRASTRUCTURE ONLY (see core.hpp header).
// Restatement of phy/mod_remap.F90 (incremental remapping, advmth='remap'): triint :53-102,
// penint :104-199, remap :205-1522, and of its driver in phy/mod_advect.F90:96-153.
// Build options restated: use_TRC = (ntr > 0), use_ATRC = use_TKE = .false. (the defaults,
// phy/mod_ifdefs.F90), so the age-tracer and TKE branches are not present.
//
// The reference spells the flux integral of each donor cell out twelve times (u/v face x flow
// direction x two corner triangles and one pentagon); here the common tail of those blocks is the
// local `donor()` below — same expressions, same left-to-right accumulation order.
#include "core.hpp"

namespace orc {

namespace {

constexpr double dpeps = 1.e-12;  // :39

struct Mom { double a, ax, ay, axx, ayy, axy; };

// :53-102 (third-order moments are only needed with use_ATRC)
Mom triint(double ac, double x1, double y1, double x2, double y2, double x3, double y3) {
  const double r1_3 = 1. / 3., r1_6 = 1. / 6., r1_12 = 1. / 12.;
  Mom m;
  double xx = x1 * x2 + x2 * x3 + x1 * x3;
  double yy = y1 * y2 + y2 * y3 + y1 * y3;
  double xy1 = x1 * y1, xy2 = x2 * y2, xy3 = x3 * y3;
  double xy = xy1 + xy2 + xy3;
  m.a = .5 * ((x2 - x1) * (y3 - y1) - (y2 - y1) * (x3 - x1)) * ac;
  m.ax = r1_3 * (x1 + x2 + x3);
  m.ay = r1_3 * (y1 + y2 + y3);
  m.axx = r1_6 * (9. * m.ax * m.ax - xx);
  m.ayy = r1_6 * (9. * m.ay * m.ay - yy);
  m.axy = r1_12 * (9. * m.ax * m.ay + xy);
  m.ax = m.ax * m.a;
  m.ay = m.ay * m.a;
  m.axx = m.axx * m.a;
  m.ayy = m.ayy * m.a;
  m.axy = m.axy * m.a;
  return m;
}

// :104-199
Mom penint(double ac, double x1, double y1, double x2, double y2, double x3, double y3, double x4,
           double y4, double x5, double y5) {
  const double r1_3 = 1. / 3., r1_6 = 1. / 6., r1_12 = 1. / 12.;
  double xx123 = x1 * x2 + x2 * x3 + x1 * x3;
  double yy123 = y1 * y2 + y2 * y3 + y1 * y3;
  double xx135 = x1 * x3 + x3 * x5 + x1 * x5;
  double yy135 = y1 * y3 + y3 * y5 + y1 * y5;
  double xx345 = x3 * x4 + x4 * x5 + x3 * x5;
  double yy345 = y3 * y4 + y4 * y5 + y3 * y5;
  double xy1 = x1 * y1, xy2 = x2 * y2, xy3 = x3 * y3, xy4 = x4 * y4, xy5 = x5 * y5;
  double xy123 = xy1 + xy2 + xy3;
  double xy135 = xy1 + xy3 + xy5;
  double xy345 = xy3 + xy4 + xy5;
  double a123 = .5 * ((x2 - x1) * (y3 - y1) - (y2 - y1) * (x3 - x1)) * ac;
  double a135 = .5 * ((x3 - x1) * (y5 - y1) - (y3 - y1) * (x5 - x1)) * ac;
  double a345 = .5 * ((x4 - x3) * (y5 - y3) - (y4 - y3) * (x5 - x3)) * ac;
  double ax123 = r1_3 * (x1 + x2 + x3), ay123 = r1_3 * (y1 + y2 + y3);
  double ax135 = r1_3 * (x1 + x3 + x5), ay135 = r1_3 * (y1 + y3 + y5);
  double ax345 = r1_3 * (x3 + x4 + x5), ay345 = r1_3 * (y3 + y4 + y5);
  double axx123 = r1_6 * (9. * ax123 * ax123 - xx123);
  double ayy123 = r1_6 * (9. * ay123 * ay123 - yy123);
  double axy123 = r1_12 * (9. * ax123 * ay123 + xy123);
  double axx135 = r1_6 * (9. * ax135 * ax135 - xx135);
  double ayy135 = r1_6 * (9. * ay135 * ay135 - yy135);
  double axy135 = r1_12 * (9. * ax135 * ay135 + xy135);
  double axx345 = r1_6 * (9. * ax345 * ax345 - xx345);
  double ayy345 = r1_6 * (9. * ay345 * ay345 - yy345);
  double axy345 = r1_12 * (9. * ax345 * ay345 + xy345);
  Mom m;
  m.a = a123 + a135 + a345;
  m.ax = ax123 * a123 + ax135 * a135 + ax345 * a345;
  m.ay = ay123 * a123 + ay135 * a135 + ay345 * a345;
  m.axx = axx123 * a123 + axx135 * a135 + axx345 * a345;
  m.ayy = ayy123 * a123 + ayy135 * a135 + ayy345 * a345;
  m.axy = axy123 * a123 + axy135 * a135 + axy345 * a345;
  return m;
}

inline double max8(double a, double b, double c, double d, double e, double f, double g, double h) {
  return std::max(std::max(std::max(a, b), std::max(c, d)), std::max(std::max(e, f), std::max(g, h)));
}
inline double min8(double a, double b, double c, double d, double e, double f, double g, double h) {
  return std::min(std::min(std::min(a, b), std::min(c, d)), std::min(std::min(e, f), std::min(g, h)));
}

// wet-neighbour indices (:368-381, phy/mod_advect.F90:103-114)
struct Nbr { int iw, ie, js, jn, isw, jsw, ise, jse, inw, jnw, ine, jne; };
inline Nbr neighbours(int i, int j, I2 ip, I2 iu, I2 iv) {
  Nbr q;
  q.iw = i - iu(i, j);
  q.ie = i + iu(i + 1, j);
  q.js = j - iv(i, j);
  q.jn = j + iv(i, j + 1);
  q.isw = i * (1 - ip(q.iw, q.js)) + q.iw * ip(q.iw, q.js);
  q.jsw = j * (1 - ip(q.iw, q.js)) + q.js * ip(q.iw, q.js);
  q.ise = i * (1 - ip(q.ie, q.js)) + q.ie * ip(q.ie, q.js);
  q.jse = j * (1 - ip(q.ie, q.js)) + q.js * ip(q.ie, q.js);
  q.inw = i * (1 - ip(q.iw, q.jn)) + q.iw * ip(q.iw, q.jn);
  q.jnw = j * (1 - ip(q.iw, q.jn)) + q.jn * ip(q.iw, q.jn);
  q.ine = i * (1 - ip(q.ie, q.jn)) + q.ie * ip(q.ie, q.jn);
  q.jne = j * (1 - ip(q.ie, q.jn)) + q.jn * ip(q.ie, q.jn);
  return q;
}

// (nt,i,j) scratch, nt fastest like the reference's trx(nt,i,j)
struct T3 {
  double* p; int ntr, ldi, nb;
  inline double& operator()(int nt, int i, int j) const {
    return p[((size_t)(j + nb - 1) * ldi + (i + nb - 1)) * ntr + (nt - 1)];
  }
};

// :205-1522.  `trc` is the tracer array viewed at level kn (trc(i,j,k,nt) with k = kn of the call
// at phy/mod_advect.F90:137-147); tracer nt lives trc_stride levels further on.
void remap(A2 scp2i, A2 scp2, A2 pbmin, A2 pbu, A2 pbv, A2 plo, A2 cau, A2 cav, int mrg, A2 dp, A2 temp,
           A2 saln, A2 uflx, A2 vflx, A2 utflx, A2 vtflx, A2 usflx, A2 vsflx, double* trc_kn,
           size_t trc_stride) {
  Oracle& o = O(); const Dims& d = o.d;
  const int ii = d.ii, jj = d.jj, nb = d.nbdy, ntr = d.ntr;
  I2 ip = o.i2("ip"), iu = o.i2("iu"), iv = o.i2("iv");
  const size_t L = d.lev;
  std::vector<double> w(L * 21, 0.0), wt((size_t)std::max(1, ntr) * L * 5, 0.0);
  auto W = [&](int n) { return A2{w.data() + L * n, d.ldi, nb}; };
  A2 pup = W(0), dx = W(1), dy = W(2), xd = W(3), yd = W(4), tx = W(5), ty = W(6), td = W(7), sx = W(8),
     sy = W(9), sd = W(10), cu = W(11), cv = W(12), cuc = W(13), cvc = W(14), fdu = W(15), fdv = W(16),
     ftu = W(17), ftv = W(18), fsu = W(19), fsv = W(20);
  auto WT = [&](int n) { return T3{wt.data() + (size_t)std::max(1, ntr) * L * n, std::max(1, ntr), d.ldi, nb}; };
  T3 trx = WT(0), try_ = WT(1), trd = WT(2), ftru = WT(3), ftrv = WT(4);
  auto trc = [&](int i, int j, int nt) -> double& {
    return trc_kn[(size_t)(nt - 1) * trc_stride + (size_t)(j + nb - 1) * d.ldi + (i + nb - 1)];
  };
  double dxi, dyi, dpw, dpe, dps, dpn, dpsw, dpse, dpc, dpnw, dpne, dgmx, dfmx, dfmn, q, q1, q2, q3, q4,
         tgmx, tgmn, tfmx, tfmn, sgmx, sgmn, sfmx, sfmn, xm, ym, xc0, xc1, yc0, yc1, x2, y2, x4, y4;

  // :304-337
  for (int j = 1 - mrg - 2; j <= jj + mrg + 2; ++j) {
    for (int i = 1 - mrg - 2; i <= ii + mrg + 2; ++i) if (ip(i, j) == 1) {
      dp(i, j) = std::max(0., dp(i, j)) + dpeps;
      pup(i, j) = plo(i, j) - dp(i, j);
    }
    for (int i = 1 - mrg - 1; i <= ii + mrg + 1; ++i) {
      fdu(i, j) = 0.; fdv(i, j) = 0.; ftu(i, j) = 0.; ftv(i, j) = 0.; fsu(i, j) = 0.; fsv(i, j) = 0.;
      if (ntr > 0) {
        for (int nt = 1; nt <= ntr; ++nt) { ftru(nt, i, j) = 0.; ftrv(nt, i, j) = 0.; }
        cu(i, j) = 0.;
        cv(i, j) = 0.;
      }
    }
  }
  // NB (:333-334): cu, cv are zeroed only under use_TRC in the reference (they are automatic arrays,
  // so without tracers land faces hold whatever the stack held).  The corner loop below reads cu,cv
  // on land faces only when nw==2 selects a wet pair, whose shared face is wet, or when nw==4; both
  // cases touch wet faces only, so the restatement zero-fills (vector init) without changing results.

  // :361-600 limited gradients and centre-of-mass coordinates
  for (int j = 1 - mrg - 1; j <= jj + mrg + 1; ++j)
    for (int i = 1 - mrg - 1; i <= ii + mrg + 1; ++i) if (ip(i, j) == 1) {
      const Nbr b = neighbours(i, j, ip, iu, iv);
      dxi = 1. / std::max(1, b.ie - b.iw);
      dyi = 1. / std::max(1, b.jn - b.js);
      auto lim = [&](int a, int c) { return std::max(dpeps, std::min(pbmin(i, j) - pup(a, c), dp(a, c))); };
      dpsw = lim(b.isw, b.jsw); dps = lim(i, b.js); dpse = lim(b.ise, b.jse);
      dpw = lim(b.iw, j); dpc = lim(i, j); dpe = lim(b.ie, j);
      dpnw = lim(b.inw, b.jnw); dpn = lim(i, b.jn); dpne = lim(b.ine, b.jne);
      dx(i, j) = (dpe - dpw) * dxi;
      dy(i, j) = (dpn - dps) * dyi;
      dgmx = .5 * (std::fabs(dx(i, j)) + std::fabs(dy(i, j)));
      dfmx = std::max(0., max8(dpsw, dps, dpse, dpw, dpe, dpnw, dpn, dpne) - dpc);
      dfmn = std::min(0., min8(dpsw, dps, dpse, dpw, dpe, dpnw, dpn, dpne) - dpc);
      if (dfmx > 0. && dfmn < 0.) {
        q = std::min(dfmx / std::max(dfmx, dgmx), dfmn / std::min(dfmn, -dgmx));
        dx(i, j) = dx(i, j) * q;
        dy(i, j) = dy(i, j) * q;
        xd(i, j) = dx(i, j) / (12. * dp(i, j));
        yd(i, j) = dy(i, j) / (12. * dp(i, j));
      } else {
        dx(i, j) = 0.; dy(i, j) = 0.; xd(i, j) = 0.; yd(i, j) = 0.;
      }
      // one scalar: slopes gx,gy and centre value gd from field f(i,j) (:417-474, :558-594)
      auto scalar = [&](auto f, double& gx, double& gy, double& gd) {
        gx = (f(b.ie, j) - f(b.iw, j)) * dxi;
        gy = (f(i, b.jn) - f(i, b.js)) * dyi;
        q1 = gx * (-.5 - xd(i, j));
        q2 = gx * (.5 - xd(i, j));
        q3 = gy * (-.5 - yd(i, j));
        q4 = gy * (.5 - yd(i, j));
        tgmx = std::max(q1, q2) + std::max(q3, q4);
        tgmn = std::min(q1, q2) + std::min(q3, q4);
        tfmx = std::max(0., max8(f(b.isw, b.jsw), f(i, b.js), f(b.ise, b.jse), f(b.iw, j), f(b.ie, j),
                                 f(b.inw, b.jnw), f(i, b.jn), f(b.ine, b.jne)) - f(i, j));
        tfmn = std::min(0., min8(f(b.isw, b.jsw), f(i, b.js), f(b.ise, b.jse), f(b.iw, j), f(b.ie, j),
                                 f(b.inw, b.jnw), f(i, b.jn), f(b.ine, b.jne)) - f(i, j));
        if (tfmx > 0. && tfmn < 0.) {
          q = std::min(tfmx / std::max(tfmx, tgmx), tfmn / std::min(tfmn, tgmn));
          gx = gx * q;
          gy = gy * q;
          gd = f(i, j) - gx * xd(i, j) - gy * yd(i, j);
        } else {
          gx = 0.; gy = 0.; gd = f(i, j);
        }
      };
      scalar([&](int a, int c) { return temp(a, c); }, tx(i, j), ty(i, j), td(i, j));
      scalar([&](int a, int c) { return saln(a, c); }, sx(i, j), sy(i, j), sd(i, j));
      for (int nt = 1; nt <= ntr; ++nt)
        scalar([&](int a, int c) { return trc(a, c, nt); }, trx(nt, i, j), try_(nt, i, j), trd(nt, i, j));
      (void)sgmx; (void)sgmn; (void)sfmx; (void)sfmn;
    }

  // :604-626 non-dimensional velocities
  for (int j = 1 - mrg - 1; j <= jj + mrg + 1; ++j)
    for (int i = 1 - mrg; i <= ii + mrg + 1; ++i) if (iu(i, j) == 1) {
      if (cau(i, j) > 0.) cu(i, j) = cau(i, j) * scp2i(i - 1, j);
      else cu(i, j) = cau(i, j) * scp2i(i, j);
    }
  for (int j = 1 - mrg; j <= jj + mrg + 1; ++j)
    for (int i = 1 - mrg - 1; i <= ii + mrg + 1; ++i) if (iv(i, j) == 1) {
      if (cav(i, j) > 0.) cv(i, j) = cav(i, j) * scp2i(i, j - 1);
      else cv(i, j) = cav(i, j) * scp2i(i, j);
    }

  // :639-680 corner velocities
  for (int j = 1 - mrg; j <= jj + mrg + 1; ++j)
    for (int i = 1 - mrg; i <= ii + mrg + 1; ++i) {
      const int nw = ip(i - 1, j - 1) + ip(i, j - 1) + ip(i - 1, j) + ip(i, j);
      if (nw == 4) {
        if (cu(i, j - 1) * cu(i, j) <= 0.) cuc(i, j) = 0.;
        else cuc(i, j) = 2. * cu(i, j - 1) * cu(i, j) / (cu(i, j - 1) + cu(i, j));
        if (cv(i - 1, j) * cv(i, j) <= 0.) cvc(i, j) = 0.;
        else cvc(i, j) = 2. * cv(i - 1, j) * cv(i, j) / (cv(i - 1, j) + cv(i, j));
      } else if (nw == 2) {
        if (ip(i - 1, j - 1) + ip(i, j - 1) == 2) { cuc(i, j) = cu(i, j - 1); cvc(i, j) = 0.; }
        else if (ip(i - 1, j) + ip(i, j) == 2) { cuc(i, j) = cu(i, j); cvc(i, j) = 0.; }
        else if (ip(i - 1, j - 1) + ip(i - 1, j) == 2) { cuc(i, j) = 0.; cvc(i, j) = cv(i - 1, j); }
        else if (ip(i, j - 1) + ip(i, j) == 2) { cuc(i, j) = 0.; cvc(i, j) = cv(i, j); }
        else { cuc(i, j) = 0.; cvc(i, j) = 0.; }
      } else {
        cuc(i, j) = 0.; cvc(i, j) = 0.;
      }
    }

  // common tail of every donor-cell block (e.g. :715-760): flux integrals of thickness and of the
  // linear tracer reconstructions over the polygon whose moments are `g`
  auto donor = [&](int ic, int jc, const Mom& g, double pbf, double& fd_acc, double& ft_acc, double& fs_acc,
                   T3 ftr, int fi, int fj) {
    const double dl = std::min(dp(ic, jc), std::max(0., pbf - pup(ic, jc)));
    const double fd = g.a * dl + g.ax * dx(ic, jc) + g.ay * dy(ic, jc);
    fd_acc = fd_acc + fd;
    const double qx = g.ax * dl + g.axx * dx(ic, jc) + g.axy * dy(ic, jc);
    const double qy = g.ay * dl + g.axy * dx(ic, jc) + g.ayy * dy(ic, jc);
    ft_acc = ft_acc + fd * td(ic, jc) + qx * tx(ic, jc) + qy * ty(ic, jc);
    fs_acc = fs_acc + fd * sd(ic, jc) + qx * sx(ic, jc) + qy * sy(ic, jc);
    for (int nt = 1; nt <= ntr; ++nt)
      ftr(nt, fi, fj) = ftr(nt, fi, fj) + fd * trd(nt, ic, jc) + qx * trx(nt, ic, jc) + qy * try_(nt, ic, jc);
  };

  // :688-1063 u-components of fluxes
  for (int j = 1 - mrg; j <= jj + mrg; ++j)
    for (int i = 1 - mrg; i <= ii + mrg + 1; ++i) if (iu(i, j) == 1) {
      ym = -.5 * (cvc(i, j) + cvc(i, j + 1));
      xm = ((ym + .5) * cuc(i, j) - (ym - .5) * cuc(i, j + 1) - 2. * cu(i, j)) / (1. + cvc(i, j) - cvc(i, j + 1));
      if (cu(i, j) > 0.) {
        if (cvc(i, j) > 0.) {
          xc0 = (xm * cvc(i, j) - cuc(i, j) * (ym + .5)) / (cvc(i, j) + ym + .5);
          xc1 = xc0 * scp2(i - 1, j) * scp2i(i - 1, j - 1);
          x4 = xc0 + .5;
          y4 = -.5;
          donor(i - 1, j - 1, triint(scp2(i - 1, j - 1), xc1 + .5, .5, -cuc(i, j) + .5, -cvc(i, j) + .5, .5, .5),
                pbu(i, j), fdu(i, j), ftu(i, j), fsu(i, j), ftru, i, j);
        } else {
          x4 = -cuc(i, j) + .5;
          y4 = -cvc(i, j) - .5;
        }
        if (cvc(i, j + 1) < 0.) {
          xc0 = (xm * cvc(i, j + 1) - cuc(i, j + 1) * (ym - .5)) / (cvc(i, j + 1) + ym - .5);
          xc1 = xc0 * scp2(i - 1, j) * scp2i(i - 1, j + 1);
          x2 = xc0 + .5;
          y2 = .5;
          donor(i - 1, j + 1,
                triint(scp2(i - 1, j + 1), xc1 + .5, -.5, .5, -.5, -cuc(i, j + 1) + .5, -cvc(i, j + 1) - .5),
                pbu(i, j), fdu(i, j), ftu(i, j), fsu(i, j), ftru, i, j);
        } else {
          x2 = -cuc(i, j + 1) + .5;
          y2 = -cvc(i, j + 1) + .5;
        }
        donor(i - 1, j, penint(scp2(i - 1, j), .5, .5, x2, y2, xm + .5, ym, x4, y4, .5, -.5), pbu(i, j),
              fdu(i, j), ftu(i, j), fsu(i, j), ftru, i, j);
      } else {
        if (cvc(i, j) > 0.) {
          xc0 = (xm * cvc(i, j) - cuc(i, j) * (ym + .5)) / (cvc(i, j) + ym + .5);
          xc1 = xc0 * scp2(i, j) * scp2i(i, j - 1);
          x4 = xc0 - .5;
          y4 = -.5;
          donor(i, j - 1, triint(scp2(i, j - 1), xc1 - .5, .5, -cuc(i, j) - .5, -cvc(i, j) + .5, -.5, .5),
                pbu(i, j), fdu(i, j), ftu(i, j), fsu(i, j), ftru, i, j);
        } else {
          x4 = -cuc(i, j) - .5;
          y4 = -cvc(i, j) - .5;
        }
        if (cvc(i, j + 1) < 0.) {
          xc0 = (xm * cvc(i, j + 1) - cuc(i, j + 1) * (ym - .5)) / (cvc(i, j + 1) + ym - .5);
          xc1 = xc0 * scp2(i, j) * scp2i(i, j + 1);
          x2 = xc0 - .5;
          y2 = .5;
          donor(i, j + 1,
                triint(scp2(i, j + 1), xc1 - .5, -.5, -.5, -.5, -cuc(i, j + 1) - .5, -cvc(i, j + 1) - .5),
                pbu(i, j), fdu(i, j), ftu(i, j), fsu(i, j), ftru, i, j);
        } else {
          x2 = -cuc(i, j + 1) - .5;
          y2 = -cvc(i, j + 1) + .5;
        }
        donor(i, j, penint(scp2(i, j), -.5, .5, x2, y2, xm - .5, ym, x4, y4, -.5, -.5), pbu(i, j), fdu(i, j),
              ftu(i, j), fsu(i, j), ftru, i, j);
      }
      uflx(i, j) = uflx(i, j) + fdu(i, j);
      utflx(i, j) = utflx(i, j) + ftu(i, j);
      usflx(i, j) = usflx(i, j) + fsu(i, j);
    }

  // :1067-1462 v-components of fluxes
  for (int j = 1 - mrg; j <= jj + mrg + 1; ++j)
    for (int i = 1 - mrg; i <= ii + mrg; ++i) if (iv(i, j) == 1) {
      xm = -.5 * (cuc(i, j) + cuc(i + 1, j));
      ym = ((xm + .5) * cvc(i, j) - (xm - .5) * cvc(i + 1, j) - 2. * cv(i, j)) / (1. + cuc(i, j) - cuc(i + 1, j));
      if (cv(i, j) > 0) {
        if (cuc(i, j) > 0.) {
          yc0 = (ym * cuc(i, j) - cvc(i, j) * (xm + .5)) / (cuc(i, j) + xm + .5);
          yc1 = yc0 * scp2(i, j - 1) * scp2i(i - 1, j - 1);
          x2 = -.5;
          y2 = yc0 + .5;
          donor(i - 1, j - 1, triint(scp2(i - 1, j - 1), .5, yc1 + .5, .5, .5, -cuc(i, j) + .5, -cvc(i, j) + .5),
                pbv(i, j), fdv(i, j), ftv(i, j), fsv(i, j), ftrv, i, j);
        } else {
          x2 = -cuc(i, j) - .5;
          y2 = -cvc(i, j) + .5;
        }
        if (cuc(i + 1, j) < 0.) {
          yc0 = (ym * cuc(i + 1, j) - cvc(i + 1, j) * (xm - .5)) / (cuc(i + 1, j) + xm - .5);
          yc1 = yc0 * scp2(i, j - 1) * scp2i(i + 1, j - 1);
          x4 = .5;
          y4 = yc0 + .5;
          donor(i + 1, j - 1,
                triint(scp2(i + 1, j - 1), -.5, yc1 + .5, -cuc(i + 1, j) - .5, -cvc(i + 1, j) + .5, -.5, .5),
                pbv(i, j), fdv(i, j), ftv(i, j), fsv(i, j), ftrv, i, j);
        } else {
          x4 = -cuc(i + 1, j) + .5;
          y4 = -cvc(i + 1, j) + .5;
        }
        donor(i, j - 1, penint(scp2(i, j - 1), -.5, .5, x2, y2, xm, ym + .5, x4, y4, .5, .5), pbv(i, j),
              fdv(i, j), ftv(i, j), fsv(i, j), ftrv, i, j);
      } else {
        if (cuc(i, j) > 0.) {
          yc0 = (ym * cuc(i, j) - cvc(i, j) * (xm + .5)) / (cuc(i, j) + xm + .5);
          yc1 = yc0 * scp2(i, j) * scp2i(i - 1, j);
          x2 = -.5;
          y2 = yc0 - .5;
          donor(i - 1, j, triint(scp2(i - 1, j), .5, yc1 - .5, .5, -.5, -cuc(i, j) + .5, -cvc(i, j) - .5),
                pbv(i, j), fdv(i, j), ftv(i, j), fsv(i, j), ftrv, i, j);
        } else {
          x2 = -cuc(i, j) - .5;
          y2 = -cvc(i, j) - .5;
        }
        if (cuc(i + 1, j) < 0.) {
          yc0 = (ym * cuc(i + 1, j) - cvc(i + 1, j) * (xm - .5)) / (cuc(i + 1, j) + xm - .5);
          yc1 = yc0 * scp2(i, j) * scp2i(i + 1, j);
          x4 = .5;
          y4 = yc0 - .5;
          donor(i + 1, j,
                triint(scp2(i + 1, j), -.5, yc1 - .5, -cuc(i + 1, j) - .5, -cvc(i + 1, j) - .5, -.5, -.5),
                pbv(i, j), fdv(i, j), ftv(i, j), fsv(i, j), ftrv, i, j);
        } else {
          x4 = -cuc(i + 1, j) + .5;
          y4 = -cvc(i + 1, j) - .5;
        }
        donor(i, j, penint(scp2(i, j), -.5, -.5, x2, y2, xm, ym - .5, x4, y4, .5, -.5), pbv(i, j), fdv(i, j),
              ftv(i, j), fsv(i, j), ftrv, i, j);
      }
      // reference quirk (:1455-1457): the v fluxes are assigned, the u fluxes accumulated (:1054-1056)
      vflx(i, j) = fdv(i, j);
      vtflx(i, j) = ftv(i, j);
      vsflx(i, j) = fsv(i, j);
    }

  // :1468-1520 update
  for (int j = 1 - mrg; j <= jj + mrg; ++j)
    for (int i = 1 - mrg; i <= ii + mrg; ++i) if (ip(i, j) == 1) {
      q = dp(i, j);
      dp(i, j) = q - (fdu(i + 1, j) - fdu(i, j) + fdv(i, j + 1) - fdv(i, j)) * scp2i(i, j);
      temp(i, j) = (q * temp(i, j) - (ftu(i + 1, j) - ftu(i, j) + ftv(i, j + 1) - ftv(i, j)) * scp2i(i, j)) / dp(i, j);
      saln(i, j) = (q * saln(i, j) - (fsu(i + 1, j) - fsu(i, j) + fsv(i, j + 1) - fsv(i, j)) * scp2i(i, j)) / dp(i, j);
      for (int nt = 1; nt <= ntr; ++nt)
        trc(i, j, nt) = (q * trc(i, j, nt) -
                         (ftru(nt, i + 1, j) - ftru(nt, i, j) + ftrv(nt, i, j + 1) - ftrv(nt, i, j)) * scp2i(i, j)) /
                        dp(i, j);
      dp(i, j) = std::max(0., dp(i, j) - dpeps);
    }
}

}  // namespace

// phy/mod_advect.F90:96-153 (advmth='remap'), called after the flux-area prelude :71-94
void advect_remap(int m, int n, int mm, int nn, int k1m, int k1n) {
  (void)m; (void)k1m;
  Oracle& o = O(); const Dims& d = o.d;
  const int ii = d.ii, jj = d.jj, kk = d.kk;
  I2 ip = o.i2("ip"), iu = o.i2("iu"), iv = o.i2("iv");
  A3 p = o.a3("p");
  A2 pbmin = o.scratch("_pbmin", 1).level(1);
  for (int j = -1; j <= jj + 2; ++j)
    for (int i = -1; i <= ii + 2; ++i) if (ip(i, j) == 1) {
      const Nbr b = neighbours(i, j, ip, iu, iv);
      pbmin(i, j) = std::min(
          std::min(std::min(std::min(p(b.isw, b.jsw, kk + 1), p(i, b.js, kk + 1)), p(b.ise, b.jse, kk + 1)),
                   std::min(std::min(p(b.iw, j, kk + 1), p(i, j, kk + 1)), p(b.ie, j, kk + 1))),
          std::min(std::min(p(b.inw, b.jnw, kk + 1), p(i, b.jn, kk + 1)), p(b.ine, b.jne, kk + 1)));
    }
  xctilr(o.a3("cau"), 1, kk, 3, 3, halo_uv);
  xctilr(o.a3("cav"), 1, kk, 3, 3, halo_vv);
  for (int nt = 1; nt <= d.ntr; ++nt) xctilr(o.a3("trc").from(k1n + (nt - 1) * 2 * d.kdm), 1, kk, 3, 3, halo_ps);
  A3 dp = o.a3("dp"), temp = o.a3("temp"), saln = o.a3("saln"), cau = o.a3("cau"), cav = o.a3("cav");
  A3 uflx = o.a3("uflx"), vflx = o.a3("vflx"), utflx = o.a3("utflx"), vtflx = o.a3("vtflx"),
     usflx = o.a3("usflx"), vsflx = o.a3("vsflx"), pbu = o.a3("pbu"), pbv = o.a3("pbv");
  double* trc = d.ntr > 0 ? o.a3("trc").p : nullptr;
  for (int k = 1; k <= kk; ++k) {
    const int km = k + mm, kn = k + nn;
    remap(o.a2("scp2i"), o.a2("scp2"), pbmin, pbu.level(n), pbv.level(n), p.level(k + 1), cau.level(k),
          cav.level(k), 1, dp.level(kn), temp.level(kn), saln.level(kn), uflx.level(km), vflx.level(km),
          utflx.level(km), vtflx.level(km), usflx.level(km), vsflx.level(km),
          trc ? trc + (size_t)(kn - 1) * d.lev : nullptr, (size_t)2 * d.kdm * d.lev);
  }
}

}  // namespace orc
