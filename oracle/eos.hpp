// ORACLE — TEST INFRASTRUCTURE ONLY (see core.hpp header).
// Restatement of phy/mod_eos.F90: coefficients :36-54, inieos :83-155 and the
// pure functions used on the hot path.
#pragma once
#include "core.hpp"

namespace orc { namespace eos {

constexpr double a11 = 9.9985372432159340e+02, a12 = 1.0380621928183473e+01,
  a13 = 1.7073577195684715e+00, a14 = -3.6570490496333680e-02, a15 = -7.3677944503527477e-03,
  a16 = -3.5529175999643348e-03, b11 = 1.7083494994335439e-06, b12 = 7.1567921402953455e-09,
  b13 = 1.2821026080049485e-09, a21 = 1.0, a22 = 1.0316374535350838e-02,
  a23 = 8.9521792365142522e-04, a24 = -2.8438341552142710e-05, a25 = -1.1887778959461776e-05,
  a26 = -4.0163964812921489e-06, b21 = 1.1995545126831476e-09, b22 = 5.5234008384648383e-12,
  b23 = 8.4310335919950873e-13;

struct Coef {
  double pref = 0;
  double ap11, ap12, ap13, ap14, ap15, ap16, ap21, ap22, ap23, ap24, ap25, ap26;
  double ap110, ap120, ap130, ap140, ap150, ap160, ap210, ap220, ap230, ap240, ap250, ap260;
};
Coef& K();
void inieos_pref(double pref);  // :105-129

inline double rho(double p, double th, double s) {  // :157-172
  return (a11 + (a12 + a14 * th + a15 * s) * th + (a13 + a16 * s) * s + (b11 + b12 * th + b13 * s) * p) /
         (a21 + (a22 + a24 * th + a25 * s) * th + (a23 + a26 * s) * s + (b21 + b22 * th + b23 * s) * p);
}
inline double alp(double p, double th, double s) {  // :174-189
  return (a21 + (a22 + a24 * th + a25 * s) * th + (a23 + a26 * s) * s + (b21 + b22 * th + b23 * s) * p) /
         (a11 + (a12 + a14 * th + a15 * s) * th + (a13 + a16 * s) * s + (b11 + b12 * th + b13 * s) * p);
}
inline double sig(double th, double s) {  // :191-203
  const Coef& c = K();
  return (c.ap11 + (c.ap12 + c.ap14 * th + c.ap15 * s) * th + (c.ap13 + c.ap16 * s) * s) /
         (c.ap21 + (c.ap22 + c.ap24 * th + c.ap25 * s) * th + (c.ap23 + c.ap26 * s) * s);
}
inline double sig0(double th, double s) {  // :205-218
  const Coef& c = K();
  return (c.ap110 + (c.ap120 + c.ap140 * th + c.ap150 * s) * th + (c.ap130 + c.ap160 * s) * s) /
         (c.ap210 + (c.ap220 + c.ap240 * th + c.ap250 * s) * th + (c.ap230 + c.ap260 * s) * s);
}
inline double p_alpha(double p1, double p2, double th, double s) {  // :386-428
  const double r1_3 = 1. / 3., r1_5 = 1. / 5., r1_7 = 1. / 7., r1_9 = 1. / 9.;
  double a1 = a11 + (a12 + a14 * th + a15 * s) * th + (a13 + a16 * s) * s;
  double a2 = a21 + (a22 + a24 * th + a25 * s) * th + (a23 + a26 * s) * s;
  double b1 = b11 + b12 * th + b13 * s;
  double b2 = b21 + b22 * th + b23 * s;
  double pm = .5 * (p2 + p1);
  double r = .5 * (p2 - p1) / (a1 + b1 * pm);
  double q = b1 * r;
  double qq = q * q;
  return 2. * r * (a2 + b2 * pm + (a2 - a1 * b2 / b1) * qq * (r1_3 + qq * (r1_5 + qq * (r1_7 + qq * r1_9))));
}
inline void delphi(double p1, double p2, double th, double s, double& dphi, double& alp1, double& alp2) {  // :478-529
  const double r1_3 = 1. / 3., r1_5 = 1. / 5., r1_7 = 1. / 7., r1_9 = 1. / 9.;
  double a1 = a11 + (a12 + a14 * th + a15 * s) * th + (a13 + a16 * s) * s;
  double a2 = a21 + (a22 + a24 * th + a25 * s) * th + (a23 + a26 * s) * s;
  double b1 = b11 + b12 * th + b13 * s;
  double b2 = b21 + b22 * th + b23 * s;
  double pm = .5 * (p2 + p1);
  double r = .5 * (p2 - p1) / (a1 + b1 * pm);
  double q = b1 * r;
  double qq = q * q;
  dphi = -2. * r * (a2 + b2 * pm + (a2 - a1 * b2 / b1) * qq * (r1_3 + qq * (r1_5 + qq * (r1_7 + qq * r1_9))));
  alp1 = (a2 + b2 * p1) / (a1 + b1 * p1);
  alp2 = (a2 + b2 * p2) / (a1 + b1 * p2);
}
inline double dalpdt(double p, double th, double s) {  // :531-552
  double r1 = a21 + (a22 + a24 * th + a25 * s) * th + (a23 + a26 * s) * s + (b21 + b22 * th + b23 * s) * p;
  double r2i = 1. / (a11 + (a12 + a14 * th + a15 * s) * th + (a13 + a16 * s) * s + (b11 + b12 * th + b13 * s) * p);
  return (a22 + 2. * a24 * th + a25 * s + b22 * p - (a12 + 2. * a14 * th + a15 * s + b12 * p) * r1 * r2i) * r2i;
}
inline double dalpds(double p, double th, double s) {  // :554-574
  double r1 = a21 + (a22 + a24 * th + a25 * s) * th + (a23 + a26 * s) * s + (b21 + b22 * th + b23 * s) * p;
  double r2i = 1. / (a11 + (a12 + a14 * th + a15 * s) * th + (a13 + a16 * s) * s + (b11 + b12 * th + b13 * s) * p);
  return (a23 + a25 * th + 2. * a26 * s + b23 * p - (a13 + a15 * th + 2. * a16 * s + b13 * p) * r1 * r2i) * r2i;
}
inline void dynh_derivatives(double p0, double p1, double p2, double th, double s, double& dynh_th,
                             double& dynh_s) {  // :576-695
  const double r1_2 = 1. / 2., r1_3 = 1. / 3., r1_4 = 1. / 4., r1_5 = 1. / 5., r1_6 = 1. / 6.,
               r1_7 = 1. / 7., r1_8 = 1. / 8., r1_9 = 1. / 9., r1_10 = 1. / 10., r1_11 = 1. / 11.;
  double b1i = 1. / (b11 + b12 * th + b13 * s);
  double a1 = (a11 + (a12 + a14 * th + a15 * s) * th + (a13 + a16 * s) * s) * b1i;
  double a2 = (a21 + (a22 + a24 * th + a25 * s) * th + (a23 + a26 * s) * s) * b1i;
  double b2 = (b21 + b22 * th + b23 * s) * b1i;
  double a1_th = (a12 + 2. * a14 * th + a15 * s - a1 * b12) * b1i;
  double a2_th = (a22 + 2. * a24 * th + a25 * s - a2 * b12) * b1i;
  double b2_th = (b22 - b2 * b12) * b1i;
  double a1_s = (a13 + a15 * th + 2. * a16 * s - a1 * b13) * b1i;
  double a2_s = (a23 + a25 * th + 2. * a26 * s - a2 * b13) * b1i;
  double b2_s = (b23 - b2 * b13) * b1i;
  double pm1 = r1_2 * (p2 + p1), pp1 = r1_2 * (p2 - p1), pm0 = r1_2 * (pm1 + p0), pp0 = r1_2 * (pm1 - p0);
  double t1 = 1. / (a1 + pm1), t0 = 1. / (a1 + pm0);
  double q1 = pp1 * t1, q0 = pp0 * t0, qq1 = q1 * q1, qq0 = q0 * q0;
  double f = (a2 - a1 * b2) * a1_th;
  double c1 = a2_th - a1 * b2_th - b2 * a1_th;
  double c2 = f * t1, c3 = f * t0;
  dynh_th = 2. * (pp0 * b2_th + ((((((r1_11 * c1 - c3) * qq0 + (r1_9 * c1 - c3)) * qq0 + (r1_7 * c1 - c3)) * qq0 +
                                   (r1_5 * c1 - c3)) * qq0 + (r1_3 * c1 - c3)) * qq0 + (c1 - c3)) * q0) -
            ((((r1_11 * (r1_10 * c1 - c2) * qq1 + r1_9 * (r1_8 * c1 - c2)) * qq1 + r1_7 * (r1_6 * c1 - c2)) * qq1 +
              r1_5 * (r1_4 * c1 - c2)) * qq1 + r1_3 * (r1_2 * c1 - c2)) * qq1;
  f = (a2 - a1 * b2) * a1_s;
  c1 = a2_s - a1 * b2_s - b2 * a1_s;
  c2 = f * t1; c3 = f * t0;
  dynh_s = 2. * (pp0 * b2_s + ((((((r1_11 * c1 - c3) * qq0 + (r1_9 * c1 - c3)) * qq0 + (r1_7 * c1 - c3)) * qq0 +
                                  (r1_5 * c1 - c3)) * qq0 + (r1_3 * c1 - c3)) * qq0 + (c1 - c3)) * q0) -
           ((((r1_11 * (r1_10 * c1 - c2) * qq1 + r1_9 * (r1_8 * c1 - c2)) * qq1 + r1_7 * (r1_6 * c1 - c2)) * qq1 +
             r1_5 * (r1_4 * c1 - c2)) * qq1 + r1_3 * (r1_2 * c1 - c2)) * qq1;
}

}}  // namespace orc::eos
