// Device functions of the equation of state (phy/mod_eos.F90): 18-coefficient
// rational fit rho = P1/P2 (:36-54), potential density sig/sig0 (:191-218),
// truncated-series pressure integrals p_alpha/delphi (:386-529), thermal and
// haline derivatives of specific volume (:531-574) and of dynamic enthalpy
// (:576-695).  All inlined into the kernels that need them; the pref-dependent
// coefficients of inieos (:83-155) are passed to kernels by value.
#pragma once
#include "common.cuh"

namespace blom { namespace eos {

#define EA11 9.9985372432159340e+02
#define EA12 1.0380621928183473e+01
#define EA13 1.7073577195684715e+00
#define EA14 (-3.6570490496333680e-02)
#define EA15 (-7.3677944503527477e-03)
#define EA16 (-3.5529175999643348e-03)
#define EB11 1.7083494994335439e-06
#define EB12 7.1567921402953455e-09
#define EB13 1.2821026080049485e-09
#define EA21 1.0
#define EA22 1.0316374535350838e-02
#define EA23 8.9521792365142522e-04
#define EA24 (-2.8438341552142710e-05)
#define EA25 (-1.1887778959461776e-05)
#define EA26 (-4.0163964812921489e-06)
#define EB21 1.1995545126831476e-09
#define EB22 5.5234008384648383e-12
#define EB23 8.4310335919950873e-13

struct Coef {
  double pref;
  double ap11, ap12, ap13, ap14, ap15, ap16, ap21, ap22, ap23, ap24, ap25, ap26;
  double ap110, ap120, ap130, ap140, ap150, ap160, ap210, ap220, ap230, ap240, ap250, ap260;
};
// host copy filled by inieos_dev(); kernels take it by value
const Coef& host_coef();

__device__ __forceinline__ double P1(double p, double th, double s) {
  return EA11 + (EA12 + EA14 * th + EA15 * s) * th + (EA13 + EA16 * s) * s + (EB11 + EB12 * th + EB13 * s) * p;
}
__device__ __forceinline__ double P2(double p, double th, double s) {
  return EA21 + (EA22 + EA24 * th + EA25 * s) * th + (EA23 + EA26 * s) * s + (EB21 + EB22 * th + EB23 * s) * p;
}
__device__ __forceinline__ double rho(double p, double th, double s) { return P1(p, th, s) / P2(p, th, s); }
__device__ __forceinline__ double alp(double p, double th, double s) { return P2(p, th, s) / P1(p, th, s); }
__device__ __forceinline__ double sig(const Coef& c, double th, double s) {
  return (c.ap11 + (c.ap12 + c.ap14 * th + c.ap15 * s) * th + (c.ap13 + c.ap16 * s) * s) /
         (c.ap21 + (c.ap22 + c.ap24 * th + c.ap25 * s) * th + (c.ap23 + c.ap26 * s) * s);
}
__device__ __forceinline__ double sig0(const Coef& c, double th, double s) {
  return (c.ap110 + (c.ap120 + c.ap140 * th + c.ap150 * s) * th + (c.ap130 + c.ap160 * s) * s) /
         (c.ap210 + (c.ap220 + c.ap240 * th + c.ap250 * s) * th + (c.ap230 + c.ap260 * s) * s);
}
__device__ __forceinline__ double p_alpha(double p1, double p2, double th, double s) {
  const double r1_3 = 1. / 3., r1_5 = 1. / 5., r1_7 = 1. / 7., r1_9 = 1. / 9.;
  const double a1 = EA11 + (EA12 + EA14 * th + EA15 * s) * th + (EA13 + EA16 * s) * s;
  const double a2 = EA21 + (EA22 + EA24 * th + EA25 * s) * th + (EA23 + EA26 * s) * s;
  const double b1 = EB11 + EB12 * th + EB13 * s;
  const double b2 = EB21 + EB22 * th + EB23 * s;
  const double pm = .5 * (p2 + p1);
  const double r = .5 * (p2 - p1) / (a1 + b1 * pm);
  const double q = b1 * r;
  const double qq = q * q;
  return 2. * r * (a2 + b2 * pm + (a2 - a1 * b2 / b1) * qq * (r1_3 + qq * (r1_5 + qq * (r1_7 + qq * r1_9))));
}
__device__ __forceinline__ void delphi(double p1, double p2, double th, double s, double& dphi, double& alp1,
                                       double& alp2) {
  const double r1_3 = 1. / 3., r1_5 = 1. / 5., r1_7 = 1. / 7., r1_9 = 1. / 9.;
  const double a1 = EA11 + (EA12 + EA14 * th + EA15 * s) * th + (EA13 + EA16 * s) * s;
  const double a2 = EA21 + (EA22 + EA24 * th + EA25 * s) * th + (EA23 + EA26 * s) * s;
  const double b1 = EB11 + EB12 * th + EB13 * s;
  const double b2 = EB21 + EB22 * th + EB23 * s;
  const double pm = .5 * (p2 + p1);
  const double r = .5 * (p2 - p1) / (a1 + b1 * pm);
  const double q = b1 * r;
  const double qq = q * q;
  dphi = -2. * r * (a2 + b2 * pm + (a2 - a1 * b2 / b1) * qq * (r1_3 + qq * (r1_5 + qq * (r1_7 + qq * r1_9))));
  alp1 = (a2 + b2 * p1) / (a1 + b1 * p1);
  alp2 = (a2 + b2 * p2) / (a1 + b1 * p2);
}
__device__ __forceinline__ double dalpdt(double p, double th, double s) {
  const double r1 = P2(p, th, s);
  const double r2i = 1. / P1(p, th, s);
  return (EA22 + 2. * EA24 * th + EA25 * s + EB22 * p - (EA12 + 2. * EA14 * th + EA15 * s + EB12 * p) * r1 * r2i) * r2i;
}
__device__ __forceinline__ double dalpds(double p, double th, double s) {
  const double r1 = P2(p, th, s);
  const double r2i = 1. / P1(p, th, s);
  return (EA23 + EA25 * th + 2. * EA26 * s + EB23 * p - (EA13 + EA15 * th + 2. * EA16 * s + EB13 * p) * r1 * r2i) * r2i;
}
__device__ __forceinline__ void dynh_derivatives(double p0, double p1, double p2, double th, double s,
                                                 double& dynh_th, double& dynh_s) {
  const double r1_2 = 1. / 2., r1_3 = 1. / 3., r1_4 = 1. / 4., r1_5 = 1. / 5., r1_6 = 1. / 6.,
               r1_7 = 1. / 7., r1_8 = 1. / 8., r1_9 = 1. / 9., r1_10 = 1. / 10., r1_11 = 1. / 11.;
  const double b1i = 1. / (EB11 + EB12 * th + EB13 * s);
  const double a1 = (EA11 + (EA12 + EA14 * th + EA15 * s) * th + (EA13 + EA16 * s) * s) * b1i;
  const double a2 = (EA21 + (EA22 + EA24 * th + EA25 * s) * th + (EA23 + EA26 * s) * s) * b1i;
  const double b2 = (EB21 + EB22 * th + EB23 * s) * b1i;
  const double a1_th = (EA12 + 2. * EA14 * th + EA15 * s - a1 * EB12) * b1i;
  const double a2_th = (EA22 + 2. * EA24 * th + EA25 * s - a2 * EB12) * b1i;
  const double b2_th = (EB22 - b2 * EB12) * b1i;
  const double a1_s = (EA13 + EA15 * th + 2. * EA16 * s - a1 * EB13) * b1i;
  const double a2_s = (EA23 + EA25 * th + 2. * EA26 * s - a2 * EB13) * b1i;
  const double b2_s = (EB23 - b2 * EB13) * b1i;
  const double pm1 = r1_2 * (p2 + p1), pp1 = r1_2 * (p2 - p1), pm0 = r1_2 * (pm1 + p0), pp0 = r1_2 * (pm1 - p0);
  const double t1 = 1. / (a1 + pm1), t0 = 1. / (a1 + pm0);
  const double q1 = pp1 * t1, q0 = pp0 * t0, qq1 = q1 * q1, qq0 = q0 * q0;
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

}}  // namespace blom::eos
