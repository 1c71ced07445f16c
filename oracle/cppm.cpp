// ORACLE — TEST INFRASTRUCTURE ONLY (see core.hpp header).
// Restatement of phy/mod_cppm.F90 (all four variants: cppm_compatibility =
// 'full'|'partial' x cppm_limiting = 'non_oscillatory'|'monotonic', :44-48) and of
// phy/mod_advect.F90:59-189 (advmth='cppm').
#include "core.hpp"

namespace orc {

void advect_remap(int m, int n, int mm, int nn, int k1m, int k1n);  // remap.cpp

namespace {

// stencil tags, phy/mod_cppm.F90:60-68
enum { stencil_0000 = 0, stencil_1111 = 1, stencil_1110 = 2, stencil_0111 = 3,
       stencil_1100 = 4, stencil_0110 = 5, stencil_0011 = 6, stencil_0100 = 7,
       stencil_0010 = 8 };

constexpr double c0 = 0., c1 = 1., c2 = 2., c3 = 3., c4 = 4., c5 = 5., c6 = 6.,
  c12 = 12., c18 = 18., c42 = 42., c60 = 60., c1_2 = 1. / 2., c1_3 = 1. / 3.,
  c2_3 = 2. / 3., c1_4 = 1. / 4., c3_4 = 3. / 4., c1_5 = 1. / 5., c1_6 = 1. / 6.,
  c1_10 = 1. / 10., c1_12 = 1. / 12., c1_15 = 1. / 15., c1_20 = 1. / 20.,
  dpeps = 1.e-12;  // :70-77

// 1-D pencil view with Fortran lower bound 1-nbdy
struct P1 {
  double* p; int nb;
  inline double& operator()(int i) const { return p[i + nb - 1]; }
};
struct PI1 {
  int* p; int nb;
  inline int& operator()(int i) const { return p[i + nb - 1]; }
};
// (nt, i) pencil view, nt fastest, nt from 1
struct P2 {
  double* p; int nb; int nt;
  inline double& operator()(int t, int i) const { return p[(size_t)(i + nb - 1) * nt + (t - 1)]; }
};
// (12, i) coefficient pencil
struct PC {
  double* p; int nb;
  inline double& operator()(int r, int i) const { return p[(size_t)(i + nb - 1) * 12 + (r - 1)]; }
};

// module-private tables of mod_cppm (:79-91).  i-direction tables are stored
// (i,j) like any 2-D field; j-direction tables are stored transposed (j,i).
struct Tables {
  int ntr_loc = 0;
  std::vector<int> stencili, stencilj;
  std::vector<double> tmc0i, tmcli, tmcri, tmc0j, tmclj, tmcrj;
  std::vector<double> hevc1i, hevc2i, hevc3i, hevc4i, ssci, scci, d2mi;
  std::vector<double> hevc1j, hevc2j, hevc3j, hevc4j, sscj, sccj, d2mj;
  std::vector<double> hel_3d, her_3d;
} T;

// :101-320
void set_stencil_coeffs(const int* sm, const double* dx, int& stencil,
                        double& hevc1, double& hevc2, double& hevc3, double& hevc4,
                        double* tmc0, double* tmcl, double* tmcr) {
  // dx[0..3] = dx(1..4); tmc*[0..11] = tmc*(1..12)
  double a12, a22, a32, a42, a13, a23, a33, a43, a14, a24, a34, a44;
  const double d1 = dx[0], d2 = dx[1], d3 = dx[2], d4 = dx[3];
  a12 = -d2 - c1_2 * d1;
  a22 = -c1_2 * d2;
  a32 = c1_2 * d3;
  a42 = d3 + c1_2 * d4;
  a13 = a12 * a12 + c1_12 * d1 * d1;
  a23 = -c2_3 * a22 * d2;
  a33 = c2_3 * a32 * d3;
  a43 = a42 * a42 + c1_12 * d4 * d4;
  a14 = (a13 + c1_6 * d1 * d1) * a12;
  a24 = -c3_4 * a23 * d2;
  a34 = c3_4 * a33 * d3;
  a44 = (a43 + c1_6 * d4 * d4) * a42;

  tmcl[0] = -c1_12 * d1;
  tmcl[1] = (c1_10 * d1 + c1_6 * d2) * d1;
  tmcl[2] = -(c1_10 * (d1 + c3 * d2) * d1 + c1_4 * (d2 * d2)) * d1;
  tmcl[3] = -c1_12 * d2;
  tmcl[4] = c1_10 * (d2 * d2);
  tmcl[5] = -c1_10 * (d2 * d2 * d2);
  tmcl[6] = -c1_12 * d3;
  tmcl[7] = -c1_15 * (d3 * d3);
  tmcl[8] = -c1_20 * (d3 * d3 * d3);
  tmcl[9] = -c1_12 * d4;
  tmcl[10] = -(c1_15 * d4 + c1_6 * d3) * d4;
  tmcl[11] = -(c1_5 * (c1_4 * d4 + d3) * d4 + c1_4 * (d3 * d3)) * d4;

  tmcr[0] = c1_12 * d1;
  tmcr[1] = -(c1_15 * d1 + c1_6 * d2) * d1;
  tmcr[2] = (c1_5 * (c1_4 * d1 + d2) * d1 + c1_4 * (d2 * d2)) * d1;
  tmcr[3] = c1_12 * d2;
  tmcr[4] = -c1_15 * (d2 * d2);
  tmcr[5] = c1_20 * (d2 * d2 * d2);
  tmcr[6] = c1_12 * d3;
  tmcr[7] = c1_10 * (d3 * d3);
  tmcr[8] = c1_10 * (d3 * d3 * d3);
  tmcr[9] = c1_12 * d4;
  tmcr[10] = (c1_10 * d4 + c1_6 * d3) * d4;
  tmcr[11] = (c1_10 * (d4 + c3 * d3) * d4 + c1_4 * (d3 * d3)) * d4;

  tmc0[0] = a12;
  tmc0[1] = a13 - tmcl[1] - tmcr[1];
  tmc0[2] = a14 - tmcl[2] - tmcr[2];
  tmc0[3] = a22;
  tmc0[4] = a23 - tmcl[4] - tmcr[4];
  tmc0[5] = a24 - tmcl[5] - tmcr[5];
  tmc0[6] = a32;
  tmc0[7] = a33 - tmcl[7] - tmcr[7];
  tmc0[8] = a34 - tmcl[8] - tmcr[8];
  tmc0[9] = a42;
  tmc0[10] = a43 - tmcl[10] - tmcr[10];
  tmc0[11] = a44 - tmcl[11] - tmcr[11];

  auto eq4 = [&](int a, int b, int c, int e) {
    return sm[0] == a && sm[1] == b && sm[2] == c && sm[3] == e;
  };
  if (eq4(1, 1, 1, 1)) {
    stencil = stencil_1111;
    a22 = a22 - a12; a32 = a32 - a12; a42 = a42 - a12;
    a23 = (a23 - a13) / a22;
    a33 = a33 - a13 - a23 * a32;
    a43 = a43 - a13 - a23 * a42;
    a24 = (a24 - a14) / a22;
    a34 = a34 - a14 - a24 * a32;
    a44 = a44 - a14 - a24 * a42;
    a34 = a34 / a33;
    a44 = a44 - a34 * a43;
    hevc2 = -a12;
    hevc3 = -a13 - a23 * hevc2;
    hevc4 = -a14 - a24 * hevc2 - a34 * hevc3;
    hevc4 = hevc4 / a44;
    hevc3 = (hevc3 - a43 * hevc4) / a33;
    hevc2 = (hevc2 - a32 * hevc3 - a42 * hevc4) / a22;
    hevc1 = c1 - hevc2 - hevc3 - hevc4;
  } else if (eq4(1, 1, 1, 0)) {
    stencil = stencil_1110;
    a22 = a22 - a12; a32 = a32 - a12;
    a23 = (a23 - a13) / a22;
    a33 = a33 - a13 - a23 * a32;
    hevc2 = -a12;
    hevc3 = -a13 - a23 * hevc2;
    hevc3 = hevc3 / a33;
    hevc2 = (hevc2 - a32 * hevc3) / a22;
    hevc1 = c1 - hevc2 - hevc3;
    hevc4 = c0;
  } else if (eq4(0, 1, 1, 1)) {
    stencil = stencil_0111;
    a32 = a32 - a22; a42 = a42 - a22;
    a33 = (a33 - a23) / a32;
    a43 = a43 - a23 - a33 * a42;
    hevc3 = -a22;
    hevc4 = -a23 - a33 * hevc3;
    hevc4 = hevc4 / a43;
    hevc3 = (hevc3 - a42 * hevc4) / a32;
    hevc2 = c1 - hevc3 - hevc4;
    hevc1 = c0;
  } else if (eq4(0, 1, 1, 0)) {
    stencil = stencil_0110;
    a32 = a32 - a22;
    hevc3 = -a22 / a32;
    hevc2 = c1 - hevc3;
    hevc1 = c0; hevc4 = c0;
  } else if (sm[0] == 1 && sm[1] == 1) {
    stencil = stencil_1100;
    a22 = a22 - a12;
    hevc2 = -a12 / a22;
    hevc1 = c1 - hevc2;
    hevc3 = c0; hevc4 = c0;
  } else if (sm[2] == 1 && sm[3] == 1) {
    stencil = stencil_0011;
    a42 = a42 - a32;
    hevc4 = -a32 / a42;
    hevc3 = c1 - hevc4;
    hevc1 = c0; hevc2 = c0;
  } else if (sm[1] == 1) {
    stencil = stencil_0100;
    hevc1 = c0; hevc2 = c1; hevc3 = c0; hevc4 = c0;
  } else if (sm[2] == 1) {
    stencil = stencil_0010;
    hevc1 = c0; hevc2 = c0; hevc3 = c1; hevc4 = c0;
  } else {
    stencil = stencil_0000;
    hevc1 = c0; hevc2 = c0; hevc3 = c0; hevc4 = c0;
  }
}

// :322-341
void set_slope_coeffs(const int* sm, const double* dx, double& ssc, double& scc) {
  if (sm[0] == 0 || sm[1] == 0 || sm[2] == 0) { ssc = c0; scc = c0; }
  else { ssc = c2; scc = c2 * dx[1] / (dx[0] + c2 * dx[1] + dx[2]); }
}
// :343-359
void set_d2_mask(const int* sm, double& d2m) {
  d2m = (sm[0] == 0 || sm[1] == 0 || sm[2] == 0) ? c0 : c1;
}

// :361-434
void h_edges_nosc(int ijdm, int ijs, int ije, P1 hevc1, P1 hevc2, P1 hevc3, P1 hevc4,
                  P1 ssc, P1 scc, P1 d2m, P1 hm, P1 hel, P1 her) {
  const int nb = hm.nb;
  std::vector<double> d2hv(ijdm + 2 * nb);
  P1 d2h{d2hv.data(), nb};
  double he, sl, sr, sc, d, q, r, a2;
  for (int i = ijs - 1; i <= ije + 2; ++i) {
    he = hevc1(i) * hm(i - 2) + hevc2(i) * hm(i - 1) + hevc3(i) * hm(i) + hevc4(i) * hm(i + 1);
    hel(i) = he;
    her(i - 1) = he;
  }
  for (int i = ijs - 1; i <= ije + 1; ++i) d2h(i) = d2m(i) * (hel(i) - c2 * hm(i) + her(i));
  for (int i = ijs; i <= ije; ++i) {
    if (d2h(i - 1) * d2h(i) <= c0 || d2h(i) * d2h(i + 1) <= c0) {
      sl = ssc(i) * (hm(i) - hm(i - 1));
      sr = ssc(i) * (hm(i + 1) - hm(i));
      if (sl * sr > c0) {
        sc = scc(i) * (hm(i + 1) - hm(i - 1));
        sc = fsign(std::min(std::min(std::fabs(sl), std::fabs(sr)), std::fabs(sc)), sc);
        if ((hm(i - 1) - hel(i)) * (hm(i) - hel(i)) > c0)
          hel(i) = hm(i) - fsign(std::min(c1_2 * std::fabs(sc), std::fabs(hel(i) - hm(i))), sc);
        if ((hm(i + 1) - her(i)) * (hm(i) - her(i)) > c0)
          her(i) = hm(i) + fsign(std::min(c1_2 * std::fabs(sc), std::fabs(her(i) - hm(i))), sc);
        d = her(i) - hel(i);
        q = d * (c2 * hm(i) - hel(i) - her(i));
        r = c1_3 * d * d;
        if (q > r) hel(i) = c3 * hm(i) - c2 * her(i);
        else if (-r > q) her(i) = c3 * hm(i) - c2 * hel(i);
      } else {
        hel(i) = hm(i);
        her(i) = hm(i);
      }
    }
    hel(i) = std::max(hel(i), dpeps);
    her(i) = std::max(her(i), dpeps);
    sl = c2 * (c3 * hm(i) - c2 * hel(i) - her(i));
    a2 = c3 * (hel(i) - c2 * hm(i) + her(i));
    sr = sl + c2 * a2;
    if (sl < c0 && sr > c0) {
      if (a2 * hel(i) - c1_4 * sl * sl < a2 * dpeps) {
        q = c3 * hm(i) / (c3 * sl * sr + c4 * a2 * a2);
        hel(i) = sl * sl * q;
        her(i) = sr * sr * q;
      }
    }
  }
}

// compatible tracer edge-value weights of interface i: the per-interface LU solve that
// parabola_coeffs_fc_nosc (:519-722) and parabola_coeffs_fc_mono (:849-1052) both spell out.
// tevc1..4 are in/out: a Fortran select without a matching case keeps the previous values.
inline void compat_edge_weights(int i, PI1 stencil, PC tmc0, PC tmcl, PC tmcr, P1 hm, P1 hel, P1 her,
                                double& tevc1, double& tevc2, double& tevc3, double& tevc4) {
  double h1i, h2i, h3i, h4i, a12, a22, a32, a42, a13, a23, a33, a43, a14, a24, a34, a44, q;
    switch (stencil(i)) {
      case stencil_1111:
        h1i = c1 / hm(i - 2); h2i = c1 / hm(i - 1); h3i = c1 / hm(i); h4i = c1 / hm(i + 1);
        a12 = tmc0(1, i) + (tmcl(1, i) * hel(i - 2) + tmcr(1, i) * her(i - 2)) * h1i;
        a13 = tmc0(2, i) + (tmcl(2, i) * hel(i - 2) + tmcr(2, i) * her(i - 2)) * h1i;
        a14 = tmc0(3, i) + (tmcl(3, i) * hel(i - 2) + tmcr(3, i) * her(i - 2)) * h1i;
        a22 = tmc0(4, i) + (tmcl(4, i) * hel(i - 1) + tmcr(4, i) * her(i - 1)) * h2i - a12;
        a23 = tmc0(5, i) + (tmcl(5, i) * hel(i - 1) + tmcr(5, i) * her(i - 1)) * h2i - a13;
        a24 = tmc0(6, i) + (tmcl(6, i) * hel(i - 1) + tmcr(6, i) * her(i - 1)) * h2i - a14;
        a32 = tmc0(7, i) + (tmcl(7, i) * hel(i) + tmcr(7, i) * her(i)) * h3i - a12;
        a33 = tmc0(8, i) + (tmcl(8, i) * hel(i) + tmcr(8, i) * her(i)) * h3i - a13;
        a34 = tmc0(9, i) + (tmcl(9, i) * hel(i) + tmcr(9, i) * her(i)) * h3i - a14;
        a42 = tmc0(10, i) + (tmcl(10, i) * hel(i + 1) + tmcr(10, i) * her(i + 1)) * h4i - a12;
        a43 = tmc0(11, i) + (tmcl(11, i) * hel(i + 1) + tmcr(11, i) * her(i + 1)) * h4i - a13;
        a44 = tmc0(12, i) + (tmcl(12, i) * hel(i + 1) + tmcr(12, i) * her(i + 1)) * h4i - a14;
        q = c1 / a22;
        a23 = a23 * q;
        a33 = a33 - a23 * a32;
        a43 = a43 - a23 * a42;
        a24 = a24 * q;
        a34 = a34 - a24 * a32;
        a44 = a44 - a24 * a42;
        a34 = a34 / a33;
        a44 = a44 - a34 * a43;
        tevc2 = -a12;
        tevc3 = -a13 - a23 * tevc2;
        tevc4 = -a14 - a24 * tevc2 - a34 * tevc3;
        tevc4 = tevc4 / a44;
        tevc3 = (tevc3 - a43 * tevc4) / a33;
        tevc2 = (tevc2 - a32 * tevc3 - a42 * tevc4) / a22;
        tevc1 = c1 - tevc2 - tevc3 - tevc4;
        break;
      case stencil_0000:
        tevc1 = c0; tevc2 = c0; tevc3 = c0; tevc4 = c0;
        break;
      case stencil_1110:
        h1i = c1 / hm(i - 2); h2i = c1 / hm(i - 1); h3i = c1 / hm(i);
        a12 = tmc0(1, i) + (tmcl(1, i) * hel(i - 2) + tmcr(1, i) * her(i - 2)) * h1i;
        a13 = tmc0(2, i) + (tmcl(2, i) * hel(i - 2) + tmcr(2, i) * her(i - 2)) * h1i;
        a22 = tmc0(4, i) + (tmcl(4, i) * hel(i - 1) + tmcr(4, i) * her(i - 1)) * h2i - a12;
        a23 = tmc0(5, i) + (tmcl(5, i) * hel(i - 1) + tmcr(5, i) * her(i - 1)) * h2i - a13;
        a32 = tmc0(7, i) + (tmcl(7, i) * hel(i) + tmcr(7, i) * her(i)) * h3i - a12;
        a33 = tmc0(8, i) + (tmcl(8, i) * hel(i) + tmcr(8, i) * her(i)) * h3i - a13;
        a23 = a23 / a22;
        a33 = a33 - a23 * a32;
        tevc2 = -a12;
        tevc3 = -a13 - a23 * tevc2;
        tevc3 = tevc3 / a33;
        tevc2 = (tevc2 - a32 * tevc3) / a22;
        tevc1 = c1 - tevc2 - tevc3;
        tevc4 = c0;
        break;
      case stencil_0111:
        h2i = c1 / hm(i - 1); h3i = c1 / hm(i); h4i = c1 / hm(i + 1);
        a22 = tmc0(4, i) + (tmcl(4, i) * hel(i - 1) + tmcr(4, i) * her(i - 1)) * h2i;
        a23 = tmc0(5, i) + (tmcl(5, i) * hel(i - 1) + tmcr(5, i) * her(i - 1)) * h2i;
        a32 = tmc0(7, i) + (tmcl(7, i) * hel(i) + tmcr(7, i) * her(i)) * h3i - a22;
        a33 = tmc0(8, i) + (tmcl(8, i) * hel(i) + tmcr(8, i) * her(i)) * h3i - a23;
        a42 = tmc0(10, i) + (tmcl(10, i) * hel(i + 1) + tmcr(10, i) * her(i + 1)) * h4i - a22;
        a43 = tmc0(11, i) + (tmcl(11, i) * hel(i + 1) + tmcr(11, i) * her(i + 1)) * h4i - a23;
        a33 = a33 / a32;
        a43 = a43 - a33 * a42;
        tevc3 = -a22;
        tevc4 = -a23 - a33 * tevc3;
        tevc4 = tevc4 / a43;
        tevc3 = (tevc3 - a42 * tevc4) / a32;
        tevc2 = c1 - tevc3 - tevc4;
        tevc1 = c0;
        break;
      case stencil_1100:
        h1i = c1 / hm(i - 2); h2i = c1 / hm(i - 1);
        a12 = tmc0(1, i) + (tmcl(1, i) * hel(i - 2) + tmcr(1, i) * her(i - 2)) * h1i;
        a22 = tmc0(4, i) + (tmcl(4, i) * hel(i - 1) + tmcr(4, i) * her(i - 1)) * h2i - a12;
        tevc2 = -a12 / a22;
        tevc1 = c1 - tevc2;
        tevc3 = c0; tevc4 = c0;
        break;
      case stencil_0110:
        h2i = c1 / hm(i - 1); h3i = c1 / hm(i);
        a22 = tmc0(4, i) + (tmcl(4, i) * hel(i - 1) + tmcr(4, i) * her(i - 1)) * h2i;
        a32 = tmc0(7, i) + (tmcl(7, i) * hel(i) + tmcr(7, i) * her(i)) * h3i - a22;
        tevc3 = -a22 / a32;
        tevc2 = c1 - tevc3;
        tevc1 = c0; tevc4 = c0;
        break;
      case stencil_0011:
        h3i = c1 / hm(i); h4i = c1 / hm(i + 1);
        a32 = tmc0(7, i) + (tmcl(7, i) * hel(i) + tmcr(7, i) * her(i)) * h3i;
        a42 = tmc0(10, i) + (tmcl(10, i) * hel(i + 1) + tmcr(10, i) * her(i + 1)) * h4i - a32;
        tevc4 = -a32 / a42;
        tevc3 = c1 - tevc4;
        tevc1 = c0; tevc2 = c0;
        break;
      case stencil_0100:
        tevc1 = c0; tevc2 = c1; tevc3 = c0; tevc4 = c0;
        break;
      case stencil_0010:
        tevc1 = c0; tevc2 = c0; tevc3 = c1; tevc4 = c0;
        break;
      default:
        break;  // Fortran select with no matching case: coefficients keep previous values
    }
}

// :490-818
void parabola_coeffs_fc_nosc(int ijdm, int ijs, int ije, PI1 stencil, PC tmc0, PC tmcl, PC tmcr,
                             P1 ssc, P1 scc, P1 d2m, P1 hm, P2 tm, P1 hel, P1 her,
                             P1 hpc0, P1 hpc1, P1 hpc2, P2 tpc0, P2 tpc1, P2 tpc2) {
  const int nb = hm.nb, ntl = tm.nt, n1 = ijdm + 2 * nb;
  std::vector<double> w((size_t)n1 * (3 * ntl + 6), 0.0);
  P2 d2t{w.data(), nb, ntl}, tel{w.data() + (size_t)n1 * ntl, nb, ntl},
     ter{w.data() + (size_t)2 * n1 * ntl, nb, ntl};
  double* b = w.data() + (size_t)3 * n1 * ntl;
  P1 hf1m{b, nb}, hf1l{b + n1, nb}, hf1r{b + 2 * n1, nb}, hf2m{b + 3 * n1, nb},
     hf2l{b + 4 * n1, nb}, hf2r{b + 5 * n1, nb};
  double q, tevc1 = 0, tevc2 = 0, tevc3 = 0, tevc4 = 0, te, sl, sr, sc, a2;

  for (int i = ijs - 1; i <= ije + 2; ++i) {
    compat_edge_weights(i, stencil, tmc0, tmcl, tmcr, hm, hel, her, tevc1, tevc2, tevc3, tevc4);
    for (int nt = 1; nt <= ntl; ++nt) {
      te = tevc1 * tm(nt, i - 2) + tevc2 * tm(nt, i - 1) + tevc3 * tm(nt, i) + tevc4 * tm(nt, i + 1);
      tel(nt, i) = te;
      ter(nt, i - 1) = te;
    }
  }

  for (int i = ijs - 1; i <= ije + 1; ++i) {
    q = c1 / (c12 * hm(i) - hel(i) - her(i));
    hf1m(i) = c60 * hm(i) * q;
    hf1l(i) = -(c42 * hm(i) + c4 * hel(i) - c6 * her(i)) * q;
    hf1r(i) = -(c18 * hm(i) - c4 * hel(i) + c6 * her(i)) * q;
    hf2m(i) = -hf1m(i);
    hf2l(i) = c5 * (c6 * hm(i) + hel(i) - her(i)) * q;
    hf2r(i) = c5 * (c6 * hm(i) - hel(i) + her(i)) * q;
    for (int nt = 1; nt <= ntl; ++nt)
      d2t(nt, i) = d2m(i) * (hf2m(i) * tm(nt, i) + hf2l(i) * tel(nt, i) + hf2r(i) * ter(nt, i));
  }

  for (int i = ijs; i <= ije; ++i) {
    for (int nt = 1; nt <= ntl; ++nt) {
      if (d2t(nt, i - 1) * d2t(nt, i) <= c0 || d2t(nt, i) * d2t(nt, i + 1) <= c0) {
        sl = ssc(i) * (tm(nt, i) - tm(nt, i - 1));
        sr = ssc(i) * (tm(nt, i + 1) - tm(nt, i));
        if (sl * sr > c0) {
          sc = scc(i) * (tm(nt, i + 1) - tm(nt, i - 1));
          sc = fsign(std::min(std::min(std::fabs(sl), std::fabs(sr)), std::fabs(sc)), sc);
          if ((tm(nt, i - 1) - tel(nt, i)) * (tm(nt, i) - tel(nt, i)) > c0)
            tel(nt, i) = tm(nt, i) -
                         fsign(std::min(c1_2 * std::fabs(sc), std::fabs(tel(nt, i) - tm(nt, i))), sc);
          if ((tm(nt, i + 1) - ter(nt, i)) * (tm(nt, i) - ter(nt, i)) > c0)
            ter(nt, i) = tm(nt, i) +
                         fsign(std::min(c1_2 * std::fabs(sc), std::fabs(ter(nt, i) - tm(nt, i))), sc);
          sl = hf1m(i) * tm(nt, i) + hf1l(i) * tel(nt, i) + hf1r(i) * ter(nt, i);
          a2 = hf2m(i) * tm(nt, i) + hf2l(i) * tel(nt, i) + hf2r(i) * ter(nt, i);
          sr = sl + c2 * a2;
          if (sl * sr < c0) {
            if ((ter(nt, i) - tel(nt, i)) * a2 < c0) {
              tel(nt, i) = -((hf1m(i) + c2 * hf2m(i)) * tm(nt, i) + (hf1r(i) + c2 * hf2r(i)) * ter(nt, i)) /
                           (hf1l(i) + c2 * hf2l(i));
            } else {
              ter(nt, i) = -(hf1m(i) * tm(nt, i) + hf1l(i) * tel(nt, i)) / hf1r(i);
            }
          }
        } else {
          tel(nt, i) = tm(nt, i);
          ter(nt, i) = tm(nt, i);
        }
      }
    }
    for (int nt = 2; nt <= ntl; ++nt) {
      tel(nt, i) = std::max(tel(nt, i), c0);
      ter(nt, i) = std::max(ter(nt, i), c0);
      sl = hf1m(i) * tm(nt, i) + hf1l(i) * tel(nt, i) + hf1r(i) * ter(nt, i);
      a2 = hf2m(i) * tm(nt, i) + hf2l(i) * tel(nt, i) + hf2r(i) * ter(nt, i);
      sr = sl + c2 * a2;
      if (sl < c0 && sr > c0) {
        if (a2 * tel(nt, i) - c1_4 * sl * sl < c0) {
          q = c3 * tm(nt, i) / (c3 * sl * sr + c4 * a2 * a2);
          tel(nt, i) = sl * sl * q;
          ter(nt, i) = sr * sr * q;
        }
      }
    }
    hpc0(i) = hel(i);
    hpc1(i) = c6 * hm(i) - c4 * hel(i) - c2 * her(i);
    hpc2(i) = c3 * (hel(i) - c2 * hm(i) + her(i));
    for (int nt = 1; nt <= ntl; ++nt) {
      tpc0(nt, i) = tel(nt, i);
      tpc1(nt, i) = hf1m(i) * tm(nt, i) + hf1l(i) * tel(nt, i) + hf1r(i) * ter(nt, i);
      tpc2(nt, i) = hf2m(i) * tm(nt, i) + hf2l(i) * tel(nt, i) + hf2r(i) * ter(nt, i);
    }
  }
}

// :436-488
void h_edges_mono(int ijdm, int ijs, int ije, P1 hevc1, P1 hevc2, P1 hevc3, P1 hevc4, P1 ssc, P1 scc,
                  P1 hm, P1 hel, P1 her) {
  (void)ijdm;
  double he, sl, sr, sc, d, q, r;
  for (int i = ijs; i <= ije + 1; ++i) {
    he = hevc1(i) * hm(i - 2) + hevc2(i) * hm(i - 1) + hevc3(i) * hm(i) + hevc4(i) * hm(i + 1);
    hel(i) = he;
    her(i - 1) = he;
  }
  for (int i = ijs; i <= ije; ++i) {
    sl = ssc(i) * (hm(i) - hm(i - 1));
    sr = ssc(i) * (hm(i + 1) - hm(i));
    if (sl * sr > c0) {
      sc = scc(i) * (hm(i + 1) - hm(i - 1));
      sc = fsign(std::min(std::min(std::fabs(sl), std::fabs(sr)), std::fabs(sc)), sc);
      if ((hm(i - 1) - hel(i)) * (hm(i) - hel(i)) > c0)
        hel(i) = hm(i) - fsign(std::min(c1_2 * std::fabs(sc), std::fabs(hel(i) - hm(i))), sc);
      if ((hm(i + 1) - her(i)) * (hm(i) - her(i)) > c0)
        her(i) = hm(i) + fsign(std::min(c1_2 * std::fabs(sc), std::fabs(her(i) - hm(i))), sc);
      d = her(i) - hel(i);
      q = d * (c2 * hm(i) - hel(i) - her(i));
      r = c1_3 * d * d;
      if (q > r) hel(i) = c3 * hm(i) - c2 * her(i);
      else if (-r > q) her(i) = c3 * hm(i) - c2 * hel(i);
    } else {
      hel(i) = hm(i);
      her(i) = hm(i);
    }
  }
}

// :820-1116
void parabola_coeffs_fc_mono(int ijdm, int ijs, int ije, PI1 stencil, PC tmc0, PC tmcl, PC tmcr,
                             P1 ssc, P1 scc, P1 hm, P2 tm, P1 hel, P1 her,
                             P1 hpc0, P1 hpc1, P1 hpc2, P2 tpc0, P2 tpc1, P2 tpc2) {
  const int nb = hm.nb, ntl = tm.nt, n1 = ijdm + 2 * nb;
  std::vector<double> w((size_t)n1 * 2 * ntl, 0.0);
  P2 tel{w.data(), nb, ntl}, ter{w.data() + (size_t)n1 * ntl, nb, ntl};
  double q, tevc1 = 0, tevc2 = 0, tevc3 = 0, tevc4 = 0, te, hf1m, hf1l, hf1r, hf2m, hf2l, hf2r, sl, sr,
         sc, a2;
  for (int i = ijs; i <= ije + 1; ++i) {
    compat_edge_weights(i, stencil, tmc0, tmcl, tmcr, hm, hel, her, tevc1, tevc2, tevc3, tevc4);
    for (int nt = 1; nt <= ntl; ++nt) {
      te = tevc1 * tm(nt, i - 2) + tevc2 * tm(nt, i - 1) + tevc3 * tm(nt, i) + tevc4 * tm(nt, i + 1);
      tel(nt, i) = te;
      ter(nt, i - 1) = te;
    }
  }
  for (int i = ijs; i <= ije; ++i) {
    q = c1 / (c12 * hm(i) - hel(i) - her(i));
    hf1m = c60 * hm(i) * q;
    hf1l = -(c42 * hm(i) + c4 * hel(i) - c6 * her(i)) * q;
    hf1r = -(c18 * hm(i) - c4 * hel(i) + c6 * her(i)) * q;
    hf2m = -hf1m;
    hf2l = c5 * (c6 * hm(i) + hel(i) - her(i)) * q;
    hf2r = c5 * (c6 * hm(i) - hel(i) + her(i)) * q;
    for (int nt = 1; nt <= ntl; ++nt) {
      sl = ssc(i) * (tm(nt, i) - tm(nt, i - 1));
      sr = ssc(i) * (tm(nt, i + 1) - tm(nt, i));
      if (sl * sr > c0) {
        sc = scc(i) * (tm(nt, i + 1) - tm(nt, i - 1));
        sc = fsign(std::min(std::min(std::fabs(sl), std::fabs(sr)), std::fabs(sc)), sc);
        if ((tm(nt, i - 1) - tel(nt, i)) * (tm(nt, i) - tel(nt, i)) > c0)
          tel(nt, i) = tm(nt, i) - fsign(std::min(c1_2 * std::fabs(sc), std::fabs(tel(nt, i) - tm(nt, i))), sc);
        if ((tm(nt, i + 1) - ter(nt, i)) * (tm(nt, i) - ter(nt, i)) > c0)
          ter(nt, i) = tm(nt, i) + fsign(std::min(c1_2 * std::fabs(sc), std::fabs(ter(nt, i) - tm(nt, i))), sc);
        sl = hf1m * tm(nt, i) + hf1l * tel(nt, i) + hf1r * ter(nt, i);
        a2 = hf2m * tm(nt, i) + hf2l * tel(nt, i) + hf2r * ter(nt, i);
        sr = sl + c2 * a2;
        if (sl * sr < c0) {
          if ((ter(nt, i) - tel(nt, i)) * a2 < c0)
            tel(nt, i) = -((hf1m + c2 * hf2m) * tm(nt, i) + (hf1r + c2 * hf2r) * ter(nt, i)) / (hf1l + c2 * hf2l);
          else
            ter(nt, i) = -(hf1m * tm(nt, i) + hf1l * tel(nt, i)) / hf1r;
        }
      } else {
        tel(nt, i) = tm(nt, i);
        ter(nt, i) = tm(nt, i);
      }
    }
    hpc0(i) = hel(i);
    hpc1(i) = c6 * hm(i) - c4 * hel(i) - c2 * her(i);
    hpc2(i) = c3 * (hel(i) - c2 * hm(i) + her(i));
    for (int nt = 1; nt <= ntl; ++nt) {
      tpc0(nt, i) = tel(nt, i);
      tpc1(nt, i) = hf1m * tm(nt, i) + hf1l * tel(nt, i) + hf1r * ter(nt, i);
      tpc2(nt, i) = hf2m * tm(nt, i) + hf2l * tel(nt, i) + hf2r * ter(nt, i);
    }
  }
}

// slope limiter + parabola monotonicity fix shared verbatim by the thickness and the tracer
// branches of the partial-compatibility routines (:1168-1190, :1209-1232, :1309-1331, :1334-1358)
inline void pc_limit(double ssc, double scc, double xm, double x0, double xp, double& el, double& er) {
  double sl = ssc * (x0 - xm), sr = ssc * (xp - x0);
  if (sl * sr > c0) {
    double sc = scc * (xp - xm);
    sc = fsign(std::min(std::min(std::fabs(sl), std::fabs(sr)), std::fabs(sc)), sc);
    if ((xm - el) * (x0 - el) > c0) el = x0 - fsign(std::min(c1_2 * std::fabs(sc), std::fabs(el - x0)), sc);
    if ((xp - er) * (x0 - er) > c0) er = x0 + fsign(std::min(c1_2 * std::fabs(sc), std::fabs(er - x0)), sc);
    double d = er - el;
    double q = d * (c2 * x0 - el - er);
    double r = c1_3 * d * d;
    if (q > r) el = c3 * x0 - c2 * er;
    else if (-r > q) er = c3 * x0 - c2 * el;
  } else {
    el = x0;
    er = x0;
  }
}

// :1118-1264
void parabola_coeffs_pc_nosc(int ijdm, int ijs, int ije, P1 hevc1, P1 hevc2, P1 hevc3, P1 hevc4,
                             P1 ssc, P1 scc, P1 d2m, P1 hm, P2 tm,
                             P1 hpc0, P1 hpc1, P1 hpc2, P2 tpc0, P2 tpc1, P2 tpc2) {
  const int nb = hm.nb, ntl = tm.nt, n1 = ijdm + 2 * nb;
  std::vector<double> w((size_t)n1 * (3 * ntl + 3), 0.0);
  P2 d2t{w.data(), nb, ntl}, tel{w.data() + (size_t)n1 * ntl, nb, ntl},
     ter{w.data() + (size_t)2 * n1 * ntl, nb, ntl};
  double* b = w.data() + (size_t)3 * n1 * ntl;
  P1 hel{b, nb}, her{b + n1, nb}, d2h{b + 2 * n1, nb};
  double he, te, sl, sr, q, a2;
  for (int i = ijs - 1; i <= ije + 2; ++i) {
    he = hevc1(i) * hm(i - 2) + hevc2(i) * hm(i - 1) + hevc3(i) * hm(i) + hevc4(i) * hm(i + 1);
    hel(i) = he;
    her(i - 1) = he;
    for (int nt = 1; nt <= ntl; ++nt) {
      te = hevc1(i) * tm(nt, i - 2) + hevc2(i) * tm(nt, i - 1) + hevc3(i) * tm(nt, i) + hevc4(i) * tm(nt, i + 1);
      tel(nt, i) = te;
      ter(nt, i - 1) = te;
    }
  }
  for (int i = ijs - 1; i <= ije + 1; ++i) {
    d2h(i) = d2m(i) * (hel(i) - c2 * hm(i) + her(i));
    for (int nt = 1; nt <= ntl; ++nt) d2t(nt, i) = d2m(i) * (tel(nt, i) - c2 * tm(nt, i) + ter(nt, i));
  }
  for (int i = ijs; i <= ije; ++i) {
    if (d2h(i - 1) * d2h(i) <= c0 || d2h(i) * d2h(i + 1) <= c0)
      pc_limit(ssc(i), scc(i), hm(i - 1), hm(i), hm(i + 1), hel(i), her(i));
    hel(i) = std::max(hel(i), dpeps);
    her(i) = std::max(her(i), dpeps);
    sl = c2 * (c3 * hm(i) - c2 * hel(i) - her(i));
    a2 = c3 * (hel(i) - c2 * hm(i) + her(i));
    sr = sl + c2 * a2;
    if (sl < c0 && sr > c0) {
      if (a2 * hel(i) - c1_4 * sl * sl < a2 * dpeps) {
        q = c3 * hm(i) / (c3 * sl * sr + c4 * a2 * a2);
        hel(i) = sl * sl * q;
        her(i) = sr * sr * q;
      }
    }
    for (int nt = 1; nt <= ntl; ++nt)
      if (d2t(nt, i - 1) * d2t(nt, i) <= c0 || d2t(nt, i) * d2t(nt, i + 1) <= c0)
        pc_limit(ssc(i), scc(i), tm(nt, i - 1), tm(nt, i), tm(nt, i + 1), tel(nt, i), ter(nt, i));
    for (int nt = 2; nt <= ntl; ++nt) {
      tel(nt, i) = std::max(tel(nt, i), c0);
      ter(nt, i) = std::max(ter(nt, i), c0);
      sl = c2 * (c3 * tm(nt, i) - c2 * tel(nt, i) - ter(nt, i));
      a2 = c3 * (tel(nt, i) - c2 * tm(nt, i) + ter(nt, i));
      sr = sl + c2 * a2;
      if (sl < c0 && sr > c0) {
        if (a2 * tel(nt, i) - c1_4 * sl * sl < c0) {
          q = c3 * tm(nt, i) / (c3 * sl * sr + c4 * a2 * a2);
          tel(nt, i) = sl * sl * q;
          ter(nt, i) = sr * sr * q;
        }
      }
    }
    hpc0(i) = hel(i);
    hpc1(i) = c6 * hm(i) - c4 * hel(i) - c2 * her(i);
    hpc2(i) = c3 * (hel(i) - c2 * hm(i) + her(i));
    for (int nt = 1; nt <= ntl; ++nt) {
      tpc0(nt, i) = tel(nt, i);
      tpc1(nt, i) = c6 * tm(nt, i) - c4 * tel(nt, i) - c2 * ter(nt, i);
      tpc2(nt, i) = c3 * (tel(nt, i) - c2 * tm(nt, i) + ter(nt, i));
    }
  }
}

// :1266-1371
void parabola_coeffs_pc_mono(int ijdm, int ijs, int ije, P1 hevc1, P1 hevc2, P1 hevc3, P1 hevc4,
                             P1 ssc, P1 scc, P1 hm, P2 tm,
                             P1 hpc0, P1 hpc1, P1 hpc2, P2 tpc0, P2 tpc1, P2 tpc2) {
  const int nb = hm.nb, ntl = tm.nt, n1 = ijdm + 2 * nb;
  std::vector<double> w((size_t)n1 * (2 * ntl + 2), 0.0);
  P2 tel{w.data(), nb, ntl}, ter{w.data() + (size_t)n1 * ntl, nb, ntl};
  double* b = w.data() + (size_t)2 * n1 * ntl;
  P1 hel{b, nb}, her{b + n1, nb};
  double he, te;
  for (int i = ijs; i <= ije + 1; ++i) {
    he = hevc1(i) * hm(i - 2) + hevc2(i) * hm(i - 1) + hevc3(i) * hm(i) + hevc4(i) * hm(i + 1);
    hel(i) = he;
    her(i - 1) = he;
    for (int nt = 1; nt <= ntl; ++nt) {
      te = hevc1(i) * tm(nt, i - 2) + hevc2(i) * tm(nt, i - 1) + hevc3(i) * tm(nt, i) + hevc4(i) * tm(nt, i + 1);
      tel(nt, i) = te;
      ter(nt, i - 1) = te;
    }
  }
  for (int i = ijs; i <= ije; ++i) {
    pc_limit(ssc(i), scc(i), hm(i - 1), hm(i), hm(i + 1), hel(i), her(i));
    for (int nt = 1; nt <= ntl; ++nt)
      pc_limit(ssc(i), scc(i), tm(nt, i - 1), tm(nt, i), tm(nt, i + 1), tel(nt, i), ter(nt, i));
    hpc0(i) = hel(i);
    hpc1(i) = c6 * hm(i) - c4 * hel(i) - c2 * her(i);
    hpc2(i) = c3 * (hel(i) - c2 * hm(i) + her(i));
    for (int nt = 1; nt <= ntl; ++nt) {
      tpc0(nt, i) = tel(nt, i);
      tpc1(nt, i) = c6 * tm(nt, i) - c4 * tel(nt, i) - c2 * ter(nt, i);
      tpc2(nt, i) = c3 * (tel(nt, i) - c2 * tm(nt, i) + ter(nt, i));
    }
  }
}

// :1373-1468
void flux_integration(int ijs, int ije, P1 ca, P1 ai, P1 db, P1 du, P1 dl, P1 hpc0, P1 hpc1,
                      P1 hpc2, P2 tpc0, P2 tpc1, P2 tpc2, P1 hf, P2 htf) {
  const int ntl = tpc0.nt;
  double c, hb, p0, p1, p2, q1, q2, q3, q4;
  for (int i = ijs; i <= ije; ++i) {
    if (ca(i) < c0) {
      c = ca(i) * ai(i);
      if (dl(i) > db(i)) {
        hb = std::max(c0, db(i) - du(i));
        hf(i) = hb * ca(i);
        p0 = hb;
        p1 = -c1_2 * hb * c;
        p2 = c1_3 * hb * c * c;
      } else {
        hf(i) = (hpc0(i) - (c1_2 * hpc1(i) - c1_3 * hpc2(i) * c) * c) * ca(i);
        p0 = hpc0(i) - (c1_2 * hpc1(i) - c1_3 * hpc2(i) * c) * c;
        p1 = -(c1_2 * hpc0(i) - (c1_3 * hpc1(i) - c1_4 * hpc2(i) * c) * c) * c;
        p2 = (c1_3 * hpc0(i) - (c1_4 * hpc1(i) - c1_5 * hpc2(i) * c) * c) * c * c;
      }
      for (int nt = 1; nt <= ntl; ++nt)
        htf(nt, i) = (p0 * tpc0(nt, i) + p1 * tpc1(nt, i) + p2 * tpc2(nt, i)) * ca(i);
    } else {
      c = ca(i) * ai(i - 1);
      q1 = c1 - c1_2 * c;
      q2 = c1 - (c1 - c1_3 * c) * c;
      if (dl(i - 1) > db(i)) {
        hb = std::max(c0, db(i) - du(i - 1));
        hf(i) = hb * ca(i);
        p0 = hb;
        p1 = q1 * hb;
        p2 = q2 * hb;
      } else {
        hf(i) = (hpc0(i - 1) + q1 * hpc1(i - 1) + q2 * hpc2(i - 1)) * ca(i);
        q3 = c1_4 * (c1 + c3 * (c1 - c) * q2);
        q4 = c1_5 * (c1 + c4 * (c1 - c) * q3);
        p0 = hpc0(i - 1) + q1 * hpc1(i - 1) + q2 * hpc2(i - 1);
        p1 = q1 * hpc0(i - 1) + q2 * hpc1(i - 1) + q3 * hpc2(i - 1);
        p2 = q2 * hpc0(i - 1) + q3 * hpc1(i - 1) + q4 * hpc2(i - 1);
      }
      for (int nt = 1; nt <= ntl; ++nt)
        htf(nt, i) = (p0 * tpc0(nt, i - 1) + p1 * tpc1(nt, i - 1) + p2 * tpc2(nt, i - 1)) * ca(i);
    }
  }
}

struct Pencils {
  int n1, nb, ntl;
  std::vector<double> buf;
  P1 db, dl, du, ca, ai, ho, hm, hel, her, hpc0, hpc1, hpc2, hf;
  P2 tm, tpc0, tpc1, tpc2, htf;
  Pencils(int ijdm, int nb_, int ntl_) : n1(ijdm + 2 * nb_), nb(nb_), ntl(ntl_) {
    buf.assign((size_t)n1 * (13 + 5 * ntl), 0.0);
    double* b = buf.data();
    P1* s[] = {&db, &dl, &du, &ca, &ai, &ho, &hm, &hel, &her, &hpc0, &hpc1, &hpc2, &hf};
    for (P1* q : s) { *q = P1{b, nb}; b += n1; }
    P2* t[] = {&tm, &tpc0, &tpc1, &tpc2, &htf};
    for (P2* q : t) { *q = P2{b, nb, ntl}; b += (size_t)n1 * ntl; }
  }
};

// tracer accessor: nt=1 temp, 2 saln, >=3 trc(:,:,:,nt-2)
struct Scalars {
  A3 temp, saln, trc; int kdm2;
  inline double& at(int nt, int i, int j, int kn) const {
    if (nt == 1) return temp(i, j, kn);
    if (nt == 2) return saln(i, j, kn);
    return trc(i, j, kn + (nt - 3) * kdm2);
  }
};

// :1470-1623 (fc_nosc), :1787-1939 (fc_mono), :2102-2200 (pc_nosc), :2302-2399 (pc_mono): the four
// i-direction drivers differ only in the halo width (4 nosc / 3 mono), the pencil range that follows
// from it, whether the thickness edges are staged through hel_3d/her_3d (full compatibility) and
// which reconstruction routine they call.
void cppm_pass_i(int m, int n, int mm, int nn, int k1m, int k1n, bool second_pass, bool full, bool mono) {
  (void)m; (void)k1m;
  Oracle& o = O(); const Dims& d = o.d;
  const int idm = d.idm, jdm = d.jdm, kdm = d.kdm, nb = d.nbdy, ii = d.ii, jj = d.jj;
  const int ntl = T.ntr_loc;
  A3 dp = o.a3("dp"), temp = o.a3("temp"), saln = o.a3("saln");
  A3 trc = d.ntr > 0 ? o.a3("trc") : A3{};
  A3 uflx = o.a3("uflx"), utflx = o.a3("utflx"), usflx = o.a3("usflx");
  A3 p = o.a3("p"), cau = o.a3("cau"), cav = o.a3("cav"), pbu = o.a3("pbu");
  A2 scp2i = o.a2("scp2i");
  A3 hel_3d{T.hel_3d.data(), d.ldi, nb, d.lev}, her_3d{T.her_3d.data(), d.ldi, nb, d.lev};
  Scalars S{temp, saln, trc, 2 * kdm};
  Pencils P(idm, nb, ntl);
  auto row = [&](std::vector<double>& v, int j) { return P1{v.data() + (size_t)(j + nb - 1) * d.ldi, nb}; };

  const int hw = mono ? 3 : 4;          // halo width
  const int lo = mono ? -2 : -3, hi = idm + (mono ? 3 : 4);
  xctilr(dp.from(k1n), 1, kdm, hw, 0, halo_ps);
  xctilr(temp.from(k1n), 1, kdm, hw, 0, halo_ps);
  xctilr(saln.from(k1n), 1, kdm, hw, 0, halo_ps);
  for (int nt = 3; nt <= ntl; ++nt) xctilr(trc.from(k1n + (nt - 3) * 2 * kdm), 1, kdm, hw, 0, halo_ps);

  if (full) {
  for (int k = 1; k <= kdm; ++k) {
    const int kn = k + nn;
    for (int j = 1; j <= jdm; ++j) {
      for (int i = -2; i <= idm + 3; ++i) {
        P.ai(i) = scp2i(i, j);
        P.hm(i) = std::max(c0, dp(i, j, kn)) + dpeps;
      }
      if (second_pass)
        for (int i = -2; i <= idm + 3; ++i)
          P.hm(i) = P.hm(i) / (c1 - (cav(i, j + 1, k) - cav(i, j, k)) * P.ai(i));
      if (mono)
        h_edges_mono(idm, 1, idm, row(T.hevc1i, j), row(T.hevc2i, j), row(T.hevc3i, j),
                     row(T.hevc4i, j), row(T.ssci, j), row(T.scci, j), P.hm, P.hel, P.her);
      else
        h_edges_nosc(idm, 1, idm, row(T.hevc1i, j), row(T.hevc2i, j), row(T.hevc3i, j),
                     row(T.hevc4i, j), row(T.ssci, j), row(T.scci, j), row(T.d2mi, j), P.hm, P.hel,
                     P.her);
      for (int i = 1; i <= idm; ++i) { hel_3d(i, j, k) = P.hel(i); her_3d(i, j, k) = P.her(i); }
    }
  }
  xctilr(hel_3d, 1, kdm, hw, 0, halo_ps);
  xctilr(her_3d, 1, kdm, hw, 0, halo_ps);
  if (d.nreg == 2) {  // :1532 / :1848 (single tile: nproc == jpr)
    const int j = jj;
    for (int k = 1; k <= kdm; ++k)
      for (int i = (mono ? -2 : -3); i <= ii + (mono ? 3 : 4); ++i) std::swap(hel_3d(i, j, k), her_3d(i, j, k));
  }
  }  // full

  for (int k = 1; k <= kdm; ++k) {
    const int km = k + mm, kn = k + nn;
    for (int j = 1; j <= jdm; ++j) {
      for (int i = 1; i <= idm + 1; ++i) { P.ca(i) = cau(i, j, k); P.db(i) = pbu(i, j, n); }
      for (int i = 0; i <= idm + 1; ++i) { P.du(i) = p(i, j, k); P.dl(i) = p(i, j, k + 1); }
      for (int i = lo; i <= hi; ++i) {
        P.ai(i) = scp2i(i, j);
        P.ho(i) = std::max(c0, dp(i, j, kn)) + dpeps;
        P.hm(i) = P.ho(i);
        if (full) {
          P.hel(i) = hel_3d(i, j, k);
          P.her(i) = her_3d(i, j, k);
        }
        for (int nt = 1; nt <= ntl; ++nt) P.tm(nt, i) = S.at(nt, i, j, kn);
      }
      if (second_pass)
        for (int i = lo; i <= hi; ++i)
          P.hm(i) = P.hm(i) / (c1 - (cav(i, j + 1, k) - cav(i, j, k)) * P.ai(i));
      const size_t ro = (size_t)(j + nb - 1) * d.ldi;
      if (full && !mono)
        parabola_coeffs_fc_nosc(idm, 0, idm + 1, PI1{T.stencili.data() + ro, nb},
                                PC{T.tmc0i.data() + ro * 12, nb}, PC{T.tmcli.data() + ro * 12, nb},
                                PC{T.tmcri.data() + ro * 12, nb}, row(T.ssci, j), row(T.scci, j),
                                row(T.d2mi, j), P.hm, P.tm, P.hel, P.her, P.hpc0, P.hpc1, P.hpc2,
                                P.tpc0, P.tpc1, P.tpc2);
      else if (full)
        parabola_coeffs_fc_mono(idm, 0, idm + 1, PI1{T.stencili.data() + ro, nb},
                                PC{T.tmc0i.data() + ro * 12, nb}, PC{T.tmcli.data() + ro * 12, nb},
                                PC{T.tmcri.data() + ro * 12, nb}, row(T.ssci, j), row(T.scci, j),
                                P.hm, P.tm, P.hel, P.her, P.hpc0, P.hpc1, P.hpc2, P.tpc0, P.tpc1, P.tpc2);
      else if (!mono)
        parabola_coeffs_pc_nosc(idm, 0, idm + 1, row(T.hevc1i, j), row(T.hevc2i, j), row(T.hevc3i, j),
                                row(T.hevc4i, j), row(T.ssci, j), row(T.scci, j), row(T.d2mi, j), P.hm,
                                P.tm, P.hpc0, P.hpc1, P.hpc2, P.tpc0, P.tpc1, P.tpc2);
      else
        parabola_coeffs_pc_mono(idm, 0, idm + 1, row(T.hevc1i, j), row(T.hevc2i, j), row(T.hevc3i, j),
                                row(T.hevc4i, j), row(T.ssci, j), row(T.scci, j), P.hm, P.tm, P.hpc0,
                                P.hpc1, P.hpc2, P.tpc0, P.tpc1, P.tpc2);
      flux_integration(1, idm + 1, P.ca, P.ai, P.db, P.du, P.dl, P.hpc0, P.hpc1, P.hpc2, P.tpc0,
                       P.tpc1, P.tpc2, P.hf, P.htf);
      for (int i = 1; i <= idm; ++i) {
        double hn = P.ho(i) - (P.hf(i + 1) - P.hf(i)) * P.ai(i);
        double hni = c1 / hn;
        for (int nt = 1; nt <= ntl; ++nt)
          S.at(nt, i, j, kn) = (P.ho(i) * P.tm(nt, i) - (P.htf(nt, i + 1) - P.htf(nt, i)) * P.ai(i)) * hni;
        dp(i, j, kn) = std::max(c0, hn - dpeps);
      }
      for (int i = 1; i <= idm + 1; ++i) {
        uflx(i, j, km) = uflx(i, j, km) + P.hf(i);
        utflx(i, j, km) = utflx(i, j, km) + P.htf(1, i);
        usflx(i, j, km) = usflx(i, j, km) + P.htf(2, i);
      }
    }
  }
}

// :1625-1785 (fc_nosc), :1941-2100 (fc_mono), :2202-2300 (pc_nosc), :2401-2498 (pc_mono)
void cppm_pass_j(int m, int n, int mm, int nn, int k1m, int k1n, bool second_pass, bool full, bool mono) {
  (void)m; (void)k1m;
  Oracle& o = O(); const Dims& d = o.d;
  const int idm = d.idm, jdm = d.jdm, kdm = d.kdm, nb = d.nbdy, ii = d.ii, jj = d.jj;
  const int ntl = T.ntr_loc;
  A3 dp = o.a3("dp"), temp = o.a3("temp"), saln = o.a3("saln");
  A3 trc = d.ntr > 0 ? o.a3("trc") : A3{};
  A3 vflx = o.a3("vflx"), vtflx = o.a3("vtflx"), vsflx = o.a3("vsflx");
  A3 p = o.a3("p"), cau = o.a3("cau"), cav = o.a3("cav"), pbv = o.a3("pbv");
  A2 scp2i = o.a2("scp2i");
  A3 hel_3d{T.hel_3d.data(), d.ldi, nb, d.lev}, her_3d{T.her_3d.data(), d.ldi, nb, d.lev};
  Scalars S{temp, saln, trc, 2 * kdm};
  Pencils P(jdm, nb, ntl);
  // transposed tables: (j,i) with leading dimension ldj
  auto col = [&](std::vector<double>& v, int i) { return P1{v.data() + (size_t)(i + nb - 1) * d.ldj, nb}; };

  const int hw = mono ? 3 : 4;
  const int lo = mono ? -2 : -3, hi = jdm + (mono ? 3 : 4);
  xctilr(dp.from(k1n), 1, kdm, 0, hw, halo_ps);
  xctilr(temp.from(k1n), 1, kdm, 0, hw, halo_ps);
  xctilr(saln.from(k1n), 1, kdm, 0, hw, halo_ps);
  for (int nt = 3; nt <= ntl; ++nt) xctilr(trc.from(k1n + (nt - 3) * 2 * kdm), 1, kdm, 0, hw, halo_ps);

  if (full) {
  for (int k = 1; k <= kdm; ++k) {
    const int kn = k + nn;
    for (int i = 1; i <= idm; ++i) {
      for (int j = -2; j <= jdm + 3; ++j) {
        P.ai(j) = scp2i(i, j);
        P.hm(j) = std::max(c0, dp(i, j, kn)) + dpeps;
      }
      if (second_pass)
        for (int j = -2; j <= jdm + 3; ++j)
          P.hm(j) = P.hm(j) / (c1 - (cau(i + 1, j, k) - cau(i, j, k)) * P.ai(j));
      if (mono)
        h_edges_mono(jdm, 1, jdm, col(T.hevc1j, i), col(T.hevc2j, i), col(T.hevc3j, i),
                     col(T.hevc4j, i), col(T.sscj, i), col(T.sccj, i), P.hm, P.hel, P.her);
      else
        h_edges_nosc(jdm, 1, jdm, col(T.hevc1j, i), col(T.hevc2j, i), col(T.hevc3j, i),
                     col(T.hevc4j, i), col(T.sscj, i), col(T.sccj, i), col(T.d2mj, i), P.hm, P.hel,
                     P.her);
      for (int j = 1; j <= jdm; ++j) { hel_3d(i, j, k) = P.hel(j); her_3d(i, j, k) = P.her(j); }
    }
  }
  xctilr(hel_3d, 1, kdm, 0, hw, halo_ps);
  xctilr(her_3d, 1, kdm, 0, hw, halo_ps);
  if (d.nreg == 2) {  // :1687-1703 / :2003-2019
    const bool fold_fix = o.option("cppm_fold_fix", "0") == "1";
    for (int k = 1; k <= kdm; ++k) {
      int j = jj;
      // Reference quirk kept by default: only the right half of row jj is swapped although
      // the whole row is a mirrored duplicate (p-type); option cppm_fold_fix=1 swaps the
      // whole row, which restores round-off mass conservation across the fold.
      for (int i = (fold_fix ? 1 : std::max(1, d.itdm / 2 - d.i0 + 1)); i <= ii; ++i)
        std::swap(hel_3d(i, j, k), her_3d(i, j, k));
      for (j = jj + 1; j <= jj + hw; ++j)
        for (int i = 1; i <= ii; ++i) std::swap(hel_3d(i, j, k), her_3d(i, j, k));
    }
  }
  }  // full

  for (int k = 1; k <= kdm; ++k) {
    const int km = k + mm, kn = k + nn;
    for (int i = 1; i <= idm; ++i) {
      for (int j = 1; j <= jdm + 1; ++j) { P.ca(j) = cav(i, j, k); P.db(j) = pbv(i, j, n); }
      for (int j = 0; j <= jdm + 1; ++j) { P.du(j) = p(i, j, k); P.dl(j) = p(i, j, k + 1); }
      for (int j = lo; j <= hi; ++j) {
        P.ai(j) = scp2i(i, j);
        P.ho(j) = std::max(c0, dp(i, j, kn)) + dpeps;
        P.hm(j) = P.ho(j);
        if (full) {
          P.hel(j) = hel_3d(i, j, k);
          P.her(j) = her_3d(i, j, k);
        }
        for (int nt = 1; nt <= ntl; ++nt) P.tm(nt, j) = S.at(nt, i, j, kn);
      }
      if (second_pass)
        for (int j = lo; j <= hi; ++j)
          P.hm(j) = P.hm(j) / (c1 - (cau(i + 1, j, k) - cau(i, j, k)) * P.ai(j));
      const size_t co = (size_t)(i + nb - 1) * d.ldj;
      if (full && !mono)
        parabola_coeffs_fc_nosc(jdm, 0, jdm + 1, PI1{T.stencilj.data() + co, nb},
                                PC{T.tmc0j.data() + co * 12, nb}, PC{T.tmclj.data() + co * 12, nb},
                                PC{T.tmcrj.data() + co * 12, nb}, col(T.sscj, i), col(T.sccj, i),
                                col(T.d2mj, i), P.hm, P.tm, P.hel, P.her, P.hpc0, P.hpc1, P.hpc2,
                                P.tpc0, P.tpc1, P.tpc2);
      else if (full)
        parabola_coeffs_fc_mono(jdm, 0, jdm + 1, PI1{T.stencilj.data() + co, nb},
                                PC{T.tmc0j.data() + co * 12, nb}, PC{T.tmclj.data() + co * 12, nb},
                                PC{T.tmcrj.data() + co * 12, nb}, col(T.sscj, i), col(T.sccj, i),
                                P.hm, P.tm, P.hel, P.her, P.hpc0, P.hpc1, P.hpc2, P.tpc0, P.tpc1, P.tpc2);
      else if (!mono)
        parabola_coeffs_pc_nosc(jdm, 0, jdm + 1, col(T.hevc1j, i), col(T.hevc2j, i), col(T.hevc3j, i),
                                col(T.hevc4j, i), col(T.sscj, i), col(T.sccj, i), col(T.d2mj, i), P.hm,
                                P.tm, P.hpc0, P.hpc1, P.hpc2, P.tpc0, P.tpc1, P.tpc2);
      else
        parabola_coeffs_pc_mono(jdm, 0, jdm + 1, col(T.hevc1j, i), col(T.hevc2j, i), col(T.hevc3j, i),
                                col(T.hevc4j, i), col(T.sscj, i), col(T.sccj, i), P.hm, P.tm, P.hpc0,
                                P.hpc1, P.hpc2, P.tpc0, P.tpc1, P.tpc2);
      flux_integration(1, jdm + 1, P.ca, P.ai, P.db, P.du, P.dl, P.hpc0, P.hpc1, P.hpc2, P.tpc0,
                       P.tpc1, P.tpc2, P.hf, P.htf);
      for (int j = 1; j <= jdm; ++j) {
        double hn = P.ho(j) - (P.hf(j + 1) - P.hf(j)) * P.ai(j);
        double hni = c1 / hn;
        for (int nt = 1; nt <= ntl; ++nt)
          S.at(nt, i, j, kn) = (P.ho(j) * P.tm(nt, j) - (P.htf(nt, j + 1) - P.htf(nt, j)) * P.ai(j)) * hni;
        dp(i, j, kn) = std::max(c0, hn - dpeps);
      }
      for (int j = 1; j <= jdm + 1; ++j) {
        vflx(i, j, km) = vflx(i, j, km) + P.hf(j);
        vtflx(i, j, km) = vtflx(i, j, km) + P.htf(1, j);
        vsflx(i, j, km) = vsflx(i, j, km) + P.htf(2, j);
      }
    }
  }
}

void swap_stencil_tag(int& s) {  // :2653-2666
  switch (s) {
    case stencil_1110: s = stencil_0111; break;
    case stencil_0111: s = stencil_1110; break;
    case stencil_1100: s = stencil_0011; break;
    case stencil_0011: s = stencil_1100; break;
    case stencil_0100: s = stencil_0010; break;
    case stencil_0010: s = stencil_0100; break;
    default: break;
  }
}

}  // namespace

// :2504-2746 (the tables do not depend on the compatibility / limiting options)
void init_cppm() {
  Oracle& o = O(); const Dims& d = o.d;
  const int idm = d.idm, jdm = d.jdm, nb = d.nbdy, ii = d.ii, jj = d.jj;
  const size_t L = d.lev;
  I2 ip = o.i2("ip");
  A2 scpx = o.a2("scpx"), scpy = o.a2("scpy");
  auto Z = [&](std::vector<double>& v, size_t n) { v.assign(n, 0.0); };
  T.stencili.assign(L, 0);
  Z(T.hevc1i, L); Z(T.hevc2i, L); Z(T.hevc3i, L); Z(T.hevc4i, L);
  Z(T.tmc0i, 12 * L); Z(T.tmcli, 12 * L); Z(T.tmcri, 12 * L);
  Z(T.ssci, L); Z(T.scci, L); Z(T.d2mi, L);
  std::vector<int> stencilj_perm(L, 0);
  std::vector<double> h1p(L, 0.), h2p(L, 0.), h3p(L, 0.), h4p(L, 0.), t0p(12 * L, 0.), tlp(12 * L, 0.),
      trp(12 * L, 0.), sscp(L, 0.), sccp(L, 0.), d2mp(L, 0.), tmp2d(L, 0.);
  auto ix = [&](int i, int j) { return (size_t)(j + nb - 1) * d.ldi + (i + nb - 1); };

  for (int j = 1; j <= jj; ++j)
    for (int i = 1; i <= ii; ++i) {
      int sm4[4]; double dx4[4]; int sm3[3]; double dx3[3];
      for (int q = 0; q < 4; ++q) { sm4[q] = ip(i - 2 + q, j); dx4[q] = scpx(i - 2 + q, j); }
      size_t x = ix(i, j);
      set_stencil_coeffs(sm4, dx4, T.stencili[x], T.hevc1i[x], T.hevc2i[x], T.hevc3i[x], T.hevc4i[x],
                         &T.tmc0i[12 * x], &T.tmcli[12 * x], &T.tmcri[12 * x]);
      for (int q = 0; q < 3; ++q) { sm3[q] = ip(i - 1 + q, j); dx3[q] = scpx(i - 1 + q, j); }
      set_slope_coeffs(sm3, dx3, T.ssci[x], T.scci[x]);
      set_d2_mask(sm3, T.d2mi[x]);
      for (int q = 0; q < 4; ++q) { sm4[q] = ip(i, j - 2 + q); dx4[q] = scpy(i, j - 2 + q); }
      set_stencil_coeffs(sm4, dx4, stencilj_perm[x], h1p[x], h2p[x], h3p[x], h4p[x], &t0p[12 * x],
                         &tlp[12 * x], &trp[12 * x]);
      for (int q = 0; q < 3; ++q) { sm3[q] = ip(i, j - 1 + q); dx3[q] = scpy(i, j - 1 + q); }
      set_slope_coeffs(sm3, dx3, sscp[x], sccp[x]);
      set_d2_mask(sm3, d2mp[x]);
    }

  auto V2 = [&](std::vector<double>& v) { return A2{v.data(), d.ldi, nb}; };
  auto tile_int = [&](std::vector<int>& s, int mh, int nh, int it) {
    for (size_t x = 0; x < L; ++x) tmp2d[x] = (double)s[x];
    xctilr(V2(tmp2d), mh, nh, it);
    for (size_t x = 0; x < L; ++x) s[x] = (int)std::lround(tmp2d[x]);
  };
  auto tile12 = [&](std::vector<double>& t, int mh, int nh, int it) {
    for (int k = 0; k < 12; ++k) {
      for (size_t x = 0; x < L; ++x) tmp2d[x] = t[12 * x + k];
      xctilr(V2(tmp2d), mh, nh, it);
      for (size_t x = 0; x < L; ++x) t[12 * x + k] = tmp2d[x];
    }
  };
  // :2605-2625
  tile_int(T.stencili, nb, 0, halo_us);
  xctilr(V2(T.hevc1i), nb, 0, halo_us); xctilr(V2(T.hevc2i), nb, 0, halo_us);
  xctilr(V2(T.hevc3i), nb, 0, halo_us); xctilr(V2(T.hevc4i), nb, 0, halo_us);
  tile12(T.tmc0i, nb, 0, halo_us); tile12(T.tmcli, nb, 0, halo_us); tile12(T.tmcri, nb, 0, halo_us);
  xctilr(V2(T.ssci), nb, 0, halo_ps); xctilr(V2(T.scci), nb, 0, halo_ps); xctilr(V2(T.d2mi), nb, 0, halo_ps);
  // :2626-2646
  tile_int(stencilj_perm, 0, nb, halo_vs);
  xctilr(V2(h1p), 0, nb, halo_vs); xctilr(V2(h2p), 0, nb, halo_vs);
  xctilr(V2(h3p), 0, nb, halo_vs); xctilr(V2(h4p), 0, nb, halo_vs);
  tile12(t0p, 0, nb, halo_vs); tile12(tlp, 0, nb, halo_vs); tile12(trp, 0, nb, halo_vs);
  xctilr(V2(sscp), 0, nb, halo_ps); xctilr(V2(sccp), 0, nb, halo_ps); xctilr(V2(d2mp), 0, nb, halo_ps);

  // :2650-2720 arctic swaps
  if (d.nreg == 2) {
    int j = jj;
    for (int i = 1 - nb; i <= ii + nb; ++i) {
      size_t x = ix(i, j);
      swap_stencil_tag(T.stencili[x]);
      std::swap(T.hevc1i[x], T.hevc4i[x]);
      std::swap(T.hevc2i[x], T.hevc3i[x]);
    }
    for (int i = std::max(1, d.itdm / 2 - d.i0 + 1); i <= ii; ++i) {
      size_t x = ix(i, j);
      swap_stencil_tag(stencilj_perm[x]);
      std::swap(h1p[x], h4p[x]);
      std::swap(h2p[x], h3p[x]);
    }
    for (j = jj + 1; j <= jj + nb; ++j)
      for (int i = 1; i <= ii; ++i) {
        size_t x = ix(i, j);
        swap_stencil_tag(stencilj_perm[x]);
        std::swap(h1p[x], h4p[x]);
        std::swap(h2p[x], h3p[x]);
      }
  }

  // :2722-2736 transpose
  T.stencilj.assign(L, 0);
  Z(T.hevc1j, L); Z(T.hevc2j, L); Z(T.hevc3j, L); Z(T.hevc4j, L);
  Z(T.tmc0j, 12 * L); Z(T.tmclj, 12 * L); Z(T.tmcrj, 12 * L);
  Z(T.sscj, L); Z(T.sccj, L); Z(T.d2mj, L);
  for (int j = 1 - nb; j <= jdm + nb; ++j)
    for (int i = 1 - nb; i <= idm + nb; ++i) {
      size_t x = ix(i, j);
      size_t y = (size_t)(i + nb - 1) * d.ldj + (j + nb - 1);
      T.stencilj[y] = stencilj_perm[x];
      T.hevc1j[y] = h1p[x]; T.hevc2j[y] = h2p[x]; T.hevc3j[y] = h3p[x]; T.hevc4j[y] = h4p[x];
      for (int k = 0; k < 12; ++k) {
        T.tmc0j[12 * y + k] = t0p[12 * x + k];
        T.tmclj[12 * y + k] = tlp[12 * x + k];
        T.tmcrj[12 * y + k] = trp[12 * x + k];
      }
      T.sscj[y] = sscp[x]; T.sccj[y] = sccp[x]; T.d2mj[y] = d2mp[x];
    }
  T.hel_3d.assign(L * d.kdm, 0.0);
  T.her_3d.assign(L * d.kdm, 0.0);
  T.ntr_loc = 2 + d.ntr;
}

// debugging/test access to the tables (name -> pointer,len)
const double* cppm_table(const char* name, size_t* n) {
  std::string s(name);
  std::vector<double>* v = nullptr;
#define TB(x) if (s == #x) v = &T.x;
  TB(hevc1i) TB(hevc2i) TB(hevc3i) TB(hevc4i) TB(ssci) TB(scci) TB(d2mi)
  TB(hevc1j) TB(hevc2j) TB(hevc3j) TB(hevc4j) TB(sscj) TB(sccj) TB(d2mj)
  TB(tmc0i) TB(tmcli) TB(tmcri) TB(tmc0j) TB(tmclj) TB(tmcrj) TB(hel_3d) TB(her_3d)
#undef TB
  if (!v) { *n = 0; return nullptr; }
  *n = v->size();
  return v->data();
}
const int* cppm_stencil(const char* name, size_t* n) {
  std::string s(name);
  std::vector<int>* v = s == "stencili" ? &T.stencili : s == "stencilj" ? &T.stencilj : nullptr;
  if (!v) { *n = 0; return nullptr; }
  *n = v->size();
  return v->data();
}

// :2748-2834
void cppm(int m, int n, int mm, int nn, int k1m, int k1n) {
  Oracle& o = O(); const Dims& d = o.d;
  const int nstep = (int)o.scalar("nstep");
  const std::string comp = o.option("cppm_compatibility", "full"), lim = o.option("cppm_limiting", "non_oscillatory");
  if (comp != "full" && comp != "partial")
    throw std::runtime_error(" init_cppm: cppm_compatibility = " + comp + " is unsupported!");
  if (lim != "monotonic" && lim != "non_oscillatory")
    throw std::runtime_error(" init_cppm: cppm_limiting = " + lim + " is unsupported!");
  const bool full = comp == "full", mono = lim == "monotonic";
  const int hw = mono ? 3 : 4;
  xctilr(o.a3("cau"), 1, d.kdm, hw, hw, halo_uv);
  xctilr(o.a3("cav"), 1, d.kdm, hw, hw, halo_vv);
  if (nstep % 2 == 1) {
    cppm_pass_i(m, n, mm, nn, k1m, k1n, false, full, mono);
    cppm_pass_j(m, n, mm, nn, k1m, k1n, true, full, mono);
  } else {
    cppm_pass_j(m, n, mm, nn, k1m, k1n, false, full, mono);
    cppm_pass_i(m, n, mm, nn, k1m, k1n, true, full, mono);
  }
}

// phy/mod_advect.F90:59-189, advmth='cppm'
void advect(int m, int n, int mm, int nn, int k1m, int k1n) {
  Oracle& o = O(); const Dims& d = o.d;
  const int ii = d.ii, jj = d.jj, kk = d.kk;
  const double delt1 = o.scalar("delt1"), dlt = o.scalar("dlt");
  A3 u = o.a3("u"), v = o.a3("v"), dpu = o.a3("dpu"), dpv = o.a3("dpv");
  A3 cau = o.a3("cau"), cav = o.a3("cav"), pbu = o.a3("pbu"), pbv = o.a3("pbv");
  A3 ubflxs_p = o.a3("ubflxs_p"), vbflxs_p = o.a3("vbflxs_p");
  A3 umfltd = o.a3("umfltd"), vmfltd = o.a3("vmfltd"), umflsm = o.a3("umflsm"), vmflsm = o.a3("vmflsm");
  A2 scuy = o.a2("scuy"), scvx = o.a2("scvx"), umax = o.a2("umax"), vmax = o.a2("vmax");
  I2 iu = o.i2("iu"), iv = o.i2("iv");
  for (int j = 1; j <= jj; ++j)
    for (int k = 1; k <= kk; ++k) {
      const int km = k + mm, kn = k + nn;
      for (int i = 1; i <= ii; ++i) {
        if (iu(i, j) == 1) {  // span loop isu/ifu/ilu clipped to 1..ii (:75-76)
          double dtdl = delt1 * scuy(i, j);
          double ca_tmp = u(i, j, km) * dtdl + ubflxs_p(i, j, m) * dlt / pbu(i, j, m) +
                          (umfltd(i, j, km) + umflsm(i, j, km)) / std::max(onemm, dpu(i, j, kn));
          cau(i, j, k) = std::max(-umax(i, j) * dtdl, std::min(umax(i, j) * dtdl, ca_tmp));
        }
      }
      for (int i = 1; i <= ii; ++i) {
        if (iv(i, j) == 1) {
          double dtdl = delt1 * scvx(i, j);
          double ca_tmp = v(i, j, km) * dtdl + vbflxs_p(i, j, m) * dlt / pbv(i, j, m) +
                          (vmfltd(i, j, km) + vmflsm(i, j, km)) / std::max(onemm, dpv(i, j, kn));
          cav(i, j, k) = std::max(-vmax(i, j) * dtdl, std::min(vmax(i, j) * dtdl, ca_tmp));
        }
      }
    }
  const std::string advmth = o.option("advmth", "cppm");
  if (advmth == "remap") { advect_remap(m, n, mm, nn, k1m, k1n); return; }  // :96-153 (no trailing halo update)
  if (advmth != "cppm") throw std::runtime_error(" advmth = " + advmth + " is unsupported!");
  cppm(m, n, mm, nn, k1m, k1n);
  xctilr(o.a3("dp").from(k1n), 1, kk, 1, 1, halo_ps);
  xctilr(o.a3("temp").from(k1n), 1, kk, 1, 1, halo_ps);
  xctilr(o.a3("saln").from(k1n), 1, kk, 1, 1, halo_ps);
  for (int nt = 1; nt <= d.ntr; ++nt)
    xctilr(o.a3("trc").from(k1n + (nt - 1) * 2 * d.kdm), 1, kk, 1, 1, halo_ps);
}

}  // namespace orc
