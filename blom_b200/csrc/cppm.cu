// Layer-thickness and tracer transport: advect prelude + CPPM, all four variants
// (cppm_compatibility='full'|'partial' x cppm_limiting='non_oscillatory'|'monotonic').
//
// Reference: phy/mod_advect.F90:59-189, phy/mod_cppm.F90:101-359 (static
// tables), :361-434 (thickness edges), :490-818 (compatible tracer parabolas),
// :1373-1468 (flux integrals), :1470-1785 (directional passes), :2504-2834.
//
// B200 design (not the reference's pencil loops):
//  * tables are SoA device arrays in (i,j) layout for BOTH directions (the
//    reference transposes the j tables for its CPU pencils; lanes run along i
//    here in both passes so no transpose is wanted);
//  * each directional pass is two kernels: `hedges` (limited 4th-order
//    thickness edges) and `flux` (tracer edge LU solves, limiters, parabolas,
//    upstream flux integrals, divergence update) — `flux` runs the dependent
//    stages edge -> curvature -> parabola -> face flux -> cell update through
//    shared memory inside one thread block with a 2+3-cell skirt, so the
//    reference's tel/ter/d2t/tpc*/hf/htf pencils never touch HBM;
//  * the two Strang passes ping-pong between the state arrays and a scratch
//    set, which removes the in-place read/write hazard without extra traffic.
#include "common.cuh"

namespace blom {

namespace {

enum { stencil_0000 = 0, stencil_1111 = 1, stencil_1110 = 2, stencil_0111 = 3,
       stencil_1100 = 4, stencil_0110 = 5, stencil_0011 = 6, stencil_0100 = 7,
       stencil_0010 = 8 };

#define K0 0.
#define K1 1.
#define K2 2.
#define K3 3.
#define K4 4.
#define K5 5.
#define K6 6.
#define K12 12.
#define K18 18.
#define K42 42.
#define K60 60.
#define K1_2 (1. / 2.)
#define K1_3 (1. / 3.)
#define K2_3 (2. / 3.)
#define K1_4 (1. / 4.)
#define K3_4 (3. / 4.)
#define K1_5 (1. / 5.)
#define K1_6 (1. / 6.)
#define K1_10 (1. / 10.)
#define K1_12 (1. / 12.)
#define K1_15 (1. / 15.)
#define K1_20 (1. / 20.)
#define DPEPS 1.e-12

// table levels inside the per-direction table pack
enum { T_HEVC1 = 0, T_HEVC2 = 1, T_HEVC3 = 2, T_HEVC4 = 3, T_TMC0 = 4, T_TMCL = 16, T_TMCR = 28,
       T_SSC = 40, T_SCC = 41, T_D2M = 42, T_STEN = 43, T_NLEV = 44 };

__device__ __forceinline__ double fsign(double a, double b) { return copysign(a, b); }

// ---- static tables -----------------------------------------------------------
// One thread per interior point computes both directions' tables
// (set_stencil_coeffs / set_slope_coeffs / set_d2_mask, mod_cppm.F90:101-359).
// Moment coefficients of the 4-cell edge stencil (mod_cppm.F90:118-175): functions of the four cell
// widths only.  Used by the table kernel and recomputed inside the flux kernel (cheaper than
// re-reading 36 table words per interface and level).
struct TmCoef { double t0[12], tl[12], tr[12]; };
__device__ __forceinline__ void tm_coeffs(double d1, double d2, double d3, double d4, TmCoef& c, double a[12]) {
  double a12, a22, a32, a42, a13, a23, a33, a43, a14, a24, a34, a44;
  a12 = -d2 - K1_2 * d1;
  a22 = -K1_2 * d2;
  a32 = K1_2 * d3;
  a42 = d3 + K1_2 * d4;
  a13 = a12 * a12 + K1_12 * d1 * d1;
  a23 = -K2_3 * a22 * d2;
  a33 = K2_3 * a32 * d3;
  a43 = a42 * a42 + K1_12 * d4 * d4;
  a14 = (a13 + K1_6 * d1 * d1) * a12;
  a24 = -K3_4 * a23 * d2;
  a34 = K3_4 * a33 * d3;
  a44 = (a43 + K1_6 * d4 * d4) * a42;
  double* tl = c.tl; double* tr = c.tr; double* t0 = c.t0;
  tl[0] = -K1_12 * d1;
  tl[1] = (K1_10 * d1 + K1_6 * d2) * d1;
  tl[2] = -(K1_10 * (d1 + K3 * d2) * d1 + K1_4 * (d2 * d2)) * d1;
  tl[3] = -K1_12 * d2;
  tl[4] = K1_10 * (d2 * d2);
  tl[5] = -K1_10 * (d2 * d2 * d2);
  tl[6] = -K1_12 * d3;
  tl[7] = -K1_15 * (d3 * d3);
  tl[8] = -K1_20 * (d3 * d3 * d3);
  tl[9] = -K1_12 * d4;
  tl[10] = -(K1_15 * d4 + K1_6 * d3) * d4;
  tl[11] = -(K1_5 * (K1_4 * d4 + d3) * d4 + K1_4 * (d3 * d3)) * d4;
  tr[0] = K1_12 * d1;
  tr[1] = -(K1_15 * d1 + K1_6 * d2) * d1;
  tr[2] = (K1_5 * (K1_4 * d1 + d2) * d1 + K1_4 * (d2 * d2)) * d1;
  tr[3] = K1_12 * d2;
  tr[4] = -K1_15 * (d2 * d2);
  tr[5] = K1_20 * (d2 * d2 * d2);
  tr[6] = K1_12 * d3;
  tr[7] = K1_10 * (d3 * d3);
  tr[8] = K1_10 * (d3 * d3 * d3);
  tr[9] = K1_12 * d4;
  tr[10] = (K1_10 * d4 + K1_6 * d3) * d4;
  tr[11] = (K1_10 * (d4 + K3 * d3) * d4 + K1_4 * (d3 * d3)) * d4;
  t0[0] = a12;
  t0[1] = a13 - tl[1] - tr[1];
  t0[2] = a14 - tl[2] - tr[2];
  t0[3] = a22;
  t0[4] = a23 - tl[4] - tr[4];
  t0[5] = a24 - tl[5] - tr[5];
  t0[6] = a32;
  t0[7] = a33 - tl[7] - tr[7];
  t0[8] = a34 - tl[8] - tr[8];
  t0[9] = a42;
  t0[10] = a43 - tl[10] - tr[10];
  t0[11] = a44 - tl[11] - tr[11];
  a[0] = a12; a[1] = a22; a[2] = a32; a[3] = a42; a[4] = a13; a[5] = a23; a[6] = a33; a[7] = a43;
  a[8] = a14; a[9] = a24; a[10] = a34; a[11] = a44;
}

__device__ void stencil_coeffs(const int* sm, const double* dx, double* tab, long lev) {
  TmCoef tc; double av[12];
  tm_coeffs(dx[0], dx[1], dx[2], dx[3], tc, av);
  double a12 = av[0], a22 = av[1], a32 = av[2], a42 = av[3], a13 = av[4], a23 = av[5], a33 = av[6], a43 = av[7],
         a14 = av[8], a24 = av[9], a34 = av[10], a44 = av[11];
  const double* t0 = tc.t0; const double* tl = tc.tl; const double* tr = tc.tr;
#pragma unroll
  for (int r = 0; r < 12; ++r) {
    tab[(T_TMC0 + r) * lev] = t0[r];
    tab[(T_TMCL + r) * lev] = tl[r];
    tab[(T_TMCR + r) * lev] = tr[r];
  }
  int stencil;
  double hevc1, hevc2, hevc3, hevc4;
  const int s0 = sm[0], s1 = sm[1], s2 = sm[2], s3 = sm[3];
  if (s0 == 1 && s1 == 1 && s2 == 1 && s3 == 1) {
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
    hevc1 = K1 - hevc2 - hevc3 - hevc4;
  } else if (s0 == 1 && s1 == 1 && s2 == 1 && s3 == 0) {
    stencil = stencil_1110;
    a22 = a22 - a12; a32 = a32 - a12;
    a23 = (a23 - a13) / a22;
    a33 = a33 - a13 - a23 * a32;
    hevc2 = -a12;
    hevc3 = -a13 - a23 * hevc2;
    hevc3 = hevc3 / a33;
    hevc2 = (hevc2 - a32 * hevc3) / a22;
    hevc1 = K1 - hevc2 - hevc3;
    hevc4 = K0;
  } else if (s0 == 0 && s1 == 1 && s2 == 1 && s3 == 1) {
    stencil = stencil_0111;
    a32 = a32 - a22; a42 = a42 - a22;
    a33 = (a33 - a23) / a32;
    a43 = a43 - a23 - a33 * a42;
    hevc3 = -a22;
    hevc4 = -a23 - a33 * hevc3;
    hevc4 = hevc4 / a43;
    hevc3 = (hevc3 - a42 * hevc4) / a32;
    hevc2 = K1 - hevc3 - hevc4;
    hevc1 = K0;
  } else if (s0 == 0 && s1 == 1 && s2 == 1 && s3 == 0) {
    stencil = stencil_0110;
    a32 = a32 - a22;
    hevc3 = -a22 / a32;
    hevc2 = K1 - hevc3;
    hevc1 = K0; hevc4 = K0;
  } else if (s0 == 1 && s1 == 1) {
    stencil = stencil_1100;
    a22 = a22 - a12;
    hevc2 = -a12 / a22;
    hevc1 = K1 - hevc2;
    hevc3 = K0; hevc4 = K0;
  } else if (s2 == 1 && s3 == 1) {
    stencil = stencil_0011;
    a42 = a42 - a32;
    hevc4 = -a32 / a42;
    hevc3 = K1 - hevc4;
    hevc1 = K0; hevc2 = K0;
  } else if (s1 == 1) {
    stencil = stencil_0100;
    hevc1 = K0; hevc2 = K1; hevc3 = K0; hevc4 = K0;
  } else if (s2 == 1) {
    stencil = stencil_0010;
    hevc1 = K0; hevc2 = K0; hevc3 = K1; hevc4 = K0;
  } else {
    stencil = stencil_0000;
    hevc1 = K0; hevc2 = K0; hevc3 = K0; hevc4 = K0;
  }
  tab[T_HEVC1 * lev] = hevc1; tab[T_HEVC2 * lev] = hevc2;
  tab[T_HEVC3 * lev] = hevc3; tab[T_HEVC4 * lev] = hevc4;
  tab[T_STEN * lev] = (double)stencil;
}

__global__ void cppm_tables_kernel(Geom g, const int* __restrict__ ip, const double* __restrict__ scpx,
                                   const double* __restrict__ scpy, double* __restrict__ ti,
                                   double* __restrict__ tj) {
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long)g.ii * g.jj) return;
  const int i = (int)(t % g.ii) + 1, j = (int)(t / g.ii) + 1;
  const long x = ix2(g, i, j);
  int sm[4]; double dx[4];
  for (int q = 0; q < 4; ++q) { sm[q] = ip[ix2(g, i - 2 + q, j)]; dx[q] = scpx[ix2(g, i - 2 + q, j)]; }
  stencil_coeffs(sm, dx, ti + x, g.lev);
  {
    const bool any0 = sm[1] == 0 || sm[2] == 0 || sm[3] == 0;
    ti[T_SSC * g.lev + x] = any0 ? K0 : K2;
    ti[T_SCC * g.lev + x] = any0 ? K0 : K2 * dx[2] / (dx[1] + K2 * dx[2] + dx[3]);
    ti[T_D2M * g.lev + x] = any0 ? K0 : K1;
  }
  for (int q = 0; q < 4; ++q) { sm[q] = ip[ix2(g, i, j - 2 + q)]; dx[q] = scpy[ix2(g, i, j - 2 + q)]; }
  stencil_coeffs(sm, dx, tj + x, g.lev);
  {
    const bool any0 = sm[1] == 0 || sm[2] == 0 || sm[3] == 0;
    tj[T_SSC * g.lev + x] = any0 ? K0 : K2;
    tj[T_SCC * g.lev + x] = any0 ? K0 : K2 * dx[2] / (dx[1] + K2 * dx[2] + dx[3]);
    tj[T_D2M * g.lev + x] = any0 ? K0 : K1;
  }
}

__device__ __forceinline__ double swap_tag(double s) {
  switch ((int)s) {
    case stencil_1110: return stencil_0111;
    case stencil_0111: return stencil_1110;
    case stencil_1100: return stencil_0011;
    case stencil_0011: return stencil_1100;
    case stencil_0100: return stencil_0010;
    case stencil_0010: return stencil_0100;
    default: return s;
  }
}
// arctic swaps of tags / hevc (mod_cppm.F90:2650-2720); northern tile only
__global__ void cppm_tables_arctic(Geom g, double* ti, double* tj) {
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nrow = 1 + g.nb;  // row jj (i tables + right half of j tables) and rows jj+1..jj+nb
  if (t >= (long)g.ldi * nrow) return;
  const int i = (int)(t % g.ldi) + 1 - g.nb, r = (int)(t / g.ldi);
  const int j = g.jj + r;
  const long x = ix2(g, i, j);
  auto swp = [&](double* tb) {
    tb[T_STEN * g.lev + x] = swap_tag(tb[T_STEN * g.lev + x]);
    double a = tb[T_HEVC1 * g.lev + x]; tb[T_HEVC1 * g.lev + x] = tb[T_HEVC4 * g.lev + x]; tb[T_HEVC4 * g.lev + x] = a;
    a = tb[T_HEVC2 * g.lev + x]; tb[T_HEVC2 * g.lev + x] = tb[T_HEVC3 * g.lev + x]; tb[T_HEVC3 * g.lev + x] = a;
  };
  if (r == 0) {
    swp(ti);
    if (i >= max(1, g.itdm / 2 - g.i0 + 1) && i <= g.ii) swp(tj);
  } else if (i >= 1 && i <= g.ii) {
    swp(tj);
  }
}
__global__ void cppm_tags_to_int(Geom g, const double* __restrict__ ti, const double* __restrict__ tj,
                                 int* __restrict__ si, int* __restrict__ sj) {
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= g.lev) return;
  si[t] = __double2int_rn(ti[T_STEN * g.lev + t]);
  sj[t] = __double2int_rn(tj[T_STEN * g.lev + t]);
}

// ---- advect prelude (mod_advect.F90:71-94) ----------------------------------
template <int MINB>
__global__ void __launch_bounds__(128, MINB) advect_flux_area(Geom g, int m, int mm, int nn, double delt1, double dlt,
                                 const int* __restrict__ iu, const int* __restrict__ iv,
                                 const double* __restrict__ u, const double* __restrict__ v,
                                 const double* __restrict__ dpu, const double* __restrict__ dpv,
                                 const double* __restrict__ ubflxs_p, const double* __restrict__ vbflxs_p,
                                 const double* __restrict__ pbu, const double* __restrict__ pbv,
                                 const double* __restrict__ umfltd, const double* __restrict__ vmfltd,
                                 const double* __restrict__ umflsm, const double* __restrict__ vmflsm,
                                 const double* __restrict__ scuy, const double* __restrict__ scvx,
                                 const double* __restrict__ umax, const double* __restrict__ vmax,
                                 double* __restrict__ cau, double* __restrict__ cav) {
  const Bid b_ = bid(g);
  const int i = b_.x * blockDim.x + threadIdx.x + 1;
  const int j = b_.y + 1, k = b_.z + 1;
  if (i > g.ii) return;
  const long x = ix2(g, i, j);
  const long xm = x + (long)(k + mm - 1) * g.lev, xn = x + (long)(k + nn - 1) * g.lev;
  const long xk = x + (long)(k - 1) * g.lev, x2 = x + (long)(m - 1) * g.lev;
  if (iu[x] == 1) {
    double dtdl = delt1 * scuy[x];
    double ca_tmp = u[xm] * dtdl + ubflxs_p[x2] * dlt / pbu[x2] + (umfltd[xm] + umflsm[xm]) / fmax(onemm, dpu[xn]);
    cau[xk] = fmax(-umax[x] * dtdl, fmin(umax[x] * dtdl, ca_tmp));
  }
  if (iv[x] == 1) {
    double dtdl = delt1 * scvx[x];
    double ca_tmp = v[xm] * dtdl + vbflxs_p[x2] * dlt / pbv[x2] + (vmfltd[xm] + vmflsm[xm]) / fmax(onemm, dpv[xn]);
    cav[xk] = fmax(-vmax[x] * dtdl, fmin(vmax[x] * dtdl, ca_tmp));
  }
}

// ---- thickness edges (h_edges_nosc, mod_cppm.F90:361-434) --------------------
// DIR 0: i-pass, DIR 1: j-pass.  One thread per interior cell and level.
// sp = element stride along the pass direction, sc = along the cross direction.
// MONO: h_edges_mono (:436-488) — same edges, limiter applied unconditionally, no positivity fix.
template <int DIR, bool MONO>
__global__ void __launch_bounds__(256)
cppm_hedges(Geom g, bool second_pass, const double* __restrict__ dp /* level 1 of source set */,
            const double* __restrict__ cac /* cross-direction flux area, level 1 */,
            const double* __restrict__ scp2i, const double* __restrict__ tab,
            double* __restrict__ hel3, double* __restrict__ her3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y + 1, k = blockIdx.z + 1;
  if (i > g.ii) return;
  const long sp = DIR == 0 ? 1 : g.ldi, sc = DIR == 0 ? g.ldi : 1;
  const long x = ix2(g, i, j), xk = x + (long)(k - 1) * g.lev;
  double hm[7];  // cells c-3..c+3
#pragma unroll
  for (int q = 0; q < 7; ++q) {
    const long y = x + (q - 3) * sp, yk = xk + (q - 3) * sp;
    double h = fmax(K0, dp[yk]) + DPEPS;
    if (second_pass) h = h / (K1 - (cac[yk + sc] - cac[yk]) * scp2i[y]);
    hm[q] = h;
  }
  // edges e = c-1..c+2 -> he[0..3]; edge e uses cells e-2..e+1
  double he[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const long y = x + (q - 1) * sp;
    he[q] = tab[T_HEVC1 * g.lev + y] * hm[q] + tab[T_HEVC2 * g.lev + y] * hm[q + 1] +
            tab[T_HEVC3 * g.lev + y] * hm[q + 2] + tab[T_HEVC4 * g.lev + y] * hm[q + 3];
  }
  // d2h at c-1, c, c+1: hel(c')=he(c'), her(c')=he(c'+1)
  double d2h[3];
#pragma unroll
  for (int q = 0; q < 3; ++q)
    d2h[q] = tab[T_D2M * g.lev + x + (q - 1) * sp] * (he[q] - K2 * hm[q + 2] + he[q + 1]);
  double hel = he[1], her = he[2];
  const double hmm = hm[2], hm0 = hm[3], hmp = hm[4];
  const double ssc = tab[T_SSC * g.lev + x], scc = tab[T_SCC * g.lev + x];
  double sl, sr, sc_, d, q_, r, a2;
  if (MONO || d2h[0] * d2h[1] <= K0 || d2h[1] * d2h[2] <= K0) {
    sl = ssc * (hm0 - hmm);
    sr = ssc * (hmp - hm0);
    if (sl * sr > K0) {
      sc_ = scc * (hmp - hmm);
      sc_ = fsign(fmin(fmin(fabs(sl), fabs(sr)), fabs(sc_)), sc_);
      if ((hmm - hel) * (hm0 - hel) > K0) hel = hm0 - fsign(fmin(K1_2 * fabs(sc_), fabs(hel - hm0)), sc_);
      if ((hmp - her) * (hm0 - her) > K0) her = hm0 + fsign(fmin(K1_2 * fabs(sc_), fabs(her - hm0)), sc_);
      d = her - hel;
      q_ = d * (K2 * hm0 - hel - her);
      r = K1_3 * d * d;
      if (q_ > r) hel = K3 * hm0 - K2 * her;
      else if (-r > q_) her = K3 * hm0 - K2 * hel;
    } else {
      hel = hm0;
      her = hm0;
    }
  }
  if (!MONO) {
    hel = fmax(hel, DPEPS);
    her = fmax(her, DPEPS);
    sl = K2 * (K3 * hm0 - K2 * hel - her);
    a2 = K3 * (hel - K2 * hm0 + her);
    sr = sl + K2 * a2;
    if (sl < K0 && sr > K0) {
      if (a2 * hel - K1_4 * sl * sl < a2 * DPEPS) {
        q_ = K3 * hm0 / (K3 * sl * sr + K4 * a2 * a2);
        hel = sl * sl * q_;
        her = sr * sr * q_;
      }
    }
  }
  hel3[xk] = hel;
  her3[xk] = her;
}

// Level-marching tile form of the thickness edges.  A block owns TPO cells along the pass direction
// times TC across it and marches through a chunk of levels; per level the three stencil stages
// (cell thickness hm -> edge values he -> curvature proxy d2h) flow through shared memory, so every
// hm (one division on the second pass) and he is evaluated once per tile instead of 7/4 times per
// cell (the three d2h of a cell are three multiply-adds, redone per cell to save a barrier), and the
// static weights of a thread's own edge/cell stay in registers across the levels.
// dp (and the cross flux areas) of level k+1 are fetched into registers while level k is computed.
// Expressions and their order are those of cppm_hedges above, so both forms are bit-identical.
template <int DIR, bool MONO, int TPO, int TC, int MINB>
__global__ void __launch_bounds__(TPO* TC, MINB)
cppm_hedges_tile(Geom g, bool second_pass, int kchunk, const double* __restrict__ dp,
                 const double* __restrict__ cac, const double* __restrict__ scp2i,
                 const double* __restrict__ tab, double* __restrict__ hel3, double* __restrict__ her3) {
  constexpr int NP = TPO + 6;              // staged cells along the pass: s0-3 .. s0+TPO+2
  constexpr int NE = TPO + 3;              // edges s0-1 .. s0+TPO+1
  constexpr int NX = (NP + TPO - 1) / TPO; // staged cells per thread (2)
  // double buffered by level parity: a buffer is rewritten two levels later, after every thread has
  // passed a barrier that follows its last read, so two barriers per level are enough
  __shared__ double s_hm2[2][NP * TC], s_he2[2][NE * TC];
  const int tp = DIR == 0 ? threadIdx.x : threadIdx.y;
  const int tc = DIR == 0 ? threadIdx.y : threadIdx.x;
  auto sidx = [&](int p, int n) { return DIR == 0 ? tc * n + p : p * TC + tc; };
  const int npass = DIR == 0 ? g.ii : g.jj, ncross = DIR == 0 ? g.jj : g.ii;
  const int s0 = 1 + (DIR == 0 ? blockIdx.x : blockIdx.y) * TPO;
  const int cc = 1 + (DIR == 0 ? blockIdx.y : blockIdx.x) * TC + tc;
  const int ccl = min(cc, ncross);
  const long sc = DIR == 0 ? g.ldi : 1;
  const long lev = g.lev;
  // clamped addresses (positions beyond the last cell + 3 are never used by a stored result)
  auto addr = [&](int pi) -> long {
    const int pc = min(pi, npass + 3);
    return DIR == 0 ? ix2(g, pc, ccl) : ix2(g, ccl, pc);
  };
  const int k_first = blockIdx.z * kchunk + 1;
  const int k_last = min(g.kdm, k_first + kchunk - 1);

  // ---- per-thread invariants ----
  long ycell[NX]; double ai[NX];
#pragma unroll
  for (int r = 0; r < NX; ++r) {
    const int q = tp + r * TPO;
    ycell[r] = addr(s0 - 3 + min(q, NP - 1));
    ai[r] = second_pass ? scp2i[ycell[r]] : K0;
  }
  // own edge r = tp (cell index s0-1+tp)
  const long ye = addr(s0 - 1 + tp);
  const double w1 = tab[T_HEVC1 * lev + ye], w2 = tab[T_HEVC2 * lev + ye], w3 = tab[T_HEVC3 * lev + ye],
               w4 = tab[T_HEVC4 * lev + ye];
  // own cell c = s0+tp and the curvature masks of c-1, c, c+1
  const int c = s0 + tp;
  const long xc = addr(c);
  const double ssc = tab[T_SSC * lev + xc], scc = tab[T_SCC * lev + xc];
  const double d2m_l = tab[T_D2M * lev + ye], d2m_c = tab[T_D2M * lev + xc], d2m_r = tab[T_D2M * lev + addr(c + 1)];
  const bool store = c <= npass && cc <= ncross;

  double raw[NX], cp_[NX], cm_[NX];
  auto fetch = [&](int k) {
    const long koff = (long)(k - 1) * lev;
#pragma unroll
    for (int r = 0; r < NX; ++r) {
      if (tp + r * TPO < NP) {
        raw[r] = dp[ycell[r] + koff];
        if (second_pass) { cp_[r] = cac[ycell[r] + koff + sc]; cm_[r] = cac[ycell[r] + koff]; }
      }
    }
  };
  fetch(k_first);
  for (int k = k_first; k <= k_last; ++k) {
    double* s_hm = s_hm2[(k - k_first) & 1];
    double* s_he = s_he2[(k - k_first) & 1];
    // ---- cell thickness of the staged cells ----
#pragma unroll
    for (int r = 0; r < NX; ++r) {
      const int q = tp + r * TPO;
      if (q < NP) {
        double h = fmax(K0, raw[r]) + DPEPS;
        if (second_pass) h = h / (K1 - (cp_[r] - cm_[r]) * ai[r]);
        s_hm[sidx(q, NP)] = h;
      }
    }
    if (k < k_last) fetch(k + 1);
    __syncthreads();
    // ---- edge values ----
    s_he[sidx(tp, NE)] = w1 * s_hm[sidx(tp, NP)] + w2 * s_hm[sidx(tp + 1, NP)] + w3 * s_hm[sidx(tp + 2, NP)] +
                         w4 * s_hm[sidx(tp + 3, NP)];
    if (tp < NE - TPO) {
      const int r = tp + TPO;
      const long y = addr(s0 - 1 + r);
      s_he[sidx(r, NE)] = tab[T_HEVC1 * lev + y] * s_hm[sidx(r, NP)] + tab[T_HEVC2 * lev + y] * s_hm[sidx(r + 1, NP)] +
                          tab[T_HEVC3 * lev + y] * s_hm[sidx(r + 2, NP)] + tab[T_HEVC4 * lev + y] * s_hm[sidx(r + 3, NP)];
    }
    __syncthreads();
    // ---- curvature proxy of c-1, c, c+1 and limiter of the own cell ----
    const double e0 = s_he[sidx(tp, NE)], e3 = s_he[sidx(tp + 3, NE)];
    double hel = s_he[sidx(tp + 1, NE)], her = s_he[sidx(tp + 2, NE)];
    const double hmm = s_hm[sidx(tp + 2, NP)], hm0 = s_hm[sidx(tp + 3, NP)], hmp = s_hm[sidx(tp + 4, NP)];
    const double d2l = d2m_l * (e0 - K2 * hmm + hel);
    const double d2c = d2m_c * (hel - K2 * hm0 + her);
    const double d2r = d2m_r * (her - K2 * hmp + e3);
    double sl, sr, sc_, d, q_, r_, a2;
    if (MONO || d2l * d2c <= K0 || d2c * d2r <= K0) {
      sl = ssc * (hm0 - hmm);
      sr = ssc * (hmp - hm0);
      if (sl * sr > K0) {
        sc_ = scc * (hmp - hmm);
        sc_ = fsign(fmin(fmin(fabs(sl), fabs(sr)), fabs(sc_)), sc_);
        if ((hmm - hel) * (hm0 - hel) > K0) hel = hm0 - fsign(fmin(K1_2 * fabs(sc_), fabs(hel - hm0)), sc_);
        if ((hmp - her) * (hm0 - her) > K0) her = hm0 + fsign(fmin(K1_2 * fabs(sc_), fabs(her - hm0)), sc_);
        d = her - hel;
        q_ = d * (K2 * hm0 - hel - her);
        r_ = K1_3 * d * d;
        if (q_ > r_) hel = K3 * hm0 - K2 * her;
        else if (-r_ > q_) her = K3 * hm0 - K2 * hel;
      } else {
        hel = hm0;
        her = hm0;
      }
    }
    if (!MONO) {
      hel = fmax(hel, DPEPS);
      her = fmax(her, DPEPS);
      sl = K2 * (K3 * hm0 - K2 * hel - her);
      a2 = K3 * (hel - K2 * hm0 + her);
      sr = sl + K2 * a2;
      if (sl < K0 && sr > K0) {
        if (a2 * hel - K1_4 * sl * sl < a2 * DPEPS) {
          q_ = K3 * hm0 / (K3 * sl * sr + K4 * a2 * a2);
          hel = sl * sl * q_;
          her = sr * sr * q_;
        }
      }
    }
    if (store) {
      const long xk = xc + (long)(k - 1) * lev;
      hel3[xk] = hel;
      her3[xk] = her;
    }
  }
}

// arctic swap of hel/her after their halo update (mod_cppm.F90:1531-1541, :1686-1703)
template <int DIR>
__global__ void cppm_swap_edges(Geom g, bool fold_fix, int hw /* 4 nosc, 3 mono (:1848, :2015) */, double* hel3,
                                double* her3) {
  const int k = blockIdx.y;
  const int nrow = DIR == 0 ? 1 : 1 + hw;
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long)g.ldi * nrow) return;
  const int i = (int)(t % g.ldi) + 1 - g.nb, r = (int)(t / g.ldi);
  bool doit;
  if (DIR == 0) doit = (i >= 1 - hw && i <= g.ii + hw);
  // reference quirk (mod_cppm.F90:1690): only the right half of row jj is swapped unless
  // option cppm_fold_fix=1 asks for the whole (mirrored) row
  else doit = r == 0 ? (i >= (fold_fix ? 1 : max(1, g.itdm / 2 - g.i0 + 1)) && i <= g.ii)
                     : (i >= 1 && i <= g.ii);
  if (!doit) return;
  const long x = ix2(g, i, g.jj + r) + (long)k * g.lev;
  double a = hel3[x]; hel3[x] = her3[x]; her3[x] = a;
}

// ---- compatible tracer edge weights (per-interface LU, :519-722) --------------
// c0..c3 are the 4 cells e-2..e+1; t0/tl/tr are this interface's moment coefficients.
__device__ __forceinline__ void tracer_edge_weights(int stencil, const double* __restrict__ t0,
                                                    const double* __restrict__ tl, const double* __restrict__ tr,
                                                    const double hi[4] /* 1/hm of the four cells */,
                                                    const double hel[4],
                                                    const double her[4], double& tevc1, double& tevc2,
                                                    double& tevc3, double& tevc4) {
  double h1i, h2i, h3i, h4i, a12, a22, a32, a42, a13, a23, a33, a43, a14, a24, a34, a44, q;
#define EL(r, c, hi) (t0[(r) - 1] + (tl[(r) - 1] * hel[c] + tr[(r) - 1] * her[c]) * hi)
  switch (stencil) {
    case stencil_1111:
      h1i = hi[0]; h2i = hi[1]; h3i = hi[2]; h4i = hi[3];
      a12 = EL(1, 0, h1i); a13 = EL(2, 0, h1i); a14 = EL(3, 0, h1i);
      a22 = EL(4, 1, h2i) - a12; a23 = EL(5, 1, h2i) - a13; a24 = EL(6, 1, h2i) - a14;
      a32 = EL(7, 2, h3i) - a12; a33 = EL(8, 2, h3i) - a13; a34 = EL(9, 2, h3i) - a14;
      a42 = EL(10, 3, h4i) - a12; a43 = EL(11, 3, h4i) - a13; a44 = EL(12, 3, h4i) - a14;
      q = K1 / a22;
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
      tevc1 = K1 - tevc2 - tevc3 - tevc4;
      break;
    case stencil_1110:
      h1i = hi[0]; h2i = hi[1]; h3i = hi[2];
      a12 = EL(1, 0, h1i); a13 = EL(2, 0, h1i);
      a22 = EL(4, 1, h2i) - a12; a23 = EL(5, 1, h2i) - a13;
      a32 = EL(7, 2, h3i) - a12; a33 = EL(8, 2, h3i) - a13;
      a23 = a23 / a22;
      a33 = a33 - a23 * a32;
      tevc2 = -a12;
      tevc3 = -a13 - a23 * tevc2;
      tevc3 = tevc3 / a33;
      tevc2 = (tevc2 - a32 * tevc3) / a22;
      tevc1 = K1 - tevc2 - tevc3;
      tevc4 = K0;
      break;
    case stencil_0111:
      h2i = hi[1]; h3i = hi[2]; h4i = hi[3];
      a22 = EL(4, 1, h2i); a23 = EL(5, 1, h2i);
      a32 = EL(7, 2, h3i) - a22; a33 = EL(8, 2, h3i) - a23;
      a42 = EL(10, 3, h4i) - a22; a43 = EL(11, 3, h4i) - a23;
      a33 = a33 / a32;
      a43 = a43 - a33 * a42;
      tevc3 = -a22;
      tevc4 = -a23 - a33 * tevc3;
      tevc4 = tevc4 / a43;
      tevc3 = (tevc3 - a42 * tevc4) / a32;
      tevc2 = K1 - tevc3 - tevc4;
      tevc1 = K0;
      break;
    case stencil_1100:
      h1i = hi[0]; h2i = hi[1];
      a12 = EL(1, 0, h1i);
      a22 = EL(4, 1, h2i) - a12;
      tevc2 = -a12 / a22;
      tevc1 = K1 - tevc2;
      tevc3 = K0; tevc4 = K0;
      break;
    case stencil_0110:
      h2i = hi[1]; h3i = hi[2];
      a22 = EL(4, 1, h2i);
      a32 = EL(7, 2, h3i) - a22;
      tevc3 = -a22 / a32;
      tevc2 = K1 - tevc3;
      tevc1 = K0; tevc4 = K0;
      break;
    case stencil_0011:
      h3i = hi[2]; h4i = hi[3];
      a32 = EL(7, 2, h3i);
      a42 = EL(10, 3, h4i) - a32;
      tevc4 = -a32 / a42;
      tevc3 = K1 - tevc4;
      tevc1 = K0; tevc2 = K0;
      break;
    case stencil_0100:
      tevc1 = K0; tevc2 = K1; tevc3 = K0; tevc4 = K0;
      break;
    case stencil_0010:
      tevc1 = K0; tevc2 = K0; tevc3 = K1; tevc4 = K0;
      break;
    default:
      tevc1 = K0; tevc2 = K0; tevc3 = K0; tevc4 = K0;
      break;
  }
#undef EL
}

template <int NT>
struct ScalarPtrs {
  const double* src[NT];  // level-1 pointers of the source set (temp, saln, trc...)
  double* dst[NT];        // level-1 pointers of the destination set
};

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// ---- flux kernel ------------------------------------------------------------
// Thread block = TP positions along the pass direction x TC along the cross
// direction; it produces TP-5 updated cells per cross position and marches
// through a chunk of levels.  Everything that depends on (i,j) only (stencil
// tag, limiter coefficients, cell widths, 1/area, bottom pressure at the face)
// is loaded once per chunk; the level-dependent operands of level k+1 are
// fetched with cp.async into the second half of a double buffer while level k
// is computed, so DRAM latency is hidden by the pipeline instead of by
// occupancy.  Per level the stage data flows through shared memory:
//   cells (hm,hel,her,tm) -> A: edge values te -> B: curvature d2t
//   -> C: limited parabolas -> D: face fluxes -> E: cell update.
// The 36 moment coefficients of an interface are recomputed from the four
// cell widths (tm_coeffs) instead of being re-read per level; only interfaces
// on the tripolar fold rows, whose table entries are mirrored images
// (mod_cppm.F90:2605-2646), read the tables.
// Scheme variants (mod_cppm.F90:2748-2834): bit 0 = monotonic limiting, bit 1 = partial compatibility.
#define CPPM_J_TILE_DEFAULT "32x16"
enum { VAR_FC_NOSC = 0, VAR_FC_MONO = 1, VAR_PC_NOSC = 2, VAR_PC_MONO = 3 };

template <int NT, int TP, int VAR>
struct FluxSmem {
  static constexpr int NCELL = TP + 3;
  static constexpr int NRAW = 5 + NT;   // dp, hel, her, cross flux area (+,-), tm[NT]
  static constexpr int NOPS = 5;        // pass flux area, p(k+1), flx, tflx, sflx
  static constexpr int BUF = NRAW * NCELL + NOPS * TP;
  // hm, 1/hm, 1/area, width, E|F (1+NT rows), D (NT, +1 for the thickness curvature of the partial
  // compatibility variants), P (3+3NT)
  static constexpr int ND = NT + ((VAR & 2) ? 1 : 0);
  static constexpr int WORK = 4 * NCELL + (1 + NT + ND + 3 + 3 * NT) * TP;
  // odd stride between the cross positions: on the j pass the lanes of a warp run across tc, and an
  // even stride folds them onto half of the shared-memory banks (measured: 7.6 -> 8.6 ms)
  static constexpr int PER_TC = (2 * BUF + WORK) | 1;
};

// slope limiter + parabola monotonicity fix of the partial-compatibility routines (thickness and
// tracers alike, mod_cppm.F90:1168-1190, :1209-1232, :1309-1358)
__device__ __forceinline__ void pc_limit(double ssc, double scc, double xm, double x0, double xp, double& el,
                                         double& er) {
  const double sl = ssc * (x0 - xm), sr = ssc * (xp - x0);
  if (sl * sr > K0) {
    double scv = scc * (xp - xm);
    scv = fsign(fmin(fmin(fabs(sl), fabs(sr)), fabs(scv)), scv);
    if ((xm - el) * (x0 - el) > K0) el = x0 - fsign(fmin(K1_2 * fabs(scv), fabs(el - x0)), scv);
    if ((xp - er) * (x0 - er) > K0) er = x0 + fsign(fmin(K1_2 * fabs(scv), fabs(er - x0)), scv);
    const double d = er - el;
    const double q = d * (K2 * x0 - el - er);
    const double r = K1_3 * d * d;
    if (q > r) el = K3 * x0 - K2 * er;
    else if (-r > q) er = K3 * x0 - K2 * el;
  } else {
    el = x0;
    er = x0;
  }
}

template <int DIR, int NT, int TP, int TC, int VAR>
__global__ void __launch_bounds__(TP* TC, TP* TC <= 256 ? 2 : 1)
cppm_flux(Geom g, bool second_pass, int n_lev2d /* level (1-based) of pbu/pbv */, int kchunk,
          const double* __restrict__ dp_src, double* __restrict__ dp_dst, ScalarPtrs<NT> S,
          const double* __restrict__ hel3, const double* __restrict__ her3,
          const double* __restrict__ cad /* pass-direction flux area */,
          const double* __restrict__ cac /* cross-direction flux area */,
          const double* __restrict__ p, const double* __restrict__ pbd /* pbu or pbv */,
          const double* __restrict__ scp2i, const double* __restrict__ scpd /* scpx or scpy */,
          const double* __restrict__ tab, const int* __restrict__ sten, double* __restrict__ flx,
          double* __restrict__ tflx, double* __restrict__ sflx /* pointers at level 1+mm */,
          int extra /* 1: the scalars are a further group of passive tracers transported with the thickness
                       fluxes of the pass; thickness and the flux accumulators were written by the first group */) {
  using L = FluxSmem<NT, TP, VAR>;
  constexpr int NCELL = L::NCELL;
  constexpr bool MONO = (VAR & 1) != 0, PC = (VAR & 2) != 0;
  // 1/hm of a cell is shared by the four interfaces that use it on the i pass (5.86 -> 5.62 ms); on the
  // j pass (32-row tiles) the extra division in the three-row tail of the staging loop sits on the
  // critical path before a barrier and costs more than it saves (7.6 -> 8.5 ms), so each interface
  // keeps its own four reciprocals there.  Same values either way.
  constexpr bool SHARE_HI = !PC && DIR == 0;
  extern __shared__ double smem[];
  const int tp = DIR == 0 ? threadIdx.x : threadIdx.y;
  const int tc = DIR == 0 ? threadIdx.y : threadIdx.x;
  double* base = smem + (long)tc * L::PER_TC;
  double* bufs = base;                       // [2][BUF]
  double* s_hm = base + 2 * L::BUF;          // [NCELL]
  double* s_ai = s_hm + NCELL;               // [NCELL] 1/area of the staged cells
  double* s_dx = s_ai + NCELL;               // [NCELL] width of the staged cells along the pass
  double* s_hi = s_dx + NCELL;               // [NCELL] 1/hm, shared by the four interfaces that use a cell
  double* s_F = s_hi + NCELL;                // [1+NT][TP]; rows 1.. double as the edge values E
  double* s_E = s_F + TP;
  double* s_D = s_F + (1 + NT) * TP;         // [ND][TP]; row NT = thickness curvature (PC)
  double* s_P = s_D + L::ND * TP;            // [3+3NT][TP]

  constexpr int NOUT = TP - 5;
  const int npass = DIR == 0 ? g.idm : g.jdm;
  const int ncross = DIR == 0 ? g.jdm : g.idm;
  const int tile = DIR == 0 ? blockIdx.x : blockIdx.y;
  const int ctile = DIR == 0 ? blockIdx.y : blockIdx.x;
  const int k_first = blockIdx.z * kchunk + 1;
  const int k_last = min(g.kdm, k_first + kchunk - 1);
  const int s0 = 1 + tile * NOUT;   // first updated cell of this tile
  const int p0 = s0 - 2;            // pass index of thread tp=0
  const int cc = 1 + ctile * TC + tc;
  const bool cvalid = cc <= ncross;
  const int ccl = min(cc, ncross);
  const long sc = DIR == 0 ? g.ldi : 1;
  const int pmax = npass + g.nb;    // last addressable pass index
  auto addr = [&](int pi) -> long { return DIR == 0 ? ix2(g, pi, ccl) : ix2(g, ccl, pi); };
  const long lev = g.lev;

  // own cell / edge / face index
  const int e = p0 + tp;
  const int el = min(e, pmax);
  const long xe = addr(el);
  const int qc = tp + 2;  // smem index of own cell
  const int tm1 = max(tp - 1, 0), tp1 = min(tp + 1, TP - 1);

  // issue the level-dependent loads of level k into buffer b
  auto issue = [&](int k, int b) {
    double* B = bufs + b * L::BUF;
    const long koff = (long)(k - 1) * lev;
    for (int q = tp; q < NCELL; q += TP) {
      const long yk = addr(min(p0 - 2 + q, pmax)) + koff;
      cp_async8(B + 0 * NCELL + q, dp_src + yk);
      if (!PC) {
        cp_async8(B + 1 * NCELL + q, hel3 + yk);
        cp_async8(B + 2 * NCELL + q, her3 + yk);
      }
      if (second_pass) {
        cp_async8(B + 3 * NCELL + q, cac + yk + sc);
        cp_async8(B + 4 * NCELL + q, cac + yk);
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) cp_async8(B + (5 + nt) * NCELL + q, S.src[nt] + yk);
    }
    double* O = B + L::NRAW * NCELL;
    cp_async8(O + 0 * TP + tp, cad + xe + koff);
    cp_async8(O + 1 * TP + tp, p + xe + koff + lev);
    cp_async8(O + 2 * TP + tp, flx + xe + koff);
    cp_async8(O + 3 * TP + tp, tflx + xe + koff);
    cp_async8(O + 4 * TP + tp, sflx + xe + koff);
    cp_async_commit();
  };
  issue(k_first, 0);

  // ---- per-(i,j) invariants of the chunk ----
  for (int q = tp; q < NCELL; q += TP) {
    const long y = addr(min(p0 - 2 + q, pmax));
    s_ai[q] = scp2i[y];
    s_dx[q] = scpd[y];
  }
  const int stencil = sten[xe];
  const double d2m = tab[T_D2M * lev + xe], ssc = tab[T_SSC * lev + xe], scc = tab[T_SCC * lev + xe];
  const double db = pbd[xe + (long)(n_lev2d - 1) * lev];
  double hv1 = K0, hv2 = K0, hv3 = K0, hv4 = K0;   // thickness edge weights double as tracer weights (PC)
  if (PC) {
    hv1 = tab[T_HEVC1 * lev + xe]; hv2 = tab[T_HEVC2 * lev + xe];
    hv3 = tab[T_HEVC3 * lev + xe]; hv4 = tab[T_HEVC4 * lev + xe];
  }
  // interfaces on the fold rows carry mirrored table entries: read them instead of recomputing
  // (xctilr rewrites row jj of u-type tables with the mirror of row jj-1 even for nh=0, so the
  //  i-pass needs them on row jj; the j-pass on rows >= jj)
  const bool use_tab = !PC && g.nreg == 2 && g.north && (DIR == 1 ? e >= g.jj : cc == g.jj);
  double p_own = p[xe + (long)(k_first - 1) * lev];
  double p_up = p[addr(min(max(e - 1, 1 - g.nb), pmax)) + (long)(k_first - 1) * lev];
  const bool face_ok = cvalid && tp >= 2 && tp <= TP - 3 && e >= 1 && e <= npass + 1 &&
                       (tp <= TP - 4 || e == npass + 1);
  const bool cell_ok = cvalid && tp >= 2 && tp <= TP - 4 && e >= 1 && e <= npass;

  for (int k = k_first; k <= k_last; ++k) {
    const int b = (k - k_first) & 1;
    const double* B = bufs + b * L::BUF;
    const double* s_dp = B;
    const double* s_hel = B + NCELL;
    const double* s_her = B + 2 * NCELL;
    const double* s_tm = B + 5 * NCELL;       // [NT][NCELL]
    const double* O = B + L::NRAW * NCELL;
    cp_async_wait_all();
    __syncthreads();                           // level k landed; everybody is done with level k-1
    if (k < k_last) issue(k + 1, b ^ 1);

    // ---- cell mean thickness of the staged cells ----
    for (int q = tp; q < NCELL; q += TP) {
      double h = fmax(K0, s_dp[q]) + DPEPS;
      if (second_pass) h = h / (K1 - (B[3 * NCELL + q] - B[4 * NCELL + q]) * s_ai[q]);
      s_hm[q] = h;
      if (SHARE_HI) s_hi[q] = K1 / h;
    }
    __syncthreads();

    const double hm_c = s_hm[qc];
    double hel_c, her_c;
    if (!PC) { hel_c = s_hel[qc]; her_c = s_her[qc]; }
    double tm_c[NT];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) tm_c[nt] = s_tm[nt * NCELL + qc];

    // ---- A: tracer edge values at edge e from cells e-2..e+1 (smem tp..tp+3) ----
    if (PC) {
      // partial compatibility (:1143-1153): thickness and tracer edges share the static weights
      s_F[tp] = hv1 * s_hm[tp] + hv2 * s_hm[tp + 1] + hv3 * s_hm[tp + 2] + hv4 * s_hm[tp + 3];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
        s_E[nt * TP + tp] = hv1 * s_tm[nt * NCELL + tp] + hv2 * s_tm[nt * NCELL + tp + 1] +
                            hv3 * s_tm[nt * NCELL + tp + 2] + hv4 * s_tm[nt * NCELL + tp + 3];
    } else {
      double hi4[4], hel4[4], her4[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        hi4[q] = SHARE_HI ? s_hi[tp + q] : K1 / s_hm[tp + q];
        hel4[q] = s_hel[tp + q]; her4[q] = s_her[tp + q];
      }
      double w1, w2, w3, w4;
      TmCoef tcf;
      if (use_tab) {
#pragma unroll
        for (int r = 0; r < 12; ++r) {
          tcf.t0[r] = tab[(T_TMC0 + r) * lev + xe];
          tcf.tl[r] = tab[(T_TMCL + r) * lev + xe];
          tcf.tr[r] = tab[(T_TMCR + r) * lev + xe];
        }
      } else {
        double av[12];
        tm_coeffs(s_dx[tp], s_dx[tp + 1], s_dx[tp + 2], s_dx[tp + 3], tcf, av);
      }
      tracer_edge_weights(stencil, tcf.t0, tcf.tl, tcf.tr, hi4, hel4, her4, w1, w2, w3, w4);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
        s_E[nt * TP + tp] = w1 * s_tm[nt * NCELL + tp] + w2 * s_tm[nt * NCELL + tp + 1] +
                            w3 * s_tm[nt * NCELL + tp + 2] + w4 * s_tm[nt * NCELL + tp + 3];
    }
    __syncthreads();

    // ---- B: thickness factors and curvature proxy of own cell ----
    double tel[NT], ter[NT];
    double hf1m, hf1l, hf1r, hf2m, hf2l, hf2r;
    if (PC) {
      hel_c = s_F[tp]; her_c = s_F[tp1];
      if (!MONO) s_D[NT * TP + tp] = d2m * (hel_c - K2 * hm_c + her_c);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        tel[nt] = s_E[nt * TP + tp];
        ter[nt] = s_E[nt * TP + tp1];
        if (!MONO) s_D[nt * TP + tp] = d2m * (tel[nt] - K2 * tm_c[nt] + ter[nt]);
      }
    } else {
      const double q = K1 / (K12 * hm_c - hel_c - her_c);
      hf1m = K60 * hm_c * q;
      hf1l = -(K42 * hm_c + K4 * hel_c - K6 * her_c) * q;
      hf1r = -(K18 * hm_c - K4 * hel_c + K6 * her_c) * q;
      hf2m = -hf1m;
      hf2l = K5 * (K6 * hm_c + hel_c - her_c) * q;
      hf2r = K5 * (K6 * hm_c - hel_c + her_c) * q;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        tel[nt] = s_E[nt * TP + tp];
        ter[nt] = s_E[nt * TP + tp1];
        if (!MONO) s_D[nt * TP + tp] = d2m * (hf2m * tm_c[nt] + hf2l * tel[nt] + hf2r * ter[nt]);
      }
    }
    if (!MONO) __syncthreads();   // (uniform: MONO is a template constant)

    // ---- C: limiters and parabola coefficients of own cell ----
    double hpc0, hpc1, hpc2, tpc0[NT], tpc1[NT], tpc2[NT];
    if (PC) {
      // :1166-1263 (pc_nosc) / :1307-1369 (pc_mono)
      const double hmm = s_hm[qc - 1], hmp = s_hm[qc + 1];
      bool lim = MONO;
      if (!MONO) {
        const double d2c = s_D[NT * TP + tp], d2l = s_D[NT * TP + tm1], d2r = s_D[NT * TP + tp1];
        lim = d2l * d2c <= K0 || d2c * d2r <= K0;
      }
      if (lim) pc_limit(ssc, scc, hmm, hm_c, hmp, hel_c, her_c);
      if (!MONO) {
        hel_c = fmax(hel_c, DPEPS);
        her_c = fmax(her_c, DPEPS);
        const double sl = K2 * (K3 * hm_c - K2 * hel_c - her_c);
        const double a2 = K3 * (hel_c - K2 * hm_c + her_c);
        const double sr = sl + K2 * a2;
        if (sl < K0 && sr > K0) {
          if (a2 * hel_c - K1_4 * sl * sl < a2 * DPEPS) {
            const double q = K3 * hm_c / (K3 * sl * sr + K4 * a2 * a2);
            hel_c = sl * sl * q;
            her_c = sr * sr * q;
          }
        }
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const double tmm = s_tm[nt * NCELL + qc - 1], tmp = s_tm[nt * NCELL + qc + 1], tmc = tm_c[nt];
        bool tl_ = MONO;
        if (!MONO) {
          const double d2c = s_D[nt * TP + tp], d2l = s_D[nt * TP + tm1], d2r = s_D[nt * TP + tp1];
          tl_ = d2l * d2c <= K0 || d2c * d2r <= K0;
        }
        if (tl_) pc_limit(ssc, scc, tmm, tmc, tmp, tel[nt], ter[nt]);
        if (!MONO && nt >= 1) {  // positivity (:1234-1248)
          tel[nt] = fmax(tel[nt], K0);
          ter[nt] = fmax(ter[nt], K0);
          const double sl = K2 * (K3 * tmc - K2 * tel[nt] - ter[nt]);
          const double a2 = K3 * (tel[nt] - K2 * tmc + ter[nt]);
          const double sr = sl + K2 * a2;
          if (sl < K0 && sr > K0) {
            if (a2 * tel[nt] - K1_4 * sl * sl < K0) {
              const double q = K3 * tmc / (K3 * sl * sr + K4 * a2 * a2);
              tel[nt] = sl * sl * q;
              ter[nt] = sr * sr * q;
            }
          }
        }
        tpc0[nt] = tel[nt];
        tpc1[nt] = K6 * tmc - K4 * tel[nt] - K2 * ter[nt];
        tpc2[nt] = K3 * (tel[nt] - K2 * tmc + ter[nt]);
        s_P[(3 + 3 * nt + 0) * TP + tp] = tpc0[nt];
        s_P[(3 + 3 * nt + 1) * TP + tp] = tpc1[nt];
        s_P[(3 + 3 * nt + 2) * TP + tp] = tpc2[nt];
      }
      hpc0 = hel_c;
      hpc1 = K6 * hm_c - K4 * hel_c - K2 * her_c;
      hpc2 = K3 * (hel_c - K2 * hm_c + her_c);
      s_P[0 * TP + tp] = hpc0; s_P[1 * TP + tp] = hpc1; s_P[2 * TP + tp] = hpc2;
    } else {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const double tmm = s_tm[nt * NCELL + qc - 1], tmp = s_tm[nt * NCELL + qc + 1], tmc = tm_c[nt];
        double sl, sr, scv, a2;
        bool lim = MONO;   // fc_mono limits every cell (:1073-1099)
        if (!MONO) {
          const double d2c = s_D[nt * TP + tp], d2l = s_D[nt * TP + tm1], d2r = s_D[nt * TP + tp1];
          lim = d2l * d2c <= K0 || d2c * d2r <= K0;
        }
        if (lim) {
          sl = ssc * (tmc - tmm);
          sr = ssc * (tmp - tmc);
          if (sl * sr > K0) {
            scv = scc * (tmp - tmm);
            scv = fsign(fmin(fmin(fabs(sl), fabs(sr)), fabs(scv)), scv);
            if ((tmm - tel[nt]) * (tmc - tel[nt]) > K0)
              tel[nt] = tmc - fsign(fmin(K1_2 * fabs(scv), fabs(tel[nt] - tmc)), scv);
            if ((tmp - ter[nt]) * (tmc - ter[nt]) > K0)
              ter[nt] = tmc + fsign(fmin(K1_2 * fabs(scv), fabs(ter[nt] - tmc)), scv);
            sl = hf1m * tmc + hf1l * tel[nt] + hf1r * ter[nt];
            a2 = hf2m * tmc + hf2l * tel[nt] + hf2r * ter[nt];
            sr = sl + K2 * a2;
            if (sl * sr < K0) {
              if ((ter[nt] - tel[nt]) * a2 < K0)
                tel[nt] = -((hf1m + K2 * hf2m) * tmc + (hf1r + K2 * hf2r) * ter[nt]) / (hf1l + K2 * hf2l);
              else
                ter[nt] = -(hf1m * tmc + hf1l * tel[nt]) / hf1r;
            }
          } else {
            tel[nt] = tmc;
            ter[nt] = tmc;
          }
        }
        if (!MONO && nt >= 1) {  // positivity for everything but temperature (:788-801)
          tel[nt] = fmax(tel[nt], K0);
          ter[nt] = fmax(ter[nt], K0);
          sl = hf1m * tmc + hf1l * tel[nt] + hf1r * ter[nt];
          a2 = hf2m * tmc + hf2l * tel[nt] + hf2r * ter[nt];
          sr = sl + K2 * a2;
          if (sl < K0 && sr > K0) {
            if (a2 * tel[nt] - K1_4 * sl * sl < K0) {
              const double q = K3 * tmc / (K3 * sl * sr + K4 * a2 * a2);
              tel[nt] = sl * sl * q;
              ter[nt] = sr * sr * q;
            }
          }
        }
        tpc0[nt] = tel[nt];
        tpc1[nt] = hf1m * tmc + hf1l * tel[nt] + hf1r * ter[nt];
        tpc2[nt] = hf2m * tmc + hf2l * tel[nt] + hf2r * ter[nt];
        s_P[(3 + 3 * nt + 0) * TP + tp] = tpc0[nt];
        s_P[(3 + 3 * nt + 1) * TP + tp] = tpc1[nt];
        s_P[(3 + 3 * nt + 2) * TP + tp] = tpc2[nt];
      }
      hpc0 = hel_c;
      hpc1 = K6 * hm_c - K4 * hel_c - K2 * her_c;
      hpc2 = K3 * (hel_c - K2 * hm_c + her_c);
      s_P[0 * TP + tp] = hpc0; s_P[1 * TP + tp] = hpc1; s_P[2 * TP + tp] = hpc2;
    }
    __syncthreads();

    // ---- D: flux through face e (flux_integration, :1373-1468) ----
    const double ai_c = s_ai[qc];
    const double dl_own = O[1 * TP + tp], dl_up = O[1 * TP + tm1];
    double hf, htf[NT];
    {
      const double ca = O[0 * TP + tp];
      if (ca < K0) {
        const double c = ca * ai_c;
        const double du = p_own, dl = dl_own;
        double p0_, p1_, p2_;
        if (dl > db) {
          const double hb = fmax(K0, db - du);
          hf = hb * ca;
          p0_ = hb;
          p1_ = -K1_2 * hb * c;
          p2_ = K1_3 * hb * c * c;
        } else {
          hf = (hpc0 - (K1_2 * hpc1 - K1_3 * hpc2 * c) * c) * ca;
          p0_ = hpc0 - (K1_2 * hpc1 - K1_3 * hpc2 * c) * c;
          p1_ = -(K1_2 * hpc0 - (K1_3 * hpc1 - K1_4 * hpc2 * c) * c) * c;
          p2_ = (K1_3 * hpc0 - (K1_4 * hpc1 - K1_5 * hpc2 * c) * c) * c * c;
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) htf[nt] = (p0_ * tpc0[nt] + p1_ * tpc1[nt] + p2_ * tpc2[nt]) * ca;
      } else {
        const double c = ca * s_ai[qc - 1];  // upstream cell e-1
        const double q1 = K1 - K1_2 * c;
        const double q2 = K1 - (K1 - K1_3 * c) * c;
        const double du = p_up, dl = dl_up;
        const double u0 = s_P[0 * TP + tm1], u1 = s_P[1 * TP + tm1], u2 = s_P[2 * TP + tm1];
        double p0_, p1_, p2_;
        if (dl > db) {
          const double hb = fmax(K0, db - du);
          hf = hb * ca;
          p0_ = hb;
          p1_ = q1 * hb;
          p2_ = q2 * hb;
        } else {
          hf = (u0 + q1 * u1 + q2 * u2) * ca;
          const double q3 = K1_4 * (K1 + K3 * (K1 - c) * q2);
          const double q4 = K1_5 * (K1 + K4 * (K1 - c) * q3);
          p0_ = u0 + q1 * u1 + q2 * u2;
          p1_ = q1 * u0 + q2 * u1 + q3 * u2;
          p2_ = q2 * u0 + q3 * u1 + q4 * u2;
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
          htf[nt] = (p0_ * s_P[(3 + 3 * nt + 0) * TP + tm1] + p1_ * s_P[(3 + 3 * nt + 1) * TP + tm1] +
                     p2_ * s_P[(3 + 3 * nt + 2) * TP + tm1]) * ca;
      }
      s_F[tp] = hf;                             // (E rows are dead since stage B)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) s_F[(1 + nt) * TP + tp] = htf[nt];
    }
    p_own = dl_own; p_up = dl_up;               // p(k+1) is the upper interface of the next level
    __syncthreads();

    // ---- E: divergence update of own cell + flux accumulation at own face ----
    const long xek = xe + (long)(k - 1) * lev;
    if (cell_ok) {
      const double ho = fmax(K0, s_dp[qc]) + DPEPS;
      const double hn = ho - (s_F[tp + 1] - hf) * ai_c;
      const double hni = K1 / hn;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
        S.dst[nt][xek] = (ho * tm_c[nt] - (s_F[(1 + nt) * TP + tp + 1] - htf[nt]) * ai_c) * hni;
      if (!extra) dp_dst[xek] = fmax(K0, hn - DPEPS);
    }
    // faces s0..s0+NOUT-1 belong to this tile; the last tile also owns face npass+1
    if (face_ok && !extra) {
      flx[xek] = O[2 * TP + tp] + hf;
      tflx[xek] = O[3 * TP + tp] + htf[0];
      sflx[xek] = O[4 * TP + tp] + htf[1];
    }
  }
}

template <int DIR, int NT, int VAR, int TP, int TC>
void launch_flux_shape(bool second_pass, int n, const double* dp_src, double* dp_dst, const ScalarPtrs<NT>& S,
                 const double* hel3, const double* her3, const double* cad, const double* cac,
                 const double* p, const double* pbd, const double* scp2i, const double* scpd, const double* tab,
                 const int* sten, double* flx, double* tflx, double* sflx, int extra) {
  Ctx& c = C(); const Geom& g = c.g;
  constexpr int NOUT = TP - 5;
  const size_t smem = sizeof(double) * FluxSmem<NT, TP, VAR>::PER_TC * TC;
  auto kern = cppm_flux<DIR, NT, TP, TC, VAR>;
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  dim3 block, grid;
  if (DIR == 0) {
    block = dim3(TP, TC);
    grid = dim3(cdiv(g.idm, NOUT), cdiv(g.jdm, TC), 1);
  } else {
    block = dim3(TC, TP);
    grid = dim3(cdiv(g.idm, TC), cdiv(g.jdm, NOUT), 1);
  }
  // level chunks: enough blocks for ~4 waves of the 148 SMs, but at least 6 levels per chunk so the
  // per-chunk invariant loads stay amortised
  const long tiles = (long)grid.x * grid.y;
  int nz = (int)std::min<long>(std::max<long>(1, (148L * 2 * 4 + tiles - 1) / tiles), std::max(1, g.kdm / 6));
  const int kchunk = cdiv(g.kdm, nz);
  grid.z = cdiv(g.kdm, kchunk);
  // NB: the face npass+1 must be covered: tiles cover cells 1..ntile*NOUT >= npass and
  // thread tp = npass+1-p0 <= TP-3 of the last tile owns it.
  LAUNCH_NAMED(DIR == 0 ? "cppm_flux<i>" : "cppm_flux<j>", kern, grid, block, smem, g, second_pass, n, kchunk, dp_src,
               dp_dst, S, hel3, her3, cad, cac, p, pbd, scp2i, scpd, tab, sten, flx, tflx, sflx, extra);
}

// Tile shape of the flux kernel: TP positions along the pass (TP-5 of them updated) x TC across it.
// i pass: 128 x 2.  j pass: lanes must run along i, so the tile is TC wide in i and TP long in j;
// 32 x 16 recomputes 5 of 32 rows, 64 x 8 only 5 of 64, 32 x 8 gives two resident blocks of 256
// threads instead of one of 512 (development switch cppm_j_tile).
template <int DIR, int NT, int VAR>
void launch_flux(bool second_pass, int n, const double* dp_src, double* dp_dst, const ScalarPtrs<NT>& S,
                 const double* hel3, const double* her3, const double* cad, const double* cac,
                 const double* p, const double* pbd, const double* scp2i, const double* scpd, const double* tab,
                 const int* sten, double* flx, double* tflx, double* sflx, int extra = 0) {
  if constexpr (DIR == 0)
    launch_flux_shape<DIR, NT, VAR, 128, 2>(second_pass, n, dp_src, dp_dst, S, hel3, her3, cad, cac, p, pbd, scp2i,
                                            scpd, tab, sten, flx, tflx, sflx, extra);
  else {
    const std::string jt = C().option("cppm_j_tile", CPPM_J_TILE_DEFAULT);
    // with two passive tracers (NT = 4) a 16-wide tile needs 232 KB of shared memory, more than the
    // 227 KB a block may have: the 8-wide tile (116 KB) is used instead
    if (NT >= 4 || jt == "32x8")
      launch_flux_shape<DIR, NT, VAR, 32, 8>(second_pass, n, dp_src, dp_dst, S, hel3, her3, cad, cac, p, pbd, scp2i,
                                             scpd, tab, sten, flx, tflx, sflx, extra);
    else if (jt == "64x8")
      launch_flux_shape<DIR, NT, VAR, 64, 8>(second_pass, n, dp_src, dp_dst, S, hel3, her3, cad, cac, p, pbd, scp2i,
                                             scpd, tab, sten, flx, tflx, sflx, extra);
    else
      launch_flux_shape<DIR, NT, VAR, 32, 16>(second_pass, n, dp_src, dp_dst, S, hel3, her3, cad, cac, p, pbd,
                                              scp2i, scpd, tab, sten, flx, tflx, sflx, extra);
  }
}

template <int DIR, int NT, int VAR>
void cppm_pass(bool second_pass, int n, int mm, double* dp_src, double* dp_dst, ScalarPtrs<NT> S) {
  Ctx& c = C(); const Geom& g = c.g;
  constexpr bool MONO = (VAR & 1) != 0, PC = (VAR & 2) != 0;
  constexpr int hw = MONO ? 3 : 4;   // halo width of the variant (:1485 vs :1802)
  const int mh = DIR == 0 ? hw : 0, nh = DIR == 0 ? 0 : hw;
  // halo of the transported fields in the pass direction (:1485-1490 / :1640-1645)
  std::vector<HaloReq> reqs{{dp_src, g.kdm, halo_ps}};
  for (int nt = 0; nt < NT; ++nt) reqs.push_back({const_cast<double*>(S.src[nt]), g.kdm, halo_ps});
  halo_update(reqs, mh, nh);
  double* hel3 = c.owned("cppm_hel_3d", g.kdm);
  double* her3 = c.owned("cppm_her_3d", g.kdm);
  const double* tab = c.dev(DIR == 0 ? "cppm_tab_i" : "cppm_tab_j");
  const int* sten = c.idev(DIR == 0 ? "cppm_sten_i" : "cppm_sten_j");
  const double* cad = c.dev(DIR == 0 ? "cau" : "cav");
  const double* cac = c.dev(DIR == 0 ? "cav" : "cau");
  const double* scp2i = c.dev("scp2i");
  if (!PC) {   // full compatibility stages the limited thickness edges (:1493-1541)
    if (c.option("hedges_form", "tile") == "flat") {
      dim3 grid(cdiv(g.ii, 128), g.jj, g.kdm);
      auto hk = cppm_hedges<DIR, MONO>;
      LAUNCH_NAMED(DIR == 0 ? "cppm_hedges<i>" : "cppm_hedges<j>", hk, grid, 128, 0, g,
                   second_pass, dp_src, cac, scp2i, tab, hel3, her3);
    } else {
      constexpr int TPO = DIR == 0 ? 128 : 16, TC = DIR == 0 ? 2 : 32;
      dim3 block = DIR == 0 ? dim3(TPO, TC) : dim3(TC, TPO);
      dim3 grid = DIR == 0 ? dim3(cdiv(g.ii, TPO), cdiv(g.jj, TC), 1) : dim3(cdiv(g.ii, TC), cdiv(g.jj, TPO), 1);
      // level chunks: ~8 waves of the 148 SMs when the tile count alone does not provide them
      const long tiles = (long)grid.x * grid.y;
      const int nz = (int)std::min<long>(std::max<long>(1, (148L * 8 * 8 + tiles - 1) / tiles), std::max(1, g.kdm / 4));
      const int kchunk = cdiv(g.kdm, nz);
      grid.z = cdiv(g.kdm, kchunk);
      // resident blocks asked for: 256-thread blocks on the i pass (3 natural, 5, 8), 512-thread blocks on
      // the j pass (2 natural, 3, 4)
      OCC_DISPATCH3("hedges_minblk", 0, 0, 1, 2,
                    LAUNCH_NAMED(DIR == 0 ? "cppm_hedges<i>" : "cppm_hedges<j>",
                                 (cppm_hedges_tile<DIR, MONO, TPO, TC, DIR == 0 ? (OCC == 0 ? 3 : OCC == 1 ? 5 : 8) : OCC + 2>), grid,
                                 block, 0, g, second_pass, kchunk, dp_src, cac, scp2i, tab, hel3, her3));
    }
    halo_update(std::vector<HaloReq>{{hel3, g.kdm, halo_ps}, {her3, g.kdm, halo_ps}}, mh, nh);
    if (g.nreg == 2 && g.north) {
      const int nrow = DIR == 0 ? 1 : 1 + hw;
      dim3 grid2(cdiv((long)g.ldi * nrow, 256), g.kdm);
      LAUNCH(cppm_swap_edges<DIR>, grid2, 256, 0, g, c.option("cppm_fold_fix", "0") == "1", hw, hel3, her3);
    }
  }
  const long om = (long)mm * g.lev;
  launch_flux<DIR, NT, VAR>(second_pass, n, dp_src, dp_dst, S, hel3, her3, cad, cac, c.dev("p"),
                            c.dev(DIR == 0 ? "pbu" : "pbv"), scp2i, c.dev(DIR == 0 ? "scpx" : "scpy"), tab, sten,
                            c.dev(DIR == 0 ? "uflx" : "vflx") + om, c.dev(DIR == 0 ? "utflx" : "vtflx") + om,
                            c.dev(DIR == 0 ? "usflx" : "vsflx") + om);
}

// Further passive tracers of a pass, two per launch: the same thickness edges, flux areas and face fluxes as the
// first group (hel3/her3 and dp_src are still those of this pass), only the tracer columns change.  The reference
// loops nt = 1..ntr inside one sweep (phy/mod_cppm.F90:1599-1618); per tracer the operations are the same, so the
// result does not depend on the grouping.  `ex` holds (source, destination) level-1 pointers of the extra tracers;
// an odd one out is paired with itself (both copies write the same values).
template <int DIR, int VAR>
void cppm_pass_extra(bool second_pass, int n, int mm, const double* dp_src, double* dp_dst,
                     const std::vector<std::pair<double*, double*>>& ex) {
  if (ex.empty()) return;
  Ctx& c = C(); const Geom& g = c.g;
  constexpr int hw = (VAR & 1) ? 3 : 4;
  const int mh = DIR == 0 ? hw : 0, nh = DIR == 0 ? 0 : hw;
  std::vector<HaloReq> reqs;
  for (auto& e : ex) reqs.push_back({e.first, g.kdm, halo_ps});
  halo_update(reqs, mh, nh);
  const long om = (long)mm * g.lev;
  for (size_t q = 0; q < ex.size(); q += 2) {
    const size_t q1 = std::min(q + 1, ex.size() - 1);
    ScalarPtrs<2> S{};
    S.src[0] = ex[q].first; S.dst[0] = ex[q].second; S.src[1] = ex[q1].first; S.dst[1] = ex[q1].second;
    launch_flux<DIR, 2, VAR>(second_pass, n, dp_src, dp_dst, S, c.owned("cppm_hel_3d", g.kdm), c.owned("cppm_her_3d", g.kdm),
                             c.dev(DIR == 0 ? "cau" : "cav"), c.dev(DIR == 0 ? "cav" : "cau"), c.dev("p"),
                             c.dev(DIR == 0 ? "pbu" : "pbv"), c.dev("scp2i"), c.dev(DIR == 0 ? "scpx" : "scpy"),
                             c.dev(DIR == 0 ? "cppm_tab_i" : "cppm_tab_j"), c.idev(DIR == 0 ? "cppm_sten_i" : "cppm_sten_j"),
                             c.dev(DIR == 0 ? "uflx" : "vflx") + om, c.dev(DIR == 0 ? "utflx" : "vtflx") + om,
                             c.dev(DIR == 0 ? "usflx" : "vsflx") + om, 1);
  }
}

template <int NT, int VAR>
void cppm_run_var(int n, int mm, int nn) {
  Ctx& c = C(); const Geom& g = c.g;
  const int nstep = (int)c.scalar("nstep");
  constexpr int hw = (VAR & 1) ? 3 : 4;   // :2760-2761 / :2776-2777
  halo_update(std::vector<HaloReq>{{c.dev("cau"), g.kdm, halo_uv}, {c.dev("cav"), g.kdm, halo_vv}}, hw, hw);
  const long on = (long)nn * g.lev;
  double* dpA = c.dev("dp") + on;
  double* dpB = c.owned("cppm_tmp_dp", g.kdm);
  ScalarPtrs<NT> AB{}, BA{};
  double* a[NT]; double* b[NT];
  a[0] = c.dev("temp") + on; a[1] = c.dev("saln") + on;
  b[0] = c.owned("cppm_tmp_temp", g.kdm); b[1] = c.owned("cppm_tmp_saln", g.kdm);
  for (int nt = 2; nt < NT; ++nt) {
    a[nt] = c.dev("trc") + on + (long)(nt - 2) * 2 * g.kdm * g.lev;
    b[nt] = c.owned("cppm_tmp_trc" + std::to_string(nt - 1), g.kdm);
  }
  for (int nt = 0; nt < NT; ++nt) { AB.src[nt] = a[nt]; AB.dst[nt] = b[nt]; BA.src[nt] = b[nt]; BA.dst[nt] = a[nt]; }
  // passive tracers beyond the NT-2 that travel with T and S: transported in further launches of each pass
  std::vector<std::pair<double*, double*>> exAB, exBA;
  for (int nt = NT - 2; nt < g.ntr; ++nt) {
    double* ta = c.dev("trc") + on + (long)nt * 2 * g.kdm * g.lev;
    double* tb = c.owned("cppm_tmp_trc" + std::to_string(nt + 1), g.kdm);
    exAB.push_back({ta, tb}); exBA.push_back({tb, ta});
  }
  if (nstep % 2 == 1) {
    cppm_pass<0, NT, VAR>(false, n, mm, dpA, dpB, AB);
    cppm_pass_extra<0, VAR>(false, n, mm, dpA, dpB, exAB);
    cppm_pass<1, NT, VAR>(true, n, mm, dpB, dpA, BA);
    cppm_pass_extra<1, VAR>(true, n, mm, dpB, dpA, exBA);
  } else {
    cppm_pass<1, NT, VAR>(false, n, mm, dpA, dpB, AB);
    cppm_pass_extra<1, VAR>(false, n, mm, dpA, dpB, exAB);
    cppm_pass<0, NT, VAR>(true, n, mm, dpB, dpA, BA);
    cppm_pass_extra<0, VAR>(true, n, mm, dpB, dpA, exBA);
  }
}

// resolves the namelist options exactly like init_cppm (:2524-2549), same messages
int cppm_variant() {
  Ctx& c = C();
  const std::string comp = c.option("cppm_compatibility", "full"), lim = c.option("cppm_limiting", "non_oscillatory");
  if (comp != "full" && comp != "partial")
    throw std::runtime_error(" init_cppm: cppm_compatibility = " + comp + " is unsupported!");
  if (lim != "monotonic" && lim != "non_oscillatory")
    throw std::runtime_error(" init_cppm: cppm_limiting = " + lim + " is unsupported!");
  return (comp == "partial" ? 2 : 0) | (lim == "monotonic" ? 1 : 0);
}

template <int NT>
void cppm_run(int n, int mm, int nn) {
  switch (cppm_variant()) {
    case VAR_FC_NOSC: cppm_run_var<NT, VAR_FC_NOSC>(n, mm, nn); break;
    case VAR_FC_MONO: cppm_run_var<NT, VAR_FC_MONO>(n, mm, nn); break;
    case VAR_PC_NOSC: cppm_run_var<NT, VAR_PC_NOSC>(n, mm, nn); break;
    default: cppm_run_var<NT, VAR_PC_MONO>(n, mm, nn); break;
  }
}

}  // namespace

// init_cppm (mod_cppm.F90:2504-2746)
void init_cppm_dev() {
  Ctx& c = C(); const Geom& g = c.g;
  (void)cppm_variant();   // option check with the reference's messages (:2524-2549)
  double* ti = c.owned("cppm_tab_i", T_NLEV);
  double* tj = c.owned("cppm_tab_j", T_NLEV);
  CUDA_CHECK(cudaMemsetAsync(ti, 0, sizeof(double) * g.lev * T_NLEV, c.stream));
  CUDA_CHECK(cudaMemsetAsync(tj, 0, sizeof(double) * g.lev * T_NLEV, c.stream));
  LAUNCH(cppm_tables_kernel, cdiv((long)g.ii * g.jj, 128), 128, 0, g, c.idev("ip"), c.dev("scpx"),
         c.dev("scpy"), ti, tj);
  // halos: coefficient tables travel as u/v-type scalars, slope/curvature masks
  // as p-type (:2605-2646)
  halo_update(std::vector<HaloReq>{{ti, T_SSC, halo_us}, {ti + (long)T_STEN * g.lev, 1, halo_us},
                                   {ti + (long)T_SSC * g.lev, 3, halo_ps}}, g.nb, 0);
  halo_update(std::vector<HaloReq>{{tj, T_SSC, halo_vs}, {tj + (long)T_STEN * g.lev, 1, halo_vs},
                                   {tj + (long)T_SSC * g.lev, 3, halo_ps}}, 0, g.nb);
  if (g.nreg == 2 && g.north)
    LAUNCH(cppm_tables_arctic, cdiv((long)g.ldi * (1 + g.nb), 128), 128, 0, g, ti, tj);
  int* si = c.owned_int("cppm_sten_i", 1);
  int* sj = c.owned_int("cppm_sten_j", 1);
  LAUNCH(cppm_tags_to_int, cdiv(g.lev, 256), 256, 0, g, ti, tj, si, sj);
  c.owned("cppm_hel_3d", g.kdm);
  c.owned("cppm_her_3d", g.kdm);
}

// advect (mod_advect.F90:59-189), advmth='cppm'
void advect_dev(int m, int n, int mm, int nn, int k1m, int k1n) {
  Ctx& c = C(); const Geom& g = c.g;
  const std::string advmth = c.option("advmth", "cppm");
  if (advmth != "cppm" && advmth != "remap") throw std::runtime_error(" advmth = " + advmth + " is unsupported!");
  if (advmth == "cppm" && !c.has("cppm_tab_i")) throw std::runtime_error("advect: init_cppm has not been called");
  const dim3 grid = lgrid(g, dim3(cdiv(g.ii, 128), g.jj, g.kdm));
  OCC_DISPATCH3("advect_area_minblk", 16, 10, 12, 16,
  LAUNCH_NAMED("advect_flux_area", advect_flux_area<OCC>, grid, 128, 0, g, m, mm, nn, c.scalar("delt1"), c.scalar("dlt"), c.idev("iu"),
         c.idev("iv"), c.dev("u"), c.dev("v"), c.dev("dpu"), c.dev("dpv"), c.dev("ubflxs_p"),
         c.dev("vbflxs_p"), c.dev("pbu"), c.dev("pbv"), c.dev("umfltd"), c.dev("vmfltd"),
         c.dev("umflsm"), c.dev("vmflsm"), c.dev("scuy"), c.dev("scvx"), c.dev("umax"), c.dev("vmax"),
         c.dev("cau"), c.dev("cav")));
  if (advmth == "remap") {   // mod_advect.F90:96-153 (no trailing halo update in this branch)
    advect_remap_dev(m, n, mm, nn, k1m, k1n);
    return;
  }
  switch (std::min(4, 2 + g.ntr)) {   // T, S and up to two tracers in the first group, the rest two at a time
    case 2: cppm_run<2>(n, mm, nn); break;
    case 3: cppm_run<3>(n, mm, nn); break;
    default: cppm_run<4>(n, mm, nn); break;
  }
  const long on = (long)nn * g.lev;
  std::vector<HaloReq> reqs{{c.dev("dp") + on, g.kdm, halo_ps}, {c.dev("temp") + on, g.kdm, halo_ps},
                            {c.dev("saln") + on, g.kdm, halo_ps}};
  for (int nt = 0; nt < g.ntr; ++nt)
    reqs.push_back({c.dev("trc") + on + (long)nt * 2 * g.kdm * g.lev, g.kdm, halo_ps});
  halo_update(reqs, 1, 1);
}

}  // namespace blom
