// Layer-thickness and tracer transport by incremental remapping (advmth='remap').
//
// Reference: phy/mod_remap.F90:53-199 (triint, penint), :205-1522 (remap) and its driver
// phy/mod_advect.F90:96-153.  Build options of the reference covered: use_TRC = (ntr > 0),
// use_ATRC = use_TKE = .false. (defaults).
//
// B200 design (not the reference's layer-at-a-time 2-D temporaries): all kdm layers go through
// each stage in ONE launch (grid.z = layer), lanes along i:
//   remap_pbmin   9-point wet-neighbour minimum of the bottom pressure (2-D, once per call)
//   remap_grad    limited linear reconstructions (dx,dy | td,tx,ty | sd,sx,sy | trd,trx,try) of every
//                 cell on -1..ii+2 x -1..jj+2 -> library scratch (8+3*ntr fields)
//   remap_flux    one thread per (i,j): the u face and the v face it owns; corner velocities are
//                 recomputed from cau/cav (4 loads, L1-resident) instead of staged through memory;
//                 departure polygons (2 triangles + 1 pentagon per face) integrated in registers;
//                 writes the face fluxes to scratch and adds them to uflx.. / assigns vflx..
//   remap_update  divergence update of dp,T,S,trc on 0..ii+1 x 0..jj+1
// The in-place hazard of the reference's single array per field (dp is floored and offset by dpeps
// before use, updated at the end) is removed by evaluating max(0,dp)+dpeps on the fly in the first
// three stages; only remap_update writes state.
#include "common.cuh"

namespace blom {

namespace {

#define RM_DPEPS 1.e-12

struct Mom { double a, ax, ay, axx, ayy, axy; };

// phy/mod_remap.F90:53-102
__device__ __forceinline__ Mom triint(double ac, double x1, double y1, double x2, double y2, double x3, double y3) {
  const double r1_3 = 1. / 3., r1_6 = 1. / 6., r1_12 = 1. / 12.;
  Mom m;
  const double xx = x1 * x2 + x2 * x3 + x1 * x3;
  const double yy = y1 * y2 + y2 * y3 + y1 * y3;
  const double xy1 = x1 * y1, xy2 = x2 * y2, xy3 = x3 * y3;
  const double xy = xy1 + xy2 + xy3;
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

// phy/mod_remap.F90:104-199: pentagon = triangles 123 + 135 + 345
__device__ __forceinline__ Mom penint(double ac, double x1, double y1, double x2, double y2, double x3, double y3,
                                      double x4, double y4, double x5, double y5) {
  const double r1_3 = 1. / 3., r1_6 = 1. / 6., r1_12 = 1. / 12.;
  const double xx123 = x1 * x2 + x2 * x3 + x1 * x3, yy123 = y1 * y2 + y2 * y3 + y1 * y3;
  const double xx135 = x1 * x3 + x3 * x5 + x1 * x5, yy135 = y1 * y3 + y3 * y5 + y1 * y5;
  const double xx345 = x3 * x4 + x4 * x5 + x3 * x5, yy345 = y3 * y4 + y4 * y5 + y3 * y5;
  const double xy1 = x1 * y1, xy2 = x2 * y2, xy3 = x3 * y3, xy4 = x4 * y4, xy5 = x5 * y5;
  const double xy123 = xy1 + xy2 + xy3, xy135 = xy1 + xy3 + xy5, xy345 = xy3 + xy4 + xy5;
  const double a123 = .5 * ((x2 - x1) * (y3 - y1) - (y2 - y1) * (x3 - x1)) * ac;
  const double a135 = .5 * ((x3 - x1) * (y5 - y1) - (y3 - y1) * (x5 - x1)) * ac;
  const double a345 = .5 * ((x4 - x3) * (y5 - y3) - (y4 - y3) * (x5 - x3)) * ac;
  const double ax123 = r1_3 * (x1 + x2 + x3), ay123 = r1_3 * (y1 + y2 + y3);
  const double ax135 = r1_3 * (x1 + x3 + x5), ay135 = r1_3 * (y1 + y3 + y5);
  const double ax345 = r1_3 * (x3 + x4 + x5), ay345 = r1_3 * (y3 + y4 + y5);
  const double axx123 = r1_6 * (9. * ax123 * ax123 - xx123);
  const double ayy123 = r1_6 * (9. * ay123 * ay123 - yy123);
  const double axy123 = r1_12 * (9. * ax123 * ay123 + xy123);
  const double axx135 = r1_6 * (9. * ax135 * ax135 - xx135);
  const double ayy135 = r1_6 * (9. * ay135 * ay135 - yy135);
  const double axy135 = r1_12 * (9. * ax135 * ay135 + xy135);
  const double axx345 = r1_6 * (9. * ax345 * ax345 - xx345);
  const double ayy345 = r1_6 * (9. * ay345 * ay345 - yy345);
  const double axy345 = r1_12 * (9. * ax345 * ay345 + xy345);
  Mom m;
  m.a = a123 + a135 + a345;
  m.ax = ax123 * a123 + ax135 * a135 + ax345 * a345;
  m.ay = ay123 * a123 + ay135 * a135 + ay345 * a345;
  m.axx = axx123 * a123 + axx135 * a135 + axx345 * a345;
  m.ayy = ayy123 * a123 + ayy135 * a135 + ayy345 * a345;
  m.axy = axy123 * a123 + axy135 * a135 + axy345 * a345;
  return m;
}

// element offsets of the eight wet neighbours (:368-381): a dry neighbour is replaced by the cell itself
struct Nbr { long w, e, s, n, sw, se, nw, ne; int di, dj; };
__device__ __forceinline__ Nbr neighbours(const Geom& g, long x, const int* __restrict__ ip,
                                          const int* __restrict__ iu, const int* __restrict__ iv) {
  const long s = g.ldi;
  const int miw = iu[x], mie = iu[x + 1], mjs = iv[x], mjn = iv[x + s];
  Nbr q;
  q.w = x - miw; q.e = x + mie; q.s = x - s * mjs; q.n = x + s * mjn;
  const long xsw = x - miw - s * mjs, xse = x + mie - s * mjs, xnw = x - miw + s * mjn, xne = x + mie + s * mjn;
  q.sw = ip[xsw] ? xsw : x;
  q.se = ip[xse] ? xse : x;
  q.nw = ip[xnw] ? xnw : x;
  q.ne = ip[xne] ? xne : x;
  q.di = mie + miw;   // ie - iw
  q.dj = mjn + mjs;   // jn - js
  return q;
}

__device__ __forceinline__ double max8(double a, double b, double c, double d, double e, double f, double g,
                                       double h) {
  return fmax(fmax(fmax(a, b), fmax(c, d)), fmax(fmax(e, f), fmax(g, h)));
}
__device__ __forceinline__ double min8(double a, double b, double c, double d, double e, double f, double g,
                                       double h) {
  return fmin(fmin(fmin(a, b), fmin(c, d)), fmin(fmin(e, f), fmin(g, h)));
}

// phy/mod_advect.F90:98-121
__global__ void remap_pbmin(Geom g, const int* __restrict__ ip, const int* __restrict__ iu,
                            const int* __restrict__ iv, const double* __restrict__ pbot /* p(:,:,kk+1) */,
                            double* __restrict__ pbmin) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x - 1, j = (int)blockIdx.y - 1;
  if (i > g.ii + 2) return;
  const long x = ix2(g, i, j);
  if (ip[x] != 1) return;
  const Nbr b = neighbours(g, x, ip, iu, iv);
  pbmin[x] = fmin(min8(pbot[b.sw], pbot[b.s], pbot[b.se], pbot[b.w], pbot[x], pbot[b.e], pbot[b.nw], pbot[b.n]),
                  pbot[b.ne]);
}

// limited slopes and centre value of one scalar (:417-474 T, S; :558-594 tracers)
__device__ __forceinline__ void scalar_grad(const double* __restrict__ f, long x, const Nbr& b, double dxi,
                                            double dyi, double xd, double yd, double& gx, double& gy, double& gd) {
  const double fc = f[x], fw = f[b.w], fe = f[b.e], fs = f[b.s], fn = f[b.n];
  gx = (fe - fw) * dxi;
  gy = (fn - fs) * dyi;
  const double q1 = gx * (-.5 - xd), q2 = gx * (.5 - xd), q3 = gy * (-.5 - yd), q4 = gy * (.5 - yd);
  const double tgmx = fmax(q1, q2) + fmax(q3, q4);
  const double tgmn = fmin(q1, q2) + fmin(q3, q4);
  const double fsw = f[b.sw], fse = f[b.se], fnw = f[b.nw], fne = f[b.ne];
  const double tfmx = fmax(0., max8(fsw, fs, fse, fw, fe, fnw, fn, fne) - fc);
  const double tfmn = fmin(0., min8(fsw, fs, fse, fw, fe, fnw, fn, fne) - fc);
  if (tfmx > 0. && tfmn < 0.) {
    const double q = fmin(tfmx / fmax(tfmx, tgmx), tfmn / fmin(tfmn, tgmn));
    gx = gx * q;
    gy = gy * q;
    gd = fc - gx * xd - gy * yd;
  } else {
    gx = 0.; gy = 0.; gd = fc;
  }
}

// gradient scratch: field q of layer k at G + (q*kdm + k-1)*lev
enum { G_DX = 0, G_DY = 1, G_TD = 2, G_TX = 3, G_TY = 4, G_SD = 5, G_SX = 6, G_SY = 7, G_TR = 8 };

// :361-600 on -1..ii+2 x -1..jj+2
__global__ void __launch_bounds__(128)
remap_grad(Geom g, int nn, const int* __restrict__ ip, const int* __restrict__ iu, const int* __restrict__ iv,
           const double* __restrict__ pbmin, const double* __restrict__ p, const double* __restrict__ dp,
           const double* __restrict__ temp, const double* __restrict__ saln, const double* __restrict__ trc,
           double* __restrict__ G) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x - 1, j = (int)blockIdx.y - 1, k = blockIdx.z + 1;
  if (i > g.ii + 2) return;
  const long x = ix2(g, i, j);
  if (ip[x] != 1) return;
  const long lev = g.lev, kn = (long)(k + nn - 1) * lev;
  const double* __restrict__ dpk = dp + kn;
  const double* __restrict__ plo = p + (long)k * lev;   // p(:,:,k+1)
  const Nbr b = neighbours(g, x, ip, iu, iv);
  const double dxi = 1. / max(1, b.di), dyi = 1. / max(1, b.dj);
  const double pbm = pbmin[x];
  // dp -> max(0,dp)+dpeps, pup = plo - dp (:304-311), then the bottom-clipped thickness (:387-395)
  auto lim = [&](long y) {
    const double d = fmax(0., dpk[y]) + RM_DPEPS;
    const double pup = plo[y] - d;
    return fmax(RM_DPEPS, fmin(pbm - pup, d));
  };
  const double dpsw = lim(b.sw), dps = lim(b.s), dpse = lim(b.se), dpw = lim(b.w), dpc = lim(x), dpe = lim(b.e),
               dpnw = lim(b.nw), dpn = lim(b.n), dpne = lim(b.ne);
  double dx = (dpe - dpw) * dxi;
  double dy = (dpn - dps) * dyi;
  const double dgmx = .5 * (fabs(dx) + fabs(dy));
  const double dfmx = fmax(0., max8(dpsw, dps, dpse, dpw, dpe, dpnw, dpn, dpne) - dpc);
  const double dfmn = fmin(0., min8(dpsw, dps, dpse, dpw, dpe, dpnw, dpn, dpne) - dpc);
  double xd, yd;
  if (dfmx > 0. && dfmn < 0.) {
    const double q = fmin(dfmx / fmax(dfmx, dgmx), dfmn / fmin(dfmn, -dgmx));
    const double dpx = fmax(0., dpk[x]) + RM_DPEPS;
    dx = dx * q;
    dy = dy * q;
    xd = dx / (12. * dpx);
    yd = dy / (12. * dpx);
  } else {
    dx = 0.; dy = 0.; xd = 0.; yd = 0.;
  }
  const long ko = (long)(k - 1) * lev + x, kl = (long)g.kdm * lev;
  G[G_DX * kl + ko] = dx;
  G[G_DY * kl + ko] = dy;
  double gx, gy, gd;
  scalar_grad(temp + kn, x, b, dxi, dyi, xd, yd, gx, gy, gd);
  G[G_TD * kl + ko] = gd; G[G_TX * kl + ko] = gx; G[G_TY * kl + ko] = gy;
  scalar_grad(saln + kn, x, b, dxi, dyi, xd, yd, gx, gy, gd);
  G[G_SD * kl + ko] = gd; G[G_SX * kl + ko] = gx; G[G_SY * kl + ko] = gy;
  for (int nt = 0; nt < g.ntr; ++nt) {
    scalar_grad(trc + kn + (long)nt * 2 * g.kdm * lev, x, b, dxi, dyi, xd, yd, gx, gy, gd);
    G[(G_TR + 3 * nt) * kl + ko] = gd; G[(G_TR + 3 * nt + 1) * kl + ko] = gx; G[(G_TR + 3 * nt + 2) * kl + ko] = gy;
  }
}

constexpr int RM_MAXTR = 4;   // passive tracers carried in registers by the flux kernel

// per-layer pointers the flux stage needs
struct FluxIn {
  const int* ip;
  const double *cau, *cav, *scp2, *scp2i, *dp, *plo, *G;
  long kl;     // kdm*lev: distance between gradient fields
  int ntr;
};

// non-dimensional face velocities (:604-626)
__device__ __forceinline__ double cu_at(const FluxIn& f, long x) {
  const double c = f.cau[x];
  return c > 0. ? c * f.scp2i[x - 1] : c * f.scp2i[x];
}
__device__ __forceinline__ double cv_at(const FluxIn& f, long x, long s) {
  const double c = f.cav[x];
  return c > 0. ? c * f.scp2i[x - s] : c * f.scp2i[x];
}
// corner velocities (:639-680).  Faces read here are wet whenever they are read (nw==4, or the wet
// pair selected for nw==2), so no face mask is needed.
__device__ __forceinline__ void corner(const FluxIn& f, long x, long s, double& cuc, double& cvc) {
  const int p00 = f.ip[x - 1 - s], p10 = f.ip[x - s], p01 = f.ip[x - 1], p11 = f.ip[x];
  const int nw = p00 + p10 + p01 + p11;
  cuc = 0.; cvc = 0.;
  if (nw == 4) {
    const double cus = cu_at(f, x - s), cun = cu_at(f, x), cvw = cv_at(f, x - 1, s), cve = cv_at(f, x, s);
    if (!(cus * cun <= 0.)) cuc = 2. * cus * cun / (cus + cun);
    if (!(cvw * cve <= 0.)) cvc = 2. * cvw * cve / (cvw + cve);
  } else if (nw == 2) {
    if (p00 + p10 == 2) cuc = cu_at(f, x - s);
    else if (p01 + p11 == 2) cuc = cu_at(f, x);
    else if (p00 + p01 == 2) cvc = cv_at(f, x - 1, s);
    else if (p10 + p11 == 2) cvc = cv_at(f, x, s);
  }
}

struct FaceAcc { double fd, ft, fs, ftr[RM_MAXTR]; };

// common tail of every donor-cell block (e.g. :715-760)
__device__ __forceinline__ void donor(const FluxIn& f, long y /* donor cell */, const Mom& m, double pbf,
                                      FaceAcc& a) {
  const double d = fmax(0., f.dp[y]) + RM_DPEPS;
  const double pup = f.plo[y] - d;
  const double dl = fmin(d, fmax(0., pbf - pup));
  const double* __restrict__ G = f.G + y;
  const double dx = G[G_DX * f.kl], dy = G[G_DY * f.kl];
  const double fd = m.a * dl + m.ax * dx + m.ay * dy;
  a.fd = a.fd + fd;
  const double qx = m.ax * dl + m.axx * dx + m.axy * dy;
  const double qy = m.ay * dl + m.axy * dx + m.ayy * dy;
  a.ft = a.ft + fd * G[G_TD * f.kl] + qx * G[G_TX * f.kl] + qy * G[G_TY * f.kl];
  a.fs = a.fs + fd * G[G_SD * f.kl] + qx * G[G_SX * f.kl] + qy * G[G_SY * f.kl];
#pragma unroll
  for (int nt = 0; nt < RM_MAXTR; ++nt)
    if (nt < f.ntr)
      a.ftr[nt] = a.ftr[nt] + fd * G[(G_TR + 3 * nt) * f.kl] + qx * G[(G_TR + 3 * nt + 1) * f.kl] +
                  qy * G[(G_TR + 3 * nt + 2) * f.kl];
}

// flux scratch: field q of layer k at F + (q*kdm + k-1)*lev; q = 0..2+ntr u faces, then v faces
// :688-1462 on i = 0..ii+2, j = 0..jj+2 (u faces j <= jj+1, v faces i <= ii+1)
__global__ void __launch_bounds__(128)
remap_flux(Geom g, int n, int mm, int nn, const int* __restrict__ ip, const int* __restrict__ iu,
           const int* __restrict__ iv, const double* __restrict__ cau, const double* __restrict__ cav,
           const double* __restrict__ scp2, const double* __restrict__ scp2i, const double* __restrict__ pbu,
           const double* __restrict__ pbv, const double* __restrict__ p, const double* __restrict__ dp,
           const double* __restrict__ G, double* __restrict__ F, double* __restrict__ uflx,
           double* __restrict__ vflx, double* __restrict__ utflx, double* __restrict__ vtflx,
           double* __restrict__ usflx, double* __restrict__ vsflx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, k = blockIdx.z + 1;
  if (i > g.ii + 2) return;
  const long x = ix2(g, i, j), s = g.ldi, lev = g.lev;
  const long kk1 = (long)(k - 1) * lev, kl = (long)g.kdm * lev, km = (long)(k + mm - 1) * lev;
  FluxIn f{ip, cau + kk1, cav + kk1, scp2, scp2i, dp + (long)(k + nn - 1) * lev, p + (long)k * lev, G + kk1, kl,
           g.ntr};
  const int nf = 3 + g.ntr;
  const bool uface = j <= g.jj + 1, vface = i <= g.ii + 1;
  const bool uwet = uface && iu[x] == 1, vwet = vface && iv[x] == 1;
  double cuc00 = 0., cvc00 = 0., cuc01 = 0., cvc01 = 0., cuc10 = 0., cvc10 = 0.;
  if (uwet || vwet) corner(f, x, s, cuc00, cvc00);
  if (uwet) corner(f, x + s, s, cuc01, cvc01);   // corner (i,j+1)
  if (vwet) corner(f, x + 1, s, cuc10, cvc10);   // corner (i+1,j)

  FaceAcc a;
  a.fd = 0.; a.ft = 0.; a.fs = 0.;
#pragma unroll
  for (int nt = 0; nt < RM_MAXTR; ++nt) a.ftr[nt] = 0.;
  if (uwet) {
    const double cu = cu_at(f, x), pbf = pbu[x + (long)(n - 1) * lev];
    const double ym = -.5 * (cvc00 + cvc01);
    const double xm = ((ym + .5) * cuc00 - (ym - .5) * cuc01 - 2. * cu) / (1. + cvc00 - cvc01);
    double x2, y2, x4, y4;
    if (cu > 0.) {
      if (cvc00 > 0.) {
        const double xc0 = (xm * cvc00 - cuc00 * (ym + .5)) / (cvc00 + ym + .5);
        const double xc1 = xc0 * scp2[x - 1] * scp2i[x - 1 - s];
        x4 = xc0 + .5; y4 = -.5;
        donor(f, x - 1 - s, triint(scp2[x - 1 - s], xc1 + .5, .5, -cuc00 + .5, -cvc00 + .5, .5, .5), pbf, a);
      } else {
        x4 = -cuc00 + .5; y4 = -cvc00 - .5;
      }
      if (cvc01 < 0.) {
        const double xc0 = (xm * cvc01 - cuc01 * (ym - .5)) / (cvc01 + ym - .5);
        const double xc1 = xc0 * scp2[x - 1] * scp2i[x - 1 + s];
        x2 = xc0 + .5; y2 = .5;
        donor(f, x - 1 + s, triint(scp2[x - 1 + s], xc1 + .5, -.5, .5, -.5, -cuc01 + .5, -cvc01 - .5), pbf, a);
      } else {
        x2 = -cuc01 + .5; y2 = -cvc01 + .5;
      }
      donor(f, x - 1, penint(scp2[x - 1], .5, .5, x2, y2, xm + .5, ym, x4, y4, .5, -.5), pbf, a);
    } else {
      if (cvc00 > 0.) {
        const double xc0 = (xm * cvc00 - cuc00 * (ym + .5)) / (cvc00 + ym + .5);
        const double xc1 = xc0 * scp2[x] * scp2i[x - s];
        x4 = xc0 - .5; y4 = -.5;
        donor(f, x - s, triint(scp2[x - s], xc1 - .5, .5, -cuc00 - .5, -cvc00 + .5, -.5, .5), pbf, a);
      } else {
        x4 = -cuc00 - .5; y4 = -cvc00 - .5;
      }
      if (cvc01 < 0.) {
        const double xc0 = (xm * cvc01 - cuc01 * (ym - .5)) / (cvc01 + ym - .5);
        const double xc1 = xc0 * scp2[x] * scp2i[x + s];
        x2 = xc0 - .5; y2 = .5;
        donor(f, x + s, triint(scp2[x + s], xc1 - .5, -.5, -.5, -.5, -cuc01 - .5, -cvc01 - .5), pbf, a);
      } else {
        x2 = -cuc01 - .5; y2 = -cvc01 + .5;
      }
      donor(f, x, penint(scp2[x], -.5, .5, x2, y2, xm - .5, ym, x4, y4, -.5, -.5), pbf, a);
    }
    uflx[km + x] = uflx[km + x] + a.fd;      // accumulated (:1054-1056)
    utflx[km + x] = utflx[km + x] + a.ft;
    usflx[km + x] = usflx[km + x] + a.fs;
  }
  if (uface) {
    double* Fk = F + kk1 + x;
    Fk[0 * kl] = a.fd; Fk[1 * kl] = a.ft; Fk[2 * kl] = a.fs;
    for (int nt = 0; nt < g.ntr; ++nt) Fk[(3 + nt) * kl] = a.ftr[nt];
  }

  a.fd = 0.; a.ft = 0.; a.fs = 0.;
#pragma unroll
  for (int nt = 0; nt < RM_MAXTR; ++nt) a.ftr[nt] = 0.;
  if (vwet) {
    const double cv = cv_at(f, x, s), pbf = pbv[x + (long)(n - 1) * lev];
    const double xm = -.5 * (cuc00 + cuc10);
    const double ym = ((xm + .5) * cvc00 - (xm - .5) * cvc10 - 2. * cv) / (1. + cuc00 - cuc10);
    double x2, y2, x4, y4;
    if (cv > 0) {
      if (cuc00 > 0.) {
        const double yc0 = (ym * cuc00 - cvc00 * (xm + .5)) / (cuc00 + xm + .5);
        const double yc1 = yc0 * scp2[x - s] * scp2i[x - 1 - s];
        x2 = -.5; y2 = yc0 + .5;
        donor(f, x - 1 - s, triint(scp2[x - 1 - s], .5, yc1 + .5, .5, .5, -cuc00 + .5, -cvc00 + .5), pbf, a);
      } else {
        x2 = -cuc00 - .5; y2 = -cvc00 + .5;
      }
      if (cuc10 < 0.) {
        const double yc0 = (ym * cuc10 - cvc10 * (xm - .5)) / (cuc10 + xm - .5);
        const double yc1 = yc0 * scp2[x - s] * scp2i[x + 1 - s];
        x4 = .5; y4 = yc0 + .5;
        donor(f, x + 1 - s, triint(scp2[x + 1 - s], -.5, yc1 + .5, -cuc10 - .5, -cvc10 + .5, -.5, .5), pbf, a);
      } else {
        x4 = -cuc10 + .5; y4 = -cvc10 + .5;
      }
      donor(f, x - s, penint(scp2[x - s], -.5, .5, x2, y2, xm, ym + .5, x4, y4, .5, .5), pbf, a);
    } else {
      if (cuc00 > 0.) {
        const double yc0 = (ym * cuc00 - cvc00 * (xm + .5)) / (cuc00 + xm + .5);
        const double yc1 = yc0 * scp2[x] * scp2i[x - 1];
        x2 = -.5; y2 = yc0 - .5;
        donor(f, x - 1, triint(scp2[x - 1], .5, yc1 - .5, .5, -.5, -cuc00 + .5, -cvc00 - .5), pbf, a);
      } else {
        x2 = -cuc00 - .5; y2 = -cvc00 - .5;
      }
      if (cuc10 < 0.) {
        const double yc0 = (ym * cuc10 - cvc10 * (xm - .5)) / (cuc10 + xm - .5);
        const double yc1 = yc0 * scp2[x] * scp2i[x + 1];
        x4 = .5; y4 = yc0 - .5;
        donor(f, x + 1, triint(scp2[x + 1], -.5, yc1 - .5, -cuc10 - .5, -cvc10 - .5, -.5, -.5), pbf, a);
      } else {
        x4 = -cuc10 + .5; y4 = -cvc10 - .5;
      }
      donor(f, x, penint(scp2[x], -.5, -.5, x2, y2, xm, ym - .5, x4, y4, .5, -.5), pbf, a);
    }
    vflx[km + x] = a.fd;                     // assigned (:1455-1457; reference quirk)
    vtflx[km + x] = a.ft;
    vsflx[km + x] = a.fs;
  }
  if (vface) {
    double* Fk = F + (long)nf * kl + kk1 + x;
    Fk[0 * kl] = a.fd; Fk[1 * kl] = a.ft; Fk[2 * kl] = a.fs;
    for (int nt = 0; nt < g.ntr; ++nt) Fk[(3 + nt) * kl] = a.ftr[nt];
  }
}

// :1468-1520 on 0..ii+1 x 0..jj+1; wet cells of the two rings outside keep the dpeps offset the
// reference leaves on them (:307-308 are not undone there)
__global__ void __launch_bounds__(128)
remap_update(Geom g, int nn, const int* __restrict__ ip, const double* __restrict__ scp2i,
             const double* __restrict__ F, double* __restrict__ dp, double* __restrict__ temp,
             double* __restrict__ saln, double* __restrict__ trc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x - 2, j = (int)blockIdx.y - 2, k = blockIdx.z + 1;
  if (i > g.ii + 3) return;
  const long x = ix2(g, i, j), s = g.ldi, lev = g.lev;
  if (ip[x] != 1) return;
  const long kn = (long)(k + nn - 1) * lev + x, kl = (long)g.kdm * lev;
  const double q = fmax(0., dp[kn]) + RM_DPEPS;
  if (i < 0 || i > g.ii + 1 || j < 0 || j > g.jj + 1) { dp[kn] = q; return; }
  const int nf = 3 + g.ntr;
  const double* Fu = F + (long)(k - 1) * lev + x;
  const double* Fv = Fu + (long)nf * kl;
  const double a = scp2i[x];
  const double dpn = q - (Fu[1] - Fu[0] + Fv[s] - Fv[0]) * a;
  temp[kn] = (q * temp[kn] - (Fu[kl + 1] - Fu[kl] + Fv[kl + s] - Fv[kl]) * a) / dpn;
  saln[kn] = (q * saln[kn] - (Fu[2 * kl + 1] - Fu[2 * kl] + Fv[2 * kl + s] - Fv[2 * kl]) * a) / dpn;
  for (int nt = 0; nt < g.ntr; ++nt) {
    double* t = trc + kn + (long)nt * 2 * g.kdm * lev;
    const long o = (long)(3 + nt) * kl;
    *t = (q * *t - (Fu[o + 1] - Fu[o] + Fv[o + s] - Fv[o]) * a) / dpn;
  }
  dp[kn] = fmax(0., dpn - RM_DPEPS);
}

}  // namespace

// phy/mod_advect.F90:96-153; the flux-area prelude (:71-94) has been launched by advect_dev
void advect_remap_dev(int m, int n, int mm, int nn, int k1m, int k1n) {
  (void)m; (void)k1m; (void)k1n;
  Ctx& c = C(); const Geom& g = c.g;
  if (g.ntr > RM_MAXTR) throw std::runtime_error("advect(remap): this build transports at most 4 passive tracers");
  const int* ip = c.idev("ip"); const int* iu = c.idev("iu"); const int* iv = c.idev("iv");
  double* pbmin = c.owned("remap_pbmin", 1);
  {
    dim3 grid(cdiv(g.ii + 4, 128), g.jj + 4);
    LAUNCH(remap_pbmin, grid, 128, 0, g, ip, iu, iv, c.dev("p") + (long)g.kdm * g.lev, pbmin);
  }
  const long on = (long)nn * g.lev;
  std::vector<HaloReq> reqs{{c.dev("cau"), g.kdm, halo_uv}, {c.dev("cav"), g.kdm, halo_vv}};
  halo_update(reqs, 3, 3);
  if (g.ntr > 0) {
    std::vector<HaloReq> tr;
    for (int nt = 0; nt < g.ntr; ++nt) tr.push_back({c.dev("trc") + on + (long)nt * 2 * g.kdm * g.lev, g.kdm, halo_ps});
    halo_update(tr, 3, 3);
  }
  double* G = c.owned("remap_grad", (8 + 3 * g.ntr) * g.kdm);
  double* F = c.owned("remap_flux", 2 * (3 + g.ntr) * g.kdm);
  double* trc = g.ntr > 0 ? c.dev("trc") : nullptr;
  {
    dim3 grid(cdiv(g.ii + 4, 128), g.jj + 4, g.kdm);
    LAUNCH(remap_grad, grid, 128, 0, g, nn, ip, iu, iv, pbmin, c.dev("p"), c.dev("dp"), c.dev("temp"),
           c.dev("saln"), trc, G);
  }
  {
    dim3 grid(cdiv(g.ii + 3, 128), g.jj + 3, g.kdm);
    LAUNCH(remap_flux, grid, 128, 0, g, n, mm, nn, ip, iu, iv, c.dev("cau"), c.dev("cav"), c.dev("scp2"),
           c.dev("scp2i"), c.dev("pbu"), c.dev("pbv"), c.dev("p"), c.dev("dp"), G, F, c.dev("uflx"), c.dev("vflx"),
           c.dev("utflx"), c.dev("vtflx"), c.dev("usflx"), c.dev("vsflx"));
  }
  {
    dim3 grid(cdiv(g.ii + 6, 128), g.jj + 6, g.kdm);
    LAUNCH(remap_update, grid, 128, 0, g, nn, ip, c.dev("scp2i"), F, c.dev("dp"), c.dev("temp"), c.dev("saln"),
           trc);
  }
}

}  // namespace blom
