// Eddy-induced transport for the isopycnic bulk-mixed-layer coordinate (vcoord='isopyc_bulkml'):
// eddtra_intdif_isopyc_bulkml (phy/mod_eddtra.F90:153-226), eddtra_gm_isopyc_bulkml (:228-999) and the
// heat/salt flux diagnosis of eddtra for this coordinate (:1834-1857).
//
// B200 design: as for the hybrid coordinate (eddtra.cu) one thread owns one face column with i across
// lanes, so each level access is a coalesced row segment.  The interface transports/fluxes of the GM
// variant live in thread-local arrays; the available layer thicknesses dlm/dlp are re-evaluated from p
// when the limiter needs them instead of being staged; the reference's 2-D ptu/ptv temporaries and
// its separate diagnosis pass over umfltd are fused into the column kernels.  The reference's fatal
// conditions raise the device error flag (reported by the next sync/download, "print + xchalt").
#include "common.cuh"
#include "eos.cuh"

namespace blom {

namespace {

constexpr int KMI = 64;  // compile-time bound on kdm for the thread-local interface arrays

// heat/salt components of a layer mass flux (:1838-1853): .5*mfl*(T_m + T_p)
__device__ __forceinline__ void diag(double f, long xk, long xmk, const double* __restrict__ temp,
                                     const double* __restrict__ saln, double* __restrict__ tfl,
                                     double* __restrict__ sfl) {
  tfl[xk] = .5 * f * (temp[xmk] + temp[xk]);
  sfl[xk] = .5 * f * (saln[xmk] + saln[xk]);
}

// :153-226: interface diffusion.  Thread per face column; the reference's level loop adds q(k) to
// layer k-1 and assigns -q(k) to layer k, i.e. layer k ends up with (-q(k)) + q(k+1).
template <int DIR>
__global__ void __launch_bounds__(128)
eddtra_intdif_isopyc(Geom g, int mm, int nn, double delt1, const int* __restrict__ mask,
                     const double* __restrict__ p, const double* __restrict__ dp,
                     const double* __restrict__ difint, const double* __restrict__ scp2,
                     const double* __restrict__ sca /*scuy|scvx*/, const double* __restrict__ scbi /*scuxi|scvyi*/,
                     const double* __restrict__ temp, const double* __restrict__ saln,
                     double* __restrict__ mfltd, double* __restrict__ tfltd, double* __restrict__ sfltd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.ii) return;
  const long x = ix2(g, i, j);
  if (mask[x] != 1) return;
  const long xm = x - (DIR == 0 ? 1 : g.ldi), lev = g.lev;
  const int kk = g.kdm;
  const double am = scp2[xm], ap = scp2[x];
  auto qk = [&](int k) -> double {   // interface flux at the top of layer k, 4 <= k <= kk
    const long on = (long)(k + nn - 1) * lev, o1 = (long)(k - 2) * lev, o2 = (long)(k - 1) * lev;
    const double flxhi = .125 * fmin(dp[xm + on - lev] * am, dp[x + on] * ap);
    const double flxlo = -.125 * fmin(dp[x + on - lev] * ap, dp[xm + on] * am);
    double q = .25 * (difint[xm + o1] + difint[x + o1] + difint[xm + o2] + difint[x + o2]);
    // delt1*q*(p_m - p_p)*scuy*scuxi evaluated left to right like the reference
    q = fmin(flxhi, fmax(flxlo, delt1 * q * (p[xm + o2] - p[x + o2]) * sca[x] * scbi[x]));
    return q;
  };
  double acc = 0.;   // value layer k holds before the contribution of interface k+1 arrives
  for (int k = 1; k <= kk; ++k) {
    const double qn = (k + 1 >= 4 && k + 1 <= kk) ? qk(k + 1) : 0.;
    double f;
    if (k < 3) f = 0.;                 // layers 1,2 stay zero (:164-166)
    else if (k == kk) f = acc;         // nothing arrives from below
    else f = acc + qn;                 // umfltd(km-1) = umfltd(km-1) + q
    if (k == 3 && kk < 4) f = 0.;
    const long xk = x + (long)(k + mm - 1) * lev, xmk = xm + (long)(k + mm - 1) * lev;
    mfltd[xk] = f;
    diag(f, xk, xmk, temp, saln, tfltd, sfltd);
    acc = -qn;                         // umfltd(km) = -q for the next layer
    if (k + 1 == 3) acc = 0.;          // layer 3 starts from zero (:166)
  }
}

// :228-999: Gent-McWilliams.  One face column per thread.
template <int DIR>
__global__ void __launch_bounds__(128)
eddtra_gm_isopyc(Geom g, int n, int mm, int nn, double delt1, const int* __restrict__ mask,
                 const int* __restrict__ kfpla, const double* __restrict__ p, const double* __restrict__ dp,
                 const double* __restrict__ dpf, const double* __restrict__ temp,
                 const double* __restrict__ saln, const double* __restrict__ difint,
                 const double* __restrict__ nslp, const double* __restrict__ pbf, const double* __restrict__ sc2,
                 const double* __restrict__ scl, const double* __restrict__ scp2, double* __restrict__ mfltd,
                 double* __restrict__ tfltd, double* __restrict__ sfltd, int* __restrict__ err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.ii) return;
  const long x = ix2(g, i, j);
  if (mask[x] != 1) return;
  const long xm = x - (DIR == 0 ? 1 : g.ldi), lev = g.lev;
  const int kk = g.kdm;
  const double ffac = .0625, fface = .99 * ffac, eps = 1.e-14;
  double mfl[KMI + 2], ups[KMI + 2];

  const double et2mf = -grav * rho0 * delt1 * scl[x];
  const double ptf = fmax(p[xm], p[x]);
  const double pb = pbf[x + (long)(n - 1) * lev];
  const double am = scp2[xm], ap = scp2[x];
  const double s2 = sc2[x];
  int kmax = 1;
  for (int k = 3; k <= kk; ++k) {
    const long o = (long)(k + nn - 1) * lev;
    if (dp[xm + o] > epsilp || dp[x + o] > epsilp) kmax = k;
  }
  const int kfm = kfpla[xm + (long)(n - 1) * lev], kfp = kfpla[x + (long)(n - 1) * lev];
  auto T = [&](long base, int kl) { return temp[base + (long)(kl - 1) * lev]; };
  auto S = [&](long base, int kl) { return saln[base + (long)(kl - 1) * lev]; };
  auto kappa_ml = [&]() { return .5 * (difint[xm + lev] + difint[x + lev]); };
  auto upsk = [&](int k) {
    const long o1 = (long)(k - 2) * lev, o2 = (long)(k - 1) * lev;
    const double kappa = .25 * (difint[xm + o1] + difint[x + o1] + difint[xm + o2] + difint[x + o2]);
    return -kappa * nslp[x + o2];
  };
  const double p3m = p[xm + 2 * lev], p3p = p[x + 2 * lev];
  int kintr = 0, kmin = 0;
  bool active = true;   // false: keep the initial zero mass fluxes for this column ("cycle")
  if (kfm > kk && kfp > kk) {
    active = false;                                            // case 1
  } else if (kfm <= kk && kfp > kk) {                           // case 2 (:320-358)
    const int km = 2 + nn;
    kintr = kfm;
    int kn = kintr + nn;
    while (eos::rho(p3p, T(xm, kn), S(xm, kn)) < eos::rho(p3p, T(x, km), S(x, km)) ||
           dp[xm + (long)(kn - 1) * lev] < epsilp) {
      kintr = kintr + 1;
      if (kintr == kmax + 1 || kintr > kk) break;
      kn = kintr + nn;
    }
    if (kintr == kmax + 1 || kintr > kk) active = false;
    else {
      ups[3] = -kappa_ml() * nslp[x + 2 * lev];
      if (ups[3] <= 0.) active = false;
      else {
        kmin = kintr - 1;
        mfl[kmin] = 0.;
        mfl[kintr] = et2mf * ups[3];
        for (int k = kintr + 1; k <= kmax + 1; ++k) mfl[k] = 0.;
      }
    }
  } else if (kfm > kk && kfp <= kk) {                           // case 3 (:360-398)
    const int km = 2 + nn;
    kintr = kfp;
    int kn = kintr + nn;
    while (eos::rho(p3m, T(x, kn), S(x, kn)) < eos::rho(p3m, T(xm, km), S(xm, km)) ||
           dp[x + (long)(kn - 1) * lev] < epsilp) {
      kintr = kintr + 1;
      if (kintr == kmax + 1 || kintr > kk) break;
      kn = kintr + nn;
    }
    if (kintr == kmax + 1 || kintr > kk) active = false;
    else {
      ups[3] = -kappa_ml() * nslp[x + 2 * lev];
      if (ups[3] >= 0.) active = false;
      else {
        kmin = kintr - 1;
        mfl[kmin] = 0.;
        mfl[kintr] = et2mf * ups[3];
        for (int k = kintr + 1; k <= kmax + 1; ++k) mfl[k] = 0.;
      }
    }
  } else {                                                      // case 4 (:400-456)
    kintr = max(kfm, kfp);
    ups[3] = -kappa_ml() * nslp[x + 2 * lev];
    for (int k = kintr + 1; k <= kmax; ++k) ups[k] = upsk(k);
    ups[kmax + 1] = 0.;
    const int km = 2 + nn, kn = kintr - 1 + nn;
    // (a state produced by the bulk mixed layer scheme has kintr <= kmax, so ups[kintr+1] is defined)
    const double du = kintr + 1 <= kmax + 1 ? ups[3] - ups[kintr + 1] : 0.;
    if ((kfm < kintr && du > 0. && eos::rho(p3p, T(xm, kn), S(xm, kn)) > eos::rho(p3p, T(x, km), S(x, km))) ||
        (kfp < kintr && du < 0. && eos::rho(p3m, T(x, kn), S(x, kn)) > eos::rho(p3m, T(xm, km), S(xm, km)))) {
      kintr = kintr - 1;
      ups[kintr + 1] = ups[kintr + 2];
    }
    kmin = kintr - 1;
    mfl[kmin] = 0.;
    mfl[kintr] = et2mf * ups[3];
    for (int k = kintr + 1; k <= kmax; ++k) mfl[k] = et2mf * ups[k];
    mfl[kmax + 1] = 0.;
  }

  if (!active) {
    for (int k = 1; k <= kk; ++k) {
      const long xk = x + (long)(k + mm - 1) * lev, xmk = xm + (long)(k + mm - 1) * lev;
      mfltd[xk] = 0.;
      diag(0., xk, xmk, temp, saln, tfltd, sfltd);
    }
    return;
  }

  // available thicknesses (:463-473); index kmin stands for the mixed layer (layers 1+2)
  auto dl_m = [&](int k) {
    const double lo = k == kmin ? p3m : p[xm + (long)k * lev], up = k == kmin ? p[xm] : p[xm + (long)(k - 1) * lev];
    return fmax(0., fmin(lo, pb) - fmax(up, ptf));
  };
  auto dl_p = [&](int k) {
    const double lo = k == kmin ? p3p : p[x + (long)k * lev], up = k == kmin ? p[x] : p[x + (long)(k - 1) * lev];
    return fmax(0., fmin(lo, pb) - fmax(up, ptf));
  };
  // first guess below the mixed layer base (:478-495)
  {
    const double fhi = fface * fmax(0., fmin((p3m - ptf) * am, (pb - p[x + (long)(kintr - 1) * lev]) * ap));
    const double flo = -fface * fmax(0., fmin((p3p - ptf) * ap, (pb - p[xm + (long)(kintr - 1) * lev]) * am));
    mfl[kmin + 1] = fmin(fhi, fmax(flo, mfl[kmin + 1]));
    for (int k = kmin + 1; k <= kmax - 1; ++k) {
      const double dlm = dl_m(k), dlp = dl_p(k);
      if (mfl[k + 1] - mfl[k] > ffac * fmax(epsilp, dlm) * am) mfl[k + 1] = mfl[k] + fface * dlm * am;
      else if (mfl[k + 1] - mfl[k] < -ffac * fmax(epsilp, dlp) * ap) mfl[k + 1] = mfl[k] - fface * dlp * ap;
      else break;
    }
  }
  auto signif = [&](int k) {
    return fabs(mfl[k + 1] - mfl[k]) > eps * fmax(epsilp * s2, fabs(mfl[k + 1] + mfl[k]));
  };
  // alternate downward/upward limiter sweeps (:500-577)
  bool changed = true;
  int niter = 0, kdir = 1;
  while (changed) {
    niter++;
    if (niter == 1000) { atomicMax(err, 1); break; }
    changed = false;
    kdir = -kdir;
    const int k0 = ((1 - kdir) * kmax + (1 + kdir) * kmin) / 2, nk = kmax - kmin + 1;
    for (int s = 0, k = k0; s < nk; ++s, k += kdir) {
      if (!signif(k)) continue;
      const double lo = mfl[k], hi = mfl[k + 1];
      const double dlm = dl_m(k), dlp = dl_p(k);
      if (hi - lo > ffac * fmax(epsilp, dlm) * am) {
        const double q = fface * dlm * am;
        if (hi > -lo) {
          if (lo > -.5 * q) mfl[k + 1] = lo + q;
          else { mfl[k + 1] = .5 * q; mfl[k] = -mfl[k + 1]; }
        } else {
          if (hi < .5 * q) mfl[k] = hi - q;
          else { mfl[k] = -.5 * q; mfl[k + 1] = -mfl[k]; }
        }
        changed = true;
      } else if (hi - lo < -ffac * fmax(epsilp, dlp) * ap) {
        const double q = fface * dlp * ap;
        if (hi < -lo) {
          if (lo < .5 * q) mfl[k + 1] = lo - q;
          else { mfl[k + 1] = -.5 * q; mfl[k] = -mfl[k + 1]; }
        } else {
          if (hi > -.5 * q) mfl[k] = hi + q;
          else { mfl[k] = .5 * q; mfl[k + 1] = -mfl[k]; }
        }
        changed = true;
      }
    }
  }
  // final layer fluxes (:583-633) + diagnosis
  double f1 = 0., f2 = 0.;
  if (signif(kmin)) {
    f2 = mfl[kmin + 1] - mfl[kmin];
    const double d1 = dpf[x + (long)nn * lev], d2 = dpf[x + (long)(nn + 1) * lev];
    f1 = f2 * d1 / (d1 + d2);
    f2 = f2 - f1;
  }
  for (int k = 1; k <= kk; ++k) {
    const long xk = x + (long)(k + mm - 1) * lev, xmk = xm + (long)(k + mm - 1) * lev;
    double f = 0.;
    if (k == 1) f = f1;
    else if (k == 2) f = f2;
    else if (k >= kintr && k <= kmax) {
      if (signif(k)) f = mfl[k + 1] - mfl[k];
      if (f > ffac * fmax(epsilp, dl_m(k)) * am) atomicMax(err, 2);
      if (f < -ffac * fmax(epsilp, dl_p(k)) * ap) atomicMax(err, 3);
    }
    mfltd[xk] = f;
    diag(f, xk, xmk, temp, saln, tfltd, sfltd);
  }
}

}  // namespace

// eddtra for vcoord='isopyc_bulkml' (:1818-1857)
void eddtra_isopyc_dev(int m, int n, int mm, int nn, int k1m, int k1n) {
  (void)m; (void)k1m; (void)k1n;
  Ctx& c = C(); const Geom& g = c.g;
  const std::string eitmth = c.option("eitmth", "gm");
  const double delt1 = c.scalar("delt1");
  if (g.kdm > KMI) throw std::runtime_error("eddtra: kdm exceeds the compiled column bound (64)");
  dim3 grid(cdiv(g.ii, 128), g.jj);
  if (eitmth == "intdif") {
    LAUNCH_NAMED("eddtra_intdif_isopyc<u>", eddtra_intdif_isopyc<0>, grid, 128, 0, g, mm, nn, delt1, c.idev("iu"),
                 c.dev("p"), c.dev("dp"), c.dev("difint"), c.dev("scp2"), c.dev("scuy"), c.dev("scuxi"),
                 c.dev("temp"), c.dev("saln"), c.dev("umfltd"), c.dev("utfltd"), c.dev("usfltd"));
    LAUNCH_NAMED("eddtra_intdif_isopyc<v>", eddtra_intdif_isopyc<1>, grid, 128, 0, g, mm, nn, delt1, c.idev("iv"),
                 c.dev("p"), c.dev("dp"), c.dev("difint"), c.dev("scp2"), c.dev("scvx"), c.dev("scvyi"),
                 c.dev("temp"), c.dev("saln"), c.dev("vmfltd"), c.dev("vtfltd"), c.dev("vsfltd"));
  } else if (eitmth == "gm") {
    int* err = c.error_flag();
    LAUNCH_NAMED("eddtra_gm_isopyc<u>", eddtra_gm_isopyc<0>, grid, 128, 0, g, n, mm, nn, delt1, c.idev("iu"),
                 c.idev("kfpla"), c.dev("p"), c.dev("dp"), c.dev("dpu"), c.dev("temp"), c.dev("saln"),
                 c.dev("difint"), c.dev("nslpx"), c.dev("pbu"), c.dev("scu2"), c.dev("scuy"), c.dev("scp2"),
                 c.dev("umfltd"), c.dev("utfltd"), c.dev("usfltd"), err);
    LAUNCH_NAMED("eddtra_gm_isopyc<v>", eddtra_gm_isopyc<1>, grid, 128, 0, g, n, mm, nn, delt1, c.idev("iv"),
                 c.idev("kfpla"), c.dev("p"), c.dev("dp"), c.dev("dpv"), c.dev("temp"), c.dev("saln"),
                 c.dev("difint"), c.dev("nslpy"), c.dev("pbv"), c.dev("scv2"), c.dev("scvx"), c.dev("scp2"),
                 c.dev("vmfltd"), c.dev("vtfltd"), c.dev("vsfltd"), err);
    c.error_source = "(eddtra_gm_isopyc_bulkml) 1: no convergence, 2: flux exceeds +ffac*mass, 3: flux exceeds -ffac*mass";
  } else {
    throw std::runtime_error("(eddtra) eitmth_opt is unsupported for vcoord = 'isopyc_bulkml'!");
  }
}

}  // namespace blom
