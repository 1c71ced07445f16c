// Eddy-induced transport for the hybrid coordinate (the isopycnic variants live in
// eddtra_isopyc.cu): eddtra -> eddtra_ale
// (phy/mod_eddtra.F90:1808-1928, :1001-1739; rmeanfilt :121-151), eitmth='gm',
// mlrmth none|fox08|bod23.
//
// B200 design: the reference walks j rows and, per face, a column loop with a
// data-dependent iterative limiter.  Here one thread owns one face column with
// i across lanes, so every level access `a[x + k*lev]` is a coalesced row
// segment; the interface work arrays live in thread-local memory (interleaved
// by the hardware, i.e. also coalesced).  The depth-invariant submesoscale
// factor (upssmx/upssmy), the face top pressure (ptu/ptv) and the heat/salt
// flux diagnosis of `eddtra` are fused into the column kernel, so the 2-D
// temporaries of the reference never exist and temp/saln(km) are read once.
// Fatal conditions of the reference (no convergence after 1000 sweeps, the
// '>'/'<' consistency checks) raise a device error flag that the next
// blomgpu_sync/download reports, matching "print + xchalt".
#include "common.cuh"
#include "eos.cuh"

namespace blom {

namespace {

constexpr int KM = 64;  // compile-time bound on kdm for the thread-local interface arrays

__device__ __forceinline__ void rmeanfilt(double& filtered, double signal, double wg, double wd) {
  const double wf = signal >= filtered ? wg : wd;
  filtered = wf * filtered + (1. - wf) * signal;
}

struct MlParams {
  int mode;  // 0 none, 1 fox08, 2 bod23
  double wg_hbl, wd_hbl, wg_hml, wd_hml, mstar, nstar, wpup_min, mlbl_max_ratio;
  double csm, rtau, lfmin, dbcl82;
};

// running-mean filters of the boundary/mixed layer depths (:1054-1101)
__global__ void eddtra_mlfilter(Geom g, MlParams P, const int* __restrict__ ip,
                                const double* __restrict__ OBLdepth, const double* __restrict__ mld,
                                const double* __restrict__ ustar3, const double* __restrict__ wstar3,
                                double* __restrict__ hbl_tf, double* __restrict__ wpup_tf,
                                double* __restrict__ hml_tf1, double* __restrict__ hml_tf,
                                double* __restrict__ hml_tfbnd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.ii) return;
  const long x = ix2(g, i, j);
  if (ip[x] != 1) return;
  double hb = hbl_tf[x], h1 = hml_tf1[x], h = hml_tf[x];
  rmeanfilt(hb, OBLdepth[x], P.wg_hbl, P.wd_hbl);
  if (P.mode == 2) {
    double w = wpup_tf[x];
    const double wpup = fmax(P.wpup_min, pow(P.mstar * ustar3[x] + P.nstar * wstar3[x], 2. / 3.));
    rmeanfilt(w, wpup, P.wg_hbl, P.wd_hbl);
    wpup_tf[x] = w;
  }
  rmeanfilt(h1, mld[x], P.wg_hbl, P.wd_hbl);
  rmeanfilt(h, h1, P.wg_hml, P.wd_hml);
  hbl_tf[x] = hb; hml_tf1[x] = h1; hml_tf[x] = h;
  hml_tfbnd[x] = fmin(h, P.mlbl_max_ratio * hb);
}

// vertically averaged mixed layer potential density (:1105-1127)
__global__ void eddtra_mldens(Geom g, eos::Coef ec, int nn, const int* __restrict__ ip,
                              const double* __restrict__ p, const double* __restrict__ dp,
                              const double* __restrict__ temp, const double* __restrict__ saln,
                              const double* __restrict__ hml_tfbnd, double* __restrict__ util1) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.ii) return;
  const long x = ix2(g, i, j);
  if (ip[x] != 1) return;
  const int kk = g.kdm;
  const double p1 = p[x];
  const double pml = fmin(p1 + hml_tfbnd[x] * onem, p[x + (long)kk * g.lev]);
  const double dpmli = 1. / (pml - p1);
  double tmldp = 0., smldp = 0., pk = p1;
  for (int k = 1; k <= kk; ++k) {
    const long xn = x + (long)(k + nn - 1) * g.lev;
    const double pk1 = p[x + (long)k * g.lev];
    if (pk1 < pml) {
      const double d = dp[xn];
      tmldp = tmldp + temp[xn] * d;
      smldp = smldp + saln[xn] * d;
    } else {
      tmldp = tmldp + temp[xn] * (pml - pk);
      smldp = smldp + saln[xn] * (pml - pk);
      break;
    }
    pk = pk1;
  }
  util1[x] = eos::sig0(ec, tmldp * dpmli, smldp * dpmli);
}

// One thread per face column.  DIR 0: u faces (minus point i-1), DIR 1: v faces (j-1).
// Only the limited total interface flux mfl(k) is kept in a thread-local array; the interface
// pressures, the GM / submesoscale interface fluxes and the available thicknesses dlm/dlp are
// recomputed from their (cached) inputs where they are needed — the same expressions on the same
// operands, so the values are identical to the reference's stored work arrays, at a fifth of the
// local-memory traffic.
// SMEM (development switch eddtra_mfl=smem): the limited interface fluxes mfl(1..kmax+1) of the block's 128
// columns in shared memory, interleaved by column (mfl(k) of thread t at [k*128 + t], conflict-free), instead of
// in a thread-local array.  The limiter sweeps index the array dynamically, so the local array sits in local
// memory and at 2048 resident threads per SM its 528 bytes per thread spill through L1/L2 into HBM (round 1:
// 32 GB of DRAM traffic for 14.8 GB of algorithmic bytes) - but the shared-memory form measured 1.5x SLOWER
// (see eddtra_dev), so the local array stays the default.
template <int DIR, int MINB, bool SMEM>
__global__ void __launch_bounds__(128, MINB)
eddtra_column(Geom g, MlParams P, int n, int mm, int nn, double delt1, const int* __restrict__ mask,
              const double* __restrict__ p, const double* __restrict__ dp, const double* __restrict__ dpf,
              const double* __restrict__ temp, const double* __restrict__ saln,
              const double* __restrict__ difint, const double* __restrict__ nslp,
              const double* __restrict__ pbf, const double* __restrict__ sc2, const double* __restrict__ scl,
              const double* __restrict__ scp2, const double* __restrict__ coriop,
              const double* __restrict__ hbl_tf, const double* __restrict__ wpup_tf,
              const double* __restrict__ hml_tfbnd, const double* __restrict__ util1,
              double* __restrict__ mfltd, double* __restrict__ mflsm_o, double* __restrict__ tfltd,
              double* __restrict__ tflsm, double* __restrict__ sfltd, double* __restrict__ sflsm,
              int* __restrict__ err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.ii) return;
  const long x = ix2(g, i, j);
  if (mask[x] != 1) return;
  const long xm = x - (DIR == 0 ? 1 : g.ldi);
  const long lev = g.lev;
  const int kk = g.kdm;
  const double ffac = .0625, fface = .99 * ffac, eps = 1.e-14, c5_21 = 5. / 21.;

  extern __shared__ double mfl_sm[];
  double mfl_loc[SMEM ? 1 : KM + 2];
#define mfl(k) (*(SMEM ? &mfl_sm[(k) * 128 + threadIdx.x] : &mfl_loc[SMEM ? 0 : (k)]))

  const double hml = .5 * (hml_tfbnd[xm] + hml_tfbnd[x]);
  // depth-invariant submesoscale transport component (:1130-1185)
  double upssm = 0.;
  if (P.mode == 2) {
    const double hbl = .5 * (hbl_tf[xm] + hbl_tf[x]);
    const double absf = .5 * fabs(coriop[xm] + coriop[x]);
    const double wpup = .5 * (wpup_tf[xm] + wpup_tf[x]);
    const double drho = util1[x] - util1[xm];
    upssm = P.csm * absf * hbl * hml * hml * drho / wpup;
  } else if (P.mode == 1) {
    const double f = .5 * (coriop[xm] + coriop[x]);
    const double absfi = 1. / sqrt(f * f + P.rtau * P.rtau);
    const double lfi = 1. / fmax(sqrt(P.dbcl82 * hml) * absfi, P.lfmin);
    const double drho = util1[x] - util1[xm];
    upssm = P.csm * hml * hml * drho * lfi * absfi;
  }
  const double ptf = fmax(p[xm], p[x]);
  const double mfleps = eps * epsilp * sc2[x];
  const double et2mf = -grav * rho0 * delt1 * scl[x];
  const double pb = pbf[x + (long)(n - 1) * lev];
  const double am = scp2[xm], ap = scp2[x];

  // last layer with mass and the interface pressure below it (:1222-1229)
  int kmax = 1;
  double pbot;
  {
    double pk = ptf;
    pbot = ptf;
    for (int k = 1; k <= kk; ++k) {
      const long o = (long)(k + nn - 1) * lev;
      pk = pk + dpf[x + o];
      if (k == 1 || dp[xm + o] > epsilp || dp[x + o] > epsilp) { kmax = k; pbot = pk; }
    }
  }
  const double pml = fmin(ptf + hml * onem, pbot);
  const double dpmli = 1. / (pml - ptf);
  // first interface below the mixed layer base (:1241-1248); puv is non-decreasing in k
  int kml = kmax + 1;
  {
    double pk = ptf;
    for (int k = 2; k <= kmax; ++k) {
      pk = pk + dpf[x + (long)(k - 1 + nn - 1) * lev];
      if (pk > pml) { kml = k; break; }
    }
  }
  // GM interface flux below the mixed layer (:1252-1256)
  auto gm_below = [&](int k) -> double {
    const long o1 = (long)(k - 2) * lev, o2 = (long)(k - 1) * lev;
    const double kappa = .25 * (difint[xm + o1] + difint[x + o1] + difint[xm + o2] + difint[x + o2]);
    return -kappa * nslp[x + o2] * et2mf;
  };
  const double gm_kml = kml <= kmax ? gm_below(kml) : 0.;   // mflgm(kmax+1) = 0
  // GM and submesoscale flux at interface k with pressure puv_k (:1252-1288)
  auto gm_sm = [&](int k, double puv_k, double& gm, double& sm) {
    if (k == 1 || k == kmax + 1) { gm = 0.; sm = 0.; }
    else if (k >= kml) { gm = k == kml ? gm_kml : gm_below(k); sm = 0.; }
    else {
      gm = gm_kml * (puv_k - ptf) * dpmli;
      double q = 2. * (ptf - puv_k) * dpmli + 1.;
      q = q * q;
      sm = -upssm * (1. - q) * (1. + c5_21 * q) * et2mf;
    }
  };
  {
    double pk = ptf;
    for (int k = 1; k <= kmax + 1; ++k) {
      double gm, sm;
      gm_sm(k, pk, gm, sm);
      mfl(k) = gm + sm;
      if (k <= kmax) pk = pk + dpf[x + (long)(k + nn - 1) * lev];
    }
  }
  // available thicknesses at the two scalar points (:1304-1309)
  auto dl_m = [&](int k) { return fmax(0., fmin(p[xm + (long)k * lev], pb) - fmax(p[xm + (long)(k - 1) * lev], ptf)); };
  auto dl_p = [&](int k) { return fmax(0., fmin(p[x + (long)k * lev], pb) - fmax(p[x + (long)(k - 1) * lev], ptf)); };

  // alternate downward/upward limiter sweeps (:1318-1394)
  bool changed = true;
  int niter = 0, kdir = 1;
  while (changed) {
    niter++;
    if (niter == 1000) { atomicMax(err, 1); break; }
    changed = false;
    kdir = -kdir;
    const int k0 = (1 + kdir + (1 - kdir) * kmax) / 2;
    for (int s = 0, k = k0; s < kmax; ++s, k += kdir) {
      const double lo = mfl(k), hi = mfl(k + 1);
      if (fabs(hi - lo) > fmax(mfleps, eps * fabs(hi + lo))) {
        const double dlm = dl_m(k), dlp = dl_p(k);
        if (hi - lo > ffac * fmax(epsilp, dlm) * am) {
          const double q = fface * dlm * am;
          if (hi > -lo) {
            if (lo > -.5 * q) mfl(k + 1) = lo + q;
            else { mfl(k + 1) = .5 * q; mfl(k) = -mfl(k + 1); }
          } else {
            if (hi < .5 * q) mfl(k) = hi - q;
            else { mfl(k) = -.5 * q; mfl(k + 1) = -mfl(k); }
          }
          changed = true;
        } else if (hi - lo < -ffac * fmax(epsilp, dlp) * ap) {
          const double q = fface * dlp * ap;
          if (hi < -lo) {
            if (lo < .5 * q) mfl(k + 1) = lo - q;
            else { mfl(k + 1) = -.5 * q; mfl(k) = -mfl(k + 1); }
          } else {
            if (hi > -.5 * q) mfl(k) = hi + q;
            else { mfl(k) = .5 * q; mfl(k + 1) = -mfl(k); }
          }
          changed = true;
        }
      }
    }
  }

  // split the limited total back into GM and submesoscale parts (:1398-1436), one interface at a time
  auto split = [&](int k, double puv_k, double& f, double& gm, double& sm) {
    gm_sm(k, puv_k, gm, sm);
    f = mfl(k);
    if (fabs(f) < mfleps) {
      f = 0.; gm = 0.; sm = 0.;
    } else if (f > 0.) {
      if (gm > sm) {
        if (f > 2. * sm) gm = f - sm; else { gm = .5 * f; sm = gm; }
      } else {
        if (f > 2. * gm) sm = f - gm; else { sm = .5 * f; gm = sm; }
      }
    } else {
      if (gm < sm) {
        if (f < 2. * sm) gm = f - sm; else { gm = .5 * f; sm = gm; }
      } else {
        if (f < 2. * gm) sm = f - gm; else { sm = .5 * f; gm = sm; }
      }
    }
  };

  // layer fluxes + heat/salt components (:1442-1468, :1876-1902)
  double pk = ptf, f_lo, gm_lo, sm_lo;
  split(1, pk, f_lo, gm_lo, sm_lo);
  for (int k = 1; k <= kk; ++k) {
    const long xk = x + (long)(k + mm - 1) * lev, xmk = xm + (long)(k + mm - 1) * lev;
    double fgm = 0., fsm = 0.;
    if (k <= kmax) {
      pk = pk + dpf[x + (long)(k + nn - 1) * lev];
      double f_hi, gm_hi, sm_hi;
      split(k + 1, pk, f_hi, gm_hi, sm_hi);
      if (fabs(f_hi - f_lo) > fmax(mfleps, eps * fabs(f_hi + f_lo))) {
        fgm = gm_hi - gm_lo;
        fsm = sm_hi - sm_lo;
      }
      if (fgm + fsm > ffac * fmax(epsilp, dl_m(k)) * am) atomicMax(err, 2);
      if (fgm + fsm < -ffac * fmax(epsilp, dl_p(k)) * ap) atomicMax(err, 3);
      f_lo = f_hi; gm_lo = gm_hi; sm_lo = sm_hi;
    }
    const double qt = .5 * (temp[xmk] + temp[xk]);
    const double qs = .5 * (saln[xmk] + saln[xk]);
    mfltd[xk] = fgm; mflsm_o[xk] = fsm;
    tfltd[xk] = fgm * qt; tflsm[xk] = fsm * qt;
    sfltd[xk] = fgm * qs; sflsm[xk] = fsm * qs;
  }
#undef mfl
}

}  // namespace

void eddtra_dev(int m, int n, int mm, int nn, int k1m, int k1n) {
  Ctx& c = C(); const Geom& g = c.g;
  if (c.option("vcoord", "cntiso_hybrid") == "isopyc_bulkml") {   // :1818-1857, eddtra_isopyc.cu
    eddtra_isopyc_dev(m, n, mm, nn, k1m, k1n);
    return;
  }
  if (c.option("eitmth", "gm") != "gm")
    throw std::runtime_error("(eddtra) eitmth_opt is unsupported for vcoord = 'cntiso_hybrid'!");
  if (g.kdm > KM) throw std::runtime_error("eddtra: kdm exceeds the compiled column bound (64)");
  const std::string mlrmth = c.option("mlrmth", "fox08");
  const double delt1 = c.scalar("delt1");
  MlParams P{};
  if (mlrmth == "none") P.mode = 0;
  else if (mlrmth == "fox08") P.mode = 1;
  else if (mlrmth == "bod23") P.mode = 2;
  else throw std::runtime_error(" init_eddtra: mlrmth = " + mlrmth + " is unsupported!");
  // namelist defaults, phy/mod_eddtra.F90:54-98; dbcl82 phy/mod_cmnfld.F90:48
  const double ce = c.scalar("ce", .06), cl = c.scalar("cl", .25), tau_mlr = c.scalar("tau_mlr", 86400.);
  const double tg_hbl = c.scalar("tau_growing_hbl", 300.), td_hbl = c.scalar("tau_decaying_hbl", 86400.);
  const double tg_hml = c.scalar("tau_growing_hml", 3600.), td_hml = c.scalar("tau_decaying_hml", 259200.);
  P.wg_hbl = tg_hbl / (tg_hbl + delt1); P.wd_hbl = td_hbl / (td_hbl + delt1);
  P.wg_hml = tg_hml / (tg_hml + delt1); P.wd_hml = td_hml / (td_hml + delt1);
  P.mstar = c.scalar("mstar", .5); P.nstar = c.scalar("nstar", .066);
  P.wpup_min = c.scalar("wpup_min", 1.e-3); P.mlbl_max_ratio = c.scalar("mlbl_max_ratio", 3.);
  P.lfmin = c.scalar("lfmin", 5.e3); P.dbcl82 = c.scalar("dbcl82", .0003);
  P.rtau = 1. / tau_mlr;
  P.csm = P.mode == 2 ? grav * alpha0 * ce / cl : grav * alpha0 * ce;

  double* hml_tfbnd = c.has("hml_tfbnd") ? c.dev("hml_tfbnd") : c.owned("hml_tfbnd", 1);
  double *hbl_tf = nullptr, *wpup_tf = nullptr, *util1 = nullptr, *coriop = nullptr;
  dim3 grid2(cdiv(g.ii, 128), g.jj);
  if (P.mode != 0) {
    hbl_tf = c.dev("hbl_tf");
    wpup_tf = P.mode == 2 ? c.dev("wpup_tf") : hbl_tf;
    util1 = c.dev("util1");
    coriop = c.dev("coriop");
    const double* u3 = P.mode == 2 ? c.dev("ustar3") : hbl_tf;
    const double* w3 = P.mode == 2 ? c.dev("wstar3") : hbl_tf;
    LAUNCH(eddtra_mlfilter, grid2, 128, 0, g, P, c.idev("ip"), c.dev("OBLdepth"), c.dev("mld"), u3, w3, hbl_tf,
           wpup_tf, c.dev("hml_tf1"), c.dev("hml_tf"), hml_tfbnd);
    if (P.mode == 2)
      halo_update(std::vector<HaloReq>{{hbl_tf, 1, halo_ps}, {wpup_tf, 1, halo_ps}, {hml_tfbnd, 1, halo_ps}}, 1, 1);
    else
      halo_update(hml_tfbnd, 1, 1, 1, halo_ps);
    LAUNCH(eddtra_mldens, grid2, 128, 0, g, eos::host_coef(), nn, c.idev("ip"), c.dev("p"), c.dev("dp"),
           c.dev("temp"), c.dev("saln"), hml_tfbnd, util1);
    halo_update(util1, 1, 1, 1, halo_ps);
  }
  int* err = c.error_flag();
  // eddtra_mfl = local (default) | smem: where the limiter's interface-flux column lives (see eddtra_column).
  // Measured at tnx0.25v4 (round 2, gpurun_out/r2d_kt_a.json): shared memory 7.1 + 7.9 ms (u + v), local array
  // 4.4 + 5.4 ms - the 56 KB per block leave 16 warps per SM against 64, and the kernel is latency-bound on its
  // global operands, not on the flux column.  Kept as a switch, off.
  const bool mfl_smem = c.option("eddtra_mfl", "local") == "smem";
  const size_t smem = mfl_smem ? (size_t)(g.kdm + 2) * 128 * sizeof(double) : 0;
#define EDDTRA_LAUNCH(SM_, OCC_)                                                                                       \
  do {                                                                                                                 \
    if (SM_) {                                                                                                         \
      static bool attr_set = false;                                                                                    \
      if (!attr_set) {                                                                                                 \
        CUDA_CHECK(cudaFuncSetAttribute(eddtra_column<0, OCC_, SM_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 128 * 8)); \
        CUDA_CHECK(cudaFuncSetAttribute(eddtra_column<1, OCC_, SM_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 128 * 8)); \
        attr_set = true;                                                                                               \
      }                                                                                                                \
    }                                                                                                                  \
    LAUNCH_NAMED("eddtra_column<u>", (eddtra_column<0, OCC_, SM_>), grid2, 128, smem, g, P, n, mm, nn, delt1, c.idev("iu"), \
                 c.dev("p"), c.dev("dp"), c.dev("dpu"), c.dev("temp"), c.dev("saln"), c.dev("difint"),                  \
                 c.dev("nslpx"), c.dev("pbu"), c.dev("scu2"), c.dev("scuy"), c.dev("scp2"), coriop, hbl_tf, wpup_tf,    \
                 hml_tfbnd, util1, c.dev("umfltd"), c.dev("umflsm"), c.dev("utfltd"), c.dev("utflsm"),                  \
                 c.dev("usfltd"), c.dev("usflsm"), err);                                                               \
    LAUNCH_NAMED("eddtra_column<v>", (eddtra_column<1, OCC_, SM_>), grid2, 128, smem, g, P, n, mm, nn, delt1, c.idev("iv"), \
                 c.dev("p"), c.dev("dp"), c.dev("dpv"), c.dev("temp"), c.dev("saln"), c.dev("difint"),                  \
                 c.dev("nslpy"), c.dev("pbv"), c.dev("scv2"), c.dev("scvx"), c.dev("scp2"), coriop, hbl_tf, wpup_tf,    \
                 hml_tfbnd, util1, c.dev("vmfltd"), c.dev("vmflsm"), c.dev("vtfltd"), c.dev("vtflsm"),                  \
                 c.dev("vsfltd"), c.dev("vsflsm"), err);                                                               \
  } while (0)
  if (mfl_smem) EDDTRA_LAUNCH(true, 4);      // 4 blocks of 56 KB per SM; the register budget is no longer the limit
  else { OCC_DISPATCH3("eddtra_minblk", 16, 9, 12, 16, EDDTRA_LAUNCH(false, OCC)); }
#undef EDDTRA_LAUNCH
  c.error_source = "(eddtra_ale) 1: no convergence, 2: flux exceeds +ffac*mass, 3: flux exceeds -ffac*mass";
}

}  // namespace blom
