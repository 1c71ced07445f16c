// ORACLE — TEST INFRASTRUCTURE ONLY (see core.hpp header).
// Restatement of phy/mod_eddtra.F90: rmeanfilt :121-151, eddtra_ale :1001-1739
// (hybrid coordinate, eitmth='gm', mlrmth none|fox08|bod23), the isopycnic bulk-mixed-layer
// variants eddtra_intdif_isopyc_bulkml :153-226 and eddtra_gm_isopyc_bulkml :228-999, and the
// heat/salt flux diagnosis of eddtra :1808-1928.  Namelist defaults :54-98.
#include "core.hpp"
#include "eos.hpp"

namespace orc {

namespace {

// phy/mod_eddtra.F90:121-151
inline void rmeanfilt(double& filtered, double signal, double wg, double wd) {
  double wf = signal >= filtered ? wg : wd;
  filtered = wf * filtered + (1. - wf) * signal;
}

struct FaceGeom {  // what differs between the u and the v face columns
  int di, dj;      // offset of the "minus" scalar point
};

// One face column: phy/mod_eddtra.F90:1208-1468 (u) / :1470-1730 (v).
// Returns 0, or an error code (1: no convergence, 2: '>' check, 3: '<' check).
int face_column(const Dims& d, int i, int j, int di, int dj, int n, int mm, int nn, double delt1,
                double sc2, double scl /*scuy|scvx*/, double ptf, double upssm, double hml,
                A3 dp, A3 dpf /*dpu|dpv*/, A3 p, A3 difint, A3 nslp, A2 scp2, double pbf /*pbu|pbv(i,j,n)*/,
                A3 mfltd, A3 mflsm_out) {
  (void)n;
  const int kk = d.kk;
  const double ffac = .0625, fface = .99 * ffac, eps = 1.e-14, c5_21 = 5. / 21.;
  std::vector<double> puv(kk + 2), mflgm(kk + 2), mflsm(kk + 2), mfl(kk + 2), dlm(kk + 1), dlp(kk + 1);
  const int im = i - di, jm = j - dj;
  for (int k = 1; k <= kk; ++k) { mfltd(i, j, k + mm) = 0.; mflsm_out(i, j, k + mm) = 0.; }
  const double mfleps = eps * epsilp * sc2;
  const double et2mf = -grav * rho0 * delt1 * scl;
  int kmax = 1;
  puv[1] = ptf;
  for (int k = 1; k <= kk; ++k) {
    const int kn = k + nn;
    puv[k + 1] = puv[k] + dpf(i, j, kn);
    if (dp(im, jm, kn) > epsilp || dp(i, j, kn) > epsilp) kmax = k;
  }
  const double pml = std::min(puv[1] + hml * onem, puv[kmax + 1]);
  const double dpmli = 1. / (pml - puv[1]);
  int kml = kmax + 1;
  for (int k = kmax; k >= 2; --k) {
    if (puv[k] > pml) kml = k; else break;
  }
  for (int k = kml; k <= kmax; ++k) {
    double kappa = .25 * (difint(im, jm, k - 1) + difint(i, j, k - 1) + difint(im, jm, k) + difint(i, j, k));
    mflgm[k] = -kappa * nslp(i, j, k) * et2mf;
  }
  mflgm[kmax + 1] = 0.;
  mflgm[1] = 0.;
  for (int k = 2; k <= kml - 1; ++k) mflgm[k] = mflgm[kml] * (puv[k] - puv[1]) * dpmli;
  mflsm[1] = 0.;
  for (int k = 2; k <= kml - 1; ++k) {
    double q = (2. * (puv[1] - puv[k]) * dpmli + 1.);
    q = q * q;
    mflsm[k] = -upssm * (1. - q) * (1. + c5_21 * q) * et2mf;
  }
  for (int k = kml; k <= kmax + 1; ++k) mflsm[k] = 0.;
  for (int k = 1; k <= kmax + 1; ++k) mfl[k] = mflgm[k] + mflsm[k];
  for (int k = 1; k <= kmax; ++k) {
    dlm[k] = std::max(0., std::min(p(im, jm, k + 1), pbf) - std::max(p(im, jm, k), ptf));
    dlp[k] = std::max(0., std::min(p(i, j, k + 1), pbf) - std::max(p(i, j, k), ptf));
  }
  const double am = scp2(im, jm), ap = scp2(i, j);
  bool changed = true;
  int niter = 0, kdir = 1;
  while (changed) {
    niter++;
    if (niter == 1000) return 1;
    changed = false;
    kdir = -kdir;
    const int k0 = (1 + kdir + (1 - kdir) * kmax) / 2, k1 = (1 - kdir + (1 + kdir) * kmax) / 2;
    for (int k = k0; kdir > 0 ? k <= k1 : k >= k1; k += kdir) {
      if (std::fabs(mfl[k + 1] - mfl[k]) > std::max(mfleps, eps * std::fabs(mfl[k + 1] + mfl[k]))) {
        if (mfl[k + 1] - mfl[k] > ffac * std::max(epsilp, dlm[k]) * am) {
          double q = fface * dlm[k] * am;
          if (mfl[k + 1] > -mfl[k]) {
            if (mfl[k] > -.5 * q) mfl[k + 1] = mfl[k] + q;
            else { mfl[k + 1] = .5 * q; mfl[k] = -mfl[k + 1]; }
          } else {
            if (mfl[k + 1] < .5 * q) mfl[k] = mfl[k + 1] - q;
            else { mfl[k] = -.5 * q; mfl[k + 1] = -mfl[k]; }
          }
          changed = true;
        } else if (mfl[k + 1] - mfl[k] < -ffac * std::max(epsilp, dlp[k]) * ap) {
          double q = fface * dlp[k] * ap;
          if (mfl[k + 1] < -mfl[k]) {
            if (mfl[k] < .5 * q) mfl[k + 1] = mfl[k] - q;
            else { mfl[k + 1] = -.5 * q; mfl[k] = -mfl[k + 1]; }
          } else {
            if (mfl[k + 1] > -.5 * q) mfl[k] = mfl[k + 1] + q;
            else { mfl[k] = .5 * q; mfl[k + 1] = -mfl[k]; }
          }
          changed = true;
        }
      }
    }
  }
  for (int k = 1; k <= kmax + 1; ++k) {
    if (std::fabs(mfl[k]) < mfleps) {
      mfl[k] = 0.; mflgm[k] = 0.; mflsm[k] = 0.;
    } else if (mfl[k] > 0.) {
      if (mflgm[k] > mflsm[k]) {
        if (mfl[k] > 2. * mflsm[k]) mflgm[k] = mfl[k] - mflsm[k];
        else { mflgm[k] = .5 * mfl[k]; mflsm[k] = mflgm[k]; }
      } else {
        if (mfl[k] > 2. * mflgm[k]) mflsm[k] = mfl[k] - mflgm[k];
        else { mflsm[k] = .5 * mfl[k]; mflgm[k] = mflsm[k]; }
      }
    } else {
      if (mflgm[k] < mflsm[k]) {
        if (mfl[k] < 2. * mflsm[k]) mflgm[k] = mfl[k] - mflsm[k];
        else { mflgm[k] = .5 * mfl[k]; mflsm[k] = mflgm[k]; }
      } else {
        if (mfl[k] < 2. * mflgm[k]) mflsm[k] = mfl[k] - mflgm[k];
        else { mflsm[k] = .5 * mfl[k]; mflgm[k] = mflsm[k]; }
      }
    }
  }
  for (int k = 1; k <= kmax; ++k) {
    const int km = k + mm;
    if (std::fabs(mfl[k + 1] - mfl[k]) > std::max(mfleps, eps * std::fabs(mfl[k + 1] + mfl[k]))) {
      mfltd(i, j, km) = mflgm[k + 1] - mflgm[k];
      mflsm_out(i, j, km) = mflsm[k + 1] - mflsm[k];
    } else {
      mfltd(i, j, km) = 0.;
      mflsm_out(i, j, km) = 0.;
    }
    if (mfltd(i, j, km) + mflsm_out(i, j, km) > ffac * std::max(epsilp, dlm[k]) * am) return 2;
    if (mfltd(i, j, km) + mflsm_out(i, j, km) < -ffac * std::max(epsilp, dlp[k]) * ap) return 3;
  }
  return 0;
}

// phy/mod_eddtra.F90:1001-1739
void eddtra_ale(int m, int n, int mm, int nn, int k1m, int k1n) {
  (void)m; (void)k1m; (void)k1n;
  Oracle& o = O(); const Dims& d = o.d;
  const int ii = d.ii, jj = d.jj, kk = d.kk;
  const double delt1 = o.scalar("delt1");
  const std::string mlrmth = o.option("mlrmth", "fox08");
  // namelist defaults, phy/mod_eddtra.F90:54-98
  const double ce = o.scalar("ce", .06), cl = o.scalar("cl", .25), tau_mlr = o.scalar("tau_mlr", 86400.),
               tau_growing_hbl = o.scalar("tau_growing_hbl", 300.),
               tau_decaying_hbl = o.scalar("tau_decaying_hbl", 86400.),
               tau_growing_hml = o.scalar("tau_growing_hml", 3600.),
               tau_decaying_hml = o.scalar("tau_decaying_hml", 259200.), lfmin = o.scalar("lfmin", 5.e3),
               mstar = o.scalar("mstar", .5), nstar = o.scalar("nstar", .066),
               wpup_min = o.scalar("wpup_min", 1.e-3), mlbl_max_ratio = o.scalar("mlbl_max_ratio", 3.),
               dbcl82 = o.scalar("dbcl82", .0003);  // phy/mod_cmnfld.F90:48
  const double c2_3 = 2. / 3.;
  I2 ip = o.i2("ip"), iu = o.i2("iu"), iv = o.i2("iv");
  A3 p = o.a3("p"), dp = o.a3("dp"), dpu = o.a3("dpu"), dpv = o.a3("dpv"), temp = o.a3("temp"),
     saln = o.a3("saln"), difint = o.a3("difint"), nslpx = o.a3("nslpx"), nslpy = o.a3("nslpy"),
     pbu = o.a3("pbu"), pbv = o.a3("pbv");
  A3 umfltd = o.a3("umfltd"), vmfltd = o.a3("vmfltd"), umflsm = o.a3("umflsm"), vmflsm = o.a3("vmflsm");
  A2 scu2 = o.a2("scu2"), scv2 = o.a2("scv2"), scuy = o.a2("scuy"), scvx = o.a2("scvx"), scp2 = o.a2("scp2");
  A2 upssmx = o.scratch("_eddtra_upssmx", 1).level(1), upssmy = o.scratch("_eddtra_upssmy", 1).level(1);
  A2 ptu = o.scratch("_eddtra_ptu", 1).level(1), ptv = o.scratch("_eddtra_ptv", 1).level(1);
  A2 hml_tfbnd = o.has("hml_tfbnd") ? o.a2("hml_tfbnd") : o.scratch("hml_tfbnd", 1).level(1);

  if (mlrmth == "none") {
    for (int j = 1; j <= jj; ++j)
      for (int i = 1; i <= ii; ++i) {
        if (iu(i, j) == 1) upssmx(i, j) = 0.;
        if (iv(i, j) == 1) upssmy(i, j) = 0.;
      }
  } else {
    A2 hbl_tf = o.a2("hbl_tf"), wpup_tf = o.a2("wpup_tf"), hml_tf1 = o.a2("hml_tf1"), hml_tf = o.a2("hml_tf");
    A2 OBLdepth = o.a2("OBLdepth"), mld = o.a2("mld"), util1 = o.a2("util1"), coriop = o.a2("coriop");
    const double wf_growing_hbl = tau_growing_hbl / (tau_growing_hbl + delt1);
    const double wf_decaying_hbl = tau_decaying_hbl / (tau_decaying_hbl + delt1);
    const double wf_growing_hml = tau_growing_hml / (tau_growing_hml + delt1);
    const double wf_decaying_hml = tau_decaying_hml / (tau_decaying_hml + delt1);
    if (mlrmth == "bod23") {
      A2 ustar3 = o.a2("ustar3"), wstar3 = o.a2("wstar3");
      for (int j = 1; j <= jj; ++j)
        for (int i = 1; i <= ii; ++i) {
          if (ip(i, j) != 1) continue;
          double hbl = OBLdepth(i, j);
          double wpup = std::max(wpup_min, std::pow(mstar * ustar3(i, j) + nstar * wstar3(i, j), c2_3));
          rmeanfilt(hbl_tf(i, j), hbl, wf_growing_hbl, wf_decaying_hbl);
          rmeanfilt(wpup_tf(i, j), wpup, wf_growing_hbl, wf_decaying_hbl);
          rmeanfilt(hml_tf1(i, j), mld(i, j), wf_growing_hbl, wf_decaying_hbl);
          rmeanfilt(hml_tf(i, j), hml_tf1(i, j), wf_growing_hml, wf_decaying_hml);
          hml_tfbnd(i, j) = std::min(hml_tf(i, j), mlbl_max_ratio * hbl_tf(i, j));
        }
      xctilr(hbl_tf, 1, 1, halo_ps);
      xctilr(wpup_tf, 1, 1, halo_ps);
      xctilr(hml_tfbnd, 1, 1, halo_ps);
    } else if (mlrmth == "fox08") {
      for (int j = 1; j <= jj; ++j)
        for (int i = 1; i <= ii; ++i) {
          if (ip(i, j) != 1) continue;
          double hbl = OBLdepth(i, j);
          rmeanfilt(hbl_tf(i, j), hbl, wf_growing_hbl, wf_decaying_hbl);
          rmeanfilt(hml_tf1(i, j), mld(i, j), wf_growing_hbl, wf_decaying_hbl);
          rmeanfilt(hml_tf(i, j), hml_tf1(i, j), wf_growing_hml, wf_decaying_hml);
          hml_tfbnd(i, j) = std::min(hml_tf(i, j), mlbl_max_ratio * hbl_tf(i, j));
        }
      xctilr(hml_tfbnd, 1, 1, halo_ps);
    } else {
      throw std::runtime_error(" init_eddtra: mlrmth = " + mlrmth + " is unsupported!");
    }
    // vertically averaged mixed layer density (:1105-1127)
    for (int j = 1; j <= jj; ++j)
      for (int i = 1; i <= ii; ++i) {
        if (ip(i, j) != 1) continue;
        double pml = std::min(p(i, j, 1) + hml_tfbnd(i, j) * onem, p(i, j, kk + 1));
        double dpmli = 1. / (pml - p(i, j, 1));
        double tmldp = 0., smldp = 0.;
        for (int k = 1; k <= kk; ++k) {
          const int kn = k + nn;
          if (p(i, j, k + 1) < pml) {
            tmldp = tmldp + temp(i, j, kn) * dp(i, j, kn);
            smldp = smldp + saln(i, j, kn) * dp(i, j, kn);
          } else {
            tmldp = tmldp + temp(i, j, kn) * (pml - p(i, j, k));
            smldp = smldp + saln(i, j, kn) * (pml - p(i, j, k));
            break;
          }
        }
        util1(i, j) = eos::sig0(tmldp * dpmli, smldp * dpmli);
      }
    xctilr(util1, 1, 1, halo_ps);
    if (mlrmth == "bod23") {
      const double csm = grav * alpha0 * ce / cl;
      for (int j = 1; j <= jj; ++j)
        for (int i = 1; i <= ii; ++i) {
          if (iu(i, j) == 1) {
            double hbl = .5 * (hbl_tf(i - 1, j) + hbl_tf(i, j));
            double hml = .5 * (hml_tfbnd(i - 1, j) + hml_tfbnd(i, j));
            double absf = .5 * std::fabs(coriop(i - 1, j) + coriop(i, j));
            double wpup = .5 * (wpup_tf(i - 1, j) + wpup_tf(i, j));
            double drho = util1(i, j) - util1(i - 1, j);
            upssmx(i, j) = csm * absf * hbl * hml * hml * drho / wpup;
          }
          if (iv(i, j) == 1) {
            double hbl = .5 * (hbl_tf(i, j - 1) + hbl_tf(i, j));
            double hml = .5 * (hml_tfbnd(i, j - 1) + hml_tfbnd(i, j));
            double absf = .5 * std::fabs(coriop(i, j - 1) + coriop(i, j));
            double wpup = .5 * (wpup_tf(i, j - 1) + wpup_tf(i, j));
            double drho = util1(i, j) - util1(i, j - 1);
            upssmy(i, j) = csm * absf * hbl * hml * hml * drho / wpup;
          }
        }
    } else {
      const double rtau = 1. / tau_mlr, csm = grav * alpha0 * ce;
      for (int j = 1; j <= jj; ++j)
        for (int i = 1; i <= ii; ++i) {
          if (iu(i, j) == 1) {
            double hml = .5 * (hml_tfbnd(i - 1, j) + hml_tfbnd(i, j));
            double f = .5 * (coriop(i - 1, j) + coriop(i, j));
            double absfi = 1. / std::sqrt(f * f + rtau * rtau);
            double lfi = 1. / std::max(std::sqrt(dbcl82 * hml) * absfi, lfmin);
            double drho = util1(i, j) - util1(i - 1, j);
            upssmx(i, j) = csm * hml * hml * drho * lfi * absfi;
          }
          if (iv(i, j) == 1) {
            double hml = .5 * (hml_tfbnd(i, j - 1) + hml_tfbnd(i, j));
            double f = .5 * (coriop(i, j - 1) + coriop(i, j));
            double absfi = 1. / std::sqrt(f * f + rtau * rtau);
            double lfi = 1. / std::max(std::sqrt(dbcl82 * hml) * absfi, lfmin);
            double drho = util1(i, j) - util1(i, j - 1);
            upssmy(i, j) = csm * hml * hml * drho * lfi * absfi;
          }
        }
    }
  }
  // top pressure at velocity points (:1191-1205)
  for (int j = 1; j <= jj; ++j)
    for (int i = 1; i <= ii; ++i) {
      if (iu(i, j) == 1) ptu(i, j) = std::max(p(i - 1, j, 1), p(i, j, 1));
      if (iv(i, j) == 1) ptv(i, j) = std::max(p(i, j - 1, 1), p(i, j, 1));
    }
  int err = 0, ei = 0, ej = 0; char ec = ' ';
#pragma omp parallel for
  for (int j = 1; j <= jj; ++j) {
    for (int i = 1; i <= ii; ++i) {
      if (iu(i, j) != 1) continue;
      double hml = .5 * (hml_tfbnd(i - 1, j) + hml_tfbnd(i, j));
      int e = face_column(d, i, j, 1, 0, n, mm, nn, delt1, scu2(i, j), scuy(i, j), ptu(i, j), upssmx(i, j), hml,
                          dp, dpu, p, difint, nslpx, scp2, pbu(i, j, n), umfltd, umflsm);
      if (e) {
#pragma omp critical
        { err = e; ei = i; ej = j; ec = 'u'; }
      }
    }
    for (int i = 1; i <= ii; ++i) {
      if (iv(i, j) != 1) continue;
      double hml = .5 * (hml_tfbnd(i, j - 1) + hml_tfbnd(i, j));
      int e = face_column(d, i, j, 0, 1, n, mm, nn, delt1, scv2(i, j), scvx(i, j), ptv(i, j), upssmy(i, j), hml,
                          dp, dpv, p, difint, nslpy, scp2, pbv(i, j, n), vmfltd, vmflsm);
      if (e) {
#pragma omp critical
        { err = e; ei = i; ej = j; ec = 'v'; }
      }
    }
  }
  if (err) {
    const char* what = err == 1 ? "no convergence " : (err == 2 ? "eddtra_ale > " : "eddtra_ale < ");
    throw std::runtime_error(std::string("(eddtra_ale) ") + what + ec + " at " + std::to_string(ei + d.i0) + "," +
                             std::to_string(ej + d.j0));
  }
}

// One face column of eddtra_gm_isopyc_bulkml: phy/mod_eddtra.F90:262-636 (u) / :640-998 (v).
// "m" is the scalar point on the minus side (i-1,j) or (i,j-1), "p" the point (i,j).
// Returns 0, or an error code (1: no convergence, 2: '>' check, 3: '<' check).
int face_column_isopyc(const Dims& d, int i, int j, int di, int dj, int n, int mm, int nn, double delt1,
                       double sc2, double scl /*scuy|scvx*/, double ptf, A3 dp, A3 dpf /*dpu|dpv*/, A3 p,
                       A3 temp, A3 saln, A3 difint, A3 nslp, A2 scp2, double pbf, I2 kfpla_n, A3 mfltd) {
  (void)n;
  using eos::rho;
  const int kk = d.kk;
  const double ffac = .0625, fface = .99 * ffac, eps = 1.e-14;
  std::vector<double> upsilon(kk + 2), mfl(kk + 2), dlm(kk + 1), dlp(kk + 1);
  const int im = i - di, jm = j - dj;
  for (int k = 1; k <= kk; ++k) mfltd(i, j, k + mm) = 0.;
  const double et2mf = -grav * rho0 * delt1 * scl;
  int kmax = 1;
  for (int k = 3; k <= kk; ++k) {
    const int kn = k + nn;
    if (dp(im, jm, kn) > epsilp || dp(i, j, kn) > epsilp) kmax = k;
  }
  const int kfm = kfpla_n(im, jm), kfp = kfpla_n(i, j);
  int kintr, kmin, km, kn;
  double kappa;
  if (kfm > kk && kfp > kk) {
    return 0;                                       // case 1
  } else if (kfm <= kk && kfp > kk) {               // case 2
    km = 2 + nn;
    kintr = kfm;
    kn = kintr + nn;
    while (rho(p(i, j, 3), temp(im, jm, kn), saln(im, jm, kn)) < rho(p(i, j, 3), temp(i, j, km), saln(i, j, km)) ||
           dp(im, jm, kn) < epsilp) {
      kintr = kintr + 1;
      if (kintr == kmax + 1) break;
      kn = kintr + nn;
    }
    if (kintr == kmax + 1) return 0;
    kappa = .5 * (difint(im, jm, 2) + difint(i, j, 2));
    upsilon[3] = -kappa * nslp(i, j, 3);
    if (upsilon[3] <= 0.) return 0;
    kmin = kintr - 1;
    mfl[kmin] = 0.;
    mfl[kintr] = et2mf * upsilon[3];
    for (int k = kintr + 1; k <= kmax + 1; ++k) mfl[k] = 0.;
  } else if (kfm > kk && kfp <= kk) {               // case 3
    km = 2 + nn;
    kintr = kfp;
    kn = kintr + nn;
    while (rho(p(im, jm, 3), temp(i, j, kn), saln(i, j, kn)) < rho(p(im, jm, 3), temp(im, jm, km), saln(im, jm, km)) ||
           dp(i, j, kn) < epsilp) {
      kintr = kintr + 1;
      if (kintr == kmax + 1) break;
      kn = kintr + nn;
    }
    if (kintr == kmax + 1) return 0;
    kappa = .5 * (difint(im, jm, 2) + difint(i, j, 2));
    upsilon[3] = -kappa * nslp(i, j, 3);
    if (upsilon[3] >= 0.) return 0;
    kmin = kintr - 1;
    mfl[kmin] = 0.;
    mfl[kintr] = et2mf * upsilon[3];
    for (int k = kintr + 1; k <= kmax + 1; ++k) mfl[k] = 0.;
  } else {                                          // case 4
    kintr = std::max(kfm, kfp);
    kappa = .5 * (difint(im, jm, 2) + difint(i, j, 2));
    upsilon[3] = -kappa * nslp(i, j, 3);
    for (int k = kintr + 1; k <= kmax; ++k) {
      kappa = .25 * (difint(im, jm, k - 1) + difint(i, j, k - 1) + difint(im, jm, k) + difint(i, j, k));
      upsilon[k] = -kappa * nslp(i, j, k);
    }
    upsilon[kmax + 1] = 0.;
    km = 2 + nn;
    kn = kintr - 1 + nn;
    if ((kfm < kintr && upsilon[3] - upsilon[kintr + 1] > 0. &&
         rho(p(i, j, 3), temp(im, jm, kn), saln(im, jm, kn)) > rho(p(i, j, 3), temp(i, j, km), saln(i, j, km))) ||
        (kfp < kintr && upsilon[3] - upsilon[kintr + 1] < 0. &&
         rho(p(im, jm, 3), temp(i, j, kn), saln(i, j, kn)) > rho(p(im, jm, 3), temp(im, jm, km), saln(im, jm, km)))) {
      kintr = kintr - 1;
      upsilon[kintr + 1] = upsilon[kintr + 2];
    }
    kmin = kintr - 1;
    mfl[kmin] = 0.;
    mfl[kintr] = et2mf * upsilon[3];
    for (int k = kintr + 1; k <= kmax; ++k) mfl[k] = et2mf * upsilon[k];
    mfl[kmax + 1] = 0.;
  }
  const double am = scp2(im, jm), ap = scp2(i, j);
  dlm[kmin] = std::max(0., std::min(p(im, jm, 3), pbf) - std::max(p(im, jm, 1), ptf));
  dlp[kmin] = std::max(0., std::min(p(i, j, 3), pbf) - std::max(p(i, j, 1), ptf));
  for (int k = kintr; k <= kmax; ++k) {
    dlm[k] = std::max(0., std::min(p(im, jm, k + 1), pbf) - std::max(p(im, jm, k), ptf));
    dlp[k] = std::max(0., std::min(p(i, j, k + 1), pbf) - std::max(p(i, j, k), ptf));
  }
  const double fhi = fface * std::max(0., std::min((p(im, jm, 3) - ptf) * am, (pbf - p(i, j, kintr)) * ap));
  const double flo = -fface * std::max(0., std::min((p(i, j, 3) - ptf) * ap, (pbf - p(im, jm, kintr)) * am));
  mfl[kmin + 1] = std::min(fhi, std::max(flo, mfl[kmin + 1]));
  for (int k = kmin + 1; k <= kmax - 1; ++k) {
    if (mfl[k + 1] - mfl[k] > ffac * std::max(epsilp, dlm[k]) * am) mfl[k + 1] = mfl[k] + fface * dlm[k] * am;
    else if (mfl[k + 1] - mfl[k] < -ffac * std::max(epsilp, dlp[k]) * ap) mfl[k + 1] = mfl[k] - fface * dlp[k] * ap;
    else break;
  }
  auto signif = [&](int k) {
    return std::fabs(mfl[k + 1] - mfl[k]) > eps * std::max(epsilp * sc2, std::fabs(mfl[k + 1] + mfl[k]));
  };
  bool changed = true;
  int niter = 0, kdir = 1;
  while (changed) {
    niter++;
    if (niter == 1000) return 1;
    changed = false;
    kdir = -kdir;
    const int k0 = ((1 - kdir) * kmax + (1 + kdir) * kmin) / 2, k1 = ((1 - kdir) * kmin + (1 + kdir) * kmax) / 2;
    for (int k = k0; kdir > 0 ? k <= k1 : k >= k1; k += kdir) {
      if (signif(k)) {
        if (mfl[k + 1] - mfl[k] > ffac * std::max(epsilp, dlm[k]) * am) {
          double q = fface * dlm[k] * am;
          if (mfl[k + 1] > -mfl[k]) {
            if (mfl[k] > -.5 * q) mfl[k + 1] = mfl[k] + q;
            else { mfl[k + 1] = .5 * q; mfl[k] = -mfl[k + 1]; }
          } else {
            if (mfl[k + 1] < .5 * q) mfl[k] = mfl[k + 1] - q;
            else { mfl[k] = -.5 * q; mfl[k + 1] = -mfl[k]; }
          }
          changed = true;
        } else if (mfl[k + 1] - mfl[k] < -ffac * std::max(epsilp, dlp[k]) * ap) {
          double q = fface * dlp[k] * ap;
          if (mfl[k + 1] < -mfl[k]) {
            if (mfl[k] < .5 * q) mfl[k + 1] = mfl[k] - q;
            else { mfl[k + 1] = -.5 * q; mfl[k] = -mfl[k + 1]; }
          } else {
            if (mfl[k + 1] > -.5 * q) mfl[k] = mfl[k + 1] + q;
            else { mfl[k] = .5 * q; mfl[k + 1] = -mfl[k]; }
          }
          changed = true;
        }
      }
    }
  }
  // final mass fluxes (:583-633): the mixed-layer flux is split over layers 1 and 2 by dpu
  if (signif(kmin)) {
    mfltd(i, j, 2 + mm) = mfl[kmin + 1] - mfl[kmin];
    mfltd(i, j, 1 + mm) = mfltd(i, j, 2 + mm) * dpf(i, j, 1 + nn) / (dpf(i, j, 1 + nn) + dpf(i, j, 2 + nn));
    mfltd(i, j, 2 + mm) = mfltd(i, j, 2 + mm) - mfltd(i, j, 1 + mm);
  } else {
    mfltd(i, j, 1 + mm) = 0.;
    mfltd(i, j, 2 + mm) = 0.;
  }
  for (int k = kintr; k <= kmax; ++k) {
    const int kmk = k + mm;
    if (signif(k)) mfltd(i, j, kmk) = mfl[k + 1] - mfl[k];
    else mfltd(i, j, kmk) = 0.;
    if (mfltd(i, j, kmk) > ffac * std::max(epsilp, dlm[k]) * am) return 2;
    if (mfltd(i, j, kmk) < -ffac * std::max(epsilp, dlp[k]) * ap) return 3;
  }
  return 0;
}

// phy/mod_eddtra.F90:228-999
void eddtra_gm_isopyc_bulkml(int m, int n, int mm, int nn, int k1m, int k1n) {
  (void)m; (void)k1m; (void)k1n;
  Oracle& o = O(); const Dims& d = o.d;
  const int ii = d.ii, jj = d.jj;
  const double delt1 = o.scalar("delt1");
  I2 iu = o.i2("iu"), iv = o.i2("iv");
  A3 p = o.a3("p"), dp = o.a3("dp"), dpu = o.a3("dpu"), dpv = o.a3("dpv"), temp = o.a3("temp"), saln = o.a3("saln");
  A3 difint = o.a3("difint"), nslpx = o.a3("nslpx"), nslpy = o.a3("nslpy"), pbu = o.a3("pbu"), pbv = o.a3("pbv");
  A3 umfltd = o.a3("umfltd"), vmfltd = o.a3("vmfltd");
  A2 scp2 = o.a2("scp2"), scu2 = o.a2("scu2"), scv2 = o.a2("scv2"), scuy = o.a2("scuy"), scvx = o.a2("scvx");
  I2 kf{o.fi.at("kfpla").p + (size_t)(n - 1) * d.lev, d.ldi, d.nbdy};
  A2 ptu = o.scratch("_ptu", 1).level(1), ptv = o.scratch("_ptv", 1).level(1);
  for (int j = 1; j <= jj; ++j) {
    for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) ptu(i, j) = std::max(p(i - 1, j, 1), p(i, j, 1));
    for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) ptv(i, j) = std::max(p(i, j - 1, 1), p(i, j, 1));
  }
  int err = 0, ei = 0, ej = 0; char ec = ' ';
  for (int j = 1; j <= jj; ++j) {
    for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) {
      int e = face_column_isopyc(d, i, j, 1, 0, n, mm, nn, delt1, scu2(i, j), scuy(i, j), ptu(i, j), dp, dpu, p, temp,
                                 saln, difint, nslpx, scp2, pbu(i, j, n), kf, umfltd);
      if (e && !err) { err = e; ei = i; ej = j; ec = 'u'; }
    }
    for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) {
      int e = face_column_isopyc(d, i, j, 0, 1, n, mm, nn, delt1, scv2(i, j), scvx(i, j), ptv(i, j), dp, dpv, p, temp,
                                 saln, difint, nslpy, scp2, pbv(i, j, n), kf, vmfltd);
      if (e && !err) { err = e; ei = i; ej = j; ec = 'v'; }
    }
  }
  if (err) {
    const char* what = err == 1 ? "no convergence " : (err == 2 ? "eddtra_gm_isopyc_bulkml > " : "eddtra_gm_isopyc_bulkml < ");
    throw std::runtime_error(std::string("(eddtra_gm_isopyc_bulkml) ") + what + ec + " at " +
                             std::to_string(ei + d.i0) + "," + std::to_string(ej + d.j0));
  }
}

// phy/mod_eddtra.F90:153-226
void eddtra_intdif_isopyc_bulkml(int m, int n, int mm, int nn, int k1m, int k1n) {
  (void)m; (void)n; (void)k1m; (void)k1n;
  Oracle& o = O(); const Dims& d = o.d;
  const int ii = d.ii, jj = d.jj, kk = d.kk;
  const double delt1 = o.scalar("delt1");
  I2 iu = o.i2("iu"), iv = o.i2("iv");
  A3 p = o.a3("p"), dp = o.a3("dp"), difint = o.a3("difint"), umfltd = o.a3("umfltd"), vmfltd = o.a3("vmfltd");
  A2 scp2 = o.a2("scp2"), scuy = o.a2("scuy"), scuxi = o.a2("scuxi"), scvx = o.a2("scvx"), scvyi = o.a2("scvyi");
  for (int j = 1; j <= jj; ++j) {
    for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) {
      umfltd(i, j, 1 + mm) = 0.; umfltd(i, j, 2 + mm) = 0.; umfltd(i, j, 3 + mm) = 0.;
    }
    for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) {
      vmfltd(i, j, 1 + mm) = 0.; vmfltd(i, j, 2 + mm) = 0.; vmfltd(i, j, 3 + mm) = 0.;
    }
  }
  for (int k = 4; k <= kk; ++k) {
    const int km = k + mm, kn = k + nn;
    for (int j = 1; j <= jj; ++j) {
      for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) {
        double flxhi = .125 * std::min(dp(i - 1, j, kn - 1) * scp2(i - 1, j), dp(i, j, kn) * scp2(i, j));
        double flxlo = -.125 * std::min(dp(i, j, kn - 1) * scp2(i, j), dp(i - 1, j, kn) * scp2(i - 1, j));
        double q = .25 * (difint(i - 1, j, k - 1) + difint(i, j, k - 1) + difint(i - 1, j, k) + difint(i, j, k));
        q = std::min(flxhi, std::max(flxlo, delt1 * q * (p(i - 1, j, k) - p(i, j, k)) * scuy(i, j) * scuxi(i, j)));
        umfltd(i, j, km - 1) = umfltd(i, j, km - 1) + q;
        umfltd(i, j, km) = -q;
      }
      for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) {
        double flxhi = .125 * std::min(dp(i, j - 1, kn - 1) * scp2(i, j - 1), dp(i, j, kn) * scp2(i, j));
        double flxlo = -.125 * std::min(dp(i, j, kn - 1) * scp2(i, j), dp(i, j - 1, kn) * scp2(i, j - 1));
        double q = .25 * (difint(i, j - 1, k - 1) + difint(i, j, k - 1) + difint(i, j - 1, k) + difint(i, j, k));
        q = std::min(flxhi, std::max(flxlo, delt1 * q * (p(i, j - 1, k) - p(i, j, k)) * scvx(i, j) * scvyi(i, j)));
        vmfltd(i, j, km - 1) = vmfltd(i, j, km - 1) + q;
        vmfltd(i, j, km) = -q;
      }
    }
  }
}

}  // namespace

// phy/mod_eddtra.F90:1808-1928
void eddtra(int m, int n, int mm, int nn, int k1m, int k1n) {
  Oracle& o = O(); const Dims& d = o.d;
  const int ii = d.ii, jj = d.jj, kk = d.kk;
  I2 iu = o.i2("iu"), iv = o.i2("iv");
  A3 temp = o.a3("temp"), saln = o.a3("saln");
  A3 umfltd = o.a3("umfltd"), vmfltd = o.a3("vmfltd");
  A3 utfltd = o.a3("utfltd"), vtfltd = o.a3("vtfltd"), usfltd = o.a3("usfltd"), vsfltd = o.a3("vsfltd");
  const std::string eitmth = o.option("eitmth", "gm");
  if (o.option("vcoord", "cntiso_hybrid") == "isopyc_bulkml") {  // :1818-1857
    if (eitmth == "intdif") eddtra_intdif_isopyc_bulkml(m, n, mm, nn, k1m, k1n);
    else if (eitmth == "gm") eddtra_gm_isopyc_bulkml(m, n, mm, nn, k1m, k1n);
    else throw std::runtime_error("(eddtra) eitmth_opt is unsupported for vcoord = 'isopyc_bulkml'!");
    for (int j = 1; j <= jj; ++j)
      for (int k = 1; k <= kk; ++k) {
        const int km = k + mm;
        for (int i = 1; i <= ii; ++i) if (iu(i, j) == 1) {
          utfltd(i, j, km) = .5 * umfltd(i, j, km) * (temp(i - 1, j, km) + temp(i, j, km));
          usfltd(i, j, km) = .5 * umfltd(i, j, km) * (saln(i - 1, j, km) + saln(i, j, km));
        }
        for (int i = 1; i <= ii; ++i) if (iv(i, j) == 1) {
          vtfltd(i, j, km) = .5 * vmfltd(i, j, km) * (temp(i, j - 1, km) + temp(i, j, km));
          vsfltd(i, j, km) = .5 * vmfltd(i, j, km) * (saln(i, j - 1, km) + saln(i, j, km));
        }
      }
    return;
  }
  if (eitmth != "gm")
    throw std::runtime_error("(eddtra) eitmth_opt is unsupported for vcoord = 'cntiso_hybrid'!");
  eddtra_ale(m, n, mm, nn, k1m, k1n);
  A3 umflsm = o.a3("umflsm"), vmflsm = o.a3("vmflsm"), utflsm = o.a3("utflsm"), vtflsm = o.a3("vtflsm");
  A3 usflsm = o.a3("usflsm"), vsflsm = o.a3("vsflsm");
  for (int j = 1; j <= jj; ++j)
    for (int k = 1; k <= kk; ++k) {
      const int km = k + mm;
      for (int i = 1; i <= ii; ++i) {
        if (iu(i, j) == 1) {
          double q = .5 * (temp(i - 1, j, km) + temp(i, j, km));
          utfltd(i, j, km) = umfltd(i, j, km) * q;
          utflsm(i, j, km) = umflsm(i, j, km) * q;
          q = .5 * (saln(i - 1, j, km) + saln(i, j, km));
          usfltd(i, j, km) = umfltd(i, j, km) * q;
          usflsm(i, j, km) = umflsm(i, j, km) * q;
        }
        if (iv(i, j) == 1) {
          double q = .5 * (temp(i, j - 1, km) + temp(i, j, km));
          vtfltd(i, j, km) = vmfltd(i, j, km) * q;
          vtflsm(i, j, km) = vmflsm(i, j, km) * q;
          q = .5 * (saln(i, j - 1, km) + saln(i, j, km));
          vsfltd(i, j, km) = vmfltd(i, j, km) * q;
          vsflsm(i, j, km) = vmflsm(i, j, km) * q;
        }
      }
    }
}

}  // namespace orc
