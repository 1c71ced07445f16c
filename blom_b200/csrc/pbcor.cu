// Baroclinic mass flux correction: pbcor1 (phy/mod_pbcor.F90:66-412) and pbcor2
// (:416-743), bmcmth 'uc' (upstream column) and 'dluc' (depth limited upstream
// column).  dpeps1=1e-5, dpeps2=1e-7 (:58-59).
//
// The reference loops k outside three masked 2-D sweeps (u faces, v faces, cell
// update) that communicate through the 2-D work arrays uflux,uflux2,uflux3,
// uflxtr of mod_utility.  Every level is in fact independent (the flux of level
// k depends only on the 2-D residual, the pre-update state of level k and the
// column total), so the GPU form is three launches over all levels at once:
//   pbcor_prep    per column: (pbcor2: dp=max(0,dp)+epsilp) p(k+1)=p(k)+dp on the
//                 1-wide ring; per face: residual utot=dlt*ub - sum_k uflx
//   pbcor_update  per (i,j,k): the four face fluxes of the cell are recomputed in
//                 registers (bit-identical on both sides of a face: products are
//                 taken with __dmul_rn so FMA contraction cannot split them),
//                 own faces accumulated into uflx/usflx/utflx, new dp/T/S/trc
//                 written to the ping-pong set shared with cppm (no in-place
//                 hazard with neighbours reading the pre-update state)
//   pbcor_finish  per column: p rebuilt, column rescaled to the barotropic bottom
//                 pressure, T/S/trc moved back
// The 2-D work arrays never exist; the residuals utotm/vtotm (pbcor1) and
// utotn/vtotn (pbcor2) are written like the reference does.
#include "common.cuh"
#include "eos.cuh"

namespace blom {

namespace {

constexpr int MAXTR = 4;
struct PbTr { double* t[MAXTR]; double* tb[MAXTR]; int n; };

template <int WHICH>
__global__ void pbcor_prep(Geom g, double dlt, int ks, int kf, int lt, const int* __restrict__ ip,
                           const int* __restrict__ iu, const int* __restrict__ iv, double* __restrict__ dp,
                           double* __restrict__ p, const double* __restrict__ ubt, const double* __restrict__ vbt,
                           const double* __restrict__ uflx, const double* __restrict__ vflx,
                           double* __restrict__ utot, double* __restrict__ vtot) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;  // 0..ii+1, 0..jj+1
  if (i > g.ii + 1) return;
  const long x = ix2(g, i, j), lev = g.lev;
  const int kk = g.kdm;
  if (ip[x] == 1) {
    double pk = p[x];
    for (int k = 1; k <= kk; ++k) {
      double d = dp[x + (long)(k + ks - 1) * lev];
      if (WHICH == 2) { d = fmax(0., d) + epsilp; dp[x + (long)(k + ks - 1) * lev] = d; }
      pk = pk + d;
      p[x + (long)k * lev] = pk;
    }
  }
  if (i >= 1 && j >= 1 && j <= g.jj && iu[x] == 1) {
    double t = dlt * ubt[x + (long)(lt - 1) * lev];
    for (int k = 1; k <= kk; ++k) t = t - uflx[x + (long)(k + kf - 1) * lev];
    utot[x] = t;
  }
  if (i >= 1 && i <= g.ii && j >= 1 && iv[x] == 1) {
    double t = dlt * vbt[x + (long)(lt - 1) * lev];
    for (int k = 1; k <= kk; ++k) t = t - vflx[x + (long)(k + kf - 1) * lev];
    vtot[x] = t;
  }
}

struct Flux3 { double f, f2, f3, ftr[MAXTR]; };

// Level-independent part of one face (mod_pbcor.F90:166-187 'uc', :240-265 'dluc'): residual,
// upstream cell and the column total / limited bottom pressure the flux is scaled with.
struct FaceInv { bool on; double tot, div; long up; };
template <bool DLUC>
__device__ __forceinline__ FaceInv face_prepare(long x, long s, int kk, long lev, const int* __restrict__ mask,
                                                const double* __restrict__ tot, const double* __restrict__ p) {
  FaceInv f; f.on = mask[x] == 1; f.tot = 0.; f.div = 1.; f.up = x;
  if (!f.on) return f;
  f.tot = tot[x];
  f.up = f.tot > 0. ? x - s : x;
  f.div = DLUC ? fmin(p[x + (long)kk * lev], p[x - s + (long)kk * lev]) : p[f.up + (long)kk * lev];
  return f;
}
// level-dependent operands of one (i,j): own cell, the upstream cell of each of the four faces and the
// old values of the six flux accumulators.  Loaded one level ahead (register double buffer) so that
// the loads of level k+1 are in flight while level k is computed.
struct LvlOps {
  double dpc, tc, sc;
  double dpf[4], tf[4], sf[4];   // upstream cell of faces u, v, east, north
  double pk0[4], pk1[4];         // dluc: p(k), p(k+1) of those cells
  double acc[6];                 // uflx, usflx, utflx, vflx, vsflx, vtflx
};

template <bool DLUC>
__device__ __forceinline__ Flux3 face_flux(const FaceInv& F, int f, const LvlOps& L, long ol,
                                           const PbTr& T) {
  Flux3 r{};
  if (!F.on) return r;
  if (!DLUC) r.f = __dmul_rn(F.tot, L.dpf[f]) / F.div;
  else r.f = __dmul_rn(F.tot, fmax(0., fmin(F.div, L.pk1[f]) - L.pk0[f])) / F.div;
  r.f2 = __dmul_rn(r.f, L.sf[f]);
  r.f3 = __dmul_rn(r.f, L.tf[f]);
#pragma unroll
  for (int nt = 0; nt < MAXTR; ++nt) if (nt < T.n) r.ftr[nt] = __dmul_rn(r.f, T.t[nt][F.up + ol]);
  return r;
}

// One thread per (i,j) and chunk of levels: the face invariants are prepared once, the level loop
// streams dp/T/S of the cell and its four upstream candidates.
template <int WHICH, bool DLUC, int MINB>
__global__ void __launch_bounds__(128, MINB)
pbcor_update(Geom g, eos::Coef ec, int ks, int kf, int kchunk, const int* __restrict__ ip, const int* __restrict__ iu,
             const int* __restrict__ iv, const double* __restrict__ utot, const double* __restrict__ vtot,
             const double* __restrict__ dp, const double* __restrict__ p, const double* __restrict__ temp,
             const double* __restrict__ saln, const double* __restrict__ scp2i, double* __restrict__ uflx,
             double* __restrict__ usflx, double* __restrict__ utflx, double* __restrict__ vflx,
             double* __restrict__ vsflx, double* __restrict__ vtflx, double* __restrict__ dpB,
             double* __restrict__ tB, double* __restrict__ sB, double* __restrict__ sigma, PbTr T,
             int extra /* 1: a further group of passive tracers; thickness, T, S and the flux arrays belong to the
                          first launch and are only read */) {
  const double dpeps1 = 1.e-5, dpeps2 = 1.e-7;
  const Bid b_ = bid(g);
  const int i = b_.x * blockDim.x + threadIdx.x + 1, j = b_.y + 1;  // 1..ii+1, 1..jj+1
  if (i > g.ii + 1) return;
  const long x = ix2(g, i, j), lev = g.lev, s = g.ldi;
  const int kk = g.kdm;
  const int k0 = b_.z * kchunk + 1, k1 = min(kk, k0 + kchunk - 1);
  const bool uok = j <= g.jj, vok = i <= g.ii;
  const bool cell = uok && vok && ip[x] == 1;
  FaceInv F[4];
#pragma unroll
  for (int f = 0; f < 4; ++f) { F[f].on = false; F[f].tot = 0.; F[f].div = 1.; F[f].up = x; }
  if (uok) F[0] = face_prepare<DLUC>(x, 1, kk, lev, iu, utot, p);
  if (vok) F[1] = face_prepare<DLUC>(x, s, kk, lev, iv, vtot, p);
  if (cell) {
    F[2] = face_prepare<DLUC>(x + 1, 1, kk, lev, iu, utot, p);
    F[3] = face_prepare<DLUC>(x + s, s, kk, lev, iv, vtot, p);
  }
  if (!F[0].on && !F[1].on && !cell) return;
  const double a = cell ? scp2i[x] : 0.;

  auto load = [&](int k, LvlOps& L) {
    const long ol = (long)(k + ks - 1) * lev, oa = (long)(k + kf - 1) * lev;
    L.dpc = dp[x + ol]; L.tc = temp[x + ol]; L.sc = saln[x + ol];
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      L.dpf[f] = dp[F[f].up + ol]; L.tf[f] = temp[F[f].up + ol]; L.sf[f] = saln[F[f].up + ol];
      if (DLUC) { L.pk0[f] = p[F[f].up + (long)(k - 1) * lev]; L.pk1[f] = p[F[f].up + (long)k * lev]; }
    }
    L.acc[0] = uflx[x + oa]; L.acc[1] = usflx[x + oa]; L.acc[2] = utflx[x + oa];
    L.acc[3] = vflx[x + oa]; L.acc[4] = vsflx[x + oa]; L.acc[5] = vtflx[x + oa];
  };
  LvlOps cur, nxt;
  load(k0, cur);
  for (int k = k0; k <= k1; ++k) {
    if (k < k1) load(k + 1, nxt);
    const long ol = (long)(k + ks - 1) * lev, oa = (long)(k + kf - 1) * lev, ok = (long)(k - 1) * lev;
    const Flux3 fu = face_flux<DLUC>(F[0], 0, cur, ol, T);
    const Flux3 fv = face_flux<DLUC>(F[1], 1, cur, ol, T);
    if (F[0].on && !extra) {
      uflx[x + oa] = cur.acc[0] + fu.f;
      usflx[x + oa] = cur.acc[1] + fu.f2;
      utflx[x + oa] = cur.acc[2] + fu.f3;
    }
    if (F[1].on && !extra) {
      vflx[x + oa] = cur.acc[3] + fv.f;
      vsflx[x + oa] = cur.acc[4] + fv.f2;
      vtflx[x + oa] = cur.acc[5] + fv.f3;
    }
    if (cell) {
      const Flux3 fue = face_flux<DLUC>(F[2], 2, cur, ol, T);
      const Flux3 fvn = face_flux<DLUC>(F[3], 3, cur, ol, T);
      double dpo = cur.dpc, dpn, dpni;
      const double dm = fue.f - fu.f + fvn.f - fv.f;
      const double ds = fue.f2 - fu.f2 + fvn.f2 - fv.f2;
      const double dt = fue.f3 - fu.f3 + fvn.f3 - fv.f3;
      if (WHICH == 1) {
        dpn = fmax(0., dpo - dm * a);
        dpo = dpo + dpeps1;
        dpni = 1. / (dpn + dpeps1);
      } else {
        dpn = dpo - a * dm;
        dpni = 1. / dpn;
      }
      const double sn = (dpo * cur.sc - ds * a) * dpni;
      const double tn = (dpo * cur.tc - dt * a) * dpni;
#pragma unroll
      for (int nt = 0; nt < MAXTR; ++nt) if (nt < T.n) {
        const double dtr = fue.ftr[nt] - fu.ftr[nt] + fvn.ftr[nt] - fv.ftr[nt];
        T.tb[nt][x + ok] = (dpo * T.t[nt][x + ol] - dtr * a) * dpni;
      }
      if (!extra) {
        if (WHICH == 2) {
          sigma[x + ol] = eos::sig(ec, tn, sn);
          dpn = dpn - epsilp;
        }
        if (dpn < dpeps2) dpn = 0.;
        dpB[x + ok] = dpn; tB[x + ok] = tn; sB[x + ok] = sn;
      }
    }
    cur = nxt;
  }
}

// copy-back of a further tracer group (the first group's copy-back is part of pbcor_finish)
__global__ void pbcor_finish_tracers(Geom g, int ks, const int* __restrict__ ip, PbTr T) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1, k = blockIdx.z + 1;
  if (i > g.ii) return;
  const long x = ix2(g, i, j), lev = g.lev;
  if (ip[x] != 1) return;
  const long ok = (long)(k - 1) * lev, ol = (long)(k + ks - 1) * lev;
#pragma unroll
  for (int nt = 0; nt < MAXTR; ++nt)
    if (nt < T.n) T.t[nt][x + ol] = T.tb[nt][x + ok];
}

template <int WHICH>
__global__ void pbcor_finish(Geom g, int ks, const int* __restrict__ ip, const double* __restrict__ pbt,
                             const double* __restrict__ dpB, const double* __restrict__ tB,
                             const double* __restrict__ sB, double* __restrict__ dp, double* __restrict__ temp,
                             double* __restrict__ saln, double* __restrict__ p, PbTr T) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.ii) return;
  const long x = ix2(g, i, j), lev = g.lev;
  if (ip[x] != 1) return;
  const int kk = g.kdm;
  const double p1 = p[x];
  double pk = p1;
  for (int k = 1; k <= kk; ++k) {
    pk = pk + dpB[x + (long)(k - 1) * lev];
    if (WHICH == 1) p[x + (long)k * lev] = pk;
  }
  const double pbfac = pbt[x] / pk;
  pk = p1;
  for (int k = 1; k <= kk; ++k) {
    const long ok = (long)(k - 1) * lev, ol = (long)(k + ks - 1) * lev;
    const double d = dpB[x + ok] * pbfac;
    dp[x + ol] = d;
    if (WHICH == 2) { pk = pk + d; p[x + (long)k * lev] = pk; }
    temp[x + ol] = tB[x + ok];
    saln[x + ol] = sB[x + ok];
#pragma unroll
    for (int nt = 0; nt < MAXTR; ++nt)   // static index: the pointer table stays in the constant bank
      if (nt < T.n) T.t[nt][x + ol] = T.tb[nt][x + ok];
  }
}

template <int WHICH>
void pbcor_run(int m, int n, int mm, int nn, int k1m) {
  Ctx& c = C(); const Geom& g = c.g;
  const std::string bmcmth = c.option("bmcmth", "uc");
  if (bmcmth != "uc" && bmcmth != "dluc")
    throw std::runtime_error(" bmcmth = " + bmcmth + " is unsupported! " + (WHICH == 1 ? "(pbcor1)" : "(pbcor2)"));
  const bool dluc = bmcmth == "dluc";
  const double dlt = c.scalar("dlt");
  const int ks = WHICH == 1 ? nn : mm, kf = WHICH == 1 ? mm : nn, lt = WHICH == 1 ? m : n;
  const int kk = g.kdm;
  // passive tracers in groups of MAXTR: the first group is corrected together with dp, T and S, further groups by
  // extra launches of the update kernel that only touch their tracers (the reference loops nt = 1..ntr per cell,
  // phy/mod_pbcor.F90:330-336; per tracer the operations are the same)
  std::vector<PbTr> groups;
  for (int n0 = 0; n0 < std::max(1, g.ntr); n0 += MAXTR) {
    PbTr G{}; G.n = std::max(0, std::min(MAXTR, g.ntr - n0));
    for (int q = 0; q < G.n; ++q) {
      G.t[q] = c.dev("trc") + (long)(n0 + q) * 2 * kk * g.lev;
      G.tb[q] = c.owned("cppm_tmp_trc" + std::to_string(n0 + q + 1), kk);
    }
    groups.push_back(G);
  }
  const PbTr T = groups[0];
  if (WHICH == 2) {  // :433-440
    halo_update(std::vector<HaloReq>{{c.dev("ubflxs") + (long)(n - 1) * g.lev, 1, halo_uv},
                                     {c.dev("vbflxs") + (long)(n - 1) * g.lev, 1, halo_vv}}, 1, 1);
    if (g.ntr > 0) {
      std::vector<HaloReq> r;
      for (int nt = 0; nt < g.ntr; ++nt)
        r.push_back({c.dev("trc") + ((long)nt * 2 * kk + k1m - 1) * g.lev, kk, halo_ps});
      halo_update(r, 1, 1);
    }
  }
  const char* un = WHICH == 1 ? "utotm" : "utotn"; const char* vn = WHICH == 1 ? "vtotm" : "vtotn";
  double* utot = c.has(un) ? c.dev(un) : c.owned(un, 1);
  double* vtot = c.has(vn) ? c.dev(vn) : c.owned(vn, 1);
  double *dp = c.dev("dp"), *p = c.dev("p"), *temp = c.dev("temp"), *saln = c.dev("saln");
  double *dpB = c.owned("cppm_tmp_dp", kk), *tB = c.owned("cppm_tmp_temp", kk), *sB = c.owned("cppm_tmp_saln", kk);
  const int *ip = c.idev("ip"), *iu = c.idev("iu"), *iv = c.idev("iv");
  {
    dim3 grid(cdiv(g.ii + 2, 128), g.jj + 2);
    LAUNCH_NAMED(WHICH == 1 ? "pbcor_prep<1>" : "pbcor_prep<2>", pbcor_prep<WHICH>, grid, 128, 0, g, dlt, ks, kf, lt,
                 ip, iu, iv, dp, p, c.dev(WHICH == 1 ? "ubflxs_p" : "ubflxs"), c.dev(WHICH == 1 ? "vbflxs_p" : "vbflxs"),
                 c.dev("uflx"), c.dev("vflx"), utot, vtot);
  }
  {
    // levels marched per block (development switch pbcor_kchunk; 0 = default)
    const int kc_opt = std::stoi(c.option("pbcor_kchunk", "0"));
    const int kchunk = kc_opt > 0 ? std::min(kc_opt, kk) : (kk >= 16 ? cdiv(kk, 2) : kk);   // 27 beat 8 and 14 at tnx0.25v4
    const dim3 grid = lgrid(g, dim3(cdiv(g.ii + 1, 128), g.jj + 1, cdiv(kk, kchunk)));
    const char* nm = WHICH == 1 ? "pbcor_update<1>" : "pbcor_update<2>";
    const eos::Coef ec = WHICH == 2 ? eos::host_coef() : eos::Coef{};  // only pbcor2 refreshes sigma
#define PB_LAUNCH(D)                                                                                            \
    LAUNCH_NAMED(nm, (pbcor_update<WHICH, D, OCC>), grid, 128, 0, g, ec, ks, kf, kchunk, ip, iu, iv, utot, vtot, dp, \
                 p, temp, saln, c.dev("scp2i"), c.dev("uflx"), c.dev("usflx"), c.dev("utflx"), c.dev("vflx"),     \
                 c.dev("vsflx"), c.dev("vtflx"), dpB, tB, sB, c.dev("sigma"), TG, EXTRA)
    for (size_t gi = 0; gi < groups.size(); ++gi) {
      const PbTr TG = groups[gi];
      const int EXTRA = gi > 0 ? 1 : 0;
      // resident blocks: the depth-limited branch (dluc, the hybrid default) carries the four faces' interface
      // pressures as well and wants the registers of 3 blocks (tnx0.25v4, pbcor1/2: 3 blocks 3.76/3.89 ms, 4 blocks
      // 4.41/4.83, 5 blocks 6.12/6.67); uc runs best with 4 (round 1)
      if (dluc) { OCC_DISPATCH3("pbcor_minblk", 3, 2, 3, 4, PB_LAUNCH(true)); }
      else { OCC_DISPATCH3("pbcor_minblk", 4, 3, 4, 5, PB_LAUNCH(false)); }
      if (EXTRA) {
        dim3 gridf(cdiv(g.ii, 128), g.jj, kk);
        LAUNCH(pbcor_finish_tracers, gridf, 128, 0, g, ks, ip, TG);
      }
    }
#undef PB_LAUNCH
  }
  {
    dim3 grid(cdiv(g.ii, 128), g.jj);
    const double* pbt = WHICH == 1 ? c.dev("pb_p") : c.dev("pb") + (long)(m - 1) * g.lev;
    LAUNCH_NAMED(WHICH == 1 ? "pbcor_finish<1>" : "pbcor_finish<2>", pbcor_finish<WHICH>, grid, 128, 0, g, ks, ip, pbt,
                 dpB, tB, sB, dp, temp, saln, p, T);
  }
}

}  // namespace

void pbcor1_dev(int m, int n, int mm, int nn, int k1m, int k1n) { (void)k1n; pbcor_run<1>(m, n, mm, nn, k1m); }
void pbcor2_dev(int m, int n, int mm, int nn, int k1m, int k1n) { (void)k1n; pbcor_run<2>(m, n, mm, nn, k1m); }

}  // namespace blom
