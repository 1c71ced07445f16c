// Split-explicit barotropic subcycle (phy/mod_barotp.F90:148-1003).
//
// Five blocks of lstep/2 forward-backward substeps; each substep is continuity
// followed by the two momentum equations, u before v on odd substeps and v
// before u on even ones (the second equation uses the first one's new flux).
// Halos of the three subcycled fields are refreshed once per two substeps with
// a 2(3)-wide halo exactly like the reference (:395-397), so the odd substep
// runs on the widened ranges and the even one on the interior.
//
// Kernel structure: one persistent cooperative launch per block of substeps
// (bt_subcycle: phases separated by grid barriers, in-kernel E/W + fold halos and,
// on several GPUs, in-kernel peer-to-peer band edges), i across lanes.  Two
// phases per substep: the continuity equation is evaluated inside the first
// momentum phase.  Arrays whose time weight is zero in a block are not read.
// The launch-per-phase form (bt_continuity / bt_ueq / bt_veq, option
// barotp_kernel=phases) is kept as the plain statement of the same operations.
// All pointers travel in one by-value parameter block.
#include "common.cuh"
#include "halo.cuh"
#include "p2p.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace blom {

namespace {

struct BtP {
  double *pb_ml, *pb_nl, *ub_ml, *ub_nl, *vb_ml, *vb_nl;
  const double *scp2i, *scvxi, *scuyi, *scuxi, *scvyi, *scuy, *scvx;
  const double *pvo, *pvm, *pvn;
  const double *pgfxm_o, *xixp_o, *xixm_o, *pgfxm_m, *xixp_m, *xixm_m, *pgfxm_n, *xixp_n, *xixm_n;
  const double *pgfym_o, *xiyp_o, *xiym_o, *pgfym_m, *xiyp_m, *xiym_m, *pgfym_n, *xiyp_n, *xiym_n;
  const double *utotn, *vtotn, *uglue, *vglue, *umaxb, *uminb, *vmaxb, *vminb;
  double *ubflxs_t, *ubcors_t, *vbflxs_t, *vbcors_t;
  const int *ip, *iu, *iv;
  double wo, wm, wn, dlt;
  int enscon;
};

constexpr double WBARO = .125;  // phy/mod_tmsmt.F90:51

// :177-224  clamp fluxes and coastal damping coefficients
__global__ void bt_prologue(Geom g, int m, int nn, double cwbdts, double cwbdls, const int* __restrict__ iu,
                            const int* __restrict__ iv, const double* __restrict__ u,
                            const double* __restrict__ v, const double* __restrict__ pbu,
                            const double* __restrict__ pbv, const double* __restrict__ umax,
                            const double* __restrict__ vmax, const double* __restrict__ scuy,
                            const double* __restrict__ scvx, double* __restrict__ umaxb,
                            double* __restrict__ uminb, double* __restrict__ uglue, double* __restrict__ vmaxb,
                            double* __restrict__ vminb, double* __restrict__ vglue) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.ii) return;
  const long x = ix2(g, i, j), x2 = x + (long)(m - 1) * g.lev;
  if (iu[x] == 1) {
    double mx = 0., mn = 0.;
    for (int k = 1; k <= g.kdm; ++k) {
      const double uk = u[x + (long)(k + nn - 1) * g.lev];
      mx = fmax(mx, uk); mn = fmin(mn, uk);
    }
    uglue[x] = cwbdts * exp(1. - pbu[x2] / (cwbdls * onem));
    umaxb[x] = (umax[x] - mx) * pbu[x2] * scuy[x];
    uminb[x] = (umax[x] + mn) * pbu[x2] * scuy[x];
  }
  if (iv[x] == 1) {
    double mx = 0., mn = 0.;
    for (int k = 1; k <= g.kdm; ++k) {
      const double vk = v[x + (long)(k + nn - 1) * g.lev];
      mx = fmax(mx, vk); mn = fmin(mn, vk);
    }
    vglue[x] = cwbdts * exp(1. - pbv[x2] / (cwbdls * onem));
    vmaxb[x] = (vmax[x] - mx) * pbv[x2] * scvx[x];
    vminb[x] = (vmax[x] + mn) * pbv[x2] * scvx[x];
  }
}

// :230-269  pvtrop_o <- pvtrop(n); new pvtrop(n).  Gather form of the reference's
// three scatter loops: the last writer in the reference's sequential order wins.
__global__ void bt_pvtrop(Geom g, const int* __restrict__ iu, const int* __restrict__ iv,
                          const int* __restrict__ iq, const double* __restrict__ corioq,
                          const double* __restrict__ pb_p, double* __restrict__ pvn, double* __restrict__ pvo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // 0..ii+1
  const int j = (int)blockIdx.y - 2;                   // -2..jj+3
  if (i > g.ii + 1) return;
  const long x = ix2(g, i, j), s = g.ldi;
  double val = pvn[x];
  pvo[x] = val;
  const double cq = corioq[x];
  const bool iin = i >= 1 && i <= g.ii;
  if (iin && j - 1 >= 0 && j - 1 <= g.jj && iu[x - s] == 1) val = cq * (2. / (pb_p[x - s] + pb_p[x - s - 1]));
  if (iin && j >= 0 && j <= g.jj && iu[x] == 1) val = cq * (2. / (pb_p[x] + pb_p[x - 1]));
  if (j >= 1 && j <= g.jj) {
    if (i - 1 >= 0 && i - 1 <= g.ii && iv[x - 1] == 1) val = cq * (2. / (pb_p[x - 1] + pb_p[x - 1 - s]));
    if (i <= g.ii && iv[x] == 1) val = cq * (2. / (pb_p[x] + pb_p[x - s]));
    if (iin && iq[x] == 1) val = cq * 4. / (pb_p[x] + pb_p[x - 1] + pb_p[x - s] + pb_p[x - s - 1]);
  }
  pvn[x] = val;
}

// :290-319 arctic switches in the halo region next to the fold
__global__ void bt_arctic_swap(Geom g, double* umaxb, double* uminb, double* xixp, double* xixm, double* vmaxb,
                               double* vminb, double* xiyp, double* xiym) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // 0..ii+1
  const int j = g.jj + blockIdx.y;                     // jj..jj+2
  if (i > g.ii + 1) return;
  const long x = ix2(g, i, j);
  double q = umaxb[x]; umaxb[x] = uminb[x]; uminb[x] = q;
  q = xixp[x]; xixp[x] = xixm[x]; xixm[x] = q;
  if (j > g.jj || i >= max(0, g.itdm / 2 - g.i0 + 1)) {
    q = vmaxb[x]; vmaxb[x] = vminb[x]; vminb[x] = q;
    q = xiyp[x]; xiyp[x] = xiym[x]; xiym[x] = q;
  }
}

__global__ void bt_init_block1(Geom g, const double* __restrict__ pb_mn, const double* __restrict__ ub_mn,
                               const double* __restrict__ vb_mn, double* pb_t, double* ub_t, double* vb_t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.ii) return;
  const long x = ix2(g, i, j);
  pb_t[x] = pb_mn[x]; pb_t[x + g.lev] = pb_mn[x + g.lev];
  ub_t[x] = ub_mn[x]; ub_t[x + g.lev] = ub_mn[x + g.lev];
  vb_t[x] = vb_mn[x]; vb_t[x + g.lev] = vb_mn[x + g.lev];
}

__global__ void bt_zero_acc(Geom g, BtP P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // 0..ii+1
  const int j = (int)blockIdx.y - 1;                   // -1..jj+2
  if (i > g.ii + 1) return;
  const long x = ix2(g, i, j);
  if (P.iu[x] == 1) { P.ubflxs_t[x] = 0.; P.ubcors_t[x] = 0.; }
  if (j >= 0 && i <= g.ii && P.iv[x] == 1) { P.vbflxs_t[x] = 0.; P.vbcors_t[x] = 0.; }
}

__global__ void bt_continuity(Geom g, BtP P, int i0, int i1, int j0, int j1) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + i0, j = blockIdx.y + j0;
  if (i > i1 || j > j1) return;
  const long x = ix2(g, i, j);
  if (P.ip[x] != 1) return;
  P.pb_nl[x] = (1. - WBARO) * P.pb_ml[x] + WBARO * P.pb_nl[x] -
               (1. + WBARO) * P.dlt * (P.ub_ml[x + 1] - P.ub_ml[x] + P.vb_ml[x + g.ldi] - P.vb_ml[x]) * P.scp2i[x];
}

// vb: level of vbflx_t entering the Coriolis term (ml on odd, nl on even substeps)
__global__ void bt_ueq(Geom g, BtP P, const double* __restrict__ vb, int i0, int i1, int j0, int j1) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + i0, j = blockIdx.y + j0;
  if (i > i1 || j > j1) return;
  const long x = ix2(g, i, j), s = g.ldi;
  if (P.iu[x] != 1) return;
  const double uml = P.ub_ml[x], unl = P.ub_nl[x];
  P.ubflxs_t[x] = P.ubflxs_t[x] - WBARO * unl + (1. + WBARO) * uml;
  double q;
  if (P.enscon)
    q = (vb[x] * P.scvxi[x] + vb[x + s] * P.scvxi[x + s] + vb[x - 1] * P.scvxi[x - 1] +
         vb[x - 1 + s] * P.scvxi[x - 1 + s]) *
        (P.wo * (P.pvo[x] + P.pvo[x + s]) + P.wm * (P.pvm[x] + P.pvm[x + s]) + P.wn * (P.pvn[x] + P.pvn[x + s])) * .125;
  else
    q = .25 * ((vb[x] * P.scvxi[x] + vb[x - 1] * P.scvxi[x - 1]) * (P.wo * P.pvo[x] + P.wm * P.pvm[x] + P.wn * P.pvn[x]) +
               (vb[x + s] * P.scvxi[x + s] + vb[x - 1 + s] * P.scvxi[x - 1 + s]) *
                   (P.wo * P.pvo[x + s] + P.wm * P.pvm[x + s] + P.wn * P.pvn[x + s]));
  P.ubcors_t[x] = P.ubcors_t[x] + q;
  const double pbc = P.pb_nl[x], pbw = P.pb_nl[x - 1];
  const double utndcy = q + (P.wo * (P.pgfxm_o[x] - (P.xixp_o[x] * pbc - P.xixm_o[x] * pbw)) +
                             P.wm * (P.pgfxm_m[x] - (P.xixp_m[x] * pbc - P.xixm_m[x] * pbw)) +
                             P.wn * (P.pgfxm_n[x] - (P.xixp_n[x] * pbc - P.xixm_n[x] * pbw))) * P.scuxi[x];
  double un = (1. - WBARO) * uml + WBARO * unl +
              (1. + WBARO) * P.dlt * ((utndcy + P.utotn[x]) * P.scuy[x] * fmin(pbw, pbc) - P.uglue[x] * uml);
  P.ub_nl[x] = fmax(-P.uminb[x], fmin(P.umaxb[x], un));
}

__global__ void bt_veq(Geom g, BtP P, const double* __restrict__ ub, int i0, int i1, int j0, int j1) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + i0, j = blockIdx.y + j0;
  if (i > i1 || j > j1) return;
  const long x = ix2(g, i, j), s = g.ldi;
  if (P.iv[x] != 1) return;
  const double vml = P.vb_ml[x], vnl = P.vb_nl[x];
  P.vbflxs_t[x] = P.vbflxs_t[x] - WBARO * vnl + (1. + WBARO) * vml;
  double q;
  if (P.enscon)
    q = -(ub[x] * P.scuyi[x] + ub[x + 1] * P.scuyi[x + 1] + ub[x - s] * P.scuyi[x - s] +
          ub[x + 1 - s] * P.scuyi[x + 1 - s]) *
        (P.wo * (P.pvo[x] + P.pvo[x + 1]) + P.wm * (P.pvm[x] + P.pvm[x + 1]) + P.wn * (P.pvn[x] + P.pvn[x + 1])) * .125;
  else
    q = -.25 * ((ub[x] * P.scuyi[x] + ub[x - s] * P.scuyi[x - s]) * (P.wo * P.pvo[x] + P.wm * P.pvm[x] + P.wn * P.pvn[x]) +
                (ub[x + 1] * P.scuyi[x + 1] + ub[x + 1 - s] * P.scuyi[x + 1 - s]) *
                    (P.wo * P.pvo[x + 1] + P.wm * P.pvm[x + 1] + P.wn * P.pvn[x + 1]));
  P.vbcors_t[x] = P.vbcors_t[x] + q;
  const double pbc = P.pb_nl[x], pbs = P.pb_nl[x - s];
  const double vtndcy = q + (P.wo * (P.pgfym_o[x] - (P.xiyp_o[x] * pbc - P.xiym_o[x] * pbs)) +
                             P.wm * (P.pgfym_m[x] - (P.xiyp_m[x] * pbc - P.xiym_m[x] * pbs)) +
                             P.wn * (P.pgfym_n[x] - (P.xiyp_n[x] * pbc - P.xiym_n[x] * pbs))) * P.scvyi[x];
  double vn = (1. - WBARO) * vml + WBARO * vnl +
              (1. + WBARO) * P.dlt * ((vtndcy + P.vtotn[x]) * P.scvx[x] * fmin(pbs, pbc) - P.vglue[x] * vml);
  P.vb_nl[x] = fmax(-P.vminb[x], fmin(P.vmaxb[x], vn));
}

// ---- persistent form of the subcycle ------------------------------------------------------------
// One cooperative launch runs a whole block of lstep/2 substeps: the phases of a substep (continuity
// + first momentum equation, second momentum equation; three phases with barotp_fuse=0) and, on one
// tile, the halo refresh of the subcycled fields are separated by grid-wide barriers instead of kernel
// boundaries.  Read-write
// fields are read with ld.global.cg (L1 is not coherent between the SMs of one launch); the ~40
// read-only coefficient arrays use the normal cached path.  Operation order per cell is the one of
// bt_continuity / bt_ueq / bt_veq above, so results are bit-identical to the launch-per-phase form.
#define BT_SHAPE_DEFAULT "768x2"   // tnx0.25v4, round 1: 512x2 22.7 ms, 640x2 22.2, 768x2 21.5, 1024x2 21.6, 256x2 29.6;
                                   // round 2 (zero-weight skip + two phases): 512x2 18.5, 768x2 18.6, 1024x2 27.7

struct BtSched {
  int lll0, nsub, ml, nl;
  // levels (1-based) of the three bottom-pressure buffers that hold the mid level, the old/new level and the
  // spare the fused form writes the new level into; fuse = 1: continuity is evaluated inside the first momentum
  // phase of a substep (see bt_subcycle)
  int pml, pnl, psp, fuse;
  double woa, wob, wna, wnb;
  int inkernel_halo;
  // band edges exchanged in-kernel through the peer mailboxes (multi-GPU); seq0 = sequence number of
  // the first exchange of this launch
  int p2p;
  unsigned long long seq0;
};

// time-level pointers and time weights of the current substep
struct BtLv { double *pb_ml, *pb_nl, *ub_ml, *ub_nl, *vb_ml, *vb_nl; double wo, wm, wn; };

// Time-weight pattern of a block of substeps (phy/mod_barotp.F90:330-358): the weights wo/wm/wn of the
// old/mid/new baroclinic forcing are fixed per block, and in every block at least one of them is exactly
// zero for all its substeps: block 1 has wn = 0, blocks 2-3 have wo = 0, blocks 4-5 have wo = wm = 0, wn = 1.
// A term `0 * a` contributes +-0 to its sum, so the arrays that only enter through a zero weight are not
// read at all (7 of the 53 words per point and substep in blocks 1-3, 14 in blocks 4-5); the result is the
// same number (at most the sign of an exact zero differs).
enum { W_ALL = 0, W_NO_N = 1, W_NO_O = 2, W_ONLY_N = 3 };
template <int WM> struct Wsel {
  static constexpr bool o = (WM == W_ALL || WM == W_NO_N), m = (WM != W_ONLY_N), n = (WM != W_NO_N);
};
// wo*a_o + wm*a_m + wn*a_n in the reference's order, without the terms whose weight is zero
template <int WM>
__device__ __forceinline__ double wsum3(const BtLv& V, double a_o, double a_m, double a_n) {
  if (WM == W_ALL) return V.wo * a_o + V.wm * a_m + V.wn * a_n;
  if (WM == W_NO_N) return V.wo * a_o + V.wm * a_m;
  if (WM == W_NO_O) return V.wm * a_m + V.wn * a_n;
  return V.wn * a_n;
}

// (the callers test the mask of the cell: P.ip for btp_continuity, P.iu for btp_ueq, P.iv for btp_veq)
__device__ __forceinline__ void btp_continuity(const Geom& g, const BtP& P, const BtLv& V, long x) {
  V.pb_nl[x] = (1. - WBARO) * __ldcg(V.pb_ml + x) + WBARO * __ldcg(V.pb_nl + x) -
               (1. + WBARO) * P.dlt * (__ldcg(V.ub_ml + x + 1) - __ldcg(V.ub_ml + x) + __ldcg(V.vb_ml + x + g.ldi) -
                                       __ldcg(V.vb_ml + x)) * P.scp2i[x];
}
__device__ __forceinline__ double btp_continuity_value(const Geom& g, const BtP& P, const BtLv& V, long x) {
  return (1. - WBARO) * __ldcg(V.pb_ml + x) + WBARO * __ldcg(V.pb_nl + x) -
         (1. + WBARO) * P.dlt * (__ldcg(V.ub_ml + x + 1) - __ldcg(V.ub_ml + x) + __ldcg(V.vb_ml + x + g.ldi) -
                                 __ldcg(V.vb_ml + x)) * P.scp2i[x];
}
// pbc, pbw: new bottom pressure at the u point's two mass points (read from the array by the staged form,
// recomputed from the continuity equation by the fused form - the same expression on the same operands)
template <int WM>
__device__ __forceinline__ void btp_ueq(const Geom& g, const BtP& P, const BtLv& V, const double* __restrict__ vb, long x,
                                        double pbc, double pbw) {
  using W = Wsel<WM>;
  const long s = g.ldi;
  const double uml = __ldcg(V.ub_ml + x), unl = __ldcg(V.ub_nl + x);
  P.ubflxs_t[x] = __ldcg(P.ubflxs_t + x) - WBARO * unl + (1. + WBARO) * uml;
  const double v00 = __ldcg(vb + x), v01 = __ldcg(vb + x + s), vm0 = __ldcg(vb + x - 1), vm1 = __ldcg(vb + x - 1 + s);
  const double pvo0 = W::o ? P.pvo[x] : 0., pvo1 = W::o ? P.pvo[x + s] : 0.;
  const double pvm0 = W::m ? P.pvm[x] : 0., pvm1 = W::m ? P.pvm[x + s] : 0.;
  const double pvn0 = W::n ? P.pvn[x] : 0., pvn1 = W::n ? P.pvn[x + s] : 0.;
  double q;
  if (P.enscon)
    q = (v00 * P.scvxi[x] + v01 * P.scvxi[x + s] + vm0 * P.scvxi[x - 1] + vm1 * P.scvxi[x - 1 + s]) *
        wsum3<WM>(V, pvo0 + pvo1, pvm0 + pvm1, pvn0 + pvn1) * .125;
  else
    q = .25 * ((v00 * P.scvxi[x] + vm0 * P.scvxi[x - 1]) * wsum3<WM>(V, pvo0, pvm0, pvn0) +
               (v01 * P.scvxi[x + s] + vm1 * P.scvxi[x - 1 + s]) * wsum3<WM>(V, pvo1, pvm1, pvn1));
  P.ubcors_t[x] = __ldcg(P.ubcors_t + x) + q;
  const double t_o = W::o ? P.pgfxm_o[x] - (P.xixp_o[x] * pbc - P.xixm_o[x] * pbw) : 0.;
  const double t_m = W::m ? P.pgfxm_m[x] - (P.xixp_m[x] * pbc - P.xixm_m[x] * pbw) : 0.;
  const double t_n = W::n ? P.pgfxm_n[x] - (P.xixp_n[x] * pbc - P.xixm_n[x] * pbw) : 0.;
  const double utndcy = q + wsum3<WM>(V, t_o, t_m, t_n) * P.scuxi[x];
  const double un = (1. - WBARO) * uml + WBARO * unl +
                    (1. + WBARO) * P.dlt * ((utndcy + P.utotn[x]) * P.scuy[x] * fmin(pbw, pbc) - P.uglue[x] * uml);
  V.ub_nl[x] = fmax(-P.uminb[x], fmin(P.umaxb[x], un));
}
template <int WM>
__device__ __forceinline__ void btp_veq(const Geom& g, const BtP& P, const BtLv& V, const double* __restrict__ ub, long x,
                                        double pbc, double pbs) {
  using W = Wsel<WM>;
  const long s = g.ldi;
  const double vml = __ldcg(V.vb_ml + x), vnl = __ldcg(V.vb_nl + x);
  P.vbflxs_t[x] = __ldcg(P.vbflxs_t + x) - WBARO * vnl + (1. + WBARO) * vml;
  const double u00 = __ldcg(ub + x), u10 = __ldcg(ub + x + 1), u0m = __ldcg(ub + x - s), u1m = __ldcg(ub + x + 1 - s);
  const double pvo0 = W::o ? P.pvo[x] : 0., pvo1 = W::o ? P.pvo[x + 1] : 0.;
  const double pvm0 = W::m ? P.pvm[x] : 0., pvm1 = W::m ? P.pvm[x + 1] : 0.;
  const double pvn0 = W::n ? P.pvn[x] : 0., pvn1 = W::n ? P.pvn[x + 1] : 0.;
  double q;
  if (P.enscon)
    q = -(u00 * P.scuyi[x] + u10 * P.scuyi[x + 1] + u0m * P.scuyi[x - s] + u1m * P.scuyi[x + 1 - s]) *
        wsum3<WM>(V, pvo0 + pvo1, pvm0 + pvm1, pvn0 + pvn1) * .125;
  else
    q = -.25 * ((u00 * P.scuyi[x] + u0m * P.scuyi[x - s]) * wsum3<WM>(V, pvo0, pvm0, pvn0) +
                (u10 * P.scuyi[x + 1] + u1m * P.scuyi[x + 1 - s]) * wsum3<WM>(V, pvo1, pvm1, pvn1));
  P.vbcors_t[x] = __ldcg(P.vbcors_t + x) + q;
  const double t_o = W::o ? P.pgfym_o[x] - (P.xiyp_o[x] * pbc - P.xiym_o[x] * pbs) : 0.;
  const double t_m = W::m ? P.pgfym_m[x] - (P.xiyp_m[x] * pbc - P.xiym_m[x] * pbs) : 0.;
  const double t_n = W::n ? P.pgfym_n[x] - (P.xiyp_n[x] * pbc - P.xiym_n[x] * pbs) : 0.;
  const double vtndcy = q + wsum3<WM>(V, t_o, t_m, t_n) * P.scvyi[x];
  const double vn = (1. - WBARO) * vml + WBARO * vnl +
                    (1. + WBARO) * P.dlt * ((vtndcy + P.vtotn[x]) * P.scvx[x] * fmin(pbs, pbc) - P.vglue[x] * vml);
  V.vb_nl[x] = fmax(-P.vminb[x], fmin(P.vmaxb[x], vn));
}

// Grid-wide barrier on a monotonically increasing counter (one atomic per block and barrier); cheaper
// than the generic cooperative-groups barrier for the ~440 barriers of one call.  The cooperative
// launch guarantees co-residency.  Writes of the block are made visible by bar.sync + the cumulative
// gpu-scope fence of thread 0; read-write data is read with ld.global.cg afterwards.
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(ctr, 1u);
    while (*(volatile unsigned*)ctr < target) {}
    __threadfence();
  }
  __syncthreads();
}

template <int THREADS, int MINBLK, int WM, bool FUSE>
__global__ void __launch_bounds__(THREADS, MINBLK)
bt_subcycle(Geom g, const BtP P, BtSched S, P2PView X, double* pb_t, double* ub_t, double* vb_t, unsigned* ctr) {
  // FUSED FORM (template FUSE, option barotp_fuse): a substep is two phases instead of three.  The continuity equation is evaluated inside
  // the first momentum phase: the thread of a cell computes the cell's new bottom pressure, stores it, and - if
  // the cell's u (odd substeps) or v (even substeps) point is wet - recomputes the new pressure of the western
  // (southern) neighbour from the same old fields to form the pressure gradient.  Same expression, same
  // operands, so the values are those of the three-phase form; one grid barrier and one sweep over the plane
  // per substep are gone.  The new pressure cannot overwrite the old/new level in place (the neighbour still
  // needs the old value), so pb_t has a third level and the three rotate: (mid, old, spare) -> (spare, mid, old).
  unsigned target = 0;
  unsigned long long seq = S.seq0;
  const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long)gridDim.x * blockDim.x;
  const long L = g.lev;
  int ml = S.ml, nl = S.nl, pml = S.pml, pnl = S.pnl, psp = S.psp;
  BtLv V;
  // cells idx = tid, tid+nthr, ... of the range; body runs where mask == 1.  (i,j) of the thread's cells is
  // advanced incrementally: one 32-bit division per phase instead of a 64-bit division and modulo in front
  // of every cell's load chain.  (An L2 prefetch of the next cell's operands and a mask look-ahead were
  // measured 5 % slower in round 1 and removed.)
  auto for_range = [&](int i0, int i1, int j0, int j1, const int* __restrict__ mask, auto&& body) {
    const int ni = i1 - i0 + 1;
    const long n = (long)ni * (j1 - j0 + 1);
    if (tid >= n) return;
    const unsigned t0 = (unsigned)tid, un = (unsigned)ni, st = (unsigned)nthr;
    int j = (int)(t0 / un), i = (int)(t0 - (unsigned)j * un);
    const int dj = (int)(st / un), di = (int)(st - (unsigned)dj * un);
    const int nj = j1 - j0 + 1;
    while (j < nj) {
      const long x = ix2(g, i0 + i, j0 + j);
      if (mask[x] == 1) body(x);
      i += di; j += dj;
      if (i >= ni) { i -= ni; ++j; }
    }
  };
  // the same walk without a mask test, handing (x, i, j) to the body (fused first phase: two masks)
  auto for_range_ij = [&](int i0, int i1, int j0, int j1, auto&& body) {
    const int ni = i1 - i0 + 1;
    const long n = (long)ni * (j1 - j0 + 1);
    if (tid >= n) return;
    const unsigned t0 = (unsigned)tid, un = (unsigned)ni, st = (unsigned)nthr;
    int j = (int)(t0 / un), i = (int)(t0 - (unsigned)j * un);
    const int dj = (int)(st / un), di = (int)(st - (unsigned)dj * un);
    const int nj = j1 - j0 + 1;
    while (j < nj) {
      body(ix2(g, i0 + i, j0 + j), i0 + i, j0 + j);
      i += di; j += dj;
      if (i >= ni) { i -= ni; ++j; }
    }
  };
  for (int lll = S.lll0; lll < S.lll0 + S.nsub; ++lll) {
    V.wo = S.woa * lll + S.wob; V.wn = S.wna * lll + S.wnb; V.wm = 1. - V.wo - V.wn;
    V.pb_ml = pb_t + (long)(pml - 1) * L; V.pb_nl = pb_t + (long)(pnl - 1) * L;
    double* const pb_new = FUSE ? pb_t + (long)(psp - 1) * L : V.pb_nl;
    V.ub_ml = ub_t + (long)(ml - 1) * L; V.ub_nl = ub_t + (long)(nl - 1) * L;
    V.vb_ml = vb_t + (long)(ml - 1) * L; V.vb_nl = vb_t + (long)(nl - 1) * L;
    if (lll % 2 == 1) {
      if (S.p2p) {
        // band edges: pack the nh edge rows of the three fields (2 levels each) straight into the
        // neighbours' mailboxes over NVLink, publish the sequence number, wait for theirs, unpack
        const int parity = (int)(seq & 1ull);
        const int ii = g.ii;
        const long npay = 14L * ii;   // (2+2+3) rows x 2 levels
        for (int dir = 0; dir < 2; ++dir) {
          if (!(dir == 0 ? X.has_s : X.has_n)) continue;
          double* q = p2p_slot(X.peer[dir], X.cap, 1 - dir, parity);
          for (long t = tid; t < npay; t += nthr) {
            const int row = (int)(t / ii), i = (int)(t % ii) + 1;
            // rows 0..3: pb_t (lev1 r0,r1, lev2 r0,r1); 4..7: ub_t; 8..13: vb_t (3 rows per level)
            const double* a; int nh, rl;
            if (row < 4) { a = nullptr; nh = 2; rl = row; } else if (row < 8) { a = ub_t; nh = 2; rl = row - 4; }
            else { a = vb_t; nh = 3; rl = row - 8; }
            const int k = rl / nh, rr = rl % nh;
            const int j = dir == 0 ? 1 + rr : g.jj - nh + 1 + rr;
            // bottom pressure: "level 1" = the mid level, "level 2" = the old/new level (every rank rotates alike)
            const double* lvl = a ? a + (long)k * L : (k == 0 ? V.pb_ml : V.pb_nl);
            q[t] = __ldcg(lvl + ix2(g, i, j));
          }
        }
        if (tid < npay) __threadfence_system();   // only threads that stored to a peer
        grid_barrier(ctr, target);
        if (tid == 0) {
          // release: the grid barrier made every block's (system-fenced) mailbox stores visible to this
          // thread; the system-scope fence orders them before the flag stores the peers acquire on
          __threadfence_system();
          if (X.has_s) *(volatile unsigned long long*)p2p_word(X.peer[0], 1) = seq;
          if (X.has_n) *(volatile unsigned long long*)p2p_word(X.peer[1], 0) = seq;
        }
        if (threadIdx.x == 0) {
          if (X.has_s) { volatile unsigned long long* f = p2p_word(X.my_block, 0); while (*f < seq) {} }
          if (X.has_n) { volatile unsigned long long* f = p2p_word(X.my_block, 1); while (*f < seq) {} }
          __threadfence_system();
        }
        __syncthreads();
        for (int dir = 0; dir < 2; ++dir) {
          if (!(dir == 0 ? X.has_s : X.has_n)) continue;
          const double* q = p2p_slot(X.my_block, X.cap, dir, parity);
          for (long t = tid; t < npay; t += nthr) {
            const int row = (int)(t / ii), i = (int)(t % ii) + 1;
            double* a; int nh, rl;
            if (row < 4) { a = nullptr; nh = 2; rl = row; } else if (row < 8) { a = ub_t; nh = 2; rl = row - 4; }
            else { a = vb_t; nh = 3; rl = row - 8; }
            const int k = rl / nh, rr = rl % nh;
            const int j = dir == 0 ? 1 - nh + rr : g.jj + 1 + rr;
            double* lvl = a ? a + (long)k * L : (k == 0 ? V.pb_ml : V.pb_nl);
            lvl[ix2(g, i, j)] = __ldcg(q + t);
          }
        }
        grid_barrier(ctr, target);
        ++seq;
      }
      if (S.inkernel_halo) {
        // xctilr(pb_t,1,2,2,2,halo_ps), (ubflx_t,..,halo_uv), (vbflx_t,1,2,2,3,halo_vv)  (:395-397)
        for (int k = 1; k <= 2; ++k) {
          halo_level<true>(g, k == 1 ? V.pb_ml : V.pb_nl, halo_ps, k, 2, 2, 1, 1, tid, nthr);
          halo_level<true>(g, ub_t + (long)(k - 1) * L, halo_uv, k, 2, 2, 1, 1, tid, nthr);
          halo_level<true>(g, vb_t + (long)(k - 1) * L, halo_vv, k, 2, 3, 1, 1, tid, nthr);
        }
        grid_barrier(ctr, target);
      }
      if (FUSE) {
        for_range_ij(-1, g.ii + 1, -1, g.jj + 2, [&](long x, int i, int j) {
          const bool wu = i >= 0 && P.iu[x] == 1;
          if (P.ip[x] != 1 && !wu) return;
          const double pbc = btp_continuity_value(g, P, V, x);
          if (P.ip[x] == 1) pb_new[x] = pbc;
          if (wu) btp_ueq<WM>(g, P, V, V.vb_ml, x, pbc, btp_continuity_value(g, P, V, x - 1));
        });
        grid_barrier(ctr, target);
      } else {
        for_range(-1, g.ii + 1, -1, g.jj + 2, P.ip, [&](long x) { btp_continuity(g, P, V, x); });
        grid_barrier(ctr, target);
        for_range(0, g.ii + 1, -1, g.jj + 2, P.iu, [&](long x) {
          btp_ueq<WM>(g, P, V, V.vb_ml, x, __ldcg(V.pb_nl + x), __ldcg(V.pb_nl + x - 1)); });
        grid_barrier(ctr, target);
      }
      for_range(0, g.ii, 0, g.jj + 2, P.iv, [&](long x) {
        btp_veq<WM>(g, P, V, V.ub_nl, x, __ldcg(pb_new + x), __ldcg(pb_new + x - g.ldi)); });
      grid_barrier(ctr, target);
    } else {
      if (FUSE) {
        for_range_ij(0, g.ii, 0, g.jj + 1, [&](long x, int i, int j) {
          const bool wv = j >= 1 && P.iv[x] == 1;
          if (P.ip[x] != 1 && !wv) return;
          const double pbc = btp_continuity_value(g, P, V, x);
          if (P.ip[x] == 1) pb_new[x] = pbc;
          if (wv) btp_veq<WM>(g, P, V, V.ub_ml, x, pbc, btp_continuity_value(g, P, V, x - g.ldi));
        });
        grid_barrier(ctr, target);
      } else {
        for_range(0, g.ii, 0, g.jj + 1, P.ip, [&](long x) { btp_continuity(g, P, V, x); });
        grid_barrier(ctr, target);
        for_range(0, g.ii, 1, g.jj + 1, P.iv, [&](long x) {
          btp_veq<WM>(g, P, V, V.ub_ml, x, __ldcg(V.pb_nl + x), __ldcg(V.pb_nl + x - g.ldi)); });
        grid_barrier(ctr, target);
      }
      for_range(1, g.ii, 1, g.jj, P.iu, [&](long x) {
        btp_ueq<WM>(g, P, V, V.vb_nl, x, __ldcg(pb_new + x), __ldcg(pb_new + x - 1)); });
      grid_barrier(ctr, target);
    }
    if (FUSE) { const int t3 = psp; psp = pnl; pnl = pml; pml = t3; }   // (mid, old, spare) -> (spare, mid, old)
    else { const int t2 = pml; pml = pnl; pnl = t2; }
    const int t = ml; ml = nl; nl = t;
  }
}

struct HvP {
  double *pb, *pbu, *pbv, *ub, *vb, *ubflx, *vbflx, *ubflxs, *vbflxs, *ubflxs_p, *vbflxs_p;
  double *pb_p, *pbu_p, *pbv_p, *ubcors_p, *vbcors_p, *pb_mn, *ub_mn, *vb_mn;
  const double *scuy, *scvx;
};

// :847-977
__global__ void bt_harvest(Geom g, BtP P, HvP H, int nb, int m, int n, int ml, int nl) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
  if (i > g.ii) return;
  const long x = ix2(g, i, j), s = g.ldi, L = g.lev;
  const long xm = x + (long)(m - 1) * L, xn = x + (long)(n - 1) * L, x3 = x + 2 * L;
  // after the last swap P.pb_ml is level `ml`
  const bool isp = P.ip[x] == 1, isu = P.iu[x] == 1, isv = P.iv[x] == 1;
  const double pbc = P.pb_ml[x];
  if (nb == 1 || nb == 3) {
    const long xl = nb == 1 ? xm : xn;
    if (isp) H.pb[xl] = pbc;
    if (isu) {
      const double pu = fmin(pbc, P.pb_ml[x - 1]);
      H.pbu[xl] = pu;
      const double f = P.ub_ml[x];
      H.ubflx[xl] = f;
      H.ub[xl] = f / (pu * H.scuy[x]);
      if (nb == 1) {
        H.ubflxs[xn] = H.ubflxs[xn] + P.ubflxs_t[x];
        H.ubflxs[xm] = H.ubflxs[x3] + P.ubflxs_t[x];
      } else {
        H.ubflxs_p[xm] = H.ubflxs[xm] + P.ubflxs_t[x];
        H.ubflxs_p[xn] = H.ubflxs_p[xn] + P.ubflxs_t[x];
        H.ubcors_p[x] = H.ubcors_p[x] + P.ubcors_t[x];
      }
    }
    if (isv) {
      const double pv = fmin(pbc, P.pb_ml[x - s]);
      H.pbv[xl] = pv;
      const double f = P.vb_ml[x];
      H.vbflx[xl] = f;
      H.vb[xl] = f / (pv * H.scvx[x]);
      if (nb == 1) {
        H.vbflxs[xn] = H.vbflxs[xn] + P.vbflxs_t[x];
        H.vbflxs[xm] = H.vbflxs[x3] + P.vbflxs_t[x];
      } else {
        H.vbflxs_p[xm] = H.vbflxs[xm] + P.vbflxs_t[x];
        H.vbflxs_p[xn] = H.vbflxs_p[xn] + P.vbflxs_t[x];
        H.vbcors_p[x] = H.vbcors_p[x] + P.vbcors_t[x];
      }
    }
  } else if (nb == 2) {
    const long xml = x + (long)(ml - 1) * L, xnl = x + (long)(nl - 1) * L;
    if (isp) { H.pb_mn[xml] = P.pb_ml[x]; H.pb_mn[xnl] = P.pb_nl[x]; }
    if (isu) {
      H.ub_mn[xml] = P.ub_ml[x]; H.ub_mn[xnl] = P.ub_nl[x];
      H.ubflxs[xm] = H.ubflxs[xm] + P.ubflxs_t[x];
      H.ubflxs[x3] = P.ubflxs_t[x];
      H.ubflxs_p[xn] = P.ubflxs_t[x];
      H.ubcors_p[x] = P.ubcors_t[x];
    }
    if (isv) {
      H.vb_mn[xml] = P.vb_ml[x]; H.vb_mn[xnl] = P.vb_nl[x];
      H.vbflxs[xm] = H.vbflxs[xm] + P.vbflxs_t[x];
      H.vbflxs[x3] = P.vbflxs_t[x];
      H.vbflxs_p[xn] = P.vbflxs_t[x];
      H.vbcors_p[x] = P.vbcors_t[x];
    }
  } else {
    if (nb == 5 && isp) H.pb_p[x] = pbc;
    if (isu) {
      if (nb == 5) H.pbu_p[x] = fmin(pbc, P.pb_ml[x - 1]);
      H.ubflxs_p[xn] = H.ubflxs_p[xn] + P.ubflxs_t[x];
      H.ubcors_p[x] = H.ubcors_p[x] + P.ubcors_t[x];
    }
    if (isv) {
      if (nb == 5) H.pbv_p[x] = fmin(pbc, P.pb_ml[x - s]);
      H.vbflxs_p[xn] = H.vbflxs_p[xn] + P.vbflxs_t[x];
      H.vbcors_p[x] = H.vbcors_p[x] + P.vbcors_t[x];
    }
  }
}

}  // namespace

void barotp_dev(int m, int n, int mm, int nn, int k1m, int k1n) {
  (void)mm; (void)k1m; (void)k1n;
  Ctx& c = C(); const Geom& g = c.g;
  const long L = g.lev;
  const int lstep = (int)c.scalar("lstep");
  const std::string mommth = c.option("mommth", "enscon");
  if (mommth != "enscon" && mommth != "enecon" && mommth != "enedis")
    throw std::runtime_error(" mommth = " + mommth + " is unsupported!");
  // the three subcycled fields (pb_t, ubflx_t, vbflx_t of phy/mod_barotp.F90:155-157, two time levels each) live in
  // one allocation so that one L2 access-policy window can cover them (see below)
  double* bt_state = c.owned("barotp_state", 7);
  // pb_t has three levels: the fused form of the persistent kernel writes the new bottom pressure into a spare
  // level and rotates (bt_subcycle); pbl = levels holding (mid, old/new, spare)
  double *pb_t = bt_state, *ub_t = bt_state + 3 * L, *vb_t = bt_state + 5 * L;
  int pbl[3] = {1, 2, 3};
  double *umaxb = c.owned("barotp_umaxb", 1), *uminb = c.owned("barotp_uminb", 1), *vmaxb = c.owned("barotp_vmaxb", 1),
         *vminb = c.owned("barotp_vminb", 1), *uglue = c.owned("barotp_uglue", 1), *vglue = c.owned("barotp_vglue", 1);
  const dim3 gint(cdiv(g.ii, 128), g.jj);
  LAUNCH(bt_prologue, gint, 128, 0, g, m, nn, c.scalar("cwbdts", 0.0), c.scalar("cwbdls", 25.0), c.idev("iu"),
         c.idev("iv"), c.dev("u"), c.dev("v"), c.dev("pbu"), c.dev("pbv"), c.dev("umax"), c.dev("vmax"),
         c.dev("scuy"), c.dev("scvx"), umaxb, uminb, uglue, vmaxb, vminb, vglue);
  double* pvn = c.dev("pvtrop") + (long)(n - 1) * L;
  {
    dim3 grid(cdiv(g.ii + 2, 128), g.jj + 6);
    LAUNCH(bt_pvtrop, grid, 128, 0, g, c.idev("iu"), c.idev("iv"), c.idev("iq"), c.dev("corioq"), c.dev("pb_p"), pvn,
           c.dev("pvtrop_o"));
  }
  const long on2 = (long)(n - 1) * L, om2 = (long)(m - 1) * L;
  halo_update(std::vector<HaloReq>{{uglue, 1, halo_us}, {c.dev("utotn"), 1, halo_uv}, {umaxb, 1, halo_us},
                                   {uminb, 1, halo_us}, {vglue, 1, halo_vs}, {c.dev("vtotn"), 1, halo_vv},
                                   {vmaxb, 1, halo_vs}, {vminb, 1, halo_vs}, {c.dev("pgfxm") + on2, 1, halo_uv},
                                   {c.dev("xixp") + on2, 1, halo_us}, {c.dev("xixm") + on2, 1, halo_us},
                                   {c.dev("pgfym") + on2, 1, halo_vv}}, 1, 2);
  halo_update(std::vector<HaloReq>{{c.dev("xiyp") + on2, 1, halo_vs}, {c.dev("xiym") + on2, 1, halo_vs}}, 1, 2);
  halo_update(pvn, 1, 1, 3, halo_qs);
  if (g.nreg == 2 && g.north) {
    dim3 grid(cdiv(g.ii + 2, 128), 3);
    LAUNCH(bt_arctic_swap, grid, 128, 0, g, umaxb, uminb, c.dev("xixp") + on2, c.dev("xixm") + on2, vmaxb, vminb,
           c.dev("xiyp") + on2, c.dev("xiym") + on2);
  }

  BtP P{};
  P.scp2i = c.dev("scp2i"); P.scvxi = c.dev("scvxi"); P.scuyi = c.dev("scuyi"); P.scuxi = c.dev("scuxi");
  P.scvyi = c.dev("scvyi"); P.scuy = c.dev("scuy"); P.scvx = c.dev("scvx");
  P.pvo = c.dev("pvtrop_o"); P.pvm = c.dev("pvtrop") + om2; P.pvn = pvn;
  P.pgfxm_o = c.dev("pgfxm_o"); P.xixp_o = c.dev("xixp_o"); P.xixm_o = c.dev("xixm_o");
  P.pgfxm_m = c.dev("pgfxm") + om2; P.xixp_m = c.dev("xixp") + om2; P.xixm_m = c.dev("xixm") + om2;
  P.pgfxm_n = c.dev("pgfxm") + on2; P.xixp_n = c.dev("xixp") + on2; P.xixm_n = c.dev("xixm") + on2;
  P.pgfym_o = c.dev("pgfym_o"); P.xiyp_o = c.dev("xiyp_o"); P.xiym_o = c.dev("xiym_o");
  P.pgfym_m = c.dev("pgfym") + om2; P.xiyp_m = c.dev("xiyp") + om2; P.xiym_m = c.dev("xiym") + om2;
  P.pgfym_n = c.dev("pgfym") + on2; P.xiyp_n = c.dev("xiyp") + on2; P.xiym_n = c.dev("xiym") + on2;
  P.utotn = c.dev("utotn"); P.vtotn = c.dev("vtotn"); P.uglue = uglue; P.vglue = vglue;
  P.umaxb = umaxb; P.uminb = uminb; P.vmaxb = vmaxb; P.vminb = vminb;
  P.ubflxs_t = c.owned("barotp_ubflxs_t", 1); P.ubcors_t = c.owned("barotp_ubcors_t", 1);
  P.vbflxs_t = c.owned("barotp_vbflxs_t", 1); P.vbcors_t = c.owned("barotp_vbcors_t", 1);
  P.ip = c.idev("ip"); P.iu = c.idev("iu"); P.iv = c.idev("iv");
  P.dlt = c.scalar("dlt");
  P.enscon = mommth == "enscon";
  HvP H{};
  H.pb = c.dev("pb"); H.pbu = c.dev("pbu"); H.pbv = c.dev("pbv"); H.ub = c.dev("ub"); H.vb = c.dev("vb");
  H.ubflx = c.dev("ubflx"); H.vbflx = c.dev("vbflx"); H.ubflxs = c.dev("ubflxs"); H.vbflxs = c.dev("vbflxs");
  H.ubflxs_p = c.dev("ubflxs_p"); H.vbflxs_p = c.dev("vbflxs_p"); H.pb_p = c.dev("pb_p"); H.pbu_p = c.dev("pbu_p");
  H.pbv_p = c.dev("pbv_p"); H.ubcors_p = c.dev("ubcors_p"); H.vbcors_p = c.dev("vbcors_p");
  H.pb_mn = c.dev("pb_mn"); H.ub_mn = c.dev("ubflx_mn"); H.vb_mn = c.dev("vbflx_mn");
  H.scuy = c.dev("scuy"); H.scvx = c.dev("scvx");

  auto set_levels = [&](int ml, int nl) {
    P.pb_ml = pb_t + (long)(pbl[0] - 1) * L; P.pb_nl = pb_t + (long)(pbl[1] - 1) * L;
    P.ub_ml = ub_t + (long)(ml - 1) * L; P.ub_nl = ub_t + (long)(nl - 1) * L;
    P.vb_ml = vb_t + (long)(ml - 1) * L; P.vb_nl = vb_t + (long)(nl - 1) * L;
  };
  auto range_launch = [&](const char* name, auto kern, const double* extra, int i0, int i1, int j0, int j1) {
    dim3 grid(cdiv(i1 - i0 + 1, 128), j1 - j0 + 1);
    LAUNCH_NAMED(name, kern, grid, 128, 0, g, P, extra, i0, i1, j0, j1);
  };

  // cooperative persistent form unless option barotp_kernel=phases asks for one launch per phase
  const bool persistent = c.option("barotp_kernel", "persistent") != "phases";
  // two phases per substep (continuity inside the first momentum phase) unless barotp_fuse=0
  const bool fuse = persistent && c.option("barotp_fuse", "1") != "0";
  // block shape of the persistent kernel: threads x resident blocks per SM fixes the register budget
  // (65536 / (threads*blocks)); development switch barotp_shape = "512x2" (64 regs) | "768x2" (40) | "1024x2" (32).
  // Measured at tnx0.25v4 in round 1: 256x2 29.6 ms, 512x2 22.7, 640x2 22.2, 768x2 21.5, 1024x2 21.6.
  struct Shape { const char* name; const void* fn[8]; int threads; };   // fn[wmode + 4*fuse]
#define BT_SHAPE(T, B) {#T "x" #B, {(const void*)bt_subcycle<T, B, W_ALL, false>, (const void*)bt_subcycle<T, B, W_NO_N, false>,   \
                                    (const void*)bt_subcycle<T, B, W_NO_O, false>, (const void*)bt_subcycle<T, B, W_ONLY_N, false>, \
                                    (const void*)bt_subcycle<T, B, W_ALL, true>, (const void*)bt_subcycle<T, B, W_NO_N, true>,     \
                                    (const void*)bt_subcycle<T, B, W_NO_O, true>, (const void*)bt_subcycle<T, B, W_ONLY_N, true>}, T}
  static const Shape shapes[] = {BT_SHAPE(512, 2), BT_SHAPE(768, 2), BT_SHAPE(1024, 2)};
  // (small grids, tnx1v4, two-phase form: 512x2 1.76 ms, 1024x1 1.93, 384x2 1.99, 768x2 2.28)
#undef BT_SHAPE
  // default: 768x2 where every thread walks several cells per phase (1.67 M points at tnx0.25v4: 21.5 ms
  // against 22.7 for 512x2), 512x2 (no spills) where a phase is a single cell per thread and the time
  // goes into the chain of one cell plus the grid barrier (tnx1v4: 2.05 ms against 2.46)
  const bool many_cells = (long)(g.ii + 3) * (g.jj + 4) >= 4L * 148 * 1536;
  const std::string shape_opt = c.option("barotp_shape", many_cells ? BT_SHAPE_DEFAULT : "512x2");
  const Shape* shape = nullptr;
  for (const Shape& sh : shapes) if (shape_opt == sh.name) shape = &sh;
  if (!shape) throw std::runtime_error("barotp: unknown barotp_shape " + shape_opt);
  // zero-weight arrays are skipped unless barotp_wskip=0 (development switch)
  const bool wskip = c.option("barotp_wskip", "1") != "0";
  static std::map<std::string, int> coop_grids;
  auto coop_grid_of = [&](int wm) {
    const std::string key = std::string(shape->name) + "w" + std::to_string(wm);
    int& cg_ = coop_grids[key];
    if (cg_ == 0) {
      int per_sm = 0, nsm = 0;
      CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, shape->fn[wm], shape->threads, 0));
      CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c.device));
      cg_ = std::max(1, per_sm) * nsm;
    }
    return cg_;
  };
  // (Round 2 negative result: pinning the six levels of the subcycled state in L2 with a persisting
  // access-policy window - they carry 15 of the 53 words per point and substep - changed nothing at tnx0.25v4,
  // 18.26 ms with the window against 18.35 without, gpurun_out/r2d_kt_a/b.json, while a carve-out left in place
  // halved the speed of every other kernel of the step.  Removed.)
  unsigned* bar_ctr = reinterpret_cast<unsigned*>(c.owned("barotp_barrier", 1));
  int lll0 = 1, ml = 1, nl = 2;
  double woa = 0, wob = 0, wna = 0, wnb = 0;
  for (int nb = 1; nb <= 5; ++nb) {
    if (nb == 1) {
      lll0 = 1; ml = 1; nl = 2;
      pbl[0] = 1; pbl[1] = 2; pbl[2] = 3;
      woa = -1. / lstep; wob = .5 + (lll0 - .5) / lstep; wna = 0.; wnb = 0.;
      LAUNCH(bt_init_block1, gint, 128, 0, g, c.dev("pb_mn"), c.dev("ubflx_mn"), c.dev("vbflx_mn"), pb_t, ub_t, vb_t);
    } else if (nb == 2) {
      woa = 0.; wob = 0.; wna = 1. / lstep; wnb = -(lll0 - .5) / lstep;
    } else if (nb == 4) {
      wna = 0.; wnb = 1.;
    }
    {
      dim3 grid(cdiv(g.ii + 2, 128), g.jj + 4);
      LAUNCH(bt_zero_acc, grid, 128, 0, g, P);
    }
    if (persistent) {
      // one cooperative launch per halo interval: the whole block on one tile (halos refreshed
      // in-kernel), two substeps per launch when band edges have to be exchanged in between
      int lll = lll0;
      const int lend = lll0 + lstep / 2;
      while (lll < lend) {
        BtSched S{};
        S.lll0 = lll; S.ml = ml; S.nl = nl; S.woa = woa; S.wob = wob; S.wna = wna; S.wnb = wnb;
        S.pml = pbl[0]; S.pnl = pbl[1]; S.psp = pbl[2]; S.fuse = fuse ? 1 : 0;
        P2PView X{};
        if (g.nranks == 1) { S.nsub = lend - lll; S.inkernel_halo = 1; }
        else if (p2p_view(&X, (size_t)14 * g.ii)) {
          S.nsub = lend - lll; S.inkernel_halo = 1; S.p2p = 1;
          int nexch = 0;
          for (int l = lll; l < lend; ++l) nexch += l % 2;
          S.seq0 = p2p_reserve_seq(nexch);
        } else {
          S.inkernel_halo = 0;
          if (lll % 2 == 1) {
            halo_update(std::vector<HaloReq>{{pb_t + (long)(pbl[0] - 1) * L, 1, halo_ps}, {pb_t + (long)(pbl[1] - 1) * L, 1, halo_ps},
                                             {ub_t, 2, halo_uv}}, 2, 2);
            halo_update(vb_t, 2, 2, 3, halo_vv);
            S.nsub = std::min(2, lend - lll);
          } else S.nsub = 1;
        }
        CUDA_CHECK(cudaMemsetAsync(bar_ctr, 0, sizeof(unsigned), c.stream));
        void* args[] = {(void*)&g, (void*)&P, (void*)&S, (void*)&X, (void*)&pb_t, (void*)&ub_t, (void*)&vb_t, (void*)&bar_ctr};
        // time-weight pattern of this block (see Wsel): block 1 wn = 0; blocks 2,3 wo = 0; blocks 4,5 wn = 1
        const int wmode = (!wskip ? W_ALL : (nb == 1 ? W_NO_N : (nb <= 3 ? W_NO_O : W_ONLY_N))) + (fuse ? 4 : 0);
        launch_cooperative("bt_subcycle", shape->fn[wmode], coop_grid_of(wmode), shape->threads, args);
        if (S.nsub % 2 == 1) std::swap(ml, nl);
        for (int q = 0; q < S.nsub; ++q) {
          if (fuse) { const int t3 = pbl[2]; pbl[2] = pbl[1]; pbl[1] = pbl[0]; pbl[0] = t3; }
          else std::swap(pbl[0], pbl[1]);
        }
        lll += S.nsub;
      }
    } else {
      for (int lll = lll0; lll <= lll0 + lstep / 2 - 1; ++lll) {
        P.wo = woa * lll + wob; P.wn = wna * lll + wnb; P.wm = 1. - P.wo - P.wn;
        set_levels(ml, nl);
        if (lll % 2 == 1) {
          halo_update(std::vector<HaloReq>{{pb_t + (long)(pbl[0] - 1) * L, 1, halo_ps}, {pb_t + (long)(pbl[1] - 1) * L, 1, halo_ps},
                                           {ub_t, 2, halo_uv}}, 2, 2);
          halo_update(vb_t, 2, 2, 3, halo_vv);
          {
            dim3 grid(cdiv(g.ii + 3, 128), g.jj + 4);
            LAUNCH(bt_continuity, grid, 128, 0, g, P, -1, g.ii + 1, -1, g.jj + 2);
          }
          range_launch("bt_ueq", bt_ueq, P.vb_ml, 0, g.ii + 1, -1, g.jj + 2);
          range_launch("bt_veq", bt_veq, P.ub_nl, 0, g.ii, 0, g.jj + 2);
        } else {
          {
            dim3 grid(cdiv(g.ii + 1, 128), g.jj + 2);
            LAUNCH(bt_continuity, grid, 128, 0, g, P, 0, g.ii, 0, g.jj + 1);
          }
          range_launch("bt_veq", bt_veq, P.ub_ml, 0, g.ii, 1, g.jj + 1);
          range_launch("bt_ueq", bt_ueq, P.vb_nl, 1, g.ii, 1, g.jj);
        }
        std::swap(ml, nl);
        std::swap(pbl[0], pbl[1]);
      }
    }
    lll0 = lll0 + lstep / 2;
    set_levels(ml, nl);
    LAUNCH(bt_harvest, gint, 128, 0, g, P, H, nb, m, n, ml, nl);
  }
}

}  // namespace blom
