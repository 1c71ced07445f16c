// Neutral diffusion of tracers (phy/mod_ndiff.F90): ndiff_prep_jslice :959-1026, ndiff_flx :160-953
// (with peval :62-74, pmeval :76-102, drhoroot :104-148, drho :150-158), ndiff_uflx_jslice :1028-1088,
// ndiff_vflx_jslice :1090-1150 and ndiff_update_trc_jslice :1152-1175.
//
// The reference walks rotating j-slices inside the ALE regrid-remap pipeline
// (phy/mod_ale_regrid_remap.F90:1614-1690); its slice arrays are whole-domain arrays here, in the common
// (i,j,level) layout (names and level order: include/blomgpu.h, "neutral diffusion inputs").
//
// B200 design: the search for neutral sublayers between two columns is sequential and data dependent,
// so ONE THREAD OWNS ONE FACE COLUMN.  Three launches instead of the reference's slice pipeline:
//   ndiff_prep    per cell column: kdmx, drhodt/drhods at the source interfaces, the snapped destination
//                 interfaces (a pure function of the cell, so it is evaluated once per cell instead of
//                 once per face as in the reference), zero of the face accumulators - and the TRANSPOSE of
//                 everything the searches read into per-column records (below)
//   ndiff_face<u|v>  per face column: both searches, fluxes, layer binning of the face fluxes and the
//                 neutral slope.  The reference scatters flux convergences into the two cells
//                 (flxconv(kd,nt,i-1|i)); scattering would race between faces, so each face writes
//                 its contribution to its minus-side and plus-side cell into two face-owned buffers.
//                 The destination index only moves down the column, so the running sum of the current
//                 destination layer sits in registers and every buffer level is written exactly once
//                 (no memset, no read-modify-write).
//   ndiff_update  per cell and level: gathers the four face contributions in the reference's pipeline
//                 order (south v face, west u face, east u face, north v face) and updates trc_rm.
// The only floating-point reassociation against the reference is that several contributions of one
// face to the same destination layer are summed before they meet the cell's running total.
//
// What bounds ndiff_face and the layout that follows from it.  Every lane walks down its own two columns at
// its own pace, so in the (i,j,level) layout a warp's load touches ~10 different 32-byte sectors of which each
// lane uses 8 bytes, a thread meets a new sector for every array and level (~25 arrays x 53 levels x 2
// columns), and each of those first touches is an L2 or DRAM round trip in the middle of a dependent chain:
// ncu shows 52 % of the stall samples on the long scoreboard at 16 resident warps per SM, 9 % of the DRAM peak
// and issue slots 29 % busy; removing a quarter of the instructions (32-bit index arithmetic, round 2) gained
// 4 %.  The kernel therefore reads COLUMN RECORDS: ndiff_prep writes, per cell and source layer, one contiguous
// record with everything both searches need of that layer (interface records, interface pressures, polynomial
// coefficients, diffusivity, layer means; 192 bytes for T and S), and per destination interface the pair
// {p_dst, snapped p_dst}.  A thread that steps to the next source layer of a column loads that layer's
// record in ONE batch of independent 16-byte loads into its private column of shared memory and works from
// there, and it prefetches the NEXT layer's record into the L2: one memory round trip per layer and column
// instead of one per array, and that one an L2 hit; every fetched sector used completely, no coefficient /
// layer-mean reloads.  The thread-local partner tables of the searches get presence masks in shared memory (their
// scans are bit operations), the destination layers are walked with shift registers.  The model's own arrays keep
// their layout; the transpose costs one streaming pass inside ndiff_prep (written with 32-byte sector stores).
// Measurements: profiles/r02_tuning_log.md, profiles/r02_ncu_full_ndiff_face_records.txt.  Every step of this
// rebuild was checked bit for bit on the CPU first: tests/emul compiles THIS file for the host (BLOM_HOST_EMUL).
#include "common.cuh"
#include "eos.cuh"

namespace blom {

namespace {

constexpr int KMN = 64;    // compile-time bound on kdm for the thread-local arrays
constexpr int NTMAX = 8;   // bound on the number of diffused scalars (2 + ntr)
constexpr double ndiff_dstsnp_fac = .01, rho_eps = 1.e-5, dp_eps = 1.e-5;  // :39-42
constexpr double mval = 1.e30;                                            // :210
constexpr int IT = 1, IS = 2;                                             // :43-45

// phy/mod_eos.F90:220-241, :284-304
__device__ __forceinline__ double eos_drhodt(double p, double th, double s) {
  const double r1 = eos::P1(p, th, s);
  const double r2i = 1. / eos::P2(p, th, s);
  return (EA12 + 2. * EA14 * th + EA15 * s + EB12 * p - (EA22 + 2. * EA24 * th + EA25 * s + EB22 * p) * r1 * r2i) * r2i;
}
__device__ __forceinline__ double eos_drhods(double p, double th, double s) {
  const double r1 = eos::P1(p, th, s);
  const double r2i = 1. / eos::P2(p, th, s);
  return (EA13 + EA15 * th + 2. * EA16 * s + EB13 * p - (EA23 + EA25 * th + 2. * EA26 * s + EB23 * p) * r1 * r2i) * r2i;
}

// Column record of source layer k of one cell (RS = nd_rs(T) = 24 + 8*(T-2) doubles; where it starts: nd_col below):
//    0.. 3  {drhodt, drhods, T, S} at the upper interface (is = 1)        t_srcdi(1,k,1:2) and mod_eos derivatives
//    4.. 7  the same at the lower interface (is = 2)
//    8, 9   p_src(k), p_src(k+1)                                           p_srcdi(1:2,k)
//   10..14  tpc_src(1:5,k,T)      15..19  tpc_src(1:5,k,S)
//   20      difiso(k)             21, 22  temp(k), saln(k) at the new time level          23  unused
//   24 + 8*(nt-3) ..  passive tracer nt >= 3: trc(k), tpc_src(1:5,k,nt), t_srcdi(1:2,k,nt)
// The first ND_RSB = 24 doubles (192 bytes, six sectors) are what a thread loads when it stages a layer (20 of them
// are kept per search, see ndiff_face).
// Cells are grouped in blocks of CB consecutive cells (linear (i,j) offsets); inside a block the records are ordered
// (layer, cell), i.e. the record of layer k of cell x starts at src[(((x/CB)*kk + k-1)*CB + x%CB) * RS]: the
// layer-k records of 32 neighbouring cells are one contiguous 6 KB piece, which is what a warp of ndiff_prep writes
// and what the lanes of a warp of ndiff_face read while they are at the same depth (DRAM pages instead of 192-byte
// pieces a whole column apart).  CB = ND_CB = 32.
constexpr int ND_RSB = 24, F_REC = 0, F_P = 8, F_TPC = 10, F_DIF = 20, F_TLEV = 21;
constexpr int X_TLEV = 0, X_TPC = 1, X_TSD = 6;   // inside a passive tracer's 8 doubles
__host__ __device__ constexpr int nd_rs(int T) { return ND_RSB + 8 * (T - 2); }
constexpr int ND_CB = 32;
__host__ __device__ __forceinline__ long nd_col(long x, int kk, int RS) {   // record of layer 1 of cell x
  return ((x / ND_CB) * kk * ND_CB + x % ND_CB) * RS;
}
// Destination record of interface k (1..kk+1) of cell x: dst[(x*(kk+1) + k-1)*2] = {p_dst(k), p_dstsnp(k)}.

// 32-byte (one sector) global store of sm_100 (STG.E.ENL2.256).  ndiff_prep writes a record sector by sector: a lane's
// 16-byte stores leave half-written sectors behind (every lane writes to its own record), and those cost the pass a factor
// of two (tnx1v4: 1.17 ms with sixteen-byte stores, 0.55 ms with these; a writer with one thread per 16-byte piece, whose
// warps do write whole sectors, was as slow because its loads were scattered instead).
struct Quad { double a, b, c, d; };
#ifdef BLOM_HOST_EMUL
__device__ __forceinline__ void st_quad(double* p, Quad q) { p[0] = q.a; p[1] = q.b; p[2] = q.c; p[3] = q.d; }
#else
__device__ __forceinline__ void st_quad(double* p, Quad q) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(q.a), "d"(q.b), "d"(q.c), "d"(q.d) : "memory");
}
#endif

struct NdArgs {
  const double *src, *dst;     // column records written by ndiff_prep
  const int *ksmx, *kdmx, *mask;
  const int* faces; int nfaces;   // offsets ix2(i,j) of the wet faces of this direction
  const double* dpml;
  const double *sca, *scbi;    // scuy,scuxi | scvx,scvyi
  const double* puv;           // pu | pv
  double *tflld, *sflld, *tflx, *sflx, *nslp;
  double *cvm, *cvp;           // face contributions to the minus / plus side cell, level (nt-1)*kk+kd
  double delt1;
  int mm, T, surface_align;
};

struct Rec { double drdt, drds, t, s; };

// :104-148: Newton search for the position in a layer that is neutral to (tf,sf); c(q), q = 0..9, are the layer's
// polynomial coefficients of T and S (loaded once instead of once per iteration)
template <class CF>
__device__ __forceinline__ double drhoroot(CF c, double tf, double sf,
                                           double drhodt_l, double drhodt_u, double drhods_l, double drhods_u) {
  const double eps = 1.e-14, x_tol = 1.e-4;
  double x = .5;
  const double ddrdtdx = drhodt_l - drhodt_u, ddrdsdx = drhods_l - drhods_u;
  const double T1 = c(0), T2 = c(1), T3 = c(2), T4 = c(3), T5 = c(4);
  const double S1 = c(5), S2 = c(6), S3 = c(7), S4 = c(8), S5 = c(9);
  for (int n = 1; n <= 10; ++n) {
    const double dt = tf - (T1 + (T2 + (T3 + (T4 + T5 * x) * x) * x) * x);
    const double ds = sf - (S1 + (S2 + (S3 + (S4 + S5 * x) * x) * x) * x);
    const double drdt = drhodt_l * x + drhodt_u * (1. - x);
    const double drds = drhods_l * x + drhods_u * (1. - x);
    const double dtdx = -(T2 + (2. * T3 + (3. * T4 + 4. * T5 * x) * x) * x);
    const double dsdx = -(S2 + (2. * S3 + (3. * S4 + 4. * S5 * x) * x) * x);
    const double dr = drdt * dt + drds * ds;
    const double ddrdx = ddrdtdx * dt + drdt * dtdx + ddrdsdx * ds + drds * dsdx;
    const double x_old = x;
    x = fmax(0., fmin(1., x_old - dr / copysign(fmax(eps, fabs(ddrdx)), ddrdx)));
    if (fabs(x - x_old) < x_tol) return x;
  }
  return x;
}

// ndiff_prep_jslice (:959-1026) on 0..ii+1 x 0..jj+1 + the destination snapping of ndiff_flx (:491-523) + the
// column records
struct PrepIn {
  const int *ip, *iu, *iv, *ksmx;
  const double *p_src, *tsd, *tpc, *p_dst, *difiso;
  const double* tlev[NTMAX];   // scalar nt at time level nn, level 1
};
__global__ void __launch_bounds__(128)
ndiff_prep(Geom g, int mm, int T, PrepIn I, int* __restrict__ kdmx, double* __restrict__ src, double* __restrict__ dst,
           double* __restrict__ utflld, double* __restrict__ usflld, double* __restrict__ vtflld,
           double* __restrict__ vsflld, double* __restrict__ ucm, double* __restrict__ ucp, double* __restrict__ vcm,
           double* __restrict__ vcp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i > g.ii + 1) return;
  const long x = ix2(g, i, j), lev = g.lev;
  const int kk = g.kdm;
  const bool wu = I.iu[x] == 1, wv = I.iv[x] == 1;
  if (wu || wv)
    for (int k = 1; k <= kk; ++k) {
      const long o = x + (long)(k + mm - 1) * lev;
      if (wu) { utflld[o] = 0.; usflld[o] = 0.; }
      if (wv) { vtflld[o] = 0.; vsflld[o] = 0.; }
    }
  if (wu || wv)   // face-owned convergence buffers (SideAcc)
    for (int q = 0; q < T * kk; ++q) {
      const long o = x + (long)q * lev;
      if (wu) { ucm[o] = 0.; ucp[o] = 0.; }
      if (wv) { vcm[o] = 0.; vcp[o] = 0.; }
    }
  if (I.ip[x] != 1) return;
  const double pbot = I.p_dst[x + (long)kk * lev];
  int kd = kk;
  for (int k = kk; k >= 1; --k)
    if (I.p_dst[x + (long)(k - 1) * lev] == pbot) kd = k - 1;
  kdmx[x] = kd;
  const int ks = I.ksmx[x], RS = nd_rs(T);
  double* col = src + nd_col(x, kk, RS);
  const long LS = (long)ND_CB * RS;   // from one layer's record to the next
  double p_up = I.p_src[x];
  for (int k = 1; k <= ks; ++k) {
    double* r = col + (k - 1) * LS;
    const double p_lo = I.p_src[x + (long)k * lev];
    const long ot = x + (long)(((IT - 1) * kk + k - 1) * 2) * lev, os = x + (long)(((IS - 1) * kk + k - 1) * 2) * lev;
    const double t1 = I.tsd[ot], t2 = I.tsd[ot + lev], s1 = I.tsd[os], s2 = I.tsd[os + lev];
    st_quad(r, Quad{eos_drhodt(p_up, t1, s1), eos_drhods(p_up, t1, s1), t1, s1});
    st_quad(r + 4, Quad{eos_drhodt(p_lo, t2, s2), eos_drhods(p_lo, t2, s2), t2, s2});
    const double* ct = I.tpc + x + (long)(((IT - 1) * kk + k - 1) * 5) * lev;
    const double* cs = I.tpc + x + (long)(((IS - 1) * kk + k - 1) * 5) * lev;
    st_quad(r + 8, Quad{p_up, p_lo, ct[0], ct[lev]});
    st_quad(r + 12, Quad{ct[2 * lev], ct[3 * lev], ct[4 * lev], cs[0]});
    st_quad(r + 16, Quad{cs[lev], cs[2 * lev], cs[3 * lev], cs[4 * lev]});
    const long ol = x + (long)(k - 1) * lev;
    st_quad(r + 20, Quad{I.difiso[ol], I.tlev[0][ol], I.tlev[1][ol], 0.});
    for (int nt = 3; nt <= T; ++nt) {
      const double* cn = I.tpc + x + (long)(((nt - 1) * kk + k - 1) * 5) * lev;
      const long on = x + (long)(((nt - 1) * kk + k - 1) * 2) * lev;
      double* rn = r + ND_RSB + 8 * (nt - 3);
      st_quad(rn, Quad{I.tlev[nt - 1][ol], cn[0], cn[lev], cn[2 * lev]});
      st_quad(rn + 4, Quad{cn[3 * lev], cn[4 * lev], I.tsd[on], I.tsd[on + lev]});
    }
    p_up = p_lo;
  }
  // {p_dst(k), p_dstsnp(k)}: p_dstsnp(1..kdmx+1) as in the reference, p_dst below
  double2* d = reinterpret_cast<double2*>(dst + (long)x * (kk + 1) * 2);
  double pk = I.p_dst[x], pk1 = I.p_dst[x + lev];
  d[0] = make_double2(pk, pk);
  double dp_dst_u = pk1 - pk;
  const int kl = min(ks, kd);
  for (int k = 2; k <= kl; ++k) {
    pk = pk1; pk1 = I.p_dst[x + (long)k * lev];
    const double dp_dst_l = pk1 - pk;
    const double ps = I.p_src[x + (long)(k - 1) * lev];
    d[k - 1] = make_double2(pk, fabs(pk - ps) < fmin(dp_dst_u, dp_dst_l) * ndiff_dstsnp_fac ? ps : pk);
    dp_dst_u = dp_dst_l;
  }
  for (int k = kl + 1; k <= kk + 1; ++k) {
    const double p = I.p_dst[x + (long)(k - 1) * lev];
    d[k - 1] = make_double2(p, p);
  }
}

// face-owned running sums of the flux convergence of the current destination layer of one side.  The buffers are
// zeroed by ndiff_prep (coalesced); a face writes the layers it has contributions for, each exactly once (the zero
// fill of the layers in between and below used to be a tenth of the warp instructions of ndiff_face, issued by the
// last lanes alive of every warp, 8 bytes per store)
template <int NT, class IX>
struct SideAcc {
  double a[NT > 0 ? NT : NTMAX];
  double* buf; IX x, lev; int kk, T, cur;
  __device__ __forceinline__ void init(double* b, IX x_, IX l, int k, int t) {
    buf = b; x = x_; lev = l; kk = k; T = t; cur = 0;
#pragma unroll
    for (int q = 0; q < (NT > 0 ? NT : NTMAX); ++q) a[q] = 0.;
  }
  __device__ __forceinline__ void advance(int kd) {   // kd never decreases
    if (kd == cur) return;
#pragma unroll
    for (int q = 0; q < (NT > 0 ? NT : NTMAX); ++q)
      if (q < T) {
        if (cur > 0) buf[x + (IX)(q * kk + cur - 1) * lev] = a[q];
        a[q] = 0.;
      }
    cur = kd;
  }
  __device__ __forceinline__ void finish() { advance(kk + 1); }
};

// ndiff_flx (:160-953) for the face between cell M (i-1|j-1) and cell P (i,j)
// 128 threads per block at a register budget of 128 per thread (512 resident threads per SM).
// IX is the type the index arithmetic of the level-strided OUTPUT arrays is done in (`unsigned` when every
// element index fits 32 bits - one IMAD and one IMAD.WIDE per address instead of a 64-bit product - else `long`).
// STG: how the record of a column's current source layer is held (option ndiff_stage):
//   1  staged in shared memory when the column steps to the layer (one batch of loads)
//   0  read from the column record in global memory whenever needed (no shared memory, the whole L1 is cache)
//   2  as 0, with an L1 prefetch of the record's lines when the column steps to the layer
//   4  as 0, with an L2 prefetch of the next layer's record
//   3  as 1, with an L2 prefetch of the NEXT layer's record (default; tnx1v4, u + v faces: 12.9 / 12.2 / 12.5 / 10.3 ms
//      for 0 / 1 / 2 / 3: the prefetch turns the DRAM round trip of every layer step into an L2 hit)
constexpr int ND_BS = 128;
#ifdef BLOM_HOST_EMUL
__device__ __forceinline__ void nd_prefetch_l1(const void*) {}
__device__ __forceinline__ void nd_prefetch_l2(const void*) {}
#else
__device__ __forceinline__ void nd_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void nd_prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#endif
template <int DIR, int NT, class IX, int STG>
__global__ void __launch_bounds__(ND_BS, 512 / ND_BS)
ndiff_face(Geom g, NdArgs A) {
  // wet faces only: thread t owns face A.faces[t] (linear (i,j) offset of the face's plus-side cell).  A thread
  // of a land face would idle for the whole life of its warp - the kernel is latency-bound, so the
  // compacted list (built once, the masks are static) removes that share of the warps outright.
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= A.nfaces) return;
  const int tx = threadIdx.x;
  const IX x = (IX)A.faces[t];
  const IX xm = x - (IX)(DIR == 0 ? 1 : g.ldi), lev = (IX)g.lev;
  const int kk = g.kdm, T = NT > 0 ? NT : A.T, mm = A.mm;
  const int RS = NT > 0 ? nd_rs(NT) : nd_rs(T);
  // column records of the two cells (side 0 = M, 1 = P)
  const double* const col_m = A.src + nd_col((long)xm, kk, RS);
  const double* const col_p = A.src + nd_col((long)x, kk, RS);
  const int LS = ND_CB * RS;   // from one layer's record to the next
  const double* const dst_m = A.dst + (long)xm * (kk + 1) * 2;
  const double* const dst_p = A.dst + (long)x * (kk + 1) * 2;
  // The records of the current source layers of the two columns, staged: sm[side*ND_RSS + row][thread].  The first
  // search needs fields 0..19 (row = field); the second search needs neither drhodt nor drhods (fields 0, 1, 4, 5)
  // but the diffusivity and the layer means (20, 21, 22), which take rows 0, 1, 4 then: 20 rows per side, 40 KB
  // per block.  The remaining 8 KB of the 48 KB hold per-thread state that is used too rarely to deserve
  // registers at a budget of 128 (the presence masks of the partner tables, the upper interface of the current
  // destination layers).
  constexpr int ND_RSS = 20;
  __shared__ double sm[(STG & 1) ? 2 * ND_RSS : 1][ND_BS];
  __shared__ unsigned long long smk[4][ND_BS];   // [side*2 + word]
  __shared__ double smd[4][ND_BS];               // [side*2 + {p_dst, p_dstsnp}]
  const double* lrec_m = col_m;   // records of the current layers (STG even)
  const double* lrec_p = col_p;
#define ND_ROW(f) ((f) == F_DIF ? 0 : (f) == F_TLEV ? 1 : (f) == F_TLEV + 1 ? 4 : (f))
#define SM(side, f) ((STG & 1) ? sm[(STG & 1) ? (side) * ND_RSS + ND_ROW(f) : 0][tx] : ((side) ? lrec_p : lrec_m)[f])
  auto stage = [&](int side, int k, bool second) {
    const double* rec = (side ? col_p : col_m) + (k - 1) * LS;
    if (STG & 1) {
      // (16-byte loads: staging with six 32-byte loads measured 5 - 15 % slower, tnx1v4 u / v 4.88 / 5.42 against
      // 4.66 / 4.72 ms)
      const double2* r = reinterpret_cast<const double2*>(rec);
      double2 v[ND_RSB / 2];
#pragma unroll
      for (int q = 0; q < ND_RSB / 2; ++q) v[q] = r[q];
#pragma unroll
      for (int q = 0; q < ND_RSB / 2; ++q) {
        const int f0 = 2 * q, f1 = 2 * q + 1;
        const bool skip0 = second ? (f0 == 0 || f0 == 4) : f0 >= ND_RSS;   // fields 0,1 / 4,5 and 23 in the second
        const bool skip1 = second ? (f1 == 1 || f1 == 5 || f1 == 23) : f1 >= ND_RSS;   // search, 20..23 in the first
        if (!skip0) sm[(STG & 1) ? side * ND_RSS + ND_ROW(f0) : 0][tx] = v[q].x;
        if (!skip1) sm[(STG & 1) ? side * ND_RSS + ND_ROW(f1) : 0][tx] = v[q].y;
      }
      if (STG == 3 && k < kk) { nd_prefetch_l2(rec + LS); nd_prefetch_l2(rec + LS + 16); }
    } else {
      if (side) lrec_p = rec; else lrec_m = rec;
      if (STG == 2) { nd_prefetch_l1(rec); nd_prefetch_l1(rec + 16); }
      if (STG == 4 && k < kk) { nd_prefetch_l2(rec + LS); nd_prefetch_l2(rec + LS + 16); }
    }
  };
  // unstaged reads (layers other than the current ones): p_srcdi(s,k) = p_src(k+s-1), p_dst(k), p_dstsnp(k)
  auto psd = [&](int side, int s, int k) { return (side ? col_p : col_m)[(k - 1) * LS + F_P + s - 1]; };
  auto pdst = [&](int side, int k) { return (side ? dst_p : dst_m)[(k - 1) * 2]; };
  auto srec = [&](int side, int is) {   // interface record (is) of the staged layer
    const int f = F_REC + 4 * (is - 1);
    return Rec{SM(side, f), SM(side, f + 1), SM(side, f + 2), SM(side, f + 3)};
  };
  const int ksmx_m = A.ksmx[xm], ksmx_p = A.ksmx[x], kdmx_m = A.kdmx[xm], kdmx_p = A.kdmx[x];
  if (STG >= 3) {   // the destination records of both columns (16 bytes per interface) are read front to back
    for (int q = 0; q < 2 * (kk + 1); q += 16) { nd_prefetch_l2(dst_m + q); nd_prefetch_l2(dst_p + q); }
  }
  const double cdiff = A.delt1 * A.sca[x] * A.scbi[x];          // :1064 / :1126
  const double cnslp = alpha0 * A.scbi[x] / grav;

  // Partner tables PN(s,k), thread-local.  They are NOT initialised: an entry is only read where its presence bit
  // (below) is set, a clear bit stands for the reference's "mval"; filling 2 x 110 doubles per thread with mval
  // first cost 216 local stores and their share of the L1/L2.
  double pnm[2 * (KMN + 1) + 2], pnp[2 * (KMN + 1) + 2];
#ifdef BLOM_HOST_EMUL   // tests/emul: poison, so that a read of an entry that was never written cannot go unnoticed
  for (int q = 0; q < 2 * (KMN + 1) + 2; ++q) { pnm[q] = __builtin_nan(""); pnp[q] = __builtin_nan(""); }
#endif
#define PNM(s, k) pnm[2 * (k) + (s) - 1]
#define PNP(s, k) pnp[2 * (k) + (s) - 1]
  unsigned long long stab_m = 0ull, stab_p = 0ull;   // bit k-1 <-> stab(k), k = 1..64
  // Presence masks of the partner tables: bit q = 2*k + s - 1 (2..127 for kk <= 63) of hm / hp is set exactly when
  // PNM(s,k) / PNP(s,k) != mval.  The second search scans the tables for the next interface that has a partner;
  // with the masks the scans are bit operations and only the values actually used are read from local memory.
  smk[0][tx] = 0ull; smk[1][tx] = 0ull; smk[2][tx] = 0ull; smk[3][tx] = 0ull;
  auto hset = [&](int side, int q) { smk[side * 2 + (q >> 6)][tx] |= 1ull << (q & 63); };
  auto hget = [&](int side, int q) { return ((smk[side * 2 + (q >> 6)][tx] >> (q & 63)) & 1ull) != 0ull; };
#define PNM_SET(s, k, v) do { PNM(s, k) = (v); hset(0, 2 * (k) + (s) - 1); } while (0)
#define PNP_SET(s, k, v) do { PNP(s, k) = (v); hset(1, 2 * (k) + (s) - 1); } while (0)
  double pml = 0., drho_curr = 0., p_ni_m_prev, p_ni_p_prev;
  int nns = 0, kssa_m = 0, kssa_p = 0, is_m, is_p, ks_m, ks_p;

  // records of the current interfaces (is_m,ks_m) and (is_p,ks_p); only the side that moved is reloaded
  Rec rm{0., 0., 0., 0.}, rp{0., 0., 0., 0.};
  auto drho_at = [&]() {
    return .5 * (rm.drdt + rp.drdt) * (rp.t - rm.t) + .5 * (rm.drds + rp.drds) * (rp.s - rm.s);
  };

  // Neutral slope at the destination interfaces (:913-951), evaluated while the first search produces the
  // (slope, pressure) pairs instead of from stored lists afterwards.  The reference scans the destination
  // interfaces kd = 1..kk in order and for each one advances a source pointer ks to the first pair at or
  // below it, never moving ks back; handing every new pair to the still-unfilled destination interfaces
  // visits exactly the same (kd, ks) combinations, so the interpolated values are identical and the two
  // lists of up to 4*(kk+1) doubles per thread never exist.
  int kd_sl = 1;
  double pd_sl = .5 * (pdst(0, 1) + pdst(1, 1)), s_prev = 0., p_prev = 0.;
  auto emit_slope = [&](double sl, double pr) {
    nns = nns + 1;
    while (kd_sl <= kk && !(pd_sl > pr)) {
      if (nns == 1) A.nslp[x + (IX)(kd_sl - 1) * lev] = sl;
      else {
        const double q = (pr - pd_sl) / fmax(pr - p_prev, epsilp);
        A.nslp[x + (IX)(kd_sl - 1) * lev] = q * s_prev + (1. - q) * sl;
      }
      kd_sl = kd_sl + 1;
      if (kd_sl <= kk) pd_sl = .5 * (pdst(0, kd_sl) + pdst(1, kd_sl));
    }
    s_prev = sl; p_prev = pr;
  };

  // ---- first search: neutral interfaces anchored at source layer interfaces (:212-406)
  if (A.surface_align) {
    pml = .5 * (psd(0, 1, 1) + A.dpml[xm] + psd(1, 1, 1) + A.dpml[x]);
    kssa_m = 2;
    while (kssa_m <= ksmx_m) {
      if (psd(0, 1, kssa_m) > pml) break;
      kssa_m = kssa_m + 1;
    }
    kssa_p = 2;
    while (kssa_p <= ksmx_p) {
      if (psd(1, 1, kssa_p) > pml) break;
      kssa_p = kssa_p + 1;
    }
    is_m = 1; ks_m = kssa_m; is_p = 1; ks_p = kssa_p;
    p_ni_m_prev = pml; p_ni_p_prev = pml;
  } else {
    is_m = 1; ks_m = 1; is_p = 1; ks_p = 1;
    p_ni_m_prev = psd(0, 1, 1); p_ni_p_prev = psd(1, 1, 1);
  }
  if (ks_m <= ksmx_m && ks_p <= ksmx_p) {
    stage(0, ks_m, false); stage(1, ks_p, false);
    rm = srec(0, is_m); rp = srec(1, is_p); drho_curr = drho_at();
  }

  // search_loop1.  The reference handles the minus and the plus column in separate, mirrored code blocks
  // (root search in M when drho < 0, in P when drho > 0; advance M, then advance P).  Lanes of a warp sit in
  // different blocks at the same time, so the mirrored blocks are written ONCE with the column chosen per
  // lane (`side`: 0 = M, 1 = P): lanes that advance M and lanes that advance P, or that solve for a root in
  // M and in P, then execute together instead of one after the other.  The arithmetic per lane is the
  // reference's (sums are commuted only where a + b == b + a exactly).  Every read of this loop is of the
  // current layer of a column, i.e. of the staged records.
  {
    bool done1 = false;
    while (!done1 && ks_m <= ksmx_m && ks_p <= ksmx_p) {
      const bool drho_neg = drho_curr <= -rho_eps;
      const bool drho_pos = drho_curr >= rho_eps;
      const bool drho_zero = !(drho_neg || drho_pos);
      if (is_m + ks_m > 2 && is_p + ks_p > 2) {
        const bool rootm = drho_neg && is_m == 2, rootp = drho_pos && is_p == 2;
        if (rootm || rootp) {
          // the current layer of the searched column (side sr) against the fixed interface fx of the other column
          const int sr = rootm ? 0 : 1;
          const Rec fx = rootm ? rp : rm;
          const double drhodt_x0 = .5 * (SM(sr, F_REC) + fx.drdt), drhodt_x1 = .5 * (SM(sr, F_REC + 4) + fx.drdt);
          const double drhods_x0 = .5 * (SM(sr, F_REC + 1) + fx.drds), drhods_x1 = .5 * (SM(sr, F_REC + 5) + fx.drds);
          const double x_ni = drhoroot([&](int q) { return SM(sr, F_TPC + q); }, fx.t, fx.s, drhodt_x1, drhodt_x0,
                                       drhods_x1, drhods_x0);
          const double p_ni = SM(sr, F_P + 1) * x_ni + SM(sr, F_P) * (1. - x_ni);
          if (p_ni > (rootm ? p_ni_m_prev : p_ni_p_prev)) {
            // pressure of the fixed interface in its own column
            const double pe = rootm ? SM(1, F_P + is_p - 1) : SM(0, F_P + is_m - 1);
            if (rootm) { p_ni_m_prev = p_ni; PNP_SET(is_p, ks_p, p_ni); }
            else { p_ni_p_prev = p_ni; PNM_SET(is_m, ks_m, p_ni); }
            const double pa = rootm ? pe : p_ni, pb = rootm ? p_ni : pe;   // (plus side) - (minus side)
            emit_slope(-cnslp * (pa - pb), .5 * (pa + pb));
          }
        } else if (drho_zero) {
          const double pm = SM(0, F_P + is_m - 1), pp = SM(1, F_P + is_p - 1);
          PNP_SET(is_p, ks_p, pm);
          PNM_SET(is_m, ks_m, pp);
          emit_slope(-cnslp * (pp - pm), .5 * (pp + pm));
        }
      }
      // advance the minus column (drho >= 0), then the plus column (drho <= 0)
      int side = (drho_zero || drho_pos) ? 0 : 1;
      bool then_p = drho_zero;
      for (;;) {
        const double drho_prev = drho_curr;
        int is = side ? is_p : is_m, ks = side ? ks_p : ks_m;
        if (is == 1) is = 2;
        else {
          ks = ks + 1;
          if (ks > (side ? ksmx_p : ksmx_m)) { if (side) ks_p = ks; else ks_m = ks; done1 = true; break; }
          is = 1;
          stage(side, ks, false);
        }
        const Rec r = srec(side, is);
        if (side) { rp = r; is_p = is; ks_p = ks; } else { rm = r; is_m = is; ks_m = ks; }
        drho_curr = drho_at();
        if ((side ? drho_curr - drho_prev : drho_prev - drho_curr) > rho_eps) {
          if (is == 2 && SM(side, F_P + 1) - SM(side, F_P) > onemm) {
            if (side) stab_p |= 1ull << (ks - 1); else stab_m |= 1ull << (ks - 1);
          }
          if (then_p) { then_p = false; side = 1; continue; }
          break;
        }
        if (is == 1 && hget(side, 2 * ks - 1)) {   // PN(1,ks) = PN(2,ks-1); both are mval when the bit is clear
          double* pn = side ? pnp : pnm;
          pn[2 * ks] = pn[2 * ks - 1];
          hset(side, 2 * ks);
        }
      }
    }
  }

  if (A.surface_align) {  // :408-479
    int issa_m = 1;
    while (kssa_m <= ksmx_m) {
      if (hget(0, 2 * kssa_m + issa_m - 1)) break;
      if (issa_m == 1) issa_m = 2;
      else { kssa_m = kssa_m + 1; issa_m = 1; }
    }
    int issa_p = 1;
    while (kssa_p <= ksmx_p) {
      if (hget(1, 2 * kssa_p + issa_p - 1)) break;
      if (issa_p == 1) issa_p = 2;
      else { kssa_p = kssa_p + 1; issa_p = 1; }
    }
    if (kssa_m > ksmx_m || kssa_p > ksmx_p) {
      const double pbm = psd(0, 2, ksmx_m), pbp = psd(1, 2, ksmx_p);
      PNM_SET(1, 1, psd(0, 1, 1));
      for (ks_m = 1; ks_m <= ksmx_m - 1; ++ks_m) {
        if (psd(0, 1, ks_m) > pbp) break;
        const double p_ni = fmin(psd(0, 2, ks_m), pbp);
        PNM_SET(1, ks_m + 1, p_ni);
        PNM_SET(2, ks_m, p_ni);
        stab_m |= 1ull << (ks_m - 1);
      }
      PNP_SET(1, 1, psd(1, 1, 1));
      for (ks_p = 1; ks_p <= ksmx_p - 1; ++ks_p) {
        if (psd(1, 1, ks_p) > pbm) break;
        const double p_ni = fmin(psd(1, 2, ks_p), pbm);
        PNP_SET(1, ks_p + 1, p_ni);
        PNP_SET(2, ks_p, p_ni);
        stab_p |= 1ull << (ks_p - 1);
      }
    } else {
      double p1_m, p2_m, p1_p, p2_p;
      if (psd(0, issa_m, kssa_m) < PNP(issa_p, kssa_p)) {
        p1_m = psd(0, 1, 1); p2_m = psd(0, issa_m, kssa_m);
        p1_p = psd(1, 1, 1); p2_p = PNM(issa_m, kssa_m);
      } else {
        p1_m = psd(0, 1, 1); p2_m = PNP(issa_p, kssa_p);
        p1_p = psd(1, 1, 1); p2_p = psd(1, issa_p, kssa_p);
      }
      PNM_SET(1, 1, p1_p);
      for (ks_m = 1; ks_m <= kssa_m - 1; ++ks_m) {
        const double pl = psd(0, 2, ks_m);
        const double p_ni = ((pl - p1_m) * p2_p + (p2_m - pl) * p1_p) / (p2_m - p1_m);
        PNM_SET(1, ks_m + 1, p_ni);
        PNM_SET(2, ks_m, p_ni);
        stab_m |= 1ull << (ks_m - 1);
      }
      PNP_SET(1, 1, p1_m);
      for (ks_p = 1; ks_p <= kssa_p - 1; ++ks_p) {
        const double pl = psd(1, 2, ks_p);
        const double p_ni = ((pl - p1_p) * p2_m + (p2_p - pl) * p1_m) / (p2_p - p1_p);
        PNP_SET(1, ks_p + 1, p_ni);
        PNP_SET(2, ks_p, p_ni);
        stab_p |= 1ull << (ks_p - 1);
      }
    }
  }

  // ---- second search: neutral layers and their fluxes (:525-911)
  // The reference keeps the previous/current neutral interface in two slots that swap (nip/nic); here they
  // are plain "prev"/"cur" registers and cur is copied to prev when an interface has been found (a slot
  // index would put them in local memory).  Whenever a column lands on a new source layer that layer's record
  // is staged, so the polynomial coefficients, interface values, diffusivity and layer means of the current
  // layers are shared-memory reads (peval and pmeval are evaluated for every neutral interface found inside a
  // layer).  The branches of the case analysis only decide HOW the interface values are obtained (ev_m/ev_p);
  // the evaluation itself and the flux computation run after the branches have reconverged.
  SideAcc<NT, IX> accm, accp;
  accm.init(A.cvm, x, lev, kk, T);
  accp.init(A.cvp, x, lev, kk, T);
  {
    constexpr int NTC = NT > 0 ? NT : NTMAX;
    // scalar nt of the current layer of a column: T and S from the staged record, passive tracers from the
    // column record itself (their 8 doubles follow the staged part)
    struct CoefRef {   // the five polynomial coefficients, as the polynomial helpers read them
      const double (*col)[ND_BS]; const double* g5; int t;
      __device__ __forceinline__ double operator[](int c5) const { return g5 ? g5[c5] : col[c5][t]; }
    };
    auto lrec = [&](int side) { return side ? lrec_p : lrec_m; };
    auto xrec = [&](int side, int nt) {   // passive tracer nt >= 3 in the current layer's record
      return (side ? col_p + (ks_p - 1) * LS : col_m + (ks_m - 1) * LS) + ND_RSB + 8 * (nt - 3);
    };
    auto cf = [&](int side, int nt) {
      return nt > 2    ? CoefRef{nullptr, xrec(side, nt) + X_TPC, tx}
             : (STG & 1) ? CoefRef{&sm[(STG & 1) ? side * ND_RSS + F_TPC + (nt - 1) * 5 : 0], nullptr, tx}
                        : CoefRef{nullptr, lrec(side) + F_TPC + (nt - 1) * 5, tx};
    };
    auto tsrcdi = [&](int side, int is, int nt) {   // t_srcdi(is, ks, nt)
      return nt <= 2 ? SM(side, F_REC + 4 * (is - 1) + 1 + nt) : xrec(side, nt)[X_TSD + is - 1];
    };
    auto lmean = [&](int side, int nt) { return nt <= 2 ? SM(side, F_TLEV + nt - 1) : xrec(side, nt)[X_TLEV]; };
    int kc_m = 0, kc_p = 0;                        // layers whose records are staged
    auto pe = [&](const CoefRef c, double xx) { return (((c[4] * xx + c[3]) * xx + c[2]) * xx + c[1]) * xx + c[0]; };
    auto pme = [&](const CoefRef c, double x0, double x1) {
      const double c1_2 = 1. / 2., c1_3 = 1. / 3., c1_4 = 1. / 4., c1_5 = 1. / 5.;
      const double b5 = c1_5 * c[4];
      const double b4 = b5 * x1 + c1_4 * c[3];
      const double b3 = b4 * x1 + c1_3 * c[2];
      const double b2 = b3 * x1 + c1_2 * c[1];
      const double b1 = b2 * x1 + c[0];
      return (((b5 * x0 + b4) * x0 + b3) * x0 + b2) * x0 + b1;
    };

    is_m = 2; ks_m = 0; is_p = 2; ks_p = 0;
    int kd_m = 0, kd_p = 0, ks_m_prev = 0, ks_p_prev = 0;
    bool advance_src_m = true, advance_src_p = true, advance_dst_m = true, advance_dst_p = true;
    double p_prev_m = -mval, p_prev_p = -mval, p_cur_m = 0., p_cur_p = 0.;
    double x_prev_m = 0., x_prev_p = 0., x_cur_m = 0., x_cur_p = 0.;
    double t_prev_m[NTC], t_prev_p[NTC], t_cur_m[NTC], t_cur_p[NTC];
#pragma unroll
    for (int q = 0; q < NTC; ++q) { t_prev_m[q] = 0.; t_prev_p[q] = 0.; t_cur_m[q] = 0.; t_cur_p[q] = 0.; }
    // Values that only change when a column's source interface or destination layer moves are loaded at that
    // moment and kept: the interface pressures of the current source layer (ps1, ps2), pressure and neutral
    // partner of the next interface with a partner (ps_n, pn_n), the partner of the current interface (pn_c) and
    // the snapped lower interface of the current destination layer (snp).  An iteration of the search moves one
    // of the four pointers; re-reading all of these at its top cost ten loads where one to five are needed.
    double psm1 = 0., psm2 = 0., psp1 = 0., psp2 = 0., psm_n = 0., psp_n = 0., pnm_n = 0., pnp_n = 0.;
    double pnm_c = 0., pnp_c = 0.;
    // Destination layer kd of a column: {p_dst, p_dstsnp} of its upper (pdu, snu) and lower (pdl, snp) interface.
    // The layers are visited in order, so stepping to kd + 1 shifts lower to upper and loads ONE 16-byte pair;
    // for kd = 0 the "lower" interface is interface 1.  The upper pair lives in shared memory (smd).
    double snp_m, snp_p, pdl_m, pdl_p;
    { const double2 d0 = *reinterpret_cast<const double2*>(dst_m); pdl_m = d0.x; snp_m = d0.y; }
    { const double2 d0 = *reinterpret_cast<const double2*>(dst_p); pdl_p = d0.x; snp_p = d0.y; }
    auto step_dst_m = [&]() {   // kd_m has just been incremented (kd_m <= kdmx_m)
      const double2 d = reinterpret_cast<const double2*>(dst_m)[kd_m];
      smd[0][tx] = pdl_m; smd[1][tx] = snp_m; pdl_m = d.x; snp_m = d.y;
    };
    auto step_dst_p = [&]() {
      const double2 d = reinterpret_cast<const double2*>(dst_p)[kd_p];
      smd[2][tx] = pdl_p; smd[3][tx] = snp_p; pdl_p = d.x; snp_p = d.y;
    };
    int kuv = 1;
    auto puv = [&](int k) { return A.puv[x + (IX)(k - 1) * lev]; };
    // The face fluxes of a neutral sublayer are binned on the face's layers (:870-905).  A layer collects
    // the contributions of several sublayers one after the other; its four running sums (tflld, sflld, tflx,
    // sflx of layer kuv_acc) are kept in registers from the first contribution until the binning moves on
    // to the next layer: the additions and their order are the reference's, each array element is read
    // once and written once instead of once per contribution.
    int kuv_acc = 0;
    double a_tflld = 0., a_sflld = 0., a_tflx = 0., a_sflx = 0., pk_c = 0., pk1_c = 0.;
    auto flush_layer = [&]() {
      if (kuv_acc == 0) return;
      const IX o = x + (IX)(kuv_acc + mm - 1) * lev;
      A.tflld[o] = a_tflld; A.sflld[o] = a_sflld; A.tflx[o] = a_tflx; A.sflx[o] = a_sflx;
    };

    for (;;) {
      // advance to the next source interface of the minus and/or the plus column (mirrored blocks of the
      // reference, written once; lanes that advance different columns run together)
      if (advance_src_m || advance_src_p) {
        int side = advance_src_m ? 0 : 1;
        bool then_p = advance_src_m && advance_src_p, out = false;
        for (;;) {
          int is = side ? is_p : is_m, ks = side ? ks_p : ks_m;
          const int kmx = side ? ksmx_p : ksmx_m;
          const unsigned long long stab = side ? stab_p : stab_m;
          const double* pn = side ? pnp : pnm;
          for (;;) {
            if (is == 1) {
              is = 2;
              if (ks >= 1 && ((stab >> (ks - 1)) & 1ull)) break;
            } else {
              ks = ks + 1;
              if (ks > kmx) { out = true; break; }
              is = 1;
              if (((stab >> (ks - 1)) & 1ull) && hget(side, 2 * ks + is - 1)) break;
            }
          }
          if (out) break;
          // the next interface at or below (is,ks) that has a partner, (2,kmx) at the latest: the reference steps
          // through (is,ks) = q -> q + 1 while PN == mval, i.e. it looks for the first set bit at or above q
          int qn = 2 * ks + is - 1;
          {
            const unsigned long long w0 = smk[side * 2][tx], w1 = smk[side * 2 + 1][tx];
            int qf = 128;
            if (qn < 64) {
              const unsigned long long lo = w0 >> qn;
              if (lo != 0ull) qf = qn + __ffsll((long long)lo) - 1;
              else if (w1 != 0ull) qf = 64 + __ffsll((long long)w1) - 1;
            } else {
              const unsigned long long hi = w1 >> (qn - 64);
              if (hi != 0ull) qf = qn + __ffsll((long long)hi) - 1;
            }
            qn = min(qf, 2 * kmx + 1);
          }
          const int isn = (qn & 1) + 1, ksn = qn >> 1;
          {
            if (ks != (side ? kc_p : kc_m)) { stage(side, ks, true); if (side) kc_p = ks; else kc_m = ks; }
            const double ps1 = SM(side, F_P), ps2 = SM(side, F_P + 1);
            const double ps_n = ksn == ks ? (isn == 1 ? ps1 : ps2) : psd(side, isn, ksn);
            const double pn_n = hget(side, qn) ? pn[qn] : mval;
            const double pn_c = hget(side, 2 * ks + is - 1) ? pn[2 * ks + is - 1] : mval;
            if (side) { is_p = is; ks_p = ks; psp1 = ps1; psp2 = ps2; psp_n = ps_n; pnp_n = pn_n; pnp_c = pn_c; }
            else { is_m = is; ks_m = ks; psm1 = ps1; psm2 = ps2; psm_n = ps_n; pnm_n = pn_n; pnm_c = pn_c; }
          }
          if (then_p) { then_p = false; side = 1; continue; }
          break;
        }
        if (out) break;
      }
      // the quantities every branch below looks at
      if (p_prev_m == -mval) {
        if ((pnm_n - psp_n) < (pnp_n - psm_n)) {
          p_prev_m = psm_n;
          p_prev_p = pnm_n;
        } else {
          p_prev_m = pnp_n;
          p_prev_p = psp_n;
        }
      }
      if (advance_dst_m) {
        kd_m = kd_m + 1;
        if (kd_m > kdmx_m) break;
        step_dst_m();
      }
      if (advance_dst_p) {
        kd_p = kd_p + 1;
        if (kd_p > kdmx_p) break;
        step_dst_p();
      }
      {
        bool out = false;
        const double lim_m = fmax(psm1, p_prev_m);
        while (snp_m <= lim_m) {
          kd_m = kd_m + 1;
          if (kd_m > kdmx_m) { out = true; break; }
          step_dst_m();
        }
        if (out) break;
        const double lim_p = fmax(psp1, p_prev_p);
        while (snp_p <= lim_p) {
          kd_p = kd_p + 1;
          if (kd_p > kdmx_p) { out = true; break; }
          step_dst_p();
        }
        if (out) break;
      }
      advance_src_m = false; advance_src_p = false; advance_dst_m = false; advance_dst_p = false;

      const double psm = is_m == 1 ? psm1 : psm2, psp = is_p == 1 ? psp1 : psp2;
      int case_m = 3;
      if (psm <= pnp_n) {
        if (psm <= snp_m) case_m = 1;
      } else if (pnp_n <= snp_m) {
        case_m = 2;
      }
      int case_p = 3;
      if (psp <= pnm_n) {
        if (psp <= snp_p) case_p = 1;
      } else if (pnm_n <= snp_p) {
        case_p = 2;
      }
      bool found_ni = false;
      // how the scalar values at the new interface are obtained: 1 = polynomial at x_cur, 2 = stored
      // interface value t_srcdi(is,ks)
      int ev_m = 0, ev_p = 0;

      // The reference spells the case analysis out once per column; the two columns' branches are mirror
      // images, so they are written once with the roles chosen per lane (a = the column that decides,
      // b = the other one) and lanes on mirrored branches stay together.  The local coordinates x of the
      // interface are evaluated after the branches from p_cur (same expression as in every branch).
      if (case_m == 3 && case_p == 3) {
        if (is_p == 2 && is_m == 2) {
          p_cur_m = snp_m;
          p_cur_p = snp_p;
          const double pu_m = p_prev_m, pu_p = p_prev_p;
          double pl_m, pl_p;
          if ((pnm_n - psp_n) < (pnp_n - psm_n)) {
            pl_m = psm_n;
            pl_p = pnm_n;
          } else {
            pl_m = pnp_n;
            pl_p = psp_n;
          }
          const double pp1 = (p_cur_m - pu_m) * (pl_p - pu_p);
          const double pp2 = (p_cur_p - pu_p) * (pl_m - pu_m);
          if (fabs(pp1 - pp2) < dp_eps * fmax(dp_eps, pl_m - pu_m + pl_p - pu_p)) {
            advance_dst_m = true;
            advance_dst_p = true;
          } else if (pp1 < pp2) {
            p_cur_p = pu_p + pp1 / (pl_m - pu_m);
            advance_dst_m = true;
          } else {
            p_cur_m = pu_m + pp2 / (pl_p - pu_p);
            advance_dst_p = true;
          }
          if (p_cur_m >= psm1 && p_cur_m <= psm2 && p_cur_p >= psp1 && p_cur_p <= psp2) {
            ev_m = 1; ev_p = 1;
            found_ni = true;
          }
        } else {
          if (is_p != 2) advance_dst_m = true;
          if (is_m != 2) advance_dst_p = true;
        }
      } else if (case_m == 3 || case_p == 3) {
        // exactly one column (a) meets its next destination interface first (:640-700)
        const bool am = case_m == 3;
        const int is_b = am ? is_p : is_m, case_b = am ? case_p : case_m;
        if (is_b == 2) {
          const double snp_a = am ? snp_m : snp_p;
          const double pa_prev = am ? p_prev_m : p_prev_p, pb_prev = am ? p_prev_p : p_prev_m;
          const double px = case_b == 1 ? (am ? psp_n : psm_n) : (am ? pnm_n : pnp_n);
          const double py = case_b == 1 ? (am ? pnp_n : pnm_n) : (am ? psm_n : psp_n);
          const double pb_cur = pb_prev + (snp_a - pa_prev) * (px - pb_prev) / (py - pa_prev);
          if (am) { p_cur_m = snp_a; p_cur_p = pb_cur; } else { p_cur_p = snp_a; p_cur_m = pb_cur; }
          if (pb_cur >= (am ? psp1 : psm1) && pb_cur <= (am ? psp2 : psm2)) {
            ev_m = 1; ev_p = 1;
            found_ni = true;
            if (am) advance_dst_m = true; else advance_dst_p = true;
          } else {
            const double pn_b = am ? pnp_c : pnm_c;
            if (case_b == 1 && pn_b == mval) { if (am) advance_src_p = true; else advance_src_m = true; }
            else { if (am) advance_dst_m = true; else advance_dst_p = true; }
          }
        } else {
          if (am) advance_dst_m = true; else advance_dst_p = true;
        }
      } else if (case_m == 1 && case_p == 1) {
        if (pnm_c != mval && pnp_c != mval) {
          p_cur_m = psm;
          p_cur_p = psp;
          ev_m = 2; ev_p = 2;
          found_ni = true;
          advance_src_m = true;
          advance_src_p = true;
        } else {
          if (pnm_c == mval) advance_src_m = true;
          if (pnp_c == mval) advance_src_p = true;
        }
      } else if (case_m == 1 || case_p == 1) {
        // one column (a) sits on a source interface whose neutral partner lies inside the other's layer (:745-790)
        const bool am = case_m == 1;
        const double pn_c = am ? pnm_c : pnp_c;
        if (pn_c != mval && pn_c >= (am ? psp1 : psm1)) {
          if (am) { p_cur_m = psm; p_cur_p = pn_c; ev_m = 2; ev_p = 1; }
          else { p_cur_p = psp; p_cur_m = pn_c; ev_p = 2; ev_m = 1; }
          found_ni = true;
        }
        if (am) advance_src_m = true; else advance_src_p = true;
      } else {
        advance_src_m = true;
        advance_src_p = true;
      }

      if (found_ni) {  // :795-907
        // NOTE: advance_src_* set above only act at the top of the next iteration: is/ks are still the
        // ones the interface was found for
        {
          const double xm_ = (p_cur_m - psm1) / (psm2 - psm1), xp_ = (p_cur_p - psp1) / (psp2 - psp1);
          x_cur_m = ev_m == 2 ? (double)(is_m - 1) : xm_;
          x_cur_p = ev_p == 2 ? (double)(is_p - 1) : xp_;
        }
#pragma unroll
        for (int nt = 1; nt <= NTC; ++nt)
          if (nt <= T) {
            t_cur_m[nt - 1] = ev_m == 2 ? tsrcdi(0, is_m, nt) : pe(cf(0, nt), x_cur_m);
            t_cur_p[nt - 1] = ev_p == 2 ? tsrcdi(1, is_p, nt) : pe(cf(1, nt), x_cur_p);
          }
        const double dp_ni_m = fmin(p_cur_m - p_prev_m, pdl_m - smd[0][tx]);
        const double dp_ni_p = fmin(p_cur_p - p_prev_p, pdl_p - smd[2][tx]);
        const double dp_ni = 2. * dp_ni_m * dp_ni_p / fmax(dp_ni_m + dp_ni_p, 2. * dp_eps);
        if (ks_m == ks_m_prev && ks_p == ks_p_prev && p_prev_m >= smd[1][tx] && p_cur_m <= snp_m &&
            p_prev_p >= smd[3][tx] && p_cur_p <= snp_p && dp_ni > 2. * dp_eps) {
          accm.advance(kd_m);
          accp.advance(kd_p);
          const double q = .5 * cdiff * (SM(0, F_DIF) + SM(1, F_DIF)) * dp_ni;
          double tflx = 0., sflx = 0.;
          bool ts_ok = true;
#pragma unroll
          for (int nt = 1; nt <= NTC; ++nt)
            if (nt <= T) {
              const double d = pme(cf(0, nt), x_prev_m, x_cur_m) - pme(cf(1, nt), x_prev_p, x_cur_p);
              const double cm = lmean(0, nt), cp = lmean(1, nt);
              const bool ok = d * (cm - cp) >= 0. && d * (t_prev_m[nt - 1] - t_prev_p[nt - 1]) >= 0. &&
                              d * (t_cur_m[nt - 1] - t_cur_p[nt - 1]) >= 0.;
              if (nt == IT) { tflx = q * d; ts_ok = ok; }
              else if (nt == IS) { sflx = q * d; ts_ok = ts_ok && ok; }
              else if (ok) {
                const double f = q * d;
                accm.a[nt - 1] += f;
                accp.a[nt - 1] -= f;
              }
            }
          if (ts_ok) {
            accm.a[IT - 1] += tflx; accp.a[IT - 1] -= tflx;
            accm.a[IS - 1] += sflx; accp.a[IS - 1] -= sflx;
            const double p_ni_up = .5 * (p_prev_m + p_prev_p);
            const double p_ni_lo = .5 * (p_cur_m + p_cur_p);
            const double dp_ni_i = 1. / fmax(epsilp, p_ni_lo - p_ni_up);
            while (kuv <= kk) {
              if (kuv_acc != kuv) {   // bring layer kuv's four sums into registers (the previous layer's go out)
                flush_layer();
                const IX o = x + (IX)(kuv + mm - 1) * lev;
                a_tflld = A.tflld[o]; a_sflld = A.sflld[o]; a_tflx = A.tflx[o]; a_sflx = A.sflx[o];
                pk_c = puv(kuv); pk1_c = puv(kuv + 1);
                kuv_acc = kuv;
              }
              const double pk = pk_c, pk1 = pk1_c;
              const bool below = pk1 < p_ni_lo;
              const double mlfrac = below ? fmax(0., pk1 - fmax(p_ni_up, pk)) * dp_ni_i
                                          : (p_ni_lo - fmax(p_ni_up, pk)) * dp_ni_i;
              a_tflld = a_tflld + tflx * mlfrac;
              a_sflld = a_sflld + sflx * mlfrac;
              a_tflx = a_tflx + tflx * mlfrac;
              a_sflx = a_sflx + sflx * mlfrac;
              if (!below) break;
              kuv = kuv + 1;
            }
          }
        }
        ks_m_prev = ks_m;
        ks_p_prev = ks_p;
        p_prev_m = p_cur_m; p_prev_p = p_cur_p; x_prev_m = x_cur_m; x_prev_p = x_cur_p;
#pragma unroll
        for (int q2 = 0; q2 < NTC; ++q2) { t_prev_m[q2] = t_cur_m[q2]; t_prev_p[q2] = t_cur_p[q2]; }
      }
    }
    flush_layer();
  }
  accm.finish();
  accp.finish();

  // ---- neutral slope at the destination interfaces below the last pair (:913-951, tail of emit_slope)
  for (; kd_sl <= kk; ++kd_sl) A.nslp[x + (IX)(kd_sl - 1) * lev] = nns == 0 ? 0. : s_prev;
#undef PNM
#undef PNP
#undef SM
}

// ndiff_update_trc_jslice (:1152-1175) with the gather of the four face contributions
__global__ void __launch_bounds__(256)
ndiff_update(Geom g, int T, const int* __restrict__ ip, const int* __restrict__ iu, const int* __restrict__ iv,
             const double* __restrict__ scp2, const double* __restrict__ p_dst, const double* __restrict__ ucm,
             const double* __restrict__ ucp, const double* __restrict__ vcm, const double* __restrict__ vcp,
             double* __restrict__ trc_rm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1, k = blockIdx.z + 1;
  if (i > g.ii) return;
  const long x = ix2(g, i, j), lev = g.lev;
  if (ip[x] != 1) return;
  const int kk = g.kdm;
  const double q = 1. / (scp2[x] * fmax(p_dst[x + (long)k * lev] - p_dst[x + (long)(k - 1) * lev], dp_eps));
  const bool ws = iv[x] == 1, ww = iu[x] == 1, we = iu[x + 1] == 1, wn = iv[x + g.ldi] == 1;
  for (int nt = 0; nt < T; ++nt) {
    const long o = (long)(nt * kk + k - 1) * lev;
    double conv = 0.;
    if (ws) conv += vcp[x + o];
    if (ww) conv += ucp[x + o];
    if (we) conv += ucm[x + 1 + o];
    if (wn) conv += vcm[x + g.ldi + o];
    trc_rm[x + o] = trc_rm[x + o] - q * conv;
  }
}

}  // namespace

#ifndef BLOM_HOST_EMUL
// neutral diffusion over the whole tile in the order of the reference's slice pipeline
// (phy/mod_ale_regrid_remap.F90:1607-1690)
void ndiff_dev(int m, int n, int mm, int nn, int k1m, int k1n) {
  (void)m; (void)n; (void)k1m; (void)k1n;
  Ctx& c = C(); const Geom& g = c.g;
  const int kk = g.kdm, T = 2 + g.ntr;
  if (kk >= KMN) throw std::runtime_error("ndiff: kdm exceeds the compiled column bound (63)");
  if (T > NTMAX) throw std::runtime_error("ndiff: more than 8 diffused scalars are not compiled in");
  const bool surface_align = c.option("ndiff_surface_align", "1") == "1";   // namelist default .true.
  if (surface_align) halo_update(c.dev("dpml"), 1, 1, 1, halo_ps);
  const int RS = nd_rs(T);
  // column records: kk records of RS doubles per cell, whole cell blocks
  double* src = c.owned("_nd_src", cdiv((long)cdiv(g.lev, ND_CB) * ND_CB * kk * RS, g.lev));
  double* dst = c.owned("_nd_dst", 2 * (kk + 1));     // {p_dst, p_dstsnp} per destination interface and cell
  int* kdmx = c.owned_int("_nd_kdmx", 1);
  double* ucm = c.owned("_nd_ucm", kk * T);
  double* ucp = c.owned("_nd_ucp", kk * T);
  double* vcm = c.owned("_nd_vcm", kk * T);
  double* vcp = c.owned("_nd_vcp", kk * T);

  PrepIn I{};
  I.ip = c.idev("ip"); I.iu = c.idev("iu"); I.iv = c.idev("iv"); I.ksmx = c.idev("nd_ksmx");
  I.p_src = c.dev("nd_p_src"); I.tsd = c.dev("nd_t_srcdi"); I.tpc = c.dev("nd_tpc_src"); I.p_dst = c.dev("nd_p_dst");
  I.difiso = c.dev("difiso");
  I.tlev[0] = c.dev("temp") + (long)nn * g.lev;
  I.tlev[1] = c.dev("saln") + (long)nn * g.lev;
  for (int nt = 3; nt <= T; ++nt) I.tlev[nt - 1] = c.dev("trc") + (long)(nn + (nt - 3) * 2 * kk) * g.lev;
  LAUNCH(ndiff_prep, dim3(cdiv(g.ii + 2, 128), g.jj + 2), 128, 0, g, mm, T, I, kdmx, src, dst,
         c.dev("utflld"), c.dev("usflld"), c.dev("vtflld"), c.dev("vsflld"), ucm, ucp, vcm, vcp);

  NdArgs A{};
  A.src = src; A.dst = dst;
  A.ksmx = c.idev("nd_ksmx"); A.kdmx = kdmx;
  A.dpml = c.dev("dpml");
  A.delt1 = c.scalar("delt1"); A.mm = mm; A.T = T; A.surface_align = surface_align ? 1 : 0;

  NdArgs U = A;
  U.mask = c.idev("iu"); U.sca = c.dev("scuy"); U.scbi = c.dev("scuxi"); U.puv = c.dev("pu");
  U.tflld = c.dev("utflld"); U.sflld = c.dev("usflld"); U.tflx = c.dev("utflx"); U.sflx = c.dev("usflx");
  U.nslp = c.dev("nslpx"); U.cvm = ucm; U.cvp = ucp;
  NdArgs V = A;
  V.mask = c.idev("iv"); V.sca = c.dev("scvx"); V.scbi = c.dev("scvyi"); V.puv = c.dev("pv");
  V.tflld = c.dev("vtflld"); V.sflld = c.dev("vsflld"); V.tflx = c.dev("vtflx"); V.sflx = c.dev("vsflx");
  V.nslp = c.dev("nslpy"); V.cvm = vcm; V.cvp = vcp;

  // compacted lists of the wet u faces (1..ii+1 x 1..jj) and v faces (1..ii x 1..jj+1); the masks are static
  // after bigrid, so the lists are built once per tile
  int* list_u = c.owned_int("_nd_faces_u", 1);
  int* list_v = c.owned_int("_nd_faces_v", 1);
  if (!c.sc.count("_nd_nfaces_u")) {   // (bigrid erases the counts when the masks are rebuilt)
    std::vector<int> hu(g.lev), hv(g.lev), lu, lv;
    CUDA_CHECK(cudaMemcpyAsync(hu.data(), c.idev("iu"), sizeof(int) * g.lev, cudaMemcpyDeviceToHost, c.stream));
    CUDA_CHECK(cudaMemcpyAsync(hv.data(), c.idev("iv"), sizeof(int) * g.lev, cudaMemcpyDeviceToHost, c.stream));
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
    for (int j = 1; j <= g.jj; ++j)
      for (int i = 1; i <= g.ii + 1; ++i) {
        const long x = ix2(g, i, j);
        if (hu[x] == 1) lu.push_back((int)x);
      }
    // v faces in tiles of 32 (i) x 4 (j): a 128-thread block then owns four consecutive rows of a 32-wide strip, and
    // the cell column (i,j) that is the plus side of face (i,j) and the minus side of face (i,j+1) is fetched by one
    // block instead of by two blocks a whole grid row apart (ncu, row-major order: 75 GB of DRAM traffic per launch
    // for the v faces against 49 GB for the u faces, whose two columns sit in neighbouring lanes)
    // (walking the tiles in patches of 8 x 4, which puts the two faces of a shared column into one warp, measured equal)
    const int pw = 32;
    for (int j0 = 1; j0 <= g.jj + 1; j0 += 4)
      for (int i0 = 1; i0 <= g.ii; i0 += pw)
        for (int j = j0; j < std::min(j0 + 4, g.jj + 2); ++j)
          for (int i = i0; i < std::min(i0 + pw, g.ii + 1); ++i) {
            const long x = ix2(g, i, j);
            if (hv[x] == 1) lv.push_back((int)x);
          }
    if (!lu.empty()) CUDA_CHECK(cudaMemcpyAsync(list_u, lu.data(), sizeof(int) * lu.size(), cudaMemcpyHostToDevice, c.stream));
    if (!lv.empty()) CUDA_CHECK(cudaMemcpyAsync(list_v, lv.data(), sizeof(int) * lv.size(), cudaMemcpyHostToDevice, c.stream));
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
    c.sc["_nd_nfaces_u"] = (double)lu.size();
    c.sc["_nd_nfaces_v"] = (double)lv.size();
  }
  U.faces = list_u; U.nfaces = (int)c.sc["_nd_nfaces_u"];
  V.faces = list_v; V.nfaces = (int)c.sc["_nd_nfaces_v"];
  // 32-bit index arithmetic for the level-strided output arrays of ndiff_face when every element index fits (the
  // face buffers are the largest: kk*T levels); the 64-bit instantiation covers the rest
  const bool ix32 = ((long)kk * std::max(T, 2) + 2) * g.lev < (1l << 32);
  const dim3 gu(std::max(1, cdiv(U.nfaces, ND_BS))), gv(std::max(1, cdiv(V.nfaces, ND_BS)));
  const int stg = std::stoi(c.option("ndiff_stage", "3"));
  if (stg < 0 || stg > 4) throw std::runtime_error("ndiff: ndiff_stage must be 0 .. 4");
#define ND_LAUNCH(NT_, IX_, STG_)                                                                  \
  do {                                                                                             \
    LAUNCH_NAMED("ndiff_face<u>", (ndiff_face<0, NT_, IX_, STG_>), gu, ND_BS, 0, g, U);            \
    LAUNCH_NAMED("ndiff_face<v>", (ndiff_face<1, NT_, IX_, STG_>), gv, ND_BS, 0, g, V);            \
  } while (0)
#define ND_FACE(NT_)                                                                               \
  do {                                                                                             \
    if (!ix32) ND_LAUNCH(NT_, long, 3);                                                            \
    else if (stg == 0) ND_LAUNCH(NT_, unsigned, 0);                                                \
    else if (stg == 2) ND_LAUNCH(NT_, unsigned, 2);                                                \
    else if (stg == 4) ND_LAUNCH(NT_, unsigned, 4);                                                \
    else if (stg == 1) ND_LAUNCH(NT_, unsigned, 1);                                                \
    else ND_LAUNCH(NT_, unsigned, 3);                                                              \
  } while (0)
  if (T == 2) { ND_FACE(2); }
  else if (T == 3) { ND_FACE(3); }
  else { ND_FACE(0); }
#undef ND_FACE
#undef ND_LAUNCH
  LAUNCH(ndiff_update, dim3(cdiv(g.ii, 256), g.jj, kk), 256, 0, g, T, c.idev("ip"), c.idev("iu"), c.idev("iv"),
         c.dev("scp2"), c.dev("nd_p_dst"), ucm, ucp, vcm, vcp, c.dev("nd_trc_rm"));
}
#endif  // BLOM_HOST_EMUL

}  // namespace blom
