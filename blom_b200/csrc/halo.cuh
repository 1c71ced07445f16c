// Device-side halo update (xctilr semantics, phy/mod_xc.F90:4222-4428) shared by the stand-alone
// halo kernel (xc.cu) and the persistent barotropic kernel (barotp.cu), which refreshes its
// subcycled fields between substeps without leaving the kernel.
#pragma once
#include "common.cuh"

namespace blom {

template <bool CGL>
__device__ __forceinline__ double hld(const double* p) { return CGL ? __ldcg(p) : *p; }

constexpr int HALO_MAX = 12;
struct HaloBatch {
  double* base[HALO_MAX];
  int nlev[HALO_MAX];
  int itype[HALO_MAX];
};

// Value of a(i,j) after the N/S phase of xctilr for 1<=i<=ii.  `tgt` tells
// whether (i,j) is rewritten by the N/S phase on this tile.
template <bool CGL = false>
__device__ __forceinline__ double ns_value(const Geom& g, const double* a, int itype, int i, int j,
                                           int nhl, bool& tgt) {
  const int ii = g.ii, jj = g.jj;
  const int it = itype % 10;
  tgt = false;
  if (g.nreg == 2) {
    if (j <= 0) {
      if (g.south) { tgt = true; return 0.0; }  // closed southern boundary
      return hld<CGL>(a + ix2(g, i, j));
    }
    if (g.north && j >= jj) {
      int io, jo = 0;
      if (it == 1 || it == 3) {  // p, u: rows jj+jh <- jj-1-jh, jh=0..nhl
        tgt = true;
        jo = jj - 1 - (j - jj);
        io = (it == 1) ? ii + 1 - i : (i == 1 ? 1 : ii + 2 - i);
      } else {  // q, v: right half of row jj; rows jj+jh <- jj-jh
        io = (it == 2) ? (i == 1 ? 1 : ii + 2 - i) : ii + 1 - i;
        if (j > jj) { tgt = true; jo = jj - (j - jj); }
        else if (i >= ii / 2 + 1) { tgt = true; jo = jj; }
      }
      if (tgt) {
        double v = hld<CGL>(a + ix2(g, io, jo));
        return itype < 10 ? v : -v;
      }
    }
    return hld<CGL>(a + ix2(g, i, j));
  }
  if (j <= 0) {
    if (g.south) {
      tgt = true;
      if (g.nreg <= 2) return 0.0;
      return hld<CGL>(a + ix2(g, i, jj + j));  // periodic in j, single band only
    }
    return hld<CGL>(a + ix2(g, i, j));
  }
  if (j > jj) {
    if (g.north) {
      tgt = true;
      if (g.nreg <= 2) return 0.0;
      return hld<CGL>(a + ix2(g, i, j - jj));
    }
    return hld<CGL>(a + ix2(g, i, j));
  }
  return hld<CGL>(a + ix2(g, i, j));
}

// One thread per halo target cell, level and request.
// ns_l1: first level (1-based) that takes part in the N/S phase;
// ew_l1: first level of the E/W phase (1 for the arctic serial code, else l1).
// Update level k (pointer a) of one field: thread `tid` of `nthr` cooperating threads.
template <bool CGL = false>
__device__ __forceinline__ void halo_level(const Geom& g, double* a, int itype, int k, int mhl, int nhl, int ns_l1,
                                           int ew_l1, long tid, long nthr) {
  const int ii = g.ii, jj = g.jj;
  const bool fold = (g.nreg == 2 && g.north);
  const int rows_n = nhl + (fold ? 1 : 0);
  const long n_ns = (long)(nhl + rows_n) * ii;
  const int ew_rows = jj + 2 * nhl;
  const long n_ew = (long)2 * mhl * ew_rows;
  for (long idx = tid; idx < n_ns + n_ew; idx += nthr) {
    if (idx < n_ns) {
      if (k < ns_l1) continue;
      const int rr = (int)(idx / ii), i = (int)(idx % ii) + 1;
      int j;
      if (rr < nhl) j = -rr;
      else j = fold ? jj + (rr - nhl) : jj + 1 + (rr - nhl);
      bool tgt;
      double v = ns_value<CGL>(g, a, itype, i, j, nhl, tgt);
      if (tgt) a[ix2(g, i, j)] = v;
    } else {
      if (k < ew_l1) continue;
      const long e = idx - n_ns;
      const int c = (int)(e / ew_rows), j = (int)(e % ew_rows) + 1 - nhl;
      int itg, is;
      if (c < mhl) { itg = -c; is = ii - c; }           // a(1-i') <- a(ii+1-i')
      else { itg = ii + (c - mhl) + 1; is = c - mhl + 1; }  // a(ii+i') <- a(i')
      double v;
      if (g.nreg == 0 || g.nreg == 4) v = 0.0;
      else if (k >= ns_l1 && (j <= 0 || j >= jj)) { bool tgt; v = ns_value<CGL>(g, a, itype, is, j, nhl, tgt); }
      else v = hld<CGL>(a + ix2(g, is, j));
      a[ix2(g, itg, j)] = v;
    }
  }
}

}  // namespace blom
