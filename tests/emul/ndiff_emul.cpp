// TEST INFRASTRUCTURE ONLY.  The neutral-diffusion kernels of the product (blom_b200/csrc/ndiff.cu: ndiff_prep,
// ndiff_face<u|v>, ndiff_update) compiled for the host through cuda_host_shim.hpp and run one emulated thread
// after the other, in the launch order of ndiff_dev.  tests/test_emul_ndiff.py holds the result against the
// oracle bit for bit, so the kernel source itself (not a restatement of it) is parity-checked in the CPU suite.
// Build: g++ -O2 -ffp-contract=off (the analogue of the product flavour's -fmad=false).
#include "cuda_host_shim.hpp"
#define BLOM_HOST_EMUL 1
#include "../../blom_b200/csrc/ndiff.cu"

#include <cstring>
#include <vector>

namespace blom {

struct EmuNdiff {
  // geometry
  int ii, jj, kdm, nb, ldi, ldj, ntr;
  // time-level arguments and options
  int mm, nn, surface_align, ix64;   // ix64: kernel variant, see faces_ix
  double delt1;
  // masks and inputs
  const int *ip, *iu, *iv, *ksmx;
  const double *p_src, *tsd, *tpc, *p_dst, *dpml, *difiso, *temp, *saln, *trc;
  const double *scuy, *scuxi, *scvx, *scvyi, *scp2, *pu, *pv;
  // outputs
  double *utflld, *usflld, *vtflld, *vsflld, *utflx, *usflx, *vtflx, *vsflx, *nslpx, *nslpy, *trc_rm;
};

template <int NT, class IX, int STG>
static void faces(const Geom& g, const NdArgs& U, const NdArgs& V) {
  emu_launch(dim3(std::max(1, cdiv(U.nfaces, ND_BS))), dim3(ND_BS), [&] { ndiff_face<0, NT, IX, STG>(g, U); });
  emu_launch(dim3(std::max(1, cdiv(V.nfaces, ND_BS))), dim3(ND_BS), [&] { ndiff_face<1, NT, IX, STG>(g, V); });
}
// variant: 0..4 = 32-bit index arithmetic with ndiff_stage = variant; 9 = the 64-bit instantiation (staged)
template <int NT>
static void faces_ix(int variant, const Geom& g, const NdArgs& U, const NdArgs& V) {
  if (variant == 9) faces<NT, long, 3>(g, U, V);
  else if (variant == 4) faces<NT, unsigned, 4>(g, U, V);
  else if (variant == 3) faces<NT, unsigned, 3>(g, U, V);
  else if (variant == 0) faces<NT, unsigned, 0>(g, U, V);
  else if (variant == 2) faces<NT, unsigned, 2>(g, U, V);
  else faces<NT, unsigned, 1>(g, U, V);
}

extern "C" int emu_ndiff(const EmuNdiff* e) {
  Geom g{};
  g.kdm = e->kdm; g.nb = e->nb; g.ntr = e->ntr; g.ii = e->ii; g.jj = e->jj;
  g.idm = e->ii; g.jdm = e->jj; g.ldi = e->ldi; g.ldj = e->ldj; g.lev = (long)e->ldi * e->ldj;
  const int kk = g.kdm, T = 2 + g.ntr;
  if (kk >= KMN || T > NTMAX) return 1;
  const size_t lev = (size_t)g.lev;
  std::vector<double> src((size_t)cdiv((long)lev, ND_CB) * ND_CB * kk * nd_rs(T), 0.), dst(lev * 2 * (kk + 1), 0.);
  // (poisoned: ndiff_prep has to zero what ndiff_update reads)
  const double nan = __builtin_nan("");
  std::vector<double> ucm(lev * kk * T, nan), ucp(lev * kk * T, nan), vcm(lev * kk * T, nan), vcp(lev * kk * T, nan);
  std::vector<int> kdmx(lev, 0);

  PrepIn I{};
  I.ip = e->ip; I.iu = e->iu; I.iv = e->iv; I.ksmx = e->ksmx;
  I.p_src = e->p_src; I.tsd = e->tsd; I.tpc = e->tpc; I.p_dst = e->p_dst; I.difiso = e->difiso;
  I.tlev[0] = e->temp + (long)e->nn * g.lev;
  I.tlev[1] = e->saln + (long)e->nn * g.lev;
  for (int nt = 3; nt <= T; ++nt) I.tlev[nt - 1] = e->trc + (long)(e->nn + (nt - 3) * 2 * kk) * g.lev;
  emu_launch(dim3(cdiv(g.ii + 2, 128), g.jj + 2), dim3(128), [&] {
    ndiff_prep(g, e->mm, T, I, kdmx.data(), src.data(), dst.data(), e->utflld, e->usflld, e->vtflld, e->vsflld,
               ucm.data(), ucp.data(), vcm.data(), vcp.data());
  });

  NdArgs A{};
  A.src = src.data(); A.dst = dst.data();
  A.ksmx = e->ksmx; A.kdmx = kdmx.data(); A.dpml = e->dpml;
  A.delt1 = e->delt1; A.mm = e->mm; A.T = T; A.surface_align = e->surface_align;
  NdArgs U = A, V = A;
  U.mask = e->iu; U.sca = e->scuy; U.scbi = e->scuxi; U.puv = e->pu;
  U.tflld = e->utflld; U.sflld = e->usflld; U.tflx = e->utflx; U.sflx = e->usflx; U.nslp = e->nslpx;
  U.cvm = ucm.data(); U.cvp = ucp.data();
  V.mask = e->iv; V.sca = e->scvx; V.scbi = e->scvyi; V.puv = e->pv;
  V.tflld = e->vtflld; V.sflld = e->vsflld; V.tflx = e->vtflx; V.sflx = e->vsflx; V.nslp = e->nslpy;
  V.cvm = vcm.data(); V.cvp = vcp.data();
  std::vector<int> lu, lv;   // wet faces (ranges of ndiff_dev)
  for (int j = 1; j <= g.jj; ++j)
    for (int i = 1; i <= g.ii + 1; ++i)
      if (e->iu[ix2(g, i, j)] == 1) lu.push_back((int)ix2(g, i, j));
  for (int j = 1; j <= g.jj + 1; ++j)
    for (int i = 1; i <= g.ii; ++i)
      if (e->iv[ix2(g, i, j)] == 1) lv.push_back((int)ix2(g, i, j));
  U.faces = lu.data(); U.nfaces = (int)lu.size();
  V.faces = lv.data(); V.nfaces = (int)lv.size();
  if (T == 2) faces_ix<2>(e->ix64, g, U, V);
  else if (T == 3) faces_ix<3>(e->ix64, g, U, V);
  else faces_ix<0>(e->ix64, g, U, V);

  emu_launch(dim3(cdiv(g.ii, 256), g.jj, kk), dim3(256), [&] {
    ndiff_update(g, T, e->ip, e->iu, e->iv, e->scp2, e->p_dst, ucm.data(), ucp.data(), vcm.data(), vcp.data(), e->trc_rm);
  });
  return 0;
}

}  // namespace blom
