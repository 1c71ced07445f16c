// ORACLE — TEST INFRASTRUCTURE ONLY (see core.hpp header).
// Flat C entry points so tests / bench.py can drive the restatement with ctypes.
#include "core.hpp"
#include "eos.hpp"
#include <cstdio>
#include <cstring>

namespace orc {
void init_cppm();
void cppm(int, int, int, int, int, int);
void advect(int, int, int, int, int, int);
const double* cppm_table(const char*, size_t*);
const int* cppm_stencil(const char*, size_t*);
void diffus(int, int, int, int, int, int);
void tmsmt1(int);
void tmsmt2(int, int, int, int);
void pgforc(int, int, int, int, int, int);
void inieos();
void momtum(int, int, int, int, int, int);
void barotp(int, int, int, int, int, int);
void eddtra(int, int, int, int, int, int);
void numerical_bounds();
void pbcor1(int, int, int, int, int, int);
void pbcor2(int, int, int, int, int, int);
void init_fluxes(int, int, int, int, int, int);
void ndiff(int, int, int, int, int, int);
void cmnfld2(int, int, int, int, int, int);
void cmnfld_bfsqf_ale(int, int, int, int, int, int);
void cmnfld_nslope_ale(int, int, int, int, int, int);
void cmnfld_nnslope_ale(int, int, int, int, int, int);
void budget_init(double*);
void budget_sums(int, int, int, double*);
}

static char g_err[1024] = "";

#define GUARD(stmt)                                        \
  try { stmt; return 0; }                                  \
  catch (const std::exception& e) {                        \
    std::snprintf(g_err, sizeof g_err, "%s", e.what());    \
    return 1;                                              \
  }

extern "C" {

const char* oracle_last_error() { return g_err; }

// dims: itdm,jtdm,kdm,idm,jdm,nbdy,ntr,nreg  (single tile: idm=itdm, jdm=jtdm)
int oracle_init(const int* dims) {
  orc::Oracle& o = orc::O();
  o = orc::Oracle();
  orc::Dims& d = o.d;
  d.itdm = dims[0]; d.jtdm = dims[1]; d.kdm = dims[2]; d.idm = dims[3]; d.jdm = dims[4];
  d.nbdy = dims[5]; d.ntr = dims[6]; d.nreg = dims[7];
  d.i0 = 0; d.j0 = 0; d.ii = d.idm; d.jj = d.jdm; d.kk = d.kdm;
  d.ldi = d.idm + 2 * d.nbdy; d.ldj = d.jdm + 2 * d.nbdy;
  d.lev = (size_t)d.ldi * d.ldj;
  return 0;
}
int oracle_nreg() { return orc::O().d.nreg; }

int oracle_register(const char* name, double* p, int nlev) {
  orc::O().f[name] = orc::Field{p, nlev};
  return 0;
}
int oracle_register_int(const char* name, int* p, int nlev) {
  orc::O().fi[name] = orc::IField{p, nlev};
  return 0;
}
int oracle_set_option(const char* k, const char* v) { orc::O().opt[k] = v; return 0; }
int oracle_set_scalar(const char* k, double v) { orc::O().sc[k] = v; return 0; }

// copy an oracle-owned int array (masks, span tables) out; returns length or -1
long oracle_get_int(const char* name, int* out, long cap) {
  auto& m = orc::O().owni;
  auto it = m.find(name);
  if (it == m.end()) return -1;
  long n = (long)it->second.size();
  if (out && cap >= n) std::memcpy(out, it->second.data(), sizeof(int) * n);
  return n;
}
long oracle_get_owned(const char* name, double* out, long cap) {
  auto& m = orc::O().own;
  auto it = m.find(name);
  if (it == m.end()) return -1;
  long n = (long)it->second.size();
  if (out && cap >= n) std::memcpy(out, it->second.data(), sizeof(double) * n);
  return n;
}
long oracle_cppm_table(const char* name, double* out, long cap) {
  size_t n; const double* p = orc::cppm_table(name, &n);
  if (!p) return -1;
  if (out && cap >= (long)n) std::memcpy(out, p, 8 * n);
  return (long)n;
}
long oracle_cppm_stencil(const char* name, int* out, long cap) {
  size_t n; const int* p = orc::cppm_stencil(name, &n);
  if (!p) return -1;
  if (out && cap >= (long)n) std::memcpy(out, p, 4 * n);
  return (long)n;
}

int oracle_xctilr(const char* name, int l1, int ld, int mh, int nh, int itype) {
  GUARD(orc::xctilr(orc::O().a3(name), l1, ld, mh, nh, itype))
}
// xctilr on a sub-view starting at level `koff` (Fortran a(1-nbdy,1-nbdy,koff))
int oracle_xctilr_at(const char* name, int koff, int l1, int ld, int mh, int nh, int itype) {
  GUARD(orc::xctilr(orc::O().a3(name).from(koff), l1, ld, mh, nh, itype))
}
int oracle_xcsum(const char* name, int lev, const char* mask, double* out) {
  GUARD(*out = orc::xcsum(orc::O().a3(name).level(lev), orc::O().i2(mask)))
}
int oracle_xccrc(const char* name, int ld, const char* mask, uint32_t* out) {
  GUARD(*out = orc::xccrc(orc::O().a3(name), ld, orc::O().i2(mask)))
}
int oracle_xccrc_at(const char* name, int koff, int ld, const char* mask, uint32_t* out) {
  GUARD(*out = orc::xccrc(orc::O().a3(name).from(koff), ld, orc::O().i2(mask)))
}
uint32_t oracle_crc32(const void* p, long n, uint32_t init) { return orc::crc32_bytes(p, (size_t)n, init); }
int oracle_bigrid(const char* depth) { GUARD(orc::bigrid(orc::O().a2(depth))) }

int oracle_init_cppm() { GUARD(orc::init_cppm()) }
int oracle_advect(int m, int n, int mm, int nn, int k1m, int k1n) { GUARD(orc::advect(m, n, mm, nn, k1m, k1n)) }
int oracle_cppm(int m, int n, int mm, int nn, int k1m, int k1n) { GUARD(orc::cppm(m, n, mm, nn, k1m, k1n)) }

int oracle_inieos() { GUARD(orc::inieos()) }
int oracle_diffus(int m, int n, int mm, int nn, int k1m, int k1n) { GUARD(orc::diffus(m, n, mm, nn, k1m, k1n)) }
int oracle_tmsmt1(int nn) { GUARD(orc::tmsmt1(nn)) }
int oracle_tmsmt2(int m, int mm, int nn, int k1m) { GUARD(orc::tmsmt2(m, mm, nn, k1m)) }

int oracle_pgforc(int m, int n, int mm, int nn, int k1m, int k1n) { GUARD(orc::pgforc(m, n, mm, nn, k1m, k1n)) }

int oracle_barotp(int m, int n, int mm, int nn, int k1m, int k1n) { GUARD(orc::barotp(m, n, mm, nn, k1m, k1n)) }

int oracle_eddtra(int m, int n, int mm, int nn, int k1m, int k1n) { GUARD(orc::eddtra(m, n, mm, nn, k1m, k1n)) }
int oracle_pbcor1(int m, int n, int mm, int nn, int k1m, int k1n) { GUARD(orc::pbcor1(m, n, mm, nn, k1m, k1n)) }
int oracle_pbcor2(int m, int n, int mm, int nn, int k1m, int k1n) { GUARD(orc::pbcor2(m, n, mm, nn, k1m, k1n)) }
int oracle_momtum(int m, int n, int mm, int nn, int k1m, int k1n) { GUARD(orc::momtum(m, n, mm, nn, k1m, k1n)) }
int oracle_numerical_bounds() { GUARD(orc::numerical_bounds()) }
int oracle_init_fluxes(int m, int n, int mm, int nn, int k1m, int k1n) { GUARD(orc::init_fluxes(m, n, mm, nn, k1m, k1n)) }
int oracle_ndiff(int m, int n, int mm, int nn, int k1m, int k1n) { GUARD(orc::ndiff(m, n, mm, nn, k1m, k1n)) }
int oracle_cmnfld2(int m, int n, int mm, int nn, int k1m, int k1n) { GUARD(orc::cmnfld2(m, n, mm, nn, k1m, k1n)) }
int oracle_cmnfld_bfsqf_ale(int m, int n, int mm, int nn, int k1m, int k1n) { GUARD(orc::cmnfld_bfsqf_ale(m, n, mm, nn, k1m, k1n)) }
int oracle_cmnfld_nslope_ale(int m, int n, int mm, int nn, int k1m, int k1n) { GUARD(orc::cmnfld_nslope_ale(m, n, mm, nn, k1m, k1n)) }
int oracle_cmnfld_nnslope_ale(int m, int n, int mm, int nn, int k1m, int k1n) { GUARD(orc::cmnfld_nnslope_ale(m, n, mm, nn, k1m, k1n)) }
int oracle_budget_init(double* mass0) { GUARD(orc::budget_init(mass0)) }
int oracle_budget_sums(int ncall, int n, int nn, double* out) { GUARD(orc::budget_sums(ncall, n, nn, out)) }
double oracle_get_scalar(const char* k) { return orc::O().scalar(k, 0.0); }

// scalar access to the EOS restatement for the unit checks in tests/test_oracle_ops.py
int oracle_eos(const char* fn, const double* a, double* out) {
  using namespace orc::eos;
  std::string f(fn);
  if (f == "rho") out[0] = rho(a[0], a[1], a[2]);
  else if (f == "alp") out[0] = alp(a[0], a[1], a[2]);
  else if (f == "sig") out[0] = sig(a[0], a[1]);
  else if (f == "sig0") out[0] = sig0(a[0], a[1]);
  else if (f == "p_alpha") out[0] = p_alpha(a[0], a[1], a[2], a[3]);
  else if (f == "dalpdt") out[0] = dalpdt(a[0], a[1], a[2]);
  else if (f == "dalpds") out[0] = dalpds(a[0], a[1], a[2]);
  else if (f == "delphi") delphi(a[0], a[1], a[2], a[3], out[0], out[1], out[2]);
  else if (f == "dynh_derivatives") dynh_derivatives(a[0], a[1], a[2], a[3], a[4], out[0], out[1]);
  else return 1;
  return 0;
}

}  // extern "C"
