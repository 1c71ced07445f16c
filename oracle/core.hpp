// ORACLE — TEST INFRASTRUCTURE ONLY.
// CPU restatement of the BLOM reference algorithms for the horizontal stencil
// step.  Nothing in blom_b200/ (the product) may include, link or call this.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs use it, and only as the checker / reported baseline.
//
// Parity status: "parity unpinned" by the reference itself (the reference has
// no golden vectors and cannot be compiled in this image: no Fortran
// toolchain).  Pinned here by (i) the CRC-32 check value
// 0xCBF43926 for mod_crc32, (ii) hand-derived fold/halo index fixtures under
// tests/golden/, (iii) analytic invariants (uniform-field preservation,
// conservation) of the restated routines, (iv) the physical known answer of
// the reference's own idealized test: the fuk95 density front adjusts to a
// geostrophic jet of the speed u0 it was built for (tests/test_oracle_fuk95.py).
//
// Conventions: every `real` of the reference is real(8) (meson.build:10
// -fdefault-real-8); arrays are column-major a(1-nbdy:idm+nbdy,
// 1-nbdy:jdm+nbdy[,k]) with i fastest (phy/mod_xc.F90:45,
// phy/mod_state.F90:34-86).  Views below use the Fortran index origin so the
// restated loops read like the source.  Compile with -ffp-contract=off
// (meson.build:17-19).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cmath>
#include <map>
#include <string>
#include <vector>
#include <algorithm>
#include <stdexcept>

namespace orc {

// phy/mod_constants.F90:30-56
constexpr double grav = 9.806, alpha0 = 1.e-3, rho0 = 1.e3;
constexpr double epsilpl = 1.e-14, epsilp = 1.e-12, spval = 1.e33;
constexpr double onem = 9806., onecm = 98.06, onemm = 9.806, onemu = .009806;
constexpr double tenm = 98060.;

// xctilr itype codes (phy/mod_xc.F90:4238-4246, halo_* parameters :95-104)
enum { halo_ps = 1, halo_qs = 2, halo_us = 3, halo_vs = 4,
       halo_pv = 11, halo_qv = 12, halo_uv = 13, halo_vv = 14 };

struct Dims {
  int itdm = 0, jtdm = 0, kdm = 0, idm = 0, jdm = 0, nbdy = 4, ntr = 0, nreg = -1;
  int i0 = 0, j0 = 0, ii = 0, jj = 0, kk = 0;
  int ldi = 0, ldj = 0;
  size_t lev = 0;
};

struct A2 {
  double* p = nullptr; int ldi = 0, nb = 0;
  inline double& operator()(int i, int j) const {
    return p[(size_t)(j + nb - 1) * ldi + (i + nb - 1)];
  }
};
struct I2 {
  int* p = nullptr; int ldi = 0, nb = 0;
  inline int& operator()(int i, int j) const {
    return p[(size_t)(j + nb - 1) * ldi + (i + nb - 1)];
  }
};
struct A3 {
  double* p = nullptr; int ldi = 0, nb = 0; size_t lev = 0;
  inline double& operator()(int i, int j, int k) const {
    return p[(size_t)(k - 1) * lev + (size_t)(j + nb - 1) * ldi + (i + nb - 1)];
  }
  inline A2 level(int k) const { return A2{p + (size_t)(k - 1) * lev, ldi, nb}; }
  // view starting at level k (Fortran a(1-nbdy,1-nbdy,k) actual argument)
  inline A3 from(int k) const { return A3{p + (size_t)(k - 1) * lev, ldi, nb, lev}; }
};

struct Field { double* p; int nlev; };
struct IField { int* p; int nlev; };

struct Oracle {
  Dims d;
  std::map<std::string, Field> f;
  std::map<std::string, IField> fi;
  std::map<std::string, std::string> opt;
  std::map<std::string, double> sc;
  // owned scratch (routine-local / module-private arrays of the reference)
  std::map<std::string, std::vector<double>> own;
  std::map<std::string, std::vector<int>> owni;
  // span tables of bigrid (phy/mod_xc.F90:60-92); [j+nb-1][l] flattened, ms=100
  static constexpr int ms = 100;

  A2 a2(const std::string& n) const {
    auto it = f.find(n);
    if (it == f.end()) throw std::runtime_error("oracle: field not registered: " + n);
    return A2{it->second.p, d.ldi, d.nbdy};
  }
  A3 a3(const std::string& n) const {
    auto it = f.find(n);
    if (it == f.end()) throw std::runtime_error("oracle: field not registered: " + n);
    return A3{it->second.p, d.ldi, d.nbdy, d.lev};
  }
  bool has(const std::string& n) const { return f.count(n) != 0; }
  I2 i2(const std::string& n) const {
    auto it = fi.find(n);
    if (it == fi.end()) throw std::runtime_error("oracle: int field not registered: " + n);
    return I2{it->second.p, d.ldi, d.nbdy};
  }
  // allocate (or fetch) an oracle-owned array of nlev levels, zero-filled on creation
  A3 scratch(const std::string& n, int nlev) {
    auto& v = own[n];
    if (v.size() != d.lev * (size_t)nlev) v.assign(d.lev * (size_t)nlev, 0.0);
    f[n] = Field{v.data(), nlev};
    return A3{v.data(), d.ldi, d.nbdy, d.lev};
  }
  I2 iscratch(const std::string& n, int nlev = 1) {
    auto& v = owni[n];
    if (v.size() != d.lev * (size_t)nlev) v.assign(d.lev * (size_t)nlev, 0);
    fi[n] = IField{v.data(), nlev};
    return I2{v.data(), d.ldi, d.nbdy};
  }
  double scalar(const std::string& k) const {
    auto it = sc.find(k);
    if (it == sc.end()) throw std::runtime_error("oracle: scalar not set: " + k);
    return it->second;
  }
  double scalar(const std::string& k, double dflt) const {
    auto it = sc.find(k);
    return it == sc.end() ? dflt : it->second;
  }
  std::string option(const std::string& k, const std::string& dflt) const {
    auto it = opt.find(k);
    return it == opt.end() ? dflt : it->second;
  }
};

Oracle& O();

// mod_xc serial primitives (xc.cpp)
void xctilr(A3 a, int l1, int ld, int mh, int nh, int itype);
inline void xctilr(A2 a, int mh, int nh, int itype) {
  const Dims& d = O().d;
  xctilr(A3{a.p, a.ldi, a.nb, d.lev}, 1, 1, mh, nh, itype);
}
double xcsum(A2 a, I2 mask);
uint32_t xccrc(A3 a, int ld, I2 mask);
uint32_t crc32_bytes(const void* data, size_t n, uint32_t crc_init);
void bigrid(A2 depth);

inline double fsign(double a, double b) {  // Fortran sign(a,b)
  double m = std::fabs(a);
  return std::signbit(b) ? -m : m;
}

}  // namespace orc
