// ORACLE — TEST INFRASTRUCTURE ONLY (see core.hpp header).
// Restatement of the serial half of mod_xc plus mod_crc32 and mod_bigrid.
#include "core.hpp"
#include <cstring>
#include <cstdio>

namespace orc {

Oracle& O() { static Oracle o; return o; }

// ---------------------------------------------------------------------------
// xctilr, serial version.  phy/mod_xc.F90:4222-4428.
// use_ARCTIC  <=>  nreg == 2 (meson.build:141-143).
// ---------------------------------------------------------------------------
void xctilr(A3 a, int l1, int ld, int mh, int nh, int itype) {
  const Dims& d = O().d;
  const int ii = d.ii, jj = d.jj, nbdy = d.nbdy, nreg = d.nreg;
  const double vland = 0.0;                       // :4077
  const int mhl = std::max(0, std::min(mh, nbdy));  // :4258
  const int nhl = std::max(0, std::min(nh, nbdy));  // :4259

  if (nreg == 2) {  // :4262 use_ARCTIC
    for (int k = l1; k <= ld; ++k) {
      // southern boundary is closed (:4267-4271)
      for (int j = 1; j <= nhl; ++j)
        for (int i = 1; i <= ii; ++i) a(i, 1 - j, k) = vland;

      const double sg = itype < 10 ? 1.0 : -1.0;
      const int it = itype % 10;
      if (it == 1) {  // p-grid (:4277-4282, :4320-4325)
        for (int j = 0; j <= nhl; ++j)
          for (int i = 1; i <= ii; ++i) {
            int io = ii - ((i - 1) % ii);
            a(i, jj + j, k) = itype < 10 ? a(io, jj - 1 - j, k) : -a(io, jj - 1 - j, k);
          }
      } else if (it == 2) {  // q-grid (:4285-4294, :4328-4337)
        for (int i = ii / 2 + 1; i <= ii; ++i) {
          int io = ((ii - (i - 1)) % ii) + 1;
          a(i, jj, k) = itype < 10 ? a(io, jj, k) : -a(io, jj, k);
        }
        for (int j = 1; j <= nhl; ++j)
          for (int i = 1; i <= ii; ++i) {
            int io = ((ii - (i - 1)) % ii) + 1;
            a(i, jj + j, k) = itype < 10 ? a(io, jj - j, k) : -a(io, jj - j, k);
          }
      } else if (it == 3) {  // u-grid (:4297-4302, :4340-4345)
        for (int j = 0; j <= nhl; ++j)
          for (int i = 1; i <= ii; ++i) {
            int io = ((ii - (i - 1)) % ii) + 1;
            a(i, jj + j, k) = itype < 10 ? a(io, jj - 1 - j, k) : -a(io, jj - 1 - j, k);
          }
      } else {  // v-grid (:4305-4314, :4348-4357)
        for (int i = ii / 2 + 1; i <= ii; ++i) {
          int io = ii - ((i - 1) % ii);
          a(i, jj, k) = itype < 10 ? a(io, jj, k) : -a(io, jj, k);
        }
        for (int j = 1; j <= nhl; ++j)
          for (int i = 1; i <= ii; ++i) {
            int io = ii - ((i - 1) % ii);
            a(i, jj + j, k) = itype < 10 ? a(io, jj - j, k) : -a(io, jj - j, k);
          }
      }
      (void)sg;
    }
    if (mhl > 0) {
      for (int k = 1; k <= ld; ++k)  // NB: reference loops from 1, not l1 (:4363)
        for (int j = 1 - nhl; j <= jj + nhl; ++j)
          for (int i = 1; i <= mhl; ++i) {
            a(1 - i, j, k) = a(ii + 1 - i, j, k);
            a(ii + i, j, k) = a(i, j, k);
          }
    }
  } else {  // :4374 NOT use_ARCTIC
    if (nhl > 0) {
      if (nreg <= 2) {  // closed in latitude (:4378-4386)
        for (int k = l1; k <= ld; ++k)
          for (int j = 1; j <= nhl; ++j)
            for (int i = 1; i <= ii; ++i) {
              a(i, 1 - j, k) = vland;
              a(i, jj + j, k) = vland;
            }
      } else {  // periodic in latitude (:4388-4395)
        for (int k = l1; k <= ld; ++k)
          for (int j = 1; j <= nhl; ++j)
            for (int i = 1; i <= ii; ++i) {
              a(i, 1 - j, k) = a(i, jj + 1 - j, k);
              a(i, jj + j, k) = a(i, j, k);
            }
      }
    }
    if (mhl > 0) {
      if (nreg == 0 || nreg == 4) {  // closed in longitude (:4400-4408)
        for (int k = l1; k <= ld; ++k)
          for (int j = 1 - nhl; j <= jj + nhl; ++j)
            for (int i = 1; i <= mhl; ++i) {
              a(1 - i, j, k) = vland;
              a(ii + i, j, k) = vland;
            }
      } else {  // periodic in longitude (:4410-4417)
        for (int k = l1; k <= ld; ++k)
          for (int j = 1 - nhl; j <= jj + nhl; ++j)
            for (int i = 1; i <= mhl; ++i) {
              a(1 - i, j, k) = a(ii + 1 - i, j, k);
              a(ii + i, j, k) = a(i, j, k);
            }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// xcsum, serial.  phy/mod_xc.F90:4116-4161.  Association order is the contract.
// ---------------------------------------------------------------------------
double xcsum(A2 a, I2 mask) {
  const Dims& d = O().d;
  const int idm = d.idm, jdm = d.jdm, nbdy = d.nbdy;
  std::vector<double> sum8j(jdm + 1);
  for (int j = 1; j <= jdm; ++j) {
    double sum8 = 0.0;
    for (int i1 = 1; i1 <= idm; i1 += 2 * nbdy + 1) {
      double sum8p = 0.0;
      for (int i = i1; i <= std::min(i1 + 2 * nbdy, idm); ++i)
        if (mask(i, j) == 1) sum8p = sum8p + a(i, j);
      sum8 = sum8 + sum8p;
    }
    sum8j[j] = sum8;
  }
  double sum8 = sum8j[1];
  for (int j = 2; j <= jdm; ++j) sum8 = sum8 + sum8j[j];
  return sum8;
}

// ---------------------------------------------------------------------------
// mod_crc32: table-driven CRC-32 (poly 0xEDB88320 == -306674912),
// phy/mod_crc32.F90:69-88 (table), :600-653 (real64), :305-331 (int32).
// ---------------------------------------------------------------------------
static uint32_t crc_table[256];
static bool table_initialized = false;
static void init_table() {
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t k = i;
    for (int j = 0; j < 8; ++j) k = (k & 1u) ? ((k >> 1) ^ 0xEDB88320u) : (k >> 1);
    crc_table[i] = k;
  }
  table_initialized = true;
}
uint32_t crc32_bytes(const void* data, size_t n, uint32_t crc_init) {
  if (!table_initialized) init_table();
  const unsigned char* b = static_cast<const unsigned char*>(data);
  uint32_t crc = ~crc_init;
  for (size_t j = 0; j < n; ++j) crc = (crc >> 8) ^ crc_table[(crc ^ b[j]) & 255u];
  return ~crc;
}

// ---------------------------------------------------------------------------
// xccrc, serial.  phy/mod_xc.F90:4164-4205.
// ---------------------------------------------------------------------------
uint32_t xccrc(A3 a, int ld, I2 mask) {
  const Dims& d = O().d;
  const int idm = d.idm, jdm = d.jdm, nbdy = d.nbdy;
  std::vector<uint32_t> crc8j(jdm);
  std::vector<double> col(ld);
  for (int j = 1; j <= jdm; ++j) {
    uint32_t crc8 = 0;
    for (int i1 = 1; i1 <= idm; i1 += 2 * nbdy + 1) {
      uint32_t crc8p = 0;
      for (int i = i1; i <= std::min(i1 + 2 * nbdy, idm); ++i)
        if (mask(i, j) == 1) {
          for (int k = 1; k <= ld; ++k) col[k - 1] = a(i, j, k);
          crc8p = crc32_bytes(col.data(), 8 * (size_t)ld, crc8p);
        }
      crc8 = crc32_bytes(&crc8p, 4, crc8);
    }
    crc8j[j - 1] = crc8;
  }
  return crc32_bytes(crc8j.data(), 4 * (size_t)jdm, 0);
}

// ---------------------------------------------------------------------------
// bigrid + indxi/indxj.  phy/mod_bigrid.F90:44-429 (single tile: i0=j0=0).
// Span tables are stored as int fields "ifp","ilp" ([ (j+nb-1)*ms + l-1 ]),
// "isp" etc. in Oracle::owni.
// ---------------------------------------------------------------------------
static void indxi(I2 ipt, const char* nf, const char* nl, const char* ns) {
  Oracle& o = O(); const Dims& d = o.d;
  const int nb = d.nbdy, ii = d.ii, jj = d.jj, ms = Oracle::ms;
  auto& vf = o.owni[nf]; auto& vl = o.owni[nl]; auto& vs = o.owni[ns];
  vf.assign((size_t)d.ldj * ms, 0); vl.assign((size_t)d.ldj * ms, 0); vs.assign(d.ldj, 0);
  for (int j = 1 - nb; j <= jj + nb; ++j) {
    int* f = &vf[(size_t)(j + nb - 1) * ms]; int* l = &vl[(size_t)(j + nb - 1) * ms];
    int k = 1;
    int last = ipt(1 - nb, j);
    if (last == 1) f[k - 1] = 1 - nb;
    for (int i = 2 - nb; i <= ii + nb; ++i) {
      if (last == 1 && ipt(i, j) == 0) { l[k - 1] = i - 1; k = k + 1; }
      else if (last == 0 && ipt(i, j) == 1) {
        if (k > ms) throw std::runtime_error("indxi -- ms too small");
        f[k - 1] = i;
      }
      last = ipt(i, j);
    }
    if (last == 1) { l[k - 1] = ii + nb; vs[j + nb - 1] = k; }
    else vs[j + nb - 1] = k - 1;
  }
}
static void indxj(I2 jpt, const char* nf, const char* nl, const char* ns) {
  Oracle& o = O(); const Dims& d = o.d;
  const int nb = d.nbdy, ii = d.ii, jj = d.jj, ms = Oracle::ms;
  auto& vf = o.owni[nf]; auto& vl = o.owni[nl]; auto& vs = o.owni[ns];
  vf.assign((size_t)d.ldi * ms, 0); vl.assign((size_t)d.ldi * ms, 0); vs.assign(d.ldi, 0);
  for (int i = 1 - nb; i <= ii + nb; ++i) {
    int* f = &vf[(size_t)(i + nb - 1) * ms]; int* l = &vl[(size_t)(i + nb - 1) * ms];
    int k = 1;
    int last = jpt(i, 1 - nb);
    if (last == 1) f[k - 1] = 1 - nb;
    for (int j = 2 - nb; j <= jj + nb; ++j) {
      if (last == 1 && jpt(i, j) == 0) { l[k - 1] = j - 1; k = k + 1; }
      else if (last == 0 && jpt(i, j) == 1) {
        if (k > ms) throw std::runtime_error("indxj -- ms too small");
        f[k - 1] = j;
      }
      last = jpt(i, j);
    }
    if (last == 1) { l[k - 1] = jj + nb; vs[i + nb - 1] = k; }
    else vs[i + nb - 1] = k - 1;
  }
}

void bigrid(A2 depth) {
  Oracle& o = O(); Dims& d = o.d;
  const int nb = d.nbdy, ii = d.ii, jj = d.jj, idm = d.idm, jdm = d.jdm;
  // :59-78 periodicity detection (single tile => i0+ii==itdm, j0+jj==jtdm)
  double depmax = 0.0;
  for (int j = 1; j <= jj; ++j) depmax = std::max(depmax, depth(ii, j));
  const bool lperiodi = depmax > 0.0;
  depmax = 0.0;
  for (int i = 1; i <= ii; ++i) depmax = std::max(depmax, depth(i, jj));
  const bool larctic = depmax > 0.0 && d.nreg == 2;
  const bool lperiodj = depmax > 0.0 && d.nreg != 2;
  // :81-107
  int& nreg = d.nreg;
  if (!lperiodi && !lperiodj && (nreg == 0 || nreg == -1)) nreg = 0;
  else if (lperiodi && !lperiodj && (nreg == 1 || nreg == -1)) nreg = 1;
  else if (lperiodi && larctic && (nreg == 2 || nreg == -1)) nreg = 2;
  else if (lperiodi && lperiodj && (nreg == 3 || nreg == -1)) nreg = 3;
  else if (!lperiodi && lperiodj && (nreg == 4 || nreg == -1)) nreg = 4;
  else throw std::runtime_error("bigrid: basin depth array inconsistent with nreg");

  xctilr(depth, nb, nb, halo_ps);  // :126

  // :129-163 part I
  if (!lperiodj)
    for (int j = 1 - nb; j <= 0; ++j)
      for (int i = 1 - nb; i <= ii + nb; ++i) depth(i, j) = 0.0;
  if (!lperiodj && !larctic)
    for (int j = jj + 1; j <= jj + nb; ++j)
      for (int i = 1 - nb; i <= ii + nb; ++i) depth(i, j) = 0.0;
  if (!lperiodi) {
    for (int j = 1 - nb; j <= jj + nb; ++j)
      for (int i = 1 - nb; i <= 0; ++i) depth(i, j) = 0.0;
    for (int j = 1 - nb; j <= jj + nb; ++j)
      for (int i = ii + 1; i <= ii + nb; ++i) depth(i, j) = 0.0;
  }
  // :165-193 single-width inlets / 1-point seas
  int nfill = 0;
  for (int j = 1; j <= jj; ++j)
    for (int i = 1; i <= ii; ++i) {
      int nzero = 0;
      if (depth(i, j) > 0.0) {
        if (depth(i - 1, j) <= 0.0) nzero++;
        if (depth(i + 1, j) <= 0.0) nzero++;
        if (depth(i, j - 1) <= 0.0) nzero++;
        if (depth(i, j + 1) <= 0.0) nzero++;
        if (nzero >= 3) nfill++;
      }
    }
  if (nfill > 0) throw std::runtime_error("bigrid: Must correct bathymetry before running BLOM");

  I2 ip = o.iscratch("ip"), iq = o.iscratch("iq"), iu = o.iscratch("iu"), iv = o.iscratch("iv");
  std::vector<double> b1(d.lev, 0.0), b2(d.lev, 0.0), b3(d.lev, 0.0);
  A2 util1{b1.data(), d.ldi, nb}, util2{b2.data(), d.ldi, nb}, util3{b3.data(), d.ldi, nb};
  for (int j = 1 - nb; j <= jdm + nb; ++j)
    for (int i = 1 - nb; i <= idm + nb; ++i) {
      ip(i, j) = 0; iq(i, j) = 0; iu(i, j) = 0; iv(i, j) = 0;
    }
  for (int j = 1 - nb; j <= jj + nb; ++j)
    for (int i = 1 - nb; i <= ii + nb; ++i)
      if (depth(i, j) > 0.) ip(i, j) = 1;
  for (int j = 1; j <= jj; ++j)
    for (int i = 1; i <= ii; ++i) {
      if (ip(i - 1, j) > 0 && ip(i, j) > 0) iu(i, j) = 1;
      if (ip(i, j - 1) > 0 && ip(i, j) > 0) iv(i, j) = 1;
      if (std::min(std::min(ip(i, j), ip(i - 1, j)), std::min(ip(i, j - 1), ip(i - 1, j - 1))) > 0)
        iq(i, j) = 1;
      else if ((ip(i, j) > 0 && ip(i - 1, j - 1) > 0) || (ip(i - 1, j) > 0 && ip(i, j - 1) > 0))
        iq(i, j) = 1;
      util1(i, j) = iu(i, j); util2(i, j) = iv(i, j); util3(i, j) = iq(i, j);
    }
  xctilr(util1, nb, nb, halo_us);
  xctilr(util2, nb, nb, halo_vs);
  xctilr(util3, nb, nb, halo_qs);
  for (int j = 1 - nb; j <= jj + nb; ++j)
    for (int i = 1 - nb; i <= ii + nb; ++i) {
      iu(i, j) = (int)std::lround(util1(i, j));
      iv(i, j) = (int)std::lround(util2(i, j));
      iq(i, j) = (int)std::lround(util3(i, j));
    }
  // :259-302 part II
  auto zero = [&](int j0, int j1, int i0, int i1) {
    for (int j = j0; j <= j1; ++j)
      for (int i = i0; i <= i1; ++i) { iq(i, j) = 0; iu(i, j) = 0; iv(i, j) = 0; }
  };
  if (!lperiodj) zero(1 - nb, 0, 1 - nb, ii + nb);
  if (!lperiodj && !larctic) zero(jj + 1, jj + nb, 1 - nb, ii + nb);
  if (!lperiodi) { zero(1 - nb, jj + nb, 1 - nb, 0); zero(1 - nb, jj + nb, ii + 1, ii + nb); }

  indxi(iq, "ifq", "ilq", "isq"); indxj(iq, "jfq", "jlq", "jsq");
  indxi(ip, "ifp", "ilp", "isp"); indxj(ip, "jfp", "jlp", "jsp");
  indxi(iu, "ifu", "ilu", "isu"); indxj(iu, "jfu", "jlu", "jsu");
  indxi(iv, "ifv", "ilv", "isv"); indxj(iv, "jfv", "jlv", "jsv");
}

}  // namespace orc
