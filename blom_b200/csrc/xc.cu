// mod_xc on the device: halo update (xctilr incl. the tripolar fold), the
// bit-reproducible strip-ordered global sum (xcsum), exact max/min, the masked
// CRC-32 field checksum (xccrc / chksum) and the mask builder (bigrid).
//
// Reference semantics: phy/mod_xc.F90:4222-4428 (xctilr, one tile), :4116-4161
// (xcsum), :4164-4205 (xccrc), phy/mod_crc32.F90:69-88, phy/mod_bigrid.F90:44-317.
// With j-band tiles the fold partner of the northern tile is the tile itself
// (phy/mod_xc.F90:1608-1615), so the fold is always an on-device gather; band
// edges between GPUs are filled by exchange_ns() in comm.cu before the kernel
// below runs.  The result is bit-identical to the one-tile result.
#include "common.cuh"
#include "halo.cuh"

namespace blom {

void exchange_ns(const std::vector<HaloReq>& reqs, int nhl);  // comm.cu

__global__ void halo_kernel(Geom g, HaloBatch b, int mhl, int nhl, int ns_l1, int ew_l1) {
  const int r = blockIdx.z;
  const int k = blockIdx.y + 1;
  if (k > b.nlev[r]) return;
  halo_level(g, b.base[r] + (long)(k - 1) * g.lev, b.itype[r], k, mhl, nhl, ns_l1, ew_l1,
             (long)blockIdx.x * blockDim.x + threadIdx.x, (long)gridDim.x * blockDim.x);
}

static void halo_launch(const std::vector<HaloReq>& reqs, int mh, int nh, int l1) {
  Ctx& c = C();
  const Geom& g = c.g;
  const int mhl = std::max(0, std::min(mh, g.nb)), nhl = std::max(0, std::min(nh, g.nb));
  if (g.nranks > 1 && nhl > 0) exchange_ns(reqs, nhl);
  const bool fold = (g.nreg == 2 && g.north);
  if (mhl == 0 && nhl == 0 && !fold) return;
  for (size_t s = 0; s < reqs.size(); s += HALO_MAX) {
    HaloBatch b{};
    int n = (int)std::min((size_t)HALO_MAX, reqs.size() - s), maxlev = 0;
    for (int r = 0; r < n; ++r) {
      b.base[r] = reqs[s + r].base; b.nlev[r] = reqs[s + r].nlev; b.itype[r] = reqs[s + r].itype;
      maxlev = std::max(maxlev, b.nlev[r]);
    }
    long cells = (long)(2 * nhl + 1) * g.ii + (long)2 * mhl * (g.jj + 2 * nhl);
    dim3 grid(std::max(1, std::min(cdiv(cells, 256), 64)), maxlev, n);
    const int ew_l1 = (g.nreg == 2) ? 1 : l1;
    LAUNCH(halo_kernel, grid, 256, 0, g, b, mhl, nhl, l1, ew_l1);
  }
}

void halo_update(const std::vector<HaloReq>& reqs, int mh, int nh) { halo_launch(reqs, mh, nh, 1); }

void xctilr_exact(double* base, int l1, int ld, int mh, int nh, int itype) {
  halo_launch(std::vector<HaloReq>{HaloReq{base, ld, itype}}, mh, nh, l1);
}

// ---------------------------------------------------------------------------
// xcsum: 9-wide strips (2*nbdy+1) in fixed global positions, strips added
// left to right, rows south to north.  Strip partials are the parallel stage;
// the two serial tails (<=320 and <=2165 adds) keep the reference order.
// ---------------------------------------------------------------------------
__global__ void xcsum_strips(Geom g, const double* __restrict__ a, const int* __restrict__ mask,
                             int nstrip, double* __restrict__ strip) {
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long)nstrip * g.jj) return;
  int j = (int)(t / nstrip) + 1, s = (int)(t % nstrip);
  int w = 2 * g.nb + 1, i1 = 1 + s * w;
  double sum8p = 0.0;
  for (int i = i1; i <= min(i1 + 2 * g.nb, g.idm); ++i)
    if (mask[ix2(g, i, j)] == 1) sum8p = sum8p + a[ix2(g, i, j)];
  strip[t] = sum8p;
}
__global__ void xcsum_rows(int jj, int nstrip, const double* __restrict__ strip, double* __restrict__ rows) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= jj) return;
  double sum8 = 0.0;
  for (int s = 0; s < nstrip; ++s) sum8 = sum8 + strip[(long)j * nstrip + s];
  rows[j] = sum8;
}
__global__ void xcsum_total(int n, const double* __restrict__ rows, double* __restrict__ out) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double sum8 = rows[0];
    for (int j = 1; j < n; ++j) sum8 = sum8 + rows[j];
    out[0] = sum8;
  }
}

double* gather_rows(double* rows_local, int jj_local, int* n_total);  // comm.cu

static void ensure_red(size_t n) {
  Ctx& c = C();
  if (n <= c.red_cap) return;
  CUDA_CHECK(cudaStreamSynchronize(c.stream));
  cudaFree(c.d_red); cudaFreeHost(c.h_red);
  c.red_cap = n;
  CUDA_CHECK(cudaMalloc(&c.d_red, n * sizeof(double)));
  CUDA_CHECK(cudaMallocHost(&c.h_red, n * sizeof(double)));
}

double xcsum_dev(const double* a, const int* mask) {
  Ctx& c = C(); const Geom& g = c.g;
  const int w = 2 * g.nb + 1, nstrip = (g.idm + w - 1) / w;
  ensure_red((size_t)nstrip * g.jj + g.jtdm + 8);
  double* strip = c.d_red;
  double* rows = c.d_red + (size_t)nstrip * g.jj;
  LAUNCH(xcsum_strips, cdiv((long)nstrip * g.jj, 256), 256, 0, g, a, mask, nstrip, strip);
  LAUNCH(xcsum_rows, cdiv(g.jj, 128), 128, 0, g.jj, nstrip, strip, rows);
  int ntot = g.jj;
  double* allrows = rows;
  if (g.nranks > 1) allrows = gather_rows(rows, g.jj, &ntot);
  double* out = c.d_red + c.red_cap - 1;
  LAUNCH(xcsum_total, 1, 32, 0, ntot, allrows, out);
  CUDA_CHECK(cudaMemcpyAsync(c.h_red, out, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  CUDA_CHECK(cudaStreamSynchronize(c.stream));
  return c.h_red[0];
}

// ---------------------------------------------------------------------------
// xcmax / xcmin: exactly associative, so a shuffle tree is bit-identical.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double warp_minmax(double v, bool is_max) {
  for (int o = 16; o > 0; o >>= 1) {
    double w = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmax(v, w) : fmin(v, w);
  }
  return v;
}
__global__ void minmax_kernel(Geom g, const double* __restrict__ a, const int* __restrict__ mask,
                              bool is_max, double* __restrict__ part) {
  __shared__ double sm[32];
  const double init = is_max ? -1.7976931348623157e308 : 1.7976931348623157e308;
  double v = init;
  const long n = (long)g.ii * g.jj;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x) {
    int j = (int)(t / g.ii) + 1, i = (int)(t % g.ii) + 1;
    long x = ix2(g, i, j);
    if (mask == nullptr || mask[x] == 1) v = is_max ? fmax(v, a[x]) : fmin(v, a[x]);
  }
  v = warp_minmax(v, is_max);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : init;
    v = warp_minmax(v, is_max);
    if (threadIdx.x == 0) part[blockIdx.x] = v;
  }
}
__global__ void minmax_final(int n, bool is_max, const double* __restrict__ part, double* __restrict__ out) {
  const double init = is_max ? -1.7976931348623157e308 : 1.7976931348623157e308;
  double v = init;
  for (int t = threadIdx.x; t < n; t += 32) v = is_max ? fmax(v, part[t]) : fmin(v, part[t]);
  v = warp_minmax(v, is_max);
  if (threadIdx.x == 0) out[0] = v;
}

void allreduce_minmax(double* d_val, bool is_max);  // comm.cu

double xcmax_dev(const double* a, const int* mask, bool is_max) {
  Ctx& c = C(); const Geom& g = c.g;
  const int nblk = 296;
  ensure_red(nblk + 8);
  LAUNCH(minmax_kernel, nblk, 256, 0, g, a, mask, is_max, c.d_red);
  LAUNCH(minmax_final, 1, 32, 0, nblk, is_max, c.d_red, c.d_red + nblk);
  if (g.nranks > 1) allreduce_minmax(c.d_red + nblk, is_max);
  CUDA_CHECK(cudaMemcpyAsync(c.h_red, c.d_red + nblk, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  CUDA_CHECK(cudaStreamSynchronize(c.stream));
  return c.h_red[0];
}

// ---------------------------------------------------------------------------
// xccrc: CRC-32 (poly 0xEDB88320), per point chained over the ld levels,
// chained over the points of a strip, strips chained into the row value,
// final CRC over the row values.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t crc_table_entry(uint32_t i) {
  uint32_t k = i;
#pragma unroll
  for (int j = 0; j < 8; ++j) k = (k & 1u) ? ((k >> 1) ^ 0xEDB88320u) : (k >> 1);
  return k;
}
__device__ __forceinline__ uint32_t crc_bytes8(const uint32_t* tab, uint32_t crc, unsigned long long bits) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    crc = (crc >> 8) ^ tab[(crc ^ (uint32_t)(bits & 0xffu)) & 255u];
    bits >>= 8;
  }
  return crc;
}
__device__ __forceinline__ uint32_t crc_bytes4(const uint32_t* tab, uint32_t crc, uint32_t bits) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    crc = (crc >> 8) ^ tab[(crc ^ (bits & 0xffu)) & 255u];
    bits >>= 8;
  }
  return crc;
}
__global__ void crc_strips(Geom g, const double* __restrict__ a, int ld, const int* __restrict__ mask,
                           int nstrip, uint32_t* __restrict__ strip) {
  __shared__ uint32_t tab[256];
  for (int t = threadIdx.x; t < 256; t += blockDim.x) tab[t] = crc_table_entry(t);
  __syncthreads();
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long)nstrip * g.jj) return;
  int j = (int)(t / nstrip) + 1, s = (int)(t % nstrip);
  int w = 2 * g.nb + 1, i1 = 1 + s * w;
  uint32_t crc8p = 0;
  for (int i = i1; i <= min(i1 + 2 * g.nb, g.idm); ++i)
    if (mask[ix2(g, i, j)] == 1) {
      uint32_t crc = ~crc8p;
      for (int k = 1; k <= ld; ++k)
        crc = crc_bytes8(tab, crc, (unsigned long long)__double_as_longlong(a[ix3(g, i, j, k)]));
      crc8p = ~crc;
    }
  strip[t] = crc8p;
}
__global__ void crc_rows(int jj, int nstrip, const uint32_t* __restrict__ strip, uint32_t* __restrict__ rows) {
  __shared__ uint32_t tab[256];
  for (int t = threadIdx.x; t < 256; t += blockDim.x) tab[t] = crc_table_entry(t);
  __syncthreads();
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= jj) return;
  uint32_t crc8 = 0;
  for (int s = 0; s < nstrip; ++s) crc8 = ~crc_bytes4(tab, ~crc8, strip[(long)j * nstrip + s]);
  rows[j] = crc8;
}
__global__ void crc_total(int n, const uint32_t* __restrict__ rows, uint32_t* __restrict__ out) {
  __shared__ uint32_t tab[256];
  for (int t = threadIdx.x; t < 256; t += blockDim.x) tab[t] = crc_table_entry(t);
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t crc = ~0u;
    for (int j = 0; j < n; ++j) crc = crc_bytes4(tab, crc, rows[j]);
    out[0] = ~crc;
  }
}

uint32_t* gather_rows_u32(uint32_t* rows_local, int jj_local, int* n_total);  // comm.cu

uint32_t xccrc_dev(const double* a, int ld, const int* mask) {
  Ctx& c = C(); const Geom& g = c.g;
  const int w = 2 * g.nb + 1, nstrip = (g.idm + w - 1) / w;
  ensure_red(((size_t)nstrip * g.jj + g.jtdm) / 2 + 16);
  uint32_t* strip = reinterpret_cast<uint32_t*>(c.d_red);
  uint32_t* rows = strip + (size_t)nstrip * g.jj;
  LAUNCH(crc_strips, cdiv((long)nstrip * g.jj, 128), 128, 0, g, a, ld, mask, nstrip, strip);
  LAUNCH(crc_rows, cdiv(g.jj, 128), 128, 0, g.jj, nstrip, strip, rows);
  int ntot = g.jj;
  uint32_t* allrows = rows;
  if (g.nranks > 1) allrows = gather_rows_u32(rows, g.jj, &ntot);
  uint32_t* out = reinterpret_cast<uint32_t*>(c.d_red + c.red_cap - 1);
  LAUNCH(crc_total, 1, 256, 0, ntot, allrows, out);
  CUDA_CHECK(cudaMemcpyAsync(c.h_red, out, sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
  CUDA_CHECK(cudaStreamSynchronize(c.stream));
  return *reinterpret_cast<uint32_t*>(c.h_red);
}

// ---------------------------------------------------------------------------
// bigrid: wet masks from depth.  phy/mod_bigrid.F90:44-317.
// ---------------------------------------------------------------------------
__global__ void edge_max_kernel(Geom g, const double* __restrict__ depth, double* __restrict__ out) {
  // out[0] = max depth(ii, 1..jj); out[1] = max depth(1..ii, jj)
  __shared__ double sm[2][32];
  double v0 = 0.0, v1 = 0.0;
  for (int j = threadIdx.x + 1; j <= g.jj; j += blockDim.x) v0 = fmax(v0, depth[ix2(g, g.ii, j)]);
  for (int i = threadIdx.x + 1; i <= g.ii; i += blockDim.x) v1 = fmax(v1, depth[ix2(g, i, g.jj)]);
  v0 = warp_minmax(v0, true); v1 = warp_minmax(v1, true);
  if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = v0; sm[1][threadIdx.x >> 5] = v1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (blockDim.x >> 5); ++w) { v0 = fmax(v0, sm[0][w]); v1 = fmax(v1, sm[1][w]); }
    out[0] = v0; out[1] = v1;
  }
}
// zero the non-periodic / non-arctic outer frames of a double array (part I)
__global__ void zero_frames_d(Geom g, double* a, int zs, int zn, int zw, int ze) {
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= g.lev) return;
  int i = (int)(t % g.ldi) + 1 - g.nb, j = (int)(t / g.ldi) + 1 - g.nb;
  if ((zs && j <= 0) || (zn && j > g.jj) || (zw && i <= 0) || (ze && i > g.ii)) a[t] = 0.0;
}
__global__ void inlet_check(Geom g, const double* __restrict__ depth, int* __restrict__ nfill) {
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long)g.ii * g.jj) return;
  int i = (int)(t % g.ii) + 1, j = (int)(t / g.ii) + 1;
  if (depth[ix2(g, i, j)] > 0.0) {
    int nzero = 0;
    if (depth[ix2(g, i - 1, j)] <= 0.0) nzero++;
    if (depth[ix2(g, i + 1, j)] <= 0.0) nzero++;
    if (depth[ix2(g, i, j - 1)] <= 0.0) nzero++;
    if (depth[ix2(g, i, j + 1)] <= 0.0) nzero++;
    if (nzero >= 3) atomicAdd(nfill, 1);
  }
}
__global__ void mask_p(Geom g, const double* __restrict__ depth, int* ip, int* iq, int* iu, int* iv) {
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= g.lev) return;
  ip[t] = depth[t] > 0. ? 1 : 0;
  iq[t] = 0; iu[t] = 0; iv[t] = 0;
}
__global__ void mask_uvq(Geom g, const int* __restrict__ ip, double* u1, double* u2, double* u3) {
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long)g.ii * g.jj) return;
  int i = (int)(t % g.ii) + 1, j = (int)(t / g.ii) + 1;
  int p00 = ip[ix2(g, i, j)], pm0 = ip[ix2(g, i - 1, j)], p0m = ip[ix2(g, i, j - 1)],
      pmm = ip[ix2(g, i - 1, j - 1)];
  int u = (pm0 > 0 && p00 > 0), v = (p0m > 0 && p00 > 0), q = 0;
  if (min(min(p00, pm0), min(p0m, pmm)) > 0) q = 1;
  else if ((p00 > 0 && pmm > 0) || (pm0 > 0 && p0m > 0)) q = 1;
  u1[ix2(g, i, j)] = u; u2[ix2(g, i, j)] = v; u3[ix2(g, i, j)] = q;
}
__global__ void mask_finish(Geom g, const double* __restrict__ u1, const double* __restrict__ u2,
                            const double* __restrict__ u3, int* iu, int* iv, int* iq, int zs, int zn,
                            int zw, int ze) {
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= g.lev) return;
  int i = (int)(t % g.ldi) + 1 - g.nb, j = (int)(t / g.ldi) + 1 - g.nb;
  bool z = (zs && j <= 0) || (zn && j > g.jj) || (zw && i <= 0) || (ze && i > g.ii);
  iu[t] = z ? 0 : __double2int_rn(u1[t]);
  iv[t] = z ? 0 : __double2int_rn(u2[t]);
  iq[t] = z ? 0 : __double2int_rn(u3[t]);
}

void bigrid_dev(const std::string& depth_name) {
  C().sc.erase("_nd_nfaces_u"); C().sc.erase("_nd_nfaces_v");   // ndiff's wet-face lists follow the masks
  Ctx& c = C(); Geom& g = c.g;
  double* depth = c.dev(depth_name);
  bool lperiodi, lperiodj, larctic;
  if (g.nranks == 1) {
    LAUNCH(edge_max_kernel, 1, 256, 0, g, depth, c.d_red);
    CUDA_CHECK(cudaMemcpyAsync(c.h_red, c.d_red, 2 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
    lperiodi = c.h_red[0] > 0.0;
    larctic = c.h_red[1] > 0.0 && g.nreg == 2;
    lperiodj = c.h_red[1] > 0.0 && g.nreg != 2;
    int& nreg = g.nreg;
    if (!lperiodi && !lperiodj && (nreg == 0 || nreg == -1)) nreg = 0;
    else if (lperiodi && !lperiodj && (nreg == 1 || nreg == -1)) nreg = 1;
    else if (lperiodi && larctic && (nreg == 2 || nreg == -1)) nreg = 2;
    else if (lperiodi && lperiodj && (nreg == 3 || nreg == -1)) nreg = 3;
    else if (!lperiodi && lperiodj && (nreg == 4 || nreg == -1)) nreg = 4;
    else throw std::runtime_error("bigrid: basin depth array inconsistent with nreg");
  } else {
    if (g.nreg < 0) throw std::runtime_error("bigrid: nreg must be given for multi-tile runs");
    lperiodi = (g.nreg == 1 || g.nreg == 2 || g.nreg == 3);
    lperiodj = (g.nreg == 3 || g.nreg == 4);
    larctic = (g.nreg == 2);
    if (lperiodj) throw std::runtime_error("bigrid: j-periodic regions are single-tile only");
  }
  halo_update(depth, 1, g.nb, g.nb, halo_ps);
  const int zs = (!lperiodj && g.south), zn = (!lperiodj && !larctic && g.north);
  const int zw = !lperiodi, ze = !lperiodi;
  const int nb_all = cdiv(g.lev, 256), nb_int = cdiv((long)g.ii * g.jj, 256);
  LAUNCH(zero_frames_d, nb_all, 256, 0, g, depth, zs, zn, zw, ze);
  int* nfill = reinterpret_cast<int*>(c.d_red);
  CUDA_CHECK(cudaMemsetAsync(nfill, 0, sizeof(int), c.stream));
  LAUNCH(inlet_check, nb_int, 256, 0, g, depth, nfill);
  CUDA_CHECK(cudaMemcpyAsync(c.h_red, nfill, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  CUDA_CHECK(cudaStreamSynchronize(c.stream));
  if (*reinterpret_cast<int*>(c.h_red) > 0)
    throw std::runtime_error("bigrid: Must correct bathymetry before running BLOM");
  int *ip = c.owned_int("ip", 1), *iq = c.owned_int("iq", 1), *iu = c.owned_int("iu", 1),
      *iv = c.owned_int("iv", 1);
  double *u1 = c.owned("_bigrid_u1", 1), *u2 = c.owned("_bigrid_u2", 1), *u3 = c.owned("_bigrid_u3", 1);
  CUDA_CHECK(cudaMemsetAsync(u1, 0, sizeof(double) * g.lev, c.stream));
  CUDA_CHECK(cudaMemsetAsync(u2, 0, sizeof(double) * g.lev, c.stream));
  CUDA_CHECK(cudaMemsetAsync(u3, 0, sizeof(double) * g.lev, c.stream));
  LAUNCH(mask_p, nb_all, 256, 0, g, depth, ip, iq, iu, iv);
  LAUNCH(mask_uvq, nb_int, 256, 0, g, ip, u1, u2, u3);
  halo_update(std::vector<HaloReq>{{u1, 1, halo_us}, {u2, 1, halo_vs}, {u3, 1, halo_qs}}, g.nb, g.nb);
  LAUNCH(mask_finish, nb_all, 256, 0, g, u1, u2, u3, iu, iv, iq, zs, zn, zw, ze);
}

}  // namespace blom
