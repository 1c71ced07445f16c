// Context, array registry and the extern "C" boundary (include/blomgpu.h).
#include "common.cuh"
#include "../../include/blomgpu.h"
#include <cstring>

namespace blom {

void comm_release_p2p();  // comm.cu

Ctx& C() { static Ctx c; return c; }

double* Ctx::owned(const std::string& n, int nlev) {
  auto it = f.find(n);
  if (it != f.end() && it->second.nlev == nlev) return it->second.d;
  if (it != f.end() && it->second.h == nullptr && it->second.d) cudaFree(it->second.d);
  DField fd; fd.nlev = nlev; fd.h = nullptr;
  size_t bytes = sizeof(double) * (size_t)g.lev * nlev;
  CUDA_CHECK(cudaMalloc(&fd.d, bytes));
  CUDA_CHECK(cudaMemsetAsync(fd.d, 0, bytes, stream));
  f[n] = fd;
  return fd.d;
}
int* Ctx::owned_int(const std::string& n, int nlev) {
  auto it = fi.find(n);
  if (it != fi.end() && it->second.nlev == nlev) return it->second.d;
  IFieldD fd; fd.nlev = nlev; fd.h = nullptr;
  size_t bytes = sizeof(int) * (size_t)g.lev * nlev;
  CUDA_CHECK(cudaMalloc(&fd.d, bytes));
  CUDA_CHECK(cudaMemsetAsync(fd.d, 0, bytes, stream));
  fi[n] = fd;
  return fd.d;
}

int* Ctx::error_flag() {
  if (!err_host) {
    CUDA_CHECK(cudaHostAlloc(&err_host, sizeof(int), cudaHostAllocMapped));
    *err_host = 0;
    CUDA_CHECK(cudaHostGetDevicePointer(&err_devptr, err_host, 0));
  }
  return err_devptr;
}
// call after the stream has been synchronised
void Ctx::check_errors() {
  if (err_host && *err_host != 0) {
    const int code = *err_host;
    *err_host = 0;
    throw std::runtime_error("blomgpu: device-side fatal condition, code " + std::to_string(code) + " " +
                             error_source);
  }
}

ScopedTimer::ScopedTimer(const char* n) : name(n), on(C().timers_on) {
  if (!on) return;
  l0 = C().launches;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, C().stream);
}
ScopedTimer::~ScopedTimer() {
  if (!on) return;
  cudaEventRecord(e1, C().stream);
  cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  Ctx& c = C();
  if (!c.timers.count(name)) c.timer_order.push_back(name);
  Timer& t = c.timers[name];
  t.ms += ms; t.calls++; t.launches += c.launches - l0;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
}

}  // namespace blom

using namespace blom;

static char g_err[2048] = "";
namespace blom {
// comm.cu reports through the same buffer, so blomgpu_last_error() covers the blomgpu_comm_* entries
void set_last_error(const char* msg) {
  std::snprintf(g_err, sizeof g_err, "%s", msg);
  std::fprintf(stderr, "blomgpu error: %s\n", g_err);
}
}

#define GUARD(...)                                          \
  try { __VA_ARGS__; return 0; }                            \
  catch (const std::exception& e) {                         \
    std::snprintf(g_err, sizeof g_err, "%s", e.what());     \
    std::fprintf(stderr, "blomgpu error: %s\n", g_err);     \
    return 1;                                               \
  }

static void do_finalize();
static void do_init(const int* dims, const int* tile, int device) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    throw std::runtime_error("blomgpu_init: no CUDA device available (this library has no CPU fallback)");
  CUDA_CHECK(cudaSetDevice(device));
  Ctx& c = C();
  // a second init without finalize (e.g. a host that aborted a run half way) must not inherit device
  // arrays sized for the previous tile: release everything first
  if (c.d_red || !c.f.empty() || !c.fi.empty()) do_finalize();
  Geom& g = c.g;
  g.itdm = dims[0]; g.jtdm = dims[1]; g.kdm = dims[2]; g.idm = dims[3]; g.jdm = dims[4];
  g.nb = dims[5]; g.ntr = dims[6]; g.nreg = dims[7];
  g.i0 = tile[0]; g.j0 = tile[1]; g.ii = tile[2]; g.jj = tile[3];
  g.rank = tile[4]; g.nranks = tile[5];
  if (g.ii != g.idm || g.jj != g.jdm)
    throw std::runtime_error("blomgpu_init: tile extent must equal idm/jdm");
  if (g.i0 != 0 || g.ii != g.itdm)
    throw std::runtime_error("blomgpu_init: only j-band decompositions (i0=0, ii=itdm) are supported");
  g.ldi = g.idm + 2 * g.nb; g.ldj = g.jdm + 2 * g.nb;
  g.lev = (long)g.ldi * g.ldj;
  g.south = (g.j0 == 0); g.north = (g.j0 + g.jj == g.jtdm);
  g.lf = 1;
  c.device = device;
  if (!c.stream) CUDA_CHECK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  c.red_cap = 1 << 16;
  CUDA_CHECK(cudaMalloc(&c.d_red, c.red_cap * sizeof(double)));
  CUDA_CHECK(cudaMallocHost(&c.h_red, c.red_cap * sizeof(double)));
}

static void do_finalize() {
  Ctx& c = C();
  if (c.stream) cudaStreamSynchronize(c.stream);
  comm_release_p2p();
  for (auto& kv : c.f) if (kv.second.d) cudaFree(kv.second.d);
  for (auto& kv : c.fi) if (kv.second.d) cudaFree(kv.second.d);
  if (c.d_red) cudaFree(c.d_red);
  if (c.h_red) cudaFreeHost(c.h_red);
  if (c.err_host) cudaFreeHost(c.err_host);
  for (int s = 0; s < 2; ++s) {
    if (c.halo_send[s]) cudaFree(c.halo_send[s]);
    if (c.halo_recv[s]) cudaFree(c.halo_recv[s]);
  }
  if (c.copy_stream) cudaStreamSynchronize(c.copy_stream);
  if (c.up_stream) cudaStreamSynchronize(c.up_stream);
  for (auto& kv : c.up_done) if (kv.second) cudaEventDestroy(kv.second);
  cudaStream_t st = c.stream, cst = c.copy_stream, ust = c.up_stream;
  void* nccl = c.nccl;
  c = Ctx();
  c.stream = st; c.copy_stream = cst; c.up_stream = ust;
  c.nccl = nccl;
}

static void do_register(const char* name, double* host, int nlev) {
  Ctx& c = C();
  auto it = c.f.find(name);
  if (it != c.f.end() && it->second.nlev == nlev) { it->second.h = host; return; }
  if (it != c.f.end() && it->second.d) cudaFree(it->second.d);
  DField fd; fd.h = host; fd.nlev = nlev;
  CUDA_CHECK(cudaMalloc(&fd.d, sizeof(double) * (size_t)c.g.lev * nlev));
  c.f[name] = fd;
}
static void do_register_int(const char* name, int* host, int nlev) {
  Ctx& c = C();
  auto it = c.fi.find(name);
  if (it != c.fi.end() && it->second.nlev == nlev) { it->second.h = host; return; }
  if (it != c.fi.end() && it->second.d) cudaFree(it->second.d);
  IFieldD fd; fd.h = host; fd.nlev = nlev;
  CUDA_CHECK(cudaMalloc(&fd.d, sizeof(int) * (size_t)c.g.lev * nlev));
  c.fi[name] = fd;
}

static void do_copy(const char* name, bool up) {
  Ctx& c = C();
  auto it = c.f.find(name);
  if (it != c.f.end()) {
    DField& fd = it->second;
    if (!fd.h) throw std::runtime_error(std::string("blomgpu: no host array bound to ") + name);
    size_t bytes = sizeof(double) * (size_t)c.g.lev * fd.nlev;
    if (up) CUDA_CHECK(cudaMemcpyAsync(fd.d, fd.h, bytes, cudaMemcpyHostToDevice, c.stream));
    else CUDA_CHECK(cudaMemcpyAsync(fd.h, fd.d, bytes, cudaMemcpyDeviceToHost, c.stream));
    return;
  }
  auto jt = c.fi.find(name);
  if (jt == c.fi.end()) throw std::runtime_error(std::string("blomgpu: field not registered: ") + name);
  IFieldD& fd = jt->second;
  if (!fd.h) throw std::runtime_error(std::string("blomgpu: no host array bound to ") + name);
  size_t bytes = sizeof(int) * (size_t)c.g.lev * fd.nlev;
  if (up) CUDA_CHECK(cudaMemcpyAsync(fd.d, fd.h, bytes, cudaMemcpyHostToDevice, c.stream));
  else CUDA_CHECK(cudaMemcpyAsync(fd.h, fd.d, bytes, cudaMemcpyDeviceToHost, c.stream));
}
// Device -> host copy of a field that the remaining routines of the step no longer write, on a
// second stream: it starts when everything enqueued so far on the library stream has finished and
// overlaps the kernels enqueued afterwards.  blomgpu_sync() waits for it.
static void do_download_async(const char* name) {
  Ctx& c = C();
  auto it = c.f.find(name);
  if (it == c.f.end()) throw std::runtime_error(std::string("blomgpu: field not registered: ") + name);
  DField& fd = it->second;
  if (!fd.h) throw std::runtime_error(std::string("blomgpu: no host array bound to ") + name);
  if (!c.copy_stream) CUDA_CHECK(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
  cudaEvent_t ev;
  CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventRecord(ev, c.stream));
  CUDA_CHECK(cudaStreamWaitEvent(c.copy_stream, ev, 0));
  CUDA_CHECK(cudaEventDestroy(ev));   // released once the wait has been satisfied
  CUDA_CHECK(cudaMemcpyAsync(fd.h, fd.d, sizeof(double) * (size_t)c.g.lev * fd.nlev, cudaMemcpyDeviceToHost,
                             c.copy_stream));
}
// Level range [koff, koff+nlev) (1-based) of a double field, checked against the registration.
static DField& field_range(const char* name, int koff, int nlev, size_t* off, size_t* bytes) {
  Ctx& c = C();
  auto it = c.f.find(name);
  if (it == c.f.end()) throw std::runtime_error(std::string("blomgpu: field not registered: ") + name);
  DField& fd = it->second;
  if (!fd.h) throw std::runtime_error(std::string("blomgpu: no host array bound to ") + name);
  if (koff < 1 || nlev < 1 || koff - 1 + nlev > fd.nlev)
    throw std::runtime_error(std::string("blomgpu: level range outside array: ") + name);
  *off = (size_t)(koff - 1) * c.g.lev;
  *bytes = sizeof(double) * (size_t)c.g.lev * nlev;
  return fd;
}
static void stream_after(cudaStream_t waiter, cudaStream_t producer) {
  cudaEvent_t ev;
  CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventRecord(ev, producer));
  CUDA_CHECK(cudaStreamWaitEvent(waiter, ev, 0));
  CUDA_CHECK(cudaEventDestroy(ev));   // released once the wait has been satisfied
}
// Host -> device copy of a level range on the upload stream.  It starts when everything enqueued so far
// (kernels and downloads that may still read the old content) has finished and overlaps the routines
// called afterwards; a routine that reads the field must be preceded by blomgpu_wait_upload(name).
static void do_upload_async(const char* name, int koff, int nlev) {
  Ctx& c = C();
  size_t off, bytes;
  DField& fd = field_range(name, koff, nlev, &off, &bytes);
  if (!c.up_stream) CUDA_CHECK(cudaStreamCreateWithFlags(&c.up_stream, cudaStreamNonBlocking));
  stream_after(c.up_stream, c.stream);
  if (c.copy_stream) stream_after(c.up_stream, c.copy_stream);
  CUDA_CHECK(cudaMemcpyAsync(fd.d + off, fd.h + off, bytes, cudaMemcpyHostToDevice, c.up_stream));
  cudaEvent_t& ev = c.up_done[name];
  if (!ev) CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventRecord(ev, c.up_stream));
}
static void do_wait_upload(const char* name) {
  Ctx& c = C();
  auto it = c.up_done.find(name);
  if (it != c.up_done.end() && it->second) CUDA_CHECK(cudaStreamWaitEvent(c.stream, it->second, 0));
}
static void do_download_levels_async(const char* name, int koff, int nlev) {
  Ctx& c = C();
  size_t off, bytes;
  DField& fd = field_range(name, koff, nlev, &off, &bytes);
  if (!c.copy_stream) CUDA_CHECK(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
  stream_after(c.copy_stream, c.stream);
  CUDA_CHECK(cudaMemcpyAsync(fd.h + off, fd.d + off, bytes, cudaMemcpyDeviceToHost, c.copy_stream));
}
static void do_sync() {
  Ctx& c = C();
  if (c.up_stream) CUDA_CHECK(cudaStreamSynchronize(c.up_stream));
  CUDA_CHECK(cudaStreamSynchronize(c.stream));
  if (c.copy_stream) CUDA_CHECK(cudaStreamSynchronize(c.copy_stream));
  c.check_errors();
}
static void do_copy_all(bool up) {
  Ctx& c = C();
  for (auto& kv : c.f) if (kv.second.h) do_copy(kv.first.c_str(), up);
  for (auto& kv : c.fi) if (kv.second.h) do_copy(kv.first.c_str(), up);
  CUDA_CHECK(cudaStreamSynchronize(c.stream));
  c.check_errors();
}

static const int* mask_for_itype(int itype) {
  switch (itype) {
    case halo_ps: case halo_pv: return C().idev("ip");
    case halo_qs: case halo_qv: return C().idev("iq");
    case halo_us: case halo_uv: return C().idev("iu");
    case halo_vs: case halo_vv: return C().idev("iv");
    default: throw std::runtime_error(" chksum: itype is unsupported!");
  }
}

// Halo refreshes that difest_lateral_hybrid / difest_isobml (phy/mod_difest.F90:826-831) and cmnfld2
// (phy/mod_cmnfld_routines.F90:1171-1172) issue between tmsmt1 and eddtra.  Those routines are out of scope
// (column physics), but momtum and eddtra rely on the halo validity they leave behind (SURVEY.md appendix A,
// validity chain 1 and 6), so a host that keeps the state device-resident calls this entry in their place.
// Two batched launches instead of eight xctilr calls; identical result (the fields are independent).
static void difest_halos_dev() {
  Ctx& c = C(); const int kk = c.g.kdm;
  halo_update(std::vector<HaloReq>{{c.dev("u"), 2 * kk, halo_uv}, {c.dev("v"), 2 * kk, halo_vv},
                                   {c.dev("ubflxs_p"), 2, halo_uv}, {c.dev("vbflxs_p"), 2, halo_vv},
                                   {c.dev("pbu"), 2, halo_us}, {c.dev("pbv"), 2, halo_vs}}, 2, 2);
  halo_update(std::vector<HaloReq>{{c.dev("temp"), 2 * kk, halo_ps}, {c.dev("saln"), 2 * kk, halo_ps}}, 3, 3);
}

static int not_impl(const char* what) {
  std::snprintf(g_err, sizeof g_err, "blomgpu: %s is not implemented in this build", what);
  std::fprintf(stderr, "%s\n", g_err);
  return 2;
}

// External-linkage C++ function behind blomgpu_parity_build(): if another flavour of this library
// were allowed to interpose the internal calls, the entry point would report the other flavour
// (tests/test_abi.py::test_flavours_are_isolated).
int build_flavour() {
#ifdef BLOM_PARITY_BUILD
  return 1;
#else
  return 0;
#endif
}

extern "C" {

const char* blomgpu_last_error(void) { return g_err; }
int blomgpu_parity_build(void) { return build_flavour(); }
int blomgpu_init(const int dims[8], const int tile[6], int device) { GUARD(do_init(dims, tile, device)) }
int blomgpu_finalize(void) { GUARD(do_finalize()) }

int blomgpu_register(const char* name, double* host, int nlev) { GUARD(do_register(name, host, nlev)) }
int blomgpu_register_int(const char* name, int* host, int nlev) { GUARD(do_register_int(name, host, nlev)) }
int blomgpu_upload(const char* name) { GUARD(do_copy(name, true)) }
int blomgpu_download(const char* name) { GUARD(do_copy(name, false); CUDA_CHECK(cudaStreamSynchronize(C().stream)); C().check_errors()) }
int blomgpu_upload_all(void) { GUARD(do_copy_all(true)) }
int blomgpu_download_all(void) { GUARD(do_copy_all(false)) }
int blomgpu_download_async(const char* name) { GUARD(do_download_async(name)) }
int blomgpu_download_levels_async(const char* name, int koff, int nlev) { GUARD(do_download_levels_async(name, koff, nlev)) }
int blomgpu_upload_async(const char* name, int koff, int nlev) { GUARD(do_upload_async(name, koff, nlev)) }
int blomgpu_wait_upload(const char* name) { GUARD(do_wait_upload(name)) }
int blomgpu_sync(void) { GUARD(do_sync()) }
int blomgpu_device_ptr(const char* name, void** dptr, int* nlev) {
  GUARD(
    Ctx& c = C();
    auto it = c.f.find(name);
    if (it != c.f.end()) { *dptr = it->second.d; if (nlev) *nlev = it->second.nlev; }
    else {
      auto jt = c.fi.find(name);
      if (jt == c.fi.end()) throw std::runtime_error(std::string("blomgpu: field not registered: ") + name);
      *dptr = jt->second.d; if (nlev) *nlev = jt->second.nlev;
    })
}

int blomgpu_set_option(const char* key, const char* value) {
  GUARD(
    C().opt[key] = value;
    // block order of the level-parallel kernels: "level" (default, level fastest) or "plane"
    if (std::string(key) == "grid_order") C().g.lf = std::string(value) == "plane" ? 0 : 1)
}
int blomgpu_set_scalar(const char* key, double value) { GUARD(C().sc[key] = value) }
int blomgpu_get_scalar(const char* key, double* value) { GUARD(*value = C().scalar(key)) }

int blomgpu_xctilr(const char* name, int koff, int l1, int ld, int mh, int nh, int itype) {
  GUARD(
    Ctx& c = C();
    if (koff < 1 || koff - 1 + ld > c.nlev(name)) throw std::runtime_error("blomgpu_xctilr: level range outside array");
    ScopedTimer t("xctilr");
    xctilr_exact(c.dev(name) + (size_t)(koff - 1) * c.g.lev, l1, ld, mh, nh, itype))
}
int blomgpu_xcsum(const char* name, int lev, const char* mask, double* sum) {
  GUARD(Ctx& c = C(); *sum = xcsum_dev(c.dev(name) + (size_t)(lev - 1) * c.g.lev, c.idev(mask)))
}
int blomgpu_xcmax(const char* name, int lev, const char* mask, double* out) {
  GUARD(Ctx& c = C(); *out = xcmax_dev(c.dev(name) + (size_t)(lev - 1) * c.g.lev, c.idev(mask), true))
}
int blomgpu_xcmin(const char* name, int lev, const char* mask, double* out) {
  GUARD(Ctx& c = C(); *out = xcmax_dev(c.dev(name) + (size_t)(lev - 1) * c.g.lev, c.idev(mask), false))
}
int blomgpu_chksum_at(const char* name, int koff, int kcsd, int itype, uint32_t* crc) {
  GUARD(
    Ctx& c = C();
    if (koff < 1 || koff - 1 + kcsd > c.nlev(name)) throw std::runtime_error("blomgpu_chksum_at: level range outside array");
    *crc = xccrc_dev(c.dev(name) + (size_t)(koff - 1) * c.g.lev, kcsd, mask_for_itype(itype)))
}
int blomgpu_chksum(const char* name, int kcsd, int itype, uint32_t* crc) {
  GUARD(Ctx& c = C(); *crc = xccrc_dev(c.dev(name), kcsd, mask_for_itype(itype)))
}

int blomgpu_bigrid(const char* depth_name) { GUARD(bigrid_dev(depth_name)) }
int blomgpu_nreg(void) { return C().g.nreg; }
int blomgpu_init_cppm(void) { GUARD(init_cppm_dev()) }
int blomgpu_inieos(void) { GUARD(inieos_dev()) }
int blomgpu_numerical_bounds(void) { GUARD(numerical_bounds_dev()) }
int blomgpu_init_fluxes(int m, int n, int mm, int nn, int k1m, int k1n) { GUARD(init_fluxes_dev(m, n, mm, nn, k1m, k1n)) }

int blomgpu_difest_halos(int m, int n, int mm, int nn, int k1m, int k1n) {
  (void)m; (void)n; (void)mm; (void)nn; (void)k1m; (void)k1n;
  GUARD(ScopedTimer t("difest_halos"); difest_halos_dev())
}
int blomgpu_tmsmt1(int nn) { GUARD(ScopedTimer t("tmsmt1"); tmsmt1_dev(nn)) }
int blomgpu_eddtra(int m, int n, int mm, int nn, int k1m, int k1n) { GUARD(ScopedTimer t("eddtra"); eddtra_dev(m, n, mm, nn, k1m, k1n)) }
int blomgpu_advect(int m, int n, int mm, int nn, int k1m, int k1n) { GUARD(ScopedTimer t("advect"); advect_dev(m, n, mm, nn, k1m, k1n)) }
int blomgpu_pbcor1(int m, int n, int mm, int nn, int k1m, int k1n) { GUARD(ScopedTimer t("pbcor1"); pbcor1_dev(m, n, mm, nn, k1m, k1n)) }
int blomgpu_diffus(int m, int n, int mm, int nn, int k1m, int k1n) { GUARD(ScopedTimer t("diffus"); diffus_dev(m, n, mm, nn, k1m, k1n)) }
int blomgpu_pgforc(int m, int n, int mm, int nn, int k1m, int k1n) { GUARD(ScopedTimer t("pgforc"); pgforc_dev(m, n, mm, nn, k1m, k1n)) }
int blomgpu_momtum(int m, int n, int mm, int nn, int k1m, int k1n) { GUARD(ScopedTimer t("momtum"); momtum_dev(m, n, mm, nn, k1m, k1n)) }
int blomgpu_barotp(int m, int n, int mm, int nn, int k1m, int k1n) { GUARD(ScopedTimer t("barotp"); barotp_dev(m, n, mm, nn, k1m, k1n)) }
int blomgpu_pbcor2(int m, int n, int mm, int nn, int k1m, int k1n) { GUARD(ScopedTimer t("pbcor2"); pbcor2_dev(m, n, mm, nn, k1m, k1n)) }
int blomgpu_ndiff(int m, int n, int mm, int nn, int k1m, int k1n) { GUARD(ScopedTimer t("ndiff"); ndiff_dev(m, n, mm, nn, k1m, k1n)) }
int blomgpu_cmnfld2(int m, int n, int mm, int nn, int k1m, int k1n) { GUARD(ScopedTimer t("cmnfld2"); cmnfld2_dev(m, n, mm, nn, k1m, k1n)) }
int blomgpu_cmnfld_bfsqf_ale(int m, int n, int mm, int nn, int k1m, int k1n) { GUARD(cmnfld_bfsqf_ale_dev(m, n, mm, nn, k1m, k1n)) }
int blomgpu_cmnfld_nslope_ale(int m, int n, int mm, int nn, int k1m, int k1n) { GUARD(cmnfld_nslope_ale_dev(m, n, mm, nn, k1m, k1n)) }
int blomgpu_cmnfld_nnslope_ale(int m, int n, int mm, int nn, int k1m, int k1n) { GUARD(cmnfld_nnslope_ale_dev(m, n, mm, nn, k1m, k1n)) }
int blomgpu_budget_init(double* mass0) { GUARD(*mass0 = budget_init_dev()) }
int blomgpu_budget_sums(int ncall, int n, int nn, double out[4]) { GUARD(budget_sums_dev(ncall, n, nn, out)) }
int blomgpu_tmsmt2(int m, int mm, int nn, int k1m) { GUARD(ScopedTimer t("tmsmt2"); tmsmt2_dev(m, mm, nn, k1m)) }

long blomgpu_launch_count(void) { return C().launches; }
void blomgpu_launch_count_reset(void) { C().launches = 0; }
int blomgpu_timers_enable(int enable) { C().timers_on = enable != 0; return 0; }
int blomgpu_timers_get(int cap, char names[][32], double* ms_total, long* calls, long* launches) {
  Ctx& c = C();
  int n = 0;
  for (auto& nm : c.timer_order) {
    if (n >= cap) break;
    Timer& t = c.timers[nm];
    std::snprintf(names[n], 32, "%s", nm.c_str());
    ms_total[n] = t.ms; calls[n] = t.calls; launches[n] = t.launches;
    ++n;
  }
  return n;
}
void blomgpu_timers_reset(void) { C().timers.clear(); C().timer_order.clear(); }
int blomgpu_ktimers_enable(int enable) {
  Ctx& c = C();
  c.ktimers_on = enable != 0;
  if (enable) { c.ktimers.clear(); }
  return 0;
}
int blomgpu_ktimers_get(int cap, char names[][64], double* ms_total, long* launches) {
  Ctx& c = C();
  cudaStreamSynchronize(c.stream);
  for (auto& e : c.kev) {
    float ms = 0; cudaEventElapsedTime(&ms, e.e0, e.e1);
    Timer& t = c.ktimers[e.name];
    t.ms += ms; t.launches++;
    cudaEventDestroy(e.e0); cudaEventDestroy(e.e1);
  }
  c.kev.clear();
  int n = 0;
  for (auto& kv : c.ktimers) {
    if (n >= cap) break;
    std::snprintf(names[n], 64, "%s", kv.first.c_str());
    ms_total[n] = kv.second.ms; launches[n] = kv.second.launches;
    ++n;
  }
  return n;
}
void* blomgpu_stream(void) { return (void*)C().stream; }

}  // extern "C"
