// Shared declarations of the B200-native BLOM stencil-step library.
// Layout contract: every field is the reference's
//   real(8) a(1-nbdy:idm+nbdy, 1-nbdy:jdm+nbdy [,nlev])   (phy/mod_xc.F90:45)
// column-major, i fastest.  Device copies keep exactly that layout so that the
// host (Fortran) arrays can be moved with a single cudaMemcpy.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>
#include <stdexcept>

namespace blom {

// phy/mod_constants.F90:30-56
constexpr double grav = 9.806, alpha0 = 1.e-3, rho0 = 1.e3;
constexpr double epsilpl = 1.e-14, epsilp = 1.e-12, spval = 1.e33;
constexpr double onem = 9806., onecm = 98.06, onemm = 9.806, onemu = .009806;
constexpr double tenm = 98060.;

enum { halo_ps = 1, halo_qs = 2, halo_us = 3, halo_vs = 4,
       halo_pv = 11, halo_qv = 12, halo_uv = 13, halo_vv = 14 };

// Geometry handed to every kernel by value.
struct Geom {
  int itdm, jtdm, kdm, idm, jdm, nb, ntr, nreg;
  int i0, j0, ii, jj;
  int ldi, ldj;
  long lev;         // elements per level = ldi*ldj
  int rank, nranks; // j-band decomposition
  int south, north; // 1 if this tile touches the global southern / northern edge
  int lf;           // 1: level-parallel kernels are launched with the level as the fastest block index
};

// Block order of the level-parallel kernels (one block = 128 cells of a row at one level).  With the
// plain order (i-block, j, k) a launch sweeps the whole horizontal plane once per level, so the 2-D
// operands of a kernel (masks, metrics, barotropic parts: 10-30 arrays of 13 MB each at tnx0.25v4,
// more than the 126 MB L2) are streamed from HBM again for every level.  With g.lf the grid is
// launched as (k, i-block, j): the blocks of one row segment run back to back through the levels and
// take the 2-D operands from L1/L2; only the 3-D fields stream.  Kernels read their block coordinates
// through bid(g) and launches permute the grid with lgrid(g, grid).
struct Bid { int x, y, z; };
__device__ __forceinline__ Bid bid(const Geom& g) {
  return g.lf ? Bid{(int)blockIdx.y, (int)blockIdx.z, (int)blockIdx.x}
              : Bid{(int)blockIdx.x, (int)blockIdx.y, (int)blockIdx.z};
}
inline dim3 lgrid(const Geom& g, dim3 grid) { return g.lf ? dim3(grid.z, grid.x, grid.y) : grid; }

__host__ __device__ __forceinline__ long ix2(const Geom& g, int i, int j) {
  return (long)(j + g.nb - 1) * g.ldi + (i + g.nb - 1);
}
__host__ __device__ __forceinline__ long ix3(const Geom& g, int i, int j, int k) {
  return (long)(k - 1) * g.lev + (long)(j + g.nb - 1) * g.ldi + (i + g.nb - 1);
}

struct DField {
  double* d = nullptr;   // device
  double* h = nullptr;   // host (caller-owned) or nullptr for library-owned
  int nlev = 0;
};
struct IFieldD {
  int* d = nullptr;
  int* h = nullptr;
  int nlev = 0;
};

struct Timer {
  double ms = 0; long calls = 0; long launches = 0;
};

struct Ctx {
  Geom g{};
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;   // blomgpu_download_async: D2H copies that overlap later kernels
  cudaStream_t up_stream = nullptr;     // blomgpu_upload_async: H2D copies that overlap earlier kernels
  std::map<std::string, cudaEvent_t> up_done;   // completion event of the last asynchronous upload per field
  std::map<std::string, DField> f;
  std::map<std::string, IFieldD> fi;
  std::map<std::string, std::string> opt;
  std::map<std::string, double> sc;
  long launches = 0;
  bool timers_on = false;
  // per-kernel device timers (bench.py roofline leg): one event pair per launch
  bool ktimers_on = false;
  struct KEv { const char* name; cudaEvent_t e0, e1; };
  std::vector<KEv> kev;
  std::map<std::string, Timer> ktimers;
  std::map<std::string, Timer> timers;
  std::vector<std::string> timer_order;
  void* nccl = nullptr;      // ncclComm_t
  double* d_red = nullptr;   // small device scratch for reductions
  double* h_red = nullptr;   // pinned host mirror
  size_t red_cap = 0;
  double* halo_send[2] = {nullptr, nullptr};  // [0]=to south, [1]=to north
  double* halo_recv[2] = {nullptr, nullptr};
  size_t halo_cap = 0;
  // device-raised fatal conditions (the reference's "print + xchalt"): kernels
  // atomicMax a code into this mapped pinned word; check_errors() reports it at
  // the next sync / download.
  int* err_host = nullptr;
  int* err_devptr = nullptr;
  std::string error_source;
  int* error_flag();
  void check_errors();

  double* dev(const std::string& n) const {
    auto it = f.find(n);
    if (it == f.end()) throw std::runtime_error("blomgpu: field not registered: " + n);
    return it->second.d;
  }
  bool has(const std::string& n) const { return f.count(n) != 0; }
  int nlev(const std::string& n) const {
    auto it = f.find(n);
    if (it == f.end()) throw std::runtime_error("blomgpu: field not registered: " + n);
    return it->second.nlev;
  }
  int* idev(const std::string& n) const {
    auto it = fi.find(n);
    if (it == fi.end()) throw std::runtime_error("blomgpu: int field not registered: " + n);
    return it->second.d;
  }
  // library-owned device array (routine-local `save` arrays / module-private
  // scratch of the reference); zero-filled on creation.
  double* owned(const std::string& n, int nlev);
  int* owned_int(const std::string& n, int nlev);
  double scalar(const std::string& k) const {
    auto it = sc.find(k);
    if (it == sc.end()) throw std::runtime_error("blomgpu: scalar not set: " + k);
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

Ctx& C();

#define CUDA_CHECK(x)                                                              \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess)                                                         \
      throw std::runtime_error(std::string("CUDA error ") + cudaGetErrorString(e_) + \
                               " at " __FILE__ ":" + std::to_string(__LINE__));    \
  } while (0)

// every kernel launch goes through this so gpu_launches is a real count
#define LAUNCH_NAMED(kname, kernel, grid, block, smem, ...)                 \
  do {                                                                      \
    blom::Ctx& c_ = blom::C();                                              \
    cudaEvent_t ke0_ = nullptr, ke1_ = nullptr;                             \
    if (c_.ktimers_on) {                                                    \
      cudaEventCreate(&ke0_); cudaEventCreate(&ke1_);                       \
      cudaEventRecord(ke0_, c_.stream);                                     \
    }                                                                       \
    kernel<<<(grid), (block), (smem), c_.stream>>>(__VA_ARGS__);            \
    if (c_.ktimers_on) {                                                    \
      cudaEventRecord(ke1_, c_.stream);                                     \
      c_.kev.push_back(blom::Ctx::KEv{kname, ke0_, ke1_});                  \
    }                                                                       \
    c_.launches++;                                                          \
    CUDA_CHECK(cudaGetLastError());                                         \
  } while (0)

#define LAUNCH(kernel, grid, block, smem, ...) \
  LAUNCH_NAMED(#kernel, kernel, grid, block, smem, __VA_ARGS__)

// cooperative (grid-synchronising) launch on the library stream, counted and timed like LAUNCH
inline void launch_cooperative(const char* kname, const void* kernel, int grid, int block, void** args) {
  blom::Ctx& c_ = blom::C();
  cudaEvent_t ke0_ = nullptr, ke1_ = nullptr;
  if (c_.ktimers_on) { cudaEventCreate(&ke0_); cudaEventCreate(&ke1_); cudaEventRecord(ke0_, c_.stream); }
  CUDA_CHECK(cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(block), args, 0, c_.stream));
  if (c_.ktimers_on) { cudaEventRecord(ke1_, c_.stream); c_.kev.push_back(blom::Ctx::KEv{kname, ke0_, ke1_}); }
  c_.launches++;
}

inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// Resident-blocks-per-SM bound of a kernel templated on it (second __launch_bounds__ argument).
// The stencil kernels of this path are latency-bound (ncu: long-scoreboard stalls, issue slots and
// HBM both far from saturated), and on B200 more resident warps beat more registers per thread even
// at the price of a few spilled values, so the defaults sit above the compilers' natural choice;
// `opt` is a development switch to measure the alternatives.
#define OCC_DISPATCH3(opt, dflt, VA_, VB_, VC_, ...)                                 \
  do {                                                                              \
    const int occ_ = std::stoi(blom::C().option(opt, #dflt));                       \
    if (occ_ == VA_) { constexpr int OCC = VA_; __VA_ARGS__; }                      \
    else if (occ_ == VB_) { constexpr int OCC = VB_; __VA_ARGS__; }                 \
    else { constexpr int OCC = VC_; __VA_ARGS__; }                                  \
  } while (0)

// RAII per-routine timer (device time via CUDA events on the library stream)
struct ScopedTimer {
  std::string name; cudaEvent_t e0 = nullptr, e1 = nullptr; long l0 = 0; bool on;
  explicit ScopedTimer(const char* n);
  ~ScopedTimer();
};

// ---- halo update (xc.cu) ----------------------------------------------------
struct HaloReq { double* base; int nlev; int itype; };
// xctilr semantics for each request: levels 1..nlev of `base`, same (mh,nh)
void halo_update(const std::vector<HaloReq>& reqs, int mh, int nh);
inline void halo_update(double* base, int nlev, int mh, int nh, int itype) {
  halo_update(std::vector<HaloReq>{HaloReq{base, nlev, itype}}, mh, nh);
}
// reference-exact variant with l1 (levels below l1 skip the N/S phase in the
// arctic serial code, phy/mod_xc.F90:4265,4363)
void xctilr_exact(double* base, int l1, int ld, int mh, int nh, int itype);

double xcsum_dev(const double* a, const int* mask);
double xcmax_dev(const double* a, const int* mask, bool is_max);
uint32_t xccrc_dev(const double* a, int ld, const int* mask);
void bigrid_dev(const std::string& depth);

// ---- routines ---------------------------------------------------------------
void init_cppm_dev();
void advect_dev(int m, int n, int mm, int nn, int k1m, int k1n);
void advect_remap_dev(int m, int n, int mm, int nn, int k1m, int k1n);  // remap.cu
void diffus_dev(int m, int n, int mm, int nn, int k1m, int k1n);
void tmsmt1_dev(int nn);
void tmsmt2_dev(int m, int mm, int nn, int k1m);
void inieos_dev();
void pgforc_dev(int m, int n, int mm, int nn, int k1m, int k1n);
void momtum_dev(int m, int n, int mm, int nn, int k1m, int k1n);
void barotp_dev(int m, int n, int mm, int nn, int k1m, int k1n);
void eddtra_dev(int m, int n, int mm, int nn, int k1m, int k1n);
void eddtra_isopyc_dev(int m, int n, int mm, int nn, int k1m, int k1n);  // eddtra_isopyc.cu
void pbcor1_dev(int m, int n, int mm, int nn, int k1m, int k1n);
void pbcor2_dev(int m, int n, int mm, int nn, int k1m, int k1n);
void numerical_bounds_dev();
void init_fluxes_dev(int m, int n, int mm, int nn, int k1m, int k1n);
void ndiff_dev(int m, int n, int mm, int nn, int k1m, int k1n);  // ndiff.cu
void cmnfld2_dev(int m, int n, int mm, int nn, int k1m, int k1n);             // cmnfld.cu
void cmnfld_bfsqf_ale_dev(int m, int n, int mm, int nn, int k1m, int k1n);
void cmnfld_nslope_ale_dev(int m, int n, int mm, int nn, int k1m, int k1n);
void cmnfld_nnslope_ale_dev(int m, int n, int mm, int nn, int k1m, int k1n);
double budget_init_dev();                                       // setup_ops.cu
void budget_sums_dev(int ncall, int n, int nn, double* out);

}  // namespace blom
