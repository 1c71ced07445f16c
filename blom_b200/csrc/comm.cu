// Multi-GPU plumbing for j-band tiles: replaces the MPI half of mod_xc
// (phy/mod_xc.F90:2954-3188 xctilr_nonarctic N/S exchange, :2071-2192 xcsum
// gather, :1157-1201 xcmax allreduce) with NCCL point-to-point / collectives on
// the library stream over NVLink.  NCCL is dlopen'ed at comm_init time so a
// single-GPU run has no NCCL dependency.
#include "common.cuh"
#include "../../include/blomgpu.h"
#include <nccl.h>
#include <dlfcn.h>
#include <cstring>

namespace blom {

namespace {
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
} N;

void load_nccl() {
  if (N.h) return;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) { N.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (N.h) break; }
  if (!N.h) throw std::runtime_error(std::string("blomgpu: cannot dlopen libnccl: ") + dlerror());
#define SYM(f) *(void**)(&N.f) = dlsym(N.h, "nccl" #f); \
  if (!N.f) throw std::runtime_error("blomgpu: libnccl lacks nccl" #f);
  SYM(GetUniqueId) SYM(CommInitRank) SYM(Send) SYM(Recv) SYM(AllReduce) SYM(GroupStart)
  SYM(GroupEnd) SYM(GetErrorString)
#undef SYM
}
#define NCCL_CHECK(x)                                                                   \
  do {                                                                                  \
    ncclResult_t r_ = (x);                                                              \
    if (r_ != ncclSuccess)                                                              \
      throw std::runtime_error(std::string("NCCL error ") + N.GetErrorString(r_) +      \
                               " at " __FILE__ ":" + std::to_string(__LINE__));         \
  } while (0)

constexpr int XR_MAX = 12;
struct XBatch { double* base[XR_MAX]; int nlev[XR_MAX]; long off[XR_MAX]; int n; };

// pack rows [j_first, j_first+nhl) x i=1..ii of every level of every request
// into buf ([req][k][r][i]); unpack does the inverse into the target rows.
__global__ void pack_rows(Geom g, XBatch b, int nhl, int j_first, double* __restrict__ buf, int unpack) {
  const int r = blockIdx.z, k = blockIdx.y;
  if (r >= b.n || k >= b.nlev[r]) return;
  double* a = b.base[r] + (long)k * g.lev;
  double* q = buf + b.off[r] + (long)k * nhl * g.ii;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < (long)nhl * g.ii;
       t += (long)gridDim.x * blockDim.x) {
    int rr = (int)(t / g.ii), i = (int)(t % g.ii) + 1;
    long x = ix2(g, i, j_first + rr);
    if (unpack) a[x] = q[t]; else q[t] = a[x];
  }
}
}  // namespace

// Fill the nhl halo rows on band edges shared with a neighbouring GPU.
void exchange_ns(const std::vector<HaloReq>& reqs, int nhl) {
  Ctx& c = C(); const Geom& g = c.g;
  if (!c.nccl) throw std::runtime_error("blomgpu: multi-tile halo update before blomgpu_comm_init");
  if (g.nreg >= 3) throw std::runtime_error("blomgpu: j-periodic regions are single-tile only");
  ncclComm_t comm = (ncclComm_t)c.nccl;
  for (size_t s = 0; s < reqs.size(); s += XR_MAX) {
    XBatch b{}; b.n = (int)std::min((size_t)XR_MAX, reqs.size() - s);
    long tot = 0; int maxlev = 0;
    for (int r = 0; r < b.n; ++r) {
      b.base[r] = reqs[s + r].base; b.nlev[r] = reqs[s + r].nlev; b.off[r] = tot;
      tot += (long)reqs[s + r].nlev * nhl * g.ii;
      maxlev = std::max(maxlev, b.nlev[r]);
    }
    if ((size_t)tot > c.halo_cap) {
      CUDA_CHECK(cudaStreamSynchronize(c.stream));
      for (int q = 0; q < 2; ++q) {
        if (c.halo_send[q]) cudaFree(c.halo_send[q]);
        if (c.halo_recv[q]) cudaFree(c.halo_recv[q]);
        CUDA_CHECK(cudaMalloc(&c.halo_send[q], sizeof(double) * tot));
        CUDA_CHECK(cudaMalloc(&c.halo_recv[q], sizeof(double) * tot));
      }
      c.halo_cap = tot;
    }
    const bool has_s = g.rank > 0, has_n = g.rank + 1 < g.nranks;
    dim3 grid(std::max(1, std::min(cdiv((long)nhl * g.ii, 256), 32)), maxlev, b.n);
    if (has_s) LAUNCH(pack_rows, grid, 256, 0, g, b, nhl, 1, c.halo_send[0], 0);
    if (has_n) LAUNCH(pack_rows, grid, 256, 0, g, b, nhl, g.jj - nhl + 1, c.halo_send[1], 0);
    NCCL_CHECK(N.GroupStart());
    if (has_s) {
      NCCL_CHECK(N.Send(c.halo_send[0], tot, ncclDouble, g.rank - 1, comm, c.stream));
      NCCL_CHECK(N.Recv(c.halo_recv[0], tot, ncclDouble, g.rank - 1, comm, c.stream));
    }
    if (has_n) {
      NCCL_CHECK(N.Send(c.halo_send[1], tot, ncclDouble, g.rank + 1, comm, c.stream));
      NCCL_CHECK(N.Recv(c.halo_recv[1], tot, ncclDouble, g.rank + 1, comm, c.stream));
    }
    NCCL_CHECK(N.GroupEnd());
    c.launches += (has_s ? 1 : 0) + (has_n ? 1 : 0);
    if (has_s) LAUNCH(pack_rows, grid, 256, 0, g, b, nhl, 1 - nhl, c.halo_recv[0], 1);
    if (has_n) LAUNCH(pack_rows, grid, 256, 0, g, b, nhl, g.jj + 1, c.halo_recv[1], 1);
  }
}

// Row partials of all bands in global row order on every rank (zero-padded
// all-reduce: x + 0.0 == x exactly, so the values are untouched).
double* gather_rows(double* rows_local, int jj_local, int* n_total) {
  Ctx& c = C(); const Geom& g = c.g;
  if (!c.nccl) throw std::runtime_error("blomgpu: reduction before blomgpu_comm_init");
  double* all = c.owned("_xcsum_rows", 1);  // lev >= jtdm always
  CUDA_CHECK(cudaMemsetAsync(all, 0, sizeof(double) * g.jtdm, c.stream));
  CUDA_CHECK(cudaMemcpyAsync(all + g.j0, rows_local, sizeof(double) * jj_local, cudaMemcpyDeviceToDevice, c.stream));
  NCCL_CHECK(N.AllReduce(all, all, g.jtdm, ncclDouble, ncclSum, (ncclComm_t)c.nccl, c.stream));
  c.launches++;
  *n_total = g.jtdm;
  return all;
}
uint32_t* gather_rows_u32(uint32_t* rows_local, int jj_local, int* n_total) {
  Ctx& c = C(); const Geom& g = c.g;
  if (!c.nccl) throw std::runtime_error("blomgpu: reduction before blomgpu_comm_init");
  uint32_t* all = reinterpret_cast<uint32_t*>(c.owned("_xccrc_rows", 1));
  CUDA_CHECK(cudaMemsetAsync(all, 0, sizeof(uint32_t) * g.jtdm, c.stream));
  CUDA_CHECK(cudaMemcpyAsync(all + g.j0, rows_local, sizeof(uint32_t) * jj_local, cudaMemcpyDeviceToDevice, c.stream));
  NCCL_CHECK(N.AllReduce(all, all, g.jtdm, ncclUint32, ncclSum, (ncclComm_t)c.nccl, c.stream));
  c.launches++;
  *n_total = g.jtdm;
  return all;
}
void allreduce_minmax(double* d_val, bool is_max) {
  Ctx& c = C();
  if (!c.nccl) throw std::runtime_error("blomgpu: reduction before blomgpu_comm_init");
  NCCL_CHECK(N.AllReduce(d_val, d_val, 1, ncclDouble, is_max ? ncclMax : ncclMin, (ncclComm_t)c.nccl, c.stream));
  c.launches++;
}

}  // namespace blom

using namespace blom;
static char g_cerr[512];
extern "C" {
int blomgpu_comm_unique_id(char id[128]) {
  try {
    load_nccl();
    ncclUniqueId u;
    NCCL_CHECK(N.GetUniqueId(&u));
    static_assert(sizeof(u) == 128, "ncclUniqueId size");
    std::memcpy(id, &u, 128);
    return 0;
  } catch (const std::exception& e) { std::fprintf(stderr, "blomgpu error: %s\n", e.what()); return 1; }
}
int blomgpu_comm_init(const char id[128], int rank, int nranks) {
  try {
    load_nccl();
    ncclUniqueId u; std::memcpy(&u, id, 128);
    ncclComm_t comm;
    NCCL_CHECK(N.CommInitRank(&comm, nranks, u, rank));
    C().nccl = comm;
    return 0;
  } catch (const std::exception& e) { std::fprintf(stderr, "blomgpu error: %s\n", e.what()); return 1; }
}
}
