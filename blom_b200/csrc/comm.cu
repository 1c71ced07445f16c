// Multi-GPU plumbing for j-band tiles: replaces the MPI half of mod_xc
// (phy/mod_xc.F90:2954-3188 xctilr_nonarctic N/S exchange, :2071-2192 xcsum
// gather, :1157-1201 xcmax allreduce).  Band edges go through peer-memory mailboxes
// over NVLink (CUDA IPC; p2p_push / p2p_unpack below and the in-kernel exchange of
// barotp.cu) when peer access exists, else through NCCL send/recv groups on the
// library stream; the reductions use NCCL collectives.  NCCL is dlopen'ed at
// comm_init time so a single-GPU run has no NCCL dependency.
#include "common.cuh"
#include "p2p.cuh"
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
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
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
  SYM(GetUniqueId) SYM(CommInitRank) SYM(Send) SYM(Recv) SYM(AllReduce) SYM(AllGather) SYM(GroupStart)
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

// ---------------------------------------------------------------------------
// Peer-to-peer band-edge exchange over NVLink (CUDA IPC): every rank owns a
// mailbox; a neighbour packs its edge rows STRAIGHT into that mailbox with
// ordinary stores on the mapped peer pointer, publishes a sequence number in
// the mailbox's flag word (system-scope fence + store), and the owner's next
// kernel spins on the flag before unpacking.  One pack+publish launch and one
// wait+unpack launch per exchange, no host round trip, no NCCL kernels.  Two
// parity slots per direction are enough: a rank can only run one exchange
// ahead of its neighbour (its wait for exchange n+1 completes after the
// neighbour's push n+1, which follows the neighbour's unpack n in stream order).
// Falls back to the NCCL path when IPC peer mapping is unavailable
// (option comm=nccl forces it).
// ---------------------------------------------------------------------------
struct P2P {
  bool tried = false, on = false;
  char* block = nullptr;                 // [flags+counters 256 B][2 dirs][2 parity][cap] doubles
  size_t cap = 0;                        // doubles per slot
  char* peer[2] = {nullptr, nullptr};    // mapped blocks of the south / north neighbour
  unsigned long long seq = 0;
  unsigned long long done_target[2] = {0, 0};   // cumulative block count of the pushes per direction
};
P2P g_p2p;

// Two-phase teardown: every rank first closes its mappings of the neighbours' mailboxes, then all ranks
// meet in an NCCL barrier, and only then is the exported block freed (freeing IPC-exported memory that
// an importer still maps is undefined).  Collective whenever the mailboxes are in use (q.on is agreed
// by all ranks in p2p_setup), which holds for the capacity-growth path and for blomgpu_finalize.
void p2p_close() {
  P2P& q = g_p2p;
  Ctx& c = C();
  for (int d = 0; d < 2; ++d) if (q.peer[d]) { cudaIpcCloseMemHandle(q.peer[d]); q.peer[d] = nullptr; }
  if (q.on && c.nccl && c.g.nranks > 1 && c.stream) {
    int* d_b = nullptr;
    if (cudaMalloc(&d_b, sizeof(int)) == cudaSuccess) {
      cudaMemsetAsync(d_b, 0, sizeof(int), c.stream);
      N.AllReduce(d_b, d_b, 1, ncclInt, ncclSum, (ncclComm_t)c.nccl, c.stream);
      cudaStreamSynchronize(c.stream);
      cudaFree(d_b);
    }
  }
  if (q.block) { cudaFree(q.block); q.block = nullptr; }
  q.cap = 0; q.on = false;
}

// (re)allocate the mailbox for `cap` doubles per slot and map the neighbours' mailboxes.  Collective:
// every rank issues the same sequence of exchanges, so all ranks arrive here together.
bool p2p_setup(size_t cap) {
  Ctx& c = C(); const Geom& g = c.g;
  P2P& q = g_p2p;
  ncclComm_t comm = (ncclComm_t)c.nccl;
  CUDA_CHECK(cudaStreamSynchronize(c.stream));
  p2p_close();
  const size_t bytes = P2P_HDR + 4 * cap * sizeof(double);
  CUDA_CHECK(cudaMalloc(&q.block, bytes));
  CUDA_CHECK(cudaMemset(q.block, 0, P2P_HDR));
  cudaIpcMemHandle_t mine;
  bool ok = cudaIpcGetMemHandle(&mine, q.block) == cudaSuccess;
  // all-gather the handles (and an ok byte) through NCCL
  const size_t rec = sizeof(cudaIpcMemHandle_t) + 8;
  std::vector<char> h_all(rec * g.nranks, 0), h_mine(rec, 0);
  std::memcpy(h_mine.data(), &mine, sizeof mine);
  h_mine[sizeof mine] = ok ? 1 : 0;
  char *d_mine = nullptr, *d_all = nullptr;
  CUDA_CHECK(cudaMalloc(&d_mine, rec)); CUDA_CHECK(cudaMalloc(&d_all, rec * g.nranks));
  CUDA_CHECK(cudaMemcpy(d_mine, h_mine.data(), rec, cudaMemcpyHostToDevice));
  NCCL_CHECK(N.AllGather(d_mine, d_all, rec, ncclChar, comm, c.stream));
  CUDA_CHECK(cudaStreamSynchronize(c.stream));
  CUDA_CHECK(cudaMemcpy(h_all.data(), d_all, rec * g.nranks, cudaMemcpyDeviceToHost));
  cudaFree(d_mine); cudaFree(d_all);
  for (int r = 0; r < g.nranks; ++r) ok = ok && h_all[r * rec + sizeof mine] == 1;
  if (ok) {
    const int nb[2] = {g.rank - 1, g.rank + 1};
    for (int d = 0; d < 2 && ok; ++d) {
      if (nb[d] < 0 || nb[d] >= g.nranks) continue;
      cudaIpcMemHandle_t h; std::memcpy(&h, h_all.data() + nb[d] * rec, sizeof h);
      void* ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; cudaGetLastError(); }
      q.peer[d] = static_cast<char*>(ptr);
    }
  }
  // agree on the outcome (a rank that could not map a neighbour takes everybody to the NCCL path)
  int* d_ok = nullptr; CUDA_CHECK(cudaMalloc(&d_ok, sizeof(int)));
  int h_ok = ok ? 1 : 0;
  CUDA_CHECK(cudaMemcpy(d_ok, &h_ok, sizeof(int), cudaMemcpyHostToDevice));
  NCCL_CHECK(N.AllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, comm, c.stream));
  CUDA_CHECK(cudaStreamSynchronize(c.stream));
  CUDA_CHECK(cudaMemcpy(&h_ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost));
  cudaFree(d_ok);
  q.seq = 0; q.done_target[0] = q.done_target[1] = 0;
  if (!h_ok) { p2p_close(); return false; }
  q.cap = cap; q.on = true;
  return true;
}

// pack the edge rows of every request into the neighbours' mailboxes; per direction the last block to finish
// publishes `seq` in that neighbour's flag word.  One launch serves both neighbours: blockIdx.z = dir * b.n +
// request, blockIdx.y = level, gridDim.x blocks per level; dir 0: rows 1..nhl go south, dir 1: rows
// jj-nhl+1..jj go north.  Blocks of a direction without neighbour return at once and are not counted.
__global__ void p2p_push(Geom g, XBatch b, int nhl, char* my_block, char* peer_s, char* peer_n, size_t cap, int parity,
                         unsigned long long seq, unsigned long long done_target_s, unsigned long long done_target_n) {
  const int dir = blockIdx.z / b.n, r = blockIdx.z % b.n, k = blockIdx.y;
  char* peer_block = dir == 0 ? peer_s : peer_n;
  if (!peer_block) return;
  if (k < b.nlev[r]) {
    const double* a = b.base[r] + (long)k * g.lev;
    // data sent south lands in the neighbour's "from north" slot (1) and vice versa
    double* q = p2p_slot(peer_block, cap, 1 - dir, parity) + b.off[r] + (long)k * nhl * g.ii;
    const int j_first = dir == 0 ? 1 : g.jj - nhl + 1;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < (long)nhl * g.ii; t += (long)gridDim.x * blockDim.x) {
      const int rr = (int)(t / g.ii), i = (int)(t % g.ii) + 1;
      q[t] = a[ix2(g, i, j_first + rr)];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long* done = p2p_word(my_block, 2 + dir);
    const unsigned long long prev = atomicAdd(done, 1ull);
    if (prev + 1 == (dir == 0 ? done_target_s : done_target_n)) {   // cumulative count of all pushes so far in this direction
      __threadfence_system();
      *(volatile unsigned long long*)p2p_word(peer_block, 1 - dir) = seq;
      __threadfence_system();
    }
  }
}

// wait for the neighbours' rows of exchange `seq`, then unpack them into the halo rows (both sides in one launch)
__global__ void p2p_unpack(Geom g, XBatch b, int nhl, int has_s, int has_n, char* my_block, size_t cap, int parity,
                           unsigned long long seq) {
  const int dir = blockIdx.z / b.n, r = blockIdx.z % b.n, k = blockIdx.y;
  if (!(dir == 0 ? has_s : has_n)) return;
  if (threadIdx.x == 0) {
    volatile unsigned long long* flag = p2p_word(my_block, dir);
    while (*flag < seq) {}
    __threadfence_system();
  }
  __syncthreads();
  if (k >= b.nlev[r]) return;
  double* a = b.base[r] + (long)k * g.lev;
  const double* q = p2p_slot(my_block, cap, dir, parity) + b.off[r] + (long)k * nhl * g.ii;
  const int j_first = dir == 0 ? 1 - nhl : g.jj + 1;   // dir 0: rows from the south neighbour
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < (long)nhl * g.ii; t += (long)gridDim.x * blockDim.x) {
    const int rr = (int)(t / g.ii), i = (int)(t % g.ii) + 1;
    a[ix2(g, i, j_first + rr)] = __ldcg(q + t);
  }
}

bool p2p_view(P2PView* v, size_t need_cap) {
  Ctx& c = C(); const Geom& g = c.g;
  P2P& q = g_p2p;
  if (g.nranks < 2 || c.option("comm", "p2p") != "p2p") return false;
  if (q.tried && !q.on) return false;
  if (!q.on || need_cap > q.cap) {
    q.tried = true;
    // room for the largest exchange of the step: 12 fields x kdm levels x nbdy rows
    const size_t want = std::max<size_t>(need_cap, (size_t)12 * g.kdm * g.nb * g.ii);
    if (!p2p_setup(want)) return false;
  }
  v->my_block = q.block; v->peer[0] = q.peer[0]; v->peer[1] = q.peer[1]; v->cap = q.cap;
  v->has_s = g.rank > 0; v->has_n = g.rank + 1 < g.nranks;
  return true;
}
unsigned long long p2p_reserve_seq(int n) {
  const unsigned long long first = g_p2p.seq + 1;
  g_p2p.seq += n;
  return first;
}

// returns false if the P2P path is not available (caller uses NCCL)
bool exchange_ns_p2p(const XBatch& b, long tot, int maxlev, int nhl) {
  Ctx& c = C(); const Geom& g = c.g;
  P2P& q = g_p2p;
  P2PView view;
  if (!p2p_view(&view, (size_t)tot)) return false;
  const bool has_s = g.rank > 0, has_n = g.rank + 1 < g.nranks;
  const unsigned long long seq = p2p_reserve_seq(1);
  const int parity = (int)(seq & 1ull);
  dim3 grid(std::max(1, std::min(cdiv((long)nhl * g.ii, 256), 32)), maxlev, 2 * b.n);   // both directions in one launch
  const unsigned nblk = grid.x * grid.y * b.n;
  if (has_s) q.done_target[0] += nblk;
  if (has_n) q.done_target[1] += nblk;
  LAUNCH(p2p_push, grid, 256, 0, g, b, nhl, q.block, has_s ? q.peer[0] : nullptr, has_n ? q.peer[1] : nullptr, q.cap,
         parity, seq, q.done_target[0], q.done_target[1]);
  LAUNCH(p2p_unpack, grid, 256, 0, g, b, nhl, has_s ? 1 : 0, has_n ? 1 : 0, q.block, q.cap, parity, seq);
  return true;
}

void comm_release_p2p() { g_p2p.tried = false; p2p_close(); }

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
    if (exchange_ns_p2p(b, tot, maxlev, nhl)) continue;
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
namespace blom { void set_last_error(const char* msg); }   // ctx.cu
extern "C" {
int blomgpu_comm_unique_id(char id[128]) {
  try {
    load_nccl();
    ncclUniqueId u;
    NCCL_CHECK(N.GetUniqueId(&u));
    static_assert(sizeof(u) == 128, "ncclUniqueId size");
    std::memcpy(id, &u, 128);
    return 0;
  } catch (const std::exception& e) { set_last_error(e.what()); return 1; }
}
int blomgpu_comm_init(const char id[128], int rank, int nranks) {
  try {
    load_nccl();
    ncclUniqueId u; std::memcpy(&u, id, 128);
    ncclComm_t comm;
    NCCL_CHECK(N.CommInitRank(&comm, nranks, u, rank));
    C().nccl = comm;
    return 0;
  } catch (const std::exception& e) { set_last_error(e.what()); return 1; }
}
}
