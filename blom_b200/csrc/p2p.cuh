// Peer-to-peer mailbox view shared between comm.cu (stand-alone band-edge exchange) and kernels that
// exchange band edges without leaving the kernel (the persistent barotropic subcycle).
#pragma once
#include "common.cuh"

namespace blom {

constexpr size_t P2P_HDR = 256;

__host__ __device__ inline double* p2p_slot(char* block, size_t cap, int dir, int parity) {
  return reinterpret_cast<double*>(block + P2P_HDR) + ((size_t)dir * 2 + parity) * cap;
}
// header words: [0],[1] flags "from south","from north"; [2],[3] block-done counters of the pushes
__host__ __device__ inline unsigned long long* p2p_word(char* block, int w) {
  return reinterpret_cast<unsigned long long*>(block) + w;
}

struct P2PView {
  char* my_block; char* peer[2]; size_t cap; int has_s, has_n;
};
// mailbox view with at least `need_cap` doubles per slot (collective on first use); false if the
// peer-to-peer path is unavailable or disabled (option comm=nccl)
bool p2p_view(P2PView* v, size_t need_cap);
// reserve n consecutive exchange sequence numbers; returns the first one
unsigned long long p2p_reserve_seq(int n);

}  // namespace blom
