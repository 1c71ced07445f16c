// TEST INFRASTRUCTURE ONLY.  Host shim that lets g++ compile the per-thread CUDA kernels of the product
// (`blom_b200/csrc/*.cu`) unchanged, so that their logic can be run one emulated thread after the other on
// the CPU and held against the oracle bit for bit in the `-m "not gpu"` suite (the kernels emulated this way
// use no cross-thread communication: no __syncthreads, no warp intrinsics, shared memory only as a private
// per-thread column).  Nothing in the product includes this file.
#pragma once
#include <cuda_runtime.h>   // vector types, make_double2, dim3; the qualifiers expand to ignored attributes
#include <cmath>

#include <algorithm>
using std::min;   // CUDA's global integer min/max
using std::max;

inline int __ffsll(long long v) { return __builtin_ffsll(v); }

#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif

namespace emu {
inline thread_local uint3 tid{0, 0, 0}, bid{0, 0, 0};
inline thread_local dim3 bdim{1, 1, 1}, gdim{1, 1, 1};
}
#define threadIdx (::emu::tid)
#define blockIdx (::emu::bid)
#define blockDim (::emu::bdim)
#define gridDim (::emu::gdim)

// one emulated launch: every thread of every block in turn
template <class F>
inline void emu_launch(dim3 grid, dim3 block, F&& body) {
  emu::gdim = grid; emu::bdim = block;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx)
        for (unsigned tz = 0; tz < block.z; ++tz)
          for (unsigned ty = 0; ty < block.y; ++ty)
            for (unsigned tx = 0; tx < block.x; ++tx) {
              emu::bid = uint3{bx, by, bz};
              emu::tid = uint3{tx, ty, tz};
              body();
            }
}
